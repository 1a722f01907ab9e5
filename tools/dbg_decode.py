import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import util, oracle
from compression_b200 import gen_ops as ops
rng = np.random.default_rng(11)
for case in range(4):
  if case == 0: spec = ((41, 3.0), (301, 40.0), (1501, 250.0), (9, 0.7)); escp = 0.01
  if case == 1: spec = ((301, 40.0),); escp = 0.0
  if case == 2: spec = ((41, 3.0),); escp = 0.01
  if case == 3: spec = ((1501, 250.0),); escp = 0.0
  cdfs = [util.laplace_cdf(n, 12, s) for n, s in spec]
  R = len(cdfs)
  lookup = util.make_lookup_1d(cdfs, [12] * R, [True] * R)
  S, N = 4, 1024
  index = rng.integers(0, R, size=(S, N)).astype(np.int32)
  value = np.empty((S, N), np.int32)
  for r, c in enumerate(cdfs):
    m = index == r
    value[m] = util.sample_symbols(rng, c, int(m.sum()))
  esc = rng.random((S, N)) < escp
  value[esc] = rng.integers(-3000, 3000, size=int(esc.sum()))
  h = ops.create_range_encoder([S], torch.from_numpy(lookup))
  h = ops.entropy_encode_index(h, torch.from_numpy(index).cuda(), torch.from_numpy(value).cuda())
  got = ops.entropy_encode_finalize(h)
  hd = ops.create_range_decoder(got, torch.from_numpy(lookup))
  hd, dec = ops.entropy_decode_index(hd, torch.from_numpy(index).cuda(), [N])
  dec = dec.cpu().numpy()
  bad = np.argwhere(dec != value)
  print(case, 'mismatches', len(bad), 'first', bad[:3].tolist())
  for s_, j in bad[:3]:
    print('   idx', index[s_, j], 'want', value[s_, j], 'got', dec[s_, j], 'prev', value[s_, max(0,j-2):j].tolist(), 'ncdf', len(cdfs[index[s_, j]]))
