"""Minimal driver for ncu: a few cfg2 encode + decode launches (no timing)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from compression_b200 import gen_ops
rng = np.random.default_rng(0)
C, S, N = 128, 256, 32768
cdfs = [util.laplace_cdf(41, 12, 0.4 + 0.05 * c) for c in range(C)]
lookup = util.make_lookup_1d(cdfs, [12] * C, [True] * C)
value = np.empty((S, N), np.int32)
for c in range(C):
  value[:, c::C] = util.sample_symbols(rng, cdfs[c], S * (N // C)).reshape(S, -1)
v = torch.from_numpy(value).cuda()
for _ in range(3):
  h = gen_ops.create_range_encoder([S], lookup)
  gen_ops.entropy_encode_channel(h, v)
  s = gen_ops.entropy_encode_finalize(h)
  hd = gen_ops.create_range_decoder(s, lookup)
  hd, out = gen_ops.entropy_decode_channel(hd, [N])
  ok = gen_ops.entropy_decode_finalize(hd)
torch.cuda.synchronize()
assert torch.equal(out, v)
