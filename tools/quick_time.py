"""Scratch timing of the hot kernels at cfg2 shapes (not the official bench)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import util
from compression_b200 import gen_ops, functional

def timeit(fn, n=5, warm=2):
  for _ in range(warm): fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(n):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return min(ts), sorted(ts)[len(ts)//2]

rng = np.random.default_rng(0)
C, S, N = 128, 256, 32768
cdfs = [util.laplace_cdf(41, 12, 0.4 + 0.05 * c) for c in range(C)]
lookup = util.make_lookup_1d(cdfs, [12] * C, [True] * C)
value = np.empty((S, N), np.int32)
for c in range(C):
  value[:, c::C] = util.sample_symbols(rng, cdfs[c], S * (N // C)).reshape(S, -1)
v = torch.from_numpy(value).cuda()
box = {}
def enc():
  h = gen_ops.create_range_encoder([S], lookup)
  gen_ops.entropy_encode_channel(h, v)
  box["s"] = gen_ops.entropy_encode_finalize(h)
t, tm = timeit(enc)
nbytes = box["s"].nbytes()
print(f"encode cfg2: {t:.3f} ms best / {tm:.3f} med -> {S*N/t/1e3:.1f} Msym/s, {8*nbytes/(S*N):.3f} bit/sym")
# kernel-only timing (handle creation / finalize outside the timed region)
def enc_only():
  hs = [gen_ops.create_range_encoder([S], lookup) for _ in range(6)]
  torch.cuda.synchronize()
  ts = []
  for h in hs:
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); gen_ops.entropy_encode_channel(h, v); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return min(ts)
t = enc_only()
print(f"encode kernel only: {t:.3f} ms -> {S*N/t/1e3:.1f} Msym/s")
def dec_only():
  hs = [gen_ops.create_range_decoder(box["s"], lookup) for _ in range(6)]
  torch.cuda.synchronize()
  ts = []
  for h in hs:
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); gen_ops.entropy_decode_channel(h, [N]); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return min(ts)
t = dec_only()
print(f"decode kernel only: {t:.3f} ms -> {S*N/t/1e3:.1f} Msym/s")
def dec():
  h = gen_ops.create_range_decoder(box["s"], lookup)
  h, out = gen_ops.entropy_decode_channel(h, [N])
  box["ok"] = gen_ops.entropy_decode_finalize(h); box["o"] = out
t, tm = timeit(dec)
assert torch.equal(box["o"], v) and bool(box["ok"].all())
print(f"decode cfg2: {t:.3f} ms best / {tm:.3f} med -> {S*N/t/1e3:.1f} Msym/s")
for Cg, npix in ((128, 256*64*64), (192, 128*128*128)):
  x = torch.randn(npix, Cg, device="cuda")
  gamma = (0.1*torch.eye(Cg) + (0.02*torch.randn(Cg, Cg)).abs()).cuda(); beta = (1+0.5*torch.rand(Cg)).cuda()
  t, tm = timeit(lambda: functional.gdn_forward(x, gamma, beta), n=3, warm=1)
  print(f"gdn fwd C={Cg} npix={npix}: {t:.3f} ms -> {8*npix*Cg/t/1e6:.1f} GB/s")
  dy = torch.randn_like(x)
  t, tm = timeit(lambda: functional.gdn_backward(x, gamma, beta, dy), n=3, warm=1)
  print(f"gdn bwd C={Cg} npix={npix}: {t:.3f} ms -> {12*npix*Cg/t/1e6:.1f} GB/s")
