"""Host-side breakdown of one cfg2 compress step (create / encode / finalize), wall clock with syncs."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from compression_b200 import gen_ops, functional
dev = torch.device("cuda", 0)
scales, ys_host = bench.synth_latents(0, 2)
model = bench.build_model(scales, dev)
ys = [y.to(dev) for y in ys_host]
S = bench.CFG["batch"]
for _ in range(3): model.compress(ys[0])
torch.cuda.synchronize()
coff = model.cdf_offset.reshape(-1); qoff = model.quantization_offset
T = {"lookup_host": [], "create": [], "encode_launch": [], "encode_wait": [], "finalize": [], "total_compress": []}
for i in range(20):
  torch.cuda.synchronize(); t0 = time.perf_counter()
  lookup = model._lookup_host(); t1 = time.perf_counter()
  h = gen_ops.create_range_encoder([S], lookup); t2 = time.perf_counter()
  functional.encode_channel_f32(h, ys[i % 2], qoff, coff); t3 = time.perf_counter()
  torch.cuda.synchronize(); t4 = time.perf_counter()
  s = gen_ops.entropy_encode_finalize(h); t5 = time.perf_counter()
  T["lookup_host"].append(t1 - t0); T["create"].append(t2 - t1); T["encode_launch"].append(t3 - t2)
  T["encode_wait"].append(t4 - t3); T["finalize"].append(t5 - t4)
  torch.cuda.synchronize(); t6 = time.perf_counter()
  model.compress(ys[i % 2]); torch.cuda.synchronize(); T["total_compress"].append(time.perf_counter() - t6)
for k, v in T.items(): print(f"{k:16s} {1e3 * float(np.median(v[3:])):.3f} ms")
