"""Encode kernel alone in its four regimes: int32 / fused f32 input, L2-resident / rotating (HBM) inputs."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from compression_b200 import gen_ops, functional
dev = torch.device("cuda", 0)
scales, ys_host = bench.synth_latents(0, 6)
model = bench.build_model(scales, dev)
ys = [y.to(dev) for y in ys_host]
S = bench.CFG["batch"]; N = ys[0].numel() // S
coff = model.cdf_offset.reshape(-1); qoff = model.quantization_offset
lookup = model._lookup_host()
q0 = 0.0 if qoff is None else qoff.reshape(1, 1, 1, -1)
vs = [(torch.round(y - q0).to(torch.int32).reshape(S, -1, coff.numel()) - coff.reshape(1, 1, -1)).reshape(S, N).contiguous() for y in ys]
def run(kind, rotate):
  ts = []
  for i in range(8):
    h = gen_ops.create_range_encoder([S], lookup)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    k = i % 6 if rotate else 0
    a.record()
    if kind == "f32": functional.encode_channel_f32(h, ys[k], qoff, coff)
    else: gen_ops.entropy_encode_channel(h, vs[k])
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b)); h.close()
  return float(np.median(ts[2:]))
for kind in ("i32", "f32"):
  for rotate in (False, True):
    t = run(kind, rotate)
    print(f"{kind} rotate={rotate}: {t:.3f} ms -> {S*N/t/1e3:.0f} Msym/s")
