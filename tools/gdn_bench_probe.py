"""Scratch: why does bench.py's GDN / decode timing differ from the standalone tools?  Times the same call
(a) one call per event pair, (b) back to back, (c) after the oracle pool exists, (d) after NUMA pinning."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from compression_b200 import functional as F
import bench

def one_by_one(fn, n=5):
  ts = []
  for _ in range(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return sorted(ts)[len(ts) // 2]

def host_time(fn, n=5):
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for _ in range(n): fn()
  t1 = time.perf_counter(); torch.cuda.synchronize()
  return (t1 - t0) / n * 1e3

def run(tag):
  C, npix = 128, 256 * 64 * 64
  gamma = (0.1 * torch.eye(C) + (0.02 * torch.randn(C, C)).abs()).cuda(); beta = (1 + 0.5 * torch.rand(C)).cuda()
  x = torch.randn(npix, C, device="cuda"); dy = torch.randn_like(x)
  fwd = lambda: F.gdn_forward(x, gamma, beta)
  bwd = lambda: F.gdn_backward(x, gamma, beta, dy)
  fwd(); bwd(); torch.cuda.synchronize()
  print(tag, "fwd one-by-one %.3f ms, back-to-back %.3f ms, host %.3f ms/call" % (one_by_one(fwd), bench._time_ms(fwd, 5, warm=2)[0], host_time(fwd)))
  print(tag, "bwd one-by-one %.3f ms, back-to-back %.3f ms, host %.3f ms/call" % (one_by_one(bwd), bench._time_ms(bwd, 3)[0], host_time(bwd)), flush=True)

run("fresh       ")
bench.warm_oracle_pool()
run("oracle pool ")
print(bench.pin_to_gpu_numa_node(0))
run("numa pinned ")
os.sched_setaffinity(0, range(os.cpu_count()))
run("unpinned    ")
