"""Scratch: attribution of the C = 128 backward kernel's time by switching pieces off (TFCB_GDN_DBG bits)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compression_b200 import functional as F
C, n_pix = 128, 256 * 64 * 64
torch.manual_seed(0)
gamma = (0.1 * torch.eye(C) + (0.02 * torch.randn(C, C)).abs()).cuda(); beta = (1 + 0.5 * torch.rand(C)).cuda()
x = torch.randn(n_pix, C, device="cuda"); dy = torch.randn_like(x)
def med_ms(fn, reps=9):
  out = fn(); out = fn()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
  torch.cuda.synchronize(); ev[0].record()
  for i in range(reps):
    out = fn(); ev[i + 1].record()
  torch.cuda.synchronize()
  return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))[reps // 2]
for dbg in (0, 8, 0, 8):
  os.environ["TFCB_GDN_DBG"] = str(dbg)
  print(f"dbg={dbg}: {med_ms(lambda: F.gdn_backward(x, gamma, beta, dy)):.3f} ms", flush=True)
