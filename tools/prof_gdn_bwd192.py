"""Minimal driver for ncu: the C = 192 GDN backward (two kernels) at 0.5 M pixels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compression_b200 import functional
C, npix = 192, 128 * 64 * 64
x = torch.randn(npix, C, device="cuda"); dy = torch.randn_like(x)
gamma = (0.1 * torch.eye(C) + (0.02 * torch.randn(C, C)).abs()).cuda(); beta = (1 + 0.5 * torch.rand(C)).cuda()
for _ in range(3):
  functional.gdn_backward(x, gamma, beta, dy)
torch.cuda.synchronize()
