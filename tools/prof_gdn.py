"""Minimal driver for ncu: GDN forward at the cfg2 shape (C=128) and a C=192 shape, backward at both."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compression_b200 import functional
for C, npix in ((128, 256 * 64 * 64), (192, 128 * 64 * 64)):
  x = torch.randn(npix, C, device="cuda")
  gamma = (0.1 * torch.eye(C) + (0.02 * torch.randn(C, C)).abs()).cuda(); beta = (1 + 0.5 * torch.rand(C)).cuda()
  for _ in range(3):
    y = functional.gdn_forward(x, gamma, beta)
  if C in (128, 192):
    dy = torch.randn_like(x)
    for _ in range(3):
      functional.gdn_backward(x, gamma, beta, dy)
torch.cuda.synchronize()
