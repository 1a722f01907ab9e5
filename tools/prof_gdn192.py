"""Minimal driver for ncu: the C = 192 GDN forward at 2 M pixels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compression_b200 import functional
C, npix = 192, 128 * 128 * 128
x = torch.randn(npix, C, device="cuda")
gamma = (0.1 * torch.eye(C) + (0.02 * torch.randn(C, C)).abs()).cuda(); beta = (1 + 0.5 * torch.rand(C)).cuda()
for _ in range(3):
  y = functional.gdn_forward(x, gamma, beta)
torch.cuda.synchronize()
