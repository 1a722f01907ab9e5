"""Functional CUDA entry points that have no op in the reference because the reference composes them from
TF primitives: GDN/IGDN forward + backward (``python/layers/gdn.py:371-421`` + TF autodiff) and the fused
quantise+encode / decode+dequantise paths of the entropy models."""
import ctypes as C

import torch

from compression_b200 import _lib
from compression_b200._lib import check

GDN_INVERSE = 1
GDN_RECTIFY = 2


def _stream() -> int:
  return torch.cuda.current_stream().cuda_stream


def _p(t):
  return None if t is None else C.c_void_p(t.data_ptr())


def _flags(inverse, rectify):
  return (GDN_INVERSE if inverse else 0) | (GDN_RECTIFY if rectify else 0)


def gdn_forward(x, gamma, beta, inverse=False, rectify=False, alpha=1.0, epsilon=1.0):
  """x: float32 CUDA [..., C] (channels-last, contiguous) -> y of the same shape."""
  assert x.is_cuda and x.dtype == torch.float32
  x = x.contiguous()
  C_ = x.shape[-1]
  gamma = gamma.to(device=x.device, dtype=torch.float32).contiguous()
  beta = beta.to(device=x.device, dtype=torch.float32).contiguous()
  assert gamma.shape == (C_, C_) and beta.shape == (C_,)
  y = torch.empty_like(x)
  n_pix = x.numel() // C_
  check(_lib.lib().tfcb_gdn_forward(_p(x), _p(gamma), _p(beta), _p(y), n_pix, C_, _flags(inverse, rectify),
                                    float(alpha), float(epsilon), _stream()))
  return y


def gdn_backward(x, gamma, beta, dy, inverse=False, rectify=False, alpha=1.0, epsilon=1.0):
  """Returns (dx, dgamma, dbeta) for upstream gradient dy."""
  assert x.is_cuda and x.dtype == torch.float32
  x = x.contiguous()
  dy = dy.to(dtype=torch.float32).contiguous()
  C_ = x.shape[-1]
  gamma = gamma.to(device=x.device, dtype=torch.float32).contiguous()
  beta = beta.to(device=x.device, dtype=torch.float32).contiguous()
  n_pix = x.numel() // C_
  dx = torch.empty_like(x)
  dgamma = torch.empty_like(gamma)
  dbeta = torch.empty_like(beta)
  ws_bytes = int(_lib.lib().tfcb_gdn_backward_workspace_bytes(n_pix, C_))
  ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
  check(_lib.lib().tfcb_gdn_backward(_p(x), _p(gamma), _p(beta), _p(dy), _p(dx), _p(dgamma), _p(dbeta), _p(ws),
                                     n_pix, C_, _flags(inverse, rectify), float(alpha), float(epsilon),
                                     _stream()))
  return dx, dgamma, dbeta


class _GDNFunction(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, gamma, beta, inverse, rectify, alpha, epsilon):
    ctx.save_for_backward(x, gamma, beta)
    ctx.cfg = (inverse, rectify, alpha, epsilon)
    return gdn_forward(x, gamma, beta, inverse, rectify, alpha, epsilon)

  @staticmethod
  def backward(ctx, dy):
    x, gamma, beta = ctx.saved_tensors
    inverse, rectify, alpha, epsilon = ctx.cfg
    dx, dgamma, dbeta = gdn_backward(x, gamma, beta, dy, inverse, rectify, alpha, epsilon)
    return dx, dgamma, dbeta, None, None, None, None


def gdn(x, gamma, beta, inverse=False, rectify=False, alpha=1.0, epsilon=1.0):
  """Differentiable GDN/IGDN on channels-last float32 CUDA tensors."""
  return _GDNFunction.apply(x, gamma, beta, bool(inverse), bool(rectify), float(alpha), float(epsilon))
