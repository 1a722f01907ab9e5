"""ORACLE (test infrastructure): CPU restatement of tfc.GDN.call, tensorflow_compression/python/layers/
gdn.py:371-421, in PyTorch.  fp64 is the ground truth, fp32 the reference-precision path; gradients come
from torch autograd of the same graph (the reference relies on TF autodiff, it has no hand-written
gradient).  Also holds the closed forms asserted by the reference's tests (layers/gdn_test.py:42-88)."""
import torch


def gdn_reference(x, gamma, beta, inverse=False, rectify=False, alpha=1.0, epsilon=1.0, dtype=torch.float64):
  x = x.detach().to("cpu", dtype)
  gamma = gamma.detach().to("cpu", dtype)
  beta = beta.detach().to("cpu", dtype)
  return _graph(x, gamma, beta, inverse, rectify, alpha, epsilon)


def _graph(x, gamma, beta, inverse, rectify, alpha, epsilon):
  u = torch.relu(x) if rectify else x
  if alpha == 1 and rectify:
    pool = u
  elif alpha == 1:
    pool = u.abs()
  elif alpha == 2:
    pool = u * u
  else:
    pool = u**alpha
  n = pool @ gamma + beta  # 1x1 convolution over the channel axis == matmul on [..., C]
  if epsilon == 1:
    pass
  elif epsilon == 0.5:
    n = n.sqrt()
  else:
    n = n**epsilon
  return u * n if inverse else u / n


def gdn_reference_grads(x, gamma, beta, dy, inverse=False, rectify=False, alpha=1.0, epsilon=1.0,
                        dtype=torch.float64):
  x = x.detach().to("cpu", dtype).requires_grad_(True)
  gamma = gamma.detach().to("cpu", dtype).requires_grad_(True)
  beta = beta.detach().to("cpu", dtype).requires_grad_(True)
  y = _graph(x, gamma, beta, inverse, rectify, alpha, epsilon)
  y.backward(dy.detach().to("cpu", dtype))
  return x.grad, gamma.grad, beta.grad
