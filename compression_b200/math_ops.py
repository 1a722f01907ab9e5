"""Bound / perturbation / rounding helpers, mirroring tensorflow_compression/python/ops/math_ops.py:27-216 and
round_ops.py:28-130 on PyTorch autograd."""
import torch

__all__ = ["upper_bound", "lower_bound", "perturb_and_apply", "round_st", "soft_round", "soft_round_inverse",
           "soft_round_conditional_mean"]

_GRADIENTS = ("identity_if_towards", "identity", "disconnected")


class _Bound(torch.autograd.Function):

  @staticmethod
  def forward(ctx, inputs, bound, is_upper, gradient):
    ctx.save_for_backward(inputs, bound)
    ctx.is_upper = is_upper
    ctx.gradient = gradient
    return torch.minimum(inputs, bound) if is_upper else torch.maximum(inputs, bound)

  @staticmethod
  def backward(ctx, grad):
    inputs, bound = ctx.saved_tensors
    if ctx.gradient == "identity":
      return grad, None, None, None
    inside = (inputs <= bound) if ctx.is_upper else (inputs >= bound)
    if ctx.gradient == "disconnected":
      return inside.to(grad.dtype) * grad, None, None, None
    towards = (grad > 0) if ctx.is_upper else (grad < 0)  # math_ops.py:71-75,132-135
    return (inside | towards).to(grad.dtype) * grad, None, None, None


def _bound(inputs, bound, gradient, is_upper):
  if gradient not in _GRADIENTS:
    raise ValueError(f"Invalid value for `gradient`: '{gradient}'.")
  inputs = torch.as_tensor(inputs)
  bound = torch.as_tensor(bound, dtype=inputs.dtype, device=inputs.device)
  return _Bound.apply(inputs, bound, is_upper, gradient)


def upper_bound(inputs, bound, gradient="identity_if_towards"):
  """`minimum(inputs, bound)` with the reference's gradient choices (math_ops.py:27-90)."""
  return _bound(inputs, bound, gradient, True)


def lower_bound(inputs, bound, gradient="identity_if_towards"):
  """`maximum(inputs, bound)` with the reference's gradient choices (math_ops.py:93-154)."""
  return _bound(inputs, bound, gradient, False)


class _RoundST(torch.autograd.Function):

  @staticmethod
  def forward(ctx, inputs, offset):
    if offset is None:
      return torch.round(inputs)  # round-half-even, like tf.round
    return torch.round(inputs - offset) + offset

  @staticmethod
  def backward(ctx, grad):
    return grad, None


def round_st(inputs, offset=None):
  """Straight-through round with optional quantization offset (round_ops.py:28-43)."""
  return _RoundST.apply(inputs, offset)


def perturb_and_apply(f, x, *args, u=None, x_plus_u=None, expected_grads=True):
  """math_ops.py:157-216: y = f(x + u, *args) with u ~ U(-.5, .5); with `expected_grads` the gradient
  w.r.t. x is the analytic expectation f(x + .5) - f(x - .5)."""
  if x_plus_u is None:
    if u is None:
      u = torch.rand_like(x) - 0.5
    x_plus_u = x + u
  elif u is not None:
    raise ValueError("Cannot provide both `u` and `x_plus_u`.")
  if not expected_grads:
    return f(x_plus_u, *args), x_plus_u

  xpu = x_plus_u.detach()
  y = f(xpu, *args)  # gradients to args / parameters flow normally
  with torch.no_grad():
    dydx = f(x.detach() + 0.5, *args) - f(x.detach() - 0.5, *args)
  # y + (x - x.detach()) * dydx has value y and d/dx = dydx
  y = y + (x - x.detach()) * dydx
  return y, x_plus_u


def _alpha_like(alpha, x):
  return torch.as_tensor(alpha, dtype=x.dtype, device=x.device)


def soft_round(x, alpha, eps=1e-3):
  """Differentiable approximation of round (round_ops.py:44-74; Agustsson & Theis 2020, sec. 4.1):
  m + tanh(alpha r) / (2 tanh(alpha / 2)) with m = floor(x) + 1/2, r = x - m; the identity for alpha < eps.  alpha
  is bounded below by eps inside the formula so that the branch not taken has finite gradients."""
  x = torch.as_tensor(x)
  alpha = _alpha_like(alpha, x)
  bounded = torch.clamp_min(alpha, eps)
  m = torch.floor(x) + .5
  y = m + torch.tanh(bounded * (x - m)) / (torch.tanh(bounded / 2.) * 2.)
  return torch.where(alpha < eps, x, y)


def soft_round_inverse(y, alpha, eps=1e-3):
  """Inverse of `soft_round` (round_ops.py:77-108): r = atanh(2 tanh(alpha / 2) (y - m)) / alpha clipped to
  [-1/2, 1/2] (atanh overflows first for large alpha), result m + r; the identity for alpha < eps."""
  y = torch.as_tensor(y)
  alpha = _alpha_like(alpha, y)
  bounded = torch.clamp_min(alpha, eps)
  m = torch.floor(y) + .5
  r = torch.atanh((y - m) * (torch.tanh(bounded / 2.) * 2.)) / bounded
  r = torch.clamp(r, -.5, .5)
  return torch.where(alpha < eps, y, m + r)


def soft_round_conditional_mean(y, alpha):
  """E[Y | soft_round(Y) + U = y], U ~ U(-1/2, 1/2), Y locally uniform (round_ops.py:111-130)."""
  return soft_round_inverse(torch.as_tensor(y) - .5, alpha) + .5
