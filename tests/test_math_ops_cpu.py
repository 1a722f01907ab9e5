"""CPU: bound / perturbation helpers and the GDN parameter reparameterisation (pure torch), with the cases of
py/ops/math_ops_test.py:25-106, py/ops/round_ops.py:28-43 and py/layers/parameters_test.py:24-90."""
import numpy as np
import pytest
import scipy.stats
import torch

from compression_b200 import math_ops
import compression_b200 as tfc
from compression_b200.gdn import GDNParameter


def _grads(fn, inputs, sign):
  x = inputs.clone().requires_grad_(True)
  y = fn(x)
  (g,) = torch.autograd.grad(y, x, sign * torch.ones_like(y))
  return y.detach(), g


@pytest.mark.parametrize("gradient", ["disconnected", "identity", "identity_if_towards"])
def test_upper_bound_outputs_and_gradients(gradient):
  inputs = torch.tensor([-1., 1.])
  out, pg = _grads(lambda x: math_ops.upper_bound(x, 0, gradient=gradient), inputs, 1.0)
  _, ng = _grads(lambda x: math_ops.upper_bound(x, 0, gradient=gradient), inputs, -1.0)
  assert out.tolist() == [-1, 0]
  want = {"disconnected": ([1, 0], [-1, 0]), "identity": ([1, 1], [-1, -1]), "identity_if_towards": ([1, 1], [-1, 0])}
  assert (pg.tolist(), ng.tolist()) == want[gradient]


@pytest.mark.parametrize("gradient", ["disconnected", "identity", "identity_if_towards"])
def test_lower_bound_outputs_and_gradients(gradient):
  inputs = torch.tensor([-1., 1.])
  out, pg = _grads(lambda x: math_ops.lower_bound(x, 0, gradient=gradient), inputs, 1.0)
  _, ng = _grads(lambda x: math_ops.lower_bound(x, 0, gradient=gradient), inputs, -1.0)
  assert out.tolist() == [0, 1]
  want = {"disconnected": ([0, 1], [0, -1]), "identity": ([1, 1], [-1, -1]), "identity_if_towards": ([0, 1], [-1, -1])}
  assert (pg.tolist(), ng.tolist()) == want[gradient]


def test_bounds_reject_unknown_gradient():
  with pytest.raises(ValueError):
    math_ops.upper_bound(torch.zeros(1, 2), 0, gradient="invalid")
  with pytest.raises(ValueError):
    math_ops.lower_bound(torch.zeros(1, 2), 0, gradient="invalid")


def test_round_st_rounds_half_to_even_with_offset_and_passes_gradients():
  x = torch.tensor([-2.5, -0.5, 0.5, 1.5, 2.4, 2.6], requires_grad=True)
  y = math_ops.round_st(x)
  assert y.tolist() == [-2., -0., 0., 2., 2., 3.]
  y.sum().backward()
  assert x.grad.tolist() == [1.] * 6
  off = torch.tensor(0.25)
  assert math_ops.round_st(torch.tensor([0.7, 0.8, -0.3]), off).tolist() == pytest.approx([0.25, 1.25, -0.75])


def test_perturb_and_apply_noise_is_uniform():
  torch.manual_seed(0)
  x = torch.randn(10000)
  y, x_plus_u = math_ops.perturb_and_apply(lambda t: t, x, expected_grads=True)
  u0, u1 = (x_plus_u - x).numpy(), (y - x).detach().numpy()
  np.testing.assert_allclose(u0, u1, atol=1e-6)
  assert np.abs(u0).max() <= 0.5
  assert scipy.stats.kstest(u0, "uniform", (-0.5, 1.0))[1] > 1e-6


def test_perturb_and_apply_expected_gradient_of_a_parabola():
  f = lambda t, a: a * t * t
  x = torch.linspace(-2.0, 2.0, 200, requires_grad=True)
  y = math_ops.perturb_and_apply(f, x, 7.0, expected_grads=True)[0]
  (dx,) = torch.autograd.grad(y.sum(), x)
  want = f(x.detach() + .5, 7.0) - f(x.detach() - .5, 7.0)
  np.testing.assert_allclose(dx.numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
  with pytest.raises(ValueError):
    math_ops.perturb_and_apply(f, x, 7.0, u=torch.zeros_like(x), x_plus_u=x)


def test_gdn_parameter_reproduces_initial_value_and_minimum():
  torch.manual_seed(1)
  init = torch.rand(4, 3)
  p = GDNParameter(init)
  np.testing.assert_allclose(p().detach().numpy(), init.numpy(), rtol=1e-5, atol=1e-6)
  pm = GDNParameter(torch.tensor([0.0, 0.05, 0.3]), minimum=0.1)       # parameters_test.py:82-90
  np.testing.assert_allclose(pm().detach().numpy(), [0.1, 0.1, 0.3], rtol=1e-5, atol=1e-6)
  cfg = pm.get_config()
  assert cfg["minimum"] == 0.1 and cfg["shape"] == (3,)


def test_gdn_parameter_gradients_propagate():
  p = GDNParameter(torch.full((5,), 0.5))
  p().sum().backward()
  assert p.variable.grad is not None and bool((p.variable.grad != 0).all())
  with pytest.raises(ValueError):
    GDNParameter(None)


@pytest.mark.parametrize("kw", [dict(epsilon_parameter=None), dict(alpha_parameter=None),
                                dict(alpha_parameter=None, epsilon_parameter=None, rectify=True),
                                dict(alpha_parameter=2, epsilon_parameter=None),
                                dict(alpha_parameter=None, epsilon_parameter=.5, inverse=True)])
def test_gdn_trainable_exponent_graph_keeps_the_fixed_special_cases(kw):
  """layers/gdn.py:377-415: |x| for a FIXED alpha == 1 (no rectify), square for a fixed 2, sqrt for a fixed
  epsilon == .5 -- also when the other exponent is a trainable GDNParameter; a trainable alpha is a plain
  `inputs ** alpha` on the signed input, as in the reference."""
  torch.manual_seed(0)
  C = 6
  layer = tfc.GDN(**kw)
  x = torch.randn(9, C)
  layer.build(x.shape)
  y = layer._torch_graph(x, x.device)   # the branch GDN.forward takes for callable exponents (device-free)
  xd, gamma, beta = x.double(), layer.gamma.detach().double(), layer.beta.detach().double()
  u = torch.relu(xd) if layer.rectify else xd
  a, e = layer.alpha_parameter, layer.epsilon_parameter
  if callable(a):
    pool = u**float(layer.alpha.detach())
  else:
    pool = {1: u if layer.rectify else u.abs(), 2: u * u}[a]
  n = pool @ gamma + beta
  if callable(e):
    n = n**float(layer.epsilon.detach())
  elif e == .5:
    n = n.sqrt()
  want = u * n if layer.inverse else u / n
  assert torch.isfinite(y).all()
  np.testing.assert_allclose(y.detach().double().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
  if not callable(a):   # the fixed-alpha cases are exactly the oracle's graph
    from oracle import gdn_oracle
    eps = float(layer.epsilon.detach()) if callable(e) else e
    ref = gdn_oracle.gdn_reference(x, gamma, beta, layer.inverse, layer.rectify, a, eps)
    np.testing.assert_allclose(y.detach().double().numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


def test_perturb_and_apply_expected_gradient_of_soft_round():
  """math_ops_test.py:88-96: soft_round(x + 1/2) - soft_round(x - 1/2) == 1 for every x, so the expected gradient
  of the soft-rounded, noise-perturbed signal is the identity's."""
  x = torch.linspace(-2.0, 2.0, 200, requires_grad=True)
  y = math_ops.perturb_and_apply(math_ops.soft_round, x, 7.0, expected_grads=True)[0]
  y.sum().backward()
  assert torch.allclose(x.grad, torch.ones_like(x.grad), atol=1e-5)
