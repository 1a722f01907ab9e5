"""CPU: soft-round ops, layers and the round adapters, after the reference's round_ops_test.py:25-87,
soft_round_test.py:22-43 and round_adapters_test.py:27-262 (closed forms and limits; no kernels involved)."""
import math

import numpy as np
import pytest
import torch

from compression_b200 import distributions as D
from compression_b200 import math_ops
from compression_b200 import soft_round_layers


def test_small_alpha_is_the_identity():
  x = torch.linspace(-2., 2., 50)
  assert torch.allclose(math_ops.soft_round(x, alpha=1e-13), x)
  assert torch.equal(math_ops.soft_round_inverse(x, alpha=1e-13), x)


def test_large_alpha_limits():
  for offset in range(-5, 5):   # away from the half-integers / integers, where the limit is discontinuous
    x = torch.linspace(offset - 0.499, offset + 0.499, 100)
    assert torch.allclose(math_ops.soft_round(x, alpha=2000.0), torch.round(x), atol=0.02)
    x = torch.linspace(offset + 0.001, offset + 0.999, 100)
    assert torch.allclose(math_ops.soft_round_inverse(x, alpha=5000.0), torch.ceil(x) - 0.5, atol=0.001)
    assert torch.allclose(math_ops.soft_round_conditional_mean(x, alpha=5000.0), torch.round(x), atol=0.001)


def test_inverse_is_the_inverse():
  x = torch.tensor([-1.25, -0.75, 0.75, 1.25])
  assert torch.allclose(math_ops.soft_round_inverse(math_ops.soft_round(x, alpha=2.0), alpha=2.0), x)
  # fixed points: integers and half-integers, for every alpha
  p = torch.arange(-3., 3.5, .5)
  for alpha in (.1, 1., 7.):
    assert torch.allclose(math_ops.soft_round(p, alpha), p, atol=1e-6)


@pytest.mark.parametrize("alpha", [0., 1e-6, 1e-2, 5., 1e6])
def test_values_and_gradients_are_finite(alpha):
  x = torch.linspace(0., 1., 11, requires_grad=True)   # exact integers and half-integers included
  y = math_ops.soft_round(x, alpha=alpha)
  y.sum().backward()
  assert bool(torch.isfinite(y).all()) and bool(torch.isfinite(x.grad).all())
  x = torch.linspace(-.5, .5, 11, requires_grad=True)
  y = math_ops.soft_round_inverse(x, alpha=alpha)
  y.sum().backward()
  assert bool(torch.isfinite(y).all())
  finite = torch.isfinite(x.grad)
  if alpha > 15:   # the function is extremely steep at 0 for large alphas (round_ops_test.py:82-86)
    finite[5] = True
  assert bool(finite.all())


def test_layers_apply_the_ops():
  x = torch.linspace(-5.0, 5.0, 50)
  assert torch.allclose(soft_round_layers.SoftRound(alpha=5.0)(x), math_ops.soft_round(x, 5.0))
  assert torch.allclose(soft_round_layers.SoftRound(alpha=5.0, inverse=True)(x), math_ops.soft_round_inverse(x, 5.0))
  assert torch.allclose(soft_round_layers.SoftRoundConditionalMean(alpha=5.0)(x),
                        math_ops.soft_round_conditional_mean(x, 5.0))
  assert soft_round_layers.SoftRound().compute_output_shape((2, 3)) == (2, 3)


ADAPTED = [
    ("softround_deepfactorized", lambda d: D.SoftRoundAdapter(d, alpha=5.0), lambda: D.DeepFactorized()),
    ("softround_logistic", lambda d: D.SoftRoundAdapter(d, alpha=5.0), lambda: D.Logistic(10.3, 1.5)),
    ("softround_normal", lambda d: D.SoftRoundAdapter(d, alpha=4.0), lambda: D.Normal(10.4, 1.5)),
    ("noisysoftround_deepfactorized", lambda d: D.NoisySoftRoundAdapter(d, alpha=5.0), lambda: D.DeepFactorized()),
    ("noisysoftround_logistic", lambda d: D.NoisySoftRoundAdapter(d, alpha=5.0), lambda: D.Logistic(10., 1.5)),
    ("noisysoftround_normal", lambda d: D.NoisySoftRoundAdapter(d, alpha=5.0), lambda: D.Normal(10., 1.5)),
    ("round_deepfactorized", D.RoundAdapter, lambda: D.DeepFactorized(init_scale=1.0)),
    ("round_logistic", D.RoundAdapter, lambda: D.Logistic(1.5, 1.5)),
    ("round_normal", D.RoundAdapter, lambda: D.Normal(1.5, 1.5)),
    ("noisyround_deepfactorized", D.NoisyRoundAdapter, lambda: D.DeepFactorized(init_scale=1.0)),
    ("noisyround_logistic", D.NoisyRoundAdapter, lambda: D.Logistic(1.5, 1.5)),
    ("noisyround_normal", D.NoisyRoundAdapter, lambda: D.Normal(1.5, 1.5)),
]


@pytest.mark.parametrize("name,adapter,base", ADAPTED, ids=[a[0] for a in ADAPTED])
def test_tails(name, adapter, base):
  """round_adapters_test.py:85-103: at most tail_mass / 2... the reference asserts <= tail_mass on either side."""
  dist = adapter(base())
  lo = dist._lower_tail(2**-8)
  try:
    left = dist.cdf(lo)
  except (NotImplementedError, AttributeError):
    left = dist.base.cdf(lo)       # the noisy adapters have no cdf: the base's mass stands in
  hi = dist._upper_tail(2**-8)
  try:
    right = dist.survival_function(hi)
  except (NotImplementedError, AttributeError):
    right = dist.base.survival_function(hi)
  assert float(left) <= 2**-8 and float(right) <= 2**-8 and float(hi) > float(lo)


@pytest.mark.parametrize("base", [lambda: D.Logistic(10., 1.5), lambda: D.Normal(10., 1.5)])
def test_mode_and_quantile_pass_through_the_transform(base):
  dist = D.SoftRoundAdapter(base(), alpha=5.0)
  assert math.isclose(float(dist.cdf(dist.mode())), 0.5, abs_tol=1e-5)
  assert math.isclose(float(dist.cdf(dist.quantile(0.75))), 0.75, abs_tol=1e-5)


def test_non_invertible_adapters_refuse_what_they_cannot_give():
  dist = D.RoundAdapter(D.Logistic(1.5, 1.5))
  with pytest.raises(NotImplementedError):
    dist.mode()
  with pytest.raises(NotImplementedError):
    dist.quantile(0.75)

  class Ceil(D.MonotonicAdapter):
    invertible = False

    def transform(self, x):
      return torch.ceil(x)

    def inverse_transform(self, y):
      return torch.floor(y)

  dist = Ceil(D.Normal(1.5, 1.5))
  with pytest.raises(NotImplementedError):
    dist._lower_tail(0.01)
  with pytest.raises(NotImplementedError):
    dist._upper_tail(0.01)


def _log_prob_gradient_is_bounded(make, values):
  x = torch.tensor(values, requires_grad=True)
  p = make().log_prob(x)
  idx = p < -32.0
  p = torch.clamp_min(p, -32.0)
  dx, = torch.autograd.grad(p.sum(), x)
  assert bool((dx[idx] == 0).all())
  assert bool(torch.isfinite(dx).all()), dx


def test_noisy_soft_rounded_deep_factorized():
  df = D.NoisySoftRoundedDeepFactorized(init_scale=1e-3)   # scale -> 0: the unit-width uniform density
  assert torch.allclose(df.prob(torch.linspace(-1., 1., 10)), torch.tensor([0, 0, 0, 1, 1, 1, 1, 0, 0, 0.]), atol=1e-5)
  _log_prob_gradient_is_bounded(D.NoisySoftRoundedDeepFactorized, [0.0, 1.0, 2.0, 1e3])


@pytest.mark.parametrize("cls", [D.NoisyRoundedNormal, D.NoisySoftRoundedNormal])
def test_location_scale_family(cls):
  """round_adapters_test.py:181-262 (LocationScaleTest)."""
  assert cls(loc=3., scale=5.).batch_shape == () and cls(loc=[3., 2.], scale=5.).batch_shape == (2,)
  loc = torch.tensor(1., requires_grad=True)
  log_scale = torch.tensor(0., requires_grad=True)
  x = torch.randn(20, generator=torch.Generator().manual_seed(0))
  loss = -cls(loc=loc, scale=torch.exp(log_scale)).log_prob(x).mean()
  grads = torch.autograd.grad(loss, [loc, log_scale])
  assert all(g is not None and bool(torch.isfinite(g)) for g in grads)
  dist = cls(loc=5.0, scale=1e-7)
  assert torch.allclose(dist.prob(torch.linspace(4., 6., 10)), torch.tensor([0, 0, 0, 1, 1, 1, 1, 0, 0, 0.]), atol=1e-5)
  assert cls(loc=0., scale=[3., 5.]).sample((5, 4)).shape == (5, 4, 2)
  dist = cls(loc=10., scale=1.5)
  assert float(dist._upper_tail(2**-8)) > float(dist._lower_tail(2**-8))
  dist = cls(loc=1., scale=2.)
  for call in (dist.mode, lambda: dist.quantile(.5), lambda: dist.survival_function(.5)):
    with pytest.raises(NotImplementedError):
      call()


def test_noisy_soft_rounded_normal_gradient_is_bounded():
  _log_prob_gradient_is_bounded(lambda: D.NoisySoftRoundedNormal(loc=0.0, scale=1.0), [0.0, 1.0, 2.0, 1e3])


def test_rounded_prior_builds_integer_aligned_tables():
  """A NoisyRounded prior has zero quantisation offset and integer tails (round_adapters.py:160-169), i.e. the
  table support of a ContinuousBatchedEntropyModel built from it starts and ends on integers."""
  d = D.NoisyRoundedNormal(loc=[0.3, -1.7], scale=[2., 3.])
  assert torch.equal(D.quantization_offset(d), torch.zeros(()))
  lo, hi = D.lower_tail(d, 2**-8), D.upper_tail(d, 2**-8)
  assert torch.equal(lo, torch.round(lo)) and torch.equal(hi, torch.round(hi))


@pytest.mark.parametrize("cls", [D.NoisyNormalMixture, D.NoisyLogisticMixture])
def test_noisy_mixtures(cls):
  """uniform_noise_test.py:105-172 (MixtureTest)."""
  assert cls(loc=[3., -3.], scale=[5., 2.5], weight=[.3, .7]).batch_shape == ()
  assert cls(loc=[[3., -3.], [2., -2.]], scale=[5., 2.5], weight=[.3, .7]).batch_shape == (2,)
  loc = torch.ones(2, requires_grad=True)
  log_scale = torch.zeros(2, requires_grad=True)
  logit = torch.tensor([.3, .7], requires_grad=True)
  x = torch.randn(20, generator=torch.Generator().manual_seed(0))
  loss = -cls(loc=loc, scale=torch.exp(log_scale), weight=torch.softmax(logit, 0)).log_prob(x).mean()
  assert all(g is not None and bool(torch.isfinite(g).all()) for g in torch.autograd.grad(loss, [loc, log_scale, logit]))
  # scales -> 0: a mixture of unit-width uniform densities
  dist = cls(loc=[2.5, -1.], scale=[1e-7, 1e-7], weight=[.3, .7])
  box = torch.tensor([0, 0, 0, 1, 1, 1, 1, 0, 0, 0.])
  assert torch.allclose(dist.prob(torch.linspace(1.5, 3.5, 10)), .3 * box, atol=1e-5)
  assert torch.allclose(dist.prob(torch.linspace(-2., 0., 10)), .7 * box, atol=1e-5)
  assert cls(loc=[[0.]], scale=[3., 5.], weight=[.2, .8]).sample((5, 4)).shape == (5, 4, 1)
  dist = cls(loc=[5.4, 8.6], scale=[1.4, 2.], weight=[.6, .4])
  assert float(D.upper_tail(dist, 2**-8)) > float(D.lower_tail(dist, 2**-8))
  assert math.isclose(float(D.quantization_offset(dist)), 0.4, abs_tol=1e-5)   # decimal part of the peakiest mode
  assert float(dist.base.cdf(D.lower_tail(dist, 2**-8))) <= 2**-8 * 0.51
  dist = cls(loc=[1., 0.], scale=2., weight=[.1, .9])
  for call in (dist.mode, lambda: dist.quantile(.5), lambda: dist.survival_function(.5)):
    with pytest.raises(NotImplementedError):
      call()
  dist = cls(loc=[0., 0.], scale=[0., 0.], weight=[.5, .5])   # all mass at 0
  assert torch.allclose(dist.prob([0.]), torch.ones(1)) and torch.allclose(dist.prob([1.]), torch.zeros(1))


def test_mixture_prior_trains_an_entropy_model_on_cpu():
  import compression_b200 as tfc
  prior = tfc.NoisyNormalMixture(loc=torch.tensor([[-2., 2.]] * 3), scale=torch.tensor([[1., 1.5]] * 3),
                                 weight=torch.tensor([.4, .6]))
  assert prior.batch_shape == (3,)
  em = tfc.ContinuousBatchedEntropyModel(prior, coding_rank=1, compression=False)
  x = torch.randn(5, 3, generator=torch.Generator().manual_seed(0)) * 2
  xt, bits = em(x, training=True)
  assert xt.shape == x.shape and bits.shape == (5,) and bool((bits > 0).all())
