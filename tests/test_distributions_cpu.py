"""CPU: host logic of the table build -- tails, quantisation offsets, noisy PMFs -- mirroring
py/distributions/helpers_test.py:25-92, uniform_noise_test.py:46-90, deep_factorized_test.py:51-64, with scipy as
the closed-form reference.  Pure torch: no CUDA library needed."""
import numpy as np
import pytest
import scipy.stats
import torch

from compression_b200 import distributions as D


def test_estimate_tails_terminates_on_nan_and_perfect_guess():
  # helpers_test.py:25-33
  D.estimate_tails(lambda x: torch.tanh(x) * float("nan"), 0.5, (), torch.float32)
  D.estimate_tails(torch.tanh, 0.0, (), torch.float32)


@pytest.mark.parametrize("cls,loc,scale", [(D.Laplace, -2., 5.), (D.Logistic, -3., 1.), (D.Normal, 3., 5.)])
def test_quantizes_to_mode_decimal_part_and_tails_in_order(cls, loc, scale):
  # helpers_test.py:43-83
  dist = cls(torch.tensor(loc), torch.tensor(scale))
  assert float(D.quantization_offset(dist)) == 0.0
  assert float(D.upper_tail(dist, 2**-8)) > float(D.lower_tail(dist, 2**-8))
  dist = cls(torch.tensor(1.4), torch.tensor(scale))
  assert abs(float(D.quantization_offset(dist)) - 0.4) < 1e-6


@pytest.mark.parametrize("cls,sp", [(D.Normal, scipy.stats.norm), (D.Laplace, scipy.stats.laplace),
                                    (D.Logistic, scipy.stats.logistic)])
def test_loc_scale_families_match_scipy(cls, sp):
  loc, scale = torch.tensor([-1.5, 0.0, 2.25]), torch.tensor([0.3, 1.0, 7.0])
  dist = cls(loc, scale)
  x = torch.linspace(-30, 30, 241).reshape(-1, 1)
  ref = sp(loc=loc.numpy(), scale=scale.numpy())
  np.testing.assert_allclose(dist.cdf(x).numpy(), ref.cdf(x.numpy()), rtol=2e-5, atol=1e-7)
  np.testing.assert_allclose(dist.survival_function(x).numpy(), ref.sf(x.numpy()), rtol=2e-5, atol=1e-7)
  lc, rl = dist.log_cdf(x).numpy(), ref.logcdf(x.numpy())
  ok = rl > -80  # beyond that fp32 log-cdf forms differ only in how they underflow
  np.testing.assert_allclose(lc[ok], rl[ok], rtol=2e-4, atol=2e-6)
  q = torch.tensor([[2.0**-8], [0.25], [0.5], [0.9]])
  np.testing.assert_allclose(dist.quantile(q).numpy(), ref.ppf(q.numpy()), rtol=2e-5, atol=1e-5)
  # tails = quantiles of tail_mass / 2 (helpers.py:160-219)
  np.testing.assert_allclose(D.lower_tail(dist, 2**-8).numpy(), ref.ppf(2**-9), rtol=1e-4, atol=1e-4)
  np.testing.assert_allclose(D.upper_tail(dist, 2**-8).numpy(), ref.isf(2**-9), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("cls,sp", [(D.NoisyNormal, scipy.stats.norm), (D.NoisyLaplace, scipy.stats.laplace),
                                    (D.NoisyLogistic, scipy.stats.logistic)])
def test_noisy_families_are_pmfs_on_the_integer_grid(cls, sp):
  # uniform_noise_test.py:56-61 and the closed form  p(y) = F(y + 1/2) - F(y - 1/2)
  loc, scale = torch.tensor([0.3, -2.0]), torch.tensor([0.7, 4.0])
  dist = cls(loc=loc, scale=scale)
  y = torch.arange(-200, 201, dtype=torch.float32).reshape(-1, 1) + 0.0
  p = dist.prob(y)
  np.testing.assert_allclose(p.sum(0).numpy(), [1.0, 1.0], atol=2e-5)
  ref = sp(loc=loc.numpy().astype(np.float64), scale=scale.numpy().astype(np.float64))
  y64 = y.numpy().astype(np.float64)
  want = np.where(y64 > loc.numpy(), ref.sf(y64 - .5) - ref.sf(y64 + .5), ref.cdf(y64 + .5) - ref.cdf(y64 - .5))
  np.testing.assert_allclose(p.numpy(), want, rtol=2e-4, atol=2e-7)
  # log_prob works in log space (uniform_noise.py:158-183): accurate far into the tails
  big = want > 1e-30
  np.testing.assert_allclose(dist.log_prob(y).numpy()[big], np.log(want[big]), rtol=2e-3, atol=2e-3)
  # tails and offset in order (uniform_noise_test.py:68-74)
  lo, hi, off = D.lower_tail(dist, 2**-8), D.upper_tail(dist, 2**-8), D.quantization_offset(dist)
  assert bool((lo < off).all()) and bool((off < hi).all())


def test_deep_factorized_is_a_distribution_and_logistic_is_special_case():
  torch.manual_seed(0)
  dist = D.DeepFactorized(batch_shape=(10,))
  x = torch.linspace(-400, 400, 801).reshape(-1, 1)
  cdf = dist.cdf(x).detach()
  assert bool((cdf[1:] >= cdf[:-1] - 1e-7).all()) and float(cdf[0].max()) < 1e-4 and float(cdf[-1].min()) > 1 - 1e-4
  np.testing.assert_allclose((cdf + dist.survival_function(x).detach()).numpy(), 1.0, atol=1e-6)
  # the estimated tails hold tail_mass / 2 each (helpers.py:160-219; found by gradient iteration, so loosely)
  lo, hi = D.lower_tail(dist, 2**-8).detach(), D.upper_tail(dist, 2**-8).detach()
  np.testing.assert_allclose(dist.cdf(lo).detach().numpy(), 2**-9, rtol=0.05)
  np.testing.assert_allclose(dist.survival_function(hi).detach().numpy(), 2**-9, rtol=0.05)
  assert bool((D.upper_tail(dist, 2**-8) - D.lower_tail(dist, 2**-8) > 0).all())       # helpers_test.py:84-87
  noisy = D.NoisyDeepFactorized(batch_shape=(10,))
  assert bool((D.upper_tail(noisy, 2**-8) - D.lower_tail(noisy, 2**-8) > 0).all())     # helpers_test.py:89-92
  # deep_factorized_test.py:51-64: no hidden units -> a logistic with scale 1 / softplus-reparameterised slope
  df = D.DeepFactorized(batch_shape=(), num_filters=(), init_scale=1)
  loc = -float(df.biases[0][0, 0, 0].detach())
  xs = torch.linspace(-5, 5, 20)
  ref = scipy.stats.logistic(loc=loc, scale=1.0)
  np.testing.assert_allclose(df.cdf(xs).detach().numpy(), ref.cdf(xs.numpy()), atol=1e-5)
  np.testing.assert_allclose(df.log_survival_function(xs).detach().numpy(), ref.logsf(xs.numpy()), atol=1e-5)
  np.testing.assert_allclose(df.prob(xs).detach().numpy(), ref.pdf(xs.numpy()), atol=1e-5)


def test_noisy_gradients_reach_the_parameters():
  # uniform_noise_test.py:35-44
  loc = torch.tensor(0.2, requires_grad=True)
  scale = torch.tensor(1.3, requires_grad=True)
  dist = D.NoisyNormal(loc=loc, scale=scale)
  dist.log_prob(torch.tensor([0.0, 1.0, -2.5])).sum().backward()
  assert loc.grad is not None and scale.grad is not None and float(loc.grad.abs()) > 0 and float(scale.grad.abs()) > 0


@pytest.mark.parametrize("cls", [D.NoisyNormal, D.NoisyLogistic, D.NoisyLaplace])
def test_noisy_location_scale_family_contract(cls):
  """uniform_noise_test.py:28-98 (LocationScaleTest): shapes, the unit-width uniform limit, sampling, tails, and the
  statistics the noisy density does not define."""
  assert cls(loc=3., scale=5.).batch_shape == () and cls(loc=[3., 2.], scale=5.).batch_shape == (2,)
  dist = cls(loc=5.0, scale=1e-7)
  assert torch.allclose(dist.prob(torch.linspace(4., 6., 10)), torch.tensor([0, 0, 0, 1, 1, 1, 1, 0, 0, 0.]), atol=1e-5)
  s = cls(loc=0., scale=[3., 5.]).sample((5, 4), generator=torch.Generator().manual_seed(0))
  assert s.shape == (5, 4, 2) and bool(torch.isfinite(s).all())
  dist = cls(loc=10., scale=1.5)
  assert float(dist._upper_tail(2**-8)) > float(dist._lower_tail(2**-8))
  dist = cls(loc=1., scale=2.)
  for call in (dist.mode, lambda: dist.quantile(.5), lambda: dist.survival_function(.5)):
    with pytest.raises(NotImplementedError):
      call()
  # samples follow the density: the empirical mean of a wide distribution is its location
  big = cls(loc=-4., scale=2.).sample((40000,), generator=torch.Generator().manual_seed(1))
  assert abs(float(big.mean()) + 4.) < 0.08


def test_deep_factorized_contract():
  """deep_factorized_test.py:32-70,81-106,126-143: defaults, broadcasting of a batch of densities over an input,
  eight trainable variables, and the statistics that have no closed form."""
  df = D.DeepFactorized()
  assert df.batch_shape == () and tuple(df.num_filters) == (3, 3) and df.init_scale == 10
  df = D.DeepFactorized(batch_shape=(2, 3))
  x = torch.linspace(-5., 5., 20).reshape(4, 5, 1, 1)
  for method in ("prob", "log_prob", "cdf", "log_cdf", "survival_function", "log_survival_function"):
    assert getattr(df, method)(x).shape == (4, 5, 2, 3)
  noisy = D.NoisyDeepFactorized(num_filters=(2, 3, 4))
  assert noisy.batch_shape == () and tuple(noisy.base.num_filters) == (2, 3, 4)
  assert noisy.prob(torch.randn(10)).shape == (10,)
  noisy = D.NoisyDeepFactorized(batch_shape=(4, 3))
  assert noisy.prob(torch.randn(10, 4, 3)).shape == (10, 4, 3)
  noisy = D.NoisyDeepFactorized()
  loss = -noisy.log_prob(torch.randn(20)).mean()
  grads = torch.autograd.grad(loss, list(noisy.parameters()))
  assert len(grads) == 8 and all(g is not None for g in grads)
  assert float(D.upper_tail(noisy, 2**-8)) > float(D.lower_tail(noisy, 2**-8))
  for call in (noisy.mode, noisy.mean, lambda: noisy.quantile(.5), lambda: noisy.survival_function(.5), noisy.sample):
    with pytest.raises(NotImplementedError):
      call()
