"""CPU: the entropy models' host logic that needs no range coder (compression=False): construction contract,
quantisation, straight-through gradients, the differentiable bit cost, index normalisation.  Cases from
py/entropy_models/continuous_batched_test.py:27-101 and continuous_indexed_test.py:96-137."""
import numpy as np
import pytest
import scipy.stats
import torch

import compression_b200 as tfc


def test_batched_model_contract_without_compression():
  noisy = tfc.NoisyNormal(loc=torch.tensor(0.), scale=torch.tensor(1.))
  em = tfc.ContinuousBatchedEntropyModel(noisy, 1)
  assert em.prior is noisy and em.coding_rank == 1 and em.tail_mass == 2**-8
  assert em.bottleneck_dtype == torch.float32 and em.compression is False
  x = torch.randn(100)
  with pytest.raises(RuntimeError):   # continuous_batched_test.py:93-101
    em.compress(x)
  with pytest.raises(RuntimeError):
    em.decompress([b""], [100])
  with pytest.raises(ValueError):     # :65-72 coding_rank must cover the prior's batch dimensions
    tfc.ContinuousBatchedEntropyModel(tfc.NoisyLaplace(loc=torch.zeros(1, 1), scale=torch.ones(1, 1)), 1)


def test_batched_quantizes_to_integers_modulo_offset_with_straight_through_gradients():
  torch.manual_seed(0)
  noisy = tfc.NoisyNormal(loc=torch.tensor(.25), scale=torch.tensor(10.))
  em = tfc.ContinuousBatchedEntropyModel(noisy, 1)
  x = (torch.randn(100) * 10 + .25).requires_grad_(True)
  q = em.quantize(x)
  np.testing.assert_allclose(((q.detach() - .25) % 1).numpy(), 0.0, atol=1e-5)   # :74-80
  assert float((q.detach() - x.detach()).abs().max()) <= 0.5 + 1e-6
  q.sum().backward()
  assert x.grad.tolist() == [1.0] * 100                                            # :82-91


def test_batched_bit_cost_matches_the_noisy_prior():
  torch.manual_seed(1)
  loc, scale = torch.tensor([.5, -1.0]), torch.tensor([2.0, 0.7])
  em = tfc.ContinuousBatchedEntropyModel(tfc.NoisyNormal(loc=loc, scale=scale), coding_rank=2)
  x = torch.randn(3, 50, 2) * scale + loc
  x_hat, bits = em(x, training=False)
  assert bits.shape == (3,) and torch.equal(x_hat, em.quantize(x))
  ref = scipy.stats.norm(loc=loc.numpy().astype(np.float64), scale=scale.numpy().astype(np.float64))
  xq = x_hat.numpy().astype(np.float64)
  want = -np.log2(ref.cdf(xq + .5) - ref.cdf(xq - .5)).sum(axis=(1, 2))
  np.testing.assert_allclose(bits.numpy(), want, rtol=2e-4)
  # training: additive uniform noise in [-1/2, 1/2], bits finite and differentiable w.r.t. the input
  xin = x.clone().requires_grad_(True)
  x_noisy, bits_t = em(xin, training=True)
  assert float((x_noisy.detach() - x).abs().max()) <= 0.5 and bool(torch.isfinite(bits_t).all())
  bits_t.sum().backward()
  assert xin.grad is not None and bool(torch.isfinite(xin.grad).all())


def _indexed(**kw):
  return tfc.ContinuousIndexedEntropyModel(
      tfc.NoisyNormal, index_ranges=(10, 10),
      parameter_fns=dict(loc=lambda i: i[..., 0] - 5.0, scale=lambda i: torch.exp(i[..., 1] / 3.0 - 2)),
      coding_rank=1, channel_axis=-1, **kw)


def test_indexes_are_clipped_and_flattened():
  em = _indexed()
  idx = torch.tensor([[[-1.0, 3.2], [4.9, 12.0]]])
  norm = em._normalize_indexes(idx)                     # continuous_indexed_test.py:96-103
  assert norm.tolist() == [[[0.0, pytest.approx(3.2)], [pytest.approx(4.9), 9.0]]]
  assert em._flatten_indexes(norm).tolist() == [[3, 49]]   # row = i0 * 10 + i1, truncated towards zero
  with pytest.raises(RuntimeError):                      # :130-137
    em.compress(torch.zeros(1, 2), idx)


def test_indexed_quantizes_to_integers_and_bits_follow_the_indexed_prior():
  torch.manual_seed(2)
  em = _indexed()
  idx = torch.stack([torch.randint(0, 10, (4, 30)).float(), torch.randint(0, 10, (4, 30)).float()], dim=-1)
  x = torch.randn(4, 30) * 3
  x_hat, bits = em(x, idx, training=False)
  assert torch.equal(x_hat, torch.round(x)) and bits.shape == (4,)   # :113-118 (no offset in indexed models)
  loc = (idx[..., 0] - 5.0).numpy().astype(np.float64)
  scale = np.exp(idx[..., 1].numpy().astype(np.float64) / 3.0 - 2)
  xq = x_hat.numpy().astype(np.float64)
  p = np.where(xq > loc, scipy.stats.norm.sf(xq - .5, loc, scale) - scipy.stats.norm.sf(xq + .5, loc, scale),
               scipy.stats.norm.cdf(xq + .5, loc, scale) - scipy.stats.norm.cdf(xq - .5, loc, scale))
  # the model floors the likelihood (likelihood bound) so that a symbol far in a narrow prior's tail costs a
  # finite number of bits: compare where the closed form is representable
  ok = p > 1e-8
  got = em(torch.where(torch.from_numpy(ok), x, torch.from_numpy(loc).float()), idx, training=False)[1]
  want = -np.log2(np.where(ok, p, scipy.stats.norm.cdf(.5, 0, scale) - scipy.stats.norm.cdf(-.5, 0, scale))).sum(axis=1)
  np.testing.assert_allclose(got.numpy(), want, rtol=2e-3)


def test_prior_handed_to_the_model_is_registered_like_a_tf_module_attribute():
  """continuous_batched.py:205 (`self._prior = prior` on a tf.Module): the prior's variables are the model's
  trainable variables and checkpoint state."""
  prior = tfc.NoisyDeepFactorized(batch_shape=(4,))
  em = tfc.ContinuousBatchedEntropyModel(prior, coding_rank=1)
  names = {n for n, _ in em.named_parameters()}
  assert len(names) == len(list(prior.parameters())) > 0
  assert all(n.startswith("_prior.") for n in names)
  assert set(em.state_dict()) >= {"_prior." + k for k in prior.state_dict()}
  eb = tfc.EntropyBottleneck(num_channels=4, compression=False)
  assert len(list(eb.parameters())) == 8         # 3 matrices + 3 biases + 2 factors for (3, 3) filters
  # an optimiser built from model.parameters() moves the prior
  x = torch.randn(16, 4)
  opt = torch.optim.SGD(em.parameters(), lr=0.1)
  before = [p.detach().clone() for p in prior.parameters()]
  _, bits = em(x, training=True)
  bits.sum().backward()
  opt.step()
  assert any(not torch.equal(a, b) for a, b in zip(before, prior.parameters()))
  del em.prior
  assert list(em.parameters()) == []
  with pytest.raises(RuntimeError):
    em.prior  # pylint:disable=pointless-statement


def test_deep_factorized_density_is_differentiable_in_its_input():
  """deep_factorized.py:195-230: prob/log_prob are functions of x (finite-difference check, factors != 0)."""
  torch.manual_seed(3)
  d = tfc.DeepFactorized(batch_shape=(3,), dtype=torch.float64)
  with torch.no_grad():
    for f in d.factors:
      f.copy_(torch.randn_like(f))
  x = torch.randn(5, 3, dtype=torch.float64).requires_grad_(True)
  for fn in (d.prob, d.log_prob):
    g, = torch.autograd.grad(fn(x).sum(), x)
    h = 1e-5
    fd = (fn((x + h).detach()) - fn((x - h).detach())) / (2 * h)
    np.testing.assert_allclose(g.numpy(), fd.detach().numpy(), rtol=1e-5, atol=1e-8)
  # and under no_grad they still evaluate
  with torch.no_grad():
    assert d.prob(x).shape == (5, 3) and not d.log_prob(x).requires_grad


def test_construction_argument_checks():
  """continuous_batched_test.py:65-72: the coding unit must cover the prior's batch dimensions."""
  noisy = tfc.NoisyLogistic(loc=0., scale=[[1.], [2.]])
  for coding_rank in (0, 1):
    with pytest.raises(ValueError):
      tfc.ContinuousBatchedEntropyModel(noisy, coding_rank)
  for coding_rank in (2, 3):
    tfc.ContinuousBatchedEntropyModel(noisy, coding_rank)
  with pytest.raises(ValueError):
    tfc.ContinuousBatchedEntropyModel(tfc.NoisyNormal(loc=0., scale=1.), 1, tail_mass=1.5)


def test_laplace_tail_mass_mixes_in_a_heavy_tail():
  """continuous_batched_test.py:244-252 and continuous_base.py:298-334: with `laplace_tail_mass` the bit cost far
  from the prior's mass follows the unit Laplace tail instead of the Gaussian one."""
  noisy = tfc.NoisyNormal(loc=0., scale=1.)
  assert tfc.ContinuousBatchedEntropyModel(noisy, 1, laplace_tail_mass=0.0).laplace_tail_mass == 0.0
  em = tfc.ContinuousBatchedEntropyModel(noisy, 1, laplace_tail_mass=torch.tensor(1e-3))
  assert float(em.laplace_tail_mass) == pytest.approx(1e-3)
  lp = em._log_prob(noisy, torch.tensor(0.0))
  assert lp.dtype == torch.float32
  x = torch.tensor([0., 30., 60.])
  with_tail, without = em._log_prob(noisy, x), tfc.ContinuousBatchedEntropyModel(noisy, 1)._log_prob(noisy, x)
  assert torch.allclose(with_tail[0], without[0], atol=2e-3)           # at the mode the mixture barely matters
  lap = tfc.NoisyLaplace(loc=0., scale=1.)
  assert torch.allclose(with_tail[1:], np.log(1e-3) + lap.log_prob(x[1:]), atol=1e-4)
  assert bool((with_tail[1:] > without[1:] + 100).all())
  with pytest.raises(ValueError):
    tfc.ContinuousBatchedEntropyModel(noisy, 1, laplace_tail_mass=1.0)._log_prob(noisy, x)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_dtypes_with_a_sixteen_bit_bottleneck(dtype):
  """continuous_batched_test.py:197-216 without the coder: a float64 prior under a 16-bit bottleneck gives a
  16-bit perturbed tensor and float64 bits; a prior of the bottleneck's own dtype works on the CPU too."""
  noisy = tfc.NoisyNormal(loc=torch.tensor(.5, dtype=torch.float64), scale=torch.tensor(1., dtype=torch.float64),
                          dtype=torch.float64)
  em = tfc.ContinuousBatchedEntropyModel(noisy, 1, bottleneck_dtype=dtype)
  assert em.bottleneck_dtype == dtype and em.prior.dtype == torch.float64
  x = torch.randn(2, 5, generator=torch.Generator().manual_seed(1)).to(dtype)
  x_tilde, bits = em(x)
  assert x_tilde.dtype == dtype and float((x.float() - x_tilde.float()).abs().max()) <= .5 + 2e-2
  assert bits.dtype == torch.float64 and bits.shape == (2,) and bool((bits >= 0).all())
  assert em.quantize(x).dtype == dtype
  for cls in (tfc.NoisyNormal, tfc.NoisyLogistic, tfc.NoisyLaplace):
    em = tfc.ContinuousBatchedEntropyModel(cls(loc=0., scale=1., dtype=dtype), 1, bottleneck_dtype=dtype)
    _, bits = em(x)
    assert bits.dtype == dtype and bool(torch.isfinite(bits).all())
