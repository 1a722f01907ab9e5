"""CPU: the universal-quantisation entropy models' host logic (py/entropy_models/universal.py:30-62,147-211,446-528;
cases from universal_test.py:29-58,121-150,349-375) and the counter-based noise-level stream that stands in for
tf.random.stateless_uniform."""
import math

import numpy as np
import pytest
import torch

import compression_b200 as tfc
from compression_b200 import entropy_models as E


def test_philox_known_answers_and_stream_properties():
  """Philox-4x32-10 known-answer vectors of the Random123 distribution (kat_vectors: zero and all-ones inputs)."""
  zero = torch.zeros(1, 4, dtype=torch.int64)
  assert [int(v) for v in E._philox4x32(zero, (0, 0))[0]] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
  ones = torch.full((1, 4), 0xFFFFFFFF, dtype=torch.int64)
  assert [int(v) for v in E._philox4x32(ones, (0xFFFFFFFF, 0xFFFFFFFF))[0]] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
  a = E.stateless_uniform_int((7, 11, 3), (1234, 1234), 15)
  b = E.stateless_uniform_int((7 * 11 * 3,), (1234, 1234), 15)
  assert a.dtype == torch.int32 and torch.equal(a.reshape(-1), b)          # depends on (seed, position) only
  assert int(a.min()) >= 0 and int(a.max()) < 15
  big = E.stateless_uniform_int((150000,), (1234, 1234), 15).numpy()
  assert np.abs(np.bincount(big, minlength=15) / 150000 - 1 / 15).max() < 0.004
  assert not torch.equal(a, E.stateless_uniform_int((7, 11, 3), (1234, 1235), 15))


def test_offsets_lie_on_the_noise_grid_and_quantisation_noise_is_uniform():
  """universal.py:45-62 (levels k -> (k + 1)/(n + 1) - 1/2) and universal_test.py:349-375."""
  off = E._range_coding_offsets(15, 1, torch.float32)
  assert off.shape == (15, 1)
  np.testing.assert_allclose(off[:, 0].numpy(), (np.arange(15) + 1) / 16 - .5, rtol=0, atol=1e-7)
  prior = tfc.NoisyLogistic(loc=torch.zeros(3), scale=torch.full((3,), 4.))
  em = tfc.UniversalBatchedEntropyModel(prior, coding_rank=2, num_noise_levels=15)
  x = torch.randn(4, 20000, 3) * 4
  x_hat, bits = em(x, training=False)
  noise = (x_hat - x).numpy().ravel()
  assert np.abs(noise).max() <= 0.5 + 1e-6 and bits.shape == (4,)
  assert abs(noise.mean()) < 5e-3 and abs(noise.var() - 1 / 12) < 2e-3
  # decoded values are integers plus the shared offset of their position
  _, offset = em._compute_indexes_and_offset((20000,), x.device)
  np.testing.assert_allclose(((x_hat - offset) - torch.round(x_hat - offset)).numpy(), 0, atol=1e-5)


def test_batched_bits_estimates_agree_between_training_and_inference_and_expected_grads_flow():
  torch.manual_seed(0)
  prior = tfc.NoisyDeepFactorized(batch_shape=(2,))
  em = tfc.UniversalBatchedEntropyModel(prior, coding_rank=2, expected_grads=True)
  x = (torch.randn(3, 4000, 2) * 6).requires_grad_(True)
  _, bits_train = em(x, training=True)
  _, bits_eval = em(x, training=False)
  np.testing.assert_allclose(bits_train.detach().numpy(), bits_eval.detach().numpy(), rtol=0.02)
  bits_train.sum().backward()
  assert x.grad is not None and all(p.grad is not None for p in prior.parameters())
  with pytest.raises(RuntimeError):
    em.compress(x)
  with pytest.raises(ValueError):
    tfc.UniversalBatchedEntropyModel(tfc.NoisyLogistic(loc=torch.zeros(2, 2), scale=torch.ones(2, 2)), coding_rank=1)


def test_indexed_model_indexes_are_clipped_flattened_with_the_noise_level_first():
  em = tfc.UniversalIndexedEntropyModel(tfc.NoisyLogistic, (5, 3), dict(loc=lambda i: i[..., 0], scale=lambda i: 1. + i[..., 1]),
                                        coding_rank=1, num_noise_levels=7)
  assert em.index_ranges == (7, 5, 3) and em.index_ranges_without_offsets == (5, 3)
  idx = torch.tensor([[[-3., 9.], [4.2, 1.7]]])
  np.testing.assert_allclose(em._normalize_indexes(idx).numpy(), [[[0., 2.], [4., 1.7]]], rtol=1e-6)
  with_off = torch.tensor([[[9., 4., 2.], [-1., 0., 0.]]])
  np.testing.assert_array_equal(em._normalize_indexes(with_off).numpy(), [[[6., 4., 2.], [0., 0., 0.]]])
  np.testing.assert_array_equal(em._flatten_indexes(torch.tensor([[6, 4, 2], [1, 0, 2]])).numpy(), [6 * 15 + 4 * 3 + 2, 15 + 2])
  x = torch.randn(2, 50) * 3
  indexes = torch.stack([torch.randint(0, 5, (2, 50)), torch.randint(0, 3, (2, 50))], -1).float()
  for training in (True, False):
    x_hat, bits = em(x, indexes, training=training)
    assert bits.shape == (2,) and float((x_hat - x).abs().max()) <= 0.5 + 1e-6
  with pytest.raises(ValueError):
    tfc.UniversalIndexedEntropyModel(tfc.NoisyLogistic, (5,), dict(loc=lambda i: i[..., 0]), coding_rank=0)


def _mixture_model(expected_grads, coding_rank=2):
  return tfc.UniversalIndexedEntropyModel(
      tfc.NoisyLogisticMixture, index_ranges=(10, 10, 5),
      parameter_fns=dict(loc=lambda i: i[..., 0:2] - 5, scale=lambda _: 1.,
                         weight=lambda i: torch.softmax((i[..., 2:3] - 2) * torch.tensor([-1., 1.]), -1)),
      coding_rank=coding_rank, expected_grads=expected_grads)


def test_n_dimensional_indexes_with_a_mixture_prior():
  """universal_test.py:152-167,377-417: three index dimensions parameterise a two-component logistic mixture; the
  quantisation noise stays within half a bin and the bit estimate does not depend on `expected_grads`."""
  em = _mixture_model(True, coding_rank=1)
  assert em.coding_rank == 1 and float(em.laplace_tail_mass) == 0.0 and em.tail_mass == 2**-8
  assert em.bottleneck_dtype == torch.float32
  g = torch.Generator().manual_seed(0)
  x = torch.randn(3, 2000, 16, generator=g)
  indexes = (10 * torch.rand(3, 2000, 16, 3, generator=g)).to(torch.int32)
  indexes[..., 2] //= 2
  em_expected, em_plain = _mixture_model(True), _mixture_model(False)
  x_hat, bits_expected = em_expected(x, indexes)
  assert float((x - x_hat).abs().max()) <= .5
  _, bits_plain = em_plain(x, indexes)
  assert bits_expected.shape == (3,)
  assert torch.allclose(bits_expected, bits_plain, rtol=0.01)
  x = x.requires_grad_(True)
  em_expected(x, indexes)[1].sum().backward()
  assert x.grad is not None and bool(torch.isfinite(x.grad).all()) and float(x.grad.abs().sum()) > 0


def test_laplace_tail_mass_bounds_the_cost_of_outliers():
  """universal_test.py:94-119: with `laplace_tail_mass` an input far outside the prior costs about |x| / ln 2 bits
  (the unit Laplace tail) instead of diverging, and nothing changes where the prior has its mass."""
  prior = tfc.NoisyDeepFactorized(batch_shape=(1,))
  em_tail = tfc.UniversalBatchedEntropyModel(prior, coding_rank=1, laplace_tail_mass=1e-3)
  em_plain = tfc.UniversalBatchedEntropyModel(prior, coding_rank=1)
  x = torch.tensor([1e3, 1e4, 1e5, 1e6])
  _, bits = em_tail(x[..., None])
  assert torch.allclose(bits, x / math.log(2.0), rtol=0.01)
  x = torch.linspace(-10.0, 10.0, 50)
  _, bits_tail = em_tail(x[..., None])
  _, bits_plain = em_plain(x[..., None])
  assert torch.allclose(bits_tail, bits_plain, rtol=0.01, atol=0.05)
