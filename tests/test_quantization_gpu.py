"""StochasticRound on the GPU: the reference's own tests (python/ops/quantization_ops_test.py:28-83) and bit equality
with the oracle's sequential stream (cc/kernels/quantization_kernels.cc:48-95) for the same seed."""
import time

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
  from compression_b200 import gen_ops
  return gen_ops


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_difference_is_at_most_one(ops, dtype):
  values = (torch.rand(100) * 200 - 100).to(dtype)
  rounded = ops.stochastic_round(values, 1., ())
  assert rounded.dtype == torch.int32
  assert float((values.float().cuda() - rounded.float()).abs().max()) <= 1


def test_identical_seed_yields_identical_output(ops):
  values = torch.rand(100) * 200 - 100
  r1 = ops.stochastic_round(values, 1., (123, 456))
  r2 = ops.stochastic_round(values, 1., (123, 456))
  r3 = ops.stochastic_round(values, 1., (456, 789))
  assert torch.equal(r1, r2) and not torch.equal(r1, r3)


def test_clock_seed_yields_different_output(ops):
  values = torch.rand(100) * 200 - 100
  r1 = ops.stochastic_round(values, 1., ())
  time.sleep(0.01)
  r2 = ops.stochastic_round(values, 1., ())
  assert not torch.equal(r1, r2)


@pytest.mark.parametrize("step_size", [1., .75, 1e-4])
def test_rounding_is_deterministic_at_integers(ops, step_size):
  values = torch.randint(-100, 100, (100,), dtype=torch.int32)
  rounded = ops.stochastic_round(step_size * values.float(), step_size, ())
  assert torch.equal(rounded.cpu(), values)


@pytest.mark.parametrize("step_size", [1., .75, 1e-4])
def test_difference_at_half_integers_is_at_most_one_half(ops, step_size):
  values = torch.arange(-10, 10, dtype=torch.float32) + .5
  rounded = ops.stochastic_round(step_size * values, step_size, ())
  assert float((values - rounded.float().cpu()).abs().max()) <= .5


def test_rounding_is_unbiased(ops):
  values = torch.rand(20) * 200 - 100
  rounded = ops.stochastic_round(values.expand(100000, 20).contiguous(), 1., ())
  assert float((rounded.float().mean(0).cpu() - values).abs().max()) <= 1e-2


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 65537, 3_000_001])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_same_integers_as_the_sequential_reference_stream(ops, n, dtype):
  """The parallel jump-ahead enters the reference's single xoshiro256+ stream at every 32nd element."""
  g = torch.Generator().manual_seed(n)
  values = ((torch.rand(n, generator=g) * 200 - 100) * 0.37).to(dtype)
  for seed, step in (((123, 456), 0.25), ((7,), 1.0), (tuple(range(-5, 6)), 3e-2)):
    got = ops.stochastic_round(values, step, seed).cpu().numpy()
    want = oracle.best().stochastic_round(values.float().numpy(), step, seed)
    assert np.array_equal(got, want)


def test_argument_errors(ops):
  with pytest.raises(ValueError, match="scalar"):
    ops.stochastic_round(torch.zeros(4), torch.ones(2), ())
  with pytest.raises(ValueError, match="dtype"):
    ops.stochastic_round(torch.zeros(4, dtype=torch.float64), 1., ())
  assert ops.stochastic_round(torch.zeros(0), 1., (1,)).shape == (0,)
