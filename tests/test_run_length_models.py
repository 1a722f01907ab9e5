"""CPU: host logic of PowerLawEntropyModel / LaplaceEntropyModel (power_law_test.py:22-120, laplace_test.py:22-116).
The coder itself is the CUDA one in the product (tests/test_run_length_gpu.py checks it against the oracle); here the
oracle's sequential restatement is injected in its place so that the model logic runs without a GPU."""
import numpy as np
import pytest
import torch

import oracle
from compression_b200 import gen_ops
from compression_b200 import run_length_models as M


@pytest.fixture
def cpu_coder(monkeypatch):
  O = oracle.port()
  monkeypatch.setattr(gen_ops, "run_length_encode",
                      lambda data, r, m, nz: O.run_length_encode(torch.as_tensor(data).cpu().numpy(), r, m, nz))
  monkeypatch.setattr(gen_ops, "run_length_decode",
                      lambda code, shape, r, m, nz: torch.from_numpy(O.run_length_decode(code, tuple(shape), r, m, nz)))


MODELS = [lambda **kw: M.PowerLawEntropyModel(coding_rank=1, **kw), lambda **kw: M.LaplaceEntropyModel(coding_rank=1, **kw)]


def _noisy_range(shape_tail=()):
  g = torch.Generator().manual_seed(5)
  x = torch.arange(-20., 20.).reshape((40,) + (1,) * len(shape_tail))
  return x, x + (torch.rand(x.shape, generator=g) - .5) * .98


@pytest.mark.parametrize("make", MODELS)
def test_instantiate_and_argument_checks(make):
  em = make()
  assert em.coding_rank == 1 and em.bottleneck_dtype == torch.float32
  for cls in (M.PowerLawEntropyModel, M.LaplaceEntropyModel):
    with pytest.raises(ValueError):
      cls(coding_rank=-1)
  with pytest.raises(ValueError):
    M.PowerLawEntropyModel(coding_rank=1, alpha=0.)
  with pytest.raises(ValueError):
    M.LaplaceEntropyModel(coding_rank=1, l1=0.)


@pytest.mark.parametrize("make", MODELS)
def test_quantizes_to_integers_with_straight_through_gradients(make):
  em = make()
  x, xp = _noisy_range()
  xp.requires_grad_(True)
  q = em.quantize(xp)
  assert torch.equal(q, x)
  q.sum().backward()
  assert torch.equal(xp.grad, torch.ones_like(xp))


@pytest.mark.parametrize("make", MODELS)
def test_compression_consistent_with_quantization(make, cpu_coder):
  em = make()
  _, xp = _noisy_range()
  strings = em.compress(xp)
  assert strings.shape == () and isinstance(strings[()], bytes)
  assert torch.equal(em.decompress(strings, xp.shape), em.quantize(xp))
  # batch axes to the left of the coding unit: one string each
  g = torch.Generator().manual_seed(6)
  y = torch.randn(3, 2, 50, generator=g) * 4
  s = em.compress(y)
  assert s.shape == (3, 2)
  assert torch.equal(em.decompress(s, (50,)), torch.round(y))


def test_coding_rank_zero_codes_every_element_on_its_own(cpu_coder):
  em = M.PowerLawEntropyModel(coding_rank=0)
  y = torch.tensor([[0., -3.2, 7.6], [1.1, 0.4, -0.6]])
  s = em.compress(y)
  assert s.shape == (2, 3)
  assert torch.equal(em.decompress(s, ()), torch.round(y))
  assert em.penalty(y).shape == (2, 3)


@pytest.mark.parametrize("make", MODELS)
def test_penalty_is_proportional_to_code_length(make, cpu_coder):
  em = make()
  _, xp = _noisy_range((1,))
  strings = em.compress(torch.broadcast_to(xp, (40, 100)))
  code_lengths = np.asarray([len(s) for s in strings], np.float32) * 8 / 100
  penalties = em.penalty(xp).numpy()
  assert np.corrcoef(code_lengths, penalties)[0, 1] > .96


@pytest.mark.parametrize("make", MODELS)
def test_penalty_is_nonnegative_and_differentiable(make):
  em = make()
  _, xp = _noisy_range((1,))
  xp.requires_grad_(True)
  p = em.penalty(xp)
  p.sum().backward()
  assert bool((p >= 0).all()) and torch.equal(torch.sign(xp.grad), torch.sign(xp.detach()))


@pytest.mark.parametrize("make", MODELS)
def test_dtypes_with_sixteen_bit_bottleneck(make, cpu_coder):
  em = make(bottleneck_dtype=torch.float16)
  x = torch.randn(2, 5, generator=torch.Generator().manual_seed(1)).to(torch.float16)
  x_tilde, penalty = em(x)
  x_hat = em.decompress(em.compress(x), (5,))
  assert x_hat.dtype == torch.float16 and x_tilde.dtype == torch.float16 and penalty.dtype == torch.float16
  assert penalty.shape == (2,)
  assert float((x - x_hat).abs().max()) <= .5 and float((x - x_tilde).abs().max()) <= .5


def test_laplace_code_parameters_reach_the_coder(cpu_coder):
  O = oracle.port()
  y = torch.round(torch.randn(200, generator=torch.Generator().manual_seed(2)) * 3)
  for r, m, nz in ((-1, 0, False), (2, 3, True), (0, -1, False)):
    em = M.LaplaceEntropyModel(coding_rank=1, run_length_code=r, magnitude_code=m, use_run_length_for_non_zeros=nz)
    s = em.compress(y)[()]
    assert s == O.run_length_encode(y.numpy().astype(np.int32), r, m, nz)
    assert torch.equal(em.decompress(s, (200,)), y)
  # PowerLaw is the gamma / gamma / False member of the family (run_length_ops.cc:34-37)
  assert M.PowerLawEntropyModel(coding_rank=1).compress(y)[()] == O.run_length_encode(y.numpy().astype(np.int32), -1, -1, False)
