"""RunLengthEncode / RunLengthDecode on the GPU against the sequential C restatement (oracle port), the reference test's
literal bit string (cc/kernels/run_length_kernels_test.cc:272-305) and its round-trip shapes (:129-183)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
CONFIGS = [(-1, -1, False), (-1, -1, True), (0, 0, False), (2, 3, True), (5, -1, False), (-1, 4, True), (3, 0, True)]


@pytest.fixture(scope="module")
def ops():
  from compression_b200 import gen_ops
  return gen_ops


def test_reference_literal_bit_string(ops):
  """EncodeConsistent / DecodeConsistent: [-6, 3, 0, 0] <-> {0b11010001, 0b01101101}."""
  code = ops.run_length_gamma_encode(torch.tensor([-6, 3, 0, 0], dtype=torch.int32))
  assert code == bytes([0b11010001, 0b01101101])
  assert ops.run_length_gamma_decode(code, [4]).tolist() == [-6, 3, 0, 0]
  assert ops.run_length_decode(bytes([0b11010001, 0b01101101]), [2, 2], -1, -1, False).tolist() == [[-6, 3], [0, 0]]


@pytest.mark.parametrize("rl,mg,nz", CONFIGS)
@pytest.mark.parametrize("kind", ["dense", "sparse", "leading", "trailing", "zeros", "nonzeros", "single", "big"])
def test_bytes_equal_the_sequential_restatement(ops, rl, mg, nz, kind):
  rng = np.random.default_rng(hash((rl, mg, nz, kind)) % (2**31))
  n = {"single": 1, "big": 300_000}.get(kind, 7001)
  d = rng.integers(-40, 41, n).astype(np.int32)
  if kind == "sparse" or kind == "big":
    d *= rng.random(n) < 0.07
  if kind == "leading":
    d[:500] = 0
  if kind == "trailing":
    d[-777:] = 0
  if kind == "zeros":
    d[:] = 0
  if kind == "nonzeros":
    d[d == 0] = 5
  if kind == "dense":
    d[::97] = np.iinfo(np.int32).min if mg < 0 else 100000   # int32 minimum is coded as its neighbour (gamma), :82-85
    d[1::97] = np.iinfo(np.int32).max if mg < 0 else -100000
  want = oracle.port().run_length_encode(d, rl, mg, nz)
  got = ops.run_length_encode(torch.from_numpy(d), rl, mg, nz)
  assert got == want
  back = ops.run_length_decode(got, [n], rl, mg, nz).cpu().numpy()
  assert np.array_equal(back, oracle.port().run_length_decode(want, (n,), rl, mg, nz))
  if not (kind == "dense" and mg < 0):
    assert np.array_equal(back, d)


def test_reference_round_trip_shapes(ops):
  """EncodeAndDecode{,LeadingZeros,TrailingZeros,InterspersedZeros}: shapes of run_length_kernels_test.cc:129-183."""
  rng = np.random.default_rng(5)
  for shape, zero in (((5, 70, 2), None), ((2, 80, 2), "lead"), ((50, 7, 2), "trail"), ((3, 7, 20), "mix")):
    d = rng.integers(-100, 100, shape).astype(np.int32)
    flat = d.reshape(-1)
    if zero == "lead":
      flat[:100] = 0
    if zero == "trail":
      flat[-100:] = 0
    if zero == "mix":
      flat[rng.random(flat.size) < 0.5] = 0
    for rl, mg, nz in CONFIGS:
      code = ops.run_length_encode(torch.from_numpy(d), rl, mg, nz)
      assert np.array_equal(ops.run_length_decode(code, list(shape), rl, mg, nz).cpu().numpy(), d)


def test_decode_errors_carry_the_reference_messages(ops):
  code = ops.run_length_gamma_encode(torch.tensor([0, 0, 7, -2, 0, 1], dtype=torch.int32))
  with pytest.raises(ValueError, match="Out of bits to read"):
    ops.run_length_gamma_decode(code[:1], [6])
  with pytest.raises(ValueError, match="Decoded past end of tensor"):
    ops.run_length_gamma_decode(code, [1])   # the first run length (2) already jumps past a one-element tensor
  with pytest.raises(ValueError, match="Exceeded maximum gamma bit width"):
    ops.run_length_gamma_decode(bytes([0, 0, 0, 0, 1]), [4])   # 32 zeros, then a one: width 33
  with pytest.raises(ValueError, match="shape"):
    ops.run_length_gamma_decode(code, [[6]])
  assert ops.run_length_gamma_encode(torch.zeros(0, dtype=torch.int32)) == b""


def test_power_law_and_laplace_entropy_models_on_the_cuda_coder():
  """power_law_test.py:50-56 / laplace_test.py:52-58 through the product path: every coding unit's string is the
  sequential restatement's, and decompress(compress(x)) == quantize(x)."""
  from compression_b200 import run_length_models as M
  O = oracle.port()
  g = torch.Generator().manual_seed(9)
  x = (torch.randn(3, 2, 500, generator=g) * 4 * (torch.rand(3, 2, 500, generator=g) < 0.3)).cuda()
  want = torch.round(x).cpu().numpy().astype(np.int32)
  for em, params in ((M.PowerLawEntropyModel(coding_rank=1), (-1, -1, False)),
                     (M.LaplaceEntropyModel(coding_rank=1), (-1, 0, False)),
                     (M.LaplaceEntropyModel(coding_rank=1, run_length_code=2, magnitude_code=3,
                                            use_run_length_for_non_zeros=True), (2, 3, True))):
    strings = em.compress(x)
    assert strings.shape == (3, 2)
    for i in range(3):
      for j in range(2):
        assert strings[i, j] == O.run_length_encode(want[i, j], *params)
    back = em.decompress(strings, (500,))
    assert back.dtype == torch.float32 and back.shape == (3, 2, 500)
    assert torch.equal(back.cpu(), em.quantize(x).cpu())


def test_golden_vectors_written_by_the_reference_bit_writer(ops):
  """tests/golden/run_length_golden.npz: strings produced by the reference's own BitWriter (compiled in place,
  oracle/make_run_length_golden.py); the CUDA coder writes the same bytes and reads them back."""
  import os
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "run_length_golden.npz"))
  at_d = at_c = 0
  for (rl, mg, nz), nd, nc in zip(g["params"], g["data_len"], g["code_len"]):
    d = np.ascontiguousarray(g["data"][at_d:at_d + nd]).astype(np.int32)
    code = bytes(g["code"][at_c:at_c + nc])
    at_d, at_c = at_d + nd, at_c + nc
    assert ops.run_length_encode(torch.from_numpy(d), int(rl), int(mg), bool(nz)) == code
    assert np.array_equal(ops.run_length_decode(code, [int(nd)], int(rl), int(mg), bool(nz)).cpu().numpy(), d)
