"""GPU tests of the entropy-model API, modelled on the reference's
python/entropy_models/continuous_batched_test.py:74-242 and continuous_indexed_test.py:139-145,
plus byte equality with the oracle for the strings each model produces."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tfc():
  import compression_b200 as m
  return m


def _oracle_strings(em, x, index=None):
  q = em.quantization_offset if hasattr(em, "quantization_offset") else None
  b = x.cpu() if q is None else x.cpu() - q.cpu()
  if index is None:
    sym = torch.round(b).to(torch.int32) - em.cdf_offset.cpu()
  else:
    sym = torch.round(b).to(torch.int32) - em.cdf_offset.cpu()[index.long().cpu()]
  S = int(np.prod(x.shape[:x.dim() - em.coding_rank])) if x.dim() > em.coding_rank else 1
  sym = sym.reshape(S, -1).numpy()
  idx = None if index is None else index.reshape(S, -1).cpu().numpy().astype(np.int32)
  return oracle.best().encode(em.cdf.cpu().numpy(), sym, idx)


def test_compression_consistent_with_quantization(tfc):
  """continuous_batched_test.py:103-109 + oracle byte equality."""
  noisy = tfc.NoisyNormal(loc=.25, scale=10.)
  em = tfc.ContinuousBatchedEntropyModel(noisy, 1, compression=True)
  x = torch.tensor(np.random.default_rng(0).normal(0.25, 10., size=(100,)), dtype=torch.float32).cuda()
  x_quantized = em.quantize(x)
  strings = em.compress(x)
  assert strings.shape == ()
  x_decompressed = em.decompress(strings, [100])
  assert torch.equal(x_decompressed, x_quantized)
  assert strings.tolist() == _oracle_strings(em, x)
  # the literal (unfused) op sequence produces the same bytes
  assert em.compress(x, fused=False).tolist() == strings.tolist()
  assert torch.equal(em.decompress(strings, [100], fused=False), x_quantized)


def test_quantizes_to_integers_modulo_offset_and_st_gradient(tfc):
  """continuous_batched_test.py:74-91."""
  noisy = tfc.NoisyNormal(loc=.25, scale=1.)
  em = tfc.ContinuousBatchedEntropyModel(noisy, 1, compression=True)
  x = torch.linspace(-20., 20., 1000).cuda().requires_grad_(True)
  xq = em.quantize(x)
  assert torch.allclose((xq.detach() - .25) - torch.round(xq.detach() - .25), torch.zeros_like(xq), atol=1e-6)
  xq.sum().backward()
  assert torch.equal(x.grad, torch.ones_like(x))


@pytest.mark.parametrize("scale", [2**-2, 2**0, 2**3, 2**7])
def test_information_bounds(tfc, scale):
  """continuous_batched_test.py:111-145: compressed bits exceed the model's own estimate by < 0.5 %...
  (1e6 samples in the reference; 2e5 here keeps the CPU oracle comparison fast).  Small scales exercise
  the Elias-gamma escape heavily."""
  noisy = tfc.NoisyNormal(loc=0., scale=float(scale))
  em = tfc.ContinuousBatchedEntropyModel(noisy, 1, compression=True)
  n = 200000
  x = torch.tensor(np.random.default_rng(1).normal(0., scale, size=(n,)), dtype=torch.float32).cuda()
  _, bits_eval = em(x, training=False)
  strings = em.compress(x)
  bits_compressed = 8 * strings.nbytes()
  assert strings.tolist() == _oracle_strings(em, x)
  assert bits_compressed > float(bits_eval) * 0.999
  assert bits_compressed < float(bits_eval) * 1.02 + 64
  assert torch.equal(em.decompress(strings, [n]), em.quantize(x))


def test_small_bitcost_for_dirac_prior(tfc):
  """continuous_batched_test.py:220-242: a near-deterministic prior -> tiny tables, <= 2 bytes/stream."""
  prior = tfc.NoisyNormal(loc=0., scale=1e-6)
  em = tfc.ContinuousBatchedEntropyModel(prior, 1, compression=True)
  assert em.cdf.numel() <= 16 * 6
  x = torch.zeros(1000).cuda()
  s = em.compress(x)
  assert s.nbytes() <= 2
  assert s.tolist() == _oracle_strings(em, x)
  assert torch.equal(em.decompress(s, [1000]), x)


def test_batched_channel_tables_bls2017_shape(tfc):
  """bls2017 usage: NoisyDeepFactorized(batch_shape=(C,)), coding_rank=3 (models/bls2017.py:103,160-161)."""
  torch.manual_seed(0)
  C = 8
  prior = tfc.NoisyDeepFactorized(batch_shape=(C,))
  em = tfc.ContinuousBatchedEntropyModel(prior, coding_rank=3, compression=True)
  assert em.range_coder_precision == 12
  y = (torch.randn(5, 6, 7, C) * 12).cuda()
  strings = em.compress(y)
  assert strings.shape == (5,)
  assert strings.tolist() == _oracle_strings(em, y)
  y_hat = em.decompress(strings, (6, 7))
  assert y_hat.shape == y.shape
  assert torch.equal(y_hat, em.quantize(y))
  _, bits = em(y, training=False)
  assert bits.shape == (5,)
  total = 8 * strings.nbytes()
  assert total > float(bits.sum()) * 0.98 and total < float(bits.sum()) * 1.05 + 5 * 32
  # EntropyBottleneck adaptor = the same thing
  eb = tfc.EntropyBottleneck(prior=prior)
  assert torch.equal(eb.cdf, em.cdf)


def test_state_dict_roundtrip_keeps_strings(tfc):
  """continuous_batched_test.py:147-173: tables and offset survive serialisation; strings identical."""
  torch.manual_seed(1)
  prior = tfc.NoisyDeepFactorized(batch_shape=(4,))
  em = tfc.ContinuousBatchedEntropyModel(prior, coding_rank=2, compression=True)
  y = (torch.randn(3, 10, 4) * 5).cuda()
  s1 = em.compress(y).tolist()
  cfg = em.get_config()
  em2 = tfc.ContinuousBatchedEntropyModel.from_config(cfg)
  em2.load_state_dict(em.state_dict())
  assert em2.compress(y).tolist() == s1
  assert torch.equal(em2.decompress(s1, (10,)), em.quantize(y))
  with pytest.raises(RuntimeError):
    tfc.ContinuousBatchedEntropyModel(prior, coding_rank=2, compression=False).compress(y)


def test_location_scale_indexed_model_bmshj2018_shape(tfc):
  """bmshj2018 usage: LocationScaleIndexedEntropyModel(NoisyNormal, 64 scales) (models/bmshj2018.py:240-245)
  and continuous_indexed_test.py:139-145 (decompress(compress(x)) == quantize(x))."""
  num_scales, smin, smax = 64, .11, 256.
  offset = np.log(smin)
  factor = (np.log(smax) - np.log(smin)) / (num_scales - 1.)
  scale_fn = lambda i: torch.exp(offset + factor * i)
  em = tfc.LocationScaleIndexedEntropyModel(tfc.NoisyNormal, num_scales, scale_fn, coding_rank=3, compression=True)
  rng = np.random.default_rng(5)
  B, H, W, C = 3, 4, 5, 6
  indexes = torch.tensor(rng.uniform(-2, 66, size=(B, H, W, C)), dtype=torch.float32).cuda()
  idx_int = torch.clamp(indexes, 0, num_scales - 1).to(torch.int32)
  sigma = scale_fn(idx_int.float().cpu())
  y = (torch.tensor(rng.normal(size=(B, H, W, C)), dtype=torch.float32) * sigma).cuda()
  loc = torch.tensor(rng.normal(size=(B, H, W, C)), dtype=torch.float32).cuda()
  for use_loc in (None, loc):
    strings = em.compress(y, indexes, loc=use_loc)
    assert strings.shape == (B,)
    xq = em.quantize(y, use_loc)
    assert torch.equal(em.decompress(strings, indexes, loc=use_loc), xq)
    assert em.compress(y, indexes, loc=use_loc, fused=False).tolist() == strings.tolist()
    shifted = y if use_loc is None else y - use_loc
    assert strings.tolist() == _oracle_strings(em, shifted, idx_int)


def test_sanity_check_raises_on_truncated_strings(tfc):
  em = tfc.ContinuousBatchedEntropyModel(tfc.NoisyNormal(loc=0., scale=5.), 1, compression=True)
  x = torch.randn(500).cuda() * 5
  s = em.compress(x).tolist()[0]
  with pytest.raises(tfc.InvalidArgumentError, match="Sanity check failed"):
    em.decompress([s + b"\x01\x02\x03\x04\x05\x06"], [500])


def test_gdn_layer_matches_reference_closed_form(tfc):
  """layers/gdn_test.py:42-66 through the layer class, both data formats, ranks 2..5."""
  for rank in (2, 3, 4, 5):
    for fmt in ("channels_last", "channels_first"):
      shape = (2,) + (3,) * (rank - 2) + (5,) if fmt == "channels_last" else (2, 5) + (3,) * (rank - 2)
      x = (torch.rand(shape) - .5).cuda()
      layer = tfc.GDN(data_format=fmt)
      y = layer(x)
      assert y.shape == x.shape
      assert torch.allclose(y, x / (1 + .1 * x.abs()), rtol=0, atol=1e-6)
      yi = tfc.GDN(inverse=True, data_format=fmt)(x)
      assert torch.allclose(yi, x * (1 + .1 * x.abs()), rtol=0, atol=1e-6)
  layer = tfc.GDN()
  x = torch.randn(4, 8, 8, 16).cuda().requires_grad_(True)
  layer(x).square().sum().backward()
  names = sorted(n for n, _ in layer.named_parameters())
  assert names == ["beta_parameter.variable", "gamma_parameter.variable"]
  assert all(p.grad is not None for p in layer.parameters()) and x.grad is not None
  with pytest.raises(RuntimeError):
    layer.inverse = True


def test_universal_models_compress_and_decompress_and_match_oracle(tfc):
  """universal_test.py:29-58 (batched) and :169-188 (indexed): decode(encode(x)) is x up to the shared uniform noise,
  string length matches the bit estimate, and the strings are the oracle's for the index-mode symbols the model
  derives (universal.py:213-251,530-566)."""
  from compression_b200 import entropy_models as E
  torch.manual_seed(0)
  prior = tfc.NoisyLogistic(loc=torch.zeros(6), scale=torch.linspace(1., 8., 6))
  em = tfc.UniversalBatchedEntropyModel(prior, coding_rank=2, compression=True, num_noise_levels=15)
  assert em.cdf_offset.numel() == 15 * 6
  x = (torch.randn(3, 4000, 6) * torch.linspace(1., 8., 6) * 1.8).cuda()
  strings = em.compress(x)
  assert strings.shape == (3,)
  x_hat = em.decompress(strings, (4000,))
  assert float((x_hat - x).abs().max()) <= 0.5 + 1e-5
  _, bits = em(x, training=False)
  got_bits = torch.tensor([8 * len(b) for b in strings.tolist()], dtype=torch.float32)
  assert torch.allclose(bits.cpu(), got_bits, rtol=0.01, atol=16)
  # literal op sequence == fused kernels; oracle bytes
  assert em.compress(x, fused=False).tolist() == strings.tolist()
  assert torch.equal(em.decompress(strings, (4000,), fused=False), x_hat)
  indexes, offset = em._compute_indexes_and_offset((4000,), x.device)
  sym = (torch.round(x - offset).to(torch.int32) - em.cdf_offset.cuda()[indexes.long()]).reshape(3, -1).cpu().numpy()
  idx = torch.broadcast_to(indexes, x.shape).reshape(3, -1).cpu().numpy().astype(np.int32)
  assert strings.tolist() == oracle.best().encode(em.cdf.cpu().numpy(), sym, idx)
  # indexed variant: one table per (noise level, scale index)
  emi = tfc.UniversalIndexedEntropyModel(tfc.NoisyLogistic, (8,), dict(loc=lambda i: 0. * i[..., 0], scale=lambda i: torch.exp(i[..., 0] / 3.)),
                                         coding_rank=1, compression=True, num_noise_levels=7)
  assert emi.cdf_offset.numel() == 7 * 8
  ind = torch.randint(0, 8, (5, 3000, 1)).float().cuda()
  xi = (torch.randn(5, 3000).cuda() * torch.exp(ind[..., 0] / 3.) * 1.8)
  si = emi.compress(xi, ind)
  xi_hat = emi.decompress(si, ind)
  assert si.shape == (5,) and float((xi_hat - xi).abs().max()) <= 0.5 + 1e-5
  assert emi.compress(xi, ind, fused=False).tolist() == si.tolist()
  flat, off = emi._coding_tensors(ind, xi.device)
  symi = (torch.round(xi - off).to(torch.int32) - emi.cdf_offset.cuda()[flat.long()]).cpu().numpy()
  assert si.tolist() == oracle.best().encode(emi.cdf.cpu().numpy(), symi, flat.cpu().numpy())
