"""GPU: the two models driven end to end (BASELINE.json configs[1]/[2] as stated): image -> analysis transform
(SignalConv2D glue + CUDA GDN) -> entropy model(s) on the CUDA range coder -> .tfci -> back to an image.
Reference: models/bls2017.py:55-190,262-321 and models/bmshj2018.py:53-264.  What is checked: the data path (shapes,
cropping, container layout), that the strings the model emits for ITS OWN latents are the oracle's bytes, and that
decompress reproduces exactly the synthesis of the quantised latents."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _image(h, w, seed):
  return torch.randint(0, 256, (h, w, 3), generator=torch.Generator().manual_seed(seed), dtype=torch.uint8)


def test_bls2017_image_to_tfci_and_back():
  from compression_b200 import models, PackedTensors
  torch.manual_seed(0)
  m = models.BLS2017Model(num_filters=32).build("cuda").fix_tables()
  for h, w in ((64, 64), (37, 50)):          # the second one needs the crop of bls2017.py:187
    x = _image(h, w, h)
    string, x_shape, y_shape = m.compress(x)
    assert string.shape == (1,) and x_shape.tolist() == [h, w] and y_shape.tolist() == [-(-h // 16), -(-w // 16)]
    # the bytes are the oracle's for the latents this model produced
    y = m.analysis_transform(x[None].cuda().float())
    em = m.entropy_model
    sym = (torch.round(y - em.quantization_offset).to(torch.int32) - em.cdf_offset).reshape(1, -1).cpu().numpy() \
        if em.quantization_offset is not None else (torch.round(y).to(torch.int32) - em.cdf_offset).reshape(1, -1).cpu().numpy()
    assert string.tolist() == oracle.best().encode(em.cdf.cpu().numpy(), sym)
    data = m.compress_to_tfci(x)
    packed = PackedTensors(data)
    s2, xs, ys = packed.unpack([bytes, torch.int32, torch.int32])     # the reference's feature order: string, x_shape, y_shape
    assert s2 == string.tolist() and xs.tolist() == [h, w] and ys.tolist() == y_shape.tolist()
    x_hat = m.decompress_from_tfci(data)
    assert x_hat.dtype == torch.uint8 and tuple(x_hat.shape) == (h, w, 3)
    want = m.synthesis_transform(em.quantize(y))[0, :h, :w]
    assert torch.equal(x_hat, torch.clamp(torch.round(want), 0, 255).to(torch.uint8))
  # training-time graph of the same model: rate-distortion terms and gradients into every variable
  loss, bpp, mse = m(torch.rand(2, 32, 32, 3).cuda() * 255, training=True)
  loss.backward()
  assert float(bpp) > 0 and float(mse) > 0
  assert all(p.grad is not None for n, p in m.named_parameters() if not n.startswith("entropy_model."))


def test_bls2017_batch_compress_codes_one_stream_per_image():
  from compression_b200 import models
  torch.manual_seed(1)
  m = models.BLS2017Model(num_filters=16).build("cuda").fix_tables()
  x = torch.stack([_image(48, 80, s) for s in range(5)])
  strings, x_shape, y_shape = m.compress_batch(x)
  assert strings.shape == (5,) and y_shape.tolist() == [3, 5]
  x_hat = m.decompress_batch(strings, x_shape, y_shape)
  assert tuple(x_hat.shape) == (5, 48, 80, 3) and x_hat.dtype == torch.uint8
  y = m.analysis_transform(x.cuda().float())
  want = m.synthesis_transform(m.entropy_model.quantize(y))[:, :48, :80]
  assert torch.equal(x_hat, torch.clamp(torch.round(want), 0, 255).to(torch.uint8))


def test_bmshj2018_two_level_image_to_tfci_and_back():
  from compression_b200 import models, PackedTensors
  torch.manual_seed(2)
  m = models.BMSHJ2018Model(num_filters=24).build("cuda", patch=(64, 64)).fix_tables()
  for h, w in ((64, 64), (70, 45)):
    x = _image(h, w, w)
    string, side_string, x_shape, y_shape, z_shape = m.compress(x)
    assert x_shape.tolist() == [h, w] and y_shape.tolist() == [-(-h // 16), -(-w // 16)]
    assert z_shape.tolist() == [-(-int(y_shape[0]) // 4), -(-int(y_shape[1]) // 4)]
    data = m.compress_to_tfci(x)
    feats = PackedTensors(data).unpack([bytes, bytes, torch.int32, torch.int32, torch.int32])
    assert feats[0] == string.tolist() and feats[1] == side_string.tolist() and feats[4].tolist() == z_shape.tolist()
    x_hat = m.decompress_from_tfci(data)
    assert x_hat.dtype == torch.uint8 and tuple(x_hat.shape) == (h, w, 3)
    # both levels against the oracle, on the latents this model produced
    y = m.analysis_transform(x[None].cuda().float())
    z = m.hyper_analysis_transform(y.abs())
    sem, em = m.side_entropy_model, m.entropy_model
    q = sem.quantization_offset
    zs = (torch.round(z if q is None else z - q).to(torch.int32) - sem.cdf_offset).reshape(1, -1).cpu().numpy()
    assert side_string.tolist() == oracle.best().encode(sem.cdf.cpu().numpy(), zs)
    idx = m.hyper_synthesis_transform(sem.quantize(z))[:, :y.shape[1], :y.shape[2], :]
    flat = torch.clamp(idx, 0, m.num_scales - 1).to(torch.int32)
    ysym = (torch.round(y).to(torch.int32) - em.cdf_offset.cuda()[flat.long()]).reshape(1, -1).cpu().numpy()
    assert string.tolist() == oracle.best().encode(em.cdf.cpu().numpy(), ysym, flat.reshape(1, -1).cpu().numpy())
    want = m.synthesis_transform(em.quantize(y))[0, :h, :w]
    assert torch.equal(x_hat, torch.clamp(torch.round(want), 0, 255).to(torch.uint8))
  loss, bpp, mse = m(torch.rand(2, 64, 64, 3).cuda() * 255, training=True)
  loss.backward()
  assert float(bpp) > 0 and torch.isfinite(loss)


def test_ms2020_slice_loop_image_to_tfci_and_back():
  """models/ms2020.py:331-433: the hyperprior string plus one string per channel slice, each slice coded in index mode
  with `loc` conditioned on the slices decoded before it.  Checked: the container layout (three shapes, then
  num_slices + 1 strings), every slice string against the oracle for the symbols this model derives, decompress ==
  the synthesis of the latents the encoder itself reconstructed, batches, and the training graph."""
  from compression_b200 import models, PackedTensors
  torch.manual_seed(4)
  S = 4
  m = models.MS2020Model(num_filters=24, latent_depth=32, hyperprior_depth=16, num_slices=S, max_support_slices=2)
  m.build("cuda", patch=(64, 64)).fix_tables()
  # (ms2020.py concatenates the un-cropped hyper-synthesis output with the slices: image sides must be multiples of 64,
  # which is what models/tfci.py pads to)
  for h, w in ((64, 64), (128, 64)):
    x = _image(h, w, h + w)
    out = m.compress(x)
    x_shape, y_shape, z_shape, z_string = out[:4]
    y_strings = out[4:]
    assert len(y_strings) == S and x_shape.tolist() == [h, w] and y_shape.tolist() == [-(-h // 16), -(-w // 16)]
    data = m.compress_to_tfci(x)
    feats = PackedTensors(data).unpack([torch.int32] * 3 + [bytes] * (S + 1))
    assert feats[0].tolist() == [h, w] and feats[3] == z_string.tolist() and feats[4 + S - 1] == y_strings[-1].tolist()
    x_hat = m.decompress_from_tfci(data)
    assert x_hat.dtype == torch.uint8 and tuple(x_hat.shape) == (h, w, 3)
    # replay the encoder's slice loop and compare every string with the oracle
    y = m.analysis_transform(x[None].cuda().float())
    z = m.hyper_analysis_transform(y)
    zs = (torch.round(z).to(torch.int32) - m.em_z.cdf_offset).reshape(1, -1).cpu().numpy()   # offset_heuristic=False
    assert z_string.tolist() == oracle.best().encode(m.em_z.cdf.cpu().numpy(), zs)
    z_hat = m.em_z.quantize(z)
    lm, ls = m.hyper_synthesis_mean_transform(z_hat), m.hyper_synthesis_scale_transform(z_hat)
    y_hat_slices = []
    for i, y_slice in enumerate(torch.chunk(y, S, dim=-1)):
      mu, sigma, support = m._slice_params(i, lm, ls, y_hat_slices, tuple(y.shape[1:-1]))
      flat = torch.clamp(sigma, 0, m.num_scales - 1).to(torch.int32)
      sym = (torch.round(y_slice - mu).to(torch.int32) - m.em_y.cdf_offset.cuda()[flat.long()]).reshape(1, -1).cpu().numpy()
      assert y_strings[i].tolist() == oracle.best().encode(m.em_y.cdf.cpu().numpy(), sym, flat.reshape(1, -1).cpu().numpy())
      y_hat_slices.append(m._lrp(i, support, m.em_y.quantize(y_slice, loc=mu)))
    want = m.synthesis_transform(torch.cat(y_hat_slices, dim=-1))[0, :h, :w]
    assert torch.equal(x_hat, torch.clamp(torch.round(want), 0, 255).to(torch.uint8))
  xb = torch.stack([_image(64, 128, s) for s in range(3)])
  outb = m.compress_batch(xb)
  assert outb[3].shape == (3,) and all(s.shape == (3,) for s in outb[4:])
  assert tuple(m.decompress_batch(*outb).shape) == (3, 64, 128, 3)
  loss, bpp, mse = m(torch.rand(2, 64, 64, 3).cuda() * 255, training=True)
  loss.backward()
  assert float(bpp) > 0 and torch.isfinite(loss)
