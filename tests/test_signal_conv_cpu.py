"""CPU: SignalConv2D (cuDNN/torch glue around the hot path) against a direct restatement of the layer's definition
(tensorflow_compression/python/layers/signal_conv.py:40-128 docstring: zero-insertion upsampling, `same_zeros` alignment
of ops/padding_ops.py:22-51, convolution vs. cross-correlation, downsampling), the RDFT kernel parameterisation
(layers/parameters.py:70-180) and `same_padding_for_kernel` itself (padding_ops_test.py)."""
import itertools

import numpy as np
import pytest
import torch

from compression_b200.signal_conv import RDFTParameter, SignalConv2D, same_padding_for_kernel


def _definition(x, kernel, corr, up, down, padding, extra_pad_end=True):
  """x [B,H,W,Ci], kernel [kh,kw,Ci,Co] -> [B,H',W',Co] in float64, by loops over the definition."""
  B, H, W, Ci = x.shape
  kh, kw, _, Co = kernel.shape
  Hu = H * up if extra_pad_end else (H - 1) * up + 1
  Wu = W * up if extra_pad_end else (W - 1) * up + 1
  u = np.zeros((B, Hu, Wu, Ci))
  u[:, ::up, ::up][:, :H, :W] = x
  if padding == "same_zeros":
    ph = (kh // 2, (kh - 1) // 2) if corr else ((kh - 1) // 2, kh // 2)
    pw = (kw // 2, (kw - 1) // 2) if corr else ((kw - 1) // 2, kw // 2)
    u = np.pad(u, ((0, 0), ph, pw, (0, 0)))
  k = kernel if corr else kernel[::-1, ::-1]
  Ho, Wo = u.shape[1] - kh + 1, u.shape[2] - kw + 1
  out = np.zeros((B, Ho, Wo, Co))
  for i in range(kh):
    for j in range(kw):
      out += np.einsum("bhwc,cd->bhwd", u[:, i:i + Ho, j:j + Wo], k[i, j])
  return out[:, ::down, ::down]


@pytest.mark.parametrize("corr,up,down,k,padding", [
    c for c in itertools.product([True, False], [1, 2, 4], [1, 2], [(3, 3), (5, 5), (9, 9), (4, 4), (3, 5)],
                                 ["same_zeros", "valid"])
    if not (c[0] and c[1] > 1 and any(s % 2 == 0 for s in c[3]))])   # the reference raises for that one too
def test_matches_the_definition(corr, up, down, k, padding):
  rng = np.random.default_rng(hash((corr, up, down, k, padding)) % 2**32)
  H, W = (7, 6) if padding == "same_zeros" else (11, 12)
  x = rng.normal(size=(2, H, W, 3))
  kernel = rng.normal(size=k + (3, 4))
  layer = SignalConv2D(4, k, corr=corr, strides_up=up, strides_down=down, padding=padding, use_bias=True,
                       kernel_parameter=torch.tensor(kernel), bias_parameter=torch.tensor([.5, -1., 0., 2.], dtype=torch.float64))
  y = layer(torch.tensor(x))
  want = _definition(x, kernel, corr, up, down, padding) + np.asarray([.5, -1., 0., 2.])
  assert tuple(y.shape) == want.shape == layer.compute_output_shape(x.shape)
  np.testing.assert_allclose(y.numpy(), want, rtol=1e-10, atol=1e-10)
  # channels_first is the same computation on the transposed tensor
  cf = SignalConv2D(4, k, corr=corr, strides_up=up, strides_down=down, padding=padding, data_format="channels_first",
                    kernel_parameter=torch.tensor(kernel))
  np.testing.assert_allclose(cf(torch.tensor(x).permute(0, 3, 1, 2)).permute(0, 2, 3, 1).numpy(),
                             want - np.asarray([.5, -1., 0., 2.]), rtol=1e-10, atol=1e-10)


def test_same_padding_for_kernel_cases():
  """padding_ops.py:22-51 by hand: kernel 5 / 4, correlation / convolution, upsampling by 2."""
  assert same_padding_for_kernel((5, 4), True) == [(2, 2), (2, 1)]
  assert same_padding_for_kernel((5, 4), False) == [(2, 2), (1, 2)]
  assert same_padding_for_kernel((9, 5), True, (4, 2)) == [(1, 1), (1, 1)]
  assert same_padding_for_kernel((3,), False, (1,)) == [(1, 1)]
  assert same_padding_for_kernel((1,), True) == [(0, 0)]


def test_rdft_parameter_round_trips_the_kernel_and_trains():
  torch.manual_seed(0)
  k = torch.randn(5, 5, 3, 8)
  p = RDFTParameter(k)
  assert p.real.shape == (3, 8, 5, 3) and p.shape == (5, 5, 3, 8)          # rfft2 over (kh, kw): kw -> kw // 2 + 1
  torch.testing.assert_close(p(), k, rtol=1e-5, atol=1e-5)
  # Parseval scaling of parameters.py:113-116: unit-variance kernel -> O(1) spectrum, whatever the support
  assert .3 < float(p.real.std()) < 3.
  p().square().sum().backward()
  assert p.real.grad is not None and p.imag.grad is not None


def test_lazy_build_default_parameters_and_activation_order():
  torch.manual_seed(1)
  layer = SignalConv2D(6, (5, 5), corr=True, strides_down=2, padding="same_zeros", use_bias=True,
                       activation=torch.relu)
  with pytest.raises(RuntimeError):
    layer.kernel  # pylint:disable=pointless-statement
  y = layer(torch.randn(2, 9, 8, 3))
  assert y.shape == (2, 5, 4, 6) and float(y.min()) >= 0
  names = sorted(n for n, _ in layer.named_parameters())
  assert names == ["bias_parameter", "kernel_parameter.imag", "kernel_parameter.real"]     # rdft kernel by default
  var = float(layer.kernel.var())
  assert .5 / 75 < var < 2. / 75                                                           # variance_scaling: 1 / fan_in
  with pytest.raises(ValueError):
    layer(torch.randn(9, 8, 3))


def test_identity_initializer_kernels():
  """initializers_test.py:22-50."""
  from compression_b200.signal_conv import IdentityInitializer, SignalConv2D
  k = IdentityInitializer(gain=3)((3, 4, 3), dtype=torch.int32)
  want = torch.zeros(3, 4, 3, dtype=torch.int32)
  for i in range(3):
    want[1, i, i] = 3
  assert torch.equal(k, want)
  k = IdentityInitializer()((4, 5, 1, 1))
  want = torch.zeros(4, 5, 1, 1)
  want[2, 2, 0, 0] = 1
  assert torch.equal(k, want) and k.dtype == torch.float32
  with pytest.raises(ValueError):
    IdentityInitializer()((2, 3))
  # what it is for: a `same`-padded layer initialised with it copies its input
  layer = SignalConv2D(3, (5, 5), corr=True, padding="same_zeros", use_bias=False, kernel_parameter="variable",
                       kernel_initializer=IdentityInitializer())
  x = torch.randn(2, 9, 8, 3)
  assert torch.allclose(layer(x), x, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# 1-D / 3-D layers, `same_reflect`, `channel_separable`: against the layer's definition computed with scipy
# ------------------------------------------------------------------------------------------------
def _definition_nd(x, kernel, corr, up, down, padding, extra_pad_end=True, separable=False):
  """x [B, *L, Ci], kernel [*k, Ci, Co | m] -> [B, *L', Co | Ci*m] (signal_conv.py:61-215 docstring): upsample by
  zero insertion, convolve / correlate with the kernel centred at K // 2, downsample.  `same_*` padding is applied to
  the signal BEFORE the zero insertion (that is what makes `same_reflect` well defined, :884-905), in the amounts of
  ops/padding_ops.py:22-51."""
  from scipy import signal
  r = x.ndim - 2
  k = kernel.shape[:r]
  mode = {"same_zeros": "constant", "same_reflect": "reflect", "valid": None}[padding]
  if corr and any(u > 1 for u in up):          # correlation == convolution with the reversed kernel (odd supports)
    assert all(s % 2 == 1 for s in k)
    kernel = np.flip(kernel, axis=tuple(range(r)))
    corr = False
  B, Ci = x.shape[0], x.shape[-1]
  nout = kernel.shape[-1]
  pairs = [(c, j, c * nout + j) for c in range(Ci) for j in range(nout)] if separable else \
      [(c, j, j) for c in range(Ci) for j in range(nout)]
  n_out_channels = Ci * nout if separable else nout
  if corr:
    xp = x
    if mode:
      xp = np.pad(x, [(0, 0)] + [(s // 2, (s - 1) // 2) for s in k] + [(0, 0)], mode=mode)
    out = None
    for c, j, o in pairs:
      for b in range(B):
        y = signal.correlate(xp[b, ..., c], kernel[..., c, j], mode="valid")
        if out is None:
          out = np.zeros((B,) + y.shape + (n_out_channels,))
        out[b, ..., o] += y
  else:
    P = r * [(0, 0)]
    if mode:
      P = [(((s - 1) // 2 - 1) // u + 1, (s // 2 - 1) // u + 1) for s, u in zip(k, up)]
    xp = np.pad(x, [(0, 0)] + P + [(0, 0)], mode=mode) if mode else x
    Lp = xp.shape[1:-1]
    Lu = [n * u if extra_pad_end else (n - 1) * u + 1 for n, u in zip(Lp, up)]
    us = np.zeros((B,) + tuple(Lu) + (Ci,))
    us[(slice(None),) + tuple(slice(0, None, u) for u in up)] = xp
    crop = []
    for i in range(r):
      if padding == "valid":
        a = z = k[i] - 1
      else:
        a, z = P[i][0] * up[i] + k[i] // 2, P[i][1] * up[i] + (k[i] - 1) // 2
      crop.append(slice(a, Lu[i] + k[i] - 1 - z))
    out = None
    for c, j, o in pairs:
      for b in range(B):
        y = signal.convolve(us[b, ..., c], kernel[..., c, j], mode="full")[tuple(crop)]
        if out is None:
          out = np.zeros((B,) + y.shape + (n_out_channels,))
        out[b, ..., o] += y
  return out[(slice(None),) + tuple(slice(None, None, d) for d in down)]


def _layer_cls(rank):
  from compression_b200 import signal_conv
  return {1: signal_conv.SignalConv1D, 2: signal_conv.SignalConv2D, 3: signal_conv.SignalConv3D}[rank]


ND_CASES = [
    # rank, support, corr, up, down, padding, extra_pad_end, separable
    (1, (5,), True, 1, 2, "same_zeros", True, False),
    (1, (4,), False, 2, 1, "same_zeros", True, False),
    (1, (3,), False, 2, 1, "valid", False, False),
    (1, (5,), True, 1, 2, "same_reflect", True, False),
    (1, (5,), False, 2, 1, "same_reflect", False, False),
    (1, (4,), False, 3, 2, "same_reflect", True, False),
    (1, (5,), True, 1, 1, "same_zeros", True, True),
    (1, (3,), False, 2, 1, "same_zeros", True, True),
    (2, (5, 3), True, 1, 2, "same_reflect", True, False),
    (2, (3, 5), False, (2, 1), 1, "same_reflect", True, False),
    (2, (5, 5), True, 2, 1, "same_reflect", False, False),
    (2, (4, 3), False, 1, 1, "same_reflect", True, False),
    (2, (5, 5), True, 1, (2, 2), "same_zeros", True, True),
    (2, (3, 3), False, 2, 1, "valid", True, True),
    (2, (3, 4), False, 1, 1, "same_reflect", True, True),
    (3, (3, 3, 3), True, 1, 2, "same_zeros", True, False),
    (3, (3, 2, 3), False, 2, 1, "same_zeros", True, False),
    (3, (3, 3, 3), True, 1, 1, "valid", True, False),
    (3, (3, 3, 3), False, (1, 2, 2), (2, 1, 1), "same_reflect", True, False),
]


@pytest.mark.parametrize("rank,k,corr,up,down,padding,extra,separable", ND_CASES)
def test_nd_layers_match_the_definition(rank, k, corr, up, down, padding, extra, separable):
  rng = np.random.default_rng(hash((rank, k, corr, str(up), str(down), padding, extra, separable)) % 2**32)
  up_t = rank * (up,) if isinstance(up, int) else up
  down_t = rank * (down,) if isinstance(down, int) else down
  L = {1: (11,), 2: (9, 8), 3: (6, 5, 7)}[rank]
  Ci, F = 3, 2
  x = rng.normal(size=(2,) + L + (Ci,))
  kernel = rng.normal(size=k + (Ci, F))
  layer = _layer_cls(rank)(F, k, corr=corr, strides_up=up, strides_down=down, padding=padding, extra_pad_end=extra,
                           channel_separable=separable, kernel_parameter=torch.tensor(kernel))
  y = layer(torch.tensor(x))
  want = _definition_nd(x, kernel, corr, up_t, down_t, padding, extra, separable)
  assert tuple(y.shape) == want.shape == layer.compute_output_shape(x.shape)
  np.testing.assert_allclose(y.numpy(), want, rtol=1e-9, atol=1e-9)
  cf = _layer_cls(rank)(F, k, corr=corr, strides_up=up, strides_down=down, padding=padding, extra_pad_end=extra,
                        channel_separable=separable, data_format="channels_first", kernel_parameter=torch.tensor(kernel))
  yc = cf(torch.tensor(x).movedim(-1, 1))
  assert tuple(yc.shape) == cf.compute_output_shape((2, Ci) + L)
  np.testing.assert_allclose(yc.movedim(1, -1).numpy(), want, rtol=1e-9, atol=1e-9)


def test_definition_agrees_with_the_2d_loop_restatement():
  """The scipy definition used for the N-D cases is the same function as the loop restatement above."""
  rng = np.random.default_rng(0)
  x, kernel = rng.normal(size=(2, 7, 6, 3)), rng.normal(size=(4, 5, 3, 4))
  for corr, up, down, padding in ((False, 2, 1, "same_zeros"), (True, 1, 2, "same_zeros"), (False, 1, 1, "valid"),
                                  (False, 4, 2, "valid")):
    np.testing.assert_allclose(_definition_nd(x, kernel, corr, (up, up), (down, down), padding),
                               _definition(x, kernel, corr, up, down, padding), rtol=1e-10, atol=1e-10)


def test_equivalent_kernels_of_the_docstring_give_the_same_output():
  """signal_conv.py:92-104: with `same_*` padding, convolving with [1,2,3], [0,1,2,3,0], [0,1,2,3] and correlating
  with [3,2,1], [0,3,2,1,0], [0,3,2,1] are all the same operation (the kernel centre is at K // 2)."""
  from compression_b200.signal_conv import SignalConv1D
  x = torch.randn(2, 12, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
  outs = []
  for corr, taps in ((False, [1, 2, 3]), (False, [0, 1, 2, 3, 0]), (False, [0, 1, 2, 3]),
                     (True, [3, 2, 1]), (True, [0, 3, 2, 1, 0]), (True, [0, 3, 2, 1])):
    for padding in ("same_zeros", "same_reflect"):
      k = torch.tensor(taps, dtype=torch.float64).reshape(-1, 1, 1)
      outs.append((padding, SignalConv1D(1, len(taps), corr=corr, padding=padding, kernel_parameter=k)(x)))
  for padding in ("same_zeros", "same_reflect"):
    same = [o for p, o in outs if p == padding]
    for o in same[1:]:
      torch.testing.assert_close(o, same[0], rtol=1e-12, atol=1e-12)


def test_rdft_parameter_for_1d_and_3d_kernels():
  for shape in ((7, 3, 4), (3, 4, 5, 2, 3)):
    k = torch.randn(shape, generator=torch.Generator().manual_seed(len(shape)))
    p = RDFTParameter(k)
    assert p.shape == shape and p.real.shape[:2] == shape[-2:] and p.real.shape[-1] == shape[len(shape) - 3] // 2 + 1
    torch.testing.assert_close(p(), k, rtol=1e-5, atol=1e-5)
  with pytest.raises(ValueError):
    RDFTParameter(torch.zeros(3, 3))


def test_argument_errors():
  from compression_b200.signal_conv import SignalConv1D, SignalConv3D
  with pytest.raises(ValueError):
    SignalConv2D(4, (3, 3), padding="same_wrap")
  with pytest.raises(ValueError):
    SignalConv3D(4, (3, 3))
  with pytest.raises(ValueError):
    SignalConv1D(4, 3)(torch.zeros(2, 5, 5, 3))
  with pytest.raises(NotImplementedError):   # correlation + upsampling of an even-length kernel (:946-947)
    SignalConv1D(2, 4, corr=True, strides_up=2, kernel_parameter="variable")(torch.zeros(1, 8, 3))
