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
