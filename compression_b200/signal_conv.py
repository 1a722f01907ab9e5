"""SignalConv2D: the thin convolution glue of the models' transforms, mirroring the subset of
tensorflow_compression/python/layers/signal_conv.py the two models use (:849-952 `call`, :778-847 the
transposed-convolution path, :470-560 parameters), `same_padding_for_kernel` (ops/padding_ops.py:22-51) and
`RDFTParameter` (layers/parameters.py:70-180).

This is GLUE, not a hand-written kernel: the convolutions run in cuDNN through torch.  What is reproduced exactly is
the layer's signal-processing contract: `corr` (cross-correlation vs. convolution), `strides_down`, `strides_up`
(zero-insertion upsampling, `extra_pad_end`), the `same_zeros` / `valid` alignment, the (kh, kw, in, out) kernel
layout, the RDFT reparameterisation of the kernel and the bias / activation order.  tests/test_signal_conv_cpu.py
checks every combination against a direct restatement of the definition.
"""
import math

import torch
import torch.nn.functional as Fnn
from torch import nn

__all__ = ["same_padding_for_kernel", "RDFTParameter", "SignalConv2D", "IdentityInitializer"]


class IdentityInitializer:
  """Initialises an n-D `SignalConv*` kernel (spatial..., in, out) to the identity (initializers.py:25-66): `gain`
  at the centre tap `s // 2` of every spatial axis, times the in x out identity matrix."""

  def __init__(self, gain=1):
    self.gain = gain

  def __call__(self, shape, dtype=None):
    shape = tuple(int(d) for d in shape)
    if len(shape) <= 2:
      raise ValueError(f"shape must be at least rank 3, got {shape}.")
    dtype = torch.float32 if dtype is None else dtype
    kernel = torch.zeros(shape, dtype=dtype)
    centre = tuple(s // 2 for s in shape[:-2])
    kernel[centre] = self.gain * torch.eye(shape[-2], shape[-1], dtype=dtype)
    return kernel

  def get_config(self):
    return dict(gain=self.gain)


def same_padding_for_kernel(shape, corr, strides_up=None):
  """padding_ops.py:22-51: (begin, end) padding per dimension so that a `valid` convolution / correlation of the
  padded (and possibly upsampled) signal has the size of the unpadded one."""
  rank = len(shape)
  if strides_up is None:
    strides_up = rank * (1,)
  if corr:
    padding = [(s // 2, (s - 1) // 2) for s in shape]
  else:
    padding = [((s - 1) // 2, s // 2) for s in shape]
  return [((padding[i][0] - 1) // strides_up[i] + 1, (padding[i][1] - 1) // strides_up[i] + 1) for i in range(rank)]


class RDFTParameter(nn.Module):
  """Kernel stored as its real-input DFT, split in real / imaginary parts and scaled by 1/sqrt(kh*kw)
  (parameters.py:70-180).  `forward()` returns the kernel in (kh, kw, in, out) layout."""

  def __init__(self, initial_value, name=None, shape=None, dtype=None):
    super().__init__()
    self.name = name
    if initial_value is None:
      if shape is None:
        raise ValueError("If initial_value is None, shape must be specified.")
      initial_value = torch.zeros(tuple(shape), dtype=dtype or torch.float32)
    initial_value = torch.as_tensor(initial_value, dtype=dtype)
    if initial_value.dim() != 4:
      raise ValueError(f"Expected kernel tensor of rank 4; received shape {tuple(initial_value.shape)}.")
    self._shape = tuple(initial_value.shape)
    rdft = torch.fft.rfft2(initial_value.permute(2, 3, 0, 1))       # transform over (kh, kw)
    rdft = rdft / math.sqrt(self._shape[0] * self._shape[1])
    self.real = nn.Parameter(rdft.real.contiguous())
    self.imag = nn.Parameter(rdft.imag.contiguous())

  shape = property(lambda self: self._shape)

  def forward(self, compute_dtype=None):
    real, imag = self.real, self.imag
    if compute_dtype in (torch.bfloat16, torch.float16):
      real, imag = real.float(), imag.float()
    rdft = torch.complex(real, imag) * math.sqrt(self._shape[0] * self._shape[1])
    kernel = torch.fft.irfft2(rdft, s=self._shape[:2]).permute(2, 3, 0, 1)
    return kernel if compute_dtype is None else kernel.to(compute_dtype)

  def get_config(self):
    return dict(name=self.name, initial_value=None, shape=self._shape, dtype="float32")


def _pair(v):
  return (int(v), int(v)) if isinstance(v, int) else tuple(int(a) for a in v)


class SignalConv2D(nn.Module):
  """2-D convolution / correlation with optional down- and upsampling (signal_conv.py:131-470 constructor
  contract; only what `call` needs is kept: no channel_separable, padding in {"valid", "same_zeros"}).

  Input and output are channels-last `[B, H, W, C]` by default (the models' layout; the GDN kernels consume it in
  place); `data_format="channels_first"` is accepted.  The layer builds lazily on the first call like the Keras
  layer does (`filters` outputs from however many input channels arrive)."""

  def __init__(self, filters, kernel_support, corr=False, strides_down=1, strides_up=1, padding="valid",
               extra_pad_end=True, channel_separable=False, data_format="channels_last", activation=None,
               use_bias=False, use_explicit=True, kernel_parameter="rdft", bias_parameter="variable",
               kernel_initializer="variance_scaling", bias_initializer="zeros", name=None):
    super().__init__()
    if channel_separable:
      raise NotImplementedError("channel_separable convolutions are not used by the models on this path.")
    if padding not in ("valid", "same_zeros"):
      raise NotImplementedError(f"padding='{padding}' is not supported here (the models use 'same_zeros').")
    if data_format not in ("channels_first", "channels_last"):
      raise ValueError(f"Unknown data format: '{data_format}'.")
    if isinstance(kernel_parameter, str) and kernel_parameter not in ("variable", "rdft"):
      raise ValueError(f"Unsupported value for kernel_parameter: '{kernel_parameter}'.")
    self.filters = int(filters)
    self.kernel_support = _pair(kernel_support)
    self.corr = bool(corr)
    self.strides_down = _pair(strides_down)
    self.strides_up = _pair(strides_up)
    self.padding = padding
    self.extra_pad_end = bool(extra_pad_end)
    self.data_format = data_format
    self.activation = activation
    self.use_bias = bool(use_bias)
    self.use_explicit = bool(use_explicit)
    self.kernel_parameter = kernel_parameter
    self.bias_parameter = bias_parameter if self.use_bias else None
    self.kernel_initializer = kernel_initializer
    self.bias_initializer = bias_initializer
    self.name = name
    self.built = False

  def build(self, in_channels, device=None):
    """signal_conv.py:600-640: creates kernel (kh, kw, in, out) and bias (out,)."""
    kh, kw = self.kernel_support
    shape = (kh, kw, int(in_channels), self.filters)
    if isinstance(self.kernel_parameter, str):
      init = self.kernel_initializer
      if callable(init):
        value = torch.as_tensor(init(shape), dtype=torch.float32)
      else:  # Keras "variance_scaling": truncated normal, variance 1 / fan_in
        fan_in = kh * kw * int(in_channels)
        value = torch.nn.init.trunc_normal_(torch.empty(shape), std=math.sqrt(1.0 / fan_in) / .87962566103423978,
                                            a=-2 * math.sqrt(1.0 / fan_in) / .87962566103423978,
                                            b=2 * math.sqrt(1.0 / fan_in) / .87962566103423978)
      if self.kernel_parameter == "rdft":
        self.kernel_parameter = RDFTParameter(value, name="kernel")
      else:
        self.kernel_parameter = nn.Parameter(value)
    if self.use_bias and isinstance(self.bias_parameter, str):
      init = self.bias_initializer
      value = torch.as_tensor(init((self.filters,)), dtype=torch.float32) if callable(init) else torch.zeros(self.filters)
      self.bias_parameter = nn.Parameter(value)
    if device is not None:
      self.to(device)
    self.built = True

  @property
  def kernel(self):
    if isinstance(self.kernel_parameter, str):
      raise RuntimeError("Kernel is not initialized yet. Call build().")
    p = self.kernel_parameter
    return p() if isinstance(p, nn.Module) else p

  @property
  def bias(self):
    if isinstance(self.bias_parameter, str):
      raise RuntimeError("Bias is not initialized yet. Call build().")
    return self.bias_parameter

  def forward(self, inputs):
    """signal_conv.py:849-952."""
    if inputs.dim() != 4:
      raise ValueError(f"Input tensor must have rank 4, received shape {tuple(inputs.shape)}.")
    x = inputs.permute(0, 3, 1, 2) if self.data_format == "channels_last" else inputs   # NCHW view
    if not self.built:
      self.build(x.shape[1], device=inputs.device)
    kernel = self.kernel.to(x.dtype)
    corr = self.corr
    odd = all(s % 2 == 1 for s in self.kernel_support)
    up = any(s != 1 for s in self.strides_up)
    # the same kernel manipulations as the reference (:861-883), so that even-length kernels align identically
    if not corr and not up and odd:
      corr, kernel = True, kernel.flip(0, 1)
    elif corr and up and odd:
      corr, kernel = False, kernel.flip(0, 1)
    if self.padding == "valid":
      pad = ((0, 0), (0, 0))
    else:
      pad = same_padding_for_kernel(self.kernel_support, corr, self.strides_up)
    if corr and not up:
      # correlate, then downsample: one strided cuDNN correlation on the zero-padded input
      if any(p != (0, 0) for p in pad):
        x = Fnn.pad(x, (pad[1][0], pad[1][1], pad[0][0], pad[0][1]))
      y = Fnn.conv2d(x, kernel.permute(3, 2, 0, 1), stride=self.strides_down)
    elif not corr:
      y = self._up_convolve(x, kernel, pad)
    else:
      raise NotImplementedError("This layer does not support cross-correlation with upsampling of even-length kernels.")
    if self.use_bias:
      y = y + self.bias.to(y.dtype).reshape(1, -1, 1, 1)
    if self.data_format == "channels_last":
      y = y.permute(0, 2, 3, 1)
    y = y.contiguous()
    if self.activation is not None:
      y = self.activation(y)
    return y

  def _up_convolve(self, x, kernel, prepad):
    """Upsample by zero insertion, convolve, crop (signal_conv.py:778-847): conv_transpose IS that convolution,
    computed without the inserted zeros.  FULL[n] = sum_i x[i] k[n - i*s]; the `same` output starts at
    prepad*s + k//2 of the pre-padded signal's full convolution, i.e. at k//2 of the unpadded one, and is
    len*s long (`extra_pad_end`) or len*s - (s-1)."""
    w = kernel.permute(2, 3, 0, 1)                       # (in, out, kh, kw): true convolution with `kernel`
    s, k = self.strides_up, self.kernel_support
    L = x.shape[2:]
    full = Fnn.conv_transpose2d(x, w, stride=s)          # [(L-1)*s + k] per dimension
    sl = []
    for i in range(2):
      n_up = L[i] * s[i] if self.extra_pad_end else (L[i] - 1) * s[i] + 1
      if self.padding == "valid":
        start, length = k[i] - 1, n_up - (k[i] - 1)
      else:
        start, length = k[i] // 2, n_up
      short = start + length - full.shape[2 + i]
      if short > 0:                                       # positions past the data: zeros of the end padding
        padspec = [0, 0, 0, 0]
        padspec[2 * (1 - i) + 1] = short
        full = Fnn.pad(full, padspec)
      sl.append(slice(start, start + length, self.strides_down[i]))
    return full[:, :, sl[0], sl[1]]

  def compute_output_shape(self, input_shape):
    b, h, w, c = input_shape if self.data_format == "channels_last" else (input_shape[0], *input_shape[2:], input_shape[1])
    out = []
    for i, n in enumerate((h, w)):
      n = n * self.strides_up[i] if self.extra_pad_end else (n - 1) * self.strides_up[i] + 1
      if self.padding == "valid":
        n = n - (self.kernel_support[i] - 1)
      out.append(-(-n // self.strides_down[i]))
    return (b, out[0], out[1], self.filters) if self.data_format == "channels_last" else (b, self.filters, out[0], out[1])
