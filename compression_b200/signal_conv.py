"""SignalConv1D / 2D / 3D: the thin convolution glue of the models' transforms, mirroring
tensorflow_compression/python/layers/signal_conv.py (:849-952 `call`, :697-776 the transposed-convolution path and
its crop, :470-560 parameters), `same_padding_for_kernel` (ops/padding_ops.py:22-51), `RDFTParameter`
(layers/parameters.py:70-180) and `IdentityInitializer` (layers/initializers.py:25-66).

This is GLUE, not a hand-written kernel: the convolutions run in cuDNN through torch.  What is reproduced exactly is
the layer's signal-processing contract: `corr` (cross-correlation vs. convolution), `strides_down`, `strides_up`
(zero-insertion upsampling, `extra_pad_end`), the `same_zeros` / `same_reflect` / `valid` alignment,
`channel_separable`, the (support..., in, out) kernel layout, the RDFT reparameterisation of the kernel and the bias /
activation order.  tests/test_signal_conv_cpu.py checks the combinations against a direct restatement of the
definition.
"""
import math

import torch
import torch.nn.functional as Fnn
from torch import nn

from compression_b200.parameters import Parameter

__all__ = ["same_padding_for_kernel", "RDFTParameter", "SignalConv1D", "SignalConv2D", "SignalConv3D",
           "IdentityInitializer"]


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


class RDFTParameter(Parameter):
  """Kernel stored as its real-input DFT over the spatial axes, split in real / imaginary parts and scaled by
  1/sqrt(prod(support)) (parameters.py:70-180).  `forward()` returns the kernel in (support..., in, out) layout."""

  def __init__(self, initial_value, name=None, shape=None, dtype=None):
    super().__init__()
    self.name = name
    if initial_value is None:
      if shape is None:
        raise ValueError("If initial_value is None, shape must be specified.")
      initial_value = torch.zeros(tuple(shape), dtype=dtype or torch.float32)
    initial_value = torch.as_tensor(initial_value, dtype=dtype)
    if initial_value.dim() not in (3, 4, 5):
      raise ValueError(f"Expected kernel tensor of rank 3, 4, or 5; received shape {tuple(initial_value.shape)}.")
    self._shape = tuple(initial_value.shape)
    r = len(self._shape) - 2
    self._to_spectral = (r, r + 1) + tuple(range(r))            # (in, out, support...)
    self._to_kernel = tuple(range(2, r + 2)) + (0, 1)
    self._dims = tuple(range(-r, 0))
    self._norm = math.sqrt(math.prod(self._shape[:r]))
    rdft = torch.fft.rfftn(initial_value.permute(self._to_spectral), dim=self._dims) / self._norm
    self.real = nn.Parameter(rdft.real.contiguous())
    self.imag = nn.Parameter(rdft.imag.contiguous())

  shape = property(lambda self: self._shape)

  def forward(self, compute_dtype=None):
    real, imag = self.real, self.imag
    if compute_dtype in (torch.bfloat16, torch.float16):
      real, imag = real.float(), imag.float()
    rdft = torch.complex(real, imag) * self._norm
    r = len(self._shape) - 2
    kernel = torch.fft.irfftn(rdft, s=self._shape[:r], dim=self._dims).permute(self._to_kernel)
    return kernel if compute_dtype is None else kernel.to(compute_dtype)

  def get_config(self):
    return dict(name=self.name, initial_value=None, shape=self._shape, dtype="float32")


_CONV = {1: Fnn.conv1d, 2: Fnn.conv2d, 3: Fnn.conv3d}
_CONV_TRANSPOSE = {1: Fnn.conv_transpose1d, 2: Fnn.conv_transpose2d, 3: Fnn.conv_transpose3d}


def _pad_spec(per_dim):
  """[(begin, end)] per spatial axis, outermost first -> torch.nn.functional.pad's innermost-first flat list."""
  spec = []
  for begin, end in reversed(list(per_dim)):
    spec += [int(begin), int(end)]
  return spec


class _SignalConv(nn.Module):
  """{1,2,3}-D convolution / correlation with optional down- and upsampling (signal_conv.py:131-470 constructor
  contract, :849-952 `call`).

  Input and output are channels-last `[B, spatial..., C]` by default (the models' layout; the GDN kernels consume it
  in place); `data_format="channels_first"` is accepted.  The layer builds lazily on the first call like the Keras
  layer does (`filters` outputs from however many input channels arrive; with `channel_separable`, `filters` outputs
  PER input channel, ordered channel-major as `tf.nn.depthwise_conv2d` orders them)."""
  _rank = None

  def __init__(self, filters, kernel_support, corr=False, strides_down=1, strides_up=1, padding="valid",
               extra_pad_end=True, channel_separable=False, data_format="channels_last", activation=None,
               use_bias=False, use_explicit=True, kernel_parameter="rdft", bias_parameter="variable",
               kernel_initializer="variance_scaling", bias_initializer="zeros", name=None):
    super().__init__()
    padding = str(padding).lower()
    if padding not in ("valid", "same_zeros", "same_reflect"):
      raise ValueError(f"Unsupported padding mode: '{padding}'")
    if data_format not in ("channels_first", "channels_last"):
      raise ValueError(f"Unknown data format: '{data_format}'.")
    if isinstance(kernel_parameter, str) and kernel_parameter not in ("variable", "rdft"):
      raise ValueError(f"Unsupported value for kernel_parameter: '{kernel_parameter}'.")
    self.filters = int(filters)
    self.kernel_support = self._tuple(kernel_support, "kernel_support")
    self.corr = bool(corr)
    self.strides_down = self._tuple(strides_down, "strides_down")
    self.strides_up = self._tuple(strides_up, "strides_up")
    self.padding = padding
    self.extra_pad_end = bool(extra_pad_end)
    self.channel_separable = bool(channel_separable)
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

  def _tuple(self, v, what):
    if isinstance(v, int):
      return self._rank * (int(v),)
    v = tuple(int(a) for a in v)
    if len(v) != self._rank:
      raise ValueError(f"`{what}` must be an integer or a sequence of {self._rank} integers, received {v}.")
    return v

  def build(self, in_channels, device=None):
    """signal_conv.py:600-640: creates kernel (support..., in, filters) and bias (output channels,)."""
    in_channels = int(in_channels)
    shape = self.kernel_support + (in_channels, self.filters)
    out_channels = self.filters * in_channels if self.channel_separable else self.filters
    if isinstance(self.kernel_parameter, str):
      init = self.kernel_initializer
      if callable(init):
        value = torch.as_tensor(init(shape), dtype=torch.float32)
      else:  # Keras "variance_scaling": truncated normal, variance 1 / fan_in
        fan_in = math.prod(self.kernel_support) * in_channels
        std = math.sqrt(1.0 / fan_in) / .87962566103423978
        value = torch.nn.init.trunc_normal_(torch.empty(shape), std=std, a=-2 * std, b=2 * std)
      if self.kernel_parameter == "rdft":
        self.kernel_parameter = RDFTParameter(value, name="kernel")
      else:
        self.kernel_parameter = nn.Parameter(value)
    if self.use_bias and isinstance(self.bias_parameter, str):
      init = self.bias_initializer
      value = torch.as_tensor(init((out_channels,)), dtype=torch.float32) if callable(init) else torch.zeros(out_channels)
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
    r = self._rank
    if inputs.dim() != r + 2:
      raise ValueError(f"Input tensor must have rank {r + 2}, received shape {tuple(inputs.shape)}.")
    x = inputs.permute(0, r + 1, *range(1, r + 1)) if self.data_format == "channels_last" else inputs   # N C spatial
    if not self.built:
      self.build(x.shape[1], device=inputs.device)
    kernel = self.kernel.to(x.dtype)
    corr = self.corr
    odd = all(s % 2 == 1 for s in self.kernel_support)
    up = any(s != 1 for s in self.strides_up)
    # the same kernel manipulations as the reference (:861-883), so that even-length kernels align identically
    if not corr and not up and odd:
      corr, kernel = True, kernel.flip(*range(r))
    elif corr and up and odd:
      corr, kernel = False, kernel.flip(*range(r))
    zero = r * ((0, 0),)
    prepad = zero
    if self.padding != "valid":
      pad = same_padding_for_kernel(self.kernel_support, corr, self.strides_up)
      if self.padding == "same_reflect":
        if any(p != (0, 0) for p in pad):
          x = Fnn.pad(x, _pad_spec(pad), mode="reflect")
        prepad = tuple(pad)
      elif corr and any(p != (0, 0) for p in pad):
        x = Fnn.pad(x, _pad_spec(pad))
      # same_zeros in the convolution branch: zeros are what the transposed convolution sees past the ends anyway
    if corr and not up:
      y = self._correlate_down(x, kernel)
    elif not corr:
      y = self._up_convolve(x, kernel, prepad)
    else:
      raise NotImplementedError("This layer does not support cross-correlation with upsampling of even-length kernels.")
    if self.use_bias:
      y = y + self.bias.to(y.dtype).reshape(1, -1, *(r * (1,)))
    if self.data_format == "channels_last":
      y = y.permute(0, *range(2, r + 2), 1)
    y = y.contiguous()
    if self.activation is not None:
      y = self.activation(y)
    return y

  def _correlate_down(self, x, kernel):
    """Correlate the (already padded) signal, then downsample: one strided cuDNN correlation (:600-660)."""
    r = self._rank
    if self.channel_separable:   # (support..., in, m) -> (in * m, 1, support...), one group per input channel
      w = kernel.permute(r, r + 1, *range(r)).reshape(-1, 1, *self.kernel_support)
      return _CONV[r](x, w, stride=self.strides_down, groups=x.shape[1])
    return _CONV[r](x, kernel.permute(r + 1, r, *range(r)), stride=self.strides_down)

  def _up_convolve(self, x, kernel, prepad):
    """Upsample by zero insertion, convolve, crop (signal_conv.py:697-776): conv_transpose IS that convolution,
    computed without the inserted zeros.  FULL[n] = sum_i x[i] k[n - i*s] has (L-1)*s + k samples, s - 1 more zeros
    with `extra_pad_end`; `valid` drops k - 1 samples from both ends, `same_*` drops prepad*s + k//2 from the start
    and prepad*s + (k-1)//2 from the end (prepad = what was reflect-padded before the call; 0 for zeros), leaving
    len*s samples (or len*s - (s-1) without `extra_pad_end`)."""
    r = self._rank
    w = kernel.permute(r, r + 1, *range(r))              # (in, out | m, support...): true convolution with `kernel`
    s, k = self.strides_up, self.kernel_support
    L = x.shape[2:]
    full = _CONV_TRANSPOSE[r](x, w, stride=s, groups=x.shape[1] if self.channel_separable else 1)
    sl = []
    extend = []
    for i in range(r):
      total = L[i] * s[i] + (k[i] - 1) if self.extra_pad_end else (L[i] - 1) * s[i] + k[i]
      if self.padding == "valid":
        start = stop = k[i] - 1
      else:
        start = prepad[i][0] * s[i] + k[i] // 2
        stop = prepad[i][1] * s[i] + (k[i] - 1) // 2
      # the zeros of `extra_pad_end` are materialised only where the crop reaches into them (short kernels)
      extend.append((0, max(0, total - stop - full.shape[2 + i])))
      sl.append(slice(start, total - stop, self.strides_down[i]))
    if any(e != (0, 0) for e in extend):
      full = Fnn.pad(full, _pad_spec(extend))
    return full[(slice(None), slice(None)) + tuple(sl)]

  def compute_output_shape(self, input_shape):
    r = self._rank
    input_shape = tuple(input_shape)
    spatial = input_shape[1:r + 1] if self.data_format == "channels_last" else input_shape[2:]
    channels = input_shape[-1] if self.data_format == "channels_last" else input_shape[1]
    out = []
    for i, n in enumerate(spatial):
      n = n * self.strides_up[i] if self.extra_pad_end else (n - 1) * self.strides_up[i] + 1
      if self.padding == "valid":
        n = n - (self.kernel_support[i] - 1)
      out.append(-(-n // self.strides_down[i]))
    c = self.filters * channels if self.channel_separable else self.filters
    if self.data_format == "channels_last":
      return (input_shape[0],) + tuple(out) + (c,)
    return (input_shape[0], c) + tuple(out)


class SignalConv1D(_SignalConv):
  """1-D layer (signal_conv.py:955-985)."""
  _rank = 1


class SignalConv2D(_SignalConv):
  """2-D layer (signal_conv.py:988-1016): what the models' transforms are made of."""
  _rank = 2


class SignalConv3D(_SignalConv):
  """3-D layer (signal_conv.py:1019-1047)."""
  _rank = 3
