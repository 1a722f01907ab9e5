"""Entropy models of the reference on the B200 range coder.

Mirrors tensorflow_compression/python/entropy_models:
  continuous_base.py:36-370     ContinuousEntropyModelBase (table build, storage, config)
  continuous_batched.py:30-436  ContinuousBatchedEntropyModel
  continuous_indexed.py:30-633  ContinuousIndexedEntropyModel, LocationScaleIndexedEntropyModel
Same constructor keywords, properties and method semantics; tensors are CUDA torch tensors,
`tf.string` results are `gen_ops.Strings`.  `EntropyBottleneck` (the TFC-1.x name used by the task
statement) is provided as a thin alias, see the bottom of the file.
"""
import functools
import math
import warnings

import numpy as np
import torch
from torch import nn

from compression_b200 import distributions as D
from compression_b200 import functional as F
from compression_b200 import gen_ops, math_ops

__all__ = [
    "ContinuousEntropyModelBase", "ContinuousBatchedEntropyModel", "ContinuousIndexedEntropyModel",
    "LocationScaleIndexedEntropyModel", "UniversalBatchedEntropyModel", "UniversalIndexedEntropyModel",
    "EntropyBottleneck",
]


def _cuda():
  return torch.device("cuda", torch.cuda.current_device())


class ContinuousEntropyModelBase(nn.Module):
  """continuous_base.py:36-370."""

  def __init__(self, coding_rank=None, compression=False, stateless=False, expected_grads=False,
               tail_mass=2**-8, bottleneck_dtype=None, laplace_tail_mass=0):
    super().__init__()
    self._prior = None  # set by the subclasses; an nn.Module prior registers as a submodule (see _set_prior)
    self._coding_rank = int(coding_rank)
    self._compression = bool(compression)
    self._stateless = bool(stateless)
    self._expected_grads = bool(expected_grads)
    self._tail_mass = float(tail_mass)
    self._bottleneck_dtype = bottleneck_dtype or torch.float32
    self._laplace_tail_mass = laplace_tail_mass
    if self.coding_rank < 0:
      raise ValueError("`coding_rank` must be at least 0.")
    if not 0 < self.tail_mass < 1:
      raise ValueError("`tail_mass` must be between 0 and 1.")

  def _check_compression(self):
    if not self.compression:
      raise RuntimeError(
          "For range coding, the entropy model must be instantiated with `compression=True`.")

  @property
  def prior(self):
    if self._prior is None:
      raise RuntimeError(
          "This entropy model doesn't hold a reference to its prior distribution. This can happen "
          "depending on how it is instantiated, (e.g., if it is unserialized).")
    return self._prior

  @prior.deleter
  def prior(self):
    self._prior = None

  def _set_prior(self, prior, register=True):
    """A prior the caller hands in is part of the model exactly as in the reference, where assigning it on the
    tf.Module makes its variables trainable_variables / checkpoint state (continuous_batched.py:205): an
    nn.Module prior becomes a registered submodule (parameters(), state_dict(), .to()).  Priors the model
    derives itself from `indexes` (continuous_indexed.py:226-232) hold no variables and stay plain attributes."""
    if register or not isinstance(prior, nn.Module):
      self._prior = prior
    else:
      self._modules.pop("_prior", None)
      object.__setattr__(self, "_prior", prior)

  @property
  def cdf(self):
    self._check_compression()
    return self._cdf

  @property
  def cdf_offset(self):
    self._check_compression()
    return self._cdf_offset

  bottleneck_dtype = property(lambda self: self._bottleneck_dtype)
  expected_grads = property(lambda self: self._expected_grads)
  laplace_tail_mass = property(lambda self: self._laplace_tail_mass)
  coding_rank = property(lambda self: self._coding_rank)
  compression = property(lambda self: self._compression)
  stateless = property(lambda self: self._stateless)
  tail_mass = property(lambda self: self._tail_mass)

  @property
  def range_coder_precision(self):
    return -int(self.cdf[0])

  def _init_compression(self, cdf, cdf_offset, cdf_shapes):
    """continuous_base.py:167-215: tables are stored (buffers), never rebuilt on the receiving side."""
    if not ((cdf is None) == (cdf_offset is None) == (cdf_shapes is not None)):
      raise ValueError("Either both `cdf` and `cdf_offset`, or `cdf_shapes` must be provided.")
    if cdf_shapes is not None:
      if self.stateless:
        raise ValueError("With `stateless=True`, can't provide `cdf_shapes`.")
      cdf_shapes = tuple(map(int, cdf_shapes))
      if len(cdf_shapes) != 2:
        raise ValueError("`cdf_shapes` must have two elements.")
      cdf = torch.zeros(cdf_shapes[0], dtype=torch.int32)
      cdf_offset = torch.zeros(cdf_shapes[1], dtype=torch.int32)
    cdf = torch.as_tensor(cdf).to(torch.int32)
    cdf_offset = torch.as_tensor(cdf_offset).to(torch.int32)
    if self.stateless:
      self._cdf, self._cdf_offset = cdf, cdf_offset
    else:
      self.register_buffer("_cdf", cdf)
      self.register_buffer("_cdf_offset", cdf_offset)
    self._cdf_host = None

  def _lookup_host(self):
    """Host copy of the table for handle creation (cached; tables are immutable once built)."""
    if self._cdf_host is None or self._cdf_host[0] is not self._cdf:
      self._cdf_host = (self._cdf, np.ascontiguousarray(self._cdf.detach().cpu().numpy(), dtype=np.int32))
    return self._cdf_host[1]

  def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
    # table buffers are restored with whatever shape was saved (validate_shape=False in the reference)
    for name in ("_cdf", "_cdf_offset", "_quantization_offset"):
      key = prefix + name
      if key in state_dict and getattr(self, name, None) is not None:
        setattr(self, name, state_dict[key].clone())
    self._cdf_host = None
    super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
    if self._prior is None:
      # a model rebuilt from its config holds tables, not a prior (continuous_base.py:336-360): prior weights
      # saved next to the tables are ignored instead of being reported as unexpected keys
      unexpected = args[3] if len(args) > 3 else kwargs.get("unexpected_keys")
      if unexpected is not None:
        unexpected[:] = [k for k in unexpected if not k.startswith(prefix + "_prior.")]

  @torch.no_grad()
  def _build_tables(self, prior, precision, offset=None):
    """continuous_base.py:217-296.  Returns (cdf 1-D int32 [-p, cdf...]*, cdf_offset int32)."""
    precision = int(precision)
    dev = _cuda()
    dtype = prior.dtype
    offset = torch.zeros((), dtype=dtype) if offset is None else torch.as_tensor(offset, dtype=dtype)
    lower = D.lower_tail(prior, self.tail_mass).to("cpu")
    upper = D.upper_tail(prior, self.tail_mass).to("cpu")
    offset = offset.to("cpu")
    minima = torch.floor(lower - offset).to(torch.int32)
    maxima = torch.ceil(upper - offset).to(torch.int32)
    pmf_start = minima.to(dtype) + offset
    pmf_length = maxima - minima + 1
    max_length = int(pmf_length.max())
    if max_length > 2048:
      warnings.warn(f"Very wide PMF with {max_length} elements may lead to out of memory issues. Consider "
                    "priors with smaller variance, or increasing `tail_mass` parameter.")
    prior_dev = getattr(prior, "device", torch.device("cpu"))
    samples = torch.arange(max_length, dtype=dtype).reshape([-1] + pmf_length.dim() * [1]) + pmf_start
    pmf = prior.prob(samples.to(prior_dev))
    pmf_shape = tuple(pmf.shape[1:])
    num_pmfs = int(np.prod(pmf_shape)) if pmf_shape else 1
    pmf = pmf.reshape(max_length, num_pmfs).t().contiguous()
    pmf_length = torch.broadcast_to(pmf_length, pmf_shape).reshape(num_pmfs)
    cdf_offset = torch.broadcast_to(minima, pmf_shape).reshape(num_pmfs)
    cdf = F.build_lookup(pmf.to(dev, torch.float32), pmf_length, precision)
    return cdf, cdf_offset.to(dev)

  def _log_prob(self, prior, bottleneck_perturbed):
    """continuous_base.py:298-334."""
    x = bottleneck_perturbed.to(prior.dtype)
    ltm = float(self.laplace_tail_mass)
    if ltm > 0:
      if not ltm < 1:
        raise ValueError("`laplace_tail_mass` must be less than 1.")
      lap = D.NoisyLaplace(loc=torch.zeros((), device=x.device), scale=torch.ones((), device=x.device))
      probs = (1 - ltm) * prior.prob(x) + ltm * lap.prob(x)
      too_small = probs < 1e-10
      return torch.where(too_small, math.log(ltm) + lap.log_prob(x), torch.log(torch.clamp(probs, min=1e-10)))
    return prior.log_prob(x)

  def get_config(self):
    """continuous_base.py:336-360."""
    if self.stateless or not self.compression:
      raise RuntimeError(
          "Serializing entropy models with `compression=False` or `stateless=True` is not supported.")
    return dict(
        coding_rank=self.coding_rank,
        compression=True,
        stateless=False,
        expected_grads=self.expected_grads,
        tail_mass=self.tail_mass,
        cdf_shapes=(int(self.cdf.shape[0]), int(self.cdf_offset.shape[0])),
        bottleneck_dtype=str(self.bottleneck_dtype).replace("torch.", ""),
        laplace_tail_mass=float(self.laplace_tail_mass),
    )

  def get_weights(self):
    return [b.detach().cpu().numpy() for _, b in self.named_buffers()]

  def set_weights(self, weights):
    names = [n for n, _ in self.named_buffers()]
    if len(weights) != len(names):
      raise ValueError(f"`set_weights` expects a list of {len(names)} arrays, received {len(weights)}.")
    for n, w in zip(names, weights):
      old = getattr(self, n)
      setattr(self, n, torch.as_tensor(w).to(device=old.device, dtype=old.dtype))
    self._cdf_host = None


class ContinuousBatchedEntropyModel(ContinuousEntropyModelBase):
  """continuous_batched.py:30-436: one table per element of `prior.batch_shape` (channel mode)."""

  def __init__(self, prior=None, coding_rank=None, compression=False, stateless=False, expected_grads=False,
               tail_mass=2**-8, range_coder_precision=12, bottleneck_dtype=None, prior_shape=None, cdf=None,
               cdf_offset=None, cdf_shapes=None, offset_heuristic=True, quantization_offset=None,
               decode_sanity_check=True, laplace_tail_mass=0):
    if (prior is None) == (prior_shape is None):
      raise ValueError("Either `prior` or `prior_shape` must be provided.")
    if (prior is None) + (cdf_shapes is None) + (cdf is None) != 2:
      raise ValueError("Must provide exactly one of `prior`, `cdf`, or `cdf_shapes`.")
    if not compression and not (cdf is None and cdf_offset is None and cdf_shapes is None):
      raise ValueError("CDFs can't be provided with `compression=False`")
    super().__init__(coding_rank=coding_rank, compression=compression, stateless=stateless,
                     expected_grads=expected_grads, tail_mass=tail_mass, bottleneck_dtype=bottleneck_dtype,
                     laplace_tail_mass=laplace_tail_mass)
    self._set_prior(prior)
    self._offset_heuristic = bool(offset_heuristic)
    self._prior_shape = tuple(int(s) for s in (prior_shape if prior is None else prior.batch_shape))
    if self.coding_rank < len(self.prior_shape):
      raise ValueError("`coding_rank` can't be smaller than `prior_shape`.")
    self.decode_sanity_check = decode_sanity_check

    if cdf_shapes is not None:
      assert isinstance(quantization_offset, bool)
      assert self.compression
      quantization_offset = torch.zeros(self.prior_shape) if quantization_offset else None
    elif quantization_offset is not None:
      pass
    elif self.offset_heuristic and self.compression:
      if self._prior is None:
        raise ValueError("To use the offset heuristic, a `prior` needs to be provided.")
      quantization_offset = D.quantization_offset(self.prior)
      if bool(torch.all(quantization_offset == 0.)):
        quantization_offset = None
      else:
        quantization_offset = torch.broadcast_to(quantization_offset, self.prior_shape).clone()
    else:
      quantization_offset = None
    if quantization_offset is None:
      self._quantization_offset = None
    else:
      q = torch.as_tensor(quantization_offset).detach().to(self.bottleneck_dtype)
      if self.compression and not self.stateless:
        self.register_buffer("_quantization_offset", q)
      else:
        self._quantization_offset = q
    if self.compression:
      if cdf is None and cdf_shapes is None:
        cdf, cdf_offset = self._build_tables(self.prior, range_coder_precision, offset=quantization_offset)
      self._init_compression(cdf, cdf_offset, cdf_shapes)

  prior_shape = property(lambda self: self._prior_shape)
  offset_heuristic = property(lambda self: self._offset_heuristic)

  @property
  def prior_shape_tensor(self):
    return torch.tensor(self.prior_shape, dtype=torch.int32)

  @property
  def quantization_offset(self):
    """continuous_batched.py:272-289."""
    if self._quantization_offset is not None:
      return self._quantization_offset
    if self.offset_heuristic and not self.compression:
      if self._prior is None:
        raise RuntimeError("To use the offset heuristic, a `prior` needs to be provided.")
      return D.quantization_offset(self.prior).to(self.bottleneck_dtype)
    return None

  def forward(self, bottleneck, training=True):
    """continuous_batched.py:291-322 -> (bottleneck_perturbed, bits)."""
    bottleneck = torch.as_tensor(bottleneck).to(self.bottleneck_dtype)
    log_prob_fn = functools.partial(self._log_prob, self.prior)
    if training:
      log_probs, perturbed = math_ops.perturb_and_apply(log_prob_fn, bottleneck,
                                                        expected_grads=self.expected_grads)
    else:
      perturbed = self.quantize(bottleneck)
      log_probs = log_prob_fn(perturbed)
    axes = tuple(range(-self.coding_rank, 0))
    bits = (log_probs.sum(dim=axes) if axes else log_probs) / -math.log(2.)
    return perturbed, bits

  def quantize(self, bottleneck):
    """continuous_batched.py:324-345."""
    bottleneck = torch.as_tensor(bottleneck).to(self.bottleneck_dtype)
    off = self.quantization_offset
    return math_ops.round_st(bottleneck, None if off is None else off.to(bottleneck.device))

  def _flat_tables(self, device):
    coff = self.cdf_offset.to(device).reshape(-1)
    qoff = self.quantization_offset
    return coff, (None if qoff is None else qoff.to(device, torch.float32).reshape(-1))

  def compress(self, bottleneck, fused=True):
    """continuous_batched.py:347-383.  `fused=True` quantises inside the encode kernel (same arithmetic);
    `fused=False` issues the reference's op sequence literally."""
    self._check_compression()
    bottleneck = torch.as_tensor(bottleneck).to(device=_cuda(), dtype=self.bottleneck_dtype)
    shape = tuple(bottleneck.shape)
    if len(shape) < self.coding_rank:
      raise ValueError("`bottleneck` has fewer dimensions than `coding_rank`.")
    batch_shape = shape[:len(shape) - self.coding_rank]
    rank_p = len(self.prior_shape)
    if rank_p and shape[-rank_p:] != self.prior_shape:
      bottleneck = torch.broadcast_to(bottleneck, shape[:-rank_p] + self.prior_shape)
    coff, qoff = self._flat_tables(bottleneck.device)
    handle = gen_ops.create_range_encoder(batch_shape, self._lookup_host())
    if fused and bottleneck.dtype == torch.float32:
      F.encode_channel_f32(handle, bottleneck.contiguous(), qoff, coff)
    else:
      b = bottleneck.to(torch.float32)
      if qoff is not None:
        b = b - qoff.reshape(self.prior_shape)
      symbols = torch.round(b).to(torch.int32)
      iid = shape[:len(shape) - rank_p] if rank_p else shape
      symbols = symbols.reshape(iid + (-1,)) - coff
      gen_ops.entropy_encode_channel(handle, symbols)
    return gen_ops.entropy_encode_finalize(handle)

  def decompress(self, strings, broadcast_shape, fused=True):
    """continuous_batched.py:385-422."""
    self._check_compression()
    if not isinstance(strings, gen_ops.Strings):
      strings = gen_ops.Strings.from_bytes(strings)
    broadcast_shape = tuple(int(d) for d in np.asarray(broadcast_shape).reshape(-1))
    n_prior = int(np.prod(self.prior_shape)) if self.prior_shape else 1
    output_shape = tuple(strings.shape) + broadcast_shape + self.prior_shape
    dev = strings.bytes_dev.device
    coff, qoff = self._flat_tables(dev)
    handle = gen_ops.create_range_decoder(strings, self._lookup_host())
    if fused and self.bottleneck_dtype == torch.float32:
      outputs = F.decode_channel_f32(handle, output_shape, qoff, coff)
      sanity = gen_ops.entropy_decode_finalize(handle)
      if self.decode_sanity_check and not bool(sanity.all()):
        raise gen_ops.InvalidArgumentError("Sanity check failed.")
      return outputs
    handle, symbols = gen_ops.entropy_decode_channel(handle, broadcast_shape + (n_prior,))
    sanity = gen_ops.entropy_decode_finalize(handle)
    if self.decode_sanity_check and not bool(sanity.all()):
      raise gen_ops.InvalidArgumentError("Sanity check failed.")
    symbols = symbols + coff
    outputs = symbols.reshape(output_shape).to(self.bottleneck_dtype)
    if qoff is not None:
      outputs = outputs + qoff.reshape(self.prior_shape).to(outputs.dtype)
    return outputs

  def get_config(self):
    """continuous_batched.py:424-436."""
    config = super().get_config()
    config.update(prior_shape=tuple(map(int, self.prior_shape)), offset_heuristic=self.offset_heuristic,
                  quantization_offset=self.quantization_offset is not None)
    return config

  @classmethod
  def from_config(cls, config):
    config = dict(config)
    dt = config.pop("bottleneck_dtype", "float32")
    return cls(bottleneck_dtype=getattr(torch, dt) if isinstance(dt, str) else dt, **config)


class ContinuousIndexedEntropyModel(ContinuousEntropyModelBase):
  """continuous_indexed.py:30-428: the table of every element is selected by an index tensor."""

  def __init__(self, prior_fn, index_ranges, parameter_fns, coding_rank, channel_axis=-1, compression=False,
               stateless=False, expected_grads=False, tail_mass=2**-8, range_coder_precision=12,
               bottleneck_dtype=None, prior_dtype=torch.float32, decode_sanity_check=True, laplace_tail_mass=0):
    if not callable(prior_fn):
      raise TypeError("`prior_fn` must be a class or factory function.")
    for name, fn in parameter_fns.items():
      if not isinstance(name, str):
        raise TypeError("`parameter_fns` must have string keys.")
      if not callable(fn):
        raise TypeError(f"`parameter_fns['{name}']` must be callable.")
    super().__init__(coding_rank=coding_rank, compression=compression, stateless=stateless,
                     expected_grads=expected_grads, tail_mass=tail_mass, bottleneck_dtype=bottleneck_dtype,
                     laplace_tail_mass=laplace_tail_mass)
    self._index_ranges = tuple(int(r) for r in index_ranges)
    if not self.index_ranges:
      raise ValueError("`index_ranges` must have at least one element.")
    self._channel_axis = None if channel_axis is None else int(channel_axis)
    if self.channel_axis is None and len(self.index_ranges) > 1:
      raise ValueError("`channel_axis` can't be `None` for `len(index_ranges) > 1`.")
    self._prior_fn = prior_fn
    self._parameter_fns = dict(parameter_fns)
    self._prior_dtype = prior_dtype
    self.decode_sanity_check = decode_sanity_check
    if self.compression:
      if self.channel_axis is None:
        indexes = torch.arange(self.index_ranges[0], dtype=torch.int32)
      else:
        grids = torch.meshgrid(*[torch.arange(r, dtype=torch.int32) for r in self.index_ranges], indexing="ij")
        indexes = torch.stack(grids, dim=self.channel_axis)
      self._set_prior(self._make_prior(indexes), register=False)
      cdf, cdf_offset = self._build_tables(self.prior, range_coder_precision)
      self._init_compression(cdf, cdf_offset, None)

  index_ranges = property(lambda self: self._index_ranges)
  parameter_fns = property(lambda self: self._parameter_fns)
  prior_dtype = property(lambda self: self._prior_dtype)
  prior_fn = property(lambda self: self._prior_fn)
  channel_axis = property(lambda self: self._channel_axis)

  def _make_prior(self, indexes):
    indexes = indexes.to(self.prior_dtype)
    parameters = {k: f(indexes) for k, f in self.parameter_fns.items()}
    prior = self.prior_fn(**parameters)
    assert prior.dtype == self.prior_dtype
    return prior

  def _normalize_indexes(self, indexes):
    """continuous_indexed.py:272-281."""
    indexes = math_ops.lower_bound(indexes, 0.)
    if self.channel_axis is None:
      bounds = torch.tensor(self.index_ranges[0] - 1, dtype=indexes.dtype, device=indexes.device)
    else:
      axes = [1] * indexes.dim()
      axes[self.channel_axis] = len(self.index_ranges)
      bounds = torch.tensor([s - 1 for s in self.index_ranges], dtype=indexes.dtype,
                            device=indexes.device).reshape(axes)
    return math_ops.upper_bound(indexes, bounds)

  def _flatten_indexes(self, indexes):
    """continuous_indexed.py:283-289."""
    indexes = indexes.to(torch.int32)
    if self.channel_axis is None:
      return indexes
    strides = np.cumprod((self.index_ranges + (1,))[::-1])[::-1][1:]
    strides = torch.tensor(strides.copy(), dtype=torch.int32, device=indexes.device)
    return (indexes.movedim(self.channel_axis, -1) * strides).sum(dim=-1, dtype=torch.int32)  # integer matmul does not exist on CUDA

  def forward(self, bottleneck, indexes, training=True):
    """continuous_indexed.py:291-334."""
    bottleneck = torch.as_tensor(bottleneck).to(self.bottleneck_dtype)
    indexes = self._normalize_indexes(torch.as_tensor(indexes, dtype=self.prior_dtype, device=bottleneck.device))
    if training:
      def log_prob_fn(x, idx):
        return self._log_prob(self._make_prior(idx), x)
      log_probs, perturbed = math_ops.perturb_and_apply(log_prob_fn, bottleneck, indexes,
                                                        expected_grads=self.expected_grads)
    else:
      prior = self._make_prior(indexes)
      perturbed = self.quantize(bottleneck)
      log_probs = self._log_prob(prior, perturbed)
    axes = tuple(range(-self.coding_rank, 0))
    bits = (log_probs.sum(dim=axes) if axes else log_probs) / -math.log(2.)
    return perturbed, bits

  def quantize(self, bottleneck):
    return math_ops.round_st(torch.as_tensor(bottleneck).to(self.bottleneck_dtype))

  def compress(self, bottleneck, indexes, fused=True, _loc=None):
    """continuous_indexed.py:354-386."""
    self._check_compression()
    dev = _cuda()
    bottleneck = torch.as_tensor(bottleneck).to(device=dev, dtype=self.bottleneck_dtype)
    indexes = self._normalize_indexes(torch.as_tensor(indexes).to(device=dev, dtype=self.prior_dtype))
    flat = self._flatten_indexes(indexes)
    fshape = tuple(flat.shape)
    batch_shape = fshape[:len(fshape) - self.coding_rank]
    coff = self.cdf_offset.to(dev)
    handle = gen_ops.create_range_encoder(batch_shape, self._lookup_host())
    if fused and bottleneck.dtype == torch.float32:
      F.encode_index_f32(handle, flat, bottleneck.contiguous(), _loc, coff)
    else:
      b = bottleneck if _loc is None else bottleneck - _loc
      symbols = torch.round(b).to(torch.int32) - coff[flat.long()]
      gen_ops.entropy_encode_index(handle, flat, symbols)
    return gen_ops.entropy_encode_finalize(handle)

  def decompress(self, strings, indexes, fused=True, _loc=None):
    """continuous_indexed.py:388-417."""
    self._check_compression()
    if not isinstance(strings, gen_ops.Strings):
      strings = gen_ops.Strings.from_bytes(strings)
    dev = strings.bytes_dev.device
    indexes = self._normalize_indexes(torch.as_tensor(indexes).to(device=dev, dtype=self.prior_dtype))
    flat = self._flatten_indexes(indexes)
    fshape = tuple(flat.shape)
    decode_shape = fshape[len(fshape) - self.coding_rank:] if self.coding_rank else ()
    coff = self.cdf_offset.to(dev)
    handle = gen_ops.create_range_decoder(strings, self._lookup_host())
    if fused and self.bottleneck_dtype == torch.float32:
      out = F.decode_index_f32(handle, flat, _loc, coff)
      symbols = None
    else:
      handle, symbols = gen_ops.entropy_decode_index(handle, flat, decode_shape)
    sanity = gen_ops.entropy_decode_finalize(handle)
    if self.decode_sanity_check and not bool(sanity.all()):
      raise gen_ops.InvalidArgumentError("Sanity check failed.")
    if symbols is None:
      return out
    out = (symbols + coff[flat.long()]).to(self.bottleneck_dtype)
    return out if _loc is None else out + _loc

  def get_config(self):
    raise NotImplementedError("Serializing indexed entropy models is not yet implemented.")

  @classmethod
  def from_config(cls, config):
    raise NotImplementedError("Serializing indexed entropy models is not yet implemented.")


class LocationScaleIndexedEntropyModel(ContinuousIndexedEntropyModel):
  """continuous_indexed.py:431-633."""

  def __init__(self, prior_fn, num_scales, scale_fn, coding_rank, compression=False, stateless=False,
               expected_grads=False, tail_mass=2**-8, range_coder_precision=12, bottleneck_dtype=None,
               prior_dtype=torch.float32, laplace_tail_mass=0):
    num_scales = int(num_scales)
    super().__init__(prior_fn=prior_fn, index_ranges=(num_scales,),
                     parameter_fns=dict(loc=lambda _: 0., scale=scale_fn), coding_rank=coding_rank,
                     channel_axis=None, compression=compression, stateless=stateless,
                     expected_grads=expected_grads, tail_mass=tail_mass,
                     range_coder_precision=range_coder_precision, bottleneck_dtype=bottleneck_dtype,
                     prior_dtype=prior_dtype, laplace_tail_mass=laplace_tail_mass)

  def forward(self, bottleneck, scale_indexes, loc=None, training=True):
    if loc is None:
      return super().forward(bottleneck, scale_indexes, training=training)
    perturbed, bits = super().forward(bottleneck - loc, scale_indexes, training=training)
    return perturbed + loc, bits

  def quantize(self, bottleneck, loc=None):
    return math_ops.round_st(torch.as_tensor(bottleneck).to(self.bottleneck_dtype), loc)

  def compress(self, bottleneck, scale_indexes, loc=None, fused=True):
    return super().compress(bottleneck, scale_indexes, fused=fused, _loc=loc)

  def decompress(self, strings, scale_indexes, loc=None, fused=True):
    return super().decompress(strings, scale_indexes, fused=fused, _loc=loc)


# ------------------------------------------------------------------------------------------------
# Universal quantisation (universal.py:30-603): "quantisation" is additive uniform noise whose value is shared by
# sender and receiver through a stateless pseudo-random stream; coding runs in index mode with one extra leading
# index, the noise level.
# ------------------------------------------------------------------------------------------------
_M32 = 0xFFFFFFFF


def _philox4x32(counter, key, rounds=10):
  """Philox-4x32-10 (Salmon et al., SC'11) on int64 tensors holding uint32 words: counter [n, 4] -> [n, 4]."""
  c = [counter[:, i].clone() for i in range(4)]
  k0, k1 = int(key[0]) & _M32, int(key[1]) & _M32
  for _ in range(rounds):
    p0, p1 = c[0] * 0xD2511F53, c[2] * 0xCD9E8D57          # < 2^64 as unsigned: split the products by hand
    hi0 = ((c[0] >> 16) * 0xD2511F53 + (((c[0] & 0xFFFF) * 0xD2511F53) >> 16)) >> 16
    hi1 = ((c[2] >> 16) * 0xCD9E8D57 + (((c[2] & 0xFFFF) * 0xCD9E8D57) >> 16)) >> 16
    lo0, lo1 = p0 & _M32, p1 & _M32
    c = [(hi1 ^ c[1] ^ k0) & _M32, lo1, (hi0 ^ c[3] ^ k1) & _M32, lo0]
    k0, k1 = (k0 + 0x9E3779B9) & _M32, (k1 + 0xBB67AE85) & _M32
  return torch.stack(c, dim=1)


def stateless_uniform_int(shape, seed, maxval, device=None):
  """Counter-based stand-in for `tf.random.stateless_uniform(shape, seed, 0, maxval, int32)` (universal.py:33-39):
  element i is word i % 4 of Philox-4x32-10(counter = i // 4, key = seed), reduced modulo `maxval`.  It depends on
  nothing but (seed, i), so sender and receiver -- on any device -- draw the same noise levels.  TensorFlow's own
  key / counter conventions are not reproduced (there is no TF here to pin them against): strings written with
  universal quantisation decode with THIS implementation, not with the reference's."""
  shape = tuple(int(d) for d in shape)
  n = 1
  for d in shape:
    n *= d
  blocks = (n + 3) // 4
  counter = torch.zeros(blocks, 4, dtype=torch.int64, device=device)
  idx = torch.arange(blocks, dtype=torch.int64, device=device)
  counter[:, 0] = idx & _M32
  counter[:, 1] = idx >> 32
  words = _philox4x32(counter, seed).reshape(-1)[:n]
  return (words % int(maxval)).to(torch.int32).reshape(shape)


def _add_offset_indexes(indexes, num_noise_levels):
  """universal.py:30-42: prepends the shared pseudo-random noise-level index to the last axis of `indexes`."""
  offset_indexes = stateless_uniform_int(indexes.shape[:-1], (1234, 1234), num_noise_levels, indexes.device)
  return torch.cat((offset_indexes.to(indexes.dtype)[..., None], indexes), dim=-1)


def _offset_indexes_to_offset(offset_indexes, num_noise_levels, dtype):
  """universal.py:45-47: level k of n -> (k + 1) / (n + 1) - 0.5."""
  return ((offset_indexes.to(torch.float64) + 1) / (num_noise_levels + 1) - 0.5).to(dtype)


def _range_coding_offsets(num_noise_levels, prior_rank, dtype):
  """universal.py:55-62."""
  offset_indexes = torch.arange(num_noise_levels, dtype=dtype).reshape([-1] + [1] * prior_rank)
  return _offset_indexes_to_offset(offset_indexes, num_noise_levels, dtype)


class UniversalBatchedEntropyModel(ContinuousEntropyModelBase):
  """universal.py:65-330."""

  def __init__(self, prior, coding_rank, compression=False, laplace_tail_mass=0.0, expected_grads=False,
               tail_mass=2**-8, range_coder_precision=12, bottleneck_dtype=None, num_noise_levels=15, stateless=False,
               decode_sanity_check=True):
    super().__init__(coding_rank=coding_rank, compression=compression, stateless=stateless,
                     expected_grads=expected_grads, tail_mass=tail_mass, bottleneck_dtype=bottleneck_dtype,
                     laplace_tail_mass=laplace_tail_mass)
    self._set_prior(prior)
    self._num_noise_levels = int(num_noise_levels)
    self._prior_shape = tuple(int(s) for s in prior.batch_shape)
    if self.coding_rank < len(self.prior_shape):
      raise ValueError("`coding_rank` can't be smaller than `prior_shape`.")
    self.decode_sanity_check = decode_sanity_check
    if self.compression:
      offset = _range_coding_offsets(self._num_noise_levels, len(self.prior_shape), self.bottleneck_dtype)
      cdf, cdf_offset = self._build_tables(self.prior, range_coder_precision, offset=offset)
      self._init_compression(cdf, cdf_offset, None)

  prior_shape = property(lambda self: self._prior_shape)

  @property
  def prior_shape_tensor(self):
    return torch.tensor(self.prior_shape, dtype=torch.int32)

  def _compute_indexes_and_offset(self, broadcast_shape, device):
    """universal.py:147-170 -> (flat table index, quantisation offset), both of shape broadcast_shape + prior_shape."""
    broadcast_shape = tuple(int(d) for d in broadcast_shape)
    prior_size = int(np.prod(self.prior_shape)) if self.prior_shape else 1
    indexes = torch.arange(prior_size, dtype=torch.int32, device=device)
    indexes = torch.broadcast_to(indexes, broadcast_shape + (prior_size,))[..., None]
    indexes = _add_offset_indexes(indexes, self._num_noise_levels)
    offset = _offset_indexes_to_offset(indexes[..., 0], self._num_noise_levels, self.bottleneck_dtype)
    flat = indexes[..., 0] * prior_size + indexes[..., 1]          # strides of index_ranges [levels, prior_size]
    full_shape = broadcast_shape + self.prior_shape
    return flat.reshape(full_shape).to(torch.int32), offset.reshape(full_shape)

  def forward(self, bottleneck, training=True):
    """universal.py:172-211."""
    bottleneck = torch.as_tensor(bottleneck).to(self.bottleneck_dtype)
    log_prob_fn = functools.partial(self._log_prob, self.prior)
    if training:
      log_probs, perturbed = math_ops.perturb_and_apply(log_prob_fn, bottleneck, expected_grads=self.expected_grads)
    else:
      coding_shape = tuple(bottleneck.shape[bottleneck.dim() - self.coding_rank:])
      broadcast_shape = coding_shape[:self.coding_rank - len(self.prior_shape)]
      _, offset = self._compute_indexes_and_offset(broadcast_shape, bottleneck.device)
      perturbed = torch.round(bottleneck - offset) + offset
      log_probs = log_prob_fn(perturbed)
    axes = tuple(range(-self.coding_rank, 0))
    return perturbed, log_probs.sum(dim=axes) / -math.log(2.)

  def compress(self, bottleneck, fused=True):
    """universal.py:213-251."""
    self._check_compression()
    dev = _cuda()
    bottleneck = torch.as_tensor(bottleneck).to(device=dev, dtype=self.bottleneck_dtype)
    shape = tuple(bottleneck.shape)
    batch_shape, coding_shape = shape[:len(shape) - self.coding_rank], shape[len(shape) - self.coding_rank:]
    broadcast_shape = coding_shape[:self.coding_rank - len(self.prior_shape)]
    indexes, offset = self._compute_indexes_and_offset(broadcast_shape, dev)
    indexes = torch.broadcast_to(indexes, shape).contiguous()
    coff = self.cdf_offset.to(dev)
    handle = gen_ops.create_range_encoder(batch_shape, self._lookup_host())
    if fused and bottleneck.dtype == torch.float32:
      F.encode_index_f32(handle, indexes, bottleneck.contiguous(), torch.broadcast_to(offset, shape).contiguous(), coff)
    else:
      symbols = torch.round(bottleneck - offset).to(torch.int32) - coff[indexes.long()]
      gen_ops.entropy_encode_index(handle, indexes, symbols)
    return gen_ops.entropy_encode_finalize(handle)

  def decompress(self, strings, broadcast_shape, fused=True):
    """universal.py:253-289."""
    self._check_compression()
    if not isinstance(strings, gen_ops.Strings):
      strings = gen_ops.Strings.from_bytes(strings)
    dev = strings.bytes_dev.device
    broadcast_shape = tuple(int(d) for d in np.asarray(broadcast_shape).reshape(-1))
    decode_shape = broadcast_shape + self.prior_shape
    output_shape = tuple(strings.shape) + decode_shape
    indexes, offset = self._compute_indexes_and_offset(broadcast_shape, dev)
    indexes = torch.broadcast_to(indexes, output_shape).contiguous()
    offset = torch.broadcast_to(offset, output_shape).contiguous()
    coff = self.cdf_offset.to(dev)
    handle = gen_ops.create_range_decoder(strings, self._lookup_host())
    if fused and self.bottleneck_dtype == torch.float32:
      outputs = F.decode_index_f32(handle, indexes, offset, coff)
      symbols = None
    else:
      handle, symbols = gen_ops.entropy_decode_index(handle, indexes, decode_shape)
    sanity = gen_ops.entropy_decode_finalize(handle)
    if self.decode_sanity_check and not bool(sanity.all()):
      raise gen_ops.InvalidArgumentError("Sanity check failed.")
    if symbols is None:
      return outputs
    return (symbols + coff[indexes.long()]).to(self.bottleneck_dtype) + offset

  def get_config(self):
    raise NotImplementedError()


class UniversalIndexedEntropyModel(ContinuousEntropyModelBase):
  """universal.py:292-603."""

  def __init__(self, prior_fn, index_ranges, parameter_fns, coding_rank, compression=False, laplace_tail_mass=0.0,
               expected_grads=False, tail_mass=2**-8, range_coder_precision=12, bottleneck_dtype=None,
               prior_dtype=torch.float32, stateless=False, num_noise_levels=15, decode_sanity_check=True):
    if coding_rank <= 0:
      raise ValueError("`coding_rank` must be larger than 0.")
    if not callable(prior_fn):
      raise TypeError("`prior_fn` must be a class or factory function.")
    for name, fn in parameter_fns.items():
      if not isinstance(name, str):
        raise TypeError("`parameter_fns` must have string keys.")
      if not callable(fn):
        raise TypeError(f"`parameter_fns['{name}']` must be callable.")
    super().__init__(coding_rank=coding_rank, compression=compression, stateless=stateless,
                     expected_grads=expected_grads, tail_mass=tail_mass, bottleneck_dtype=bottleneck_dtype,
                     laplace_tail_mass=laplace_tail_mass)
    self._index_ranges = tuple([int(num_noise_levels)] + [int(r) for r in index_ranges])  # extra index: noise level
    if len(self._index_ranges) < 2:
      raise ValueError("`index_ranges` must have at least one element.")
    self._prior_fn = prior_fn
    self._parameter_fns = dict(parameter_fns)
    self._prior_dtype = prior_dtype
    self._num_noise_levels = int(num_noise_levels)
    self.decode_sanity_check = decode_sanity_check
    if self.compression:
      grids = torch.meshgrid(*[torch.arange(r, dtype=torch.int32) for r in self.index_ranges_without_offsets], indexing="ij")
      indexes = torch.stack(grids, dim=-1)
      self._set_prior(self._make_prior(indexes), register=False)
      offset = _range_coding_offsets(self._num_noise_levels, len(self.prior.batch_shape), self.bottleneck_dtype)
      cdf, cdf_offset = self._build_tables(self.prior, range_coder_precision, offset=offset)
      self._init_compression(cdf, cdf_offset, None)

  index_ranges = property(lambda self: self._index_ranges)
  parameter_fns = property(lambda self: self._parameter_fns)
  prior_dtype = property(lambda self: self._prior_dtype)
  prior_fn = property(lambda self: self._prior_fn)
  index_ranges_without_offsets = property(lambda self: self._index_ranges[1:])

  def _make_prior(self, indexes):
    indexes = indexes.to(self.prior_dtype)
    return self.prior_fn(**{k: f(indexes) for k, f in self.parameter_fns.items()})

  def _flatten_indexes(self, indexes):
    """universal.py:446-449."""
    strides = np.cumprod((self.index_ranges + (1,))[::-1])[::-1][1:]
    strides = torch.tensor(strides.copy(), dtype=torch.int32, device=indexes.device)
    return (indexes.to(torch.int32) * strides).sum(dim=-1, dtype=torch.int32)  # integer matmul does not exist on CUDA

  def _normalize_indexes(self, indexes):
    """universal.py:451-466: clips every index to its range (with or without the leading noise-level index)."""
    num = indexes.shape[-1]
    ranges = self.index_ranges if num == len(self.index_ranges) else self.index_ranges_without_offsets
    assert num == len(ranges)
    indexes = math_ops.lower_bound(indexes, 0.)
    bounds = torch.tensor([s - 1 for s in ranges], dtype=indexes.dtype, device=indexes.device)
    return math_ops.upper_bound(indexes, bounds.reshape([1] * (indexes.dim() - 1) + [num]))

  def _offset_from_indexes(self, indexes_with_offsets):
    return _offset_indexes_to_offset(indexes_with_offsets[..., 0], self._num_noise_levels, self.bottleneck_dtype)

  def forward(self, bottleneck, indexes, training=True):
    """universal.py:473-528."""
    bottleneck = torch.as_tensor(bottleneck).to(self.bottleneck_dtype)
    indexes = self._normalize_indexes(torch.as_tensor(indexes, dtype=self.prior_dtype, device=bottleneck.device))
    if training:
      def log_prob_fn(x, idx):
        return self._log_prob(self._make_prior(idx), x)
      log_probs, perturbed = math_ops.perturb_and_apply(log_prob_fn, bottleneck, indexes,
                                                        expected_grads=self.expected_grads)
    else:
      prior = self._make_prior(indexes)
      offset = self._offset_from_indexes(_add_offset_indexes(indexes, self._num_noise_levels))
      perturbed = torch.round(bottleneck - offset) + offset
      log_probs = self._log_prob(prior, perturbed)
    axes = tuple(range(-self.coding_rank, 0))
    return perturbed, log_probs.sum(dim=axes) / -math.log(2.)

  def _coding_tensors(self, indexes, dev):
    indexes = torch.as_tensor(indexes).to(device=dev, dtype=self.prior_dtype)
    indexes = self._normalize_indexes(_add_offset_indexes(indexes, self._num_noise_levels))
    return self._flatten_indexes(indexes).contiguous(), self._offset_from_indexes(indexes).contiguous()

  def compress(self, bottleneck, indexes, fused=True):
    """universal.py:530-566."""
    self._check_compression()
    dev = _cuda()
    bottleneck = torch.as_tensor(bottleneck).to(device=dev, dtype=self.bottleneck_dtype)
    flat, offset = self._coding_tensors(indexes, dev)
    fshape = tuple(flat.shape)
    batch_shape = fshape[:len(fshape) - self.coding_rank]
    coff = self.cdf_offset.to(dev)
    handle = gen_ops.create_range_encoder(batch_shape, self._lookup_host())
    if fused and bottleneck.dtype == torch.float32:
      F.encode_index_f32(handle, flat, bottleneck.contiguous(), offset, coff)
    else:
      symbols = torch.round(bottleneck - offset).to(torch.int32) - coff[flat.long()]
      gen_ops.entropy_encode_index(handle, flat, symbols)
    return gen_ops.entropy_encode_finalize(handle)

  def decompress(self, strings, indexes, fused=True):
    """universal.py:568-598."""
    self._check_compression()
    if not isinstance(strings, gen_ops.Strings):
      strings = gen_ops.Strings.from_bytes(strings)
    dev = strings.bytes_dev.device
    flat, offset = self._coding_tensors(indexes, dev)
    fshape = tuple(flat.shape)
    decode_shape = fshape[len(fshape) - self.coding_rank:]
    coff = self.cdf_offset.to(dev)
    handle = gen_ops.create_range_decoder(strings, self._lookup_host())
    if fused and self.bottleneck_dtype == torch.float32:
      outputs = F.decode_index_f32(handle, flat, offset, coff)
      symbols = None
    else:
      handle, symbols = gen_ops.entropy_decode_index(handle, flat, decode_shape)
    sanity = gen_ops.entropy_decode_finalize(handle)
    if self.decode_sanity_check and not bool(sanity.all()):
      raise gen_ops.InvalidArgumentError("Sanity check failed.")
    if symbols is None:
      return outputs
    return (symbols + coff[flat.long()]).to(self.bottleneck_dtype) + offset

  def get_config(self):
    raise NotImplementedError()


def EntropyBottleneck(num_channels=None, prior=None, coding_rank=3, compression=True, **kwargs):
  """TFC-1.x name.  In this snapshot of the reference its role is played by
  `ContinuousBatchedEntropyModel(NoisyDeepFactorized(batch_shape=(C,)), coding_rank=3)`
  (models/bls2017.py:103,160-161); this adaptor builds exactly that."""
  if prior is None:
    if num_channels is None:
      raise ValueError("Either `num_channels` or `prior` must be given.")
    prior = D.NoisyDeepFactorized(batch_shape=(int(num_channels),))
  return ContinuousBatchedEntropyModel(prior, coding_rank=coding_rank, compression=compression, **kwargs)
