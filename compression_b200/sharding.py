"""Multi-GPU plumbing for the hot path: the image batch (= the code streams) is sharded over ranks with no
data-path collective; the only exchange is one broadcast of the range-coding tables from the rank that built
them (tables must never be rebuilt independently, continuous_base.py:175-184).  One process per GPU,
`torch.distributed` (NCCL on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "broadcast_tables", "allreduce_gradients"]


def shard_range(n, rank, world):
  """Contiguous, balanced [lo, hi) slice of n independent units (streams / images) for `rank`."""
  if not 0 <= rank < world:
    raise ValueError("rank out of range")
  base, rem = divmod(int(n), int(world))
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def broadcast_tables(model, src=0, device=None, group=None):
  """Broadcasts `cdf`, `cdf_offset` and (if present) `quantization_offset` of an entropy model from `src`.

  On non-source ranks `model` is typically built with `cdf_shapes=(0, 0)` placeholders; shapes are sent
  first.  Returns the model (tables replaced in place)."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return model
  rank = dist.get_rank(group)
  if device is None:
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
  names = ["_cdf", "_cdf_offset", "_quantization_offset"]
  if rank == src:
    present = [getattr(model, n, None) is not None for n in names]
    meta = torch.tensor([getattr(model, n).numel() if p else -1 for n, p in zip(names, present)],
                        dtype=torch.int64, device=device)
  else:
    meta = torch.zeros(3, dtype=torch.int64, device=device)
  dist.broadcast(meta, src, group=group)
  for n, numel in zip(names, meta.tolist()):
    if numel < 0:
      setattr(model, n, None) if n == "_quantization_offset" else None
      continue
    dtype = torch.float32 if n == "_quantization_offset" else torch.int32
    if rank == src:
      t = getattr(model, n).to(device=device, dtype=dtype).contiguous().reshape(-1)
      shape = torch.tensor(list(getattr(model, n).shape) + [0] * (4 - getattr(model, n).dim()), device=device)
    else:
      t = torch.empty(numel, dtype=dtype, device=device)
      shape = torch.zeros(4, dtype=torch.int64, device=device)
    dist.broadcast(shape, src, group=group)
    dist.broadcast(t, src, group=group)
    dims = [int(d) for d in shape.tolist() if d > 0] if n == "_quantization_offset" else [numel]
    setattr(model, n, t.reshape(dims if dims else [numel]))
  model._cdf_host = None
  return model


def allreduce_gradients(modules, average=True, group=None):
  """Data-parallel training of the transforms ("training only"): sums the parameter gradients of
  `modules` (e.g. the GDN layers: dgamma [C, C] and dbeta [C] per layer, reduced per GPU by the backward kernels)
  over the ranks with ONE all-reduce of a flat buffer -- C^2 + C floats per GDN layer, microseconds over
  NVLink -- and writes the (averaged) result back.  The reference has no counterpart (no tf.distribute in its
  tree); this is the only steady-state collective a sharded training step needs.  Returns the number of elements
  reduced."""
  if isinstance(modules, torch.nn.Module):
    modules = [modules]
  params = [p for m in modules for p in m.parameters() if p.grad is not None]
  if not params:
    return 0
  flat = torch.cat([p.grad.reshape(-1).to(torch.float32) for p in params])
  if dist.is_initialized() and dist.get_world_size(group) > 1:
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
      flat /= dist.get_world_size(group)
  at = 0
  for p in params:
    n = p.grad.numel()
    p.grad.copy_(flat[at:at + n].reshape(p.grad.shape).to(p.grad.dtype))
    at += n
  return at
