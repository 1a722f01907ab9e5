"""`Parameter`: the base of the reparameterised layer variables (tensorflow_compression/python/layers/parameters.py:
31-59) -- a module whose call returns the parameter's value as a function of the variables it stores.
`GDNParameter` (gdn.py) and `RDFTParameter` (signal_conv.py) derive from it."""
from typing import Sequence

import torch
from torch import nn

__all__ = ["Parameter"]


class Parameter(nn.Module):
  """Reparameterised layer variable: `parameter(compute_dtype=None)` computes its value."""

  def forward(self, compute_dtype=None):
    raise NotImplementedError

  def get_config(self):
    return dict(name=getattr(self, "name", None))

  def get_weights(self):
    """The stored variables as numpy arrays, in `parameters()` order (parameters.py:48-49)."""
    return [p.detach().cpu().numpy() for p in self.parameters()]

  def set_weights(self, weights: Sequence):
    variables = list(self.parameters())
    if len(weights) != len(variables):
      raise ValueError(f"set_weights() expects a list of {len(variables)} arrays, received {len(weights)}.")
    with torch.no_grad():
      for p, w in zip(variables, weights):
        p.copy_(torch.as_tensor(w, dtype=p.dtype).reshape(p.shape))
