"""Soft-rounding layers, mirroring tensorflow_compression/python/layers/soft_round.py:27-56 (elementwise torch
modules over ``math_ops.soft_round*``; training-side glue, no kernels of their own)."""
from torch import nn

from compression_b200 import math_ops

__all__ = ["SoftRound", "SoftRoundConditionalMean"]


class SoftRound(nn.Module):
  """Differentiable approximation of rounding, or its inverse (soft_round.py:27-42)."""

  def __init__(self, alpha=5.0, inverse=False):
    super().__init__()
    self._alpha = alpha
    self._transform = math_ops.soft_round_inverse if inverse else math_ops.soft_round

  def forward(self, inputs):
    return self._transform(inputs, self._alpha)

  def compute_output_shape(self, input_shape):
    return input_shape


class SoftRoundConditionalMean(nn.Module):
  """Conditional mean of the inputs given noisy soft-rounded values (soft_round.py:45-56)."""

  def __init__(self, alpha=5.0):
    super().__init__()
    self._alpha = alpha

  def forward(self, inputs):
    return math_ops.soft_round_conditional_mean(inputs, alpha=self._alpha)

  def compute_output_shape(self, input_shape):
    return input_shape
