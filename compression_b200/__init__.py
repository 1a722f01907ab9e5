"""compression_b200 -- B200-native (sm_100a) implementation of tensorflow/compression's data-parallel hot
path: range coding (multi-stream and legacy ops), PMF->CDF integerisation and GDN/IGDN, behind the
reference's own operator API (``gen_ops``, ``ContinuousBatchedEntropyModel``,
``LocationScaleIndexedEntropyModel``, ``GDN``).  All compute runs in hand-written CUDA kernels reached
through the C ABI of ``include/tfcb200.h``; there is no CPU fallback.
"""
from compression_b200 import _lib
from compression_b200._lib import InvalidArgumentError


def __getattr__(name):  # lazy: importing the package must not require torch / the built library
  import importlib
  lazy = {
      "gen_ops": "compression_b200.gen_ops",
  }
  if name in lazy:
    return importlib.import_module(lazy[name])
  raise AttributeError(name)
