"""`tfc.layers` namespace (tensorflow_compression/python/layers/__init__.py:17-21): the layer classes, their
parameter reparameterisations and initialisers, gathered from the modules that implement them."""
from compression_b200.gdn import GDN, GDNParameter
from compression_b200.parameters import Parameter
from compression_b200.signal_conv import (IdentityInitializer, RDFTParameter, SignalConv1D, SignalConv2D, SignalConv3D)
from compression_b200.soft_round_layers import SoftRound, SoftRoundConditionalMean

__all__ = ["GDN", "GDNParameter", "Parameter", "IdentityInitializer", "RDFTParameter", "SignalConv1D", "SignalConv2D", "SignalConv3D",
           "SoftRound", "SoftRoundConditionalMean"]
