"""`tfc.ops` namespace (tensorflow_compression/python/ops/__init__.py:17-20): the coder / quantisation ops of
`gen_ops` (CUDA kernels behind the C ABI) and the differentiable helpers of `math_ops` / `round_ops`."""
from compression_b200.gen_ops import (create_range_decoder, create_range_encoder, entropy_decode_channel,
                                      entropy_decode_finalize, entropy_decode_index, entropy_encode_channel,
                                      entropy_encode_finalize, entropy_encode_index, pmf_to_quantized_cdf, range_decode,
                                      range_encode, run_length_decode, run_length_encode, run_length_gamma_decode,
                                      run_length_gamma_encode, stochastic_round)
from compression_b200.math_ops import (lower_bound, perturb_and_apply, round_st, soft_round, soft_round_conditional_mean,
                                       soft_round_inverse, upper_bound)
from compression_b200.signal_conv import same_padding_for_kernel

__all__ = [
    "create_range_decoder", "create_range_encoder", "entropy_decode_channel", "entropy_decode_finalize",
    "entropy_decode_index", "entropy_encode_channel", "entropy_encode_finalize", "entropy_encode_index",
    "pmf_to_quantized_cdf", "range_decode", "range_encode", "run_length_decode", "run_length_encode",
    "run_length_gamma_decode", "run_length_gamma_encode", "stochastic_round", "lower_bound", "perturb_and_apply",
    "round_st", "soft_round", "soft_round_conditional_mean", "soft_round_inverse", "upper_bound",
    "same_padding_for_kernel",
]
