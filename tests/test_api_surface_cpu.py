"""CPU: every public name of the reference package resolves here (tensorflow_compression/__init__.py:17-40 star
imports the `__all__` lists below), plus its four sub-namespaces.  Importing needs neither a GPU nor the built
library."""
import pytest

import compression_b200 as tfc

REFERENCE_ALL = {
    "datasets/y4m_dataset.py": ["Y4MDataset"],
    "distributions/deep_factorized.py": ["DeepFactorized", "NoisyDeepFactorized"],
    "distributions/helpers.py": ["estimate_tails", "quantization_offset", "lower_tail", "upper_tail"],
    "distributions/round_adapters.py": ["MonotonicAdapter", "RoundAdapter", "NoisyRoundedNormal",
                                        "NoisyRoundedDeepFactorized", "SoftRoundAdapter", "NoisySoftRoundedNormal",
                                        "NoisySoftRoundedDeepFactorized"],
    "distributions/uniform_noise.py": ["UniformNoiseAdapter", "NoisyMixtureSameFamily", "NoisyNormal", "NoisyLogistic",
                                       "NoisyLaplace", "NoisyNormalMixture", "NoisyLogisticMixture"],
    "entropy_models/continuous_batched.py": ["ContinuousBatchedEntropyModel"],
    "entropy_models/continuous_indexed.py": ["ContinuousIndexedEntropyModel", "LocationScaleIndexedEntropyModel"],
    "entropy_models/power_law.py": ["PowerLawEntropyModel"],
    "entropy_models/universal.py": ["UniversalBatchedEntropyModel", "UniversalIndexedEntropyModel"],
    "layers/gdn.py": ["GDN"],
    "layers/initializers.py": ["IdentityInitializer"],
    "layers/parameters.py": ["Parameter", "RDFTParameter", "GDNParameter"],
    "layers/signal_conv.py": ["SignalConv1D", "SignalConv2D", "SignalConv3D"],
    "layers/soft_round.py": ["SoftRound", "SoftRoundConditionalMean"],
    "ops/gen_ops.py": ["create_range_encoder", "create_range_decoder", "entropy_decode_channel", "entropy_decode_finalize",
                       "entropy_decode_index", "entropy_encode_channel", "entropy_encode_finalize",
                       "entropy_encode_index", "pmf_to_quantized_cdf", "run_length_decode", "run_length_encode",
                       "run_length_gamma_decode", "run_length_gamma_encode", "stochastic_round"],
    "ops/math_ops.py": ["upper_bound", "lower_bound", "perturb_and_apply"],
    "ops/padding_ops.py": ["same_padding_for_kernel"],
    "ops/round_ops.py": ["round_st", "soft_round", "soft_round_inverse", "soft_round_conditional_mean"],
    "util/packed_tensors.py": ["PackedTensors"],
}


@pytest.mark.parametrize("module", sorted(REFERENCE_ALL))
def test_every_public_name_of_the_reference_resolves(module):
  for name in REFERENCE_ALL[module]:
    obj = getattr(tfc, name)
    assert callable(obj), (module, name)
    sub = module.split("/")[0]
    if sub in ("layers", "ops"):                      # the sub-namespaces re-export their own names
      assert getattr(getattr(tfc, sub), name) is obj
  with pytest.raises(AttributeError):
    tfc.no_such_name  # pylint:disable=pointless-statement


def test_sub_namespaces_exist():
  for sub in ("distributions", "entropy_models", "layers", "ops"):
    assert getattr(tfc, sub).__name__ == "compression_b200." + sub
  assert tfc.entropy_models.ContinuousBatchedEntropyModel is tfc.ContinuousBatchedEntropyModel
  assert tfc.distributions.NoisyNormal is tfc.NoisyNormal
  # the legacy ops and the two run-length models the reference keeps out of its top-level star imports
  assert callable(tfc.range_encode) and callable(tfc.range_decode) and callable(tfc.LaplaceEntropyModel)
