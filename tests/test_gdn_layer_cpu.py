"""CPU: the GDN layer's host logic that runs no kernel (python/layers/gdn_test.py:27-40,90-135): argument checks,
which parameters become variables, attributes frozen by build().  The arithmetic is in tests/test_gdn_gpu.py."""
import pytest
import torch

import compression_b200 as tfc


def test_invalid_data_format_and_vector_input_raise():
  with pytest.raises(ValueError):
    tfc.GDN(data_format="NHWC")
  for data_format in ("channels_first", "channels_last"):
    with pytest.raises(ValueError):
      tfc.GDN(data_format=data_format)(torch.zeros(3))


def test_variables_are_enumerated():
  layer = tfc.GDN(alpha_parameter=None, epsilon_parameter=None)
  layer.build((None, 5))
  names = sorted(n for n, _ in layer.named_parameters())
  assert names == ["alpha_parameter.variable", "beta_parameter.variable", "epsilon_parameter.variable",
                   "gamma_parameter.variable"]
  assert all(p.requires_grad for p in layer.parameters())
  layer = tfc.GDN()                 # fixed exponents (the models' default): beta and gamma only
  layer.build((None, 5))
  assert sorted(n for n, _ in layer.named_parameters()) == ["beta_parameter.variable", "gamma_parameter.variable"]
  assert layer.beta_parameter().shape == (5,) and layer.gamma_parameter().shape == (5, 5)
  assert torch.allclose(layer.beta_parameter(), torch.ones(5), atol=1e-6)               # gdn.py:135-138 defaults
  assert torch.allclose(layer.gamma_parameter(), 0.1 * torch.eye(5), atol=1e-6)


def test_variables_are_not_enumerated_when_overridden():
  layer = tfc.GDN(beta_parameter=torch.tensor([1.]), gamma_parameter=torch.tensor([[.1]]))
  layer.build((None, 1))
  assert list(layer.parameters()) == []


def test_attributes_cannot_be_set_after_build():
  layer = tfc.GDN()
  layer.build((None, 2))
  for name, value in (("inverse", True), ("rectify", True), ("data_format", "channels_first"), ("alpha_parameter", 5),
                      ("beta_parameter", torch.ones(5)), ("gamma_parameter", torch.ones(5, 5)),
                      ("epsilon_parameter", 1 / 3)):
    with pytest.raises(RuntimeError):
      setattr(layer, name, value)
