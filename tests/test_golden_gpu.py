"""GPU vs the committed golden vectors (generated from the compiled reference, oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

import golden_util

pytestmark = pytest.mark.gpu


def test_gpu_reproduces_golden_vectors():
  from compression_b200 import gen_ops as ops
  g = golden_util.load()
  lit = ops.range_encode(torch.from_numpy(g["lit_data"]).cuda(), torch.from_numpy(g["lit_cdf"]).cuda(), 5)
  assert lit == bytes(g["lit_bytes"])
  for mode in ("chan", "index"):
    idx = g.get(f"{mode}_index")
    want = golden_util.split(g[f"{mode}_bytes"], g[f"{mode}_len"])
    val = torch.from_numpy(g[f"{mode}_value"]).cuda()
    h = ops.create_range_encoder([val.shape[0]], g["ms_lookup"])
    if idx is None:
      ops.entropy_encode_channel(h, val)
    else:
      ops.entropy_encode_index(h, torch.from_numpy(idx).cuda(), val)
    assert ops.entropy_encode_finalize(h).tolist() == want
    hd = ops.create_range_decoder(want, g["ms_lookup"])
    if idx is None:
      hd, dec = ops.entropy_decode_channel(hd, [val.shape[1]])
    else:
      hd, dec = ops.entropy_decode_index(hd, torch.from_numpy(idx).cuda(), [val.shape[1]])
    assert torch.equal(dec, val) and bool(ops.entropy_decode_finalize(hd).all())
  leg = ops.range_encode(torch.from_numpy(g["leg_data"]).cuda(), torch.from_numpy(g["leg_cdf"]).cuda(), 13)
  assert leg == bytes(g["leg_bytes"])
  dec = ops.range_decode(leg, list(g["leg_data"].shape), torch.from_numpy(g["leg_cdf"]).cuda(), 13)
  assert np.array_equal(dec.cpu().numpy(), g["leg_data"])
  cdf = ops.pmf_to_quantized_cdf(torch.from_numpy(g["pmf"]).cuda(), 10)
  assert np.array_equal(cdf.cpu().numpy(), g["pmf_cdf"])
