"""Writes tests/golden/cfg2_tables.npz: the cfg2 range-coding tables exactly as the product's table builder makes
them (ContinuousBatchedEntropyModel over the bench's NoisyLaplace priors -> tfcb_build_lookup on the GPU).
`bench.py --impl reference` codes with these tables without importing compression_b200; the GPU arm checks that
the tables it builds are identical (`config.tables_match_fixture`).

Run on a GPU box:  python tools/make_cfg_fixtures.py [out.npz]   (default: gpurun_out/cfg2_tables.npz)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
  import torch
  out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "cfg2_tables.npz")
  scales, _ = bench.synth_latents(0, 0)
  model = bench.build_model(scales, torch.device("cuda", 0))
  q = model.quantization_offset
  os.makedirs(os.path.dirname(out), exist_ok=True)
  np.savez_compressed(out, lookup=model.cdf.cpu().numpy().astype(np.int32),
                      cdf_offset=model.cdf_offset.cpu().numpy().astype(np.int32),
                      has_qoff=np.asarray(q is not None),
                      quantization_offset=(np.zeros(0, np.float32) if q is None else q.cpu().numpy().astype(np.float32)))
  print("wrote", out, "lookup ints:", model.cdf.numel())


if __name__ == "__main__":
  main()
