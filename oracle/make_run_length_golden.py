"""Generates tests/golden/run_length_golden.npz from the COMPILED REFERENCE bit coder (oracle/_ref: the reference's
cc/lib/bit_coder.cc built in place, the op loops of cc/kernels/run_length_kernels.cc restated around it).  Run in the
container where /root/reference exists:   python oracle/make_run_length_golden.py
The reference's tests hold one literal bit string for these ops (run_length_kernels_test.cc:272-305, first vector
below); the rest are fixtures so that the C port stays pinned where /root/reference (hence oracle/_ref) is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CONFIGS = [(-1, -1, False), (-1, -1, True), (0, 0, False), (2, 3, True), (5, -1, False), (-1, 4, True), (3, 0, True)]


def vectors():
  """Data kept where the reference's writer is defined (unary parts of Rice codes below 57 zeros, see
  tests/test_oracle_pin.py)."""
  rng = np.random.default_rng(20260923)
  yield (-1, -1, False), np.asarray([-6, 3, 0, 0], np.int32)
  for rl, mg, nz in CONFIGS:
    for density, mag in ((0.05, 30), (0.6, 2000), (1.0, 7)):
      n = int(rng.integers(40, 400))
      if mg >= 0:
        mag = min(mag, 56 << mg)
      d = (rng.integers(-mag, mag + 1, n) * (rng.random(n) < density)).astype(np.int32)
      if rl >= 0:
        step = 56 << rl
        d[step - 1::step] = np.where(d[step - 1::step] == 0, 1, d[step - 1::step])
        if nz:
          d[step // 2::step] = 0
      yield (rl, mg, nz), d


def main():
  R = oracle.ref()
  params, data, codes = [], [], []
  for (rl, mg, nz), d in vectors():
    code = R.run_length_encode(d, rl, mg, nz)
    assert np.array_equal(R.run_length_decode(code, d.shape, rl, mg, nz), d)
    params.append((rl, mg, int(nz)))
    data.append(d)
    codes.append(np.frombuffer(code, np.uint8))
  out = dict(params=np.asarray(params, np.int32), data_len=np.asarray([len(d) for d in data]),
             data=np.concatenate(data), code_len=np.asarray([len(c) for c in codes]), code=np.concatenate(codes))
  path = os.path.join(ROOT, "tests", "golden", "run_length_golden.npz")
  np.savez_compressed(path, **out)
  print(path, len(params), "vectors,", int(out["code_len"].sum()), "code bytes")


if __name__ == "__main__":
  main()
