"""Generates tests/golden/range_coder_golden.npz from the COMPILED REFERENCE (oracle/_ref: the reference's
cc/lib/range_coder.cc built in place + restated op glue).  Run in the container where /root/reference
exists:   python oracle/make_golden.py
The reference's own tests pin no byte values, so these vectors are the pinned fixtures."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import util  # noqa: E402


def main():
  R = oracle.ref()
  rng = np.random.default_rng(20260922)
  out = {}
  # 1. literal fixture of the reference's tests: CDF {0,16,18,32} @ precision 5
  #    (cc/kernels/range_coding_kernels_test.cc:454-483)
  cdf = np.asarray([[0, 16, 18, 32]], np.int32)
  data = np.asarray([0, 1, 2, 2, 0, 1, 0, 2], np.int16)
  out["lit_cdf"], out["lit_data"] = cdf, data
  out["lit_bytes"] = np.frombuffer(R.range_encode(data, cdf, 5), np.uint8)
  # 2. raw triples (mixed precisions, top-hugging intervals)
  trip = []
  for t in range(64):
    n = int(rng.integers(1, 120))
    prec = rng.integers(1, 17, size=n).astype(np.int32)
    lo = np.empty(n, np.int32)
    hi = np.empty(n, np.int32)
    for i in range(n):
      tot = 1 << prec[i]
      m = rng.integers(0, 3)
      if m == 0:
        a, b = tot - int(rng.integers(1, min(tot, 4) + 1)), tot
      elif m == 1:
        a, b = 0, int(rng.integers(1, tot + 1))
      else:
        a = int(rng.integers(0, tot)); b = int(rng.integers(a + 1, tot + 1))
      lo[i], hi[i] = a, b
    trip.append((lo, hi, prec, np.frombuffer(R.encode_triples(lo, hi, prec), np.uint8)))
  out["trip_n"] = np.asarray([len(t[0]) for t in trip])
  out["trip_lo"] = np.concatenate([t[0] for t in trip])
  out["trip_hi"] = np.concatenate([t[1] for t in trip])
  out["trip_p"] = np.concatenate([t[2] for t in trip])
  out["trip_len"] = np.asarray([len(t[3]) for t in trip])
  out["trip_bytes"] = np.concatenate([t[3] for t in trip])
  # 3. multi-stream channel / index mode with overflow rows
  nrows = 6
  precs = [12, 9, 16, 12, 7, 12]
  cdfs = [util.random_cdf(rng, nb, p, peaky=3) for nb, p in zip((33, 7, 120, 2, 19, 64), precs)]
  ovf = [True, False, True, True, False, True]
  lookup = util.make_lookup_1d(cdfs, precs, ovf, pad=[0, 2, 0, 1, 0, 0])
  S, N = 5, 300
  for mode in ("chan", "index"):
    index = rng.integers(0, nrows, size=(S, N)).astype(np.int32) if mode == "index" else None
    rows = index if index is not None else np.broadcast_to(np.arange(N) % nrows, (S, N))
    nb = np.asarray([len(c) - 1 for c in cdfs])[rows]
    isov = np.asarray(ovf)[rows]
    val = (rng.random((S, N)) * np.where(isov, np.maximum(nb - 1, 1), nb)).astype(np.int32)
    wild = isov & (rng.random((S, N)) < 0.15)
    val[wild] = rng.integers(-2000, 2000, size=int(wild.sum()))
    strings = R.encode(lookup, val, index)
    out[f"{mode}_value"] = val
    if index is not None:
      out[f"{mode}_index"] = index
    out[f"{mode}_len"] = np.asarray([len(s) for s in strings])
    out[f"{mode}_bytes"] = np.frombuffer(b"".join(strings), np.uint8)
  out["ms_lookup"] = lookup
  # 4. legacy op with broadcasting
  lcdf = np.stack([util.random_cdf(rng, 20, 13) for _ in range(7)]).reshape(1, 1, 7, 21)
  ldata = rng.integers(0, 20, size=(2, 9, 7)).astype(np.int16)
  out["leg_cdf"], out["leg_data"] = lcdf, ldata
  out["leg_bytes"] = np.frombuffer(R.range_encode(ldata, lcdf, 13), np.uint8)
  # 5. PmfToQuantizedCdf on tie-free rows (random masses): under-sum and over-sum
  pmf = rng.random((4, 50)).astype(np.float32)
  pmf /= pmf.sum(-1, keepdims=True)
  pmf[:2] *= 0.85
  pmf[2:] *= 1.2
  out["pmf"], out["pmf_cdf"] = pmf, R.pmf_to_cdf(pmf, 10)
  path = os.path.join(ROOT, "tests", "golden", "range_coder_golden.npz")
  np.savez_compressed(path, **out)
  print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
  main()
