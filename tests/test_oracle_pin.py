"""CPU: pins the oracle.  The C port must reproduce (a) the golden vectors generated from the compiled
reference, (b) the compiled reference itself on fresh seeded fuzz (where oracle/_ref is present)."""
import numpy as np
import pytest

import golden_util
import oracle
import util

FLAVOURS = ["port"] + (["ref"] if oracle.have_ref() else [])


def _o(name):
  return oracle.port() if name == "port" else oracle.ref()


@pytest.mark.parametrize("flavour", FLAVOURS)
def test_golden_vectors(flavour):
  O, g = _o(flavour), golden_util.load()
  assert O.range_encode(g["lit_data"], g["lit_cdf"], 5) == bytes(g["lit_bytes"])
  assert np.array_equal(O.range_decode(bytes(g["lit_bytes"]), g["lit_data"].shape, g["lit_cdf"], 5), g["lit_data"])
  at = 0
  want = golden_util.split(g["trip_bytes"], g["trip_len"])
  for n, w in zip(g["trip_n"], want):
    sl = slice(at, at + int(n))
    assert O.encode_triples(g["trip_lo"][sl], g["trip_hi"][sl], g["trip_p"][sl]) == w
    at += int(n)
  for mode in ("chan", "index"):
    idx = g.get(f"{mode}_index")
    want = golden_util.split(g[f"{mode}_bytes"], g[f"{mode}_len"])
    assert O.encode(g["ms_lookup"], g[f"{mode}_value"], idx) == want
    dec, ok = O.decode(g["ms_lookup"], want, g[f"{mode}_value"].shape[1], idx)
    assert np.array_equal(dec, g[f"{mode}_value"]) and ok.all()
  assert O.range_encode(g["leg_data"], g["leg_cdf"], 13) == bytes(g["leg_bytes"])
  assert np.array_equal(O.range_decode(bytes(g["leg_bytes"]), g["leg_data"].shape, g["leg_cdf"], 13), g["leg_data"])
  assert np.array_equal(O.pmf_to_cdf(g["pmf"], 10), g["pmf_cdf"])


@pytest.mark.skipif(not oracle.have_ref(), reason="compiled reference not present")
def test_port_equals_compiled_reference_fuzz():
  P, R = oracle.port(), oracle.ref()
  rng = np.random.default_rng(77)
  for _ in range(400):
    n = int(rng.integers(0, 150))
    prec = rng.integers(1, 17, size=n).astype(np.int32)
    tot = (1 << prec.astype(np.int64))
    lo = (rng.random(n) * tot).astype(np.int64)
    hi = lo + 1 + (rng.random(n) * (tot - lo - 1)).astype(np.int64)
    hug = rng.random(n) < 0.3
    hi[hug] = tot[hug]
    lo[hug] = tot[hug] - 1 - (rng.random(int(hug.sum())) * np.minimum(tot[hug] - 1, 3)).astype(np.int64)
    assert P.encode_triples(lo, hi, prec) == R.encode_triples(lo, hi, prec)
  for t in range(60):
    nrows, S, N = int(rng.integers(1, 6)), int(rng.integers(1, 4)), int(rng.integers(0, 200))
    precs = [int(rng.integers(5, 17)) for _ in range(nrows)]
    cdfs = [util.random_cdf(rng, int(rng.integers(2, min(40, 1 << p) + 1)), p, peaky=3) for p in precs]
    ovf = [bool(rng.integers(0, 2)) for _ in range(nrows)]
    lookup = util.make_lookup_2d(cdfs, precs, ovf) if t % 2 else util.make_lookup_1d(cdfs, precs, ovf)
    index = rng.integers(0, nrows, size=(S, N)).astype(np.int32) if t % 3 == 0 else None
    rows = index if index is not None else np.broadcast_to(np.arange(N) % nrows, (S, N))
    nb = np.asarray([len(c) - 1 for c in cdfs])[rows]
    isov = np.asarray(ovf)[rows]
    val = (rng.random((S, N)) * np.where(isov, np.maximum(nb - 1, 1), nb)).astype(np.int32)
    wild = isov & (rng.random((S, N)) < 0.2)
    val[wild] = rng.integers(-300, 300, size=int(wild.sum()))
    a, b = P.encode(lookup, val, index), R.encode(lookup, val, index, threads=2)
    assert a == b
    da, oka = P.decode(lookup, a, N, index)
    db, okb = R.decode(lookup, a, N, index, threads=2)
    assert np.array_equal(da, val) and np.array_equal(db, val) and oka.all() and okb.all()


def test_decoder_sanity_flag_semantics():
  """RangeDecoder::Finalize (range_coder.h:144-169): unread bytes -> False; exact consumption -> True."""
  O = oracle.port()
  rng = np.random.default_rng(5)
  cdf = util.random_cdf(rng, 25, 12)
  lookup = util.make_lookup_1d([cdf], [12], [False])
  val = util.sample_symbols(rng, cdf, 900).reshape(3, 300)
  s = O.encode(lookup, val)
  _, ok = O.decode(lookup, s, 300)
  assert ok.all()
  _, ok = O.decode(lookup, s, 10)
  assert not ok.any()
  _, ok = O.decode(lookup, [x + b"\x07\x07\x07\x07" for x in s], 300)
  assert not ok.any()


def test_lookup_grammar_errors():
  """ScanCDF / IndexCDFMatrix error classes (range_coder_kernels.cc:110-164)."""
  O = oracle.port()
  for bad, msg in (([4, 1, 16], "start with 0"), ([4, 0, 3, 2, 16], "monotonically"), ([4, 0, 3], "end with"),
                   ([17, 0, 4], "precision"), ([4, 0], "prematurely")):
    with pytest.raises(oracle.OracleError, match=msg):
      O.encoder(np.asarray(bad, np.int32), 1)
  with pytest.raises(oracle.OracleError, match="end with"):
    O.encoder(np.asarray([[4, 0, 16, 16, 3]], np.int32), 1)  # 2-D row not filled with padding


def test_pmf_to_cdf_invariants_and_tie_rule():
  """pmf_to_cdf_kernels_test.cc:70-143 invariants; symmetric rows (exact ties) keep the invariants and the
  port resolves ties by lowest bin index."""
  O = oracle.port()
  rng = np.random.default_rng(9)
  for n, p, scale in ((32, 10, 0.85), (100, 7, 1.0), (41, 12, 1.15)):
    pmf = rng.random((3, n)).astype(np.float32)
    pmf[1, n // 2:] = 0
    k = np.arange(n) - (n - 1) / 2
    pmf[2] = np.exp(-0.5 * (k / (n / 8))**2)  # symmetric -> exact ties
    pmf = (pmf / pmf.sum(-1, keepdims=True) * scale).astype(np.float32)
    cdf = O.pmf_to_cdf(pmf, p)
    assert (cdf[:, 0] == 0).all() and (cdf[:, -1] == 1 << p).all() and (np.diff(cdf, axis=-1) >= 1).all()
  with pytest.raises(oracle.OracleError):
    O.pmf_to_cdf(np.asarray([[0.5, np.nan]], np.float32), 8)


def test_stochastic_round_port_equals_reference_flavour():
  """quantization_kernels.cc:48-95: the C restatement of std::seed_seq + xoshiro256+ against libstdc++'s own
  seed_seq driving the same loop (oracle/ref/ref_driver.cc), and the reference's invariants
  (python/ops/quantization_ops_test.py:28-83) on the oracle itself."""
  if not oracle.have_ref():
    pytest.skip("reference flavour of the oracle not built")
  P, R = oracle.port(), oracle.ref()
  rng = np.random.default_rng(0)
  for seed in ([1], [123, 456], [5] * 9, [-7, 2**31 - 1, 0, 3, 4, 5, 6, 7, 8, 9, 10]):
    x = rng.uniform(-100, 100, 20000).astype(np.float32)
    a, b = P.stochastic_round(x, 0.75, seed), R.stochastic_round(x, 0.75, seed)
    assert np.array_equal(a, b)
    assert np.all(np.abs(a * np.float32(0.75) - x) <= 0.75 + 1e-4)
  ints = rng.integers(-100, 100, 100).astype(np.float32)
  assert np.array_equal(P.stochastic_round(ints * np.float32(0.75), 0.75, [3]), ints.astype(np.int32))
  rep = np.broadcast_to(rng.uniform(-100, 100, 20).astype(np.float32), (20000, 20))
  assert np.abs(P.stochastic_round(rep, 1.0, [9]).mean(0) - rep[0]).max() < 3e-2


def test_run_length_port_reproduces_the_reference_literal():
  """cc/kernels/run_length_kernels_test.cc:272-305 holds the one literal bit string of the run-length ops:
  [-6, 3, 0, 0] <-> {0b11010001, 0b01101101} (gamma / gamma / zeros only); plus round trips over every code flavour."""
  P = oracle.port()
  assert P.run_length_encode([-6, 3, 0, 0]) == bytes([0b11010001, 0b01101101])
  assert P.run_length_decode(bytes([0b11010001, 0b01101101]), (4,)).tolist() == [-6, 3, 0, 0]
  rng = np.random.default_rng(1)
  for rl, mg, nz in ((-1, -1, False), (-1, -1, True), (2, 3, True), (0, 0, False), (5, -1, False), (-1, 4, True)):
    for density in (0.02, 0.5, 1.0):
      d = (rng.integers(-300, 300, 5000) * (rng.random(5000) < density)).astype(np.int32)
      code = P.run_length_encode(d, rl, mg, nz)
      assert np.array_equal(P.run_length_decode(code, d.shape, rl, mg, nz), d)
  with pytest.raises(oracle.OracleError, match="Out of bits"):
    P.run_length_decode(b"\x01", (9,))


@pytest.mark.parametrize("flavour", FLAVOURS)
def test_escape_payloads_the_reference_cannot_finish_are_refused(flavour):
  """OverflowEncode's width loop `while (gamma >= (1 << n))` (range_coder_kernels.cc:310-315, "TODO Clamp gamma")
  never ends once the payload reaches 2^30.  The checker refuses such a symbol instead of hanging (bench.py found
  this through an infinite synthetic latent); the largest payloads below the limit still round-trip."""
  O = _o(flavour)
  lookup = np.asarray([-4, 0, 5, 11, 16], np.int32)  # overflow row, escape symbol 2: payload = v - 1 for v >= 2
  fine = np.asarray([[0, 1, 1 << 30, -((1 << 30) - 1), 7, -3]], np.int32)
  s = O.encode(lookup, fine)
  back, ok = O.decode(lookup, s, fine.shape[1])
  assert np.array_equal(back, fine) and ok.all()
  for v in ((1 << 30) + 1, -(1 << 30), np.iinfo(np.int32).min, np.iinfo(np.int32).max):
    with pytest.raises(oracle.OracleError, match="Elias-gamma payload"):
      O.encode(lookup, np.asarray([[0, v, 0]], np.int32))


def test_bench_latents_are_finite_on_every_rank():
  """bench.synth_latents: rand() == 0 used to give log1p(-1) = -inf, i.e. a latent that quantises to INT32_MIN;
  rank 1's first batch had one, and the per-rank parity check then waited for ever inside the oracle."""
  import torch
  import bench
  for rank in (0, 1, 7):
    _, ys = bench.synth_latents(rank, 2, batch=64)
    assert all(bool(torch.isfinite(y).all()) and float(y.abs().max()) < 8.0 * 17.0 + 1e-3 for y in ys)
  _, full = bench.synth_latents(1, 1)   # the batch that hung the 2-GPU run
  assert bool(torch.isfinite(full[0]).all())
