"""ORACLE -- test infrastructure, NOT product code.

CPU restatement of the reference's hot-path algorithm (range coder, lookup grammar, overflow coding,
legacy broadcast addressing, PmfToQuantizedCdf) in two flavours behind one Python API:

* ``port()`` -- ``oracle/port/tfc_port.c``: plain-C restatement (always available).
* ``ref()``  -- ``oracle/_ref/libtfc_ref.so``: the reference's own ``cc/lib/range_coder.cc`` compiled
  in place from ``/root/reference`` plus a restatement of the TF-entangled op glue
  (``oracle/ref/ref_driver.cc``).  Built by ``oracle/Makefile`` where the reference tree exists; the
  built ``.so`` travels to the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  ``compression_b200`` never does.

Reference citations for every routine are in the two source files.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "port", "libtfc_port.so")
_REF_SO = os.path.join(_HERE, "_ref", "libtfc_ref.so")
REFERENCE_ROOT = os.environ.get("TFCB_REFERENCE_ROOT", "/root/reference")


class OracleError(ValueError):
  """InvalidArgument-class failure reported by the oracle."""


def build(want_ref: bool = True) -> None:
  """Compiles the C port and, when the reference tree is present, the in-place reference build."""
  subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
  if want_ref and os.path.exists(
      os.path.join(REFERENCE_ROOT, "tensorflow_compression/cc/lib/range_coder.cc")):
    subprocess.run(["make", "-s", "-C", _HERE, "ref", f"REFERENCE={REFERENCE_ROOT}"], check=True)


def _i32(a) -> np.ndarray:
  return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray, t):
  return a.ctypes.data_as(C.POINTER(t))


class Oracle:
  """ctypes front-end shared by the port and the compiled reference (same entry points, other prefix)."""

  def __init__(self, path: str, prefix: str, kind: str):
    self.kind = kind
    self.path = path
    self._lib = C.CDLL(path)
    self._px = prefix
    f = self._fn
    f("last_error", C.c_char_p, [])
    f("encode_triples", C.c_int64,
      [C.POINTER(C.c_int32)] * 3 + [C.c_int64, C.POINTER(C.c_uint8), C.c_int64])
    f("encoder_create", C.c_void_p, [C.POINTER(C.c_int32), C.c_int64, C.c_int64, C.c_int64])
    f("encoder_encode", C.c_int,
      [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64, C.c_int])
    f("encoder_finalize", C.c_int, [C.c_void_p, C.POINTER(C.c_int64)])
    f("encoder_bytes", None, [C.c_void_p, C.POINTER(C.c_uint8)])
    f("encoder_free", None, [C.c_void_p])
    f("decoder_create", C.c_void_p, [
        C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int32), C.c_int64,
        C.c_int64
    ])
    f("decoder_decode", C.c_int,
      [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64, C.c_int])
    f("decoder_finalize", C.c_int, [C.c_void_p, C.POINTER(C.c_uint8)])
    f("decoder_free", None, [C.c_void_p])
    f("range_encode", C.c_int64, [
        C.POINTER(C.c_int16), C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int32),
        C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_int64
    ])
    f("range_decode", C.c_int, [
        C.POINTER(C.c_uint8), C.c_int64, C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int32),
        C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int16)
    ])
    f("pmf_to_cdf", C.c_int,
      [C.POINTER(C.c_float), C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_int32)])
    f("hardware_threads", C.c_int, [])

  def _fn(self, name, restype, argtypes):
    fn = getattr(self._lib, self._px + name)
    fn.restype = restype
    fn.argtypes = argtypes
    setattr(self, "_" + name, fn)

  def _err(self) -> str:
    return (self._last_error() or b"").decode()

  def hardware_threads(self) -> int:
    return int(self._hardware_threads())

  # ---- raw (lower, upper, precision) triples: RangeEncoder::Encode + Finalize ----
  def encode_triples(self, lower, upper, precision) -> bytes:
    lower, upper, precision = _i32(lower), _i32(upper), _i32(precision)
    n = lower.size
    cap = 2 * n + 16
    out = np.empty(cap, dtype=np.uint8)
    r = self._encode_triples(
        _p(lower, C.c_int32), _p(upper, C.c_int32), _p(precision, C.c_int32), n,
        _p(out, C.c_uint8), cap)
    assert r >= 0
    return out[:r].tobytes()

  # ---- stateful multi-stream ops ----
  def encoder(self, lookup, n_streams: int) -> "OracleEncoder":
    return OracleEncoder(self, lookup, n_streams)

  def decoder(self, strings: Sequence[bytes], lookup) -> "OracleDecoder":
    return OracleDecoder(self, strings, lookup)

  def encode(self, lookup, value, index=None, threads: int = 1) -> List[bytes]:
    """value: int32 [S, N] (+ optional index [S, N]); returns S byte strings."""
    value = _i32(value)
    assert value.ndim == 2
    enc = self.encoder(lookup, value.shape[0])
    try:
      enc.encode(value, index, threads)
      return enc.finalize()
    finally:
      enc.close()

  def decode(self, lookup, strings: Sequence[bytes], n_per_stream: int, index=None,
             threads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    dec = self.decoder(strings, lookup)
    try:
      out = dec.decode(n_per_stream, index, threads)
      return out, dec.finalize()
    finally:
      dec.close()

  # ---- legacy RangeEncode / RangeDecode ----
  def range_encode(self, data, cdf, precision: int, debug_level: int = 1) -> bytes:
    data = np.ascontiguousarray(data, dtype=np.int16)
    cdf = _i32(cdf)
    ds = np.asarray(data.shape, dtype=np.int64)
    cs = np.asarray(cdf.shape, dtype=np.int64)
    cap = 2 * data.size + 16
    out = np.empty(cap, dtype=np.uint8)
    r = self._range_encode(
        _p(data, C.c_int16), _p(ds, C.c_int64), data.ndim, _p(cdf, C.c_int32), _p(cs, C.c_int64),
        cdf.ndim, precision, debug_level, _p(out, C.c_uint8), cap)
    if r < 0:
      raise OracleError(self._err())
    return out[:r].tobytes()

  def range_decode(self, encoded: bytes, shape, cdf, precision: int,
                   debug_level: int = 1) -> np.ndarray:
    cdf = _i32(cdf)
    ds = np.asarray(list(shape), dtype=np.int64)
    cs = np.asarray(cdf.shape, dtype=np.int64)
    buf = np.frombuffer(encoded, dtype=np.uint8) if len(encoded) else np.zeros(1, np.uint8)
    buf = np.ascontiguousarray(buf)
    out = np.empty(tuple(int(d) for d in ds), dtype=np.int16)
    rc = self._range_decode(
        _p(buf, C.c_uint8), len(encoded), _p(ds, C.c_int64), len(ds), _p(cdf, C.c_int32),
        _p(cs, C.c_int64), cdf.ndim, precision, debug_level, _p(out, C.c_int16))
    if rc != 0:
      raise OracleError(self._err())
    return out

  # ---- PmfToQuantizedCdf ----
  def stochastic_round(self, inputs, step_size: float, seed) -> np.ndarray:
    """StochasticRound (quantization_kernels.cc:48-95) with an explicit seed; `inputs` float32 (already promoted)."""
    x = np.ascontiguousarray(np.asarray(inputs, dtype=np.float32))
    sd = _i32(np.asarray(seed).reshape(-1))
    out = np.empty(x.shape, np.int32)
    if not hasattr(self, "_stochastic_round"):
      self._fn("stochastic_round", C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_void_p])
    if self._stochastic_round(x.ctypes.data, x.size, float(step_size), sd.ctypes.data, sd.size, out.ctypes.data) != 0:
      raise OracleError(self._err())
    return out

  def run_length_encode(self, data, run_length_code=-1, magnitude_code=-1, use_run_length_for_non_zeros=False) -> bytes:
    """RunLengthEncode (cc/kernels/run_length_kernels.cc:52-139): the C port restates the bit packing too; the
    reference flavour runs the op loop around the reference's own BitWriter (cc/lib/bit_coder.cc compiled in place)."""
    d = _i32(np.asarray(data).reshape(-1))
    cap = int(16 + 20 * d.size)
    while True:
      out = np.zeros(cap, np.uint8)
      if not hasattr(self, "_run_length_encode"):
        self._fn("run_length_encode", C.c_int64, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64])
      nb = self._run_length_encode(d.ctypes.data, d.size, int(run_length_code), int(magnitude_code),
                                   int(bool(use_run_length_for_non_zeros)), out.ctypes.data, cap)
      if nb >= 0:
        return out[:nb].tobytes()
      cap *= 8

  def run_length_decode(self, code: bytes, shape, run_length_code=-1, magnitude_code=-1,
                        use_run_length_for_non_zeros=False) -> np.ndarray:
    """RunLengthDecode (cc/kernels/run_length_kernels.cc:141-258); raises OracleError with the reference's messages."""
    n = int(np.prod(shape))
    buf = np.frombuffer(bytes(code) + b"\0", np.uint8)
    out = np.zeros(max(n, 1), np.int32)
    if not hasattr(self, "_run_length_decode"):
      self._fn("run_length_decode", C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64])
    e = self._run_length_decode(buf.ctypes.data, len(code), int(run_length_code), int(magnitude_code),
                                int(bool(use_run_length_for_non_zeros)), out.ctypes.data, n)
    if e:
      raise OracleError({1: "Out of bits to read.", 2: "Exceeded maximum gamma bit width.", 3: "Decoded past end of tensor."}[e])
    return out[:n].reshape(shape)

  def pmf_to_cdf(self, pmf, precision: int) -> np.ndarray:
    pmf = np.ascontiguousarray(pmf, dtype=np.float32)
    n = pmf.shape[-1]
    rows = pmf.size // max(n, 1)
    cdf = np.empty(pmf.shape[:-1] + (n + 1,), dtype=np.int32)
    rc = self._pmf_to_cdf(_p(pmf, C.c_float), rows, n, precision, _p(cdf, C.c_int32))
    if rc != 0:
      raise OracleError(self._err())
    return cdf


def _lookup_args(lookup):
  lookup = _i32(lookup)
  if lookup.ndim == 1:
    return lookup, 0
  if lookup.ndim == 2:
    return lookup, lookup.shape[1]
  raise OracleError("`lookup` must be rank 1 or 2")


class OracleEncoder:

  def __init__(self, o: Oracle, lookup, n_streams: int):
    self._o = o
    lookup, cols = _lookup_args(lookup)
    self._lookup = lookup
    self.n_streams = int(n_streams)
    self._h = o._encoder_create(_p(lookup, C.c_int32), lookup.size, cols, self.n_streams)
    if not self._h:
      raise OracleError(o._err())

  def encode(self, value, index=None, threads: int = 1) -> None:
    value = _i32(value).reshape(self.n_streams, -1)
    ip = None
    if index is not None:
      index = _i32(index).reshape(self.n_streams, -1)
      assert index.shape == value.shape
      ip = _p(index, C.c_int32)
    rc = self._o._encoder_encode(self._h, ip, _p(value, C.c_int32), value.shape[1], threads)
    if rc != 0:
      raise OracleError(self._o._err())

  def finalize(self) -> List[bytes]:
    offs = np.zeros(self.n_streams + 1, dtype=np.int64)
    self._o._encoder_finalize(self._h, _p(offs, C.c_int64))
    buf = np.empty(max(int(offs[-1]), 1), dtype=np.uint8)
    self._o._encoder_bytes(self._h, _p(buf, C.c_uint8))
    return [buf[offs[i]:offs[i + 1]].tobytes() for i in range(self.n_streams)]

  def close(self):
    if self._h:
      self._o._encoder_free(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint:disable=broad-except
      pass


class OracleDecoder:

  def __init__(self, o: Oracle, strings: Sequence[bytes], lookup):
    self._o = o
    lookup, cols = _lookup_args(lookup)
    self.n_streams = len(strings)
    offs = np.zeros(self.n_streams + 1, dtype=np.int64)
    for i, s in enumerate(strings):
      offs[i + 1] = offs[i] + len(s)
    buf = np.frombuffer(b"".join(strings) + b"\0", dtype=np.uint8).copy()
    self._h = o._decoder_create(
        _p(buf, C.c_uint8), _p(offs, C.c_int64), self.n_streams, _p(lookup, C.c_int32), lookup.size,
        cols)
    if not self._h:
      raise OracleError(o._err())

  def decode(self, n_per_stream: int, index=None, threads: int = 1) -> np.ndarray:
    out = np.empty((self.n_streams, int(n_per_stream)), dtype=np.int32)
    ip = None
    if index is not None:
      index = _i32(index).reshape(self.n_streams, -1)
      assert index.shape == out.shape
      ip = _p(index, C.c_int32)
    rc = self._o._decoder_decode(self._h, ip, _p(out, C.c_int32), int(n_per_stream), threads)
    if rc != 0:
      raise OracleError(self._o._err())
    return out

  def finalize(self) -> np.ndarray:
    ok = np.zeros(max(self.n_streams, 1), dtype=np.uint8)
    self._o._decoder_finalize(self._h, _p(ok, C.c_uint8))
    return ok[:self.n_streams].astype(bool)

  def close(self):
    if self._h:
      self._o._decoder_free(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint:disable=broad-except
      pass


_cache = {}


def port() -> Oracle:
  if "port" not in _cache:
    if not os.path.exists(_PORT_SO):
      build(want_ref=False)
    _cache["port"] = Oracle(_PORT_SO, "tfcport_", "port")
  return _cache["port"]


def have_ref() -> bool:
  return os.path.exists(_REF_SO)


def ref() -> Oracle:
  if "ref" not in _cache:
    if not have_ref():
      raise FileNotFoundError(
          f"{_REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
    _cache["ref"] = Oracle(_REF_SO, "tfcref_", "reference")
  return _cache["ref"]


def best() -> Oracle:
  """The compiled reference when present, else the C port."""
  return ref() if have_ref() else port()
