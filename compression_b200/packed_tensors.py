"""PackedTensors: the reference's container for compressed tensors (tensorflow_compression/python/util/
packed_tensors.py:25-100), i.e. a serialised `tf.train.Example` whose features are named chr(1), chr(2), ... plus
an optional "MD" model identifier.  TensorFlow is not available here, so the protobuf wire format of
tensorflow/core/example/{example,feature}.proto is written and parsed directly:

  Example   { Features features = 1; }
  Features  { map<string, Feature> feature = 1; }          (map entry: key = 1, value = 2)
  Feature   { oneof kind { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; } }
  BytesList { repeated bytes value = 1; }
  FloatList { repeated float value = 1 [packed = true]; }
  Int64List { repeated int64 value = 1 [packed = true]; }

Files written here parse with the reference (`tfci.py`, `bls2017.py decompress`), and files written by the
reference parse here (packed and unpacked repeated encodings, any feature order).
"""
import struct

import numpy as np
import torch

__all__ = ["PackedTensors"]

_BYTES, _FLOAT, _INT64 = 1, 2, 3  # field numbers of the Feature oneof


# ---------------------------------------------------------------------------------------------
# protobuf wire format
# ---------------------------------------------------------------------------------------------
def _varint(value):
  value &= (1 << 64) - 1  # int64: negative numbers are ten-byte two's complement varints
  out = bytearray()
  while True:
    byte = value & 0x7F
    value >>= 7
    if value:
      out.append(byte | 0x80)
    else:
      out.append(byte)
      return bytes(out)


def _read_varint(buf, pos):
  result, shift = 0, 0
  while True:
    if pos >= len(buf):
      raise ValueError("Truncated varint.")
    byte = buf[pos]
    pos += 1
    result |= (byte & 0x7F) << shift
    if not byte & 0x80:
      return result, pos
    shift += 7
    if shift > 63:
      raise ValueError("Varint too long.")


def _len_field(number, payload):
  return _varint((number << 3) | 2) + _varint(len(payload)) + payload


def _fields(buf):
  """Yields (field number, wire type, value) for every field of a message."""
  pos = 0
  while pos < len(buf):
    key, pos = _read_varint(buf, pos)
    number, wire = key >> 3, key & 7
    if wire == 0:
      value, pos = _read_varint(buf, pos)
    elif wire == 1:
      value, pos = buf[pos:pos + 8], pos + 8
    elif wire == 2:
      size, pos = _read_varint(buf, pos)
      value, pos = buf[pos:pos + size], pos + size
      if len(value) != size:
        raise ValueError("Truncated length-delimited field.")
    elif wire == 5:
      value, pos = buf[pos:pos + 4], pos + 4
    else:
      raise ValueError(f"Unsupported wire type {wire}.")
    if pos > len(buf):
      raise ValueError("Truncated message.")
    yield number, wire, value


def _signed64(value):
  return value - (1 << 64) if value >= (1 << 63) else value


def _encode_feature(kind, values):
  if kind == _BYTES:
    inner = b"".join(_len_field(1, bytes(v)) for v in values)
  elif kind == _FLOAT:
    inner = _len_field(1, struct.pack(f"<{len(values)}f", *values)) if len(values) else b""
  else:
    inner = _len_field(1, b"".join(_varint(int(v)) for v in values)) if len(values) else b""
  return _len_field(kind, inner)


def _decode_feature(buf):
  kind, values = None, []
  for number, wire, value in _fields(buf):
    if number not in (_BYTES, _FLOAT, _INT64) or wire != 2:
      continue
    kind, values = number, []  # oneof: the last one on the wire wins
    for n2, w2, v2 in _fields(value):
      if n2 != 1:
        continue
      if number == _BYTES:
        values.append(bytes(v2))
      elif number == _FLOAT:
        if w2 == 2:  # packed
          values.extend(struct.unpack(f"<{len(v2) // 4}f", bytes(v2)))
        else:
          values.append(struct.unpack("<f", bytes(v2))[0])
      else:
        if w2 == 2:  # packed
          p = 0
          while p < len(v2):
            x, p = _read_varint(v2, p)
            values.append(_signed64(x))
        else:
          values.append(_signed64(v2))
  return kind, values


# ---------------------------------------------------------------------------------------------
class PackedTensors:
  """Packs several rank-1 tensors (integer, floating point or byte strings) into one string, optionally with a
  model identifier.  Same interface as the reference class."""

  def __init__(self, string=None):
    self._features = {}  # name -> (kind, list of values)
    if string:
      self.string = string

  # -- model identifier ("MD") --
  @property
  def model(self):
    return self._features["MD"][1][0].decode("ascii")

  @model.setter
  def model(self, value):
    self._features["MD"] = (_BYTES, [value.encode("ascii")])

  @model.deleter
  def model(self):
    del self._features["MD"]

  # -- serialised form --
  @property
  def string(self):
    entries = b"".join(
        _len_field(1, _len_field(1, name.encode("utf-8")) + _len_field(2, _encode_feature(kind, values)))
        for name, (kind, values) in sorted(self._features.items()))
    return _len_field(1, entries) if self._features else b""

  @string.setter
  def string(self, value):
    features = {}
    for number, wire, body in _fields(bytes(value)):
      if number != 1 or wire != 2:
        continue
      for n2, w2, entry in _fields(body):  # Features.feature map entries
        if n2 != 1 or w2 != 2:
          continue
        name, feature = "", (None, [])
        for n3, w3, v3 in _fields(entry):
          if n3 == 1 and w3 == 2:
            name = bytes(v3).decode("utf-8")
          elif n3 == 2 and w3 == 2:
            feature = _decode_feature(v3)
        features[name] = feature
    self._features = features

  # -- tensors --
  def pack(self, tensors):
    """Packs rank-1 values: torch / numpy integer or floating arrays, or byte strings (`gen_ops.Strings`, a
    sequence of `bytes`, or a numpy object / bytes array)."""
    i = 1
    for tensor in tensors:
      kind, values = self._classify(tensor)
      self._features[chr(i)] = (kind, values)
      i += 1
    while chr(i) in self._features:  # delete any remaining, previously set arrays
      del self._features[chr(i)]
      i += 1

  @staticmethod
  def _classify(tensor):
    if hasattr(tensor, "tolist") and hasattr(tensor, "bytes_dev"):  # gen_ops.Strings
      if len(tensor.shape) != 1:
        raise RuntimeError(f"Unexpected tensor rank: {len(tensor.shape)}.")
      return _BYTES, [bytes(b) for b in tensor.tolist()]
    if isinstance(tensor, (bytes, bytearray)):
      raise RuntimeError("Unexpected tensor rank: 0.")
    if isinstance(tensor, (list, tuple)) and all(isinstance(b, (bytes, bytearray)) for b in tensor):
      return _BYTES, [bytes(b) for b in tensor]
    if isinstance(tensor, torch.Tensor):
      array = tensor.detach().cpu().numpy()
    else:
      array = np.asarray(tensor)
    if array.ndim != 1:
      raise RuntimeError(f"Unexpected tensor rank: {array.ndim}.")
    if array.dtype.kind in "iu" or array.dtype == np.bool_:
      return _INT64, [int(v) for v in array]
    if array.dtype.kind == "f":
      return _FLOAT, [float(v) for v in array.astype(np.float32)]
    if array.dtype.kind in "SO":
      return _BYTES, [bytes(v) for v in array]
    raise RuntimeError(f"Unexpected tensor dtype: '{array.dtype}'.")

  def unpack(self, dtypes):
    """Unpacks values based on `dtypes` (torch / numpy dtypes; `bytes`, `str`, "string" or `object` for byte
    strings).  Numeric features come back as torch tensors, byte strings as a list of `bytes`."""
    tensors = []
    for i, dtype in enumerate(dtypes):
      kind, values = self._features.get(chr(i + 1), (None, []))
      if dtype in (bytes, str, object, "string", "bytes") or (isinstance(dtype, np.dtype) and dtype.kind in "SO"):
        tensors.append(list(values) if kind in (_BYTES, None) else [])
        continue
      tdtype = dtype if isinstance(dtype, torch.dtype) else torch.from_numpy(np.zeros(0, dtype=dtype)).dtype
      if tdtype.is_floating_point:
        tensors.append(torch.tensor(values if kind == _FLOAT else [], dtype=tdtype))
      elif tdtype in (torch.int8, torch.uint8, torch.int16, torch.int32, torch.int64, torch.bool):
        tensors.append(torch.tensor(values if kind == _INT64 else [], dtype=tdtype))
      else:
        raise RuntimeError(f"Unexpected dtype: '{dtype}'.")
    return tensors
