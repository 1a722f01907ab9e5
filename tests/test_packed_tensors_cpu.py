"""CPU: the `.tfci` container (py/util/packed_tensors.py:25-100).  TensorFlow is absent, so the wire format is
checked against the real protobuf runtime with the public schema of tensorflow/core/example/{example,feature}.proto
built from descriptors: what the reference writes must parse here, and what is written here must parse there."""
import numpy as np
import pytest
import torch
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

from compression_b200.packed_tensors import PackedTensors


def _example_class():
  f = descriptor_pb2.FileDescriptorProto(name="tfcb_example_test.proto", package="tensorflow", syntax="proto3")
  T = descriptor_pb2.FieldDescriptorProto

  def msg(name):
    m = f.message_type.add()
    m.name = name
    return m

  def field(m, name, number, ftype, label=T.LABEL_OPTIONAL, type_name=None, oneof=None):
    fd = m.field.add(name=name, number=number, type=ftype, label=label)
    if type_name:
      fd.type_name = type_name
    if oneof is not None:
      fd.oneof_index = oneof
    return fd

  field(msg("BytesList"), "value", 1, T.TYPE_BYTES, T.LABEL_REPEATED)
  field(msg("FloatList"), "value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED)
  field(msg("Int64List"), "value", 1, T.TYPE_INT64, T.LABEL_REPEATED)
  feat = msg("Feature")
  feat.oneof_decl.add(name="kind")
  field(feat, "bytes_list", 1, T.TYPE_MESSAGE, type_name=".tensorflow.BytesList", oneof=0)
  field(feat, "float_list", 2, T.TYPE_MESSAGE, type_name=".tensorflow.FloatList", oneof=0)
  field(feat, "int64_list", 3, T.TYPE_MESSAGE, type_name=".tensorflow.Int64List", oneof=0)
  feats = msg("Features")
  entry = feats.nested_type.add(name="FeatureEntry")
  entry.options.map_entry = True
  field(entry, "key", 1, T.TYPE_STRING)
  field(entry, "value", 2, T.TYPE_MESSAGE, type_name=".tensorflow.Feature")
  field(feats, "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, type_name=".tensorflow.Features.FeatureEntry")
  field(msg("Example"), "features", 1, T.TYPE_MESSAGE, type_name=".tensorflow.Features")
  pool = descriptor_pool.DescriptorPool()
  pool.Add(f)
  return message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow.Example"))


Example = _example_class()


def test_pack_unpack_identity():
  # packed_tensors_test.py:24-34
  packed = PackedTensors()
  packed.pack([[b"xyz"], torch.tensor([1, 3], dtype=torch.int32)])
  packed = PackedTensors(packed.string)
  string, shape = packed.unpack([bytes, torch.int32])
  assert string == [b"xyz"] and shape.tolist() == [1, 3] and shape.dtype == torch.int32


def test_set_get_model_identity():
  # packed_tensors_test.py:36-42
  packed = PackedTensors()
  packed.model = "xyz"
  packed = PackedTensors(packed.string)
  assert packed.model == "xyz"
  del packed.model
  with pytest.raises(KeyError):
    _ = packed.model


def test_what_we_write_parses_with_protobuf():
  strings = [bytes([0, 255, 7]), b"", b"abc" * 100]
  shape = np.array([768, 512, -5, 2**40], dtype=np.int64)
  floats = torch.tensor([0.5, -1.25, 3.0e-8])
  packed = PackedTensors()
  packed.model = "bmshj2018-factorized-mse-3"
  packed.pack([strings, shape, floats])
  ex = Example()
  ex.ParseFromString(packed.string)
  f = ex.features.feature
  assert sorted(f.keys()) == [chr(1), chr(2), chr(3), "MD"]
  assert list(f[chr(1)].bytes_list.value) == strings
  assert list(f[chr(2)].int64_list.value) == shape.tolist()
  assert list(f[chr(3)].float_list.value) == pytest.approx(floats.tolist())
  assert f["MD"].bytes_list.value[0] == b"bmshj2018-factorized-mse-3"
  # byte for byte what protobuf's deterministic serialisation produces
  assert packed.string == ex.SerializeToString(deterministic=True)


def test_what_protobuf_writes_parses_here():
  ex = Example()
  ex.features.feature[chr(2)].int64_list.value[:] = [16, 16, 0, -1]
  ex.features.feature[chr(1)].bytes_list.value[:] = [b"\x01\x02", b"tail"]
  ex.features.feature["MD"].bytes_list.value[:] = [b"hific-lo"]
  ex.features.feature[chr(3)].float_list.value[:] = [1.5, 2.5]
  packed = PackedTensors(ex.SerializeToString())
  assert packed.model == "hific-lo"
  strings, shape, fl = packed.unpack(["string", np.int32, torch.float32])
  assert strings == [b"\x01\x02", b"tail"] and shape.tolist() == [16, 16, 0, -1] and fl.tolist() == [1.5, 2.5]
  # unpacked repeated encodings (old writers) are accepted too: int64 as separate varints, floats as fixed32
  raw = b"\x08\x05\x08\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01"                 # Int64List {5, -1}, unpacked
  feature = b"\x1a" + bytes([len(raw)]) + raw
  entry = b"\x0a\x01\x01" + b"\x12" + bytes([len(feature)]) + feature
  features = b"\x0a" + bytes([len(entry)]) + entry
  blob = b"\x0a" + bytes([len(features)]) + features
  assert PackedTensors(blob).unpack([torch.int64])[0].tolist() == [5, -1]


def test_repacking_drops_stale_features_and_rejects_bad_ranks():
  packed = PackedTensors()
  packed.pack([[b"a"], np.arange(3), np.arange(2.0)])
  packed.pack([[b"b"]])
  ex = Example()
  ex.ParseFromString(packed.string)
  assert list(ex.features.feature.keys()) == [chr(1)]
  with pytest.raises(RuntimeError):
    packed.pack([np.zeros((2, 2), np.int32)])
  with pytest.raises(RuntimeError):
    packed.pack([np.zeros(2, np.complex64)])
  with pytest.raises(ValueError):
    PackedTensors(b"\x0a\x05\x0a")  # truncated


def test_packs_a_strings_object_like_the_models_do():
  # models/bls2017.py:281-284: packed.pack((string, x_shape, y_shape)) with `string` the entropy model's output.
  class FakeStrings:  # the attributes of gen_ops.Strings that pack() relies on (no CUDA needed)
    shape = (2,)
    bytes_dev = object()

    def tolist(self):
      return [b"\x10\x20", b"\x30"]

  packed = PackedTensors()
  packed.pack((FakeStrings(), torch.tensor([512, 768]), torch.tensor([32, 48])))
  strings, x_shape, y_shape = PackedTensors(packed.string).unpack([bytes, torch.int32, torch.int32])
  assert strings == [b"\x10\x20", b"\x30"] and x_shape.tolist() == [512, 768] and y_shape.tolist() == [32, 48]
  FakeStrings.shape = (2, 1)
  with pytest.raises(RuntimeError):
    PackedTensors().pack([FakeStrings()])
