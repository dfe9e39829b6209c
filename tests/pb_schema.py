"""The public TensorFlow message schemas the readers in this repository restate by hand -- example.proto / feature.proto
(tf.train.SequenceExample) and tensor_bundle.proto / tensor_shape.proto / versions.proto (V2 checkpoints) -- declared to
Google's protobuf runtime, so that an INDEPENDENT implementation of the wire format encodes and decodes the test
messages.  Field numbers and types as published in tensorflow/core/example/*.proto and
tensorflow/core/protobuf/tensor_bundle.proto (TensorFlow itself is not installed here)."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=F.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    if packed is not None:
        f.options.packed = packed
    if oneof is not None:
        f.oneof_index = oneof
    return f


def _map_entry(parent, name, value_type_name):
    e = parent.nested_type.add()
    e.name = name
    e.options.map_entry = True
    _field(e, "key", 1, F.TYPE_STRING)
    _field(e, "value", 2, F.TYPE_MESSAGE, type_name=value_type_name)
    return e


def build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "tf_public_schemas.proto", "tfpub", "proto3"
    m = fd.message_type.add(); m.name = "BytesList"
    _field(m, "value", 1, F.TYPE_BYTES, F.LABEL_REPEATED)
    m = fd.message_type.add(); m.name = "FloatList"
    _field(m, "value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=True)
    m = fd.message_type.add(); m.name = "Int64List"
    _field(m, "value", 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=True)
    m = fd.message_type.add(); m.name = "Feature"
    m.oneof_decl.add().name = "kind"
    _field(m, "bytes_list", 1, F.TYPE_MESSAGE, type_name=".tfpub.BytesList", oneof=0)
    _field(m, "float_list", 2, F.TYPE_MESSAGE, type_name=".tfpub.FloatList", oneof=0)
    _field(m, "int64_list", 3, F.TYPE_MESSAGE, type_name=".tfpub.Int64List", oneof=0)
    m = fd.message_type.add(); m.name = "Features"
    _map_entry(m, "FeatureEntry", ".tfpub.Feature")
    _field(m, "feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".tfpub.Features.FeatureEntry")
    m = fd.message_type.add(); m.name = "FeatureList"
    _field(m, "feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".tfpub.Feature")
    m = fd.message_type.add(); m.name = "FeatureLists"
    _map_entry(m, "FeatureListEntry", ".tfpub.FeatureList")
    _field(m, "feature_list", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".tfpub.FeatureLists.FeatureListEntry")
    m = fd.message_type.add(); m.name = "SequenceExample"
    _field(m, "context", 1, F.TYPE_MESSAGE, type_name=".tfpub.Features")
    _field(m, "feature_lists", 2, F.TYPE_MESSAGE, type_name=".tfpub.FeatureLists")
    # ---- checkpoints
    m = fd.message_type.add(); m.name = "TensorShapeProto"
    d = m.nested_type.add(); d.name = "Dim"
    _field(d, "size", 1, F.TYPE_INT64)
    _field(d, "name", 2, F.TYPE_STRING)
    _field(m, "dim", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name=".tfpub.TensorShapeProto.Dim")
    _field(m, "unknown_rank", 3, F.TYPE_BOOL)
    m = fd.message_type.add(); m.name = "VersionDef"
    _field(m, "producer", 1, F.TYPE_INT32)
    _field(m, "min_consumer", 2, F.TYPE_INT32)
    _field(m, "bad_consumers", 3, F.TYPE_INT32, F.LABEL_REPEATED, packed=True)
    m = fd.message_type.add(); m.name = "BundleHeaderProto"
    _field(m, "num_shards", 1, F.TYPE_INT32)
    _field(m, "endianness", 2, F.TYPE_INT32)            # enum {LITTLE = 0, BIG = 1}: a varint on the wire
    _field(m, "version", 3, F.TYPE_MESSAGE, type_name=".tfpub.VersionDef")
    m = fd.message_type.add(); m.name = "BundleEntryProto"
    _field(m, "dtype", 1, F.TYPE_INT32)                 # enum DataType: a varint on the wire (DT_FLOAT = 1, DT_INT64 = 9, ...)
    _field(m, "shape", 2, F.TYPE_MESSAGE, type_name=".tfpub.TensorShapeProto")
    _field(m, "shard_id", 3, F.TYPE_INT32)
    _field(m, "offset", 4, F.TYPE_INT64)
    _field(m, "size", 5, F.TYPE_INT64)
    _field(m, "crc32c", 6, F.TYPE_FIXED32)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda name: message_factory.GetMessageClass(pool.FindMessageTypeByName("tfpub." + name))
    return {n: get(n) for n in ("SequenceExample", "Feature", "BundleEntryProto", "BundleHeaderProto")}
