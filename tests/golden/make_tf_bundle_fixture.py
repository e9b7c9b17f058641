"""Builds tests/golden/tf_bundle/: a TensorFlow checkpoint-v2 bundle written INDEPENDENTLY of
matryodshka_b200/tf_checkpoint.py, to pin that reader (row f3: checkpoint ingest, reference test.py:192-202).

No TensorFlow binary exists in this sandbox and none of its checkpoints are on disk, so this script assembles the
files from the published formats with as much third-party code as the image offers:

* the protocol buffers (BundleHeaderProto / BundleEntryProto of tensorflow/core/protobuf/tensor_bundle.proto) are
  declared from the published .proto text through google.protobuf descriptors and serialized by the OFFICIAL protobuf
  runtime; TensorShapeProto, DataType and VersionDef are the generated classes TensorBoard ships
  (tensorboard.compat.proto.{tensor_shape,types,versions}_pb2) -- nothing of tf_checkpoint.py's hand-rolled wire
  encoder is used;
* crc32c and its LevelDB masking are TensorBoard's (tensorboard.compat.tensorflow_stub.pywrap_tensorflow.masked_crc32c,
  written by the TensorFlow authors for TFRecord);
* the table container (.index) follows leveldb/doc/table_format.md and table_builder.cc as TensorFlow's
  tensorflow/core/lib/io/table_builder.cc uses them: data blocks with prefix compression and a restart point every 16
  entries, flushed at `block_size`; an index block (restart interval 1) whose keys are the SHORTENED separators
  (FindShortestSeparator / FindShortSuccessor), an empty metaindex block, 5-byte block trailers (type 0 = no
  compression, masked crc32c of block + type), 48-byte footer with the magic number.  Two bundles are written: one with
  TensorFlow's table block size (256 KB: one data block) and one with 512-byte blocks (many blocks, so index
  separators and restart arrays are exercised);
* the tensors mimic what the reference's Saver writes: `net/...` variables, their Adam slots (`.../Adam`,
  `.../Adam_1`: long shared key prefixes), `beta1_power`, an int64 `global_step`, float16 / uint8 / bool odd ones.

Run from the repo root:  python tests/golden/make_tf_bundle_fixture.py  (rewrites the committed files bit for bit).
"""
import os
import struct

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2  # noqa: F401  (registers the imports)
from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import masked_crc32c

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "tf_bundle")
MAGIC = 0xDB4775248B80FB57
T = descriptor_pb2.FieldDescriptorProto


def bundle_protos():
    """BundleHeaderProto / BundleEntryProto from the published tensor_bundle.proto (TensorFlow 1.14)."""
    f = descriptor_pb2.FileDescriptorProto()
    f.name = "msi_fixture/tensor_bundle.proto"
    f.package = "tensorboard"
    f.syntax = "proto3"
    f.dependency.extend(["tensorboard/compat/proto/tensor_shape.proto", "tensorboard/compat/proto/types.proto",
                         "tensorboard/compat/proto/versions.proto"])
    h = f.message_type.add()
    h.name = "BundleHeaderProto"
    en = h.enum_type.add()
    en.name = "Endianness"
    for n, v in (("LITTLE", 0), ("BIG", 1)):
        ev = en.value.add()
        ev.name, ev.number = n, v
    for name, num, typ, tn in (("num_shards", 1, T.TYPE_INT32, None),
                               ("endianness", 2, T.TYPE_ENUM, ".tensorboard.BundleHeaderProto.Endianness"),
                               ("version", 3, T.TYPE_MESSAGE, ".tensorboard.VersionDef")):
        fd = h.field.add()
        fd.name, fd.number, fd.type, fd.label = name, num, typ, T.LABEL_OPTIONAL
        if tn:
            fd.type_name = tn
    e = f.message_type.add()
    e.name = "BundleEntryProto"
    for name, num, typ, tn in (("dtype", 1, T.TYPE_ENUM, ".tensorboard.DataType"),
                               ("shape", 2, T.TYPE_MESSAGE, ".tensorboard.TensorShapeProto"),
                               ("shard_id", 3, T.TYPE_INT32, None), ("offset", 4, T.TYPE_INT64, None),
                               ("size", 5, T.TYPE_INT64, None), ("crc32c", 6, T.TYPE_FIXED32, None)):
        fd = e.field.add()
        fd.name, fd.number, fd.type, fd.label = name, num, typ, T.LABEL_OPTIONAL
        if tn:
            fd.type_name = tn
    pool = descriptor_pool.Default()
    try:
        pool.Add(f)
    except TypeError:   # the file is already in the pool (script imported twice)
        pass
    get = getattr(message_factory, "GetMessageClass", None)
    mk = (lambda d: get(d)) if get else (lambda d: message_factory.MessageFactory(pool).GetPrototype(d))
    return (mk(pool.FindMessageTypeByName("tensorboard.BundleHeaderProto")),
            mk(pool.FindMessageTypeByName("tensorboard.BundleEntryProto")))


DT = {np.dtype(np.float32): types_pb2.DT_FLOAT, np.dtype(np.int64): types_pb2.DT_INT64, np.dtype(np.int32): types_pb2.DT_INT32,
      np.dtype(np.float16): types_pb2.DT_HALF, np.dtype(np.uint8): types_pb2.DT_UINT8, np.dtype(np.bool_): types_pb2.DT_BOOL,
      np.dtype(np.float64): types_pb2.DT_DOUBLE}


def fixture_tensors():
    rng = np.random.default_rng(20191203)
    t = {}
    for scope, shp in (("net/conv1_1", (3, 3, 3, 4)), ("net/conv1_2", (3, 3, 5, 8)), ("net/conv6_1", (4, 4, 4, 8)),
                       ("net/color_pred", (1, 1, 4, 6))):
        w = rng.normal(0, 0.1, shp).astype(np.float32)
        t[scope + "/weights"] = w
        t[scope + "/weights/Adam"] = (w * 0.01).astype(np.float32)
        t[scope + "/weights/Adam_1"] = (w * w).astype(np.float32)
        if scope != "net/color_pred":
            for p in ("gamma", "beta"):
                v = rng.uniform(0.5, 1.5, (shp[3] if "conv6" not in scope else shp[2],)).astype(np.float32)
                t[f"{scope}/LayerNorm/{p}"] = v
                t[f"{scope}/LayerNorm/{p}/Adam"] = (v * 0.5).astype(np.float32)
                t[f"{scope}/LayerNorm/{p}/Adam_1"] = (v * 0.25).astype(np.float32)
        else:
            t[scope + "/biases"] = rng.normal(0, 0.1, (shp[3],)).astype(np.float32)
    t["beta1_power"] = np.float32(0.9 ** 410000)
    t["beta2_power"] = np.float32(0.0)
    t["global_step"] = np.int64(410000)
    t["misc/half"] = rng.normal(0, 1, (5, 3)).astype(np.float16)
    t["misc/bytes"] = rng.integers(0, 256, (4, 2, 2), dtype=np.uint8)
    t["misc/flags"] = np.array([True, False, True])
    t["misc/empty"] = np.zeros((0, 4), np.float32)
    t["misc/f64"] = np.array([[np.pi, -np.e]], np.float64)
    return t


# ---- LevelDB table (leveldb/doc/table_format.md, table/{block,table}_builder.cc) ---------------------------------
def varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


class BlockBuilder:
    def __init__(self, restart_interval):
        self.buf, self.restarts, self.counter, self.last = bytearray(), [0], 0, b""
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.counter < self.interval:
            n = min(len(self.last), len(key))
            while shared < n and self.last[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last = key
        self.counter += 1

    def size_estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))

    def empty(self):
        return not self.buf


def shortest_separator(start, limit):
    """BytewiseComparator::FindShortestSeparator."""
    n = min(len(start), len(limit))
    d = 0
    while d < n and start[d] == limit[d]:
        d += 1
    if d >= n:
        return start
    b = start[d]
    if b < 0xFF and b + 1 < limit[d]:
        return start[:d] + bytes([b + 1])
    return start


def short_successor(key):
    """BytewiseComparator::FindShortSuccessor."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def write_table(path, items, block_size):
    out = bytearray()

    def write_block(content):
        handle = varint(len(out)) + varint(len(content))
        trailer_type = b"\x00"   # kNoCompression (tensor_bundle.cc)
        out.extend(content + trailer_type + struct.pack("<I", masked_crc32c(content + trailer_type)))
        return handle

    data, index = BlockBuilder(16), BlockBuilder(1)
    pending = None   # (last key of the flushed block, its handle): the index entry waits for the next key
    for key, value in items:
        if pending is not None:
            index.add(shortest_separator(pending[0], key), pending[1])
            pending = None
        data.add(key, value)
        if data.size_estimate() >= block_size:
            pending = (key, write_block(data.finish()))
            data = BlockBuilder(16)
    if not data.empty():
        pending = (data.last, write_block(data.finish()))
    if pending is not None:
        index.add(short_successor(pending[0]), pending[1])
    meta_handle = write_block(BlockBuilder(16).finish())     # metaindex block: no filter policy -> empty
    index_handle = write_block(index.finish())
    footer = meta_handle + index_handle
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))
    with open(path, "wb") as fh:
        fh.write(bytes(out))


def write_bundle(prefix, tensors, block_size):
    Header, Entry = bundle_protos()
    items = [(b"", Header(num_shards=1, endianness=0, version=versions_pb2.VersionDef(producer=1)).SerializeToString())]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[name])
            raw = a.tobytes(order="C")
            shape = tensor_shape_pb2.TensorShapeProto(dim=[tensor_shape_pb2.TensorShapeProto.Dim(size=int(d)) for d in a.shape])
            e = Entry(dtype=DT[a.dtype], shape=shape, shard_id=0, offset=offset, size=len(raw), crc32c=masked_crc32c(raw))
            items.append((name.encode(), e.SerializeToString()))
            fh.write(raw)
            offset += len(raw)
    write_table(prefix + ".index", items, block_size)


def main():
    os.makedirs(OUT, exist_ok=True)
    tensors = fixture_tensors()
    write_bundle(os.path.join(OUT, "model.ckpt-410000"), tensors, block_size=262144)     # TensorFlow's table block size
    write_bundle(os.path.join(OUT, "small_blocks.ckpt"), tensors, block_size=512)
    with open(os.path.join(OUT, "checkpoint"), "w") as fh:   # what tf.train.Saver leaves beside the bundle
        fh.write('model_checkpoint_path: "model.ckpt-410000"\nall_model_checkpoint_paths: "model.ckpt-410000"\n')
    # (the expected values are fixture_tensors() itself: the test regenerates them from the seed)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
