"""Pure-Python reader (and minimal writer) of TensorFlow checkpoint-v2 "tensor bundles".

The reference restores its weights with ``tf.train.Saver.restore(sess, tf.train.latest_checkpoint(dir))``
(test.py:192-202); TensorFlow is not available here, so the bundle is parsed directly and turned into
the ``{variable name: ndarray}`` dict that ``NetEngine`` / ``MSI`` take (``net/conv1_1/weights``,
``net/conv1_1/LayerNorm/gamma`` ... -- the names ARE the bundle keys).

Format [TensorFlow 1.14, tensorflow/core/util/tensor_bundle + tensorflow/core/lib/io/table, restated
from the published format].  Parity: no TensorFlow binary and no TensorFlow-written checkpoint exist in this
sandbox; the reader is pinned to tests/golden/tf_bundle/, a bundle assembled INDEPENDENTLY of this module
(tests/golden/make_tf_bundle_fixture.py: protocol buffers serialized by the official protobuf runtime from the
published tensor_bundle.proto and TensorBoard's generated TensorShapeProto / DataType / VersionDef, crc32c and its
masking from TensorBoard's TFRecord code, the table container written after leveldb's table_format.md with
TensorFlow's options), and to this module's own writer:

* ``<prefix>.index``  a LevelDB-format sorted table, written WITHOUT compression
  (tensor_bundle.cc: ``options.compression = table::kNoCompression``):
  ``[data blocks][metaindex block][index block][footer]``; footer = 48 bytes = two block handles
  (varint64 offset, varint64 size) padded to 40 bytes + magic 0xdb4775248b80fb57 (little endian);
  every block is followed by a 5-byte trailer (compression type, masked crc32c); a block is a run of
  prefix-compressed entries ``varint32 shared, varint32 non_shared, varint32 value_len, key delta,
  value`` followed by the restart array and its uint32 length.  The index block maps separator
  keys to data-block handles.
  Key ``""`` -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}; every other key is a
  tensor name -> BundleEntryProto {1: dtype, 2: TensorShapeProto, 3: shard_id, 4: offset, 5: size,
  6: crc32c (fixed32, masked), 7: slices}.
* ``<prefix>.data-SSSSS-of-NNNNN``  raw little-endian tensor bytes at [offset, offset + size).
* ``checkpoint``  text proto; ``model_checkpoint_path: "<name>"`` is what latest_checkpoint returns.
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---- varints / protobuf wire format ----------------------------------------------------------------
def _varint(buf, pos) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf):
    """Yields (field number, wire type, value) of one serialized message (value: int or bytes)."""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


# ---- crc32c (Castagnoli), masked the LevelDB way -----------------------------------------------------
def _make_crc_table():
    poly = 0x82F63B78
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        t.append(c)
    return t


_CRC_TABLE = _make_crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    t = _CRC_TABLE
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- LevelDB table ----------------------------------------------------------------------------------
def _block_entries(block: bytes):
    """(key, value) pairs of one table block (prefix-compressed keys; the restart array is skipped)."""
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if ctype != 0:
        raise NotImplementedError("compressed table block (type %d); TensorFlow writes bundles uncompressed" % ctype)
    if verify:
        want = struct.unpack_from("<I", data, offset + size + 1)[0]
        if masked_crc32c(data[offset:offset + size + 1]) != want:
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    return block


def read_table(path: str, verify: bool = True) -> Dict[bytes, bytes]:
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _MAGIC:
        raise ValueError("%s is not a LevelDB table (bad magic)" % path)
    footer = data[-48:]
    pos = 0
    _, pos = _varint(footer, pos)      # metaindex handle (unused: TensorFlow writes no filter block)
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    out = {}
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for k, v in _block_entries(_read_block(data, off, size, verify)):
            out[k] = v
    return out


# ---- tensor bundle ------------------------------------------------------------------------------------
def _parse_shape(buf) -> Tuple[int, ...]:
    dims = []
    for f, _, v in _proto_fields(buf):
        if f == 2:  # Dim
            size = 0
            for f2, _, v2 in _proto_fields(v):
                if f2 == 1:
                    size = _signed64(v2)
            dims.append(size)
        elif f == 3 and v:
            raise NotImplementedError("tensor of unknown rank")
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for f, _, v in _proto_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _parse_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["slices"] += 1
    return e


def list_variables(prefix: str):
    """[(name, shape, numpy dtype)] like tf.train.list_variables."""
    out = []
    for k, v in read_table(prefix + ".index").items():
        if k == b"":
            continue
        e = _parse_entry(v)
        out.append((k.decode(), e["shape"], _DTYPES.get(e["dtype"])))
    return out


def load_checkpoint(prefix: str, verify_data_crc: bool = False, name_prefix: str = None) -> Dict[str, np.ndarray]:
    """{variable name: ndarray} of the bundle ``<prefix>.index`` + ``<prefix>.data-*``.  The table
    blocks are always checksummed; ``verify_data_crc`` also checks every tensor (pure-Python crc32c:
    ~10 s for the 68 MB of this net).  ``name_prefix`` (e.g. ``"net/"``) materialises only the variables whose
    name starts with it (plus ``global_step``): a training checkpoint carries two Adam slots per weight."""
    table = read_table(prefix + ".index")
    num_shards, endian = 1, 0
    for f, _, v in _proto_fields(table.get(b"", b"")):
        if f == 1:
            num_shards = v
        elif f == 2:
            endian = v
    if endian != 0:
        raise NotImplementedError("big-endian bundle")
    shards = {}
    out = {}
    for k, v in table.items():
        if k == b"":
            continue
        name = k.decode()
        if name_prefix is not None and name != "global_step" and not name.startswith(name_prefix):
            continue
        if name_prefix is not None and re.search(r"/Adam(_1)?$", name):
            continue
        e = _parse_entry(v)
        if e["slices"]:
            raise NotImplementedError("partitioned variable %r (tensor slices)" % k.decode())
        if e["dtype"] not in _DTYPES:
            raise NotImplementedError("dtype %d of %r" % (e["dtype"], k.decode()))
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(_DTYPES[e["dtype"]])
        n = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if n * dt.itemsize != e["size"]:
            raise ValueError("%r: %d bytes for shape %s %s" % (k.decode(), e["size"], e["shape"], dt))
        if verify_data_crc and e["crc32c"] is not None and masked_crc32c(raw.tobytes()) != e["crc32c"]:
            raise ValueError("tensor %r: crc32c mismatch" % k.decode())
        out[name] = np.array(np.frombuffer(raw, dtype=dt).reshape(e["shape"]))   # one copy, out of the mapping
    return out


def latest_checkpoint(checkpoint_dir: str):
    """tf.train.latest_checkpoint: the prefix named by ``model_checkpoint_path`` in ``<dir>/checkpoint``."""
    p = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.exists(p):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"(.*)"', open(p).read(), re.M)
    if not m:
        return None
    name = m.group(1)
    return name if os.path.isabs(name) else os.path.join(checkpoint_dir, name)


# ---- writer (tests, and converting an .npz of weights into a bundle TensorFlow can restore) -------------
def _pb_varint(field, v):
    return _put_varint(field << 3) + _put_varint(v & 0xFFFFFFFFFFFFFFFF)


def _pb_bytes(field, b):
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


def _build_block(items, restart_interval=16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def save_checkpoint(prefix: str, tensors: Dict[str, np.ndarray], block_bytes: int = 4096) -> str:
    """Writes ``<prefix>.index``, ``<prefix>.data-00000-of-00001`` and ``<dir>/checkpoint``."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    entries = [(b"", _pb_varint(1, 1) + _pb_varint(2, 0) + _pb_bytes(3, _pb_varint(1, 1)))]
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        off = 0
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[name])
            a = a if a.flags.c_contiguous else a.copy(order="C")  # (ascontiguousarray would turn a scalar into shape (1,))
            raw = a.tobytes()
            shape = b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in a.shape)
            e = (_pb_varint(1, _DTYPE_IDS[a.dtype]) + _pb_bytes(2, shape) + (_pb_varint(4, off) if off else b"") +
                 _pb_varint(5, len(raw)) + _put_varint((6 << 3) | 5) + struct.pack("<I", masked_crc32c(raw)))
            entries.append((name.encode(), e))
            f.write(raw)
            off += len(raw)
    out = bytearray()

    def emit(block: bytes):
        handle = _put_varint(len(out)) + _put_varint(len(block))
        trailer = b"\x00"
        out.extend(block + trailer + struct.pack("<I", masked_crc32c(block + trailer)))
        return handle

    index_items, cur, cur_size = [], [], 0
    for k, v in entries:
        cur.append((k, v))
        cur_size += len(k) + len(v) + 3
        if cur_size >= block_bytes:
            index_items.append((cur[-1][0], emit(_build_block(cur))))
            cur, cur_size = [], 0
    if cur:
        index_items.append((cur[-1][0], emit(_build_block(cur))))
    meta = emit(_build_block([]))
    index = emit(_build_block(index_items, restart_interval=1))
    footer = meta + index
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))
    open(prefix + ".index", "wb").write(bytes(out))
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
    return prefix


def load_weights(path: str) -> Dict[str, np.ndarray]:
    """Weights for NetEngine / MSI from an ``.npz`` keyed by the TF variable names, a bundle prefix, or a
    directory holding a ``checkpoint`` file (the reference's --checkpoint_dir/--experiment_name)."""
    if os.path.isdir(path):
        prefix = latest_checkpoint(path)
        if prefix is None:
            raise FileNotFoundError("no 'checkpoint' file in %s" % path)
        return load_checkpoint(prefix, name_prefix="net/")
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    if path.endswith(".index"):
        path = path[:-len(".index")]
    return load_checkpoint(path)
