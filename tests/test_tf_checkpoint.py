"""matryodshka_b200.tf_checkpoint: checkpoint-v2 tensor bundles without TensorFlow (CPU only).
The reader is pinned to tests/golden/tf_bundle/ -- a bundle assembled independently of the module (official protobuf
runtime, TensorBoard's generated protos and crc32c; tests/golden/make_tf_bundle_fixture.py) -- and exercised against the
module's own writer, plus known answers for the pieces that have them (crc32c, varints, table framing).  No
TensorFlow-written file exists in the sandbox."""
import os
import struct

import numpy as np
import pytest

from matryodshka_b200 import synth, tf_checkpoint as tfc


def test_crc32c_known_answers():
    assert tfc.crc32c(b"123456789") == 0xE3069283          # the standard CRC-32C check value
    assert tfc.crc32c(b"") == 0
    assert tfc.crc32c(bytes(32)) == 0x8A9136AA             # RFC 3720 B.4: 32 zero bytes
    assert tfc.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43    # RFC 3720 B.4: 32 0xFF bytes
    c = tfc.crc32c(b"abc")
    assert tfc.masked_crc32c(b"abc") == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_varint_round_trip():
    for v in [0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 63 + 5]:
        b = tfc._put_varint(v)
        assert tfc._varint(b, 0) == (v, len(b))


def test_bundle_round_trip_full_net(tmp_path):
    """All 53 variables of the net (conv / deconv weights, LayerNorm gamma / beta, head bias) + global_step."""
    wts = synth.net_weights(24, 8, 8)
    wts = dict(wts)
    wts["global_step"] = np.array(410000, dtype=np.int64)
    prefix = tfc.save_checkpoint(str(tmp_path / "exp" / "model.ckpt-410000"), wts, block_bytes=512)  # many table blocks
    assert tfc.latest_checkpoint(str(tmp_path / "exp")) == prefix
    got = tfc.load_checkpoint(prefix, verify_data_crc=True)
    assert set(got) == set(wts)
    for k in wts:
        assert got[k].dtype == np.asarray(wts[k]).dtype and got[k].shape == np.asarray(wts[k]).shape
        assert np.array_equal(got[k], wts[k])
    names = [n for n, _, _ in tfc.list_variables(prefix)]
    assert names == sorted(names, key=lambda s: s.encode()) and "net/conv1_1/weights" in names
    # load_weights accepts the directory, the prefix, the .index path
    for p in (str(tmp_path / "exp"), prefix, prefix + ".index"):
        assert np.array_equal(tfc.load_weights(p)["net/color_pred/biases"], wts["net/color_pred/biases"])


def test_table_framing_and_corruption(tmp_path):
    prefix = tfc.save_checkpoint(str(tmp_path / "m.ckpt"), {"a/b": np.arange(6, dtype=np.float32).reshape(2, 3),
                                                            "a/c": np.ones((1, 1, 2, 2), np.float64)})
    data = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", data[-8:])[0] == 0xDB4775248B80FB57 and len(data) >= 48
    table = tfc.read_table(prefix + ".index")
    assert list(table)[0] == b"" and set(table) == {b"", b"a/b", b"a/c"}
    # a flipped byte in a block is caught by the block checksum, a flipped tensor byte by the entry crc
    bad = bytearray(data)
    bad[3] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        tfc.read_table(prefix + ".index")
    open(prefix + ".index", "wb").write(data)
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[5] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        tfc.load_checkpoint(prefix, verify_data_crc=True)


def test_npz_weights(tmp_path):
    wts = synth.net_weights(24, 8, 8)
    np.savez(tmp_path / "weights.npz", **wts)
    got = tfc.load_weights(str(tmp_path / "weights.npz"))
    assert all(np.array_equal(got[k], wts[k]) for k in wts)


# ---- the independently built bundle (tests/golden/tf_bundle/, written by tests/golden/make_tf_bundle_fixture.py) ----
FIX = os.path.join(os.path.dirname(__file__), "golden", "tf_bundle")


def _fixture_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_tf_bundle_fixture",
                                                  os.path.join(os.path.dirname(__file__), "golden", "make_tf_bundle_fixture.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["model.ckpt-410000", "small_blocks.ckpt"])
def test_reader_against_the_independent_bundle(name):
    """Every tensor of the committed bundle (one 256 KB table block as TensorFlow writes it / many 512-byte blocks with
    shortened index separators), data crc32c verified with OUR crc32c against TensorBoard's, values from the seed."""
    want = _fixture_module().fixture_tensors()
    got = tfc.load_checkpoint(os.path.join(FIX, name), verify_data_crc=True)
    assert set(got) == set(want)
    for k, v in want.items():
        v = np.asarray(v)
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    assert int(got["global_step"]) == 410000 and got["global_step"].dtype == np.int64
    names = [n for n, _, _ in tfc.list_variables(os.path.join(FIX, name))]
    assert names == sorted(want, key=lambda s: s.encode())
    # the Saver's directory layout and the net/ filter the driver uses (Adam slots and optimizer scalars dropped)
    assert tfc.latest_checkpoint(FIX) == os.path.join(FIX, "model.ckpt-410000")
    w = tfc.load_weights(FIX)
    assert set(w) == {k for k in want if k.startswith("net/") and not k.endswith(("/Adam", "/Adam_1"))} | {"global_step"}


def test_the_fixture_script_reproduces_the_committed_bytes(tmp_path):
    """Where TensorBoard / protobuf are importable (they are in this image) the builder script rewrites the committed
    files bit for bit -- the fixture is not a hand-edited blob."""
    pytest.importorskip("tensorboard")
    m = _fixture_module()
    t = m.fixture_tensors()
    for name, bs in (("model.ckpt-410000", 262144), ("small_blocks.ckpt", 512)):
        m.write_bundle(str(tmp_path / name), t, block_size=bs)
        for ext in (".index", ".data-00000-of-00001"):
            assert open(str(tmp_path / name) + ext, "rb").read() == open(os.path.join(FIX, name) + ext, "rb").read(), name + ext
    # and our own writer produces a table our reader AND the fixture's independent protos agree on
    tfc.save_checkpoint(str(tmp_path / "ours.ckpt"), {k: np.asarray(v) for k, v in t.items()})
    Header, Entry = m.bundle_protos()
    table = tfc.read_table(str(tmp_path / "ours.ckpt.index"))
    h = Header.FromString(table[b""])
    assert h.num_shards == 1 and h.endianness == 0 and h.version.producer == 1
    e = Entry.FromString(table[b"net/conv1_2/weights"])
    assert [d.size for d in e.shape.dim] == [3, 3, 5, 8] and e.size == 3 * 3 * 5 * 8 * 4 and e.dtype == 1
    from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import masked_crc32c as tb_crc
    assert e.crc32c == tb_crc(np.asarray(t["net/conv1_2/weights"]).tobytes())
