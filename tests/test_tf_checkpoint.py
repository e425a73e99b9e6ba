"""CPU: TensorFlow V2 checkpoint (tensor bundle) reader / writer -- the container the reference's weights ship in
(FISRnet.py:1101-1115, README.md:56-66).  TensorFlow is absent here, so the checks are the published format's known
answers (crc32c, snappy, protobuf, LevelDB table layout) and writer -> reader round trips."""
import os
import struct

import numpy as np
import pytest

from fisr_b200 import tf_checkpoint as T


def test_crc32c_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283                 # RFC 3720 B.4 check value
    assert T.crc32c(b"\x00" * 32) == 0x8A9136AA                  # RFC 3720 B.4: 32 bytes of zeros
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.mask_crc(0) == 0xA282EAD8


def test_snappy_decoder():
    # hand-assembled stream: length 20; literal "abcd" ; copy(offset 4, len 8) 1-byte-offset form; literal "XY";
    # copy(offset 14, len 6) 2-byte-offset form
    lit1 = bytes([(4 - 1) << 2]) + b"abcd"
    copy1 = bytes([((8 - 4) << 2) | 1 | (0 << 5), 4])
    lit2 = bytes([(2 - 1) << 2]) + b"XY"
    copy2 = bytes([((6 - 1) << 2) | 2, 14, 0])
    stream = bytes([20]) + lit1 + copy1 + lit2 + copy2
    assert T.snappy_decompress(stream) == b"abcdabcdabcdXYabcdab"
    with pytest.raises(ValueError):
        T.snappy_decompress(bytes([5]) + lit1)


def test_entry_and_header_protos():
    e = T.parse_entry(T.build_entry(T.DT_FLOAT, (3, 3, 64, 256), 0, 123456789, 3 * 3 * 64 * 256 * 4, 0xDEADBEEF))
    assert e == {"dtype": 1, "shape": [3, 3, 64, 256], "shard_id": 0, "offset": 123456789, "size": 589824, "crc32c": 0xDEADBEEF,
                 "slices": 0}
    # the byte string protoc would emit for BundleEntryProto{dtype: DT_FLOAT, shape{dim{size:2}}, size: 8, crc32c: 1}
    assert T.build_entry(1, (2,), 0, 0, 8, 1) == bytes([0x08, 0x01, 0x12, 0x04, 0x12, 0x02, 0x08, 0x02, 0x28, 0x08, 0x35, 1, 0, 0, 0])
    assert T.parse_header(T.build_header(1)) == {"num_shards": 1, "endianness": 0}
    assert T.parse_entry(T.build_entry(T.DT_INT64, (), 0, 0, 8, 0))["shape"] == []


def _fisr_like(rng, extra=True):
    from fisr_b200.engine import param_inventory
    t = {}
    for name, shape in param_inventory().items():
        t[name] = rng.standard_normal(shape).astype(np.float32) if name.endswith("/w") and shape[2] <= 64 else np.zeros(shape, np.float32)
        if extra:                                                # what tf.train.AdamOptimizer adds to a training checkpoint
            t[name + "/Adam"] = np.zeros(1, np.float32)
            t[name + "/Adam_1"] = np.zeros(1, np.float32)
    if extra:
        t["beta1_power"] = np.float32(0.5).reshape(())
        t["beta2_power"] = np.float32(0.9).reshape(())
        t["Variable"] = np.int32(122000).reshape(())
    return t


def test_round_trip_full_fisrnet_checkpoint(tmp_path):
    rng = np.random.default_rng(0)
    tensors = _fisr_like(rng)
    prefix = str(tmp_path / "FISRnet_exp1" / "FISRnet-122000")
    T.save_checkpoint(prefix, tensors)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    # structure: footer magic, several data blocks (830 keys do not fit one 4 KB block)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == T.TABLE_MAGIC
    listed = T.list_variables(prefix)
    assert len(listed) == len(tensors) == 276 * 3 + 3
    assert listed["FISRnet/level_3/FI-SR/conv/1/w"] == (T.DT_FLOAT, (3, 3, 64, 256))
    assert listed["Variable"] == (T.DT_INT32, ())
    keys = list(T.read_table(prefix + ".index", verify=True))
    assert keys == sorted(keys) and keys[0] == b""
    w = T.fisrnet_weights(prefix)
    assert len(w) == 276
    for k, v in w.items():
        assert v.dtype == np.float32 and np.array_equal(v, tensors[k])
    some = T.load_checkpoint(prefix, ["Variable", "beta1_power", "FISRnet/level_1/enc/level_0/conv/0/w"], verify_crc=True)
    assert int(some["Variable"]) == 122000 and some["FISRnet/level_1/enc/level_0/conv/0/w"].shape == (3, 3, 29, 64)
    with pytest.raises(KeyError):
        T.load_checkpoint(prefix, ["no/such/variable"])


def test_detects_corruption_and_missing_variables(tmp_path):
    prefix = str(tmp_path / "m")
    T.save_checkpoint(prefix, {"a/w": np.arange(12, dtype=np.float32).reshape(3, 4), "a/b": np.ones(4, np.float32)})
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[5] ^= 0xFF
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="crc32c"):
        T.load_checkpoint(prefix, verify_crc=True)
    with pytest.raises(KeyError, match="lacks"):
        T.fisrnet_weights(prefix)
    open(prefix + ".index", "ab").write(b"x")
    with pytest.raises(ValueError, match="magic"):
        T.read_table(prefix + ".index")


def test_snappy_compressed_block_is_read(tmp_path):
    """A table whose data block is stored snappy-compressed (type 1) as an all-literal stream."""
    prefix = str(tmp_path / "s")
    T.save_checkpoint(prefix, {"v": np.arange(4, dtype=np.float32)})
    raw = open(prefix + ".index", "rb").read()
    footer = raw[-48:]
    _, p = T._varint(footer, 0); _, p = T._varint(footer, p)
    idx_off, p = T._varint(footer, p); idx_size, p = T._varint(footer, p)
    (sep, handle), = list(T._block_entries(raw[idx_off:idx_off + idx_size]))
    off, q = T._varint(handle, 0); size, _ = T._varint(handle, q)
    block = raw[off:off + size]
    assert len(block) < 60 * 256
    lit = (bytes([(len(block) - 1) << 2]) if len(block) <= 60 else bytes([60 << 2, len(block) - 1])) + block
    comp = T._put_varint(len(block)) + lit
    table = bytearray(comp + b"\x01" + struct.pack("<I", T.mask_crc(T.crc32c(comp + b"\x01"))))
    meta = T._build_block([])
    mh = T._put_varint(len(table)) + T._put_varint(len(meta))
    table += meta + b"\x00" + struct.pack("<I", T.mask_crc(T.crc32c(meta + b"\x00")))
    index = T._build_block([(sep, T._put_varint(0) + T._put_varint(len(comp)))], 1)
    ih = T._put_varint(len(table)) + T._put_varint(len(index))
    table += index + b"\x00" + struct.pack("<I", T.mask_crc(T.crc32c(index + b"\x00")))
    f = mh + ih
    table += f + b"\x00" * (40 - len(f)) + struct.pack("<Q", T.TABLE_MAGIC)
    open(prefix + ".index", "wb").write(bytes(table))
    got = T.load_checkpoint(prefix, verify_crc=True)
    assert np.array_equal(got["v"], np.arange(4, dtype=np.float32))


def test_latest_checkpoint_state_file(tmp_path):
    d = tmp_path / "FISRnet_exp1"
    d.mkdir()
    (d / "checkpoint").write_text('model_checkpoint_path: "FISRnet-122000"\nall_model_checkpoint_paths: "FISRnet-122000"\n')
    assert T.latest_checkpoint(str(d)) == str(d / "FISRnet-122000")
    assert T.latest_checkpoint(str(tmp_path)) is None


def test_training_checkpoint_round_trips_optimizer_state(tmp_path):
    """What tf.train.Saver stores for a TRAINING run (FISRnet.py:489-491,1092-1099): weights + <var>/Adam + <var>/Adam_1 +
    beta powers + the global step.  An inference-only export (weights alone) reports no optimizer state."""
    from fisr_b200.engine import param_inventory
    rng = np.random.default_rng(3)
    small = lambda shape: (rng.standard_normal(shape) * 1e-3).astype(np.float32) if np.prod(shape) < 40000 else np.zeros(shape, np.float32)
    params = {k: small(s) for k, s in param_inventory().items()}
    adam = {"m": {k: small(s) for k, s in param_inventory().items()}, "v": {k: np.abs(small(s)) for k, s in param_inventory().items()}, "t": 37}
    prefix = str(tmp_path / "FISRnet_exp1" / "FISRnet-37")
    T.save_fisrnet_checkpoint(prefix, params, adam, global_step=37)
    listed = T.list_variables(prefix)
    assert len(listed) == 3 * 276 + 3 and listed[T.GLOBAL_STEP_NAME] == (T.DT_INT32, ())
    b1 = T.load_checkpoint(prefix, ["beta1_power", "beta2_power", T.GLOBAL_STEP_NAME])
    assert float(b1["beta1_power"]) == pytest.approx(0.9 ** 37, rel=1e-6) and float(b1["beta2_power"]) == pytest.approx(0.999 ** 37, rel=1e-6)
    assert int(b1[T.GLOBAL_STEP_NAME]) == 37
    m, v, t = T.fisrnet_adam_state(prefix)
    assert t == 37
    k = "FISRnet/level_2/SR/conv/2/w"
    assert np.array_equal(m[k], adam["m"][k]) and np.array_equal(v[k], adam["v"][k])
    w = T.fisrnet_weights(prefix)
    assert np.array_equal(w[k], params[k])
    prefix2 = str(tmp_path / "export" / "FISRnet-122000")
    T.save_fisrnet_checkpoint(prefix2, params, None, global_step=122000)
    assert T.fisrnet_adam_state(prefix2) is None and len(T.list_variables(prefix2)) == 277
    # a corrupted small tensor is caught by the load path FISRnet.load uses (crc32c of the index and of tensors <= 64 KB)
    path = prefix2 + ".data-00000-of-00001"
    raw = bytearray(open(path, "rb").read())
    off = T.parse_entry(T.read_table(prefix2 + ".index")[b"FISRnet/level_1/SR/conv/2/b"])["offset"]
    raw[off] ^= 0x40
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="crc32c"):
        T.fisrnet_weights(prefix2)
