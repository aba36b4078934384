"""TF-V2 checkpoint reader / writer (wave-u-net_b200/TFCheckpoint.py) - CPU tests.

TensorFlow cannot be installed here and the reference ships no checkpoint file, so there is no TF-written fixture:
the format is pinned through published known answers of its primitives (RFC 3720 CRC-32C vectors, the LevelDB mask
constant and table magic, protobuf wire bytes written out by hand below) plus round trips - and, at the end of this file,
against Google's own code where this image has it: TensorBoard's TensorFlow stub (CRC-32C, mask) and the protobuf runtime over
TensorBoard's generated TensorFlow protos (BundleEntryProto / BundleHeaderProto byte for byte, the `checkpoint` state file).
The LevelDB table layout of the .index stays "parity unpinned", and the module header says so too."""
import os
import struct

import numpy as np
import pytest

import Config
import TFCheckpoint as tfc
import wun
from Models.UnetAudioSeparator import UnetAudioSeparator


def test_crc32c_known_answers():
    # RFC 3720 (iSCSI) appendix B.4 + the classic check value
    assert wun.crc32c(b"123456789") == 0xE3069283
    assert wun.crc32c(bytes(32)) == 0x8A9136AA
    assert wun.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert wun.crc32c(bytes(range(32))) == 0x46DD794E
    assert wun.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert wun.crc32c(b"") == 0
    # extend == one shot, at unaligned split points
    a = np.random.default_rng(0).integers(0, 256, 4099, dtype=np.uint8)
    for cut in (0, 1, 7, 8, 9, 2048, 4098, 4099):
        assert wun.crc32c(a[cut:], wun.crc32c(a[:cut])) == wun.crc32c(a)


def test_crc_mask_is_leveldbs():
    # crc32c::Mask(crc) = ((crc >> 15) | (crc << 17)) + 0xa282ead8  (leveldb util/crc32c.h)
    assert tfc.mask_crc(0) == 0xA282EAD8
    assert tfc.mask_crc(0xE3069283) == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
    for c in (0, 1, 0xDEADBEEF, 0xFFFFFFFF, 0xE3069283):
        assert tfc.unmask_crc(tfc.mask_crc(c)) == c
        assert tfc.mask_crc(c) != c


def test_varint_and_proto_wire_bytes():
    for v, enc in ((0, b"\x00"), (1, b"\x01"), (127, b"\x7f"), (128, b"\x80\x01"), (300, b"\xac\x02"),
                   (2 ** 32, b"\x80\x80\x80\x80\x10")):
        out = bytearray()
        tfc.put_varint(out, v)
        assert bytes(out) == enc and tfc.get_varint(enc, 0) == (v, len(enc))
    # BundleHeaderProto{num_shards: 1, version{producer: 1}} written out by hand
    assert tfc.encode_header(1) == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    assert tfc.decode_header(tfc.encode_header(1)) == {"num_shards": 1, "endianness": 0, "producer": 1}
    # BundleEntryProto{dtype: DT_FLOAT, shape{dim{size:15} dim{size:2} dim{size:24}}, offset: 300, size: 2880, crc32c: 0x01020304}
    want = bytes([0x08, 0x01,
                  0x12, 0x0C, 0x12, 0x02, 0x08, 0x0F, 0x12, 0x02, 0x08, 0x02, 0x12, 0x02, 0x08, 0x18,
                  0x20, 0xAC, 0x02,
                  0x28, 0xC0, 0x16,
                  0x35, 0x04, 0x03, 0x02, 0x01])
    assert tfc.encode_entry(tfc.DT_FLOAT, (15, 2, 24), 300, 2880, 0x01020304) == want
    e = tfc.decode_entry(want)
    assert (e["dtype"], e["shape"], e["offset"], e["size"], e["crc32c"]) == (1, [15, 2, 24], 300, 2880, 0x01020304)
    # scalar: the shape field is present and empty; zero offset is omitted (proto3 default)
    s = tfc.encode_entry(tfc.DT_INT64, (), 0, 8, 7)
    assert s == bytes([0x08, 0x09, 0x12, 0x00, 0x28, 0x08, 0x35, 0x07, 0, 0, 0])
    assert tfc.decode_entry(s)["shape"] == []


def test_table_layout_prefix_compression_restarts_and_footer():
    keys = [b""] + [("separator/conv1d_%d/kernel" % i).encode() for i in range(40)]
    keys = sorted(set(keys))
    items = [(k, b"v" + k[-3:]) for k in keys]
    for block_size in (tfc.BLOCK_SIZE, 64):                       # one data block / many data blocks
        data = tfc.build_table(items, block_size)
        assert struct.unpack("<Q", data[-8:])[0] == 0xdb4775248b80fb57 and len(data[-48:]) == 48
        assert tfc.read_table(data) == items
    data = tfc.build_table(items)
    # first entry of the first block: shared = 0, key "" (non_shared 0), value length 1
    assert data[:4] == bytes([0, 0, 1]) + b"v"
    # second entry shares nothing with "", third shares the long common prefix with the second
    second = items[1][0]
    assert data[4:7] == bytes([0, len(second), len(items[1][1])])
    pos = 7 + len(second) + len(items[1][1])
    shared = len(os.path.commonprefix([items[1][0], items[2][0]]))
    assert data[pos] == shared and shared > 10
    # 41 entries -> restarts at entries 0, 16, 32
    footer = data[-48:]
    p = 0
    moff, p = tfc.get_varint(footer, p); msize, p = tfc.get_varint(footer, p)
    ioff, p = tfc.get_varint(footer, p); isize, p = tfc.get_varint(footer, p)
    assert msize == 8 and data[moff:moff + 8] == bytes([0, 0, 0, 0, 1, 0, 0, 0])      # empty metaindex block
    first_block_size = moff - 5
    assert struct.unpack_from("<I", data, first_block_size - 4)[0] == 3
    # block trailer: type 0 + masked crc of (block + type)
    assert data[first_block_size] == 0
    stored = struct.unpack_from("<I", data, first_block_size + 1)[0]
    assert tfc.unmask_crc(stored) == wun.crc32c(data[:first_block_size + 1])
    # corruption is detected, and unsorted keys are refused
    bad = bytearray(data); bad[10] ^= 1
    with pytest.raises(tfc.CheckpointError):
        tfc.read_table(bytes(bad))
    with pytest.raises(tfc.CheckpointError):
        tfc.build_table([(b"b", b""), (b"a", b"")])
    with pytest.raises(tfc.CheckpointError):
        tfc.read_table(data[:-1] + b"\x00")


def test_bundle_round_trip_dtypes_scalars_and_checks(tmp_path):
    rng = np.random.default_rng(3)
    tensors = {"b/w": rng.standard_normal((15, 2, 24)).astype(np.float32), "global_step": np.int64(2000),
               "a": np.float32(0.5), "z/empty": np.zeros((0, 3), np.float32), "c/d": rng.integers(0, 9, (4, 4)).astype(np.int32),
               "flag": np.array([True, False]), "dbl": rng.standard_normal(5)}
    prefix = str(tmp_path / "ck" / "42-2000")
    assert tfc.write_checkpoint(prefix, tensors) == prefix
    assert sorted(os.listdir(tmp_path / "ck")) == ["42-2000.data-00000-of-00001", "42-2000.index", "checkpoint"]
    assert tfc.latest_checkpoint(str(tmp_path / "ck")) == prefix
    got = tfc.read_checkpoint(prefix)
    assert list(got) == sorted(tensors)                                        # key order = byte order of the names
    for k, v in tensors.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape
        np.testing.assert_array_equal(got[k], v)
    assert ("b/w", np.float32, (15, 2, 24)) in tfc.list_variables(prefix)
    assert list(tfc.read_checkpoint(prefix, names=["a", "global_step"])) == ["a", "global_step"]
    # data file = tensors back to back in key order, no padding
    sizes = [np.asarray(tensors[k]).nbytes for k in sorted(tensors)]
    assert os.path.getsize(prefix + ".data-00000-of-00001") == sum(sizes)
    # a flipped data byte is caught by the per-tensor checksum
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(sum(sizes[:2]) + 1); b = f.read(1); f.seek(-1, 1); f.write(bytes([b[0] ^ 0x40]))
    with pytest.raises(tfc.CheckpointError):
        tfc.read_checkpoint(prefix)
    assert "b/w" in tfc.read_checkpoint(prefix, verify=False)
    with pytest.raises(tfc.CheckpointError):
        tfc.read_checkpoint(str(tmp_path / "missing"))


@pytest.mark.parametrize("preset", ["baseline_stereo", "full"])
def test_separator_variable_set_round_trips_with_either_optimizer_scope(tmp_path, preset):
    mc = Config.build_config([preset], dict(num_layers=3, num_frames=64), experiment_id=1)["model_config"]
    sep = UnetAudioSeparator(mc)
    table = sep.param_table(num_frames=64)
    names = [n for n, _, _, _ in table]
    assert names[0] == "separator/conv1d/kernel" and names[1] == "separator/conv1d/bias"
    rng = np.random.default_rng(5)
    var = {n: rng.standard_normal(s).astype(np.float32) for n, s, _, _ in table}
    m = {n: rng.standard_normal(s).astype(np.float32) for n, s, _, _ in table}
    v = {n: rng.random(s).astype(np.float32) for n, s, _, _ in table}
    tensors = tfc.separator_tensors(var, m, v, global_step=7)
    # what tf.train.Saver(tf.global_variables()) holds for Training.py:66-77: variables, 2 slots each, 2 powers, the step
    assert len(tensors) == 3 * len(names) + 3
    assert tensors["separator_solver/separator/conv1d/kernel/Adam_1"].shape == tuple(table[0][1])
    assert np.isclose(tensors["separator_solver/beta1_power"], 0.9 ** 8) and tensors["global_step"].dtype == np.int64
    prefix = tfc.write_checkpoint(str(tmp_path / "1-7"), tensors)
    back = tfc.read_checkpoint(prefix)
    var2, m2, v2, step = tfc.split_separator_tensors(back, names)
    assert step == 7 and list(var2) == names
    for n in names:
        np.testing.assert_array_equal(var2[n], var[n]); np.testing.assert_array_equal(m2[n], m[n]); np.testing.assert_array_equal(v2[n], v[n])
    # the other naming a TF build may have used for the slots (no optimizer scope prefix) restores as well
    alt = {(k[len("separator_solver/"):] if k.startswith("separator_solver/separator/") else k): a for k, a in tensors.items()}
    _, m3, v3, _ = tfc.split_separator_tensors(alt, names)
    np.testing.assert_array_equal(m3[names[2]], m[names[2]])
    # inference-only checkpoint (no slots) and a checkpoint of another architecture
    only = tfc.separator_tensors(var)
    _, m4, v4, step4 = tfc.split_separator_tensors(only, names)
    assert m4 is None and v4 is None and step4 == 0
    with pytest.raises(tfc.CheckpointError):
        tfc.split_separator_tensors({k: a for k, a in only.items() if k != names[3]}, names)


# ---- pinned against Google's own code where it is available here: TensorBoard's TensorFlow stub ---------------------------------
def _bundle_messages():
    """BundleHeaderProto / BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto) and CheckpointState
    (tensorflow/python/training/checkpoint_state.proto), declared with the protobuf runtime on top of the GENUINE generated
    TensorShapeProto, DataType and VersionDef that TensorBoard ships (tensorboard.compat.proto: compiled from TensorFlow's .proto
    files).  The outer field numbers are restated from tensor_bundle.proto; everything nested is Google's."""
    pytest.importorskip("tensorboard")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2      # noqa: F401 (registers the files in the default pool)
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "wun_test/tensor_bundle_restated.proto"
    fd.package = "wun_test"
    fd.syntax = "proto3"
    fd.dependency.extend([tensor_shape_pb2.DESCRIPTOR.name, types_pb2.DESCRIPTOR.name, versions_pb2.DESCRIPTOR.name])
    h = fd.message_type.add(); h.name = "BundleHeaderProto"
    for name, num, typ, tn in (("num_shards", 1, F.TYPE_INT32, None), ("endianness", 2, F.TYPE_INT32, None),
                               ("version", 3, F.TYPE_MESSAGE, ".tensorboard.VersionDef")):
        f = h.field.add(); f.name, f.number, f.type, f.label = name, num, typ, F.LABEL_OPTIONAL
        if tn:
            f.type_name = tn
    e = fd.message_type.add(); e.name = "BundleEntryProto"
    for name, num, typ, tn in (("dtype", 1, F.TYPE_ENUM, ".tensorboard.DataType"), ("shape", 2, F.TYPE_MESSAGE, ".tensorboard.TensorShapeProto"),
                               ("shard_id", 3, F.TYPE_INT32, None), ("offset", 4, F.TYPE_INT64, None), ("size", 5, F.TYPE_INT64, None),
                               ("crc32c", 6, F.TYPE_FIXED32, None)):
        f = e.field.add(); f.name, f.number, f.type, f.label = name, num, typ, F.LABEL_OPTIONAL
        if tn:
            f.type_name = tn
    c = fd.message_type.add(); c.name = "CheckpointState"
    for name, num, label in (("model_checkpoint_path", 1, F.LABEL_OPTIONAL), ("all_model_checkpoint_paths", 2, F.LABEL_REPEATED)):
        f = c.field.add(); f.name, f.number, f.type, f.label = name, num, F.TYPE_STRING, label
    pool = descriptor_pool.Default()
    try:
        filed = pool.Add(fd) or pool.FindFileByName(fd.name)
    except TypeError:                                        # second call in one process: already registered
        filed = pool.FindFileByName(fd.name)
    get = getattr(message_factory, "GetMessageClass", None)
    return tuple((get(filed.message_types_by_name[n]) if get else message_factory.MessageFactory(pool).GetPrototype(filed.message_types_by_name[n]))
                 for n in ("BundleHeaderProto", "BundleEntryProto", "CheckpointState"))


def test_hand_encoded_bundle_protos_match_the_protobuf_runtime_over_tensorboards_tf_protos():
    Header, Entry, _ = _bundle_messages()
    from tensorboard.compat.proto import types_pb2
    assert (types_pb2.DT_FLOAT, types_pb2.DT_INT32, types_pb2.DT_INT64, types_pb2.DT_DOUBLE) == (tfc.DT_FLOAT, tfc.DT_INT32, tfc.DT_INT64, tfc.DT_DOUBLE)
    h = Header(num_shards=1)
    h.version.producer = 1
    assert h.SerializeToString(deterministic=True) == tfc.encode_header(1)
    assert Header.FromString(tfc.encode_header(1)).version.producer == 1
    rng = np.random.default_rng(5)
    for dtype, shape, offset, size, crc in ((tfc.DT_FLOAT, (15, 2, 24), 300, 2880, 0x01020304), (tfc.DT_INT64, (), 0, 8, 7),
                                            (tfc.DT_FLOAT, (312,), 2 ** 33 + 5, 1248, 0xFFFFFFFF), (tfc.DT_FLOAT, (5, 600, 288), 0, 3456000, 0),
                                            (tfc.DT_DOUBLE, (1, 1, 1, 7), 123456789012, 56, int(rng.integers(0, 2 ** 32)))):
        mine = tfc.encode_entry(dtype, shape, offset, size, crc)
        m = Entry(dtype=dtype, offset=offset, size=size, crc32c=crc)
        m.shape.SetInParent()                                # TF always writes the shape message, also for scalars
        for d in shape:
            m.shape.dim.add().size = d
        assert m.SerializeToString(deterministic=True) == mine, (dtype, shape)
        back = Entry.FromString(mine)
        assert (back.dtype, [d.size for d in back.shape.dim], back.offset, back.size, back.crc32c) == (dtype, list(shape), offset, size, crc)
        d = tfc.decode_entry(m.SerializeToString())          # and the reader understands what the runtime writes
        assert (d["dtype"], d["shape"], d["offset"], d["size"], d["crc32c"]) == (dtype, list(shape), offset, size, crc)


def test_crc32c_and_its_mask_match_tensorboards_tensorflow_stub():
    """tensorboard.compat.tensorflow_stub.pywrap_tensorflow carries Google's pure-Python CRC-32C and TFRecord masking
    (masked_crc32c) - the same mask the tensor bundle stores per tensor and LevelDB per table block."""
    pytest.importorskip("tensorboard")
    from tensorboard.compat.tensorflow_stub import pywrap_tensorflow as tb
    rng = np.random.default_rng(9)
    for n in (0, 1, 3, 4, 7, 8, 9, 63, 64, 65, 1000, 4099):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert wun.crc32c(data) == (tb.crc32c(data) & 0xFFFFFFFF), n
        assert tfc.mask_crc(wun.crc32c(data)) == (tb.masked_crc32c(data) & 0xFFFFFFFF), n


def test_written_checkpoint_parses_with_the_protobuf_runtime(tmp_path):
    """Every BundleEntryProto / the header of a written .index and the `checkpoint` state file (text format) through the real
    protobuf parsers; each tensor's stored crc32c equals TensorBoard's masked CRC-32C of its bytes in the data file."""
    Header, Entry, State = _bundle_messages()
    from google.protobuf import text_format
    from tensorboard.compat.tensorflow_stub import pywrap_tensorflow as tb
    rng = np.random.default_rng(2)
    tensors = {"separator/conv1d/kernel": rng.standard_normal((15, 2, 24)).astype(np.float32),
               "separator/conv1d/bias": np.zeros(24, np.float32), "global_step": np.int64(1234),
               "separator_solver/beta1_power": np.float32(0.9 ** 1234), "separator/interp_0": rng.standard_normal(312).astype(np.float32)}
    prefix = tfc.write_checkpoint(str(tmp_path / "run" / "run-1234"), tensors)
    table = tfc.read_table(open(prefix + ".index", "rb").read())
    assert table[0][0] == b"" and Header.FromString(table[0][1]).num_shards == 1
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    seen = {}
    for key, value in table[1:]:
        e = Entry.FromString(value)
        raw = data[e.offset:e.offset + e.size]
        assert e.crc32c == (tb.masked_crc32c(raw) & 0xFFFFFFFF), key
        want = np.asarray(tensors[key.decode()])
        assert [d.size for d in e.shape.dim] == list(want.shape) and raw == want.tobytes()
        seen[key.decode()] = e.dtype
    assert sorted(seen) == sorted(tensors) and seen["global_step"] == tfc.DT_INT64 and seen["separator/conv1d/bias"] == tfc.DT_FLOAT
    st = text_format.Parse(open(os.path.join(os.path.dirname(prefix), "checkpoint")).read(), State())
    assert st.model_checkpoint_path == "run-1234" and list(st.all_model_checkpoint_paths) == ["run-1234"]
    assert tfc.latest_checkpoint(os.path.dirname(prefix)) == prefix
