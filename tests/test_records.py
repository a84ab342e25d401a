"""Record / feature I/O: the reference's GZIP TFRecord of multi-view tf.train.Examples, read
and written without TensorFlow, and the pre-extracted feature shards."""
import io
import os

import numpy as np
import pytest

from gvcnn_tf_b200 import records as rec


def test_crc32c_known_answers():
    assert rec.crc32c(b"") == 0
    assert rec.crc32c(b"123456789") == 0xE3069283            # the CRC-32C check value
    assert rec.crc32c(b"\x00" * 32) == 0x8A9136AA            # RFC 3720 B.4


def test_example_wire_format_golden():
    # Example{features{feature{"a": Int64List[1]}}}: hand-encoded per the protobuf wire spec
    assert rec.encode_example({"a": [1]}) == b"\n\x0c\n\n\n\x01a\x12\x05\x1a\x03\n\x01\x01"
    ex = rec.decode_example(rec.encode_example({"image/label": [-3], "f": [0.5, 2.0], "b": [b"xy", b""]}))
    assert ex == {"image/label": [-3], "f": [0.5, 2.0], "b": [b"xy", b""]}
    # unpacked int64 / float encodings (older writers) decode too
    unpacked = b"\n\x0d\n\x0b\n\x01a\x12\x06\x1a\x04\x08\x01\x08\x02"
    assert rec.decode_example(unpacked) == {"a": [1, 2]}


def _png(rgb):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(rgb).save(buf, format="PNG")
    return buf.getvalue()


def test_multiview_tfrecord_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    V, n = 6, 3
    path = str(tmp_path / "train.record")
    shapes = []
    with rec.TFRecordWriter(path) as w:
        for i in range(n):
            views = [rng.integers(0, 255, (8, 9, 3), dtype=np.uint8) for _ in range(V)]
            names = ["airplane_%04d.%d.png" % (i, v) for v in range(V)]
            w.write(rec.multiview_example([_png(v) for v in views], i % 5, names, [8] * V, [9] * V))
            shapes.append((views, i % 5, names))
    # the keys of dataset_tools/create_modelnet_tf_record.py:119-129
    ex = rec.decode_example(next(rec.read_tfrecords(path, verify_crc=True)))
    assert sorted(ex) == ["image/encoded", "image/filename", "image/format", "image/height", "image/key/sha256",
                          "image/label", "image/source_id", "image/width"]
    assert ex["image/format"] == [b"PNG"] * V and len(ex["image/key/sha256"][0]) == 64
    got = list(rec.read_multiview(path, V, verify_crc=True))
    assert len(got) == n
    for (views, label, names), (gv, gl, gn) in zip(shapes, got):
        assert gl == label and gn == names
        for a, b in zip(views, gv):
            np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):
        next(rec.read_multiview(path, 12))                    # FixedLenFeature([num_views]) mismatch
    # corruption is detected
    import gzip
    raw = bytearray(gzip.open(path, "rb").read())
    raw[40] ^= 0xFF
    bad = str(tmp_path / "bad.record")
    with gzip.open(bad, "wb") as f:
        f.write(bytes(raw))
    with pytest.raises(IOError):
        list(rec.read_tfrecords(bad, verify_crc=True))


def test_feature_shard_batches(tmp_path):
    rng = np.random.default_rng(1)
    N, V, Cr, D = 11, 6, 32, 64
    raw = rng.standard_normal((N, V, Cr)).astype(np.float32)
    final = rng.standard_normal((N, V, 2, 2, D // 4)).astype(np.float32)
    labels = rng.integers(0, 5, N)
    prefix = str(tmp_path / "shard0")
    rec.write_feature_shard(prefix, raw, final, labels)
    sh = rec.FeatureShard(prefix)
    assert len(sh) == N
    seen = 0
    for r, f, y in sh.batches(4, pin=False):
        np.testing.assert_array_equal(r.numpy(), raw[seen:seen + len(y)])
        np.testing.assert_array_equal(f.numpy(), final[seen:seen + len(y)])
        np.testing.assert_array_equal(y.numpy(), labels[seen:seen + len(y)])
        seen += len(y)
    assert seen == N
    with pytest.raises(ValueError):
        rec.write_feature_shard(prefix, raw[:3], final, labels)
