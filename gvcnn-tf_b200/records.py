"""Record / feature I/O on either side of the path - no TensorFlow needed.

1. The reference's multi-view dataset format: GZIP-compressed TFRecord files of ``tf.train.Example``s written
   by ``dataset_tools/create_modelnet_tf_record.py:119-129`` (``image/encoded`` = V PNGs, ``image/label``,
   ``image/filename`` ...) and read by ``train_data.py:47-54`` / ``eval_data.py:50-57``.  Implemented here from
   the public formats: TFRecord framing (length, masked CRC-32C, payload, masked CRC-32C) and the protobuf
   wire encoding of Example { Features { map<string, Feature{bytes_list|float_list|int64_list}> } }.
2. A pre-extracted feature shard format for feeding the grouping/fusion path at rate: per shard three ``.npy``
   files (raw [N, V, C_raw], final [N, V, D], labels [N]) opened memory-mapped, batches staged through reused
   pinned host buffers - the form ``gvcnn_grouping_fusion_host`` takes.
"""
from __future__ import annotations

import gzip
import io
import os
import struct
from typing import Dict, Iterator, List, Sequence, Union

import numpy as np

# ---------------------------------------------------------------- CRC-32C (Castagnoli), as TFRecord uses it
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------- TFRecord framing
class TFRecordWriter:
    """``tf.io.TFRecordWriter(path, options='GZIP')`` (create_modelnet_tf_record.py writes GZIP records)."""

    def __init__(self, path: str, gzip_compressed: bool = True):
        self._f = gzip.open(path, "wb") if gzip_compressed else open(path, "wb")

    def write(self, record: bytes):
        header = struct.pack("<Q", len(record))
        self._f.write(header + struct.pack("<I", masked_crc(header)) + record + struct.pack("<I", masked_crc(record)))

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_tfrecords(path: str, gzip_compressed: bool = True, verify_crc: bool = False) -> Iterator[bytes]:
    """Yields the serialized records of one file (``tf.data.TFRecordDataset(path, compression_type='GZIP')``,
    train_data.py:22)."""
    with (gzip.open(path, "rb") if gzip_compressed else open(path, "rb")) as f:
        while True:
            header = f.read(8)
            if not header:
                return
            if len(header) != 8:
                raise IOError("truncated TFRecord length in %s" % path)
            (n,) = struct.unpack("<Q", header)
            (hcrc,) = struct.unpack("<I", f.read(4))
            data = f.read(n)
            tail = f.read(4)
            if len(data) != n or len(tail) != 4:
                raise IOError("truncated TFRecord payload in %s" % path)
            if verify_crc:
                if hcrc != masked_crc(header) or struct.unpack("<I", tail)[0] != masked_crc(data):
                    raise IOError("TFRecord CRC mismatch in %s" % path)
            yield data


# ---------------------------------------------------------------- protobuf wire format of tf.train.Example
def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _read_varint(buf: bytes, pos: int):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _len_delim(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


Feature = Union[Sequence[bytes], Sequence[int], Sequence[float]]


def encode_example(features: Dict[str, Feature]) -> bytes:
    """Serializes {key: list of bytes | ints | floats} the way tf.train.Example does (int64/float lists packed,
    map entries in sorted key order - a deterministic serialization of the same message)."""
    entries = b""
    for key in sorted(features):
        vals = list(features[key])
        if vals and isinstance(vals[0], (bytes, bytearray)):
            feat = _len_delim(1, b"".join(_len_delim(1, bytes(v)) for v in vals))              # BytesList
        elif vals and isinstance(vals[0], float):
            feat = _len_delim(2, _len_delim(1, struct.pack("<%df" % len(vals), *vals)))         # FloatList
        else:
            feat = _len_delim(3, _len_delim(1, b"".join(_varint(int(v)) for v in vals)))        # Int64List
        entries += _len_delim(1, _len_delim(1, key.encode("utf8")) + _len_delim(2, feat))       # map entry
    return _len_delim(1, entries)                                                               # Example.features


def _fields(buf: bytes):
    pos = 0
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 2:
            n, pos = _read_varint(buf, pos)
            yield field, wt, buf[pos:pos + n]
            pos += n
        elif wt == 0:
            v, pos = _read_varint(buf, pos)
            yield field, wt, v
        elif wt == 5:
            yield field, wt, buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            yield field, wt, buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)


def decode_example(record: bytes) -> Dict[str, list]:
    """Parses a serialized tf.train.Example into {key: list} (bytes, ints or floats); accepts packed and
    unpacked repeated scalars (``tf.io.parse_single_example``, train_data.py:47-54)."""
    out: Dict[str, list] = {}
    for f, _, features in _fields(record):
        if f != 1:
            continue
        for ef, _, entry in _fields(features):
            if ef != 1:
                continue
            key, feat = None, b""
            for kf, _, v in _fields(entry):
                if kf == 1:
                    key = v.decode("utf8")
                elif kf == 2:
                    feat = v
            vals: list = []
            for kind, _, lst in _fields(feat):
                for vf, wt, v in _fields(lst):
                    if vf != 1:
                        continue
                    if kind == 1:
                        vals.append(bytes(v))
                    elif kind == 2:
                        vals.extend(struct.unpack("<%df" % (len(v) // 4), v) if wt == 2 else struct.unpack("<f", v))
                    elif kind == 3:
                        if wt == 2:
                            p = 0
                            while p < len(v):
                                x, p = _read_varint(v, p)
                                vals.append(x - (1 << 64) if x >> 63 else x)
                        else:
                            vals.append(v - (1 << 64) if v >> 63 else v)
            out[key] = vals
    return out


# ---------------------------------------------------------------- the reference's multi-view examples
def multiview_example(png_views: Sequence[bytes], label: int, filenames: Sequence[str],
                      heights: Sequence[int], widths: Sequence[int]) -> bytes:
    """One shape = V views, with exactly the keys of create_modelnet_tf_record.py:119-129."""
    import hashlib
    names = [f.encode("utf8") for f in filenames]
    return encode_example({
        "image/height": list(heights), "image/width": list(widths),
        "image/filename": names, "image/source_id": names,
        "image/key/sha256": [hashlib.sha256(p).hexdigest().encode("utf8") for p in png_views],
        "image/encoded": [bytes(p) for p in png_views],
        "image/format": [b"PNG"] * len(png_views),
        "image/label": [int(label)],
    })


def read_multiview(path: str, num_views: int, decode_png: bool = True, verify_crc: bool = False):
    """Yields (views, label, filenames) per shape; ``views`` is a list of V uint8 arrays [H, W, 3]
    (``tf.image.decode_png(img, channels=3)``, train_data.py:60) or the raw PNG bytes."""
    for rec in read_tfrecords(path, verify_crc=verify_crc):
        ex = decode_example(rec)
        pngs = ex["image/encoded"]
        if len(pngs) != num_views:                      # FixedLenFeature([num_views]) would fail the same way
            raise ValueError("example has %d views, expected %d" % (len(pngs), num_views))
        if decode_png:
            from PIL import Image
            views = [np.asarray(Image.open(io.BytesIO(p)).convert("RGB")) for p in pngs]
        else:
            views = pngs
        yield views, int(ex["image/label"][0]), [n.decode("utf8") for n in ex.get("image/filename", [])]


# ---------------------------------------------------------------- pre-extracted feature shards
def write_feature_shard(prefix: str, raw: np.ndarray, final: np.ndarray, labels: np.ndarray):
    """raw [N, V, C_raw], final [N, V, ...], labels [N] -> prefix.{raw,final,labels}.npy"""
    if raw.shape[:2] != final.shape[:2] or raw.shape[0] != labels.shape[0]:
        raise ValueError("raw, final and labels must agree in N (and V)")
    np.save(prefix + ".raw.npy", np.ascontiguousarray(raw))
    np.save(prefix + ".final.npy", np.ascontiguousarray(final))
    np.save(prefix + ".labels.npy", np.ascontiguousarray(labels))


class FeatureShard:
    """Memory-mapped shard; ``batches`` stages consecutive shapes through a small ring of reused pinned buffers so
    the host->device copies of ``gvcnn_grouping_fusion_host`` (or ``.cuda(non_blocking=True)``) run at PCIe rate.

    Buffer reuse is ordered against the consumer's copies: before buffer k is refilled, the generator waits on the
    CUDA event recorded (on the consumer's current stream) when the consumer came back for the next batch - i.e.
    after it queued its ``.cuda(non_blocking=True)`` copy of the batch that lives in buffer k.  A loop that never
    synchronises can therefore not have a pending async H2D read overwritten."""

    def __init__(self, prefix: str):
        self.raw = np.load(prefix + ".raw.npy", mmap_mode="r")
        self.final = np.load(prefix + ".final.npy", mmap_mode="r")
        self.labels = np.load(prefix + ".labels.npy", mmap_mode="r")
        if self.raw.shape[:2] != self.final.shape[:2] or self.raw.shape[0] != self.labels.shape[0]:
            raise ValueError("inconsistent shard %s" % prefix)

    def __len__(self):
        return int(self.raw.shape[0])

    def batches(self, batch_size: int, pin: bool = True, lo: int = 0, hi: int = None, depth: int = 3):
        import torch
        hi = len(self) if hi is None else hi
        pin = pin and torch.cuda.is_available()
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pin)
        rt = torch.from_numpy(np.zeros(0, self.raw.dtype)).dtype
        ft = torch.from_numpy(np.zeros(0, self.final.dtype)).dtype
        depth = max(2, depth)
        bufs = [(mk((batch_size,) + self.raw.shape[1:], rt), mk((batch_size,) + self.final.shape[1:], ft))
                for _ in range(depth)]
        events = [None] * depth                                      # consumer-side "copies of this buffer are queued"
        for n, i in enumerate(range(lo, hi, batch_size)):
            j = min(i + batch_size, hi)
            k = n % depth
            if events[k] is not None:
                events[k].synchronize()                              # the H2D copies that read buffer k have finished
            r, f = bufs[k]
            r[:j - i].copy_(torch.from_numpy(np.array(self.raw[i:j])))
            f[:j - i].copy_(torch.from_numpy(np.array(self.final[i:j])))
            yield r[:j - i], f[:j - i], torch.from_numpy(np.asarray(self.labels[i:j]).copy())
            if pin:                                                  # back from the consumer: its copies are queued
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                events[k] = ev
