"""Dependency-free reader for TensorFlow "bundle" checkpoints (``*.index`` + ``*.data-00000-of-00001``)
and for the conv attributes of a ``*.meta`` MetaGraphDef.

The reference restores its weights with ``tf.train.Saver.restore`` (chiron/chiron_eval.py:272-276);
TensorFlow is not available to this project, so the shipped checkpoints are converted offline (see
``chiron_b200/convert_weights.py``) using this hand-written parser.  Format notes: SURVEY.md App. B.

 * ``*.index`` is a LevelDB-style SSTable: 48-byte footer (metaindex handle, index handle, magic), an index
   block pointing at data blocks, every block a run of prefix-compressed ``(shared, non_shared, value_len,
   key_suffix, value)`` entries followed by a restart array and a 5-byte trailer (compression type + crc).
 * every value is a ``BundleEntryProto`` {1:dtype 2:shape{2:dim{1:size}} 3:shard_id 4:offset 5:size 6:crc32c}.
 * ``*.data-*`` holds the raw little-endian tensors.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, List, Tuple

import numpy as np

_SSTABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _proto_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) for one protobuf message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, val


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_block(buf: bytes, offset: int, size: int) -> bytes:
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError("compressed SSTable blocks are not supported (type %d)" % ctype)
    return buf[offset:offset + size]


def read_index(index_path: str) -> Dict[str, dict]:
    """Return {variable name: {dtype, shape, offset, size, shard}} for a bundle ``.index`` file."""
    with open(index_path, "rb") as f:
        buf = f.read()
    footer = buf[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != _SSTABLE_MAGIC:
        raise ValueError("%s is not an SSTable" % index_path)
    pos = 0
    _, pos = _varint(footer, pos)
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries: Dict[str, dict] = {}
    for _, handle in _block_entries(_read_block(buf, idx_off, idx_size)):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for key, value in _block_entries(_read_block(buf, off, size)):
            if key == b"":
                continue  # BundleHeaderProto
            ent = {"dtype": None, "shape": [], "shard": 0, "offset": 0, "size": 0}
            for field, _, val in _proto_fields(value):
                if field == 1:
                    ent["dtype"] = val
                elif field == 2:
                    for f2, _, dim in _proto_fields(val):
                        if f2 == 2:
                            size_ = 0
                            for f3, _, v3 in _proto_fields(dim):
                                if f3 == 1:
                                    size_ = v3
                            ent["shape"].append(size_)
                elif field == 3:
                    ent["shard"] = val
                elif field == 4:
                    ent["offset"] = val
                elif field == 5:
                    ent["size"] = val
            entries[key.decode("utf-8")] = ent
    return entries


def latest_checkpoint(model_dir: str) -> str:
    """``tf.train.latest_checkpoint``: read the ``checkpoint`` text file (chiron/chiron_eval.py:276)."""
    with open(os.path.join(model_dir, "checkpoint")) as f:
        for line in f:
            if line.startswith("model_checkpoint_path:"):
                name = line.split(":", 1)[1].strip().strip('"')
                return os.path.join(model_dir, os.path.basename(name))
    raise FileNotFoundError("no model_checkpoint_path in %s/checkpoint" % model_dir)


def read_checkpoint(prefix: str, skip_optimizer: bool = True) -> Dict[str, np.ndarray]:
    """Load every tensor of a bundle checkpoint ``prefix`` (no TensorFlow needed)."""
    entries = read_index(prefix + ".index")
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        data = f.read()
    out: Dict[str, np.ndarray] = {}
    for name, ent in entries.items():
        if skip_optimizer and (name.endswith("/Adam") or name.endswith("/Adam_1") or
                               name in ("beta1_power", "beta2_power", "global_step")):
            continue
        if ent["dtype"] not in _DTYPES:
            continue
        dt = np.dtype(_DTYPES[ent["dtype"]]).newbyteorder("<")
        arr = np.frombuffer(data, dtype=dt, count=ent["size"] // dt.itemsize, offset=ent["offset"])
        out[name] = arr.reshape(ent["shape"]).copy()
    return out


def read_conv_attrs(meta_path: str) -> Dict[str, dict]:
    """Return {node name: {strides, padding}} for every Conv2D node of a ``.meta`` MetaGraphDef.

    The graph, not ``model.json``, is the truth for RNA_default (SURVEY.md finding 3)."""
    with open(meta_path, "rb") as f:
        buf = f.read()
    graph = None
    for field, wt, val in _proto_fields(buf):
        if field == 2 and wt == 2:
            graph = val
    if graph is None:
        raise ValueError("no GraphDef in %s" % meta_path)
    convs: Dict[str, dict] = {}
    for field, wt, node in _proto_fields(graph):
        if field != 1 or wt != 2:
            continue
        name = op = None
        attrs: List[bytes] = []
        for f2, w2, v2 in _proto_fields(node):
            if f2 == 1:
                name = v2.decode()
            elif f2 == 2:
                op = v2.decode()
            elif f2 == 5:
                attrs.append(v2)
        if op != "Conv2D":
            continue
        info = {"strides": None, "padding": None}
        for entry in attrs:
            k = v = None
            for f3, _, v3 in _proto_fields(entry):
                if f3 == 1:
                    k = v3.decode()
                elif f3 == 2:
                    v = v3
            if k == "strides":
                for f4, _, v4 in _proto_fields(v):
                    if f4 == 1:  # AttrValue.list
                        ints = []
                        for f5, w5, v5 in _proto_fields(v4):
                            if f5 == 3 and w5 == 2:  # packed repeated int64
                                p = 0
                                while p < len(v5):
                                    x, p = _varint(v5, p)
                                    ints.append(x)
                            elif f5 == 3:
                                ints.append(v5)
                        info["strides"] = ints
            elif k == "padding":
                for f4, _, v4 in _proto_fields(v):
                    if f4 == 2:
                        info["padding"] = v4.decode()
        convs[name] = info
    return convs
