"""Minimal HDF5 writer -- TEST INFRASTRUCTURE ONLY.

Enough of the format to synthesise fast5 files the bundled examples do not cover (multi-read layout ``read_*/Raw/Signal``,
chiron/utils/extract_sig_ref.py:178-193): superblock v0, version-1 object headers, compact groups made of Link messages,
contiguous 1-D integer datasets, fixed-length string and scalar numeric attributes.  Written from the HDF5 file-format
specification, independently of chiron_b200/fast5.py (which it is used to test)."""
import struct

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _msg(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _dtype_msg(dt: np.dtype) -> bytes:
    if dt.kind in "iu":
        bits0 = (0x08 if dt.kind == "i" else 0) | (1 if dt.byteorder == ">" else 0)
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f":
        # IEEE little-endian float: class 1; bit field: mantissa normalisation 2 (implied msb) << 4, sign location in byte 1
        size = dt.itemsize
        sign, exp_loc, exp_sz, man_sz, bias = (31, 23, 8, 23, 127) if size == 4 else (63, 52, 11, 52, 1023)
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, sign, 0, size, 0, 8 * size, exp_loc, exp_sz, 0, man_sz, bias)
    raise TypeError(dt)


def _attr_msg(name: str, value) -> bytes:
    nm = name.encode() + b"\x00"
    if isinstance(value, str):
        raw = value.encode() + b"\x00"
        dt = struct.pack("<BBBBI", 0x13, 0, 0, 0, len(raw))                  # class 3 string, null terminated, ASCII
    else:
        arr = np.asarray(value)
        raw = arr.tobytes()
        dt = _dtype_msg(arr.dtype)
    ds = struct.pack("<BBB5x", 1, 0, 0)                                        # scalar dataspace
    return _msg(0x0C, struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + raw)


class H5Writer:
    def __init__(self):
        self.buf = bytearray(96)                   # superblock (56) + root symbol table entry (40)

    def _alloc(self, data: bytes) -> int:
        self.buf += b"\x00" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def _header(self, msgs) -> int:
        body = b"".join(msgs)
        return self._alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)

    def dataset(self, arr: np.ndarray, attrs=None) -> int:
        arr = np.ascontiguousarray(arr)
        data_addr = self._alloc(arr.tobytes()) if arr.size else _UNDEF
        msgs = [_msg(0x01, struct.pack("<BBB5xQ", 1, 1, 0, arr.shape[0])), _msg(0x03, _dtype_msg(arr.dtype)),
                _msg(0x08, struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes))]
        msgs += [_attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._header(msgs)

    def group(self, children: dict, attrs=None) -> int:
        msgs = []
        for name, addr in children.items():
            nm = name.encode()
            msgs.append(_msg(0x06, struct.pack("<BBB", 1, 0, len(nm)) + nm + struct.pack("<Q", addr)))
        msgs += [_attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._header(msgs)

    def finish(self, root_addr: int) -> bytes:
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQII16x", 0, root_addr, 0, 0)
        self.buf[:len(sb)] = sb
        return bytes(self.buf)


def write_multi_read_fast5(path: str, reads: dict):
    """reads: {read_key: (int16 signal, read_id or None)} -> ``<read_key>/Raw/Signal`` with attr read_id on Raw."""
    w = H5Writer()
    top = {}
    for key, (sig, read_id) in reads.items():
        ds = w.dataset(np.asarray(sig, dtype="<i2"))
        raw = w.group({"Signal": ds}, {"read_id": read_id} if read_id is not None else None)
        top[key] = w.group({"Raw": raw})
    with open(path, "wb") as f:
        f.write(w.finish(w.group(top)))


def write_single_read_fast5(path: str, sig, read_number: int = 7, read_id=None, channel=None):
    """``/Raw/Reads/Read_<n>/Signal`` (+ ``/UniqueGlobalKey/channel_id`` attributes) as MinKNOW writes single-read files."""
    w = H5Writer()
    ds = w.dataset(np.asarray(sig, dtype="<i2"))
    attrs = {"read_number": np.int32(read_number)}
    if read_id is not None:
        attrs["read_id"] = read_id
    read = w.group({"Signal": ds}, attrs)
    raw = w.group({"Reads": w.group({"Read_%d" % read_number: read})})
    top = {"Raw": raw}
    if channel:
        top["UniqueGlobalKey"] = w.group({"channel_id": w.group({}, channel)})
    with open(path, "wb") as f:
        f.write(w.finish(w.group(top)))
