"""Dependency-free reader for the HDF5 subset used by ONT fast5 files (h5py / libhdf5 are not available).

Replaces the h5py calls of chiron/chiron_input.py:541-547 and chiron/utils/extract_sig_ref.py:92-193.  Supported
(verified on every bundled fast5): superblock v0/v1, version-1 object headers with continuation
blocks, old-style groups (symbol table: v1 B-tree + local heap) and new-style compact groups (Link messages), datasets
with contiguous or chunked layout (v1 chunk B-tree) and the deflate / shuffle / fletcher32 filters, fixed-point and
floating-point element types, version-1/2/3 attributes with numeric or string values (fixed or variable length)."""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Fast5Error(IOError):
    pass


class _Dataset:
    def __init__(self, f, msgs):
        self.f = f
        self.msgs = msgs


class _File:
    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != _SIG:
            raise Fast5Error("%s is not an HDF5 file" % path)
        ver = b[8]
        if ver not in (0, 1):
            raise Fast5Error("unsupported HDF5 superblock version %d" % ver)
        self.O, self.L = b[13], b[14]
        if self.O != 8 or self.L != 8:
            raise Fast5Error("unsupported offset/length sizes %d/%d" % (self.O, self.L))
        pos = 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", b, pos)[0]
        pos += 4 * 8                                    # base, free-space, EOF, driver-info addresses
        # root symbol table entry: link name offset, object header address, cache type, reserved, scratch
        self.root_header = struct.unpack_from("<Q", b, pos + 8)[0]

    # ---- object headers ------------------------------------------------------------------------------------------
    def messages(self, addr: int) -> List[Tuple[int, bytes]]:
        b = self.buf
        if b[addr] != 1:
            raise Fast5Error("unsupported object header version %d at %d" % (b[addr], addr))
        n_msgs, = struct.unpack_from("<H", b, addr + 2)
        hdr_size, = struct.unpack_from("<I", b, addr + 8)
        out: List[Tuple[int, bytes]] = []
        blocks = [(addr + 16, hdr_size)]
        while blocks and len(out) < n_msgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < n_msgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:                        # continuation
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((caddr, clen))
                out.append((mtype, data))
        return out

    # ---- groups ------------------------------------------------------------------------------------------------------
    def children(self, addr: int) -> Dict[str, int]:
        kids: Dict[str, int] = {}
        for mtype, data in self.messages(addr):
            if mtype == 0x11:                            # symbol table message: B-tree + local heap
                btree, heap = struct.unpack_from("<QQ", data, 0)
                heap_data = self._local_heap(heap)
                self._walk_group_btree(btree, heap_data, kids)
            elif mtype == 0x06:                          # link message (new-style compact group)
                name, target = self._link(data)
                if target is not None:
                    kids[name] = target
        return kids

    def _local_heap(self, addr: int) -> int:
        if self.buf[addr:addr + 4] != b"HEAP":
            raise Fast5Error("bad local heap signature")
        return struct.unpack_from("<Q", self.buf, addr + 8 + 2 * 8)[0]

    def _walk_group_btree(self, addr: int, heap_data: int, kids: Dict[str, int]):
        b = self.buf
        if b[addr:addr + 4] == b"TREE":
            level, n = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
            pos = addr + 8 + 16                          # skip sibling addresses
            for i in range(n):
                child = struct.unpack_from("<Q", b, pos + 8)[0]      # key(L) child(O)
                pos += 16
                if level > 0:
                    self._walk_group_btree(child, heap_data, kids)
                else:
                    self._snod(child, heap_data, kids)
        elif b[addr:addr + 4] == b"SNOD":
            self._snod(addr, heap_data, kids)
        else:
            raise Fast5Error("bad group B-tree node")

    def _snod(self, addr: int, heap_data: int, kids: Dict[str, int]):
        b = self.buf
        if b[addr:addr + 4] != b"SNOD":
            raise Fast5Error("bad symbol table node")
        n, = struct.unpack_from("<H", b, addr + 6)
        pos = addr + 8
        for _ in range(n):
            name_off, header = struct.unpack_from("<QQ", b, pos)
            pos += 2 * 8 + 24
            s = heap_data + name_off
            e = b.index(b"\x00", s)
            kids[b[s:e].decode("utf-8")] = header

    def _link(self, data: bytes) -> Tuple[str, Optional[int]]:
        flags = data[1]
        pos = 2
        ltype = 0
        if flags & 0x08:
            ltype = data[pos]
            pos += 1
        if flags & 0x04:
            pos += 8
        if flags & 0x10:
            pos += 1
        nlen_size = 1 << (flags & 3)
        nlen = int.from_bytes(data[pos:pos + nlen_size], "little")
        pos += nlen_size
        name = data[pos:pos + nlen].decode("utf-8")
        pos += nlen
        if ltype != 0:
            return name, None                            # soft / external links are not followed
        return name, struct.unpack_from("<Q", data, pos)[0]

    def resolve(self, path: str) -> int:
        addr = self.root_header
        for part in [p for p in path.split("/") if p]:
            kids = self.children(addr)
            if part not in kids:
                raise KeyError(path)
            addr = kids[part]
        return addr

    # ---- datasets ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _dtype(data: bytes) -> Tuple[Optional[np.dtype], int, int]:
        cls = data[0] & 0x0F
        bits0 = data[1]
        size, = struct.unpack_from("<I", data, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            signed = (bits0 >> 3) & 1
            return np.dtype("%s%s%d" % (order, "i" if signed else "u", size)), cls, size
        if cls == 1:
            return np.dtype("%sf%d" % (order, size)), cls, size
        return None, cls, size                           # 3 = fixed string, 9 = variable length

    @staticmethod
    def _dataspace(data: bytes) -> List[int]:
        ver, rank = data[0], data[1]
        pos = 8 if ver == 1 else 4
        return [struct.unpack_from("<Q", data, pos + 8 * i)[0] for i in range(rank)]

    def read_dataset(self, addr: int) -> np.ndarray:
        msgs = self.messages(addr)
        dims = dtype = layout = None
        filters: List[int] = []
        for mtype, data in msgs:
            if mtype == 0x01:
                dims = self._dataspace(data)
            elif mtype == 0x03:
                dtype, _, esize = self._dtype(data)
            elif mtype == 0x08:
                layout = data
            elif mtype == 0x0B:
                filters = self._filters(data)
        if dims is None or dtype is None or layout is None:
            raise Fast5Error("object at %d is not a numeric dataset" % addr)
        count = int(np.prod(dims)) if dims else 1
        if layout[0] != 3:
            raise Fast5Error("unsupported data layout message version %d" % layout[0])
        lclass = layout[1]
        if lclass == 0:                                  # compact
            size, = struct.unpack_from("<H", layout, 2)
            raw = layout[4:4 + size]
        elif lclass == 1:                                # contiguous
            daddr, size = struct.unpack_from("<QQ", layout, 2)
            raw = b"" if daddr == _UNDEF else self.buf[daddr:daddr + size]
        elif lclass == 2:                                # chunked
            ndims = layout[2]
            btree, = struct.unpack_from("<Q", layout, 3)
            cdims = struct.unpack_from("<%dI" % ndims, layout, 11)
            raw = self._read_chunks(btree, ndims, cdims, dims, dtype.itemsize, filters)
        else:
            raise Fast5Error("unsupported layout class %d" % lclass)
        return np.frombuffer(raw, dtype=dtype, count=count).reshape(dims if dims else ())

    @staticmethod
    def _filters(data: bytes) -> List[int]:
        ver, n = data[0], data[1]
        pos = 8 if ver == 1 else 2
        ids = []
        for _ in range(n):
            fid, = struct.unpack_from("<H", data, pos)
            if ver == 1 or fid >= 256:
                nlen, = struct.unpack_from("<H", data, pos + 2)
                pos += 4
            else:
                nlen = 0
                pos += 2
            _flags, ncd = struct.unpack_from("<HH", data, pos)
            pos += 4
            pos += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            pos += 4 * ncd
            if ver == 1 and ncd % 2:
                pos += 4
            ids.append(fid)
        return ids

    def _read_chunks(self, btree: int, ndims: int, cdims, dims, esize: int, filters: List[int]) -> bytes:
        if len(dims) != 1:
            raise Fast5Error("only 1-D chunked datasets are supported")
        total = dims[0]
        out = bytearray(total * esize)
        chunk_elems = cdims[0]

        def walk(addr):
            b = self.buf
            if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 1:
                raise Fast5Error("bad chunk B-tree node")
            level, n = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
            pos = addr + 8 + 16
            key_size = 8 + 8 * ndims
            for _ in range(n):
                csize, fmask = struct.unpack_from("<II", b, pos)
                offs = struct.unpack_from("<%dQ" % ndims, b, pos + 8)
                child, = struct.unpack_from("<Q", b, pos + key_size)
                pos += key_size + 8
                if level > 0:
                    walk(child)
                    continue
                data = b[child:child + csize]
                for i, fid in reversed(list(enumerate(filters))):
                    if fmask >> i & 1:
                        continue
                    if fid == 1:
                        data = zlib.decompress(data)
                    elif fid == 2:                           # shuffle
                        arr = np.frombuffer(data, dtype=np.uint8)
                        nel = len(arr) // esize
                        data = arr[:nel * esize].reshape(esize, nel).T.tobytes()
                    elif fid == 3:                           # fletcher32: checksum trails the data
                        data = data[:-4]
                    else:
                        raise Fast5Error("unsupported HDF5 filter %d" % fid)
                start = offs[0]
                nvalid = min(chunk_elems, total - start)
                out[start * esize:(start + nvalid) * esize] = data[:nvalid * esize]

        walk(btree)
        return bytes(out)

    # ---- attributes --------------------------------------------------------------------------------------------------
    def attrs(self, addr: int) -> Dict[str, object]:
        out: Dict[str, object] = {}
        for mtype, data in self.messages(addr):
            if mtype != 0x0C:
                continue
            ver = data[0]
            name_sz, dt_sz, ds_sz = struct.unpack_from("<HHH", data, 2)
            pos = 8 + (1 if ver == 3 else 0)
            pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
            name = data[pos:pos + name_sz].split(b"\x00")[0].decode("utf-8")
            pos += pad(name_sz)
            dt = data[pos:pos + dt_sz]
            pos += pad(dt_sz)
            ds = data[pos:pos + ds_sz]
            pos += pad(ds_sz)
            val = data[pos:]
            dtype, cls, size = self._dtype(dt)
            dims = self._dataspace(ds) if ds_sz >= 2 else []
            count = int(np.prod(dims)) if dims else 1
            if dtype is not None:
                arr = np.frombuffer(val, dtype=dtype, count=count)
                out[name] = arr[0].item() if not dims else arr.copy()
            elif cls == 3:
                out[name] = val[:size].split(b"\x00")[0].decode("utf-8", "replace")
            elif cls == 9:                               # variable length: (length u32, global heap address, index u32)
                try:
                    out[name] = self._vlen_string(val)
                except Exception:
                    out[name] = None
        return out

    def _vlen_string(self, val: bytes) -> str:
        length, gaddr, index = struct.unpack_from("<IQI", val, 0)
        b = self.buf
        if b[gaddr:gaddr + 4] != b"GCOL":
            raise Fast5Error("bad global heap")
        csize, = struct.unpack_from("<Q", b, gaddr + 8)
        pos, end = gaddr + 16, gaddr + csize
        while pos + 16 <= end:
            idx, _ref, _res, osize = struct.unpack_from("<HHIQ", b, pos)
            if idx == index:
                return b[pos + 16:pos + 16 + length].decode("utf-8", "replace")
            if idx == 0:
                break
            pos += 16 + (osize + 7) // 8 * 8
        raise KeyError("global heap object %d" % index)


# ---- fast5 conventions ------------------------------------------------------------------------------------------------
def _first_read_group(f: _File) -> Tuple[str, int]:
    reads = f.children(f.resolve("/Raw/Reads"))
    if not reads:
        raise Fast5Error("no reads under /Raw/Reads")
    name = sorted(reads)[0]
    return name, reads[name]


def read_raw_signal(path: str) -> np.ndarray:
    """``list(root['/Raw/Reads'].values())[0]['Signal']`` (chiron_input.py:546-547): raw DAC integers."""
    f = _File(path)
    _, grp = _first_read_group(f)
    return f.read_dataset(f.children(grp)["Signal"])


def read_fast5(path: str) -> List[dict]:
    """Every read of a single-read (``/Raw/Reads/Read_N``, extract_sig_ref.py:149-175) or multi-read
    (``read_*/Raw``, :178-193) fast5: [{read_key, signal, read_id (None if the attribute is absent), channel}]."""
    f = _File(path)
    root = f.children(f.root_header)
    out = []
    if "Raw" in root:
        name, grp = _first_read_group(f)
        attrs = f.attrs(grp)
        chan = {}
        try:
            chan = f.attrs(f.resolve("/UniqueGlobalKey/channel_id"))
        except KeyError:
            pass
        out.append({"read_key": name, "signal": f.read_dataset(f.children(grp)["Signal"]),
                    "read_id": attrs.get("read_id"), "attrs": attrs, "channel": chan})
    else:
        for key in sorted(root):
            kids = f.children(root[key])
            if "Raw" not in kids:
                continue
            raw = kids["Raw"]
            attrs = f.attrs(raw)
            out.append({"read_key": key, "signal": f.read_dataset(f.children(raw)["Signal"]),
                        "read_id": attrs.get("read_id"), "attrs": attrs, "channel": {}})
    return out
