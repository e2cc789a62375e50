"""Minimal NetCDF-4 (= HDF5) reader for the repwvl lookup tables.

Test/fixture infrastructure only.  It understands exactly the subset the
`Reduced{10,20,100}Forcing.nc` family uses:
superblock v0, version-2 object headers (OHDR/OCHK), dense link storage in a
fractal heap (link messages are recovered by walking the heap's direct blocks),
and contiguous, unfiltered, little-endian IEEE-754 f64 datasets.

It is an independent implementation of the product's C++ reader
(`our_first_climate_model_b200/csrc/host/nc4lite.cpp`); tests cross-check one
against the other.
"""
from __future__ import annotations

import re
import struct
from dataclasses import dataclass, field

import numpy as np


@dataclass
class _Obj:
    addr: int
    shape: tuple | None = None
    dtype: str | None = None
    data_addr: int | None = None
    data_size: int | None = None
    links: dict = field(default_factory=dict)


def _parse_messages(buf: bytes, start: int, end: int, crt_order: bool, obj: _Obj, file: bytes):
    p = start
    while p + 4 <= end:
        mtype = buf[p]
        msize = struct.unpack_from("<H", buf, p + 1)[0]
        p += 4
        if crt_order:
            p += 2
        body = p
        if mtype == 0x01:  # dataspace
            ver, rank, flags = buf[body], buf[body + 1], buf[body + 2]
            q = body + (4 if ver == 2 else 8)
            obj.shape = tuple(struct.unpack_from("<Q", buf, q + 8 * i)[0] for i in range(rank))
        elif mtype == 0x03:  # datatype
            cls = buf[body] & 0x0F
            bits0 = buf[body + 1]
            size = struct.unpack_from("<I", buf, body + 4)[0]
            if cls == 1 and size == 8 and (bits0 & 1) == 0:
                obj.dtype = "<f8"
            elif cls == 0 and (bits0 & 1) == 0:
                obj.dtype = "<i%d" % size if (bits0 & 8) else "<u%d" % size
            else:
                obj.dtype = "other"
        elif mtype == 0x08:  # data layout
            ver, cls = buf[body], buf[body + 1]
            if ver == 3 and cls == 1:
                obj.data_addr, obj.data_size = struct.unpack_from("<QQ", buf, body + 2)
        elif mtype == 0x06:  # link (compact storage)
            name, addr = _parse_link(buf, body)[:2]
            if name is not None:
                obj.links[name] = addr
        elif mtype == 0x10:  # continuation -> OCHK block
            off, length = struct.unpack_from("<QQ", buf, body)
            assert file[off:off + 4] == b"OCHK", "bad continuation block"
            _parse_messages(file, off + 4, off + length - 4, crt_order, obj, file)
        p = body + msize


def _parse_link(buf: bytes, p: int):
    """Link message, version 1.  Returns (name, target address, next offset)."""
    if buf[p] != 1:
        return None, None, p
    flags = buf[p + 1]
    q = p + 2
    ltype = 0
    if flags & 0x08:
        ltype = buf[q]; q += 1
    if flags & 0x04:
        q += 8
    if flags & 0x10:
        q += 1
    lsz = 1 << (flags & 3)
    nlen = int.from_bytes(buf[q:q + lsz], "little"); q += lsz
    name = buf[q:q + nlen].decode("utf-8", "replace"); q += nlen
    if ltype != 0:
        return None, None, q
    addr = struct.unpack_from("<Q", buf, q)[0]
    return name, addr, q + 8


def _object(file: bytes, addr: int) -> _Obj:
    assert file[addr:addr + 4] == b"OHDR" and file[addr + 4] == 2
    flags = file[addr + 5]
    p = addr + 6
    if flags & 0x20:
        p += 16
    if flags & 0x10:
        p += 4
    csz = 1 << (flags & 3)
    chunk0 = int.from_bytes(file[p:p + csz], "little"); p += csz
    obj = _Obj(addr)
    _parse_messages(file, p, p + chunk0, bool(flags & 0x04), obj, file)
    return obj


def _heap_links(file: bytes) -> dict:
    """Walk every fractal-heap direct block and decode the link messages in it."""
    links = {}
    for m in re.finditer(b"FRHP", file):
        h = m.start()
        if file[h + 4] != 0:
            continue
        # header: sig4 ver1 idlen2 filtlen2 flags1 maxmanaged4 ... maxheapsize(bits) at +118? -> parse fields
        flags = file[h + 9]
        # fixed layout up to "maximum heap size" (2 bytes) for 8-byte offsets/lengths
        max_heap_bits = struct.unpack_from("<H", file, h + 4 + 1 + 2 + 2 + 1 + 4 + 8 * 12 + 2 + 8 + 8)[0]
        boff = (max_heap_bits + 7) // 8
        for d in re.finditer(b"FHDB", file):
            b = d.start()
            if file[b + 4] != 0 or struct.unpack_from("<Q", file, b + 5)[0] != h:
                continue
            p = b + 5 + 8 + boff + (4 if flags & 0x02 else 0)
            while p < len(file) and file[p] == 1:
                name, addr, nxt = _parse_link(file, p)
                if name is None or nxt == p:
                    break
                links[name] = addr
                p = nxt
    return links


def read_nc(path: str) -> dict:
    """Return {variable name: float64 ndarray} for every contiguous f64 dataset."""
    file = open(path, "rb").read()
    assert file[:8] == b"\x89HDF\r\n\x1a\n" and file[8] == 0, "not an HDF5 superblock-v0 file"
    links = _heap_links(file)
    out = {}
    for name, addr in links.items():
        if file[addr:addr + 4] != b"OHDR":
            continue
        obj = _object(file, addr)
        if obj.dtype == "<f8" and obj.data_addr not in (None, 0xFFFFFFFFFFFFFFFF) and obj.shape:
            n = int(np.prod(obj.shape))
            if obj.data_size != 8 * n:
                continue
            out[name] = np.frombuffer(file, dtype="<f8", count=n, offset=obj.data_addr).reshape(obj.shape).copy()
            out[name + "@offset"] = obj.data_addr
    return out


if __name__ == "__main__":
    import sys
    for k, v in read_nc(sys.argv[1]).items():
        if isinstance(v, np.ndarray):
            print(k, v.shape, float(v.min()), float(v.max()))
        else:
            print(k, v)
