"""Minimal HDF5 reader / writer for the MATLAB v7.3 ``.mat`` files either side of the hot path, without h5py.

The reference reads its training / test tensors with ``h5py.File(fname, 'r')[key][()]`` (utils.py:29-54) and writes the warped
frames with ``hdf5storage.write(..., matlab_compatible=True)`` (FISR_tfoptflow/FISR_for_video_warp_img_with_flo.py:131-137).
Neither package is installable here, so this module restates the part of the published HDF5 file format those files use:

  reader  superblock v0/v1 (any user-block offset: MATLAB prepends 512 bytes) and v2/v3; version-1 object headers with
          continuation blocks (and the version-2 "OHDR" form); old-style groups (symbol table = local heap + B-tree v1 + SNOD)
          and link messages; dataspace v1/v2; fixed-point and IEEE float datatypes; data layout v1-v3: compact, contiguous,
          chunked through the B-tree v1 chunk index; filters deflate (1), shuffle (2) and fletcher32 (3).
  writer  a MATLAB-compatible file: 512-byte MATLAB header, superblock v0, one old-style root group, one dataset per array --
          contiguous, or chunked + (shuffle, deflate, fletcher32) like hdf5storage's defaults -- each with a ``MATLAB_class``
          attribute.

PARITY NOTE.  The reader is pinned by a file MATLAB itself wrote: ``testhdf5_7.4_GLNX86.mat`` from scipy's test data (a
``double`` 9 x 1 ramp 0:pi/4:2pi; old-style group, version-1 object header, layout v2) -- tests/test_hdf5_min.py.  The chunked /
filtered path and the writer are validated against this module's own counterpart only (no libhdf5 in the image to cross-check).
"""
from __future__ import annotations

import struct
import time
import zlib
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF

MSG_DATASPACE, MSG_LINKINFO, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LINK, MSG_LAYOUT = 0x1, 0x2, 0x3, 0x4, 0x5, 0x6, 0x8
MSG_FILTERS, MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMTAB, MSG_MTIME = 0xB, 0xC, 0x10, 0x11, 0x12


# =========================================================================================== reader
class H5Error(ValueError):
    pass


class H5File:
    """Read-only view of the datasets of an HDF5 file: ``H5File(path)[name]`` -> ndarray (dims as h5py reports them)."""

    def __init__(self, path: str):
        self.path = path
        self.buf = np.memmap(path, dtype=np.uint8, mode="r")
        self._find_superblock()
        self.links = self._group_links(self.root_header)

    # ---- low level
    def _u(self, off: int, n: int) -> int:
        return int.from_bytes(bytes(self.buf[off:off + n]), "little")

    def _bytes(self, off: int, n: int) -> bytes:
        return bytes(self.buf[off:off + n])

    def _addr(self, rel: int) -> int:
        return rel + self.base

    def _find_superblock(self) -> None:
        off, size = 0, len(self.buf)
        while off + 8 <= size:
            if self._bytes(off, 8) == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
        else:
            raise H5Error(f"{self.path}: no HDF5 signature (is it a MATLAB v5 .mat? use scipy.io.loadmat)")
        ver = self.buf[off + 8]
        if ver in (0, 1):
            self.so, self.sl = int(self.buf[off + 13]), int(self.buf[off + 14])
            p = off + 24 + (4 if ver == 1 else 0)
            self.base = self._u(p, self.so)
            p += 4 * self.so                                   # base, free-space info, end of file, driver info
            # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch pad
            self.root_header = self._u(p + self.so, self.so)
        elif ver in (2, 3):
            self.so, self.sl = int(self.buf[off + 9]), int(self.buf[off + 10])
            p = off + 12
            self.base = self._u(p, self.so)
            self.root_header = self._u(p + 3 * self.so, self.so)
        else:
            raise H5Error(f"unsupported superblock version {ver}")
        if self.base == 0 and off:                             # files whose addresses are relative to the signature
            self.base = off if ver in (0, 1) and self._looks_relative(off) else 0

    def _looks_relative(self, off: int) -> bool:
        a = self.root_header + off
        return a + 4 <= len(self.buf) and (self.buf[a] == 1 or self._bytes(a, 4) == b"OHDR")

    # ---- object headers
    def _messages(self, header: int) -> List[Tuple[int, int, bytes]]:
        """[(type, flags, payload)] of the object header at relative address ``header`` (v1 or v2), continuations followed."""
        a = self._addr(header)
        out: List[Tuple[int, int, bytes]] = []
        if self._bytes(a, 4) == b"OHDR":
            flags = int(self.buf[a + 5])
            p = a + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            csz = 1 << (flags & 3)
            chunk0 = self._u(p, csz)
            p += csz
            blocks = [(p, chunk0)]
            track = bool(flags & 0x04)
            while blocks:
                p, n = blocks.pop(0)
                end = p + n
                while p + 4 + (2 if track else 0) <= end:
                    t, sz, fl = int(self.buf[p]), self._u(p + 1, 2), int(self.buf[p + 3])
                    p += 4 + (2 if track else 0)
                    body = self._bytes(p, sz)
                    p += sz
                    if t == MSG_CONTINUATION:
                        off, ln = int.from_bytes(body[:self.so], "little"), int.from_bytes(body[self.so:self.so + self.sl], "little")
                        blocks.append((self._addr(off) + 4, ln - 8))          # skip "OCHK", drop the checksum
                    elif t != 0:
                        out.append((t, fl, body))
            return out
        if self.buf[a] != 1:
            raise H5Error(f"object header at {a:#x}: unknown version {int(self.buf[a])}")
        nmsg, hsize = self._u(a + 2, 2), self._u(a + 8, 4)
        blocks = [(a + 16, hsize)]
        while blocks and len(out) < nmsg + 64:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end:
                t, sz, fl = self._u(p, 2), self._u(p + 2, 2), int(self.buf[p + 4])
                body = self._bytes(p + 8, sz)
                p += 8 + sz
                if t == MSG_CONTINUATION:
                    off, ln = int.from_bytes(body[:self.so], "little"), int.from_bytes(body[self.so:self.so + self.sl], "little")
                    blocks.append((self._addr(off), ln))
                elif t != 0:
                    out.append((t, fl, body))
        return out

    # ---- groups
    def _group_links(self, header: int) -> "OrderedDict[str, int]":
        links: "OrderedDict[str, int]" = OrderedDict()
        for t, _, body in self._messages(header):
            if t == MSG_SYMTAB:
                btree = int.from_bytes(body[:self.so], "little")
                heap = int.from_bytes(body[self.so:2 * self.so], "little")
                h = self._addr(heap)
                if self._bytes(h, 4) != b"HEAP":
                    raise H5Error("local heap signature missing")
                heap_data = self._addr(self._u(h + 8 + 2 * self.sl, self.so))
                self._walk_group_btree(btree, heap_data, links)
            elif t == MSG_LINK:
                self._parse_link(body, links)
        return links

    def _walk_group_btree(self, node: int, heap_data: int, links) -> None:
        a = self._addr(node)
        if self._bytes(a, 4) == b"SNOD":
            n = self._u(a + 6, 2)
            p = a + 8
            for _ in range(n):
                name_off, hdr = self._u(p, self.so), self._u(p + self.so, self.so)
                q = heap_data + name_off
                e = q
                while self.buf[e] != 0:
                    e += 1
                links[self._bytes(q, e - q).decode()] = hdr
                p += 2 * self.so + 4 + 4 + 16
            return
        if self._bytes(a, 4) != b"TREE":
            raise H5Error(f"group B-tree node at {a:#x}: bad signature")
        n = self._u(a + 6, 2)
        p = a + 8 + 2 * self.so + self.sl                      # skip siblings and key 0
        for _ in range(n):
            self._walk_group_btree(self._u(p, self.so), heap_data, links)
            p += self.so + self.sl

    def _parse_link(self, body: bytes, links) -> None:
        flags = body[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[p]; p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nl = 1 << (flags & 3)
        n = int.from_bytes(body[p:p + nl], "little"); p += nl
        name = body[p:p + n].decode(); p += n
        if ltype == 0:
            links[name] = int.from_bytes(body[p:p + self.so], "little")

    # ---- datasets
    def keys(self):
        return list(self.links)

    def __contains__(self, name: str) -> bool:
        return name.strip("/") in self.links

    def __getitem__(self, name: str) -> np.ndarray:
        name = name.strip("/")
        if name not in self.links:
            raise KeyError(f"{name!r} not in {self.path} (has {self.keys()})")
        return self._read_dataset(self.links[name])

    def attrs(self, name: str) -> Dict[str, object]:
        out: Dict[str, object] = {}
        for t, _, body in self._messages(self.links[name.strip("/")]):
            if t == MSG_ATTRIBUTE and body[0] == 1:
                nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
                pad = lambda v: (v + 7) // 8 * 8
                p = 8
                aname = body[p:p + nsz].split(b"\0")[0].decode(); p += pad(nsz)
                dt = self._datatype(body[p:p + tsz]); p += pad(tsz)
                shape = self._dataspace(body[p:p + ssz]); p += pad(ssz)
                n = int(np.prod(shape)) if shape else 1
                raw = body[p:p + n * dt.itemsize]
                out[aname] = raw.split(b"\0")[0].decode() if dt.kind == "S" else np.frombuffer(raw, dt).reshape(shape)
        return out

    def _dataspace(self, body: bytes) -> Tuple[int, ...]:
        ver, rank = body[0], body[1]
        p = 8 if ver == 1 else 4
        return tuple(int.from_bytes(body[p + i * self.sl:p + (i + 1) * self.sl], "little") for i in range(rank))

    @staticmethod
    def _datatype(body: bytes) -> np.dtype:
        cls, bits0 = body[0] & 0x0F, body[1]
        size = struct.unpack_from("<I", body, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise H5Error(f"unsupported datatype class {cls}")

    def _read_dataset(self, header: int) -> np.ndarray:
        shape = dtype = layout = None
        filters: List[Tuple[int, List[int]]] = []
        for t, _, body in self._messages(header):
            if t == MSG_DATASPACE:
                shape = self._dataspace(body)
            elif t == MSG_DATATYPE:
                dtype = self._datatype(body)
            elif t == MSG_LAYOUT:
                layout = body
            elif t == MSG_FILTERS:
                filters = self._filters(body)
        if shape is None or dtype is None or layout is None:
            raise H5Error("dataset lacks a dataspace, datatype or layout message")
        n = int(np.prod(shape)) if shape else 1
        ver = layout[0]
        if ver in (1, 2):
            rank, cls = layout[1], layout[2]
            p = 8
            addr = None
            if cls != 0:
                addr = int.from_bytes(layout[p:p + self.so], "little"); p += self.so
            dims = struct.unpack_from(f"<{rank}I", layout, p); p += 4 * rank
            if cls == 1:
                return self._contiguous(addr, n, dtype, shape)
            if cls == 2:
                return self._chunked(addr, dims[:-1], shape, dtype, filters)
            size = struct.unpack_from("<I", layout, p)[0]
            return np.frombuffer(layout[p + 4:p + 4 + size], dtype, n).reshape(shape).copy()
        if ver == 3:
            cls = layout[1]
            if cls == 0:
                size = struct.unpack_from("<H", layout, 2)[0]
                return np.frombuffer(layout[4:4 + size], dtype, n).reshape(shape).copy()
            if cls == 1:
                return self._contiguous(int.from_bytes(layout[2:2 + self.so], "little"), n, dtype, shape)
            if cls == 2:
                rank = layout[2]
                addr = int.from_bytes(layout[3:3 + self.so], "little")
                dims = struct.unpack_from(f"<{rank}I", layout, 3 + self.so)
                return self._chunked(addr, dims[:-1], shape, dtype, filters)
        raise H5Error(f"unsupported data layout version {ver} / class {layout[1]}")

    def _contiguous(self, addr: int, n: int, dtype: np.dtype, shape) -> np.ndarray:
        if addr == UNDEF & ((1 << (8 * self.so)) - 1):
            return np.zeros(shape, dtype.newbyteorder("="))
        a = self._addr(addr)
        return np.frombuffer(self.buf[a:a + n * dtype.itemsize].tobytes(), dtype, n).reshape(shape).astype(dtype.newbyteorder("="))

    @staticmethod
    def _filters(body: bytes) -> List[Tuple[int, List[int]]]:
        ver, nf = body[0], body[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid = struct.unpack_from("<H", body, p)[0]; p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", body, p)[0]; p += 2
            _flags, ncd = struct.unpack_from("<HH", body, p); p += 4
            p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = list(struct.unpack_from(f"<{ncd}I", body, p)); p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _chunked(self, btree: int, chunk: Tuple[int, ...], shape, dtype: np.dtype, filters) -> np.ndarray:
        out = np.zeros(shape, dtype.newbyteorder("="))
        if btree == UNDEF & ((1 << (8 * self.so)) - 1):
            return out
        rank = len(shape)
        csize = int(np.prod(chunk)) * dtype.itemsize

        def leaf(addr: int, nbytes: int, mask: int, offs: Tuple[int, ...]) -> None:
            raw = self._bytes(self._addr(addr), nbytes)
            for k in range(len(filters) - 1, -1, -1):                      # undo the pipeline back to front
                if mask & (1 << k):
                    continue
                fid, cd = filters[k]
                if fid == 3:
                    raw = raw[:-4]
                elif fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dtype.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                else:
                    raise H5Error(f"unsupported HDF5 filter id {fid}")
            if len(raw) != csize:
                raise H5Error(f"chunk at {offs}: {len(raw)} bytes after filters, expected {csize}")
            block = np.frombuffer(raw, dtype).reshape(chunk)
            sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
            sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = block[sl_in]

        def walk(node: int) -> None:
            a = self._addr(node)
            if self._bytes(a, 4) != b"TREE" or self.buf[a + 4] != 1:
                raise H5Error(f"chunk B-tree node at {a:#x}: bad signature / type")
            level, n = int(self.buf[a + 5]), self._u(a + 6, 2)
            p = a + 8 + 2 * self.so
            key = 8 + 8 * (rank + 1)
            for _ in range(n):
                nbytes, mask = struct.unpack_from("<II", self.buf, p)
                offs = struct.unpack_from(f"<{rank}Q", self.buf, p + 8)
                child = self._u(p + key, self.so)
                if level == 0:
                    leaf(child, nbytes, mask, offs)
                else:
                    walk(child)
                p += key + self.so

        walk(btree)
        return out


def read_dataset(path: str, name: str) -> np.ndarray:
    """``h5py.File(path, 'r')[name][()]``."""
    return H5File(path)[name]


# =========================================================================================== writer
def fletcher32(data: bytes) -> int:
    """HDF5's Fletcher-32 (H5_checksum_fletcher32): big-endian 16-bit words, sums folded every 360 words."""
    a = np.frombuffer(data[:len(data) // 2 * 2], ">u2").astype(np.uint64)
    s1 = s2 = 0
    for i in range(0, len(a), 360):
        blk = a[i:i + 360]
        c = np.cumsum(blk, dtype=np.uint64)
        s2 += int(c.sum()) + s1 * len(blk)
        s1 += int(c[-1]) if len(blk) else 0
        s1 = (s1 & 0xFFFF) + (s1 >> 16)
        s2 = (s2 & 0xFFFF) + (s2 >> 16)
    if len(data) % 2:
        s1 += data[-1] << 8
        s2 += s1
        s1 = (s1 & 0xFFFF) + (s1 >> 16)
        s2 = (s2 & 0xFFFF) + (s2 >> 16)
    s1 = (s1 & 0xFFFF) + (s1 >> 16)
    s2 = (s2 & 0xFFFF) + (s2 >> 16)
    return (s2 << 16) | s1


_MATLAB_CLASS = {"float32": "single", "float64": "double", "uint8": "uint8", "int8": "int8", "uint16": "uint16", "int16": "int16",
                 "uint32": "uint32", "int32": "int32", "uint64": "uint64", "int64": "int64"}


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _datatype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        bits, eb, mb, bias = (32, 8, 23, 127) if dt.itemsize == 4 else (64, 11, 52, 1023)
        return struct.pack("<BBBBI", 0x11, 0x20, bits - 1, 0, dt.itemsize) + struct.pack("<HHBBBBI", 0, bits, mb, eb, 0, mb, bias)
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise H5Error(f"unsupported dtype {dt}")


def _dataspace_msg(shape) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _string_attr(name: str, value: str) -> bytes:
    nm = name.encode() + b"\0"
    val = value.encode()
    dt = struct.pack("<BBBBI", 0x13, 0x00, 0, 0, len(val))                 # class 3 (string), null-terminated, ASCII
    ds = struct.pack("<BBB5x", 1, 0, 0)                                      # scalar dataspace
    return struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + val


def write_mat73(path: str, arrays: Dict[str, np.ndarray], chunks: Optional[Tuple[int, ...]] = None, compress: bool = False,
                shuffle: bool = True, fletcher: bool = True, btree_k: int = 32) -> None:
    """Writes ``arrays`` as the datasets of a MATLAB v7.3 file.  Arrays are stored with the dims given (``hdf5storage`` writes
    the TRANSPOSE of the Python array, i.e. MATLAB's column-major view: callers that mirror it pass ``a.T``).  ``chunks`` (applied
    to every array of that rank) selects the chunked layout, with the hdf5storage filter stack when ``compress``."""
    out = bytearray()
    head = ("MATLAB 7.3 MAT-file, Platform: fisr_b200, Created on: %s HDF5 schema 1.00 ." % time.strftime("%a %b %d %H:%M:%S %Y")).encode()
    out += head.ljust(116, b" ") + b"\0" * 8 + b"\x00\x02IM"
    out = out.ljust(512, b"\0")
    BASE = 512
    names = list(arrays)
    # ---- fixed layout (relative addresses): superblock 0..96, heap 96, B-tree, root header, SNOD, then datasets
    heap_data = b"\0" * 8 + b"".join(n.encode() + b"\0" + b"\0" * (-(len(n) + 1) % 8) for n in names)
    name_off, o = {}, 8
    for n in names:
        name_off[n] = o
        o += len(n) + 1 + (-(len(n) + 1) % 8)
    heap_data = heap_data.ljust(max(128, (len(heap_data) + 7) // 8 * 8), b"\0")
    a_heap = 96
    a_heap_data = a_heap + 32
    a_btree = a_heap_data + len(heap_data)
    btree_size = 8 + 16 + (2 * 16 + 1) * 8 + 2 * 16 * 8                      # node sized for group internal K = 16
    a_root = a_btree + btree_size
    root_size = 16 + 8 + 16
    a_snod = a_root + root_size
    leaf_k = max(4, (len(names) + 1) // 2)
    snod_size = 8 + 2 * leaf_k * 40
    cursor = a_snod + snod_size
    blobs: List[Tuple[int, bytes]] = []

    def alloc(data: bytes, align: int = 8) -> int:
        nonlocal cursor
        cursor = (cursor + align - 1) // align * align
        addr = cursor
        blobs.append((addr, data))
        cursor += len(data)
        return addr

    headers = {}
    for n in names:
        a = np.ascontiguousarray(arrays[n])
        dt = a.dtype.newbyteorder("<")
        a = a.astype(dt, copy=False)
        # fill value v2: allocation time late (contiguous) / incremental (chunked), write time "if set", no fill value defined
        msgs = [_msg(MSG_FILL, struct.pack("<BBBB", 2, 2 if chunks is None else 3, 2, 0)), _msg(MSG_DATATYPE, _datatype_msg(dt), 1),
                _msg(MSG_DATASPACE, _dataspace_msg(a.shape))]
        if chunks is None or a.ndim != len(chunks):
            addr = alloc(a.tobytes())
            msgs.append(_msg(MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, addr, a.nbytes)))
        else:
            filt = []
            if compress:
                if shuffle:
                    filt.append((2, [dt.itemsize]))
                filt.append((1, [7]))
                if fletcher:
                    filt.append((3, []))
            entries = []
            grid = [range(0, s, c) for s, c in zip(a.shape, chunks)]
            for offs in np.ndindex(*[len(g) for g in grid]):
                o0 = tuple(g[i] for g, i in zip(grid, offs))
                block = np.zeros(chunks, dt)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(o0, chunks, a.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
                raw = block.tobytes()
                for fid, cd in filt:
                    if fid == 2:
                        raw = np.frombuffer(raw, np.uint8).reshape(-1, cd[0]).T.tobytes()
                    elif fid == 1:
                        raw = zlib.compress(raw, cd[0])
                    elif fid == 3:
                        raw = raw + struct.pack("<I", fletcher32(raw))
                entries.append((o0, len(raw), alloc(raw)))
            rank = a.ndim

            def key(offs, nbytes):
                return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", int(v)) for v in offs) + struct.pack("<Q", 0)

            node_bytes = 24 + (2 * btree_k + 1) * (8 + 8 * (rank + 1)) + 2 * btree_k * 8       # the library's fixed node size

            def build(nodes, level):
                """nodes: [(first offsets, nbytes, address)] -> address of the (sub)tree root over them."""
                if len(nodes) <= 2 * btree_k:
                    body = b"TREE" + struct.pack("<BBH", 1, level, len(nodes)) + struct.pack("<QQ", UNDEF, UNDEF)
                    for offs, nbytes, addr in nodes:
                        body += key(offs, nbytes) + struct.pack("<Q", addr)
                    body += key(tuple(s for s in a.shape), 0)                    # final key: one past the last chunk
                    return alloc(body.ljust(node_bytes, b"\0"))
                groups = [nodes[i:i + 2 * btree_k] for i in range(0, len(nodes), 2 * btree_k)]
                return build([(g[0][0], g[0][1], build(g, level)) for g in groups], level + 1)

            root = build(entries, 0)
            msgs.append(_msg(MSG_LAYOUT, struct.pack("<BBBQ", 3, 2, rank + 1, root) + struct.pack(f"<{rank + 1}I", *chunks, dt.itemsize)))
            if filt:
                fb = struct.pack("<BB6x", 1, len(filt))
                for fid, cd in filt:
                    fb += struct.pack("<HHHH", fid, 0, 0 if fid != 3 else 0, len(cd)) + b"".join(struct.pack("<I", v) for v in cd)
                    if len(cd) % 2:
                        fb += b"\0" * 4
                msgs.append(_msg(MSG_FILTERS, fb))
        msgs.append(_msg(MSG_ATTRIBUTE, _string_attr("MATLAB_class", _MATLAB_CLASS.get(dt.name, dt.name))))
        body = b"".join(msgs)
        hdr = struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body
        headers[n] = alloc(hdr)
    eof = (cursor + 7) // 8 * 8
    # ---- superblock v0 (layout of a MATLAB-written file: scipy's testhdf5_7.4_GLNX86.mat)
    sb = SIGNATURE + struct.pack("<BBBBBBBxHHI", 0, 0, 0, 0, 0, 8, 8, leaf_k, 16, 0)
    sb += struct.pack("<QQQQ", BASE, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQI4x", 0, a_root, 1) + struct.pack("<QQ", a_btree, a_heap)
    sb = sb.ljust(96, b"\0")
    heap = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, a_heap_data)          # free-list head 1 = H5HL_FREE_NULL
    btree = b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF) + struct.pack("<QQQ", 0, a_snod, name_off[sorted(names)[-1]])
    btree = btree.ljust(btree_size, b"\0")
    root = struct.pack("<BxHII4x", 1, 1, 1, 24) + _msg(MSG_SYMTAB, struct.pack("<QQ", a_btree, a_heap))
    snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
    for n in sorted(names):                                                   # entries sorted by name, like the library keeps them
        snod += struct.pack("<QQI4x16x", name_off[n], headers[n], 0)
    snod = snod.ljust(snod_size, b"\0")
    body = bytearray(eof)
    body[0:len(sb)] = sb
    body[a_heap:a_heap + len(heap)] = heap
    body[a_heap_data:a_heap_data + len(heap_data)] = heap_data
    body[a_btree:a_btree + len(btree)] = btree
    body[a_root:a_root + len(root)] = root
    body[a_snod:a_snod + len(snod)] = snod
    for addr, data in blobs:
        body[addr:addr + len(data)] = data
    with open(path, "wb") as f:
        f.write(bytes(out))
        f.write(bytes(body))
