"""A small, dependency-free subset of HDF5, used only when ``h5py`` cannot be imported.

Why it exists: the reference's input (ipyrad ``*.snps.hdf5``: datasets ``snps`` u8[n,S] with
attribute ``names``, ``snpsmap``) and its working database (``seqarr``, ``spans``, ``tmparr``,
``tmpmap``) are HDF5 files (tetrad/src/write_database.py:132-168), and this image ships neither
h5py nor libhdf5.

What it does (HDF5 File Format Specification 3.0):

* **write**: superblock v0, symbol-table root group (v1 B-tree + local heap + SNOD), v1 object
  headers, contiguous little-endian integer / float / fixed-length-string datasets, v1 attributes.
* **read**: superblock v0-v3; object headers v1 and v2 (with continuation blocks); groups stored
  as symbol tables or as compact link messages; layouts compact / contiguous / chunked with a v1
  B-tree index (layout v3) or a single chunk / implicit / fixed-array index (layout v4); the
  ``deflate`` and ``shuffle`` filters; fixed-point, floating-point and fixed-length string types;
  attributes v1-v3 with fixed-size types or variable-length strings (global heap; what h5py writes for a
  ``list[str]`` such as ``names``); a user block in front of the superblock.
* not supported (raises ``NotImplementedError``): variable-length sequences, dense (fractal-heap)
  attribute or link storage, extensible-array / v2-B-tree chunk indexes, external links.

The API mirrors the few h5py calls the code base makes: ``File(path, mode)`` as a context
manager, ``f[name][:]``, ``name in f``, ``f.keys()``, ``f[name].shape/.dtype/.attrs``,
``f.create_dataset(name, data=...)`` and ``dset.attrs[key] = array``.

Validation status: the reader is checked against the one libhdf5-written file this image holds (scipy's MATLAB
v7.3 test file, copied to ``tests/golden/libhdf5_matlab73_testdouble.mat``: user block, superblock v0, symbol-table
group, v1 object header, contiguous f64 data, fixed-length string attribute -- the structures an h5py file written
with default settings has; that file found the reader's one real defect, addresses not rebased past a user block).
Chunked / filtered datasets, v2 headers and variable-length strings are read from files assembled by the tests from
the format specification, not by libhdf5; written files are re-read by this module.  Prefer h5py whenever it is
installed (``tetrad_b200/database.py`` does).
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class _VlenStr:
    """Marker for the variable-length string datatype (class 9, type 1); elements are 16-byte global-heap references."""
    itemsize = 16

    def __init__(self, utf8: bool):
        self.utf8 = utf8

    def __repr__(self):
        return f"vlen-str({'utf-8' if self.utf8 else 'ascii'})"


_VLEN_STR_UTF8, _VLEN_STR_ASCII = _VlenStr(True), _VlenStr(False)


# =============================================================================================
# reading
# =============================================================================================
class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        pos = 0
        while buf[pos:pos + 8] != SIGNATURE:          # the superblock may sit at 0, 512, 1024, ...
            pos = 512 if pos == 0 else pos * 2
            if pos + 8 > len(buf):
                raise ValueError("not an HDF5 file")
        # A file with a user block (MATLAB v7.3 files, ``h5py.File(userblock_size=...)``) has its superblock at 512, 1024,
        # ...; every address in it is relative to the base address, which libhdf5 takes to be the superblock's own
        # position (H5F__super_read adjusts ``base_addr`` to the address it found the signature at).
        self.base = pos
        if pos:
            buf = self.buf = buf[pos:]
            pos = 0
        version = buf[pos + 8]
        if version in (0, 1):
            self.so, self.sl = buf[pos + 13], buf[pos + 14]
            p = pos + 24 + (4 if version == 1 else 0)
            p += 4 * self.so                           # base, free-space, eof, driver addresses
            # root group symbol table entry
            self.root = self._addr(p + self.so)
        elif version in (2, 3):
            self.so, self.sl = buf[pos + 9], buf[pos + 10]
            self.root = self._addr(pos + 12 + 3 * self.so)
        else:
            raise NotImplementedError(f"superblock version {version}")
        if self.so != 8 or self.sl != 8:
            raise NotImplementedError("only 8-byte offsets/lengths are supported")

    def _addr(self, p: int) -> int:
        return struct.unpack_from("<Q", self.buf, p)[0]

    # ---- object headers -------------------------------------------------------------------
    def messages(self, addr: int):
        """[(type, flags, bytes)] of the object header at ``addr``."""
        buf = self.buf
        out = []
        if buf[addr:addr + 4] == b"OHDR":
            flags = buf[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            nbytes = 1 << (flags & 3)
            size = int.from_bytes(buf[p:p + nbytes], "little")
            p += nbytes
            blocks = [(p, size)]
            while blocks:
                p, size = blocks.pop(0)
                end = p + size
                while p + 4 <= end:
                    mtype = buf[p]
                    msize = struct.unpack_from("<H", buf, p + 1)[0]
                    mflags = buf[p + 3]
                    p += 4 + (2 if flags & 0x04 else 0)
                    data = buf[p:p + msize]
                    p += msize
                    if mtype == 0x10:
                        off, length = struct.unpack_from("<QQ", data, 0)
                        blocks.append((off + 4, length - 8))      # skip "OCHK", drop checksum
                    elif mtype != 0:
                        out.append((mtype, mflags, data))
            return out
        version = buf[addr]
        if version != 1:
            raise NotImplementedError(f"object header version {version}")
        nmsg = struct.unpack_from("<H", buf, addr + 2)[0]
        size = struct.unpack_from("<I", buf, addr + 8)[0]
        blocks = [(addr + 16, size)]
        while blocks and len(out) < nmsg + 64:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end:
                mtype, msize, mflags = struct.unpack_from("<HHB", buf, p)
                data = buf[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    off, length = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off, length))
                elif mtype != 0:
                    out.append((mtype, mflags, data))
        return out

    # ---- groups ----------------------------------------------------------------------------
    def links(self, addr: int) -> dict:
        """{name: object header address} of the group at ``addr``."""
        out = {}
        for mtype, _, data in self.messages(addr):
            if mtype == 0x11:                                   # symbol table
                btree, heap = struct.unpack_from("<QQ", data, 0)
                heap_data = self._addr(heap + 8 + 2 * self.sl)
                self._walk_group_btree(btree, heap_data, out)
            elif mtype == 0x06:                                 # link message
                flags = data[1]
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = data[p]; p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                nb = 1 << (flags & 3)
                nlen = int.from_bytes(data[p:p + nb], "little"); p += nb
                name = bytes(data[p:p + nlen]).decode("utf-8"); p += nlen
                if ltype == 0:
                    out[name] = struct.unpack_from("<Q", data, p)[0]
            elif mtype == 0x02:                                 # link info: dense storage?
                flags = data[1]
                p = 2 + (8 if flags & 1 else 0)
                fheap = struct.unpack_from("<Q", data, p)[0]
                if fheap != UNDEF:
                    raise NotImplementedError("dense (fractal heap) link storage")
        return out

    def _walk_group_btree(self, addr: int, heap_data: int, out: dict):
        buf = self.buf
        if buf[addr:addr + 4] != b"TREE":
            raise ValueError("bad group B-tree node")
        level, used = buf[addr + 5], struct.unpack_from("<H", buf, addr + 6)[0]
        p = addr + 8 + 2 * self.so
        for i in range(used):
            child = self._addr(p + self.sl + i * (self.sl + self.so))
            if level > 0:
                self._walk_group_btree(child, heap_data, out)
                continue
            if buf[child:child + 4] != b"SNOD":
                raise ValueError("bad symbol table node")
            nsym = struct.unpack_from("<H", buf, child + 6)[0]
            for s in range(nsym):
                e = child + 8 + s * (2 * self.so + 24)
                name_off, obj = struct.unpack_from("<QQ", buf, e)
                q = heap_data + name_off
                name = bytes(buf[q:buf.index(b"\0", q)]).decode("utf-8")
                out[name] = obj

    # ---- datatypes -------------------------------------------------------------------------
    @staticmethod
    def dtype(data: bytes) -> np.dtype:
        cls = data[0] & 0x0F
        bits0 = data[1]
        size = struct.unpack_from("<I", data, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        if cls == 9:
            if (bits0 & 0x0F) == 1:                              # variable-length STRING: what h5py writes for list[str]
                return _VLEN_STR_UTF8 if (data[2] & 0x0F) == 1 else _VLEN_STR_ASCII
            raise NotImplementedError("variable-length sequence datatypes")
        raise NotImplementedError(f"datatype class {cls}")

    def _vlen_strings(self, raw: bytes, count: int, utf8: bool) -> np.ndarray:
        """``count`` variable-length strings: each element is {u32 length, global-heap collection address, u32 object
        index}; the bytes live in a global heap collection (``GCOL``, spec III.E).  Returned like h5py does: an object
        array of ``str`` (UTF-8) or ``bytes`` (ASCII)."""
        out = np.empty(count, dtype=object)
        for i in range(count):
            length, addr, index = struct.unpack_from("<IQI", raw, 16 * i)
            if length == 0 or addr in (0, UNDEF):
                b = b""
            else:
                b = bytes(self._global_heap_object(addr, index)[:length])
            out[i] = b.decode("utf-8") if utf8 else b
        return out

    def _global_heap_object(self, addr: int, index: int) -> bytes:
        buf = self.buf
        if buf[addr:addr + 4] != b"GCOL":
            raise ValueError("bad global heap collection")
        end = addr + struct.unpack_from("<Q", buf, addr + 8)[0]
        p = addr + 16
        while p + 16 <= end:
            idx, _ref, _res, size = struct.unpack_from("<HHIQ", buf, p)
            if idx == 0:                                         # free space: nothing follows
                break
            if idx == index:
                return buf[p + 16:p + 16 + size]
            p += 16 + (size + 7) // 8 * 8
        raise KeyError(f"global heap object {index} not found")

    def dataspace(self, data: bytes):
        version, rank, flags = data[0], data[1], data[2]
        p = 8 if version == 1 else 4
        if version == 2 and data[3] == 2:
            return None                                          # null dataspace
        return tuple(struct.unpack_from("<Q", data, p + 8 * i)[0] for i in range(rank))

    # ---- datasets --------------------------------------------------------------------------
    def dataset(self, addr: int):
        shape = dtype = layout = None
        filters = []
        attrs = {}
        for mtype, _, data in self.messages(addr):
            if mtype == 0x01:
                shape = self.dataspace(data)
            elif mtype == 0x03:
                dtype = self.dtype(data)
            elif mtype == 0x08:
                layout = data
            elif mtype == 0x0B:
                filters = self._filters(data)
            elif mtype == 0x0C:
                name, value = self._attribute(data)
                attrs[name] = value
            elif mtype == 0x15:
                fheap = struct.unpack_from("<Q", data, 2 + (2 if data[1] & 1 else 0))[0]
                if fheap != UNDEF:
                    raise NotImplementedError("dense (fractal heap) attribute storage")
        if shape is None or dtype is None or layout is None:
            raise ValueError("object is not a dataset")
        return shape, dtype, layout, filters, attrs

    def _filters(self, data: bytes):
        version, nfilters = data[0], data[1]
        p = 8 if version == 1 else 2
        out = []
        for _ in range(nfilters):
            fid = struct.unpack_from("<H", data, p)[0]; p += 2
            nlen = 0
            if version == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", data, p)[0]; p += 2
            _flags, ncd = struct.unpack_from("<HH", data, p); p += 4
            if nlen:
                p += (nlen + 7) // 8 * 8 if version == 1 else nlen
            cd = struct.unpack_from(f"<{ncd}I", data, p); p += 4 * ncd
            if version == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _attribute(self, data: bytes):
        version = data[0]
        nsize, tsize, ssize = struct.unpack_from("<HHH", data, 2)
        p = 8 + (1 if version == 3 else 0)
        pad = (lambda n: (n + 7) // 8 * 8) if version == 1 else (lambda n: n)
        name = bytes(data[p:p + nsize]).split(b"\0")[0].decode("utf-8"); p += pad(nsize)
        dtype = self.dtype(data[p:p + tsize]); p += pad(tsize)
        shape = self.dataspace(data[p:p + ssize]); p += pad(ssize)
        count = int(np.prod(shape)) if shape else 1
        if dtype is _VLEN_STR_UTF8 or dtype is _VLEN_STR_ASCII:
            value = self._vlen_strings(data[p:p + 16 * count], count, dtype is _VLEN_STR_UTF8)
        else:
            value = np.frombuffer(data, dtype=dtype, count=count, offset=p)
        return name, (value.reshape(shape) if shape else value[0])

    def read(self, shape, dtype, layout: bytes, filters) -> np.ndarray:
        version = layout[0]
        n = int(np.prod(shape)) if shape else 1
        if version in (1, 2):
            rank, cls = layout[1], layout[2]
            p = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", layout, p)[0]; p += 8
            dims = struct.unpack_from(f"<{rank}I", layout, p); p += 4 * rank
            if cls == 1:
                return self._contiguous(addr, shape, dtype, n)
            if cls == 0:
                size = struct.unpack_from("<I", layout, p)[0]
                return np.frombuffer(layout, dtype=dtype, count=n, offset=p + 4).reshape(shape).copy()
            return self._chunked_btree(addr, shape, dtype, dims[:-1], filters)
        cls = layout[1]
        if cls == 0:
            size = struct.unpack_from("<H", layout, 2)[0]
            return np.frombuffer(layout, dtype=dtype, count=n, offset=4).reshape(shape).copy()
        if cls == 1:
            addr = struct.unpack_from("<Q", layout, 2)[0]
            return self._contiguous(addr, shape, dtype, n)
        if cls != 2:
            raise NotImplementedError(f"layout class {cls}")
        if version == 3:
            rank = layout[2]
            addr = struct.unpack_from("<Q", layout, 3)[0]
            dims = struct.unpack_from(f"<{rank}I", layout, 11)
            return self._chunked_btree(addr, shape, dtype, dims[:-1], filters)
        # version 4
        flags, rank, enc = layout[2], layout[3], layout[4]
        dims = [int.from_bytes(layout[5 + i * enc:5 + (i + 1) * enc], "little") for i in range(rank)]
        p = 5 + rank * enc
        index_type = layout[p]; p += 1
        chunk = tuple(dims[:-1])
        if index_type == 1:                                      # single chunk
            if flags & 2:
                size, _mask = struct.unpack_from("<QI", layout, p); p += 12
            else:
                size = int(np.prod(chunk)) * dtype.itemsize
            addr = struct.unpack_from("<Q", layout, p)[0]
            raw = self._unfilter(self.buf[addr:addr + size], filters, dtype) if filters else self.buf[addr:addr + size]
            arr = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk))).reshape(chunk)
            return arr[tuple(slice(0, s) for s in shape)].copy()
        if index_type == 2:                                      # implicit: chunks back to back
            addr = struct.unpack_from("<Q", layout, p)[0]
            return self._assemble(shape, dtype, chunk, self._implicit_chunks(addr, shape, dtype, chunk), [])
        if index_type == 3:                                      # fixed array
            addr = struct.unpack_from("<Q", layout, p + 1)[0]
            return self._assemble(shape, dtype, chunk, self._fixed_array_chunks(addr, shape, dtype, chunk, bool(filters)), filters)
        raise NotImplementedError(f"chunk index type {index_type}")

    def _contiguous(self, addr, shape, dtype, n):
        if addr == UNDEF:
            return np.zeros(shape, dtype=dtype)
        return np.frombuffer(self.buf, dtype=dtype, count=n, offset=addr).reshape(shape).copy()

    def _unfilter(self, raw: bytes, filters, dtype) -> bytes:
        for fid, cd in reversed(filters):
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:                                       # shuffle
                size = cd[0] if cd else dtype.itemsize
                arr = np.frombuffer(raw, dtype=np.uint8)
                nel = len(arr) // size
                raw = arr[:nel * size].reshape(size, nel).T.tobytes() + arr[nel * size:].tobytes()
            elif fid == 3:                                       # fletcher32: drop the checksum
                raw = raw[:-4]
            else:
                raise NotImplementedError(f"filter id {fid}")
        return raw

    def _chunked_btree(self, addr, shape, dtype, chunk, filters):
        chunks = []
        if addr != UNDEF:
            self._walk_chunk_btree(addr, len(shape), chunks)
        return self._assemble(shape, dtype, tuple(chunk), chunks, filters)

    def _walk_chunk_btree(self, addr, rank, out):
        buf = self.buf
        if buf[addr:addr + 4] != b"TREE":
            raise ValueError("bad chunk B-tree node")
        level, used = buf[addr + 5], struct.unpack_from("<H", buf, addr + 6)[0]
        keysize = 8 + 8 * (rank + 1)
        p = addr + 8 + 2 * self.so
        for i in range(used):
            k = p + i * (keysize + self.so)
            size, mask = struct.unpack_from("<II", buf, k)
            offs = struct.unpack_from(f"<{rank}Q", buf, k + 8)
            child = self._addr(k + keysize)
            if level > 0:
                self._walk_chunk_btree(child, rank, out)
            else:
                out.append((offs, child, size, mask))

    def _implicit_chunks(self, addr, shape, dtype, chunk):
        nbytes = int(np.prod(chunk)) * dtype.itemsize
        grid = [-(-s // c) for s, c in zip(shape, chunk)]
        out = []
        for i, idx in enumerate(np.ndindex(*grid)):
            out.append((tuple(a * c for a, c in zip(idx, chunk)), addr + i * nbytes, nbytes, 0))
        return out

    def _fixed_array_chunks(self, addr, shape, dtype, chunk, filtered):
        buf = self.buf
        if buf[addr:addr + 4] != b"FAHD":
            raise ValueError("bad fixed array header")
        entry_size, page_bits = buf[addr + 6], buf[addr + 7]
        nentries = struct.unpack_from("<Q", buf, addr + 8)[0]
        dblk = self._addr(addr + 16)
        if buf[dblk:dblk + 4] != b"FADB":
            raise ValueError("bad fixed array data block")
        if nentries > (1 << page_bits):
            raise NotImplementedError("paged fixed-array chunk index")
        p = dblk + 6 + self.so
        nbytes = int(np.prod(chunk)) * dtype.itemsize
        grid = [-(-s // c) for s, c in zip(shape, chunk)]
        out = []
        for i, idx in enumerate(np.ndindex(*grid)):
            e = p + i * entry_size
            caddr = self._addr(e)
            size, mask = nbytes, 0
            if filtered:
                width = entry_size - self.so - 4
                size = int.from_bytes(buf[e + 8:e + 8 + width], "little")
                mask = struct.unpack_from("<I", buf, e + 8 + width)[0]
            if caddr != UNDEF:
                out.append((tuple(a * c for a, c in zip(idx, chunk)), caddr, size, mask))
        return out

    def _assemble(self, shape, dtype, chunk, chunks, filters):
        out = np.zeros(shape, dtype=dtype)
        nel = int(np.prod(chunk))
        for offs, addr, size, mask in chunks:
            raw = self.buf[addr:addr + size]
            active = [f for i, f in enumerate(filters) if not (mask >> i) & 1]
            if active:
                raw = self._unfilter(raw, active, dtype)
            block = np.frombuffer(raw, dtype=dtype, count=nel).reshape(chunk)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
            out[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]
        return out


class Dataset:
    """Read view of one dataset (``shape``, ``dtype``, ``attrs``, ``dset[:]``)."""

    def __init__(self, reader: _Reader, addr: int):
        self._reader = reader
        self.shape, self.dtype, self._layout, self._filters, self.attrs = reader.dataset(addr)

    def __getitem__(self, key):
        data = self._reader.read(self.shape, self.dtype, self._layout, self._filters)
        return data[key]

    def __len__(self):
        return self.shape[0]


# =============================================================================================
# writing
# =============================================================================================
class _PendingDataset:
    def __init__(self, data: np.ndarray):
        self.data = data
        self.shape, self.dtype = data.shape, data.dtype
        self.attrs = {}

    def __getitem__(self, key):
        return self.data[key]


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _datatype_msg(dtype: np.dtype) -> bytes:
    dtype = np.dtype(dtype)
    if dtype.kind in "iu":
        bits = 0x08 if dtype.kind == "i" else 0
        return struct.pack("<BBBBIHH", 0x10 | 0, bits, 0, 0, dtype.itemsize, 0, 8 * dtype.itemsize)
    if dtype.kind == "f":
        if dtype.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            return struct.pack("<BBBBI", 0x10 | 1, 0x20, 31, 0, 4) + props
        props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, 63, 0, 8) + props
    if dtype.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, dtype.itemsize)      # null-padded ASCII (numpy S)
    raise TypeError(f"cannot store dtype {dtype}")


def _dataspace_msg(shape) -> bytes:
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), 0, 0, 0, 0) + data


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0\0\0\0" + body


class _GlobalHeap:
    """One global heap collection (spec III.E) holding the bytes of every variable-length string of the file."""

    def __init__(self):
        self.addr = 0
        self.objects = []

    def add(self, b: bytes) -> int:
        self.objects.append(b)
        return len(self.objects)                                  # object indices start at 1

    def size(self) -> int:
        used = 16 + sum(16 + len(b) + (-len(b) % 8) for b in self.objects) + 16      # header, objects, free-space object
        return max(4096, used + (-used % 8))

    def tobytes(self) -> bytes:
        total = self.size()
        out = bytearray(total)
        out[0:4] = b"GCOL"
        struct.pack_into("<BBBBQ", out, 4, 1, 0, 0, 0, total)
        p = 16
        for i, b in enumerate(self.objects):
            struct.pack_into("<HHIQ", out, p, i + 1, 1, 0, len(b))
            out[p + 16:p + 16 + len(b)] = b
            p += 16 + len(b) + (-len(b) % 8)
        struct.pack_into("<HHIQ", out, p, 0, 0, 0, total - p)    # object 0: the free space, header included
        return bytes(out)


def _is_vlen_value(value) -> bool:
    value = np.asarray(value)
    return value.dtype.kind == "O" and all(isinstance(v, (str, bytes)) for v in value.ravel())


def _attribute_msg(name: str, value, gheap: "_GlobalHeap | None" = None) -> bytes:
    value = np.asarray(value)
    if value.dtype.kind == "U":
        value = np.char.encode(value, "utf-8")
    if value.dtype.kind == "O":
        # list[str] as h5py stores it: variable-length UTF-8 strings, 16-byte global-heap references in the message
        if gheap is None or not _is_vlen_value(value):
            raise TypeError("object arrays other than lists of str/bytes cannot be stored")
        nm = name.encode("utf-8") + b"\0"
        base = struct.pack("<BBBBI", 0x10 | 3, 0x10, 0, 0, 1)     # base type: 1-byte string
        dt = struct.pack("<BBBBI", 0x10 | 9, 0x01, 0x01, 0, 16) + base   # class 9, type = string, charset UTF-8
        ds = _dataspace_msg(value.shape)
        body = b""
        for v in value.ravel():
            b = v.encode("utf-8") if isinstance(v, str) else bytes(v)
            body += struct.pack("<IQI", len(b), gheap.addr, gheap.add(b))
        head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
        return _message(0x0C, head + _pad8(nm) + _pad8(dt) + _pad8(ds) + body)
    value = np.ascontiguousarray(value).astype(value.dtype.newbyteorder("<"), copy=False)
    nm = name.encode("utf-8") + b"\0"
    dt, ds = _datatype_msg(value.dtype), _dataspace_msg(value.shape)
    head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
    return _message(0x0C, head + _pad8(nm) + _pad8(dt) + _pad8(ds) + value.tobytes())


def _write_file(path: Path, datasets: dict) -> None:
    names = sorted(datasets, key=lambda s: s.encode("utf-8"))
    # local heap data: "" at offset 0, then the names
    heap = bytearray(b"\0" * 8)
    name_off = {}
    for name in names:
        name_off[name] = len(heap)
        heap += _pad8(name.encode("utf-8") + b"\0")
    heap_data = bytes(heap)

    pos = 96                                                    # after the v0 superblock
    root_header_addr = pos
    root_header_size = 16 + 8 + 16                              # prefix + one symbol-table message
    pos += root_header_size
    heap_addr = pos
    pos += 32
    heap_data_addr = pos
    pos += len(heap_data)
    groups = [names[i:i + 8] for i in range(0, len(names), 8)] or [[]]
    if len(groups) > 32:
        raise NotImplementedError("more than 256 datasets in one file")
    btree_addr = pos
    btree_size = 8 + 16 + (2 * 32 + 1) * 8 + 2 * 32 * 8        # sized for 2K = 32 children (K = 16)
    pos += btree_size
    snod_addrs = []
    snod_size = 8 + 8 * 40
    for _ in groups:
        snod_addrs.append(pos)
        pos += snod_size

    gheap = _GlobalHeap()
    nvlen = sum(1 for name in names for v in datasets[name].attrs.values() if _is_vlen_value(v))
    if nvlen:
        for name in names:                                       # size the collection before anything refers to it
            for v in datasets[name].attrs.values():
                if _is_vlen_value(v):
                    for x in np.asarray(v).ravel():
                        gheap.add(x.encode("utf-8") if isinstance(x, str) else bytes(x))
        gheap.addr = pos
        pos += gheap.size()
        gheap.objects = []                                       # refilled, in the same order, by _attribute_msg

    blobs, headers = {}, {}
    for name in names:
        ds = datasets[name]
        data = np.ascontiguousarray(ds.data)
        if data.dtype.kind == "U":
            data = np.char.encode(data, "utf-8")
        data = data.astype(data.dtype.newbyteorder("<"), copy=False)
        raw = data.tobytes()
        data_addr = pos
        pos += len(raw) + (-len(raw) % 8)
        msgs = [_message(0x01, _dataspace_msg(data.shape)), _message(0x03, _datatype_msg(data.dtype)),
                _message(0x08, struct.pack("<BBQQ", 3, 1, data_addr if raw else UNDEF, len(raw)))]
        msgs += [_attribute_msg(k, v, gheap) for k, v in ds.attrs.items()]
        hdr = _object_header(msgs)
        headers[name] = (pos, hdr)
        pos += len(hdr) + (-len(hdr) % 8)
        blobs[name] = (data_addr, raw)
    eof = pos

    out = bytearray(eof)
    out[0:8] = SIGNATURE
    struct.pack_into("<BBBBBBBBHHI", out, 8, 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    struct.pack_into("<QQQQ", out, 24, 0, UNDEF, eof, UNDEF)
    struct.pack_into("<QQII", out, 56, 0, root_header_addr, 1, 0)
    struct.pack_into("<QQ", out, 80, btree_addr, heap_addr)
    root = _object_header([_message(0x11, struct.pack("<QQ", btree_addr, heap_addr))])
    out[root_header_addr:root_header_addr + len(root)] = root
    out[heap_addr:heap_addr + 4] = b"HEAP"
    struct.pack_into("<QQQ", out, heap_addr + 8, len(heap_data), 1, heap_data_addr)     # 1 = no free block
    out[heap_data_addr:heap_data_addr + len(heap_data)] = heap_data
    out[btree_addr:btree_addr + 4] = b"TREE"
    nchild = len(groups) if names else 0
    struct.pack_into("<BBHQQ", out, btree_addr + 4, 0, 0, nchild, UNDEF, UNDEF)
    p = btree_addr + 24
    struct.pack_into("<Q", out, p, 0); p += 8                   # key 0: the empty string
    for grp, saddr in zip(groups, snod_addrs):
        if not grp:
            break
        struct.pack_into("<QQ", out, p, saddr, name_off[grp[-1]]); p += 16
    for grp, saddr in zip(groups, snod_addrs):
        out[saddr:saddr + 4] = b"SNOD"
        struct.pack_into("<BBH", out, saddr + 4, 1, 0, len(grp))
        for i, name in enumerate(grp):
            struct.pack_into("<QQII", out, saddr + 8 + i * 40, name_off[name], headers[name][0], 0, 0)
    if nvlen:
        blob = gheap.tobytes()
        out[gheap.addr:gheap.addr + len(blob)] = blob
    for name in names:
        addr, hdr = headers[name]
        out[addr:addr + len(hdr)] = hdr
        daddr, raw = blobs[name]
        out[daddr:daddr + len(raw)] = raw
    Path(path).write_bytes(bytes(out))


# =============================================================================================
# h5py-like front end
# =============================================================================================
class File:
    """``with File(path, "r") as f: f["snps"][:]`` / ``with File(path, "w") as f: f.create_dataset(...)``.
    Extra keyword arguments of ``h5py.File`` (``swmr``, ``libver``) are accepted and ignored."""

    def __init__(self, path, mode: str = "r", **_ignored):
        self.path, self.mode = Path(path), mode
        self._pending = {}
        self._reader = None
        if mode in ("r", "r+", "a") and self.path.exists():
            self._reader = _Reader(self.path.read_bytes())
            self._links = self._reader.links(self._reader.root)
            if mode != "r":                                      # load everything so close() can rewrite
                for name in list(self._links):
                    ds = Dataset(self._reader, self._links[name])
                    pend = _PendingDataset(ds[:])
                    pend.attrs.update(ds.attrs)
                    self._pending[name] = pend
        elif mode in ("w", "a", "r+"):
            self._links = {}
        else:
            raise FileNotFoundError(path)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self.mode != "r" and self._pending is not None:
            _write_file(self.path, self._pending)
        self._pending = None

    def keys(self):
        return list(self._pending) if self.mode != "r" else list(self._links)

    def __contains__(self, name):
        return name in self.keys()

    def __getitem__(self, name):
        if self.mode != "r":
            return self._pending[name]
        if name not in self._links:
            raise KeyError(name)
        return Dataset(self._reader, self._links[name])

    def __delitem__(self, name):
        if self.mode == "r":
            raise IOError("file is read-only")
        del self._pending[name]

    def create_dataset(self, name, data=None, shape=None, dtype=None, **_ignored):
        if self.mode == "r":
            raise IOError("file is read-only")
        if data is None:
            data = np.zeros(shape, dtype=dtype or np.float64)
        data = np.asarray(data)
        if dtype is not None:
            data = data.astype(dtype)
        self._pending[name] = _PendingDataset(data)
        return self._pending[name]
