"""Minimal HDF5 reader / writer for the reference's dataset files (host side).

The reference reads ``data_prepro.h5`` / ``data_img.h5`` through the torch-hdf5 rock
(``002_train_vqa_arch1/002_train_baseline.lua:88-110``: ``hdf5.open(path,'r'):read('/ques_train'):all()``) and writes
them with h5py (``000_prepro_vqa.py:353-366``: ``f.create_dataset("ques_train", dtype='uint32', data=...)``) and
torch-hdf5 (``prepro_img.lua``: ``h5_file:write('/images_train', feat:float())``).  No HDF5 library exists in this image
(no h5py / PyTables / libhdf5), so this module restates the part of the *HDF5 File Format Specification* (version 2.0 /
3.0, public) those files use:

  reader   superblock versions 0-3 (with a user block), version-1 object headers with continuation blocks and
           version-2 ("OHDR") headers, old-style groups (symbol-table message -> version-1 B-tree -> SNOD -> local heap)
           and new-style compact groups (link messages), dataspace messages v1 / v2, fixed-point / floating-point /
           fixed-length-string datatypes, data layout message v1-v3 (compact, contiguous, chunked through the version-1
           chunk B-tree) and the deflate / shuffle / fletcher32 filters.  What it does not know it refuses loudly.
  writer   the classic layout libhdf5 itself produces for ``libver='earliest'``: superblock v0, root group with one
           symbol-table node, version-1 object headers, contiguous (or chunked + deflate) datasets of numeric arrays.

Pinned (tests/test_h5lite.py) against the one file in this image that libhdf5 wrote -- scipy's
``io/matlab/tests/data/testhdf5_7.4_GLNX86.mat`` (MATLAB 7.3 = HDF5 behind a 512-byte user block), whose dataset must
equal the same variable of the MAT-5 twin file read by ``scipy.io.loadmat`` -- and by writer -> reader round trips.  The
writer's output could not be opened with libhdf5 here; it follows the specification field by field.

I/O plumbing only: nothing here is on the timed path.
"""
import mmap
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------
class Dataset:
    def __init__(self, f, name, msgs):
        self._f, self.name, self._msgs = f, name, msgs
        self.shape, self.dtype = f._dataspace(msgs), f._datatype(msgs)

    def __repr__(self):
        return f"<h5lite.Dataset {self.name!r} shape={self.shape} dtype={self.dtype}>"

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, key):
        return self.read()[key]

    def read(self):
        """The whole dataset as a native-endian C-contiguous array (``:all()`` of torch-hdf5, ``[...]`` of h5py)."""
        a = self._f._read_data(self._msgs, self.shape, self.dtype)
        return np.array(a, dtype=a.dtype.newbyteorder("="), order="C", copy=True)     # ONE copy out of the mapped file


class Group:
    def __init__(self, f, name, links):
        self._f, self.name, self._links = f, name, links

    def keys(self):
        return list(self._links)

    def __contains__(self, k):
        return k in self._links

    def __iter__(self):
        return iter(self._links)

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(f"{path!r}: no object {part!r} in group {node.name!r}")
            node = node._f._open(node._links[part], (node.name.rstrip("/") + "/" + part))
        return node


class File(Group):
    """``h5lite.File(path)['/ques_train'].read()`` -- read-only."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            try:                                          # data_img.h5 is gigabytes: map it, copy only what is read
                self._b = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            except ValueError:                            # empty file
                self._b = fh.read()
        self.path = path
        b = self._b
        off = 0
        while off + 8 <= len(b) and b[off:off + 8] != SIGNATURE:          # user block: 512, 1024, 2048, ...
            off = 512 if off == 0 else off * 2
        if off + 8 > len(b):
            raise H5Error(f"{path}: not an HDF5 file (no superblock signature)")
        self._sb = off
        ver = b[off + 8]
        if ver in (0, 1):
            so, sl = b[off + 13], b[off + 14]
            if so != 8 or sl != 8:
                raise H5Error(f"{path}: unsupported offset/length sizes {so}/{sl}")
            p = off + 24 + (4 if ver == 1 else 0)
            self._base, _, _, _ = struct.unpack_from("<QQQQ", b, p)
            p += 32
            _, ohdr, cache, _ = struct.unpack_from("<QQII", b, p)            # root symbol-table entry
            root = ohdr
        elif ver in (2, 3):
            so, sl = b[off + 9], b[off + 10]
            if so != 8 or sl != 8:
                raise H5Error(f"{path}: unsupported offset/length sizes {so}/{sl}")
            self._base, _, _, root = struct.unpack_from("<QQQQ", b, off + 12)
        else:
            raise H5Error(f"{path}: unknown superblock version {ver}")
        if self._base == 0 and off:
            self._base = off                      # addresses are relative to the superblock when a user block precedes it
        self.superblock_version = ver
        Group.__init__(self, self, "/", self._links_of(self._messages(root)))

    def close(self):
        if isinstance(self._b, mmap.mmap):
            self._b.close()
        self._b = b""

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- object headers ----
    def _messages(self, addr):
        """[(type, flags, bytes)] of the object header at ``addr`` (v1 or v2, continuation blocks followed)."""
        b, a = self._b, self._base + addr
        out = []
        if b[a:a + 4] == b"OHDR":
            if b[a + 4] != 2:
                raise H5Error("unknown object header version")
            flags = b[a + 5]
            p = a + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            nb = 1 << (flags & 3)
            size = int.from_bytes(b[p:p + nb], "little")
            p += nb
            blocks = [(p, p + size)]
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype, msize, mflags = b[p], struct.unpack_from("<H", b, p + 1)[0], b[p + 3]
                    p += 4 + (2 if flags & 0x04 else 0)
                    data = b[p:p + msize]
                    p += msize
                    if mtype == 0x10:
                        coff, clen = struct.unpack_from("<QQ", data)
                        ca = self._base + coff
                        if b[ca:ca + 4] != b"OCHK":
                            raise H5Error("bad object header continuation block")
                        blocks.append((ca + 4, ca + clen - 4))
                    elif mtype != 0:
                        out.append((mtype, mflags, data))
            return out
        ver, _, nmsgs, _, hsize = struct.unpack_from("<BBHII", b, a)
        if ver != 1:
            raise H5Error(f"object header at {addr}: unknown version {ver}")
        blocks = [(a + 16, a + 16 + hsize)]
        seen = 0
        while blocks and seen < nmsgs:
            p, end = blocks.pop(0)
            while p + 8 <= end and seen < nmsgs:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                seen += 1
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", data)
                    blocks.append((self._base + coff, self._base + coff + clen))
                elif mtype != 0:
                    out.append((mtype, mflags, data))
        return out

    @staticmethod
    def _find(msgs, mtype):
        for t, _, d in msgs:
            if t == mtype:
                return d
        return None

    def _open(self, addr, name):
        msgs = self._messages(addr)
        if self._find(msgs, 0x08) is not None:
            return Dataset(self, name, msgs)
        return Group(self, name, self._links_of(msgs))

    # ---- groups ----
    def _links_of(self, msgs):
        links = {}
        st = self._find(msgs, 0x11)
        if st is not None:                                   # old style: symbol table
            btree, heap = struct.unpack_from("<QQ", st)
            hb = self._base + heap
            if self._b[hb:hb + 4] != b"HEAP":
                raise H5Error("bad local heap signature")
            seg = self._base + struct.unpack_from("<Q", self._b, hb + 24)[0]
            self._walk_group_btree(btree, seg, links)
            return links
        for t, _, d in msgs:                                 # new style, compact: one link message per member
            if t == 0x02:
                fr = struct.unpack_from("<Q", d, 2 + (8 if d[1] & 1 else 0))[0]
                if fr != UNDEF:
                    raise H5Error("dense link storage (fractal heap) is not supported")
            if t == 0x06:
                ver, fl = d[0], d[1]
                p = 2
                ltype = 0
                if fl & 0x08:
                    ltype = d[p]; p += 1
                if fl & 0x04:
                    p += 8
                if fl & 0x10:
                    p += 1
                nb = 1 << (fl & 3)
                n = int.from_bytes(d[p:p + nb], "little"); p += nb
                nm = d[p:p + n].decode(); p += n
                if ver == 1 and ltype == 0:
                    links[nm] = struct.unpack_from("<Q", d, p)[0]
        return links

    def _walk_group_btree(self, addr, seg, links):
        b, a = self._b, self._base + addr
        if b[a:a + 4] == b"SNOD":
            n = struct.unpack_from("<H", b, a + 6)[0]
            for i in range(n):
                noff, ohdr = struct.unpack_from("<QQ", b, a + 8 + 40 * i)
                s = seg + noff
                links[b[s:b.find(b"\0", s)].decode()] = ohdr
            return
        if b[a:a + 4] != b"TREE" or b[a + 4] != 0:
            raise H5Error("bad group B-tree node")
        n = struct.unpack_from("<H", b, a + 6)[0]
        p = a + 24 + 8                                        # skip key 0
        for _ in range(n):
            child = struct.unpack_from("<Q", b, p)[0]
            self._walk_group_btree(child, seg, links)
            p += 16

    # ---- dataset messages ----
    def _dataspace(self, msgs):
        d = self._find(msgs, 0x01)
        if d is None:
            raise H5Error("dataset without a dataspace message")
        ver, rank = d[0], d[1]
        if ver == 1:
            p = 8
        elif ver == 2:
            if d[3] == 2:
                raise H5Error("null dataspace")
            p = 4
        else:
            raise H5Error(f"dataspace message version {ver}")
        return tuple(struct.unpack_from("<Q", d, p + 8 * i)[0] for i in range(rank))

    def _datatype(self, msgs):
        d = self._find(msgs, 0x03)
        if d is None:
            raise H5Error("dataset without a datatype message")
        cls, bits0, size = d[0] & 0x0F, d[1], struct.unpack_from("<I", d, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        if cls == 1:
            if size not in (2, 4, 8):
                raise H5Error(f"floating-point type of {size} bytes")
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise H5Error(f"datatype class {cls} is not supported (numeric and fixed-length string types only)")

    def _filters(self, msgs):
        d = self._find(msgs, 0x0B)
        if d is None:
            return []
        ver, n = d[0], d[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack_from("<H", d, p)[0]; p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", d, p)[0]; p += 2
            _, ncd = struct.unpack_from("<HH", d, p); p += 4
            p += _pad8(nlen) if ver == 1 else nlen
            cd = struct.unpack_from(f"<{ncd}I", d, p); p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _read_data(self, msgs, shape, dtype):
        d = self._find(msgs, 0x08)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        nbytes = n * dtype.itemsize
        ver = d[0]
        if ver == 3:
            cls = d[1]
            if cls == 0:
                size = struct.unpack_from("<H", d, 2)[0]
                raw = d[4:4 + size]
            elif cls == 1:
                addr, size = struct.unpack_from("<QQ", d, 2)
                if addr == UNDEF:
                    return np.zeros(shape, dtype)
                if self._base + addr + nbytes > len(self._b):
                    raise H5Error("dataset extends past the end of the file")
                return np.frombuffer(self._b, dtype=dtype, count=n, offset=self._base + addr).reshape(shape)   # a view
            elif cls == 2:
                ndim = d[2]
                btree = struct.unpack_from("<Q", d, 3)[0]
                cdims = struct.unpack_from(f"<{ndim}I", d, 11)
                return self._read_chunked(btree, cdims[:-1], shape, dtype, self._filters(msgs))
            else:
                raise H5Error(f"data layout class {cls}")
        elif ver in (1, 2):
            ndim, cls = d[1], d[2]
            p = 8
            addr = UNDEF
            if cls != 0:
                addr = struct.unpack_from("<Q", d, p)[0]; p += 8
            dims = struct.unpack_from(f"<{ndim}I", d, p); p += 4 * ndim
            if cls == 1:
                raw = None if addr == UNDEF else self._b[self._base + addr:self._base + addr + nbytes]
            elif cls == 2:
                return self._read_chunked(addr, dims[:-1] if len(dims) == len(shape) + 1 else dims, shape, dtype, self._filters(msgs))
            else:
                size = struct.unpack_from("<I", d, p)[0]
                raw = d[p + 4:p + 4 + size]
        else:
            raise H5Error(f"data layout message version {ver} is not supported (write the file with libver='earliest')")
        if raw is None:
            return np.zeros(shape, dtype)
        if len(raw) < nbytes:
            raise H5Error("dataset extends past the end of the file")
        return np.frombuffer(raw[:nbytes], dtype).reshape(shape)

    def _read_chunked(self, btree, cdims, shape, dtype, filters):
        out = np.zeros(shape, dtype)
        if btree == UNDEF or 0 in shape:
            return out
        rank = len(shape)
        csize = int(np.prod(cdims)) * dtype.itemsize

        def walk(addr):
            b, a = self._b, self._base + addr
            if b[a:a + 4] != b"TREE" or b[a + 4] != 1:
                raise H5Error("bad chunk B-tree node")
            level, n = b[a + 5], struct.unpack_from("<H", b, a + 6)[0]
            ksz = 8 + 8 * (rank + 1)
            p = a + 24
            for _ in range(n):
                nbytes, mask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from(f"<{rank}Q", b, p + 8)
                child = struct.unpack_from("<Q", b, p + ksz)[0]
                p += ksz + 8
                if level > 0:
                    walk(child)
                    continue
                raw = b[self._base + child:self._base + child + nbytes]
                for i, (fid, cd) in reversed(list(enumerate(filters))):
                    if mask & (1 << i):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else dtype.itemsize
                        raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                    elif fid == 3:
                        raw = raw[:-4]
                    else:
                        raise H5Error(f"filter {fid} is not supported (deflate, shuffle, fletcher32 only)")
                chunk = np.frombuffer(raw[:csize], dtype).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

        walk(btree)
        return out


# ------------------------------------------------------------------------------------------------
# writer (classic layout: superblock v0, one symbol-table node, v1 object headers)
# ------------------------------------------------------------------------------------------------
def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits0 = (0x08 if dt.kind == "i" else 0)
        return struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        # IEEE little-endian: sign position, exponent location/size, mantissa location/size, exponent bias
        if dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            sign = 31
        else:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            sign = 63
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, dt.itemsize) + props
    raise H5Error(f"h5lite.write: dtype {dt} is not supported")


def _message(mtype, data, flags=0):
    data = data + b"\0" * (_pad8(len(data)) - len(data))
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4 + body


def write(path, datasets, chunks=None, compression=None):
    """Writes ``{name: array}`` as top-level datasets (what ``f.create_dataset(name, data=a)`` produces).  ``chunks``
    (rows per chunk along the first dimension) with ``compression='gzip'`` writes chunked, deflated datasets instead of
    contiguous ones."""
    names = sorted(datasets)                                 # symbol-table entries are ordered by name
    arrays = {k: np.ascontiguousarray(datasets[k]) for k in names}
    for k, a in arrays.items():
        if a.dtype.byteorder == ">":
            arrays[k] = a.astype(a.dtype.newbyteorder("<"))
    K = max(4, (len(names) + 1) // 2)                        # group leaf node K: one SNOD holds up to 2K entries
    buf = bytearray(96)                                      # superblock v0 (8 + 8 + 8 + 32 + 40 bytes), filled last

    def alloc(data):
        off = len(buf)
        buf.extend(data)
        buf.extend(b"\0" * (_pad8(len(buf)) - len(buf)))
        return off

    # local heap data segment: "" at offset 0, then the names
    seg = bytearray(8)
    name_off = {}
    for k in names:
        name_off[k] = len(seg)
        e = k.encode() + b"\0"
        seg.extend(e + b"\0" * (_pad8(len(e)) - len(e)))
    free_off = len(seg)
    seg.extend(struct.pack("<QQ", 1, 16))                   # one free block: next = 1 (none), size 16
    ohdr = {}
    for k in names:
        a = arrays[k]
        space = struct.pack("<BBBBI", 1, a.ndim, 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in a.shape)
        msgs = [_message(0x01, space), _message(0x03, _dtype_message(a.dtype), flags=1)]
        if chunks and a.ndim >= 1 and a.shape[0] > 0:
            rows = max(1, int(chunks))
            cdims = (rows,) + a.shape[1:]
            level0 = []
            for r0 in range(0, a.shape[0], rows):
                blk = np.zeros(cdims, a.dtype)
                part = a[r0:r0 + rows]
                blk[:part.shape[0]] = part
                raw = blk.tobytes()
                if compression == "gzip":
                    raw = zlib.compress(raw, 4)
                level0.append((len(raw), (r0,) + (0,) * (a.ndim - 1), alloc(raw)))
            node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(level0), UNDEF, UNDEF))
            for nbytes, offs, addr in level0:
                node += struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)
                node += struct.pack("<Q", addr)
            node += struct.pack("<II", 0, 0) + struct.pack("<Q", a.shape[0] + (-a.shape[0]) % rows) + b"".join(
                struct.pack("<Q", 0) for _ in range(a.ndim))           # final key: one past the last chunk
            bt = alloc(bytes(node))
            if compression == "gzip":
                msgs.append(_message(0x0B, struct.pack("<BBHI", 1, 1, 0, 0) + struct.pack("<HHHH", 1, 0, 1, 1) +
                                     struct.pack("<II", 4, 0), flags=1))
            layout = struct.pack("<BBB", 3, 2, a.ndim + 1) + struct.pack("<Q", bt) + b"".join(
                struct.pack("<I", c) for c in cdims) + struct.pack("<I", a.dtype.itemsize)
        else:
            addr = alloc(a.tobytes()) if a.size else UNDEF
            layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", addr, a.nbytes)
        msgs.append(_message(0x08, layout))
        ohdr[k] = alloc(_object_header(msgs))
    heap_seg = alloc(bytes(seg))
    heap = alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(seg), free_off, heap_seg))
    snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
    for k in names:
        snod += struct.pack("<QQII", name_off[k], ohdr[k], 0, 0) + b"\0" * 16
    snod += b"\0" * (8 + 40 * 2 * K - len(snod))
    snod_off = alloc(bytes(snod))
    tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
    tree += struct.pack("<QQQ", 0, snod_off, name_off[names[-1]] if names else 0)
    tree += b"\0" * (24 + 8 * (2 * 16 + 1) + 8 * 2 * 16 - len(tree))      # internal node K = 16
    tree_off = alloc(bytes(tree))
    root = alloc(_object_header([_message(0x11, struct.pack("<QQ", tree_off, heap))]))
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, K, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(buf), UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_off, heap)
    assert len(sb) == 96
    buf[:96] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(buf))
