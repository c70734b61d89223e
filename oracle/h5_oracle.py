"""h5_oracle.py -- TEST INFRASTRUCTURE.  An independent pure-Python restatement of the HDF5 File Format Specification
(version 3.0) for the subset MG-CFD decks need, used to cross-check the C++ implementation
(mg-cfd-app-op2_b200/host/h5lite.hpp) in both directions: files written here are read by the C++ reader, files written
by the C++ writer are read here.  The reference reads its level files through OP2's op_decl_*_hdf5 (euler3d.cpp:248-327)
and writes solutions through op_fetch_data_hdf5_file (euler3d.cpp:564,740-770); libhdf5 / h5py do not exist in this
image.  Pinning: the one genuine libhdf5-written file the image holds (tests/golden/libhdf5_matlab73_testdouble.mat, a
MATLAB 7.3 MAT-file from scipy's test data) is read here and by the C++ reader with identical, known results; the
chunked / filtered / version-2 structures have no genuine sample: parity unpinned for those (stated in DESIGN.md).

write_h5(path, datasets, ...) can produce every structural variant the C++ reader claims to handle:
  superblock 0 (symbol-table root group: B-tree "TREE" + "SNOD" + local heap) or 2 ("OHDR" v2 root with Link messages),
  contiguous / compact / chunked storage (chunk B-tree with optional shuffle + deflate + fletcher32 trailer),
  little- or big-endian elements, nested groups ("a/b"), a user block, attributes (ints, doubles, strings).
read_h5(path) -> {name: {"data": ndarray, "attrs": {...}}} walks the same structures generically."""
import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# --------------------------------------------------------------------------------------------- messages
def _type_msg(dtype):
    dt = np.dtype(dtype)
    big = dt.byteorder == ">"
    if dt.kind in "iu":
        bits = (1 if big else 0) | (8 if dt.kind == "i" else 0)
        return struct.pack("<BBBBI", 0x10, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f":
        dbl = dt.itemsize == 8
        return (struct.pack("<BBBBI", 0x11, 0x20 | (1 if big else 0), 63 if dbl else 31, 0, dt.itemsize) +
                struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, 52 if dbl else 23, 11 if dbl else 8, 0, 52 if dbl else 23,
                            1023 if dbl else 127))
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0, 0, 0, dt.itemsize)
    raise TypeError(dt)


def _space_msg(shape, version=1):
    if version == 1:
        return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", d) for d in shape)
    return struct.pack("<BBBB", 2, len(shape), 0, 1 if shape else 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _attr_msg(name, value, version=1):
    if isinstance(value, str):
        raw = value.encode() + b"\0"
        t, s = _type_msg(f"S{len(raw)}"), _space_msg(())
    else:
        a = np.atleast_1d(np.asarray(value))
        raw, t, s = a.tobytes(), _type_msg(a.dtype), _space_msg(a.shape)
    nm = name.encode() + b"\0"
    if version == 1:
        return struct.pack("<BBHHH", 1, 0, len(nm), len(t), len(s)) + _pad8(nm) + _pad8(t) + _pad8(s) + raw
    return struct.pack("<BBHHHB", 3, 0, len(nm), len(t), len(s), 0) + nm + t + s + raw


def _v1_header(messages):
    body = b""
    for mtype, data in messages:
        data = _pad8(data)
        body += struct.pack("<HHBBH", mtype, len(data), 0, 0, 0) + data
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4 + body


def _v2_header(messages):
    body = b""
    for mtype, data in messages:
        body += struct.pack("<BHB", mtype, len(data), 0) + data
    head = b"OHDR" + struct.pack("<BB", 2, 0x02) + struct.pack("<I", len(body))     # flags: 4-byte chunk size
    blob = head + body
    return blob + struct.pack("<I", zlib.crc32(blob) & 0xFFFFFFFF)      # checksum field (not validated by the readers)


class _Alloc:
    """file image under construction: append-only, 8-byte aligned pieces, addresses relative to the base address"""

    def __init__(self):
        self.buf = bytearray()

    def put(self, blob):
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += blob
        return addr

    def patch(self, addr, blob):
        self.buf[addr:addr + len(blob)] = blob


def _chunk_btree(img, arr, chunk, filters):
    """chunks + one-level v1 B-tree (node type 1); returns the B-tree address"""
    rank, es = arr.ndim, arr.dtype.itemsize
    grid = [range(0, arr.shape[r], chunk[r]) for r in range(rank)]
    entries = []
    for off in np.ndindex(*[len(g) for g in grid]):
        start = [grid[r][off[r]] for r in range(rank)]
        block = np.zeros(chunk, dtype=arr.dtype)
        sl = tuple(slice(start[r], min(start[r] + chunk[r], arr.shape[r])) for r in range(rank))
        part = arr[sl]
        block[tuple(slice(0, s) for s in part.shape)] = part
        raw = block.tobytes()
        for f in filters:
            if f == "shuffle":
                raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, es).T.tobytes()
            elif f == "deflate":
                raw = zlib.compress(raw, 6)
            elif f == "fletcher32":
                raw = raw + b"\0\0\0\0"          # trailer (readers strip it without validating)
        entries.append((start, len(raw), img.put(raw)))
    assert len(entries) <= 64, "test writer: one B-tree node only"
    node = b"TREE" + struct.pack("<BBH", 1, 0, len(entries)) + struct.pack("<QQ", UNDEF, UNDEF)
    for start, size, addr in entries:
        node += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", s) for s in start) + struct.pack("<Q", 0)
        node += struct.pack("<Q", addr)
    node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape) + struct.pack("<Q", 0)   # last key
    return img.put(node)


def _dataset_header(img, arr, layout, chunk, filters, attrs, v2):
    msgs = [(0x01, _space_msg(arr.shape, 2 if v2 else 1)), (0x03, _type_msg(arr.dtype)), (0x05, struct.pack("<BBBB", 2, 2, 2, 0))]
    if layout == "contiguous":
        addr = img.put(arr.tobytes()) if arr.size else UNDEF
        msgs.append((0x08, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)))
    elif layout == "compact":
        msgs.append((0x08, struct.pack("<BBH", 3, 0, arr.nbytes) + arr.tobytes()))
    else:
        if filters:
            ids = {"deflate": 1, "shuffle": 2, "fletcher32": 3}
            pipe = struct.pack("<BB", 1, len(filters)) + b"\0" * 6
            for f in filters:
                client = [6] if f == "deflate" else ([arr.dtype.itemsize] if f == "shuffle" else [])
                pipe += struct.pack("<HHHH", ids[f], 0, 0, len(client)) + b"".join(struct.pack("<I", c) for c in client)
                if len(client) % 2:
                    pipe += b"\0" * 4
            msgs.append((0x0B, pipe))
        bt = _chunk_btree(img, arr, chunk, filters)
        msgs.append((0x08, struct.pack("<BBB", 3, 2, arr.ndim + 1) + struct.pack("<Q", bt) +
                     b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", arr.dtype.itemsize)))
    for k, v in (attrs or {}).items():
        msgs.append((0x0C, _attr_msg(k, v, 3 if v2 else 1)))
    return img.put(_v2_header(msgs) if v2 else _v1_header(msgs))


def _old_group(img, entries):
    """entries: sorted [(name, object header address)] -> object header address of an old-style group"""
    heap = bytearray(b"\0" * 8)
    offs = []
    for name, _ in entries:
        offs.append(len(heap))
        heap += _pad8(name.encode() + b"\0")
    seg = img.put(bytes(heap))
    hp = img.put(b"HEAP" + struct.pack("<BBH", 0, 0, 0) + struct.pack("<QQQ", len(heap), 1, seg))
    snods, keys = [], [0]
    for lo in range(0, max(len(entries), 1), 8):
        part = entries[lo:lo + 8]
        node = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
        for i, (name, addr) in enumerate(part):
            node += struct.pack("<QQIIQQ", offs[lo + i], addr, 0, 0, 0, 0)
        node += b"\0" * (40 * (8 - len(part)))
        snods.append(img.put(node))
        keys.append(offs[lo + len(part) - 1] if part else 0)
    tree = b"TREE" + struct.pack("<BBH", 0, 0, len(snods)) + struct.pack("<QQ", UNDEF, UNDEF) + struct.pack("<Q", keys[0])
    for s, k in zip(snods, keys[1:]):
        tree += struct.pack("<QQ", s, k)
    tree += b"\0" * (16 * (32 - len(snods)))
    bt = img.put(tree)
    return img.put(_v1_header([(0x11, struct.pack("<QQ", bt, hp))])), bt, hp


def write_h5(path, datasets, superblock=0, layout="contiguous", chunk=None, filters=(), attrs=None, userblock=0):
    """datasets: {name or "group/name": ndarray}; attrs: {dataset name: {attr: value}}"""
    img = _Alloc()
    v2 = superblock >= 2
    img.put(b"\0" * (48 if v2 else 96))                 # superblock placeholder
    tree = {}
    for name, arr in datasets.items():
        arr = np.ascontiguousarray(arr)
        ch = tuple(chunk) if chunk else tuple(max(1, (d + 1) // 2) for d in arr.shape)
        addr = _dataset_header(img, arr, layout, ch, list(filters), (attrs or {}).get(name), v2)
        node = tree
        parts = name.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = addr

    def build(node):
        entries = sorted((k, build(v) if isinstance(v, dict) else v) for k, v in node.items())
        if v2:
            msgs = [(0x02, struct.pack("<BB", 0, 0) + struct.pack("<QQ", UNDEF, UNDEF))]     # link info: no dense storage
            for name, addr in entries:
                nm = name.encode()
                msgs.append((0x06, struct.pack("<BBB", 1, 0, len(nm)) + nm + struct.pack("<Q", addr)))
            return img.put(_v2_header(msgs))
        return _old_group(img, entries)[0]

    if v2:
        root = build(tree)
        eof = len(img.buf)
        sb = SIG + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", userblock, UNDEF, eof, root)
        sb += struct.pack("<I", zlib.crc32(sb) & 0xFFFFFFFF)
    else:
        entries = sorted((k, build(v) if isinstance(v, dict) else v) for k, v in tree.items())
        root, bt, hp = _old_group(img, entries)
        eof = len(img.buf)
        sb = (SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0) +
              struct.pack("<QQQQ", userblock, UNDEF, eof, UNDEF) + struct.pack("<QQIIQQ", 0, root, 1, 0, bt, hp))
    img.patch(0, sb)
    with open(path, "wb") as f:
        f.write(b"\0" * userblock)                      # a user block: the superblock sits at 512, 1024, ...
        f.write(bytes(img.buf))


# --------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        off = 0
        while self.b[off:off + 8] != SIG:
            off = off * 2 if off else 512
            if off + 8 > len(self.b):
                raise ValueError(f"{path}: not an HDF5 file")
        self.base = off
        v = self.b[off + 8]
        if v in (0, 1):
            assert self.b[off + 13] == 8 and self.b[off + 14] == 8
            p = off + 24 + (4 if v == 1 else 0)
            self.root = struct.unpack_from("<Q", self.b, p + 32 + 8)[0]
        elif v in (2, 3):
            self.root = struct.unpack_from("<Q", self.b, off + 12 + 24)[0]
        else:
            raise ValueError("superblock version")
        self.version = v

    def at(self, addr, n):
        return self.b[self.base + addr:self.base + addr + n]

    def messages(self, addr):
        out = []
        if self.at(addr, 4) == b"OHDR":
            flags = self.at(addr, 6)[5]
            p = addr + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
            szlen = 1 << (flags & 3)
            size = int.from_bytes(self.at(p, szlen), "little")
            blocks = [(p + szlen, size)]
            for start, length in blocks:
                q = start
                while q + 4 <= start + length:
                    mtype, msize, _ = struct.unpack("<BHB", self.at(q, 4))
                    q += 4 + (2 if flags & 4 else 0)
                    data = self.at(q, msize)
                    q += msize
                    if mtype == 0x10:
                        caddr, clen = struct.unpack("<QQ", data)
                        blocks.append((caddr + 4, clen - 8))
                    elif mtype:
                        out.append((mtype, data))
            return out
        _, _, n, _, size = struct.unpack("<BBHII", self.at(addr, 12))
        blocks, seen = [(addr + 16, size)], 0
        for start, length in blocks:
            q = start
            while q + 8 <= start + length and seen < n:
                mtype, msize = struct.unpack("<HH", self.at(q, 4))
                data = self.at(q + 8, msize)
                q += 8 + msize
                seen += 1
                if mtype == 0x10:
                    blocks.append(struct.unpack("<QQ", data))
                elif mtype:
                    out.append((mtype, data))
        return out

    def children(self, msgs):
        kids = []
        for mtype, data in msgs:
            if mtype == 0x11:
                bt, hp = struct.unpack("<QQ", data[:16])
                _, _, _, seg_size, _, seg = struct.unpack("<4sB3sQQQ", self.at(hp, 32))
                heap = self.at(seg, seg_size)

                def walk(node):
                    sig = self.at(node, 4)
                    if sig == b"SNOD":
                        n = struct.unpack("<H", self.at(node + 6, 2))[0]
                        for i in range(n):
                            noff, addr = struct.unpack("<QQ", self.at(node + 8 + 40 * i, 16))
                            kids.append((heap[noff:heap.index(b"\0", noff)].decode(), addr))
                    else:
                        assert sig == b"TREE"
                        used = struct.unpack("<H", self.at(node + 6, 2))[0]
                        for i in range(used):
                            walk(struct.unpack("<Q", self.at(node + 24 + 16 * i + 8, 8))[0])
                walk(bt)
            elif mtype == 0x06:
                flags, q, ltype = data[1], 2, 0
                if flags & 8:
                    ltype = data[q]; q += 1
                if flags & 4:
                    q += 8
                if flags & 16:
                    q += 1
                nl = 1 << (flags & 3)
                ln = int.from_bytes(data[q:q + nl], "little"); q += nl
                name = data[q:q + ln].decode(); q += ln
                if ltype == 0:
                    kids.append((name, struct.unpack("<Q", data[q:q + 8])[0]))
        return kids

    @staticmethod
    def dtype(msg):
        cls, bits, b2 = msg[0] & 15, msg[1], msg[2]
        size = struct.unpack("<I", msg[4:8])[0]
        order = ">" if bits & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits & 8 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise TypeError(cls)

    @staticmethod
    def shape(msg):
        rank = msg[1]
        q = 8 if msg[0] == 1 else 4
        return tuple(struct.unpack_from("<Q", msg, q + 8 * i)[0] for i in range(rank))

    def dataset(self, msgs):
        shape = dt = None
        data, attrs, filters = None, {}, []
        for mtype, m in msgs:
            if mtype == 0x01:
                shape = self.shape(m)
            elif mtype == 0x03:
                dt = self.dtype(m)
        for mtype, m in msgs:
            if mtype == 0x0B:
                assert m[0] == 1
                q = 8
                for _ in range(m[1]):
                    fid, nlen, _, nc = struct.unpack_from("<HHHH", m, q)
                    q += 8 + ((nlen + 7) & ~7) + 4 * nc + (4 if nc % 2 else 0)
                    filters.append(fid)
            elif mtype == 0x0C:
                ver = m[0]
                nsz, tsz, ssz = struct.unpack_from("<HHH", m, 2)
                q = 9 if ver == 3 else 8
                pad = (lambda x: (x + 7) & ~7) if ver == 1 else (lambda x: x)
                name = m[q:q + nsz].split(b"\0")[0].decode(); q += pad(nsz)
                adt = self.dtype(m[q:q + tsz]); q += pad(tsz)
                ashape = self.shape(m[q:q + ssz]); q += pad(ssz)
                cnt = int(np.prod(ashape)) if ashape else 1
                val = np.frombuffer(m[q:q + cnt * adt.itemsize], dtype=adt)
                attrs[name] = val[0].split(b"\0")[0].decode() if adt.kind == "S" else val.reshape(ashape or (1,))
        for mtype, m in msgs:
            if mtype != 0x08:
                continue
            n = int(np.prod(shape)) if shape else 1
            if m[0] in (1, 2):
                # versions 1 and 2 (HDF5 1.6 and older writers): dimensionality, class, 5 reserved bytes, the address (not
                # for compact storage), `dimensionality` 4-byte sizes (chunked: the chunk shape followed by the element size)
                ndim, cls, q = m[1], m[2], 8
                addr = UNDEF
                if cls != 0:
                    addr = struct.unpack_from("<Q", m, q)[0]
                    q += 8
                dims = struct.unpack_from(f"<{ndim}I", m, q)
                q += 4 * ndim
                if cls == 0:
                    size = struct.unpack_from("<I", m, q)[0]
                    compact = m[q + 4:q + 4 + size]
                else:
                    rank, bt, chunk = ndim - 1, addr, dims[:-1]
            else:
                assert m[0] == 3
                cls = m[1]
                if cls == 1:
                    addr = struct.unpack_from("<Q", m, 2)[0]
                elif cls == 0:
                    size = struct.unpack_from("<H", m, 2)[0]
                    compact = m[4:4 + size]
                else:
                    rank = m[2] - 1
                    bt = struct.unpack_from("<Q", m, 3)[0]
                    chunk = struct.unpack_from(f"<{rank}I", m, 11)
            if cls == 1:
                raw = self.at(addr, n * dt.itemsize) if addr != UNDEF else b"\0" * (n * dt.itemsize)
                data = np.frombuffer(raw, dtype=dt).reshape(shape)
            elif cls == 0:
                data = np.frombuffer(compact, dtype=dt).reshape(shape)
            else:
                data = np.zeros(shape, dtype=dt)

                def walk(node):
                    _, ntype, level, used = struct.unpack("<4sBBH", self.at(node, 8))
                    ksz = 8 + 8 * (rank + 1)
                    for i in range(used):
                        k = self.at(node + 24 + i * (ksz + 8), ksz + 8)
                        size, mask = struct.unpack_from("<II", k)
                        off = struct.unpack_from(f"<{rank}Q", k, 8)
                        child = struct.unpack_from("<Q", k, ksz)[0]
                        if level:
                            walk(child)
                            continue
                        raw = self.at(child, size)
                        for fi in reversed(range(len(filters))):
                            if mask >> fi & 1:
                                continue
                            if filters[fi] == 1:
                                raw = zlib.decompress(raw)
                            elif filters[fi] == 2:
                                raw = np.frombuffer(raw, dtype=np.uint8).reshape(dt.itemsize, -1).T.tobytes()
                            elif filters[fi] == 3:
                                raw = raw[:-4]
                        block = np.frombuffer(raw, dtype=dt).reshape(chunk)
                        sl = tuple(slice(off[r], min(off[r] + chunk[r], shape[r])) for r in range(rank))
                        data[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
                walk(bt)
        return {"data": data, "attrs": attrs}

    def walk(self, addr, prefix, out):
        for name, child in self.children(self.messages(addr)):
            msgs = self.messages(child)
            full = f"{prefix}/{name}" if prefix else name
            if any(t == 0x08 for t, _ in msgs):
                out[full] = self.dataset(msgs)
            else:
                self.walk(child, full, out)


def read_h5(path):
    r = _Reader(path)
    out = {}
    r.walk(r.root, "", out)
    return out
