"""HDF5 subset reader / writer (mg-cfd-app-op2_b200/host/h5lite.hpp behind include/mgcfd_h5.h) against the independent
pure-Python restatement of the file-format specification (oracle/h5_oracle.py), in both directions, over every
structural variant the reader claims: superblock 0 / 2, object headers v1 / v2, old-style and link-message groups,
contiguous / compact / chunked storage with shuffle + deflate + fletcher32, either byte order, user blocks, nested
groups, attributes.  No libhdf5 exists in the image; the one genuine libhdf5-written file it holds (a MATLAB 7.3 MAT-file
among scipy's test data, tests/golden/libhdf5_matlab73_testdouble.mat) pins the structures OP2's level files use; the
chunked / filtered / version-2 variants stay unpinned against the real library."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import h5_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mg-cfd-app-op2_b200", "libmgcfd_h5.so")


@pytest.fixture(scope="module")
def h5():
    lib = C.CDLL(LIB)
    lib.mgcfd_h5_open.restype = C.c_void_p
    lib.mgcfd_h5_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    lib.mgcfd_h5_close.argtypes = [C.c_void_p]
    lib.mgcfd_h5_count.argtypes = [C.c_void_p]
    lib.mgcfd_h5_superblock_version.argtypes = [C.c_void_p]
    lib.mgcfd_h5_name.restype = C.c_char_p
    lib.mgcfd_h5_name.argtypes = [C.c_void_p, C.c_int]
    lib.mgcfd_h5_info.argtypes = [C.c_void_p, C.c_char_p] + [C.POINTER(C.c_int)] * 5 + [C.POINTER(C.c_ulonglong)]
    lib.mgcfd_h5_read_f64.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int]
    lib.mgcfd_h5_read_i32.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int]
    lib.mgcfd_h5_attr_int.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_longlong)]
    lib.mgcfd_h5_attr_str.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    lib.mgcfd_h5_create.restype = C.c_void_p
    lib.mgcfd_h5_create.argtypes = [C.c_char_p]
    lib.mgcfd_h5_add.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.c_void_p]
    lib.mgcfd_h5_finish.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.mgcfd_h5_is_hdf5.argtypes = [C.c_char_p]
    return lib


def c_read(lib, path):
    """every dataset of a file through the C-ABI: {name: (ndarray, info)}"""
    err = C.create_string_buffer(512)
    r = lib.mgcfd_h5_open(path.encode(), err, 512)
    assert r, err.value.decode()
    out = {}
    try:
        for i in range(lib.mgcfd_h5_count(r)):
            name = lib.mgcfd_h5_name(r, i)
            cls, es, sg, lay, rank = (C.c_int() for _ in range(5))
            dims = (C.c_ulonglong * 8)()
            assert lib.mgcfd_h5_info(r, name, cls, es, sg, lay, rank, dims) == 0
            shape = tuple(dims[k] for k in range(rank.value))
            if cls.value == 1:
                a = np.empty(shape, dtype=np.float64)
                assert lib.mgcfd_h5_read_f64(r, name, a.ctypes.data, err, 512) == 0, err.value.decode()
            else:
                a = np.empty(shape, dtype=np.int32)
                assert lib.mgcfd_h5_read_i32(r, name, a.ctypes.data, err, 512) == 0, err.value.decode()
            out[name.decode()] = (a, {"class": cls.value, "bytes": es.value, "layout": lay.value})
        out["__superblock__"] = lib.mgcfd_h5_superblock_version(r)
    finally:
        lib.mgcfd_h5_close(r)
    return out


def deck_like(rng, n=57, e=190, b=23):
    return {
        "node_coordinates": rng.standard_normal((n, 3)),
        "edge-->node": rng.integers(1, n + 1, size=(e, 2)).astype(np.int32),
        "edge_weights": rng.standard_normal((e, 3)),
        "bnd_node-->node": rng.integers(1, n + 1, size=(b, 1)).astype(np.int32),
        "bnd_node-->group": rng.integers(0, 10, size=(b, 1)).astype(np.int32),
        "bnd_node_weights": rng.standard_normal((b, 3)),
        "node-->mg_node": rng.integers(1, n // 2, size=(n, 1)).astype(np.int32),
    }


VARIANTS = [
    dict(),
    dict(superblock=2),
    dict(layout="compact"),
    dict(layout="chunked"),
    dict(layout="chunked", chunk=None, filters=("deflate",)),
    dict(layout="chunked", filters=("shuffle", "deflate")),
    dict(layout="chunked", filters=("shuffle", "deflate", "fletcher32"), superblock=2),
    dict(userblock=512),
    dict(userblock=2048, superblock=2),
]


@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: ",".join(f"{k}={v[k]}" for k in v) or "default")
def test_cpp_reads_what_the_python_restatement_writes(h5, tmp_path, variant):
    rng = np.random.default_rng(7)
    data = deck_like(rng)
    path = str(tmp_path / "level.h5")
    h5_oracle.write_h5(path, data, **variant)
    assert h5.mgcfd_h5_is_hdf5(path.encode()) == 1
    got = c_read(h5, path)
    assert got.pop("__superblock__") == variant.get("superblock", 0)
    assert sorted(got) == sorted(data)
    for name, ref in data.items():
        a, info = got[name]
        assert a.shape == ref.shape and np.array_equal(a, ref), name
        assert info["layout"] == {"compact": 0, "contiguous": 1, "chunked": 2}[variant.get("layout", "contiguous")]


def test_element_types_byte_orders_and_ragged_chunks(h5, tmp_path):
    rng = np.random.default_rng(3)
    data = {
        "be_f64": rng.standard_normal((9, 5)).astype(">f8"),
        "le_f32": rng.standard_normal((4, 3)).astype("<f4"),
        "be_i64": rng.integers(-2**31, 2**31 - 1, size=(11,)).astype(">i8"),
        "le_u16": rng.integers(0, 65535, size=(6, 2)).astype("<u2"),
        "i8": rng.integers(-128, 127, size=(13,)).astype("i1"),
        "scalarish": np.array([42], dtype=np.int32),
        "empty": np.zeros((0, 3)),
        "rank3": rng.standard_normal((5, 4, 3)),
    }
    for variant in (dict(), dict(layout="chunked", chunk=None, filters=("shuffle", "deflate"))):
        path = str(tmp_path / "types.h5")
        if variant:
            h5_oracle.write_h5(path, {k: v for k, v in data.items() if v.size}, **variant)
        else:
            h5_oracle.write_h5(path, data)
        got = c_read(h5, path)
        for name, ref in data.items():
            if name not in got:
                continue
            a, info = got[name]
            assert info["bytes"] == ref.dtype.itemsize
            assert a.shape == ref.shape
            assert np.array_equal(a.astype(np.float64), ref.astype(np.float64)), name
    # chunks that do not divide the extents (7x3 chunks over 9x5: edge chunks are partly outside the dataset)
    path = str(tmp_path / "ragged.h5")
    h5_oracle.write_h5(path, {"a": data["be_f64"]}, layout="chunked", chunk=(7, 3), filters=("deflate",))
    assert np.array_equal(c_read(h5, path)["a"][0], data["be_f64"].astype(np.float64))


def test_nested_groups_and_attributes(h5, tmp_path):
    rng = np.random.default_rng(5)
    data = {"top": rng.standard_normal(4), "g/inner": rng.integers(0, 9, 6).astype(np.int32), "g/h/deep": rng.standard_normal((2, 2))}
    attrs = {"top": {"size": np.int32(4), "dim": np.int32(1), "type": "double", "scale": 2.5}}
    for sb in (0, 2):
        path = str(tmp_path / f"nested{sb}.h5")
        h5_oracle.write_h5(path, data, superblock=sb, attrs=attrs)
        got = c_read(h5, path)
        got.pop("__superblock__")
        assert sorted(got) == sorted(data)
        for k, v in data.items():
            assert np.array_equal(got[k][0], v)
        err = C.create_string_buffer(256)
        r = h5.mgcfd_h5_open(path.encode(), err, 256)
        v = C.c_longlong()
        assert h5.mgcfd_h5_attr_int(r, b"top", b"size", v) == 0 and v.value == 4
        s = C.create_string_buffer(32)
        assert h5.mgcfd_h5_attr_str(r, b"top", b"type", s, 32) == 0 and s.value == b"double"
        assert h5.mgcfd_h5_attr_int(r, b"top", b"missing", v) == -1
        h5.mgcfd_h5_close(r)


def test_many_datasets_span_several_symbol_nodes(h5, tmp_path):
    data = {f"dataset_{i:03d}": np.full((3,), i, dtype=np.int32) for i in range(41)}
    path = str(tmp_path / "many.h5")
    h5_oracle.write_h5(path, data)
    got = c_read(h5, path)
    got.pop("__superblock__")
    assert sorted(got) == sorted(data)
    for k, v in data.items():
        assert np.array_equal(got[k][0], v)


def c_write(lib, path, data):
    codes = {np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}
    w = lib.mgcfd_h5_create(path.encode())
    keep = []
    for name, arr in data.items():
        a = np.ascontiguousarray(arr)
        keep.append(a)
        dims = (C.c_ulonglong * 8)(*a.shape)
        assert lib.mgcfd_h5_add(w, name.encode(), codes[a.dtype], a.ndim, dims, a.ctypes.data) == 0
    err = C.create_string_buffer(256)
    assert lib.mgcfd_h5_finish(w, err, 256) == 0, err.value.decode()


@pytest.mark.parametrize("count", [1, 7, 8, 9, 30])
def test_python_restatement_reads_what_cpp_writes(h5, tmp_path, count):
    rng = np.random.default_rng(11)
    data = deck_like(rng)
    data["p_variables_result_L0"] = rng.standard_normal((57, 5))
    for i in range(max(0, count - len(data))):
        data[f"extra{i}"] = rng.standard_normal((i + 1,)).astype(np.float32 if i % 2 else np.float64)
    data = dict(list(data.items())[:count])
    path = str(tmp_path / "cpp.h5")
    c_write(h5, path, data)
    back = h5_oracle.read_h5(path)
    assert sorted(back) == sorted(data)
    for name, ref in data.items():
        assert back[name]["data"].dtype == ref.dtype and np.array_equal(back[name]["data"], ref), name
        # OP2's op_decl_*_hdf5 conventions: "size", "dim", "type" attributes
        assert int(back[name]["attrs"]["size"][0]) == ref.shape[0]
        assert int(back[name]["attrs"]["dim"][0]) == (ref.shape[1] if ref.ndim > 1 else 1)
        assert back[name]["attrs"]["type"] == {"int32": "int", "float64": "double", "float32": "float", "int64": "long"}[ref.dtype.name]
    # ... and the C++ reader reads its own files
    got = c_read(h5, path)
    for name, ref in data.items():
        assert np.array_equal(got[name][0].astype(np.float64), ref.astype(np.float64))


def test_cpp_file_structure_follows_the_specification(h5, tmp_path):
    """byte-level checks of a written file: signature, version-0 superblock fields, end-of-file address, sorted symbol
    table, 8-byte aligned raw data"""
    rng = np.random.default_rng(2)
    data = {"zeta": rng.standard_normal((5, 3)), "alpha": rng.integers(0, 5, (4, 2)).astype(np.int32), "mid": rng.standard_normal(3)}
    path = str(tmp_path / "s.h5")
    c_write(h5, path, data)
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    import struct
    leaf_k, internal_k = struct.unpack_from("<HH", b, 16)
    assert (leaf_k, internal_k) == (4, 16)
    base, free, eof, driver = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == 2**64 - 1 and driver == 2**64 - 1 and eof == len(b)
    name_off, root, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", b, 56)
    assert cache == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP" and b[root] == 1
    snod = struct.unpack_from("<Q", b, btree + 24 + 8)[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 3
    seg = struct.unpack_from("<Q", b, heap + 24)[0]
    names = []
    for i in range(3):
        off = struct.unpack_from("<Q", b, snod + 8 + 40 * i)[0]
        names.append(b[seg + off:b.index(b"\0", seg + off)].decode())
    assert names == sorted(data)                                   # symbol table entries in strcmp order


def test_errors_are_reported_not_misread(h5, tmp_path):
    err = C.create_string_buffer(512)
    p = tmp_path / "not.h5"
    p.write_bytes(b"MGCFDBIN" + b"\0" * 100)
    assert h5.mgcfd_h5_is_hdf5(str(p).encode()) == 0
    assert not h5.mgcfd_h5_open(str(p).encode(), err, 512) and b"not an HDF5 file" in err.value
    assert not h5.mgcfd_h5_open(str(tmp_path / "missing.h5").encode(), err, 512) and b"cannot open" in err.value
    # truncated file: the raw data of the last dataset is cut off
    rng = np.random.default_rng(1)
    good = tmp_path / "good.h5"
    h5_oracle.write_h5(str(good), {"a": rng.standard_normal(1000)})
    blob = good.read_bytes()
    # the dataset's raw data sits right after the superblock placeholder: keep the metadata, cut the file short
    cut = tmp_path / "cut.h5"
    cut.write_bytes(blob[:len(blob) // 2])
    r = h5.mgcfd_h5_open(str(cut).encode(), err, 512)
    if r:
        a = np.empty(1000)
        assert h5.mgcfd_h5_read_f64(r, b"a", a.ctypes.data, err, 512) != 0
        h5.mgcfd_h5_close(r)
    assert b"beyond the end" in err.value or b"truncated" in err.value or b"bad" in err.value
    # dense link storage is refused by name
    dense = tmp_path / "dense.h5"
    h5_oracle.write_h5(str(dense), {"a": rng.standard_normal(3)}, superblock=2)
    raw = bytearray(dense.read_bytes())
    i = raw.rindex(b"\x02\x12\x00\x00")          # link info message header of the root group: type 2, size 18
    raw[i + 6:i + 14] = (1234).to_bytes(8, "little")
    dense.write_bytes(bytes(raw))
    assert not h5.mgcfd_h5_open(str(dense).encode(), err, 512) and b"dense link storage" in err.value


def test_h5_library_exports_every_declared_symbol():
    text = open(os.path.join(ROOT, "include", "mgcfd_h5.h")).read()
    declared = sorted(set(re.findall(r"\b(mgcfd_h5_[a-z0-9_]+)\s*\(", text)))
    lib = C.CDLL(LIB)
    assert len(declared) >= 14
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_written_messages_use_the_standard_encodings(h5, tmp_path):
    """known-answer bytes: the datatype messages libhdf5 writes for native little-endian double and int (as seen in any
    h5dump -H / hexdump of a file written on x86-64), the version-1 dataspace and the version-3 contiguous layout"""
    a = np.arange(6.0).reshape(2, 3)
    b = np.arange(4, dtype=np.int32)
    path = str(tmp_path / "kat.h5")
    c_write(h5, path, {"a": a, "b": b})
    blob = open(path, "rb").read()
    f64 = bytes.fromhex("11203f00" "08000000" "0000" "4000" "340b0034" "ff030000")     # IEEE double, LE: sign 63, exp 52/11, mantissa 0/52, bias 1023
    i32 = bytes.fromhex("10080000" "04000000" "0000" "2000")                            # fixed-point, LE, signed, 32 bits
    assert blob.count(f64) == 1 and blob.count(i32) >= 1
    space = bytes.fromhex("01020000" "00000000") + (2).to_bytes(8, "little") + (3).to_bytes(8, "little")
    assert space in blob
    at = blob.index(a.tobytes())
    assert at % 8 == 0
    layout = bytes([3, 1]) + at.to_bytes(8, "little") + (48).to_bytes(8, "little")
    assert layout in blob


def test_corrupted_files_never_crash_the_reader(h5, tmp_path):
    """fuzz: single-byte corruptions and truncations of valid files either still open or fail with an error text"""
    rng = np.random.default_rng(9)
    data = {"a": rng.standard_normal((6, 3)), "b": rng.integers(0, 9, (5, 2)).astype(np.int32)}
    err = C.create_string_buffer(512)
    for variant in (dict(), dict(superblock=2, layout="chunked", filters=("shuffle", "deflate"))):
        good = tmp_path / "good.h5"
        h5_oracle.write_h5(str(good), data, **variant)
        blob = good.read_bytes()
        bad = tmp_path / "bad.h5"
        for trial in range(150):
            b = bytearray(blob)
            if trial % 3 == 0:
                b = b[:int(rng.integers(8, len(b)))]
            else:
                for _ in range(int(rng.integers(1, 4))):
                    b[int(rng.integers(8, len(b)))] = int(rng.integers(0, 256))
            bad.write_bytes(bytes(b))
            r = h5.mgcfd_h5_open(str(bad).encode(), err, 512)
            if not r:
                assert err.value                                     # an error text, not a crash
                continue
            for i in range(h5.mgcfd_h5_count(r)):
                name = h5.mgcfd_h5_name(r, i)
                dims = (C.c_ulonglong * 8)()
                cls, es, sg, lay, rank = (C.c_int() for _ in range(5))
                h5.mgcfd_h5_info(r, name, cls, es, sg, lay, rank, dims)
                n = 1
                for k in range(rank.value):
                    n *= dims[k]
                if n > 10**6 or cls.value not in (0, 1):
                    continue                                         # a corrupted extent: do not allocate for it
                out = np.empty(max(n, 1), dtype=np.float64)
                h5.mgcfd_h5_read_f64(r, name, out.ctypes.data, err, 512)
            h5.mgcfd_h5_close(r)


GENUINE = os.path.join(ROOT, "tests", "golden", "libhdf5_matlab73_testdouble.mat")


def test_reads_a_file_written_by_the_real_hdf5_library(h5):
    """The one genuine libhdf5 product in the image: scipy's test file `testhdf5_7.4_GLNX86.mat` (scipy/io/matlab/tests/data,
    BSD-3-Clause, committed unchanged as tests/golden/libhdf5_matlab73_testdouble.mat).  MATLAB 7.4 wrote it in 2008 with its
    bundled HDF5 1.6-era library ("MAT-file version 7.3"): a 512-byte user block, superblock 0, an old-style root group
    (B-tree + local heap + symbol node), a version-1 object header, a version-1 dataspace, an IEEE double datatype, a
    version-1/2 data-layout message (contiguous) and a string attribute.  Known content (scipy's `testdouble` case):
    0 : pi/4 : 2*pi as a 9 x 1 dataset.  Both the C++ reader and the Python restatement must return it bit for bit."""
    assert h5.mgcfd_h5_is_hdf5(GENUINE.encode()) == 1
    raw = open(GENUINE, "rb").read()
    assert raw.startswith(b"MATLAB 7.0 MAT-file") and raw[512:520] == h5_oracle.SIG and raw[:8] != h5_oracle.SIG
    got = c_read(h5, GENUINE)
    assert got.pop("__superblock__") == 0
    assert list(got) == ["testdouble"]
    a, info = got["testdouble"]
    assert a.shape == (9, 1) and info == {"class": 1, "bytes": 8, "layout": 1}
    ref = h5_oracle.read_h5(GENUINE)
    assert list(ref) == ["testdouble"] and ref["testdouble"]["data"].dtype == np.dtype("<f8")
    assert np.array_equal(a, ref["testdouble"]["data"])                          # the two readers agree bit for bit ...
    # ... and the bytes are MATLAB's 0:pi/4:2*pi stored in the file, read here without any HDF5 code at all: the
    # dataset is contiguous, so its 72 bytes sit somewhere in the file exactly as IEEE little-endian doubles
    expect = np.pi / 4 * np.arange(9)
    assert np.allclose(a.ravel(), expect, rtol=0, atol=1e-15)
    assert a.tobytes() in raw
    assert ref["testdouble"]["attrs"] == {"MATLAB_class": "double"}
    buf = C.create_string_buffer(64)
    err = C.create_string_buffer(256)
    r = h5.mgcfd_h5_open(GENUINE.encode(), err, 256)
    try:
        assert h5.mgcfd_h5_attr_str(r, b"testdouble", b"MATLAB_class", buf, 64) == 0 and buf.value == b"double"
    finally:
        h5.mgcfd_h5_close(r)


def test_writer_uses_the_encodings_the_real_library_wrote(h5, tmp_path):
    """the same 9 x 1 dataset written by the C++ writer: the fixed part of the superblock (format versions, sizes of
    offsets / lengths, group leaf / internal node K) and the root symbol-table entry (cached B-tree + heap addresses,
    cache type 1) are encoded exactly as the genuine libhdf5 file encodes them; the dataset's object header holds the
    same kinds of messages (dataspace, datatype, layout, attribute) and the same data bytes"""
    lib_raw = open(GENUINE, "rb").read()[512:]
    a = (np.pi / 4 * np.arange(9)).reshape(9, 1)
    path = str(tmp_path / "ours.h5")
    w = h5.mgcfd_h5_create(path.encode())
    dims = (C.c_ulonglong * 2)(9, 1)
    assert h5.mgcfd_h5_add(w, b"testdouble", 3, 2, dims, a.ctypes.data) == 0
    err = C.create_string_buffer(256)
    assert h5.mgcfd_h5_finish(w, err, 256) == 0, err.value
    ours = open(path, "rb").read()
    assert ours[:20] == lib_raw[:20]                      # signature, versions 0/0/0/0, 8-byte offsets and lengths, K = 4 / 16
    # root symbol-table entry at byte 56: name offset 0, header address, cache type 1, reserved, scratch = B-tree + heap
    for f in (ours, lib_raw):
        name_off, hdr, cache, _res, btree, heap = np.frombuffer(f[56:56 + 40], dtype="<u8, <u8, <u4, <u4, <u8, <u8")[0]
        assert name_off == 0 and cache == 1 and hdr > 0 and btree > 0 and heap > 0
    # the dataset headers, message types in both files (ours adds OP2's attributes and a fill-value message at most)
    class R(h5_oracle._Reader):
        def __init__(self, b):
            self.b, self.base = b, 0
    def dataset_message_types(b, root_hdr):
        r = R(b)
        (_, addr), = r.children(r.messages(root_hdr))
        return {t for t, _ in r.messages(addr)}
    t_ours = dataset_message_types(ours, int(np.frombuffer(ours[64:72], dtype="<u8")[0]))
    t_lib = dataset_message_types(lib_raw, int(np.frombuffer(lib_raw[64:72], dtype="<u8")[0]))
    assert {0x01, 0x03, 0x08, 0x0C} <= t_ours and {0x01, 0x03, 0x08, 0x0C} <= t_lib
    assert a.tobytes() in ours and a.tobytes() in lib_raw[:]
