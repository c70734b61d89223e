"""Deterministic synthetic multigrid meshes in the reference's level-file layout.

The reference reads one HDF5 file per multigrid level (euler3d.cpp:248-312) holding the
datasets named below; its real decks (README.md:97-103) are not available offline, so the
BASELINE.json configs are realised as seeded synthetic meshes with the same dataset names,
shapes, dtypes and 1-based maps:

    node_coordinates   float64 [N,3]
    edge-->node        int32   [E,2]   (base_array_index-based, default 1: euler3d.cpp:92)
    edge_weights       float64 [E,3]
    bnd_node-->node    int32   [B,1]
    bnd_node-->group   int32   [B,1]
    bnd_node_weights   float64 [B,3]
    node-->mg_node     int32   [N,1]   map into the next coarser level (absent on the coarsest)

Each level is a perturbed structured grid relabelled as an unstructured mesh: axis edges
plus seeded diagonal edges, all six faces as boundary nodes (groups chosen so that every
branch of compute_bnd_node_flux_kernel, flux.h:29-37, runs), and a seeded random file
order for nodes and edges so that the planner's renumbering has real work to do.
"""
from __future__ import annotations

import numpy as np

DATASETS = (
    "node_coordinates", "edge-->node", "edge_weights", "bnd_node-->node",
    "bnd_node-->group", "bnd_node_weights", "node-->mg_node",
)

# name -> (mesh_name for input.dat, [(nx, ny, nz, total_edges or None)], base seed)
CONFIGS = {
    # BASELINE.json configs[0], [1]: Onera-M6-shaped, 300K nodes / 930K edges, 4 levels
    "m6": ("m6wing", [(100, 60, 50, 930_000), (75, 55, 40, 643_000), (60, 50, 37, 488_000),
                      (54, 50, 30, 377_117)], 1),
    # BASELINE.json configs[2]: Rotor37-1M-shaped
    "rotor37_1m": ("rotor37", [(100, 100, 100, 3_200_000), (80, 80, 80, 1_638_400),
                               (63, 63, 63, 800_150), (50, 50, 50, 400_000)], 100),
    # BASELINE.json configs[3]: Rotor37-8M-shaped
    "rotor37_8m": ("rotor37", [(200, 200, 200, 25_600_000), (160, 160, 160, 13_107_200),
                               (126, 126, 126, 6_401_203), (100, 100, 100, 3_200_000)], 100),
    # BASELINE.json configs[4]: one eighth of the 600x500x500 single-level deck (a 75x500x500 slab),
    # the per-GPU share of the 8-way partition; axis edges only
    "rotor37_150m_slab": ("rotor37", [(75, 500, 500, None)], 100),
    # small decks for CPU tests and goldens
    "tiny": ("m6wing", [(9, 7, 6, 1_000), (7, 6, 5, 560), (5, 4, 4, 200)], 7),
    "small": ("m6wing", [(24, 18, 14, 18_000), (18, 14, 11, 8_600), (13, 11, 8, 3_500),
                         (10, 8, 6, 1_400)], 11),
    "medium": ("rotor37", [(48, 40, 36, 215_000), (36, 30, 27, 92_000), (27, 23, 20, 40_000)], 21),
}

# face -> boundary group: far-field in/out on x, walls on y and z-min,
# a no-op group (>7) on z-max
FACE_GROUPS = {"xmin": 3, "xmax": 5, "ymin": 0, "ymax": 1, "zmin": 2, "zmax": 9}


def _axis_edges(nx, ny, nz):
    idx = np.arange(nx * ny * nz, dtype=np.int64).reshape(nx, ny, nz)
    parts = [
        (idx[:-1, :, :].ravel(), idx[1:, :, :].ravel()),
        (idx[:, :-1, :].ravel(), idx[:, 1:, :].ravel()),
        (idx[:, :, :-1].ravel(), idx[:, :, 1:].ravel()),
    ]
    a = np.concatenate([p[0] for p in parts])
    b = np.concatenate([p[1] for p in parts])
    axis = np.concatenate([np.full(p[0].size, k, dtype=np.int8) for k, p in enumerate(parts)])
    return a, b, axis


_DIAG_DIRS = np.array([(1, 1, 0), (1, 0, 1), (0, 1, 1), (1, -1, 0), (1, 0, -1), (0, 1, -1)], dtype=np.int64)


def _diagonal_edges(nx, ny, nz, count, rng):
    """`count` distinct face-diagonal edges, drawn without replacement from all valid ones."""
    if count <= 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    n = nx * ny * nz
    i, j, k = np.unravel_index(np.arange(n, dtype=np.int64), (nx, ny, nz))
    cand_a, cand_b = [], []
    for d in _DIAG_DIRS:
        i2, j2, k2 = i + d[0], j + d[1], k + d[2]
        ok = (i2 >= 0) & (i2 < nx) & (j2 >= 0) & (j2 < ny) & (k2 >= 0) & (k2 < nz)
        cand_a.append(np.nonzero(ok)[0])
        cand_b.append((i2[ok] * ny + j2[ok]) * nz + k2[ok])
    cand_a = np.concatenate(cand_a)
    cand_b = np.concatenate(cand_b)
    if count > cand_a.size:
        raise ValueError(f"asked for {count} diagonal edges, only {cand_a.size} exist")
    pick = rng.choice(cand_a.size, size=count, replace=False)
    pick.sort()
    return cand_a[pick], cand_b[pick]


def make_level(nx, ny, nz, total_edges, seed, coarse_dims=None, base=1, shuffle=True, extent=None):
    """One level as a dict of the reference's datasets (maps `base`-based, default 1)."""
    rng = np.random.default_rng(seed)
    n = nx * ny * nz
    ext = extent if extent is not None else (1.0, 0.6, 0.5)
    h = np.array([ext[0] / max(nx - 1, 1), ext[1] / max(ny - 1, 1), ext[2] / max(nz - 1, 1)])
    hm = float(h.mean())

    # --- nodes: perturbed grid
    i, j, k = np.unravel_index(np.arange(n, dtype=np.int64), (nx, ny, nz))
    coords = np.stack([i * h[0], j * h[1], k * h[2]], axis=1)
    coords += rng.uniform(-0.2, 0.2, size=(n, 3)) * h

    # --- edges: all axis edges + seeded diagonals, random orientation
    ea, eb, axis = _axis_edges(nx, ny, nz)
    n_axis = ea.size
    n_diag = 0 if total_edges is None else total_edges - n_axis
    if n_diag < 0:
        raise ValueError(f"total_edges={total_edges} is below the {n_axis} axis edges")
    da, db = _diagonal_edges(nx, ny, nz, n_diag, rng)
    ea = np.concatenate([ea, da])
    eb = np.concatenate([eb, db])
    ne = ea.size
    flip = rng.random(ne) < 0.5
    ea, eb = np.where(flip, eb, ea), np.where(flip, ea, eb)
    # face-normal-like weights of magnitude ~h^2 (direction is re-aimed at init: misc.h:65-75)
    mag = hm * hm * rng.uniform(0.9, 1.1, size=ne)
    mag[n_axis:] *= 0.35
    w = np.zeros((ne, 3))
    ax_full = np.concatenate([axis.astype(np.int64), rng.integers(0, 3, size=n_diag)])
    w[np.arange(ne), ax_full] = mag

    # --- boundary nodes: six faces (edge/corner nodes appear once per face they lie on)
    faces = [
        ("xmin", i == 0, (-1, 0, 0), h[1] * h[2]), ("xmax", i == nx - 1, (1, 0, 0), h[1] * h[2]),
        ("ymin", j == 0, (0, -1, 0), h[0] * h[2]), ("ymax", j == ny - 1, (0, 1, 0), h[0] * h[2]),
        ("zmin", k == 0, (0, 0, -1), h[0] * h[1]), ("zmax", k == nz - 1, (0, 0, 1), h[0] * h[1]),
    ]
    bn, bg, bw = [], [], []
    for name, mask, normal, area in faces:
        ids = np.nonzero(mask)[0]
        bn.append(ids)
        bg.append(np.full(ids.size, FACE_GROUPS[name], dtype=np.int32))
        bw.append(np.outer(area * rng.uniform(0.9, 1.1, size=ids.size), np.array(normal, dtype=float)))
    bn = np.concatenate(bn)
    bg = np.concatenate(bg)
    bw = np.concatenate(bw)

    # --- multigrid map: nearest coarse grid index per axis, then a seeded 2 % of the coarse
    #     nodes hand their children to the +x neighbour so that childless coarse nodes exist
    mg = None
    if coarse_dims is not None:
        cx, cy, cz = coarse_dims
        ci = np.rint(i * ((cx - 1) / max(nx - 1, 1))).astype(np.int64)
        cj = np.rint(j * ((cy - 1) / max(ny - 1, 1))).astype(np.int64)
        ck = np.rint(k * ((cz - 1) / max(nz - 1, 1))).astype(np.int64)
        hole = rng.random(cx * cy * cz) < 0.02
        hole.reshape(cx, cy, cz)[cx - 1, :, :] = False
        mg = (ci * cy + cj) * cz + ck
        mg = np.where(hole[mg], mg + cy * cz, mg)

    # --- relabel: seeded random file order of nodes, edges and boundary entries
    if shuffle:
        perm = rng.permutation(n)            # new label of old node
        inv = np.empty(n, dtype=np.int64)
        inv[perm] = np.arange(n)
        coords = coords[inv]
        ea, eb = perm[ea], perm[eb]
        eorder = rng.permutation(ne)
        ea, eb, w = ea[eorder], eb[eorder], w[eorder]
        bn = perm[bn]
        border = rng.permutation(bn.size)
        bn, bg, bw = bn[border], bg[border], bw[border]
        if mg is not None:
            mg = mg[inv]                     # still indexes the coarse grid's *grid* labels
    else:
        perm = np.arange(n)

    level = {
        "node_coordinates": np.ascontiguousarray(coords, dtype=np.float64),
        "edge-->node": np.ascontiguousarray(np.stack([ea, eb], axis=1) + base, dtype=np.int32),
        "edge_weights": np.ascontiguousarray(w, dtype=np.float64),
        "bnd_node-->node": np.ascontiguousarray(bn[:, None] + base, dtype=np.int32),
        "bnd_node-->group": np.ascontiguousarray(bg[:, None], dtype=np.int32),
        "bnd_node_weights": np.ascontiguousarray(bw, dtype=np.float64),
    }
    return level, perm, mg


def make_multigrid(config, base=1, shuffle=True):
    """All levels of a named config.  Returns {"mesh_name", "base_array_index", "levels": [dict]}."""
    if isinstance(config, str):
        mesh_name, dims, seed0 = CONFIGS[config]
    else:
        mesh_name, dims, seed0 = config
    levels, perms, mgs = [], [], []
    for l, (nx, ny, nz, ne) in enumerate(dims):
        coarse = dims[l + 1][:3] if l + 1 < len(dims) else None
        lev, perm, mg = make_level(nx, ny, nz, ne, seed0 + l, coarse_dims=coarse, base=base, shuffle=shuffle)
        levels.append(lev)
        perms.append(perm)
        mgs.append(mg)
    # the fine level's mg map must point at the coarse level's *file* labels
    for l in range(len(dims) - 1):
        levels[l]["node-->mg_node"] = np.ascontiguousarray(
            perms[l + 1][mgs[l]][:, None] + base, dtype=np.int32)
    return {"mesh_name": mesh_name, "base_array_index": base, "levels": levels,
            "dims": [tuple(d[:3]) for d in dims]}


def zero_based(level, base=1):
    """The 0-based maps the C-ABI and the oracle take (OP2 subtracts OP_MAPS_BASE_INDEX at load)."""
    out = dict(level)
    for key in ("edge-->node", "bnd_node-->node", "node-->mg_node"):
        if key in out:
            out[key] = np.ascontiguousarray(out[key] - base, dtype=np.int32)
    return out


# ---------------------------------------------------------------------------------------------------
# deck files.  The reference reads HDF5 level files (euler3d.cpp:248-312); no HDF5 library exists in this image
# (SURVEY.md 7.3-H1), so decks are stored in a minimal container holding the SAME dataset names, shapes and dtypes:
#   "MGCFDBIN" | u32 version=1 | u32 n_datasets | per dataset: u32 name_len, name, u32 dtype (0 float64, 1 int32),
#   u32 ndim, u64 dims[ndim], u64 nbytes, zero padding to an 8-byte boundary, raw little-endian data
# read by host/euler3d_b200.cpp and by read_container() below.
# ---------------------------------------------------------------------------------------------------
import os
import struct

_MAGIC = b"MGCFDBIN"


def write_container(path, datasets):
    with open(path, "wb") as f:
        f.write(_MAGIC + struct.pack("<II", 1, len(datasets)))
        for name, arr in datasets.items():
            a = np.ascontiguousarray(arr)
            if a.dtype == np.float64:
                code = 0
            elif a.dtype == np.int32:
                code = 1
            else:
                raise TypeError(f"{name}: only float64 / int32 datasets exist in MG-CFD decks")
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)) + nb + struct.pack("<II", code, a.ndim))
            f.write(struct.pack(f"<{a.ndim}Q", *a.shape) + struct.pack("<Q", a.nbytes))
            f.write(b"\0" * (-f.tell() % 8))
            f.write(a.tobytes())


def read_container(path):
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != _MAGIC:
            raise ValueError(f"{path}: not an MGCFDBIN container")
        _, n = struct.unpack("<II", f.read(8))
        for _ in range(n):
            (ln,) = struct.unpack("<I", f.read(4))
            name = f.read(ln).decode()
            code, ndim = struct.unpack("<II", f.read(8))
            dims = struct.unpack(f"<{ndim}Q", f.read(8 * ndim))
            (nbytes,) = struct.unpack("<Q", f.read(8))
            f.read(-f.tell() % 8)
            out[name] = np.frombuffer(f.read(nbytes), dtype=np.float64 if code == 0 else np.int32).reshape(dims).copy()
    return out


# ---- HDF5 level / solution files through the library's own from-scratch HDF5 subset (include/mgcfd_h5.h; the image
# has no libhdf5 / h5py).  Datasets carry OP2's "size" / "dim" / "type" attributes (op_decl_*_hdf5 conventions).
_H5 = None


def _h5lib():
    global _H5
    if _H5 is None:
        import ctypes as C
        lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmgcfd_h5.so"))
        lib.mgcfd_h5_open.restype = C.c_void_p
        lib.mgcfd_h5_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        lib.mgcfd_h5_close.argtypes = [C.c_void_p]
        lib.mgcfd_h5_count.argtypes = [C.c_void_p]
        lib.mgcfd_h5_name.restype = C.c_char_p
        lib.mgcfd_h5_name.argtypes = [C.c_void_p, C.c_int]
        lib.mgcfd_h5_info.argtypes = [C.c_void_p, C.c_char_p] + [C.POINTER(C.c_int)] * 5 + [C.POINTER(C.c_ulonglong)]
        lib.mgcfd_h5_read_f64.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int]
        lib.mgcfd_h5_read_i32.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int]
        lib.mgcfd_h5_create.restype = C.c_void_p
        lib.mgcfd_h5_create.argtypes = [C.c_char_p]
        lib.mgcfd_h5_add.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.c_void_p]
        lib.mgcfd_h5_finish.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _H5 = lib
    return _H5


def write_h5(path, datasets):
    import ctypes as C
    lib = _h5lib()
    w = lib.mgcfd_h5_create(path.encode())
    keep = []
    for name, arr in datasets.items():
        a = np.ascontiguousarray(arr)
        if a.dtype not in (np.float64, np.int32):
            raise TypeError(f"{name}: only float64 / int32 datasets exist in MG-CFD decks")
        keep.append(a)
        dims = (C.c_ulonglong * 8)(*a.shape)
        if lib.mgcfd_h5_add(w, name.encode(), 3 if a.dtype == np.float64 else 0, a.ndim, dims, a.ctypes.data) != 0:
            raise ValueError(f"{path}: cannot add dataset {name}")
    err = C.create_string_buffer(512)
    if lib.mgcfd_h5_finish(w, err, 512) != 0:
        raise OSError(err.value.decode())


def read_h5(path):
    import ctypes as C
    lib = _h5lib()
    err = C.create_string_buffer(512)
    r = lib.mgcfd_h5_open(path.encode(), err, 512)
    if not r:
        raise ValueError(err.value.decode())
    out = {}
    try:
        for i in range(lib.mgcfd_h5_count(r)):
            name = lib.mgcfd_h5_name(r, i)
            cls, es, sg, lay, rank = (C.c_int() for _ in range(5))
            dims = (C.c_ulonglong * 8)()
            lib.mgcfd_h5_info(r, name, cls, es, sg, lay, rank, dims)
            shape = tuple(dims[k] for k in range(rank.value))
            a = np.empty(shape, dtype=np.float64 if cls.value == 1 else np.int32)
            read = lib.mgcfd_h5_read_f64 if cls.value == 1 else lib.mgcfd_h5_read_i32
            if read(r, name, a.ctypes.data, err, 512) != 0:
                raise ValueError(err.value.decode())
            out[name.decode()] = a
    finally:
        lib.mgcfd_h5_close(r)
    return out


def write_deck(directory, mesh, stem="mesh", fmt="mgb"):
    """input.dat (io.h:28-205 format) + one level file per multigrid level (fmt "mgb": container, "h5": HDF5)"""
    os.makedirs(directory, exist_ok=True)
    names = []
    for l, lev in enumerate(mesh["levels"]):
        names.append(f"{stem}.L{l}.{fmt}")
        (write_h5 if fmt == "h5" else write_container)(os.path.join(directory, names[-1]), lev)
    with open(os.path.join(directory, "input.dat"), "w") as f:
        f.write("# synthetic MG-CFD deck (meshgen.py)\n")
        f.write(f"size = {mesh['levels'][0]['node_coordinates'].shape[0]}\n")
        f.write(f"num_levels = {len(names)}\n")
        f.write(f"base_array_index = {mesh['base_array_index']}\n")
        f.write(f"mesh_name = {mesh['mesh_name']}\n")
        f.write("[levels]\n")
        for l, n in enumerate(names):
            f.write(f"{l} = {n}\n")
    return os.path.join(directory, "input.dat")


def write_solution(directory, level, cycles, variables, prefix="solution.", fmt="mgb"):
    """solution.variables.L<l>.cycles=<g> with dataset p_variables_result_L<l> (euler3d.cpp:315-327, 765-771)"""
    path = os.path.join(directory, f"{prefix}variables.L{level}.cycles={cycles}.{fmt}")
    (write_h5 if fmt == "h5" else write_container)(path, {f"p_variables_result_L{level}": np.ascontiguousarray(variables, dtype=np.float64)})
    return path


# ---------------------------------------------------------------------------------------------------
# Slab decks generated PER RANK (BASELINE.json configs[4]: the 600x500x500 single-level deck on 8 GPUs is never
# materialised on one host).  Every random quantity is a hash of the GLOBAL node / edge / boundary-entry
# index, so a rank can generate its x-slab (+ one halo plane on either side) alone and gets exactly the arrays
# mgcfd_local_mesh_build would cut out of the whole deck (tests/test_meshgen.py compares the two).
#   nodes     natural order, id = (i*ny + j)*nz + k, perturbed grid coordinates
#   edges     axis edges only; global order: x-edges (id = i*P + jk, P = ny*nz), then y-edges, then z-edges; hashed orientation
#   boundary  six faces in the order xmin, xmax, ymin, ymax, zmin, zmax, each ascending in node id
#   ownership x-plane i belongs to rank r with  r*nx//R <= i < (r+1)*nx//R
# ---------------------------------------------------------------------------------------------------
SLAB_CONFIGS = {
    "rotor37_150m": ("rotor37", (600, 500, 500), 100),       # BASELINE.json configs[4]
    "slab_test": ("rotor37", (12, 5, 4), 5),
}


def _hash_u01(ids, stream, seed):
    """uniform [0,1) per index: splitmix64 of (index, stream, seed)"""
    with np.errstate(over="ignore"):
        z = (np.asarray(ids).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03) +
             np.uint64(seed) * np.uint64(0x8CB92BA72F3D8DD7))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _slab_geometry(nx, ny, nz, extent):
    ext = extent if extent is not None else (1.0, 0.6, 0.5)
    h = np.array([ext[0] / max(nx - 1, 1), ext[1] / max(ny - 1, 1), ext[2] / max(nz - 1, 1)])
    return h, float(h.mean())


def _slab_coords(gid, nx, ny, nz, seed, h):
    i, j, k = np.unravel_index(gid, (nx, ny, nz))
    c = np.stack([i * h[0], j * h[1], k * h[2]], axis=1).astype(np.float64)
    for d in range(3):
        c[:, d] += (_hash_u01(gid, 10 + d, seed) - 0.5) * 0.4 * h[d]
    return c


def _slab_edges(nx, ny, nz, i_lo, i_hi):
    """global ids, endpoints (global node ids) and axis of the axis edges whose lower x-plane index lies in the given
    ranges: x-edges with i in [i_lo[0], i_hi[0]), y- and z-edges of planes [i_lo[1], i_hi[1])"""
    P = ny * nz
    Ex, Ey = (nx - 1) * P, nx * (ny - 1) * nz
    out = []
    # x
    i = np.arange(max(i_lo[0], 0), min(i_hi[0], nx - 1), dtype=np.int64)
    a = (i[:, None] * P + np.arange(P, dtype=np.int64)[None, :]).ravel()
    out.append((a, a, a + P, 0))
    # y
    i = np.arange(i_lo[1], i_hi[1], dtype=np.int64)
    jj, kk = np.meshgrid(np.arange(ny - 1, dtype=np.int64), np.arange(nz, dtype=np.int64), indexing="ij")
    eid = Ex + (i[:, None] * (ny - 1) * nz + (jj * nz + kk).ravel()[None, :]).ravel()
    a = (i[:, None] * P + (jj * nz + kk).ravel()[None, :]).ravel()
    out.append((eid, a, a + nz, 1))
    # z
    jj, kk = np.meshgrid(np.arange(ny, dtype=np.int64), np.arange(nz - 1, dtype=np.int64), indexing="ij")
    eid = Ex + Ey + (i[:, None] * ny * (nz - 1) + (jj * (nz - 1) + kk).ravel()[None, :]).ravel()
    a = (i[:, None] * P + (jj * nz + kk).ravel()[None, :]).ravel()
    out.append((eid, a, a + 1, 2))
    eid = np.concatenate([o[0] for o in out])
    ea = np.concatenate([o[1] for o in out])
    eb = np.concatenate([o[2] for o in out])
    axis = np.concatenate([np.full(o[0].size, o[3], dtype=np.int64) for o in out])
    return eid, ea, eb, axis


def _slab_edge_data(eid, ea, eb, axis, seed, hm):
    flip = _hash_u01(eid, 1, seed) < 0.5
    a, b = np.where(flip, eb, ea), np.where(flip, ea, eb)
    w = np.zeros((eid.size, 3))
    w[np.arange(eid.size), axis] = hm * hm * (0.9 + 0.2 * _hash_u01(eid, 2, seed))
    return a, b, w


def _slab_boundary(nx, ny, nz, x0, x1, seed, h):
    """boundary entries of the nodes in planes [x0, x1): (global entry index, global node id, group, weight)"""
    P = ny * nz
    i = np.arange(x0, x1, dtype=np.int64)
    faces = []
    jk = np.arange(P, dtype=np.int64)
    if x0 == 0:
        faces.append((jk, jk, "xmin", (-1, 0, 0), h[1] * h[2]))
    if x1 == nx:
        faces.append((P + jk, (nx - 1) * P + jk, "xmax", (1, 0, 0), h[1] * h[2]))
    k = np.arange(nz, dtype=np.int64)
    j = np.arange(ny, dtype=np.int64)
    base = 2 * P
    faces.append((base + (i[:, None] * nz + k[None, :]).ravel(), (i[:, None] * P + k[None, :]).ravel(), "ymin", (0, -1, 0), h[0] * h[2]))
    base += nx * nz
    faces.append((base + (i[:, None] * nz + k[None, :]).ravel(), (i[:, None] * P + (ny - 1) * nz + k[None, :]).ravel(), "ymax", (0, 1, 0), h[0] * h[2]))
    base += nx * nz
    faces.append((base + (i[:, None] * ny + j[None, :]).ravel(), (i[:, None] * P + j[None, :] * nz).ravel(), "zmin", (0, 0, -1), h[0] * h[1]))
    base += nx * ny
    faces.append((base + (i[:, None] * ny + j[None, :]).ravel(), (i[:, None] * P + j[None, :] * nz + nz - 1).ravel(), "zmax", (0, 0, 1), h[0] * h[1]))
    bid = np.concatenate([f[0] for f in faces])
    bn = np.concatenate([f[1] for f in faces])
    bg = np.concatenate([np.full(f[0].size, FACE_GROUPS[f[2]], dtype=np.int32) for f in faces])
    bw = np.concatenate([np.outer(f[4] * (0.9 + 0.2 * _hash_u01(f[0], 3, seed)), np.array(f[3], dtype=float)) for f in faces])
    return bid, bn, bg, bw


def slab_sizes(config):
    """(nodes, edges, boundary entries) of the whole deck, without generating it"""
    _, (nx, ny, nz), _ = SLAB_CONFIGS[config] if isinstance(config, str) else config
    return nx * ny * nz, (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1), 2 * (ny * nz + nx * nz + nx * ny)


def slab_owner_planes(nx, rank, n_ranks):
    return rank * nx // n_ranks, (rank + 1) * nx // n_ranks


def make_slab_global(config, base=1, extent=None):
    """the whole single-level deck (small sizes: tests, one GPU), same dict layout as make_multigrid"""
    mesh_name, (nx, ny, nz), seed = SLAB_CONFIGS[config] if isinstance(config, str) else config
    h, hm = _slab_geometry(nx, ny, nz, extent)
    n = nx * ny * nz
    coords = _slab_coords(np.arange(n, dtype=np.int64), nx, ny, nz, seed, h)
    eid, ea, eb, axis = _slab_edges(nx, ny, nz, (0, 0), (nx - 1, nx))
    a, b, w = _slab_edge_data(eid, ea, eb, axis, seed, hm)
    bid, bn, bg, bw = _slab_boundary(nx, ny, nz, 0, nx, seed, h)
    assert np.array_equal(eid, np.arange(eid.size)) and np.array_equal(bid, np.arange(bid.size))
    level = {
        "node_coordinates": coords,
        "edge-->node": np.ascontiguousarray(np.stack([a, b], axis=1) + base, dtype=np.int32),
        "edge_weights": w,
        "bnd_node-->node": np.ascontiguousarray(bn[:, None] + base, dtype=np.int32),
        "bnd_node-->group": np.ascontiguousarray(bg[:, None], dtype=np.int32),
        "bnd_node_weights": bw,
    }
    return {"mesh_name": mesh_name, "base_array_index": base, "levels": [level], "dims": [(nx, ny, nz)]}


def make_slab_rank(config, rank, n_ranks, extent=None):
    """this rank's share of the deck, generated without the rest: the arrays of mgcfd_level_host for a partition
    (local 0-based maps, [owned | lower halo plane | upper halo plane] nodes, edges with an owned endpoint in ascending
    global edge index, boundary entries of owned nodes, neighbour / export / import lists)"""
    mesh_name, (nx, ny, nz), seed = SLAB_CONFIGS[config] if isinstance(config, str) else config
    h, hm = _slab_geometry(nx, ny, nz, extent)
    P = ny * nz
    x0, x1 = slab_owner_planes(nx, rank, n_ranks)
    assert x1 > x0, "more ranks than x-planes"
    owned = np.arange(x0 * P, x1 * P, dtype=np.int64)
    lower = np.arange((x0 - 1) * P, x0 * P, dtype=np.int64) if x0 > 0 else np.zeros(0, np.int64)
    upper = np.arange(x1 * P, (x1 + 1) * P, dtype=np.int64) if x1 < nx else np.zeros(0, np.int64)
    gnode = np.concatenate([owned, lower, upper])
    n_owned = owned.size

    def local_of(g):
        loc = g - x0 * P                                        # owned
        loc = np.where(g < x0 * P, n_owned + (g - (x0 - 1) * P), loc)
        loc = np.where(g >= x1 * P, n_owned + lower.size + (g - x1 * P), loc)
        return loc

    eid, ea, eb, axis = _slab_edges(nx, ny, nz, (x0 - 1, x0), (x1, x1))
    a, b, w = _slab_edge_data(eid, ea, eb, axis, seed, hm)
    bid, bn, bg, bw = _slab_boundary(nx, ny, nz, x0, x1, seed, h)
    order = np.argsort(bid, kind="stable")                      # ascending global entry index (faces interleave per rank)
    bid, bn, bg, bw = bid[order], bn[order], bg[order], bw[order]
    nbr, exp_idx, exp_ptr, imp_ptr = [], [], [0], [0]
    if x0 > 0:
        nbr.append(rank - 1 if n_ranks > 1 else 0)
        exp_idx.append(np.arange(0, P, dtype=np.int64))                        # my first plane, the neighbour's upper halo
        exp_ptr.append(exp_ptr[-1] + P)
        imp_ptr.append(imp_ptr[-1] + lower.size)
    if x1 < nx:
        nbr.append(rank + 1)
        exp_idx.append(np.arange(n_owned - P, n_owned, dtype=np.int64))        # my last plane, the neighbour's lower halo
        exp_ptr.append(exp_ptr[-1] + P)
        imp_ptr.append(imp_ptr[-1] + upper.size)
    # ranks need not be adjacent in rank number when a rank owns no plane in between; planes are contiguous here
    if nbr:
        nbr[0] = _owner_of_plane(nx, n_ranks, x0 - 1) if x0 > 0 else nbr[0]
        if x1 < nx:
            nbr[-1] = _owner_of_plane(nx, n_ranks, x1)
    return {
        "mesh_name": mesh_name, "rank": rank, "n_ranks": n_ranks, "n_owned": n_owned,
        "node_coordinates": _slab_coords(gnode, nx, ny, nz, seed, h),
        "edge-->node": np.ascontiguousarray(np.stack([local_of(a), local_of(b)], axis=1), dtype=np.int32),
        "edge_weights": w,
        "bnd_node-->node": np.ascontiguousarray(local_of(bn)[:, None], dtype=np.int32),
        "bnd_node-->group": np.ascontiguousarray(bg[:, None], dtype=np.int32),
        "bnd_node_weights": bw,
        "global_node": gnode.astype(np.int32), "global_edge": eid.astype(np.int64), "global_bnd": bid.astype(np.int64),
        "neighbour_rank": np.array(nbr, dtype=np.int32), "export_ptr": np.array(exp_ptr, dtype=np.int32),
        "export_idx": (np.concatenate(exp_idx) if exp_idx else np.zeros(0, np.int64)).astype(np.int32),
        "import_ptr": np.array(imp_ptr, dtype=np.int32),
    }


def _owner_of_plane(nx, n_ranks, i):
    r = (i * n_ranks) // nx
    while slab_owner_planes(nx, r, n_ranks)[1] <= i:
        r += 1
    while slab_owner_planes(nx, r, n_ranks)[0] > i:
        r -= 1
    return r


def slab_part(config, n_ranks):
    """owner rank of every node of the whole deck (for tests against mgcfd_local_mesh_build)"""
    _, (nx, ny, nz), _ = SLAB_CONFIGS[config] if isinstance(config, str) else config
    plane_owner = np.array([_owner_of_plane(nx, n_ranks, i) for i in range(nx)], dtype=np.int32)
    return np.repeat(plane_owner, ny * nz)
