"""ctypes binding of include/mgcfd_b200.h (libmgcfd_b200.so) and a driver restating euler3d.cpp.

This is the Python face of the C-ABI for tests and bench.py; the native host driver is
host/euler3d_b200.cpp.  There is no fallback: if the CUDA library is missing or no device is
present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmgcfd_b200.so")

NVAR, NDIM, RK = 5, 3, 3
FLUX_ATOMIC, FLUX_COLOUR, FLUX_OWNER, FLUX_GATHER, FLUX_EMIT = 0, 1, 2, 3, 4
FLUX_VARIANTS = {"atomic": FLUX_ATOMIC, "colour": FLUX_COLOUR, "owner": FLUX_OWNER, "gather": FLUX_GATHER,
                 "emit": FLUX_EMIT}

ERR_NAMES = {0: "OK", -1: "ERR_ARG", -2: "ERR_CUDA", -3: "ERR_NODEVICE", -4: "ERR_MIN_DT",
             -5: "ERR_BAD_VALS", -6: "ERR_PLAN", -7: "ERR_COMM"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class MgcfdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Consts(C.Structure):
    _fields_ = [("smoothing_coefficient", C.c_double), ("ff_variable", C.c_double * 5),
                ("ff_flux_contribution_momentum_x", C.c_double * 3),
                ("ff_flux_contribution_momentum_y", C.c_double * 3),
                ("ff_flux_contribution_momentum_z", C.c_double * 3),
                ("ff_flux_contribution_density_energy", C.c_double * 3),
                ("mesh_name", C.c_int), ("pad_", C.c_int)]


class LevelHost(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("n_edges", C.c_int), ("n_bnd_nodes", C.c_int), ("n_owned_nodes", C.c_int),
                ("node_coordinates", _dp), ("edge_to_node", _ip), ("edge_weights", _dp),
                ("bnd_node_to_node", _ip), ("bnd_node_to_group", _ip), ("bnd_node_weights", _dp),
                ("node_to_mg_node", _ip), ("global_node_id", _ip), ("n_neighbours", C.c_int), ("pad_", C.c_int),
                ("neighbour_rank", _ip), ("export_ptr", _ip), ("export_idx", _ip), ("import_ptr", _ip)]


class Options(C.Structure):
    _fields_ = [("flux_variant", C.c_int), ("renumber", C.c_int), ("owner_chunk_nodes", C.c_int),
                ("colour_block_edges", C.c_int), ("exact_arith", C.c_int), ("no_fusion", C.c_int),
                ("rank", C.c_int), ("n_ranks", C.c_int), ("no_graphs", C.c_int), ("measure_mem_bound", C.c_int),
                ("reserved", C.c_int * 6)]


_lib = None


def load_library():
    """Load libmgcfd_b200.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MgcfdError(-3, f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    lib.mgcfd_last_error.restype = C.c_char_p
    lib.mgcfd_last_error.argtypes = [C.c_void_p]
    lib.mgcfd_version.restype = C.c_char_p
    lib.mgcfd_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(Options)]
    lib.mgcfd_destroy.argtypes = [C.c_void_p]
    lib.mgcfd_plan_query.restype = C.c_longlong
    lib.mgcfd_plan_query.argtypes = [C.c_void_p, C.c_int, C.c_char_p, _ip, C.c_longlong]
    lib.mgcfd_run_cycles_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.mgcfd_kernel_launches.restype = C.c_longlong
    lib.mgcfd_kernel_launches.argtypes = [C.c_void_p]
    lib.mgcfd_stream.restype = C.c_void_p
    lib.mgcfd_stream.argtypes = [C.c_void_p]
    lib.mgcfd_device_ptr.restype = C.c_void_p
    lib.mgcfd_device_ptr.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.mgcfd_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    lib.mgcfd_host_free.argtypes = [C.c_void_p]
    lib.mgcfd_local_mesh_level.restype = C.POINTER(LevelHost)
    lib.mgcfd_local_mesh_level.argtypes = [C.c_void_p, C.c_int]
    lib.mgcfd_local_mesh_query.restype = C.c_longlong
    lib.mgcfd_local_mesh_query.argtypes = [C.c_void_p, C.c_int, C.c_char_p, _ip, C.c_longlong]
    lib.mgcfd_local_mesh_free.argtypes = [C.c_void_p]
    lib.mgcfd_halo_bytes_sent.restype = C.c_longlong
    lib.mgcfd_halo_bytes_sent.argtypes = [C.c_void_p]
    _lib = lib
    return lib


# every symbol include/mgcfd_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "mgcfd_default_options", "mgcfd_create", "mgcfd_destroy", "mgcfd_last_error", "mgcfd_version",
    "mgcfd_compute_farfield_consts", "mgcfd_decl_consts", "mgcfd_decl_level", "mgcfd_plan",
    "mgcfd_loop_initialize_variables", "mgcfd_loop_zero_fluxes", "mgcfd_loop_zero_volumes",
    "mgcfd_loop_calculate_cell_volumes", "mgcfd_loop_dampen_ewt_edges", "mgcfd_loop_dampen_ewt_bnd",
    "mgcfd_loop_copy_double", "mgcfd_loop_calculate_dt", "mgcfd_loop_get_min_dt",
    "mgcfd_loop_compute_step_factor", "mgcfd_loop_compute_flux_edge", "mgcfd_loop_compute_bnd_node_flux",
    "mgcfd_loop_time_step", "mgcfd_loop_unstructured_stream", "mgcfd_loop_residual", "mgcfd_loop_calc_rms",
    "mgcfd_loop_count_bad_vals", "mgcfd_loop_up_pre", "mgcfd_loop_up", "mgcfd_loop_up_post", "mgcfd_loop_down",
    "mgcfd_run_cycles", "mgcfd_run_cycles_host", "mgcfd_fetch_dat", "mgcfd_set_dat", "mgcfd_sync", "mgcfd_validate_level",
    "mgcfd_plan_query", "mgcfd_timers_enable", "mgcfd_timers_reset", "mgcfd_timers_get",
    "mgcfd_kernel_launches", "mgcfd_set_flux_variant", "mgcfd_stream", "mgcfd_device_ptr",
    "mgcfd_host_alloc", "mgcfd_host_free", "mgcfd_partition_graph",
    "mgcfd_partition_rcb", "mgcfd_partition_coarse", "mgcfd_local_mesh_build", "mgcfd_local_mesh_level",
    "mgcfd_local_mesh_query", "mgcfd_local_mesh_free", "mgcfd_group_run_cycles", "mgcfd_nccl_unique_id",
    "mgcfd_comm_init_nccl", "mgcfd_halo_bytes_sent", "mgcfd_group_enable_p2p", "mgcfd_ipc_export", "mgcfd_comm_init_ipc",
]

_DAT_DIMS = {"variables": 5, "old_variables": 5, "residuals": 5, "fluxes": 5, "dummy_fluxes": 5,
             "volumes": 1, "step_factors": 1, "node_coordinates": 3}


def nccl_unique_id():
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.mgcfd_nccl_unique_id(buf)
    if rc != 0:
        raise MgcfdError(rc, "mgcfd_nccl_unique_id failed (libnccl.so.2 not loadable?)")
    return buf.raw


def farfield_consts(mesh_name=0):
    lib = load_library()
    c = Consts()
    lib.mgcfd_compute_farfield_consts(C.byref(c))
    c.mesh_name = mesh_name
    return c


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class PinnedArray:
    """float64 numpy view of a page-locked host buffer (mgcfd_host_alloc); fetch/set DMA it directly."""

    def __init__(self, shape):
        lib = load_library()
        self.ptr = C.c_void_p()
        n = int(np.prod(shape))
        rc = lib.mgcfd_host_alloc(C.byref(self.ptr), n * 8)
        if rc != 0:
            raise MgcfdError(rc, lib.mgcfd_last_error(None).decode())
        self.array = np.ctypeslib.as_array((C.c_double * n).from_address(self.ptr.value)).reshape(shape)

    def free(self):
        if self.ptr and self.ptr.value:
            self.array = None
            load_library().mgcfd_host_free(self.ptr)
            self.ptr = C.c_void_p()


def _level_struct(lev, keep):
    """mgcfd_level_host for a dict of the reference's datasets; `keep` collects the arrays that must stay alive."""
    coords = _as(lev["node_coordinates"], np.float64)
    e2n = _as(lev["edge-->node"], np.int32)
    ewt = _as(lev["edge_weights"], np.float64)
    b2n = _as(lev["bnd_node-->node"], np.int32)
    bgr = _as(lev["bnd_node-->group"], np.int32)
    bwt = _as(lev["bnd_node_weights"], np.float64)
    mg = _as(lev["node-->mg_node"], np.int32) if "node-->mg_node" in lev else None
    keep.extend([coords, e2n, ewt, b2n, bgr, bwt, mg])
    h = LevelHost()
    h.n_nodes, h.n_edges, h.n_bnd_nodes = coords.shape[0], e2n.shape[0], b2n.shape[0]
    h.n_owned_nodes = h.n_nodes
    h.node_coordinates = coords.ctypes.data_as(_dp)
    h.edge_to_node = e2n.ctypes.data_as(_ip)
    h.edge_weights = ewt.ctypes.data_as(_dp)
    h.bnd_node_to_node = b2n.ctypes.data_as(_ip)
    h.bnd_node_to_group = bgr.ctypes.data_as(_ip)
    h.bnd_node_weights = bwt.ctypes.data_as(_dp)
    h.node_to_mg_node = mg.ctypes.data_as(_ip) if mg is not None else _ip()
    return h


def partition_levels(levels, base_array_index, n_ranks, method="geom"):
    """Owner rank of every node of every level (op_partition, euler3d.cpp:340-375): level 0 by `method` ("geom":
    recursive coordinate bisection, "kway": recursive graph bisection with FM refinement, "block", "random"), coarse
    nodes follow their lowest-numbered child."""
    lib = load_library()
    parts = []
    for l, lev in enumerate(levels):
        coords = _as(lev["node_coordinates"], np.float64)
        part = np.empty(coords.shape[0], dtype=np.int32)
        if l == 0 and method != "geom":
            e2n = _as(lev["edge-->node"], np.int32)
            rc = lib.mgcfd_partition_graph(coords.shape[0], coords.ctypes.data_as(_dp), e2n.shape[0], e2n.ctypes.data_as(_ip),
                                           int(base_array_index), int(n_ranks), method.encode(), part.ctypes.data_as(_ip))
        elif l == 0:
            rc = lib.mgcfd_partition_rcb(coords.shape[0], coords.ctypes.data_as(_dp), int(n_ranks), part.ctypes.data_as(_ip))
        else:
            fine = levels[l - 1]
            f2c = _as(fine["node-->mg_node"], np.int32)
            e2n = _as(lev["edge-->node"], np.int32)
            rc = lib.mgcfd_partition_coarse(f2c.shape[0], parts[l - 1].ctypes.data_as(_ip), f2c.ctypes.data_as(_ip),
                                            int(base_array_index), coords.shape[0], e2n.shape[0], e2n.ctypes.data_as(_ip),
                                            coords.ctypes.data_as(_dp), part.ctypes.data_as(_ip))
        if rc != 0:
            raise MgcfdError(rc, "partitioning failed")
        parts.append(part)
    return parts


class LocalMesh:
    """One rank's share of a partitioned deck: [owned | import halo] nodes, edges with an owned endpoint, boundary
    entries of owned nodes, export/import lists (mgcfd_local_mesh_build)."""

    def __init__(self, levels, base_array_index, parts, rank, n_ranks):
        self.lib = load_library()
        self.n_levels, self.rank, self.n_ranks = len(levels), int(rank), int(n_ranks)
        keep = []
        glob = (LevelHost * len(levels))(*[_level_struct(lev, keep) for lev in levels])
        parts = [_as(p, np.int32) for p in parts]
        pp = (_ip * len(levels))(*[p.ctypes.data_as(_ip) for p in parts])
        self.handle = C.c_void_p()
        rc = self.lib.mgcfd_local_mesh_build(len(levels), glob, int(base_array_index), pp, self.rank, self.n_ranks,
                                             C.byref(self.handle))
        if rc != 0:
            raise MgcfdError(rc, "mgcfd_local_mesh_build failed")

    def level(self, l):
        return self.lib.mgcfd_local_mesh_level(self.handle, l)

    def query(self, l, what):
        n = self.lib.mgcfd_local_mesh_query(self.handle, l, what.encode(), None, 0)
        if n < 0:
            raise MgcfdError(int(n), f"local_mesh_query({what})")
        out = np.empty(n, dtype=np.int32)
        self.lib.mgcfd_local_mesh_query(self.handle, l, what.encode(), out.ctypes.data_as(_ip), n)
        return out

    def sizes(self, l):
        v = self.level(l).contents
        return v.n_nodes, v.n_edges, v.n_bnd_nodes, v.n_owned_nodes

    def free(self):
        if self.handle and self.handle.value:
            self.lib.mgcfd_local_mesh_free(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class RankMesh:
    """One rank's share of a single-level deck given directly as arrays (meshgen.make_slab_rank): the same interface as
    LocalMesh, without ever building the undecomposed deck -- BASELINE.json configs[4] generates the 150M-node deck
    per partition."""

    def __init__(self, d):
        self.n_levels, self.rank, self.n_ranks = 1, int(d["rank"]), int(d["n_ranks"])
        self._keep = {
            "coords": _as(d["node_coordinates"], np.float64), "e2n": _as(d["edge-->node"], np.int32),
            "ewt": _as(d["edge_weights"], np.float64), "b2n": _as(d["bnd_node-->node"], np.int32),
            "bgr": _as(d["bnd_node-->group"], np.int32), "bwt": _as(d["bnd_node_weights"], np.float64),
            "gn": _as(d["global_node"], np.int32), "nbr": _as(d["neighbour_rank"], np.int32),
            "ep": _as(d["export_ptr"], np.int32), "ei": _as(d["export_idx"], np.int32), "ip": _as(d["import_ptr"], np.int32),
        }
        k = self._keep
        h = LevelHost()
        h.n_nodes, h.n_edges, h.n_bnd_nodes, h.n_owned_nodes = k["coords"].shape[0], k["e2n"].shape[0], k["b2n"].shape[0], int(d["n_owned"])
        h.node_coordinates = k["coords"].ctypes.data_as(_dp)
        h.edge_to_node = k["e2n"].ctypes.data_as(_ip)
        h.edge_weights = k["ewt"].ctypes.data_as(_dp)
        h.bnd_node_to_node = k["b2n"].ctypes.data_as(_ip)
        h.bnd_node_to_group = k["bgr"].ctypes.data_as(_ip)
        h.bnd_node_weights = k["bwt"].ctypes.data_as(_dp)
        h.node_to_mg_node = _ip()
        h.global_node_id = k["gn"].ctypes.data_as(_ip)
        h.n_neighbours = k["nbr"].shape[0]
        h.neighbour_rank = k["nbr"].ctypes.data_as(_ip)
        h.export_ptr = k["ep"].ctypes.data_as(_ip)
        h.export_idx = k["ei"].ctypes.data_as(_ip)
        h.import_ptr = k["ip"].ctypes.data_as(_ip)
        self._h = h

    def level(self, l):
        assert l == 0
        return C.pointer(self._h)

    def sizes(self, l):
        v = self._h
        return v.n_nodes, v.n_edges, v.n_bnd_nodes, v.n_owned_nodes

    def query(self, l, what):
        return {"global_node": self._keep["gn"], "neighbour_rank": self._keep["nbr"], "export_ptr": self._keep["ep"],
                "export_idx": self._keep["ei"], "import_ptr": self._keep["ip"], "edge_to_node": self._keep["e2n"].ravel()}[what]

    def free(self):
        pass


def group_enable_p2p(ranks):
    """switch a single-process group to the direct peer-store transport (flags instead of events + peer copies)"""
    lib = load_library()
    arr = (C.c_void_p * len(ranks))(*[r.ctx for r in ranks])
    rc = lib.mgcfd_group_enable_p2p(arr, len(ranks))
    if rc != 0:
        raise MgcfdError(rc, lib.mgcfd_last_error(ranks[0].ctx).decode())


def group_run_cycles(ranks, n_cycles):
    """V-cycles over several contexts driven by this process (one per GPU, or several on one GPU)."""
    lib = load_library()
    arr = (C.c_void_p * len(ranks))(*[r.ctx for r in ranks])
    rc = lib.mgcfd_group_run_cycles(arr, len(ranks), int(n_cycles))
    if rc != 0:
        raise MgcfdError(rc, lib.mgcfd_last_error(ranks[0].ctx).decode())


class MGCFD:
    """One context = one GPU's share of the mesh.  Methods are the op_par_loop call sites of euler3d.cpp."""

    def __init__(self, levels=None, base_array_index=1, device=0, flux_variant="owner", renumber=True,
                 exact_arith=False, owner_chunk_nodes=64, colour_block_edges=256, consts=None,
                 init=True, fuse=True, local_mesh=None, graphs=True, measure_mem_bound=False):
        """levels: list of dicts keyed by the reference's dataset names (meshgen.make_multigrid()["levels"]), or
        local_mesh: a LocalMesh (this rank's share of a partitioned deck)."""
        self.lib = load_library()
        self.n_levels = local_mesh.n_levels if local_mesh is not None else len(levels)
        opt = Options()
        self.lib.mgcfd_default_options(C.byref(opt))
        opt.flux_variant = FLUX_VARIANTS[flux_variant] if isinstance(flux_variant, str) else int(flux_variant)
        opt.renumber = int(bool(renumber))
        opt.exact_arith = int(bool(exact_arith))
        opt.owner_chunk_nodes = int(owner_chunk_nodes)
        opt.colour_block_edges = int(colour_block_edges)
        opt.no_fusion = int(not fuse)
        opt.no_graphs = int(not graphs)
        opt.measure_mem_bound = int(bool(measure_mem_bound))
        opt.rank = local_mesh.rank if local_mesh is not None else 0
        opt.n_ranks = local_mesh.n_ranks if local_mesh is not None else 1
        self.rank, self.n_ranks = opt.rank, opt.n_ranks
        self.ctx = C.c_void_p()
        rc = self.lib.mgcfd_create(C.byref(self.ctx), int(device), self.n_levels, C.byref(opt))
        if rc != 0:
            raise MgcfdError(rc, self.lib.mgcfd_last_error(None).decode())
        self.consts = consts if consts is not None else farfield_consts()
        self._ck(self.lib.mgcfd_decl_consts(self.ctx, C.byref(self.consts)))
        self.sizes, self.n_owned = [], []
        keep = []
        for l in range(self.n_levels):
            if local_mesh is not None:
                hp = local_mesh.level(l)
                h = hp.contents
                self._ck(self.lib.mgcfd_decl_level(self.ctx, l, hp, 0))          # local maps are 0-based
            else:
                h = _level_struct(levels[l], keep)
                self._ck(self.lib.mgcfd_decl_level(self.ctx, l, C.byref(h), int(base_array_index)))
            self.sizes.append((h.n_nodes, h.n_edges, h.n_bnd_nodes))
            self.n_owned.append(h.n_owned_nodes)
        self._ck(self.lib.mgcfd_plan(self.ctx))
        if init:
            self.init_loops()

    def comm_init_nccl(self, unique_id):
        """unique_id: the 128 bytes mgcfd_nccl_unique_id() produced on rank 0"""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.lib.mgcfd_comm_init_nccl(self.ctx, self.n_ranks, self.rank, buf))

    def ipc_export(self):
        """4096-byte blob (CUDA IPC handle + arena layout) for mgcfd_comm_init_ipc on the other ranks"""
        buf = C.create_string_buffer(4096)
        self._ck(self.lib.mgcfd_ipc_export(self.ctx, buf))
        return buf.raw

    def comm_init_ipc(self, blobs):
        """blobs: the ipc_export() of every rank, concatenated in rank order"""
        data = bytes(blobs)
        assert len(data) == 4096 * self.n_ranks
        self._ck(self.lib.mgcfd_comm_init_ipc(self.ctx, C.create_string_buffer(data, len(data))))

    def halo_bytes_sent(self):
        return int(self.lib.mgcfd_halo_bytes_sent(self.ctx))

    # ---- plumbing
    def _ck(self, rc):
        if rc != 0:
            raise MgcfdError(rc, self.lib.mgcfd_last_error(self.ctx).decode())

    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.lib.mgcfd_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- euler3d.cpp:413-441
    def init_loops(self):
        L = self.lib
        for l in range(self.n_levels):
            self._ck(L.mgcfd_loop_initialize_variables(self.ctx, l))
            self._ck(L.mgcfd_loop_zero_fluxes(self.ctx, l))
            self._ck(L.mgcfd_loop_zero_volumes(self.ctx, l))
            self._ck(L.mgcfd_loop_calculate_cell_volumes(self.ctx, l))
        for l in range(self.n_levels):
            self._ck(L.mgcfd_loop_dampen_ewt_edges(self.ctx, l))
            self._ck(L.mgcfd_loop_dampen_ewt_bnd(self.ctx, l))

    def reinit_variables(self):
        """initialize_variables_kernel on every level again (euler3d.cpp:414-417): the flow state a fresh run starts from"""
        for l in range(self.n_levels):
            self._ck(self.lib.mgcfd_loop_initialize_variables(self.ctx, l))

    # ---- call sites of the cycle loop
    def copy_double(self, l): self._ck(self.lib.mgcfd_loop_copy_double(self.ctx, l))
    def calculate_dt(self, l): self._ck(self.lib.mgcfd_loop_calculate_dt(self.ctx, l))

    def get_min_dt(self, l, start=np.finfo(np.float64).max):
        m = C.c_double(start)
        self._ck(self.lib.mgcfd_loop_get_min_dt(self.ctx, l, C.byref(m)))
        return m.value

    def compute_step_factor(self, l, min_dt):
        m = C.c_double(min_dt)
        self._ck(self.lib.mgcfd_loop_compute_step_factor(self.ctx, l, C.byref(m)))

    def compute_flux_edge(self, l): self._ck(self.lib.mgcfd_loop_compute_flux_edge(self.ctx, l))
    def compute_bnd_node_flux(self, l): self._ck(self.lib.mgcfd_loop_compute_bnd_node_flux(self.ctx, l))

    def time_step(self, l, rk):
        r = C.c_int(rk)
        self._ck(self.lib.mgcfd_loop_time_step(self.ctx, l, C.byref(r)))

    def unstructured_stream(self, l): self._ck(self.lib.mgcfd_loop_unstructured_stream(self.ctx, l))
    def residual(self, l): self._ck(self.lib.mgcfd_loop_residual(self.ctx, l))

    def calc_rms(self, l, start=0.0):
        r = C.c_double(start)
        self._ck(self.lib.mgcfd_loop_calc_rms(self.ctx, l, C.byref(r)))
        return r.value

    def count_bad_vals(self, l, start=0):
        c = C.c_int(start)
        self._ck(self.lib.mgcfd_loop_count_bad_vals(self.ctx, l, C.byref(c)))
        return c.value

    def up_pre(self, la): self._ck(self.lib.mgcfd_loop_up_pre(self.ctx, la))
    def up(self, la): self._ck(self.lib.mgcfd_loop_up(self.ctx, la))
    def up_post(self, la): self._ck(self.lib.mgcfd_loop_up_post(self.ctx, la))
    def down(self, l): self._ck(self.lib.mgcfd_loop_down(self.ctx, l))

    def run_cycles(self, n):
        """Device-driven V-cycles (host checks deferred to one flag read)."""
        self._ck(self.lib.mgcfd_run_cycles(self.ctx, int(n)))

    def run_cycles_host(self, n, variables_in=None, variables_out=None):
        """mgcfd_run_cycles_host: upload the flow state of the given levels (file order), run n cycles, fetch it back -- one
        call, pipelined on one GPU when the arrays are page-locked (PinnedArray.array).  Lists with one entry per level;
        None entries are skipped."""
        def table(arrs, const):
            if arrs is None:
                return None
            t = (C.c_void_p * self.n_levels)()
            for l in range(self.n_levels):
                a = arrs[l] if l < len(arrs) else None
                if a is not None:
                    assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float64 and a.size == self.sizes[l][0] * 5
                    t[l] = a.ctypes.data
            return t
        tin, tout = table(variables_in, True), table(variables_out, False)
        self._ck(self.lib.mgcfd_run_cycles_host(self.ctx, int(n), tin, tout))

    def run_cycles_loopwise(self, n_cycles):
        """euler3d.cpp:458-641 call site by call site, host checks included.  Returns (rms, min_dt)."""
        nl = self.n_levels
        level, mg_dir, i = 0, 0, 0
        rms, min_dt = 0.0, np.finfo(np.float64).max
        while i < n_cycles:
            self.copy_double(level)
            self.calculate_dt(level)
            min_dt = self.get_min_dt(level)
            if min_dt < 0.0:
                raise MgcfdError(-4, f"Fatal error during 'step factor' calculation, min_dt = {min_dt:.5e}")
            self.compute_step_factor(level, min_dt)
            for rk in range(RK):
                self.compute_flux_edge(level)
                self.compute_bnd_node_flux(level)
                self.time_step(level, rk)
            self.residual(level)
            if level == 0:
                rms = np.sqrt(self.calc_rms(level) / self.sizes[level][0])
                if self.count_bad_vals(level) > 0:
                    raise MgcfdError(-5, "Bad variable values detected, aborting")
            if nl <= 1:
                i += 1
            elif mg_dir == 0:
                level += 1
                self.up_pre(level)
                self.up(level)
                self.up_post(level)
                if level == nl - 1:
                    mg_dir = 1
            else:
                level -= 1
                self.down(level)
                if level == 0:
                    mg_dir = 0
                    i += 1
        return rms, min_dt

    # ---- data access (file order)
    def _level(self, l):
        if not 0 <= l < self.n_levels:
            raise MgcfdError(-1, "level out of range")
        return self.sizes[l]

    def fetch_into(self, l, name, out):
        """fetch into a caller-provided C-contiguous float64 array (e.g. a PinnedArray.array)"""
        self._level(l)
        assert out.flags["C_CONTIGUOUS"] and out.dtype == np.float64
        self._ck(self.lib.mgcfd_fetch_dat(self.ctx, l, name.encode(), out.ctypes.data_as(C.c_void_p)))
        return out

    def fetch(self, l, name):
        n, e, b = self._level(l)
        if name == "edge_weights":
            out = np.empty((e, 3))
        elif name == "bnd_node_weights":
            out = np.empty((b, 3))
        elif name == "up_scratch":
            out = np.empty(n, dtype=np.int32)
        else:
            if name not in _DAT_DIMS:
                raise MgcfdError(-1, f"unknown dat '{name}'")
            d = _DAT_DIMS[name]
            out = np.empty((n, d) if d > 1 else n)
        self._ck(self.lib.mgcfd_fetch_dat(self.ctx, l, name.encode(), out.ctypes.data_as(C.c_void_p)))
        return out

    def set(self, l, name, arr):
        self._level(l)
        dt = np.int32 if name == "up_scratch" else np.float64
        a = _as(arr, dt)
        self._ck(self.lib.mgcfd_set_dat(self.ctx, l, name.encode(), a.ctypes.data_as(C.c_void_p)))

    def sync(self): self._ck(self.lib.mgcfd_sync(self.ctx))

    def validate(self, l, master):
        m = _as(master, np.float64)
        c = C.c_int(0)
        self._ck(self.lib.mgcfd_validate_level(self.ctx, l, m.ctypes.data_as(_dp), C.byref(c)))
        return c.value

    def plan_query(self, l, what):
        n = self.lib.mgcfd_plan_query(self.ctx, l, what.encode(), None, 0)
        if n < 0:
            raise MgcfdError(int(n), f"plan_query({what})")
        out = np.empty(n, dtype=np.int32)
        self.lib.mgcfd_plan_query(self.ctx, l, what.encode(), out.ctypes.data_as(_ip), n)
        return out

    def set_flux_variant(self, v):
        self._ck(self.lib.mgcfd_set_flux_variant(self.ctx, FLUX_VARIANTS[v] if isinstance(v, str) else int(v)))

    # ---- measurement
    def timers_enable(self, on=True): self._ck(self.lib.mgcfd_timers_enable(self.ctx, int(on)))
    def timers_reset(self): self._ck(self.lib.mgcfd_timers_reset(self.ctx))

    def timer(self, loop, level=-1):
        ms, calls, el = C.c_double(0), C.c_longlong(0), C.c_longlong(0)
        self._ck(self.lib.mgcfd_timers_get(self.ctx, loop.encode(), level, C.byref(ms), C.byref(calls), C.byref(el)))
        return ms.value, calls.value, el.value

    def kernel_launches(self):
        return int(self.lib.mgcfd_kernel_launches(self.ctx))

    def stream(self):
        return self.lib.mgcfd_stream(self.ctx)
