"""ctypes binding of the CPU checkers (TEST INFRASTRUCTURE ONLY, see oracle_api.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  `load("ref")` is the reference's own headers compiled in place
(oracle/_ref, prebuilt in the build container); `load("port")` is the plain-C restatement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NVAR, NDIM = 5, 3

_LIBS = {
    "ref": "_ref/libmgcfd_ref.so",
    "ref_fast": "_ref/libmgcfd_ref_fast.so",
    "port": "libmgcfd_oracle.so",
    "port_fast": "libmgcfd_oracle_fast.so",
}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class OrcLevel(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("n_edges", C.c_int), ("n_bnd", C.c_int), ("pad_", C.c_int),
                ("coords", _dp), ("e2n", _ip), ("ewt", _dp), ("b2n", _ip), ("bgroup", _ip), ("bwt", _dp),
                ("mg", _ip), ("var", _dp), ("old", _dp), ("res", _dp), ("flux", _dp), ("vol", _dp),
                ("sf", _dp), ("up_scratch", _ip)]


class OrcStats(C.Structure):
    _fields_ = [("wall_total", C.c_double), ("wall_flux_edge", C.c_double), ("flux_edges", C.c_longlong),
                ("cycles", C.c_longlong), ("last_min_dt", C.c_double), ("last_rms", C.c_double)]


def build(force=False):
    """(Re)build the checker libraries with oracle/Makefile (the _ref part only where /root/reference exists)."""
    if force:
        subprocess.run(["make", "-C", HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


def available(kind):
    return os.path.exists(os.path.join(HERE, _LIBS[kind]))


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Oracle:
    """One loaded checker library; methods mirror the op_par_loop call sites of euler3d.cpp."""

    def __init__(self, kind="port"):
        path = os.path.join(HERE, _LIBS[kind])
        if not os.path.exists(path):
            build()
        self.kind = kind
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_name.restype = C.c_char_p
        L.orc_validate_count.restype = C.c_int
        L.orc_run_cycles.restype = C.c_int
        L.orc_get_threads.restype = C.c_int
        self.consts = np.zeros(18)
        L.orc_set_farfield(_d(self.consts))

    # --- constants (euler3d.cpp:47,157-189)
    @property
    def smoothing(self):
        return float(self.consts[0])

    @property
    def ff_variable(self):
        return self.consts[1:6].copy()

    def name(self):
        return self.lib.orc_name().decode()

    def set_threads(self, n):
        self.lib.orc_set_threads(int(n))
        return self.lib.orc_get_threads()

    # --- single loops; arrays are float64 / int32 C-contiguous numpy, updated in place
    def initialize_variables(self, var):
        self.lib.orc_initialize_variables(C.c_int(var.shape[0]), _d(var))

    def calculate_cell_volumes(self, e2n, coords, ewt, vol):
        self.lib.orc_calculate_cell_volumes(C.c_int(e2n.shape[0]), _i(e2n), _d(coords), _d(ewt), _d(vol))

    def dampen_ewt(self, w):
        self.lib.orc_dampen_ewt(C.c_int(w.shape[0]), _d(w))

    def copy_double(self, var, old):
        self.lib.orc_copy_double(C.c_int(var.shape[0]), _d(var), _d(old))

    def calculate_dt(self, var, vol, sf):
        self.lib.orc_calculate_dt(C.c_int(var.shape[0]), _d(var), _d(vol), _d(sf))

    def get_min_dt(self, sf, start=np.finfo(np.float64).max):
        m = C.c_double(start)
        self.lib.orc_get_min_dt(C.c_int(sf.shape[0]), _d(sf), C.byref(m))
        return m.value

    def compute_step_factor(self, var, vol, min_dt, sf):
        m = C.c_double(min_dt)
        self.lib.orc_compute_step_factor(C.c_int(var.shape[0]), _d(var), _d(vol), C.byref(m), _d(sf))

    def compute_flux_edge(self, e2n, var, ewt, flux):
        self.lib.orc_compute_flux_edge(C.c_int(e2n.shape[0]), _i(e2n), _d(var), _d(ewt), _d(flux))

    def unstructured_stream(self, e2n, var, ewt, flux):
        self.lib.orc_unstructured_stream(C.c_int(e2n.shape[0]), _i(e2n), _d(var), _d(ewt), _d(flux))

    def compute_bnd_node_flux(self, bgroup, bwt, b2n, var, flux):
        self.lib.orc_compute_bnd_node_flux(C.c_int(b2n.shape[0]), _i(bgroup), _d(bwt), _i(b2n), _d(var), _d(flux))

    def time_step(self, rk, sf, flux, old, var):
        self.lib.orc_time_step(C.c_int(var.shape[0]), C.c_int(rk), _d(sf), _d(flux), _d(old), _d(var))

    def residual(self, old, var, res):
        self.lib.orc_residual(C.c_int(var.shape[0]), _d(old), _d(var), _d(res))

    def calc_rms(self, res):
        r = C.c_double(0.0)
        self.lib.orc_calc_rms(C.c_int(res.shape[0]), _d(res), C.byref(r))
        return r.value

    def count_bad_vals(self, var):
        c = C.c_int(0)
        self.lib.orc_count_bad_vals(C.c_int(var.shape[0]), _d(var), C.byref(c))
        return c.value

    def up_pre(self, mg, var_above, scratch_above):
        self.lib.orc_up_pre(C.c_int(mg.shape[0]), _i(mg), _d(var_above), _i(scratch_above))

    def up(self, mg, var, var_above, scratch_above):
        self.lib.orc_up(C.c_int(mg.shape[0]), _i(mg), _d(var), _d(var_above), _i(scratch_above))

    def up_post(self, var, scratch):
        self.lib.orc_up_post(C.c_int(var.shape[0]), _d(var), _i(scratch))

    def down(self, mg, var, res, coords, res_above, coords_above):
        self.lib.orc_down(C.c_int(mg.shape[0]), _i(mg), _d(var), _d(res), _d(coords), _d(res_above),
                          _d(coords_above))

    def validate_count(self, test, master):
        return self.lib.orc_validate_count(C.c_int(test.shape[0]), _d(test), _d(master))

    # --- whole-run driver
    def make_state(self, mesh0):
        """Allocate per-level arrays for a 0-based mesh (list of level dicts as meshgen.zero_based returns)."""
        return OracleRun(self, mesh0)


class OracleRun:
    """Per-level arrays + the restated euler3d.cpp driver (init :413-441, cycles :458-641)."""

    def __init__(self, orc, levels0):
        self.orc = orc
        self.levels = []
        self.c_levels = (OrcLevel * len(levels0))()
        for l, lev in enumerate(levels0):
            n = lev["node_coordinates"].shape[0]
            a = {
                "coords": np.array(lev["node_coordinates"], dtype=np.float64, order="C"),
                "e2n": np.array(lev["edge-->node"], dtype=np.int32, order="C").reshape(-1, 2),
                "ewt": np.array(lev["edge_weights"], dtype=np.float64, order="C"),
                "b2n": np.array(lev["bnd_node-->node"], dtype=np.int32, order="C").reshape(-1),
                "bgroup": np.array(lev["bnd_node-->group"], dtype=np.int32, order="C").reshape(-1),
                "bwt": np.array(lev["bnd_node_weights"], dtype=np.float64, order="C"),
                "mg": (np.array(lev["node-->mg_node"], dtype=np.int32, order="C").reshape(-1)
                       if "node-->mg_node" in lev else None),
                "var": np.zeros((n, NVAR)), "old": np.zeros((n, NVAR)), "res": np.zeros((n, NVAR)),
                "flux": np.zeros((n, NVAR)), "vol": np.zeros(n), "sf": np.zeros(n),
                "up_scratch": np.zeros((n, 2), dtype=np.int32),
            }
            self.levels.append(a)
            c = self.c_levels[l]
            c.n_nodes, c.n_edges, c.n_bnd = n, a["e2n"].shape[0], a["b2n"].shape[0]
            for name in ("coords", "ewt", "bwt", "var", "old", "res", "flux", "vol", "sf"):
                setattr(c, name, _d(a[name]))
            for name in ("e2n", "b2n", "bgroup", "up_scratch"):
                setattr(c, name, _i(a[name]))
            c.mg = _i(a["mg"]) if a["mg"] is not None else _ip()

    def init(self):
        self.orc.lib.orc_init_levels(self.c_levels, C.c_int(len(self.levels)))

    def run(self, n_cycles):
        st = OrcStats()
        rc = self.orc.lib.orc_run_cycles(self.c_levels, C.c_int(len(self.levels)), C.c_int(n_cycles), C.byref(st))
        return rc, st
