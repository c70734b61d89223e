"""The C-ABI library loads on a machine without a GPU and exports every symbol include/mgcfd_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mgcfd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mgcfd_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    assert header_symbols() == sorted(pkg.capi.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_struct_layouts_match_header(pkg):
    assert ctypes.sizeof(pkg.capi.Consts) == 8 * 18 + 8
    assert ctypes.sizeof(pkg.capi.LevelHost) == 16 + 7 * 8 + 8 + 8 + 4 * 8
    assert ctypes.sizeof(pkg.capi.Options) == 16 * 4


def test_host_only_entry_points(pkg):
    lib = pkg.load_library()
    assert b"sm_100a" in lib.mgcfd_version()
    c = pkg.farfield_consts()
    assert c.smoothing_coefficient == float.fromhex("0x1.99999ap-3")      # double(0.2f), euler3d.cpp:47
    assert list(c.ff_variable)[:2] == [1.4, 1.4 * 1.2]


def test_no_cpu_fallback(pkg):
    """without a CUDA device creation fails loudly with MGCFD_ERR_NODEVICE"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    mesh = pkg.meshgen.make_multigrid("tiny")
    with pytest.raises(pkg.MgcfdError) as ei:
        pkg.MGCFD(mesh["levels"])
    assert ei.value.code == -3


def test_planning_only_context_refuses_compute(pkg):
    mesh = pkg.meshgen.make_multigrid("tiny")
    with pkg.MGCFD(mesh["levels"], device=-1, init=False) as ctx:
        assert ctx.plan_query(0, "node_perm").size == mesh["levels"][0]["node_coordinates"].shape[0]
        for call in (lambda: ctx.compute_flux_edge(0), lambda: ctx.run_cycles(1), lambda: ctx.fetch(0, "variables"),
                     lambda: ctx.time_step(0, 0), lambda: ctx.init_loops()):
            with pytest.raises(pkg.MgcfdError) as ei:
                call()
            assert ei.value.code == -3


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "mg-cfd-app-op2_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import orc" not in text and "oracle_api.h" not in text and "libmgcfd_oracle" not in text, f
