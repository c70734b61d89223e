"""Round-2 preparation (branch r2-prep): the lean owner kernel (MGCFD_OWNER_LEAN=1).  NOT yet run on a GPU; enable with
MGCFD_TEST_EXPERIMENTAL=1."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("MGCFD_TEST_EXPERIMENTAL") != "1", reason="experimental kernel")]


def test_lean_kernel_matches_golden(pkg, meshgen, golden, monkeypatch):
    monkeypatch.setenv("MGCFD_OWNER_LEAN", "1")
    g = golden("small_cycles10.npz")
    mesh = meshgen.make_multigrid("small")
    with pkg.MGCFD(mesh["levels"], flux_variant="owner") as gpu:
        gpu.run_cycles(10)
        for l in range(len(mesh["levels"])):
            got, ref = gpu.fetch(l, "variables"), g[f"var_L{l}"]
            assert (np.abs(got - ref).max(axis=0) <= 1e-10 * np.abs(ref).max(axis=0)).all()
            assert gpu.validate(l, ref) == 0
