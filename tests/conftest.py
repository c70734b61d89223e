"""Shared fixtures.  GPU tests are marked @pytest.mark.gpu; everything else runs on CPU.

oracle/ is test infrastructure: it is imported here (and in bench.py's cpu_baseline leg and
__graft_entry__.smoke()) only, never by the product package.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

# virtual-rank tests put up to 16 streams with spin-wait kernels on one GPU: give every stream its own hardware queue
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_package():
    """Import the product package (its directory name has hyphens) as `mgcfd_b200`."""
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def meshgen(pkg):
    return pkg.meshgen


@pytest.fixture(scope="session")
def orc_mod():
    import orc
    return orc


@pytest.fixture(scope="session")
def oracle_port(orc_mod):
    return orc_mod.Oracle("port")


@pytest.fixture(scope="session")
def oracle_ref(orc_mod):
    if not orc_mod.available("ref") and not os.path.exists("/root/reference/flux.h"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return orc_mod.Oracle("ref")


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        return dict(np.load(os.path.join(GOLDEN, name)))
    return _load


def mesh0(meshgen, name):
    return [meshgen.zero_based(l) for l in meshgen.make_multigrid(name)["levels"]]
