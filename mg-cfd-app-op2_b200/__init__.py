"""mgcfd_b200 -- B200-native MG-CFD hot path (loaded under this name by __graft_entry__.load_package()).

Contents: csrc/ (CUDA kernels, host planner, C-ABI -> libmgcfd_b200.so), capi.py (ctypes
mirror of the op_par_loop call sites), meshgen.py (synthetic decks in the reference's
level-file layout), host/ (native euler3d driver).  Nothing here imports oracle/.
"""
from . import meshgen  # noqa: F401
from . import capi  # noqa: F401
from .capi import (MGCFD, MgcfdError, PinnedArray, LocalMesh, RankMesh, load_library, farfield_consts,  # noqa: F401
                   partition_levels, group_run_cycles, group_enable_p2p, nccl_unique_id)
