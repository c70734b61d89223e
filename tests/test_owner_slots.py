"""Bank-aware edge slotting of the owner chunks (device packing, csrc/api.cu bank_aware_slots): the slots of a
chunk are a permutation of its plan edges, and they cut the shared-memory bank conflicts the kernel's two gather
patterns see.  CPU only (planning-only context)."""
import numpy as np
import pytest


def wavefronts(ids):
    """8-byte accesses of one half-warp: distinct words per bank, worst bank"""
    banks = {}
    for v in ids:
        banks.setdefault(int(v) % 16, set()).add(int(v))
    return max(len(s) for s in banks.values()) if banks else 0


@pytest.mark.parametrize("name,chunk", [("tiny", 64), ("small", 64), ("medium", 64), ("small", 128)])
@pytest.mark.parametrize("split", [1, 2])
def test_slots_are_permutations_and_cut_conflicts(pkg, meshgen, name, chunk, split):
    mesh = meshgen.make_multigrid(name)
    with pkg.MGCFD(mesh["levels"], device=-1, init=False, owner_chunk_nodes=chunk) as g:
        for l in range(len(mesh["levels"])):
            eo = g.plan_query(l, "owner_edge_off")
            lab = g.plan_query(l, "owner_lab").view(np.uint32)
            slots = g.plan_query(l, f"owner_slots_split{split}")
            assert slots.size == lab.size == eo[-1]
            plan_cost = slot_cost = groups = 0
            for k in range(len(eo) - 1):
                s = slots[eo[k]:eo[k + 1]]
                assert np.array_equal(np.sort(s), np.arange(s.size)), (l, k)
                inv = np.argsort(s)
                for end in (lab[eo[k]:eo[k + 1]] & 0xffff, lab[eo[k]:eo[k + 1]] >> 16):
                    for h in range(0, s.size, 16):
                        plan_cost += wavefronts(end[h:h + 16])
                        slot_cost += wavefronts(end[inv][h:h + 16])
                        groups += 1
            # the edge phase's endpoint-state gathers: conflict-free would be 1.0 wavefront per half-warp
            assert slot_cost <= plan_cost
            if groups >= 200:
                assert slot_cost / groups < 1.25, (l, slot_cost / groups, plan_cost / groups)


def test_slots_are_deterministic(pkg, meshgen):
    mesh = meshgen.make_multigrid("small")
    with pkg.MGCFD(mesh["levels"], device=-1, init=False) as a, pkg.MGCFD(mesh["levels"], device=-1, init=False) as b:
        assert np.array_equal(a.plan_query(0, "owner_slots_split2"), b.plan_query(0, "owner_slots_split2"))


def test_batched_slot_accessor_equals_the_whole_plan(pkg, meshgen):
    """device packing asks for the slots chunk by chunk, computed a batch of chunks ahead: same answer for any batch size"""
    mesh = meshgen.make_multigrid("small")
    with pkg.MGCFD(mesh["levels"], device=-1, init=False) as g:
        for l in range(len(mesh["levels"])):
            assert np.array_equal(g.plan_query(l, "owner_slots_split2_batch7"), g.plan_query(l, "owner_slots_split2"))
