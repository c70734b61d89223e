"""Independent numpy/Python restatement of the host planner's index sets (TEST INFRASTRUCTURE ONLY).

The reference keeps renumbering, plan colouring and partitioning inside the OP2 library, which is not
in its tree and is unpinned: "parity unpinned" against OP2 itself.  What CAN be
pinned is that the product's C++ planner (mg-cfd-app-op2_b200/csrc/plan.cpp, partition.cpp) implements
the documented algorithms exactly: this module restates each of them from the specification in
DESIGN.md with different data structures (sets, dicts, numpy sorts) and the tests compare bit for bit.

  hilbert_renumber   3-D Hilbert curve (Skilling transpose), 16 bits/axis, ties by file index
  sort_edges         by (min internal endpoint, max internal endpoint, file index)
  colour_edges       OP2-style two-level first-fit colouring over blocks of consecutive sorted edges
  owner_chunks       greedy owner-compute chunking under (owned, local, edge) caps
"""
from __future__ import annotations

import numpy as np

HILBERT_BITS = 16


def hilbert_keys(q):
    """q: uint64 [3, N] quantised coordinates -> Hilbert index per point (Skilling 2004, AxestoTranspose)."""
    x = q.astype(np.uint64).copy()
    one = np.uint64(1)
    m = one << np.uint64(HILBERT_BITS - 1)
    qq = m
    while qq > one:
        p = qq - one
        for i in range(3):
            hit = (x[i] & qq) != 0
            x0_inv = x[0] ^ p                       # invert low bits of x[0]
            t = (x[0] ^ x[i]) & p                   # or exchange low bits of x[0] and x[i]
            x0_swp, xi_swp = x[0] ^ t, x[i] ^ t
            xi_new = np.where(hit, x[i], xi_swp)
            x0_new = np.where(hit, x0_inv, x0_swp)
            if i == 0:
                x[0] = x0_new                       # (exchange with itself is the identity)
            else:
                x[0], x[i] = x0_new, xi_new
        qq >>= one
    for i in range(1, 3):
        x[i] ^= x[i - 1]
    t = np.zeros(x.shape[1], dtype=np.uint64)
    qq = m
    while qq > one:
        t = np.where((x[2] & qq) != 0, t ^ (qq - one), t)
        qq >>= one
    for i in range(3):
        x[i] ^= t
    key = np.zeros(x.shape[1], dtype=np.uint64)
    for b in range(HILBERT_BITS - 1, -1, -1):
        for i in range(3):
            key = (key << one) | ((x[i] >> np.uint64(b)) & one)
    return key


def hilbert_renumber(coords, n_owned=None):
    """new_of_old: internal index of file node i.  Owned nodes are ordered along the curve, halo nodes keep
    their relative order behind them."""
    n = coords.shape[0]
    no = n if n_owned is None else n_owned
    new_of_old = np.arange(n, dtype=np.int64)
    if no > 1:
        c = np.asarray(coords[:no], dtype=np.float64)
        lo = c.min(axis=0)
        span = float((c.max(axis=0) - lo).max())
        if not span > 0.0:
            span = 1.0
        q = np.floor((c - lo) / span * float((1 << HILBERT_BITS) - 1) + 0.5).astype(np.uint64)
        keys = hilbert_keys(q.T)
        old_of_new = np.lexsort((np.arange(no), keys))
        new_of_old[old_of_new] = np.arange(no)
    return new_of_old.astype(np.int32)


def sort_edges(e2n0, new_of_old):
    p = np.asarray(new_of_old, dtype=np.int64)
    a, b = p[e2n0[:, 0]], p[e2n0[:, 1]]
    return np.lexsort((np.arange(e2n0.shape[0]), np.maximum(a, b), np.minimum(a, b))).astype(np.int32)


def _first_free(used):
    c = 0
    while c in used:
        c += 1
    return c


def colour_edges(e2n0, new_of_old, order, block_edges):
    """returns (thread_colour[E] by FILE edge, block_colour[E] by FILE edge, n_block_colours)"""
    p = np.asarray(new_of_old, dtype=np.int64)
    E = e2n0.shape[0]
    thread_colour = np.zeros(E, dtype=np.int32)
    block_colour = np.zeros(E, dtype=np.int32)
    node_block_colours = {}
    nbc = 0
    for lo in range(0, E, block_edges):
        blk = order[lo:lo + block_edges]
        used = {}
        nodes = set()
        for e in blk:
            a, b = int(p[e2n0[e, 0]]), int(p[e2n0[e, 1]])
            ua, ub = used.setdefault(a, set()), used.setdefault(b, set())
            c = _first_free(ua | ub)
            ua.add(c)
            ub.add(c)
            thread_colour[e] = c
            nodes.add(a)
            nodes.add(b)
        taken = set()
        for v in nodes:
            taken |= node_block_colours.get(v, set())
        bc = _first_free(taken)
        for v in nodes:
            node_block_colours.setdefault(v, set()).add(bc)
        block_colour[blk] = bc
        nbc = max(nbc, bc + 1)
    return thread_colour, block_colour, nbc


def check_colouring(e2n0, order, block_edges, thread_colour, block_colour):
    """validity: no two same-colour edges of a block share a node; no two same-colour blocks do"""
    E = e2n0.shape[0]
    seen_block = {}
    for k, lo in enumerate(range(0, E, block_edges)):
        blk = order[lo:lo + block_edges]
        seen = set()
        for e in blk:
            for v in (int(e2n0[e, 0]), int(e2n0[e, 1])):
                key = (v, int(thread_colour[e]))
                if key in seen:
                    return False
                seen.add(key)
        bc = int(block_colour[blk[0]])
        for v in set(e2n0[blk].ravel().tolist()):
            if seen_block.setdefault((v, bc), k) != k:
                return False
    return True


def owner_chunks(e2n0, new_of_old, n_owned, max_own, max_loc, max_edges):
    """returns (chunk_start[], halo lists (internal ids, ascending) per chunk, edge lists (file ids) per chunk).
    A chunk is closed only at an even number of owned nodes, so every chunk starts on an even node index."""
    p = np.asarray(new_of_old, dtype=np.int64)
    n = p.shape[0]
    a, b = p[e2n0[:, 0]], p[e2n0[:, 1]]
    inc = [[] for _ in range(n)]           # incident file edges per internal node, ascending file id
    for e in range(e2n0.shape[0]):
        inc[a[e]].append(e)
        inc[b[e]].append(e)
    starts, halos, edges = [0], [], []
    v = 0
    while v < n_owned:
        start = v
        halo, seen_edges, elist = set(), set(), []
        n_own = 0
        while v < n_owned:
            deg = len(inc[v])
            if n_own > 0 and n_own % 2 == 0 and (n_own + 1 > max_own or n_own + len(halo) + 1 + deg > max_loc or len(elist) + deg > max_edges):
                break
            halo.discard(v)
            for e in inc[v]:
                if e in seen_edges:
                    continue
                seen_edges.add(e)
                elist.append(e)
                u = int(b[e]) if int(a[e]) == v else int(a[e])
                if not (start <= u < v):
                    halo.add(u)
            n_own += 1
            v += 1
        halo = sorted(h for h in halo if not (start <= h < v))
        starts.append(v)
        halos.append(halo)
        edges.append(elist)
    return starts, halos, edges
