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


# ---------------------------------------------------------------------------------------------------
# domain decomposition (restates mg-cfd-app-op2_b200/csrc/partition.cpp from its specification)
# ---------------------------------------------------------------------------------------------------
def rcb(coords, n_parts):
    """recursive coordinate bisection: median split of the longest bounding-box axis, ties by index, shares
    proportional to the rank counts of the halves (left half gets floor(parts/2) ranks)"""
    part = np.empty(coords.shape[0], dtype=np.int32)

    def rec(ids, p0, p1):
        if p1 - p0 == 1:
            part[ids] = p0
            return
        c = coords[ids]
        ext = c.max(axis=0) - c.min(axis=0) if ids.size else np.zeros(3)
        axis = 0
        for d in (1, 2):
            if ext[d] > ext[axis]:
                axis = d
        order = ids[np.lexsort((ids, c[:, axis]))] if ids.size else ids
        left_parts = (p1 - p0) // 2
        n_left = (order.size * left_parts) // (p1 - p0)
        rec(order[:n_left], p0, p0 + left_parts)
        rec(order[n_left:], p0 + left_parts, p1)

    rec(np.arange(coords.shape[0], dtype=np.int64), 0, n_parts)
    return part


def kway(coords, e2n0, n_parts):
    """recursive graph bisection (partition.cpp "kway"): coordinate split of the subset, then Fiduccia-Mattheyses
    passes on the edges inside it -- left side kept within max(1, |S|//200) of its share; a pass moves the unlocked node
    of highest gain whose move keeps the balance (ties: lowest id), stops after 64 moves without a new best cumulative
    gain, rolls back to the first best prefix; at most 8 passes, while a pass gains something"""
    import heapq
    n = coords.shape[0]
    adj = [[] for _ in range(n)]
    for a, b in np.asarray(e2n0).tolist():
        adj[a].append(b)
        adj[b].append(a)
    part = np.empty(n, dtype=np.int32)

    def refine(S, side, n_left):
        inside = set(S)
        tol = max(1, len(S) // 200)
        c0 = sum(1 for v in S if side[v] == 0)
        for _ in range(8):
            gain, locked = {}, set()
            heaps = ([], [])
            for v in S:
                g = 0
                for u in adj[v]:
                    if u in inside:
                        g += 1 if side[u] != side[v] else -1
                gain[v] = g
                heapq.heappush(heaps[side[v]], (-g, v))

            def top(s):
                h = heaps[s]
                while h and (h[0][1] in locked or -h[0][0] != gain[h[0][1]]):
                    heapq.heappop(h)
                return h[0] if h else None

            moves, cur, best, best_len, run = [], 0, 0, 0, c0
            while len(moves) < len(S):
                t0, t1 = top(0), top(1)
                ok0 = t0 is not None and run - 1 >= n_left - tol
                ok1 = t1 is not None and run + 1 <= n_left + tol
                if not ok0 and not ok1:
                    break
                frm = 0 if ok0 else 1
                if ok0 and ok1:
                    frm = 1 if t1 < t0 else 0
                _, v = heapq.heappop(heaps[frm])
                locked.add(v)
                cur += gain[v]
                side[v] = 1 - frm
                run += -1 if frm == 0 else 1
                for u in adj[v]:
                    if u in inside and u not in locked:
                        gain[u] += 2 if side[u] == frm else -2
                        heapq.heappush(heaps[side[u]], (-gain[u], u))
                moves.append(v)
                if cur > best:
                    best, best_len = cur, len(moves)
                if len(moves) - best_len >= 64:
                    break
            for v in moves[best_len:]:
                side[v] = 1 - side[v]
            c0 = sum(1 for v in S if side[v] == 0)
            if best <= 0:
                break

    def rec(ids, p0, p1):
        if p1 - p0 == 1:
            part[ids] = p0
            return
        c = coords[ids]
        ext = c.max(axis=0) - c.min(axis=0) if ids.size else np.zeros(3)
        axis = 0
        for d in (1, 2):
            if ext[d] > ext[axis]:
                axis = d
        order = ids[np.lexsort((ids, c[:, axis]))] if ids.size else ids
        left_parts = (p1 - p0) // 2
        n_left = (order.size * left_parts) // (p1 - p0)
        side = {int(v): (0 if i < n_left else 1) for i, v in enumerate(order.tolist())}
        refine(order.tolist(), side, n_left)
        left = np.array(sorted(v for v in side if side[v] == 0), dtype=np.int64)
        right = np.array(sorted(v for v in side if side[v] == 1), dtype=np.int64)
        rec(left, p0, p0 + left_parts)
        rec(right, p0 + left_parts, p1)

    rec(np.arange(n, dtype=np.int64), 0, n_parts)
    return part


def edge_cut(part, e2n0):
    e = np.asarray(e2n0)
    return int((part[e[:, 0]] != part[e[:, 1]]).sum())


def coarse_part(fine_part, fine_to_coarse0, n_coarse, coarse_e2n0, coarse_coords):
    part = np.full(n_coarse, -1, dtype=np.int32)
    parents, first = np.unique(fine_to_coarse0, return_index=True)     # first occurrence = lowest-numbered child
    part[parents] = fine_part[first]
    orphans = np.nonzero(part < 0)[0]
    if orphans.size:
        resolved = {}
        orphan_set = set(orphans.tolist())
        nbrs = {int(o): [] for o in orphans}
        for a, b in coarse_e2n0.tolist():
            if a in orphan_set and b not in orphan_set:
                nbrs[a].append(b)
            if b in orphan_set and a not in orphan_set:
                nbrs[b].append(a)
        for o in orphans.tolist():
            best, best_d = -1, 0.0
            for nb in nbrs[o]:
                t = coarse_coords[o] - coarse_coords[nb]
                d2 = 0.0
                for k in range(3):
                    d2 += float(t[k]) * float(t[k])
                if best < 0 or d2 < best_d or (d2 == best_d and nb < best):
                    best, best_d = nb, d2
            resolved[o] = int(part[best]) if best >= 0 else 0
        for o, r in resolved.items():
            part[o] = r
    return part


def partition_levels(levels0, n_ranks, method="geom"):
    if method == "kway":
        parts = [kway(np.asarray(levels0[0]["node_coordinates"]), levels0[0]["edge-->node"], n_ranks)]
    else:
        parts = [rcb(np.asarray(levels0[0]["node_coordinates"]), n_ranks)]
    for l in range(1, len(levels0)):
        parts.append(coarse_part(parts[l - 1], levels0[l - 1]["node-->mg_node"].reshape(-1),
                                 levels0[l]["node_coordinates"].shape[0], levels0[l]["edge-->node"],
                                 np.asarray(levels0[l]["node_coordinates"])))
    return parts


def local_mesh(levels0, parts, rank):
    """per level: dict(global_node, n_owned, global_edge, global_bnd, e2n (local), mg (local or -1),
    neighbour_rank, export lists (global ids per neighbour), import lists (global ids per neighbour))"""
    nl = len(levels0)
    needs = [set() for _ in range(nl)]                 # (q, n): rank q reads remote node n of level l
    for l, lev in enumerate(levels0):
        p = parts[l]
        e = lev["edge-->node"]
        cut = p[e[:, 0]] != p[e[:, 1]]
        for a, b in e[cut].tolist():
            needs[l].add((int(p[a]), b))
            needs[l].add((int(p[b]), a))
        if l + 1 < nl:
            mg = lev["node-->mg_node"].reshape(-1)
            pc = parts[l + 1][mg]
            diff = np.nonzero(pc != p)[0]
            for f in diff.tolist():
                needs[l].add((int(pc[f]), f))                       # restrict on the parent's rank reads the child
                needs[l + 1].add((int(p[f]), int(mg[f])))           # prolong on the child's rank reads the parent
    out = []
    for l, lev in enumerate(levels0):
        p = parts[l]
        owned = np.nonzero(p == rank)[0]
        halo = sorted((n for q, n in needs[l] if q == rank), key=lambda n: (int(p[n]), n))
        exports = sorted((q, n) for q, n in needs[l] if q != rank and p[n] == rank)
        nbrs = sorted({int(p[n]) for n in halo} | {q for q, _ in exports})
        global_node = np.concatenate([owned, np.array(halo, dtype=np.int64)]).astype(np.int64)
        local_of = np.full(p.shape[0], -1, dtype=np.int64)
        local_of[global_node] = np.arange(global_node.size)
        e = lev["edge-->node"]
        emask = (p[e[:, 0]] == rank) | (p[e[:, 1]] == rank)
        b2n = lev["bnd_node-->node"].reshape(-1)
        bmask = p[b2n] == rank
        d = {
            "global_node": global_node, "n_owned": owned.size,
            "global_edge": np.nonzero(emask)[0], "global_bnd": np.nonzero(bmask)[0],
            "e2n": local_of[e[emask]],
            "neighbour_rank": nbrs,
            "exports": {q: [n for qq, n in exports if qq == q] for q in nbrs},
            "imports": {q: [n for n in halo if p[n] == q] for q in nbrs},
        }
        out.append(d)
    for l in range(nl - 1):
        mg = levels0[l]["node-->mg_node"].reshape(-1)
        local_coarse = np.full(parts[l + 1].shape[0], -1, dtype=np.int64)
        local_coarse[out[l + 1]["global_node"]] = np.arange(out[l + 1]["global_node"].size)
        out[l]["mg"] = local_coarse[mg[out[l]["global_node"]]]
    return out
