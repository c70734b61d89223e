// partition.cpp -- domain decomposition for 1/2/4/8 GPUs: node ownership per multigrid level and the per-rank
// local meshes with their owner / import-halo / export index sets (op_partition at
// euler3d.cpp:340-375 does this inside OP2 in the reference).  Host code, deterministic; restated independently in
// oracle/plan_oracle.py and compared bit for bit by tests/test_partition.py.
//
//   level 0    recursive coordinate bisection: split the node set at the median of its longest bounding-box axis
//              (ties by node index) into shares proportional to the rank counts of the two halves
//   level l+1  a coarse node belongs to the owner of its lowest-numbered child; a childless coarse node to the
//              owner of its nearest edge-neighbour that has children (ties by index), else rank 0
//   rank mesh  owned nodes (ascending file index) | import halo (by owner rank, then file index);
//              every edge with an owned endpoint (cut edges live on both ranks and are recomputed there);
//              boundary entries of owned nodes.  Halo = remote edge neighbours + remote children of owned coarse
//              nodes (restrict) + remote parents of owned fine nodes (prolong).
//              Export list to rank q = my owned nodes in q's halo, ascending file index = q's import order.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <set>
#include <string>

#include "mgcfd_b200.h"

namespace {

struct LocalLevel {
    std::vector<double> coords, ewt, bwt;
    std::vector<int> e2n, b2n, bgroup, mg;
    std::vector<int> global_node, global_edge, global_bnd;
    std::vector<int> nbr_rank, export_ptr, export_idx, import_ptr;
    int n_owned = 0;
    mgcfd_level_host view{};
};

void rcb(std::vector<int> &ids, int lo, int hi, int p0, int p1, const double *xyz, int *part)
{
    if (p1 - p0 == 1) {
        for (int i = lo; i < hi; i++) part[ids[i]] = p0;
        return;
    }
    double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int i = lo; i < hi; i++)
        for (int d = 0; d < 3; d++) {
            double v = xyz[(size_t)ids[i] * 3 + d];
            if (i == lo || v < mn[d]) mn[d] = v;
            if (i == lo || v > mx[d]) mx[d] = v;
        }
    int axis = 0;
    for (int d = 1; d < 3; d++)
        if (mx[d] - mn[d] > mx[axis] - mn[axis]) axis = d;
    std::sort(ids.begin() + lo, ids.begin() + hi, [&](int a, int b) {
        double va = xyz[(size_t)a * 3 + axis], vb = xyz[(size_t)b * 3 + axis];
        return va < vb || (va == vb && a < b);
    });
    int left_parts = (p1 - p0) / 2;
    long long n = hi - lo;
    int n_left = (int)(n * left_parts / (p1 - p0));
    rcb(ids, lo, lo + n_left, p0, p0 + left_parts, xyz, part);
    rcb(ids, lo + n_left, hi, p0 + left_parts, p1, xyz, part);
}

// ---------------------------------------------------------------------------------------
// "kway": recursive graph bisection.  Every split starts from the coordinate bisection above and is refined by
// Fiduccia-Mattheyses passes on the edges inside the subset (the role ParMETIS / PT-Scotch k-way play behind OP2's
// op_partition(..., "KWAY", ...), euler3d.cpp:340-375).  Fully specified so that oracle/plan_oracle.py can restate it:
//   * the left side must keep  n_left - tol <= size <= n_left + tol,  tol = max(1, |S| / 200);
//   * gain(v) = edges of v to the other side - edges of v to its own side (inside S);
//   * a pass moves, one at a time, the unlocked node of highest gain among those whose move keeps the balance (ties:
//     lowest node id), locks it and updates its neighbours' gains; the pass stops
//     after 64 moves without a new best cumulative gain or when no move is allowed, and is rolled back to the best
//     prefix (first occurrence of the maximum);
//   * passes repeat while the best prefix gains something, at most 8 times.
// ---------------------------------------------------------------------------------------
struct Graph {
    std::vector<int> ptr, adj;
};

void fm_refine(const std::vector<int> &S, const Graph &g, std::vector<int> &stamp, int tag, std::vector<char> &side, int n_left)
{
    const int n = (int)S.size(), tol = std::max(1, n / 200);
    static thread_local std::vector<int> gain;
    static thread_local std::vector<char> locked;
    if (gain.size() < stamp.size()) { gain.assign(stamp.size(), 0); locked.assign(stamp.size(), 0); }
    int c0 = 0;
    for (int v : S) c0 += side[v] == 0;
    for (int pass = 0; pass < 8; pass++) {
        std::set<std::pair<int, int>> heap[2];           // (-gain, id): begin() = highest gain, lowest id
        for (int v : S) {
            int gsum = 0;
            for (int j = g.ptr[v]; j < g.ptr[v + 1]; j++) {
                int u = g.adj[j];
                if (stamp[u] != tag) continue;
                gsum += side[u] != side[v] ? 1 : -1;
            }
            gain[v] = gsum;
            locked[v] = 0;
            heap[(int)side[v]].insert({-gsum, v});
        }
        std::vector<int> moves;
        int cur = 0, best = 0, best_len = 0, c0_run = c0;
        while ((int)moves.size() < n) {
            // a move from side 0 shrinks side 0, a move from side 1 grows it
            bool ok0 = !heap[0].empty() && c0_run - 1 >= n_left - tol, ok1 = !heap[1].empty() && c0_run + 1 <= n_left + tol;
            if (!ok0 && !ok1) break;
            int from = ok0 ? 0 : 1;
            if (ok0 && ok1) from = *heap[1].begin() < *heap[0].begin() ? 1 : 0;     // higher gain, then lower id
            auto top = *heap[from].begin();
            heap[from].erase(heap[from].begin());
            const int v = top.second;
            locked[v] = 1;
            cur += gain[v];
            side[v] = (char)(1 - from);
            c0_run += from == 0 ? -1 : 1;
            for (int j = g.ptr[v]; j < g.ptr[v + 1]; j++) {
                int u = g.adj[j];
                if (stamp[u] != tag || locked[u]) continue;
                heap[(int)side[u]].erase({-gain[u], u});
                gain[u] += side[u] == from ? 2 : -2;     // u was on v's old side: the edge is cut now; else it is healed
                heap[(int)side[u]].insert({-gain[u], u});
            }
            moves.push_back(v);
            if (cur > best) { best = cur; best_len = (int)moves.size(); }
            if ((int)moves.size() - best_len >= 64) break;
        }
        for (int i = (int)moves.size() - 1; i >= best_len; i--) side[moves[i]] = (char)(1 - side[moves[i]]);
        c0 = 0;
        for (int v : S) c0 += side[v] == 0;
        if (best <= 0) break;
    }
}

void graph_bisect(std::vector<int> &S, int p0, int p1, const double *xyz, const Graph &g, std::vector<int> &stamp, int &tag,
                  std::vector<char> &side, int *part)
{
    if (p1 - p0 == 1) {
        for (int v : S) part[v] = p0;
        return;
    }
    const int n = (int)S.size();
    double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            double v = xyz[(size_t)S[i] * 3 + d];
            if (i == 0 || v < mn[d]) mn[d] = v;
            if (i == 0 || v > mx[d]) mx[d] = v;
        }
    int axis = 0;
    for (int d = 1; d < 3; d++)
        if (mx[d] - mn[d] > mx[axis] - mn[axis]) axis = d;
    std::sort(S.begin(), S.end(), [&](int a, int b) {
        double va = xyz[(size_t)a * 3 + axis], vb = xyz[(size_t)b * 3 + axis];
        return va < vb || (va == vb && a < b);
    });
    const int left_parts = (p1 - p0) / 2;
    const int n_left = (int)((long long)n * left_parts / (p1 - p0));
    const int my_tag = ++tag;
    for (int i = 0; i < n; i++) { side[S[i]] = i < n_left ? 0 : 1; stamp[S[i]] = my_tag; }
    fm_refine(S, g, stamp, my_tag, side, n_left);
    std::vector<int> L, R;
    for (int v : S) (side[v] == 0 ? L : R).push_back(v);
    std::sort(L.begin(), L.end());
    std::sort(R.begin(), R.end());
    S.clear();
    S.shrink_to_fit();
    graph_bisect(L, p0, p0 + left_parts, xyz, g, stamp, tag, side, part);
    graph_bisect(R, p0 + left_parts, p1, xyz, g, stamp, tag, side, part);
}

}  // namespace

struct mgcfd_local_mesh {
    int n_levels = 0, rank = 0, n_ranks = 1;
    std::vector<LocalLevel> levels;
};

extern "C" {

int mgcfd_partition_rcb(int n_nodes, const double *node_coordinates, int n_parts, int *part_out)
{
    if (n_nodes < 0 || n_parts < 1 || !part_out || (n_nodes > 0 && !node_coordinates)) return MGCFD_ERR_ARG;
    std::vector<int> ids(n_nodes);
    std::iota(ids.begin(), ids.end(), 0);
    rcb(ids, 0, n_nodes, 0, n_parts, node_coordinates, part_out);
    return MGCFD_OK;
}

// op_partition's library / method selection (euler3d.cpp:340-375, config.h:203-240): "geom" (also "inertial") =
// recursive coordinate bisection; "kway" (also "parmetis", "ptscotch", "geomkway") = recursive graph bisection with
// Fiduccia-Mattheyses refinement; "block" = contiguous index ranges; "random" = equal shares of a hashed order.
int mgcfd_partition_graph(int n_nodes, const double *node_coordinates, int n_edges, const int *edge_to_node, int base,
                          int n_parts, const char *method, int *part_out)
{
    if (n_nodes < 0 || n_parts < 1 || !part_out || !method) return MGCFD_ERR_ARG;
    std::string m(method);
    for (char &c : m) c = (char)tolower(c);
    if (m == "geom" || m == "inertial") return mgcfd_partition_rcb(n_nodes, node_coordinates, n_parts, part_out);
    if (m == "block") {
        for (int i = 0; i < n_nodes; i++) part_out[i] = (int)((long long)i * n_parts / std::max(n_nodes, 1));
        return MGCFD_OK;
    }
    if (m == "random") {
        std::vector<std::pair<uint64_t, int>> key(n_nodes);
        for (int i = 0; i < n_nodes; i++) {
            uint64_t z = (uint64_t)i + 0x9e3779b97f4a7c15ull;       // splitmix64
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            key[i] = {z ^ (z >> 31), i};
        }
        std::sort(key.begin(), key.end());
        for (int r = 0; r < n_nodes; r++) part_out[key[r].second] = (int)((long long)r * n_parts / n_nodes);
        return MGCFD_OK;
    }
    if (m != "kway" && m != "parmetis" && m != "ptscotch" && m != "geomkway") return MGCFD_ERR_ARG;
    if ((n_nodes > 0 && !node_coordinates) || (n_edges > 0 && !edge_to_node)) return MGCFD_ERR_ARG;
    Graph g;
    g.ptr.assign(n_nodes + 1, 0);
    for (int e = 0; e < n_edges; e++) {
        int a = edge_to_node[2 * (size_t)e] - base, b = edge_to_node[2 * (size_t)e + 1] - base;
        if (a < 0 || a >= n_nodes || b < 0 || b >= n_nodes) return MGCFD_ERR_ARG;
        g.ptr[a + 1]++; g.ptr[b + 1]++;
    }
    for (int i = 0; i < n_nodes; i++) g.ptr[i + 1] += g.ptr[i];
    g.adj.resize(2 * (size_t)n_edges);
    std::vector<int> fill(g.ptr.begin(), g.ptr.end() - 1);
    for (int e = 0; e < n_edges; e++) {
        int a = edge_to_node[2 * (size_t)e] - base, b = edge_to_node[2 * (size_t)e + 1] - base;
        g.adj[fill[a]++] = b; g.adj[fill[b]++] = a;
    }
    std::vector<int> S(n_nodes), stamp(n_nodes, 0);
    std::iota(S.begin(), S.end(), 0);
    std::vector<char> side(n_nodes, 0);
    int tag = 0;
    graph_bisect(S, 0, n_parts, node_coordinates, g, stamp, tag, side, part_out);
    return MGCFD_OK;
}

int mgcfd_partition_coarse(int n_fine, const int *fine_part, const int *fine_to_coarse, int base, int n_coarse,
                           int n_coarse_edges, const int *coarse_edge_to_node, const double *coarse_coordinates,
                           int *coarse_part_out)
{
    if (!fine_part || !fine_to_coarse || !coarse_part_out) return MGCFD_ERR_ARG;
    for (int c = 0; c < n_coarse; c++) coarse_part_out[c] = -1;
    for (int f = 0; f < n_fine; f++) {            // ascending file index: the first child seen is the lowest
        int c = fine_to_coarse[f] - base;
        if (c < 0 || c >= n_coarse) return MGCFD_ERR_ARG;
        if (coarse_part_out[c] < 0) coarse_part_out[c] = fine_part[f];
    }
    std::vector<int> best(n_coarse, -1);
    std::vector<double> best_d(n_coarse, 0.0);
    auto consider = [&](int orphan, int nb) {
        if (coarse_part_out[orphan] >= 0 || coarse_part_out[nb] < 0) return;
        double d2 = 0.0;
        for (int k = 0; k < 3; k++) {
            double t = coarse_coordinates[(size_t)orphan * 3 + k] - coarse_coordinates[(size_t)nb * 3 + k];
            d2 += t * t;
        }
        if (best[orphan] < 0 || d2 < best_d[orphan] || (d2 == best_d[orphan] && nb < best[orphan])) {
            best[orphan] = nb;
            best_d[orphan] = d2;
        }
    };
    for (int e = 0; e < n_coarse_edges; e++) {
        int a = coarse_edge_to_node[2 * (size_t)e] - base, b = coarse_edge_to_node[2 * (size_t)e + 1] - base;
        consider(a, b);
        consider(b, a);
    }
    for (int c = 0; c < n_coarse; c++)
        if (coarse_part_out[c] < 0) coarse_part_out[c] = best[c] >= 0 ? coarse_part_out[best[c]] : 0;
    // note: `consider` only looks at neighbours whose owner came from children, because orphans are resolved after
    // the scan (coarse_part_out of an orphan stays -1 during the edge loop)
    return MGCFD_OK;
}

int mgcfd_local_mesh_build(int n_levels, const mgcfd_level_host *global_levels, int base, const int *const *part,
                           int rank, int n_ranks, mgcfd_local_mesh **out)
{
    if (!global_levels || !part || !out || n_levels < 1 || rank < 0 || rank >= n_ranks) return MGCFD_ERR_ARG;
    // the deck and the partition vectors are indexed below without further checks: refuse anything that does not fit
    for (int l = 0; l < n_levels; l++) {
        const mgcfd_level_host &G = global_levels[l];
        if (G.n_nodes < 0 || G.n_edges < 0 || G.n_bnd_nodes < 0 || !part[l]) return MGCFD_ERR_ARG;
        if ((G.n_nodes > 0 && !G.node_coordinates) || (G.n_edges > 0 && (!G.edge_to_node || !G.edge_weights)) ||
            (G.n_bnd_nodes > 0 && (!G.bnd_node_to_node || !G.bnd_node_to_group || !G.bnd_node_weights)) ||
            (l + 1 < n_levels && G.n_nodes > 0 && !G.node_to_mg_node))
            return MGCFD_ERR_ARG;
        for (int n = 0; n < G.n_nodes; n++)
            if (part[l][n] < 0 || part[l][n] >= n_ranks) return MGCFD_ERR_ARG;
        for (size_t i = 0; i < 2 * (size_t)G.n_edges; i++)
            if (G.edge_to_node[i] - base < 0 || G.edge_to_node[i] - base >= G.n_nodes) return MGCFD_ERR_ARG;
        for (int i = 0; i < G.n_bnd_nodes; i++)
            if (G.bnd_node_to_node[i] - base < 0 || G.bnd_node_to_node[i] - base >= G.n_nodes) return MGCFD_ERR_ARG;
        if (l + 1 < n_levels)
            for (int f = 0; f < G.n_nodes; f++)
                if (G.node_to_mg_node[f] - base < 0 || G.node_to_mg_node[f] - base >= global_levels[l + 1].n_nodes) return MGCFD_ERR_ARG;
    }
    mgcfd_local_mesh *M = new mgcfd_local_mesh();
    M->n_levels = n_levels; M->rank = rank; M->n_ranks = n_ranks;
    M->levels.resize(n_levels);
    // pass 1: halo membership flags per level: needed_by[q][n] for every rank q would be O(P*N); instead build, per
    // level, the set of (node, rank) "rank q needs node n" pairs restricted to what involves `rank`:
    //   in_halo[l][n]   = 1 if this rank needs remote node n
    //   exports[l]      = pairs (q, n) with n owned here and needed by q
    std::vector<std::vector<char>> in_halo(n_levels);
    std::vector<std::vector<std::pair<int, int>>> exports(n_levels);
    for (int l = 0; l < n_levels; l++) in_halo[l].assign(global_levels[l].n_nodes, 0);
    auto need = [&](int l, int q, int n) {      // rank q reads node n of level l
        int owner = part[l][n];
        if (owner == q) return;
        if (q == rank) in_halo[l][n] = 1;
        else if (owner == rank) exports[l].push_back({q, n});
    };
    for (int l = 0; l < n_levels; l++) {
        const mgcfd_level_host &G = global_levels[l];
        for (int e = 0; e < G.n_edges; e++) {
            int a = G.edge_to_node[2 * (size_t)e] - base, b = G.edge_to_node[2 * (size_t)e + 1] - base;
            int pa = part[l][a], pb = part[l][b];
            if (pa != pb) { need(l, pa, b); need(l, pb, a); }
        }
        if (l + 1 < n_levels) {
            for (int f = 0; f < G.n_nodes; f++) {
                int c = G.node_to_mg_node[f] - base;
                int pf = part[l][f], pc = part[l + 1][c];
                if (pf != pc) { need(l, pc, f); need(l + 1, pf, c); }   // restrict reads the child, prolong the parent
            }
        }
    }
    for (int l = 0; l < n_levels; l++) {
        const mgcfd_level_host &G = global_levels[l];
        LocalLevel &L = M->levels[l];
        // local node list: owned ascending, then halo by (owner, index)
        std::vector<int> halo;
        for (int n = 0; n < G.n_nodes; n++) {
            if (part[l][n] == rank) L.global_node.push_back(n);
            else if (in_halo[l][n]) halo.push_back(n);
        }
        L.n_owned = (int)L.global_node.size();
        std::stable_sort(halo.begin(), halo.end(), [&](int a, int b) { return part[l][a] < part[l][b]; });
        L.import_ptr.push_back(0);
        for (size_t i = 0; i < halo.size(); i++) {
            int q = part[l][halo[i]];
            if (L.nbr_rank.empty() || L.nbr_rank.back() != q) {
                if (!L.nbr_rank.empty()) L.import_ptr.push_back((int)i);
                L.nbr_rank.push_back(q);
            }
        }
        if (!L.nbr_rank.empty()) L.import_ptr.push_back((int)halo.size());
        L.global_node.insert(L.global_node.end(), halo.begin(), halo.end());
        // exports: sort by (rank, node), unique; neighbour set = union of import and export ranks
        auto &ex = exports[l];
        std::sort(ex.begin(), ex.end());
        ex.erase(std::unique(ex.begin(), ex.end()), ex.end());
        std::vector<int> nbrs = L.nbr_rank;
        for (auto &pr : ex) nbrs.push_back(pr.first);
        std::sort(nbrs.begin(), nbrs.end());
        nbrs.erase(std::unique(nbrs.begin(), nbrs.end()), nbrs.end());
        // rebuild import_ptr over the merged neighbour list (a neighbour may export-only or import-only)
        std::vector<int> imp_ptr(nbrs.size() + 1, 0), exp_ptr(nbrs.size() + 1, 0);
        for (size_t k = 0; k < nbrs.size(); k++) {
            int cnt = 0;
            for (int n : halo) cnt += part[l][n] == nbrs[k];
            imp_ptr[k + 1] = imp_ptr[k] + cnt;
        }
        std::vector<int> local_of(G.n_nodes, -1);
        for (size_t i = 0; i < L.global_node.size(); i++) local_of[L.global_node[i]] = (int)i;
        for (size_t k = 0; k < nbrs.size(); k++) {
            for (auto &pr : ex)
                if (pr.first == nbrs[k]) L.export_idx.push_back(local_of[pr.second]);
            exp_ptr[k + 1] = (int)L.export_idx.size();
        }
        L.nbr_rank = nbrs;
        L.import_ptr = imp_ptr;
        L.export_ptr = exp_ptr;
        // coordinates
        L.coords.resize(L.global_node.size() * 3);
        for (size_t i = 0; i < L.global_node.size(); i++)
            for (int d = 0; d < 3; d++) L.coords[i * 3 + d] = G.node_coordinates[(size_t)L.global_node[i] * 3 + d];
        // edges with an owned endpoint, ascending file index
        for (int e = 0; e < G.n_edges; e++) {
            int a = G.edge_to_node[2 * (size_t)e] - base, b = G.edge_to_node[2 * (size_t)e + 1] - base;
            if (part[l][a] != rank && part[l][b] != rank) continue;
            L.global_edge.push_back(e);
            L.e2n.push_back(local_of[a]);
            L.e2n.push_back(local_of[b]);
            for (int d = 0; d < 3; d++) L.ewt.push_back(G.edge_weights[(size_t)e * 3 + d]);
        }
        // boundary entries of owned nodes
        for (int i = 0; i < G.n_bnd_nodes; i++) {
            int n = G.bnd_node_to_node[i] - base;
            if (part[l][n] != rank) continue;
            L.global_bnd.push_back(i);
            L.b2n.push_back(local_of[n]);
            L.bgroup.push_back(G.bnd_node_to_group[i]);
            for (int d = 0; d < 3; d++) L.bwt.push_back(G.bnd_node_weights[(size_t)i * 3 + d]);
        }
    }
    // multigrid maps in local numbering (-1: the parent is not on this rank; only possible for halo nodes)
    for (int l = 0; l + 1 < n_levels; l++) {
        const mgcfd_level_host &G = global_levels[l];
        LocalLevel &L = M->levels[l], &C = M->levels[l + 1];
        std::vector<int> local_coarse(global_levels[l + 1].n_nodes, -1);
        for (size_t i = 0; i < C.global_node.size(); i++) local_coarse[C.global_node[i]] = (int)i;
        L.mg.resize(L.global_node.size());
        for (size_t i = 0; i < L.global_node.size(); i++) L.mg[i] = local_coarse[G.node_to_mg_node[L.global_node[i]] - base];
    }
    for (int l = 0; l < n_levels; l++) {
        LocalLevel &L = M->levels[l];
        mgcfd_level_host &v = L.view;
        v.n_nodes = (int)L.global_node.size();
        v.n_edges = (int)L.global_edge.size();
        v.n_bnd_nodes = (int)L.global_bnd.size();
        v.n_owned_nodes = L.n_owned;
        v.node_coordinates = L.coords.data();
        v.edge_to_node = L.e2n.data();
        v.edge_weights = L.ewt.data();
        v.bnd_node_to_node = L.b2n.data();
        v.bnd_node_to_group = L.bgroup.data();
        v.bnd_node_weights = L.bwt.data();
        v.node_to_mg_node = l + 1 < n_levels ? L.mg.data() : nullptr;
        v.global_node_id = L.global_node.data();
        v.n_neighbours = (int)L.nbr_rank.size();
        v.neighbour_rank = L.nbr_rank.data();
        v.export_ptr = L.export_ptr.data();
        v.export_idx = L.export_idx.data();
        v.import_ptr = L.import_ptr.data();
    }
    *out = M;
    return MGCFD_OK;
}

const mgcfd_level_host *mgcfd_local_mesh_level(const mgcfd_local_mesh *m, int level)
{
    if (!m || level < 0 || level >= m->n_levels) return nullptr;
    return &m->levels[level].view;
}

long long mgcfd_local_mesh_query(const mgcfd_local_mesh *m, int level, const char *what, int *out, long long capacity)
{
    if (!m || level < 0 || level >= m->n_levels || !what) return MGCFD_ERR_ARG;
    const LocalLevel &L = m->levels[level];
    const std::vector<int> *v = nullptr;
    std::string s(what);
    if (s == "global_node") v = &L.global_node;
    else if (s == "global_edge") v = &L.global_edge;
    else if (s == "global_bnd") v = &L.global_bnd;
    else if (s == "neighbour_rank") v = &L.nbr_rank;
    else if (s == "export_ptr") v = &L.export_ptr;
    else if (s == "export_idx") v = &L.export_idx;
    else if (s == "import_ptr") v = &L.import_ptr;
    else if (s == "edge_to_node") v = &L.e2n;
    else if (s == "node_to_mg_node") v = &L.mg;
    else return MGCFD_ERR_ARG;
    if (out) {
        if (capacity < (long long)v->size()) return MGCFD_ERR_ARG;
        std::copy(v->begin(), v->end(), out);
    }
    return (long long)v->size();
}

void mgcfd_local_mesh_free(mgcfd_local_mesh *m) { delete m; }

}  // extern "C"
