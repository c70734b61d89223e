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
