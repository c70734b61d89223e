// plan.cpp -- host-side planner: locality renumbering, edge ordering, OP2-style two-level
// colouring, owner-compute chunking.  Deterministic; oracle/plan_oracle.py restates every
// algorithm here independently in numpy and tests compare the index sets bit for bit
// (the reference keeps all of this inside the absent OP2 library: SURVEY.md 4.2, 8c).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "internal.h"

namespace mgcfd {

// ---------------------------------------------------------------------------------------
// Node renumbering: 3-D Hilbert curve over coordinates quantised to 16 bits per axis
// (Skilling's transpose algorithm), ties by file index.  Only owned nodes are reordered;
// halo nodes keep their relative order behind them.
// ---------------------------------------------------------------------------------------
static const int HILBERT_BITS = 16;

static uint64_t hilbert_key(uint32_t x0, uint32_t x1, uint32_t x2)
{
    uint32_t X[3] = {x0, x1, x2};
    const uint32_t M = 1u << (HILBERT_BITS - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        uint32_t P = Q - 1;
        for (int i = 0; i < 3; i++) {
            if (X[i] & Q) {
                X[0] ^= P;
            } else {
                uint32_t t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    for (int i = 1; i < 3; i++) X[i] ^= X[i - 1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    for (int i = 0; i < 3; i++) X[i] ^= t;
    uint64_t key = 0;
    for (int b = HILBERT_BITS - 1; b >= 0; b--)
        for (int i = 0; i < 3; i++) key = (key << 1) | ((X[i] >> b) & 1u);
    return key;
}

void plan_renumber(LevelHost &L, bool renumber)
{
    const int n = L.n_nodes, no = L.n_owned;
    L.new_of_old.resize(n);
    L.old_of_new.resize(n);
    std::iota(L.old_of_new.begin(), L.old_of_new.end(), 0);
    if (renumber && no > 1) {
        double lo[3], hi[3];
        for (int d = 0; d < 3; d++) lo[d] = hi[d] = L.coords[d];
        for (int i = 0; i < no; i++)
            for (int d = 0; d < 3; d++) {
                double v = L.coords[(size_t)i * 3 + d];
                if (v < lo[d]) lo[d] = v;
                if (v > hi[d]) hi[d] = v;
            }
        double span = 0.0;
        for (int d = 0; d < 3; d++) span = std::max(span, hi[d] - lo[d]);
        if (!(span > 0.0)) span = 1.0;
        const double scale = (double)((1u << HILBERT_BITS) - 1);
        std::vector<std::pair<uint64_t, int>> keyed(no);
        for (int i = 0; i < no; i++) {
            uint32_t q[3];
            for (int d = 0; d < 3; d++) {
                double f = std::floor((L.coords[(size_t)i * 3 + d] - lo[d]) / span * scale + 0.5);
                q[d] = (uint32_t)f;
            }
            keyed[i] = {hilbert_key(q[0], q[1], q[2]), i};
        }
        std::sort(keyed.begin(), keyed.end());
        for (int i = 0; i < no; i++) L.old_of_new[i] = keyed[i].second;
    }
    for (int i = 0; i < n; i++) L.new_of_old[L.old_of_new[i]] = i;
}

// ---------------------------------------------------------------------------------------
// Edge ordering: by (min internal endpoint, max internal endpoint, file index)
// ---------------------------------------------------------------------------------------
void plan_sort_edges(LevelHost &L)
{
    const int E = L.n_edges;
    std::vector<std::pair<uint64_t, int>> keyed(E);
    for (int e = 0; e < E; e++) {
        uint32_t a = (uint32_t)L.new_of_old[L.e2n[2 * (size_t)e]];
        uint32_t b = (uint32_t)L.new_of_old[L.e2n[2 * (size_t)e + 1]];
        uint32_t lo = std::min(a, b), hi = std::max(a, b);
        keyed[e] = {((uint64_t)lo << 32) | hi, e};
    }
    std::sort(keyed.begin(), keyed.end());
    L.sorted.order.resize(E);
    for (int e = 0; e < E; e++) L.sorted.order[e] = keyed[e].second;
    L.have_sorted = true;
}

static inline int lowest_zero_bit(uint64_t m)
{
    return m == ~0ull ? -1 : __builtin_ctzll(~m);
}

// ---------------------------------------------------------------------------------------
// OP2-style two-level colouring (first fit over per-node bit masks):
//   level 1: within a block of `block_edges` consecutive sorted edges, edges in sorted order
//            take the lowest colour unused by both endpoints;
//   level 2: blocks in order take the lowest colour unused by any of their nodes.
// ---------------------------------------------------------------------------------------
void plan_colour(LevelHost &L, int block_edges)
{
    if (!L.have_sorted) plan_sort_edges(L);
    ColourPlanHost &C = L.colour;
    const int E = L.n_edges, BS = block_edges;
    C = ColourPlanHost();
    C.block_edges = BS;
    C.n_blocks = (E + BS - 1) / BS;
    C.thread_colour.assign(E, 0);
    C.block_colour.assign(C.n_blocks, 0);
    C.block_ncol.assign(C.n_blocks, 0);

    std::vector<uint64_t> node_thread_mask(L.n_nodes, 0), node_block_mask(L.n_nodes, 0);
    std::vector<std::vector<int>> blk_nodes(C.n_blocks);
    for (int k = 0; k < C.n_blocks; k++) {
        int lo = k * BS, hi = std::min(E, lo + BS);
        std::vector<int> &nodes = blk_nodes[k];
        nodes.reserve(2 * (hi - lo));
        for (int i = lo; i < hi; i++) {
            int e = L.sorted.order[i];
            nodes.push_back(L.new_of_old[L.e2n[2 * (size_t)e]]);
            nodes.push_back(L.new_of_old[L.e2n[2 * (size_t)e + 1]]);
        }
        std::sort(nodes.begin(), nodes.end());
        nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
        for (int v : nodes) node_thread_mask[v] = 0;
        int ncol = 0;
        for (int i = lo; i < hi; i++) {
            int e = L.sorted.order[i];
            int a = L.new_of_old[L.e2n[2 * (size_t)e]], b = L.new_of_old[L.e2n[2 * (size_t)e + 1]];
            int c = lowest_zero_bit(node_thread_mask[a] | node_thread_mask[b]);
            if (c < 0) c = 63;   // > 64 incident edges of one node inside a block: not coloured properly (checked by caller)
            node_thread_mask[a] |= 1ull << c;
            node_thread_mask[b] |= 1ull << c;
            C.thread_colour[i] = c;
            ncol = std::max(ncol, c + 1);
        }
        C.block_ncol[k] = ncol;
        uint64_t m = 0;
        for (int v : nodes) m |= node_block_mask[v];
        int bc = lowest_zero_bit(m);
        if (bc < 0) bc = 63;
        for (int v : nodes) node_block_mask[v] |= 1ull << bc;
        C.block_colour[k] = bc;
        C.n_block_colours = std::max(C.n_block_colours, bc + 1);
        C.max_nodes = std::max(C.max_nodes, (int)nodes.size());
    }

    // execution layout: blocks by (block colour, id); edges in a block by (thread colour, sorted position)
    C.exec_block.resize(C.n_blocks);
    std::iota(C.exec_block.begin(), C.exec_block.end(), 0);
    std::stable_sort(C.exec_block.begin(), C.exec_block.end(),
                     [&](int x, int y) { return C.block_colour[x] < C.block_colour[y]; });
    C.colour_start.assign(C.n_block_colours + 1, 0);
    for (int k = 0; k < C.n_blocks; k++) C.colour_start[C.block_colour[k] + 1]++;
    for (int c = 0; c < C.n_block_colours; c++) C.colour_start[c + 1] += C.colour_start[c];
    C.node_off.assign(C.n_blocks + 1, 0);
    C.node_gid.clear();
    C.lab.resize(E);
    C.exec_edge.resize(E);
    C.ecol.resize(E);
    size_t pos = 0;
    std::vector<int> idx;
    for (int s = 0; s < C.n_blocks; s++) {
        int k = C.exec_block[s];
        int lo = k * BS, hi = std::min(E, lo + BS);
        const std::vector<int> &nodes = blk_nodes[k];
        C.node_off[s] = (int)C.node_gid.size();
        C.node_gid.insert(C.node_gid.end(), nodes.begin(), nodes.end());
        idx.resize(hi - lo);
        std::iota(idx.begin(), idx.end(), lo);
        std::stable_sort(idx.begin(), idx.end(),
                         [&](int x, int y) { return C.thread_colour[x] < C.thread_colour[y]; });
        // exec edge positions of block slot s are [s_lo, s_lo + hi-lo): slots keep the block's size, and only the
        // last block (by id) can be short, so position = running offset
        for (int i : idx) {
            int e = L.sorted.order[i];
            int a = L.new_of_old[L.e2n[2 * (size_t)e]], b = L.new_of_old[L.e2n[2 * (size_t)e + 1]];
            uint32_t la = (uint32_t)(std::lower_bound(nodes.begin(), nodes.end(), a) - nodes.begin());
            uint32_t lb = (uint32_t)(std::lower_bound(nodes.begin(), nodes.end(), b) - nodes.begin());
            C.lab[pos] = la | (lb << 16);
            C.exec_edge[pos] = e;
            C.ecol[pos] = (unsigned char)C.thread_colour[i];
            pos++;
        }
    }
    C.node_off[C.n_blocks] = (int)C.node_gid.size();
    L.have_colour = true;
}

// ---------------------------------------------------------------------------------------
// Owner-compute chunks: consecutive internal nodes [node0[k], node0[k+1]) own their fluxes; a
// chunk holds every edge with at least one owned endpoint (cut edges appear in both chunks).
// Greedy growth under caps on owned nodes, local nodes (owned + halo) and edges; the check
// before admitting node v is conservative: it assumes all deg(v) edges and neighbours are new.
// A chunk is only closed when it owns an even number of nodes (the caps on local nodes and edges
// are therefore soft by one node), so every chunk starts on an even node index.
// ---------------------------------------------------------------------------------------
bool plan_owner(LevelHost &L, int max_own, int max_loc, int max_edges, std::string &err)
{
    OwnerPlanHost &O = L.owner;
    O = OwnerPlanHost();
    const int n = L.n_nodes, no = L.n_owned, E = L.n_edges;
    // adjacency over internal ids, incident edges in ascending file order
    std::vector<int> adj_ptr(n + 1, 0);
    for (int e = 0; e < E; e++) {
        adj_ptr[L.new_of_old[L.e2n[2 * (size_t)e]] + 1]++;
        adj_ptr[L.new_of_old[L.e2n[2 * (size_t)e + 1]] + 1]++;
    }
    for (int i = 0; i < n; i++) adj_ptr[i + 1] += adj_ptr[i];
    std::vector<int> adj_edge(2 * (size_t)E), fill(adj_ptr.begin(), adj_ptr.end() - 1);
    for (int e = 0; e < E; e++) {
        adj_edge[fill[L.new_of_old[L.e2n[2 * (size_t)e]]]++] = e;
        adj_edge[fill[L.new_of_old[L.e2n[2 * (size_t)e + 1]]]++] = e;
    }
    auto other = [&](int e, int v) {
        int a = L.new_of_old[L.e2n[2 * (size_t)e]], b = L.new_of_old[L.e2n[2 * (size_t)e + 1]];
        return a == v ? b : a;
    };

    // pass 1: chunk boundaries
    std::vector<int> halo_stamp(n, -1);
    O.node0.push_back(0);
    int v = 0;
    while (v < no) {
        int k = (int)O.node0.size() - 1, start = v;
        int n_own = 0, n_halo = 0, n_edge = 0;
        while (v < no) {
            int deg = adj_ptr[v + 1] - adj_ptr[v];
            // chunks end on even node counts (the owned run is then a 16-byte aligned bulk-copy source)
            if (n_own > 0 && n_own % 2 == 0 &&
                (n_own + 1 > max_own || n_own + n_halo + 1 + deg > max_loc || n_edge + deg > max_edges))
                break;
            if (halo_stamp[v] == k) n_halo--;      // v was a halo node of this chunk until now
            for (int j = adj_ptr[v]; j < adj_ptr[v + 1]; j++) {
                int u = other(adj_edge[j], v);
                if (u >= start && u < v) continue;  // edge already counted from u's side
                n_edge++;
                if (halo_stamp[u] != k) { halo_stamp[u] = k; n_halo++; }
            }
            n_own++;
            v++;
        }
        O.node0.push_back(v);
    }
    O.n_chunks = (int)O.node0.size() - 1;

    // pass 2: per-chunk lists
    std::vector<int> edge_stamp(E, -1), halo_local(n, -1);
    O.halo_off.assign(1, 0);
    O.edge_off.assign(1, 0);
    O.rowptr_off.assign(1, 0);
    O.csr_off.assign(1, 0);
    long long blob = 0;
    std::vector<int> halo, local_edge_of(E, -1);
    for (int k = 0; k < O.n_chunks; k++) {
        int s = O.node0[k], t = O.node0[k + 1], n_own = t - s;
        halo.clear();
        size_t e_begin = O.edge_file.size();
        for (int u = s; u < t; u++)
            for (int j = adj_ptr[u]; j < adj_ptr[u + 1]; j++) {
                int e = adj_edge[j];
                if (edge_stamp[e] == k) continue;
                edge_stamp[e] = k;
                local_edge_of[e] = (int)(O.edge_file.size() - e_begin);
                O.edge_file.push_back(e);
                int w = other(e, u);
                if (w < s || w >= t) halo.push_back(w);
            }
        std::sort(halo.begin(), halo.end());
        halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
        for (size_t h = 0; h < halo.size(); h++) halo_local[halo[h]] = n_own + (int)h;
        int ne = (int)(O.edge_file.size() - e_begin);
        int nloc = n_own + (int)halo.size();
        if (nloc > 65535 || ne > 32767) {
            err = "owner chunk exceeds 16-bit local index limits";
            return false;
        }
        for (int i = 0; i < ne; i++) {
            int e = O.edge_file[e_begin + i];
            int a = L.new_of_old[L.e2n[2 * (size_t)e]], b = L.new_of_old[L.e2n[2 * (size_t)e + 1]];
            uint32_t la = (a >= s && a < t) ? (uint32_t)(a - s) : (uint32_t)halo_local[a];
            uint32_t lb = (b >= s && b < t) ? (uint32_t)(b - s) : (uint32_t)halo_local[b];
            O.lab.push_back(la | (lb << 16));
        }
        // CSR of owned incidences, ascending file edge id per node (= OP2-seq increment order)
        int ninc = 0;
        for (int u = s; u < t; u++) {
            O.rowptr.push_back((uint16_t)ninc);
            for (int j = adj_ptr[u]; j < adj_ptr[u + 1]; j++) {
                int e = adj_edge[j];
                bool is_b = L.new_of_old[L.e2n[2 * (size_t)e + 1]] == u;   // self edges are rejected at decl time
                O.csr.push_back((uint16_t)(local_edge_of[e] | (is_b ? 0x8000 : 0)));
                ninc++;
            }
        }
        O.rowptr.push_back((uint16_t)ninc);
        if (ninc > 65535) { err = "owner chunk has more than 65535 incidences"; return false; }
        O.halo_gid.insert(O.halo_gid.end(), halo.begin(), halo.end());
        O.halo_off.push_back((int)O.halo_gid.size());
        O.edge_off.push_back((int)O.edge_file.size());
        O.rowptr_off.push_back((int)O.rowptr.size());
        O.csr_off.push_back((int)O.csr.size());
        O.n_edges.push_back(ne);
        O.n_inc.push_back(ninc);
        O.blob_off.push_back(blob);
        int e_pad = (ne + 3) & ~3;
        long long bytes = (long long)e_pad * (4 * 8 + 4);
        bytes += (((long long)(n_own + 1) * 2 + 15) & ~15ll);
        bytes += (((long long)ninc * 2 + 15) & ~15ll);
        blob += bytes;
        O.max_blob = std::max(O.max_blob, (int)bytes);
        O.max_loc = std::max(O.max_loc, nloc);
        O.max_edges = std::max(O.max_edges, e_pad);
        O.max_own = std::max(O.max_own, n_own);
        O.max_inc = std::max(O.max_inc, ninc);
        O.total_edges += ne;
    }
    O.blob_off.push_back(blob);
    L.have_owner = true;
    return true;
}

}  // namespace mgcfd
