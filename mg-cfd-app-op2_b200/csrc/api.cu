// api.cu -- the C-ABI of libmgcfd_b200: context, declarations, planning, the op_par_loop call
// sites of euler3d.cpp, fetch, timers.  See include/mgcfd_b200.h for the contract.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <thread>

#include "internal.h"

using namespace mgcfd;

static std::string g_create_error;

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
            return MGCFD_ERR_CUDA;                                                                       \
        }                                                                                                \
    } while (0)

// (a null context has no error slot: the text goes where mgcfd_last_error(NULL) reads it)
static inline void set_error(mgcfd_ctx *ctx, std::string msg)
{
    if (ctx) ctx->err = std::move(msg);
    else g_create_error = std::move(msg);
}

#define REQUIRE(cond, msg)                                                                               \
    do {                                                                                                 \
        if (!(cond)) {                                                                                   \
            set_error(ctx, (msg));                                                                       \
            return MGCFD_ERR_ARG;                                                                        \
        }                                                                                                \
    } while (0)

#define CHECK_LEVEL(l) REQUIRE(ctx && (l) >= 0 && (l) < ctx->n_levels, "level out of range")
#define CHECK_PLANNED()                                                                                  \
    do {                                                                                                 \
        REQUIRE(ctx->planned, "mgcfd_plan() has not been called");                                       \
        if (ctx->device < 0) {                                                                           \
            ctx->err = "planning-only context (device -1): compute entry points need a CUDA device";     \
            return MGCFD_ERR_NODEVICE;                                                                   \
        }                                                                                                \
    } while (0)

int mgcfd::api_check_launch(mgcfd_ctx *ctx, const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
        return MGCFD_ERR_CUDA;
    }
    return MGCFD_OK;
}

template <typename T>
static int dev_alloc(mgcfd_ctx *ctx, T **p, size_t count, bool zero = true)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) count = 1;
    CK(cudaMalloc((void **)p, count * sizeof(T)));
    if (zero) CK(cudaMemsetAsync(*p, 0, count * sizeof(T), ctx->stream));
    return MGCFD_OK;
}

template <typename T>
static int dev_upload(mgcfd_ctx *ctx, T **p, const std::vector<T> &h)
{
    int rc = dev_alloc(ctx, p, h.size(), false);
    if (rc) return rc;
    if (!h.empty()) {
        CK(cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));   // h may be a temporary
    }
    return MGCFD_OK;
}

DevConsts mgcfd::api_dev_consts(const mgcfd_ctx *ctx)
{
    DevConsts c;
    c.smoothing = ctx->consts.smoothing_coefficient;
    for (int i = 0; i < 5; i++) c.ff_variable[i] = ctx->consts.ff_variable[i];
    for (int d = 0; d < 3; d++) {
        c.ff_fc[0][d] = 0.0;
        c.ff_fc[1][d] = ctx->consts.ff_flux_contribution_momentum_x[d];
        c.ff_fc[2][d] = ctx->consts.ff_flux_contribution_momentum_y[d];
        c.ff_fc[3][d] = ctx->consts.ff_flux_contribution_momentum_z[d];
        c.ff_fc[4][d] = ctx->consts.ff_flux_contribution_density_energy[d];
    }
    return c;
}

// ------------------------------------------------------------------------------------------
// timers
// ------------------------------------------------------------------------------------------
void mgcfd::timers_collect(mgcfd_ctx *ctx)
{
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->timers) {
        LoopTimer &t = kv.second;
        for (size_t i = 0; i < t.pending.size(); i++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, t.pending[i].first, t.pending[i].second);
            t.ms += ms;
            t.calls++;
            t.elements += t.pending_elems[i];
            ctx->event_pool.push_back(t.pending[i].first);
            ctx->event_pool.push_back(t.pending[i].second);
        }
        t.pending.clear();
        t.pending_elems.clear();
    }
}

// ------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------
extern "C" {

const char *mgcfd_version(void) { return "mgcfd_b200 0.1 (sm_100a)"; }

void mgcfd_default_options(mgcfd_options *opt)
{
    memset(opt, 0, sizeof(*opt));
    opt->flux_variant = MGCFD_FLUX_OWNER;
    opt->renumber = 1;
    opt->owner_chunk_nodes = 64;
    opt->colour_block_edges = 256;
    opt->exact_arith = 0;
}

const char *mgcfd_last_error(const mgcfd_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mgcfd_create(mgcfd_ctx **out, int device, int n_levels, const mgcfd_options *opt)
{
    if (!out || n_levels < 1 || n_levels > 64) {
        g_create_error = "mgcfd_create: bad arguments";
        return MGCFD_ERR_ARG;
    }
    *out = nullptr;
    if (device == -1) {
        // planning-only context: host planner + plan_query, no device, no compute (index-set tests on CPU boxes)
        mgcfd_ctx *ctx = new mgcfd_ctx();
        ctx->device = -1;
        ctx->n_levels = n_levels;
        if (opt) ctx->opt = *opt; else mgcfd_default_options(&ctx->opt);
        if (ctx->opt.owner_chunk_nodes <= 0) ctx->opt.owner_chunk_nodes = 256;
        ctx->opt.owner_chunk_nodes = std::max(2, ctx->opt.owner_chunk_nodes & ~1);
        if (ctx->opt.colour_block_edges <= 0 || ctx->opt.colour_block_edges > 256) ctx->opt.colour_block_edges = 256;
        ctx->H.resize(n_levels);
        ctx->D.resize(n_levels);
        ctx->halo.resize(n_levels);
        ctx->n_ranks = std::max(1, ctx->opt.n_ranks);
        ctx->rank = ctx->opt.rank;
        *out = ctx;
        return MGCFD_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count=0") +
                         "); libmgcfd_b200 has no CPU fallback";
        return MGCFD_ERR_NODEVICE;
    }
    if (device < 0 || device >= count) {
        g_create_error = "mgcfd_create: device index out of range";
        return MGCFD_ERR_ARG;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return MGCFD_ERR_CUDA;
    }
    mgcfd_ctx *ctx = new mgcfd_ctx();
    ctx->device = device;
    ctx->n_levels = n_levels;
    if (opt) ctx->opt = *opt; else mgcfd_default_options(&ctx->opt);
    if (ctx->opt.owner_chunk_nodes <= 0) ctx->opt.owner_chunk_nodes = 256;
    ctx->opt.owner_chunk_nodes = std::max(2, ctx->opt.owner_chunk_nodes & ~1);   // chunks own an even number of nodes
    if (ctx->opt.colour_block_edges <= 0) ctx->opt.colour_block_edges = 256;
    if (ctx->opt.colour_block_edges > 256) ctx->opt.colour_block_edges = 256;
    ctx->H.resize(n_levels);
    ctx->D.resize(n_levels);
    ctx->halo.resize(n_levels);
    ctx->n_ranks = std::max(1, ctx->opt.n_ranks);
    ctx->rank = ctx->opt.rank;
    if (ctx->rank < 0 || ctx->rank >= ctx->n_ranks) { g_create_error = "mgcfd_create: rank out of range"; delete ctx; return MGCFD_ERR_ARG; }
    auto fail = [&](const char *what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        delete ctx;
        return MGCFD_ERR_CUDA;
    };
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaMalloc((void **)&ctx->d_min_dt, sizeof(double) * n_levels)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&ctx->d_rms, sizeof(double))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&ctx->d_min_enc, sizeof(unsigned long long) * 2 * n_levels)) != cudaSuccess) return fail("cudaMalloc", e);
    k_reset_min_slots(ctx->stream, 2 * n_levels, ctx->d_min_enc);
    if ((e = cudaMalloc((void **)&ctx->d_flags, sizeof(int) * 4)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMemset(ctx->d_flags, 0, sizeof(int) * 4)) != cudaSuccess) return fail("cudaMemset", e);
    if ((e = cudaMallocHost((void **)&ctx->h_pinned, sizeof(double) * 16)) != cudaSuccess) return fail("cudaMallocHost", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_pack, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_k1, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_prod, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    std::string ce = flux_configure();
    if (!ce.empty()) {
        g_create_error = ce;
        delete ctx;
        return MGCFD_ERR_CUDA;
    }
    *out = ctx;
    return MGCFD_OK;
}

static void free_level(LevelDev &d)
{
    if (d.in_arena) d.var = d.var_alt = d.res = nullptr;     // owned by the p2p arena
    void *ptrs[] = {d.var_alt, d.bnd_ptr, d.var, d.old, d.res, d.flux, d.dummy_flux, d.vol, d.sf, d.coords, d.up_count, d.mg, d.child_ptr,
                    d.child_idx, d.bu_node, d.bu_ptr, d.b_group, d.b_wt, d.cbrt_vol, d.perm, d.atomic.nodes, d.atomic.w,
                    d.colour.blk_edge0, d.colour.blk_node0, d.colour.blk_ncol, d.colour.node_gid, d.colour.lab,
                    d.colour.ecol, d.colour.w, d.owner.desc, d.owner.halo_gid, d.owner.blob, d.owner.xtab, d.gather.desc, d.gather.halo_gid,
                    d.gather.row_node, d.gather.row_deg, d.gather.ent, d.gather.w0, d.gather.w1, d.gather.w2, d.gather.g,
                    d.emit.desc, d.emit.halo_gid, d.emit.row_node, d.emit.row_cnt, d.emit.blob};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    d = LevelDev();
}

void mgcfd_destroy(mgcfd_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->device < 0) { delete ctx; return; }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cycle_drop_graphs(ctx);
    for (auto &d : ctx->D) free_level(d);
    for (auto &kv : ctx->timers)
        for (auto &p : kv.second.pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->comm_stream) { cudaStreamSynchronize(ctx->comm_stream); cudaStreamDestroy(ctx->comm_stream); }
    for (cudaEvent_t e : {ctx->ev_pack, ctx->ev_done, ctx->ev_k1, ctx->ev_prod, ctx->ev_ready})
        if (e) cudaEventDestroy(e);
    for (auto &h : ctx->halo) {
        if (h.d_export_idx) cudaFree(h.d_export_idx);
        if (h.sendbuf) cudaFree(h.sendbuf);
        if (h.d_chunk_list) cudaFree(h.d_chunk_list);
        if (h.d_xp_base) cudaFree(h.d_xp_base);
        if (h.d_xp_ptr) cudaFree(h.d_xp_ptr);
        if (h.d_xp_ent) cudaFree(h.d_xp_ent);
        if (h.d_xn_ptr) cudaFree(h.d_xn_ptr);
        if (h.d_xn_ent) cudaFree(h.d_xn_ent);
    }
    if (ctx->d_min_dt) cudaFree(ctx->d_min_dt);
    if (ctx->d_rms) cudaFree(ctx->d_rms);
    if (ctx->d_min_enc) cudaFree(ctx->d_min_enc);
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (int r = 0; r < mgcfd::P2P_MAX_RANKS; r++)
        if (ctx->p2p.ipc && ctx->p2p.peer_base[r] && r != ctx->rank) cudaIpcCloseMemHandle(ctx->p2p.peer_base[r]);
    if (ctx->p2p.arena_owner && ctx->p2p.arena) cudaFree(ctx->p2p.arena);
    if (ctx->p2p.d_counters) cudaFree(ctx->p2p.d_counters);
    if (ctx->p2p.d_done) cudaFree(ctx->p2p.d_done);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    for (double *p : ctx->io_stage)
        if (p) cudaFree(p);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ------------------------------------------------------------------------------------------
// declarations
// ------------------------------------------------------------------------------------------
void mgcfd_compute_farfield_consts(mgcfd_consts *c)
{
    // euler3d.cpp:47 and :157-189, with the flux contributions of inlined_funcs.h:70-96
    const double gamma = 1.4, pi = 3.1415926535897931, ff_mach = 1.2, deg_aoa = 0.0;
    memset(c, 0, sizeof(*c));
    c->smoothing_coefficient = double(0.2f);
    const double aoa = double(pi / 180.0) * double(deg_aoa);
    c->ff_variable[0] = 1.4;
    double ff_p = 1.0;
    double ff_c = sqrt(gamma * ff_p / c->ff_variable[0]);
    double ff_speed = ff_mach * ff_c;
    double v[3] = {ff_speed * cos(aoa), ff_speed * sin(aoa), 0.0};
    for (int d = 0; d < 3; d++) c->ff_variable[1 + d] = c->ff_variable[0] * v[d];
    c->ff_variable[4] = c->ff_variable[0] * (0.5 * (ff_speed * ff_speed)) + (ff_p / (gamma - 1.0));
    const double *m = &c->ff_variable[1];
    double *fx = c->ff_flux_contribution_momentum_x, *fy = c->ff_flux_contribution_momentum_y,
           *fz = c->ff_flux_contribution_momentum_z, *fe = c->ff_flux_contribution_density_energy;
    fx[0] = v[0] * m[0] + ff_p; fx[1] = v[0] * m[1];        fx[2] = v[0] * m[2];
    fy[0] = fx[1];              fy[1] = v[1] * m[1] + ff_p; fy[2] = v[1] * m[2];
    fz[0] = fx[2];              fz[1] = fy[2];              fz[2] = v[2] * m[2] + ff_p;
    double ep = c->ff_variable[4] + ff_p;
    fe[0] = v[0] * ep; fe[1] = v[1] * ep; fe[2] = v[2] * ep;
}

int mgcfd_decl_consts(mgcfd_ctx *ctx, const mgcfd_consts *c)
{
    REQUIRE(ctx && c, "null argument");
    ctx->consts = *c;
    ctx->have_consts = true;
    for (auto &d : ctx->D) d.atomic.valid = d.colour.valid = d.owner.valid = d.gather.valid = d.emit.valid = false;   // g depends on smoothing
    return MGCFD_OK;
}

int mgcfd_decl_level(mgcfd_ctx *ctx, int level, const mgcfd_level_host *lv, int base)
{
    CHECK_LEVEL(level);
    REQUIRE(lv, "null level");
    REQUIRE(!ctx->planned, "levels must be declared before mgcfd_plan()");
    REQUIRE(lv->n_nodes >= 0 && lv->n_edges >= 0 && lv->n_bnd_nodes >= 0, "negative set size");
    REQUIRE(lv->n_owned_nodes >= 0 && lv->n_owned_nodes <= lv->n_nodes, "n_owned_nodes out of range");
    LevelHost &L = ctx->H[level];
    L = LevelHost();
    L.n_nodes = lv->n_nodes; L.n_edges = lv->n_edges; L.n_bnd = lv->n_bnd_nodes; L.n_owned = lv->n_owned_nodes;
    const size_t n = L.n_nodes, E = L.n_edges, B = L.n_bnd;
    REQUIRE(n == 0 || lv->node_coordinates, "null node_coordinates");
    REQUIRE(E == 0 || (lv->edge_to_node && lv->edge_weights), "null edge-->node / edge_weights");
    REQUIRE(B == 0 || (lv->bnd_node_to_node && lv->bnd_node_to_group && lv->bnd_node_weights), "null boundary dataset");
    L.coords.assign(lv->node_coordinates, lv->node_coordinates + n * 3);
    L.ewt.assign(lv->edge_weights, lv->edge_weights + E * 3);
    L.bwt.assign(lv->bnd_node_weights, lv->bnd_node_weights + B * 3);
    L.e2n.resize(E * 2);
    for (size_t i = 0; i < E * 2; i++) {
        int v = lv->edge_to_node[i] - base;
        REQUIRE(v >= 0 && v < L.n_nodes, "edge-->node entry out of range (check base_array_index)");
        L.e2n[i] = v;
    }
    for (size_t e = 0; e < E; e++) REQUIRE(L.e2n[2 * e] != L.e2n[2 * e + 1], "self edge in edge-->node");
    L.b2n.resize(B);
    L.bgroup.assign(lv->bnd_node_to_group, lv->bnd_node_to_group + B);
    for (size_t i = 0; i < B; i++) {
        int v = lv->bnd_node_to_node[i] - base;
        REQUIRE(v >= 0 && v < L.n_nodes, "bnd_node-->node entry out of range");
        L.b2n[i] = v;
    }
    if (lv->node_to_mg_node) {
        REQUIRE(level + 1 < ctx->n_levels, "node-->mg_node given on the coarsest level");
        L.mg.resize(n);
        for (size_t i = 0; i < n; i++) L.mg[i] = lv->node_to_mg_node[i] - base;   // range-checked in mgcfd_plan
    }
    if (lv->global_node_id) L.global_node.assign(lv->global_node_id, lv->global_node_id + n);
    if (lv->n_neighbours > 0) {
        REQUIRE(ctx->n_ranks > 1, "halo lists given to a context created with n_ranks == 1");
        REQUIRE(lv->neighbour_rank && lv->export_ptr && lv->import_ptr, "null halo list");
        const int nn = lv->n_neighbours;
        L.nbr_rank.assign(lv->neighbour_rank, lv->neighbour_rank + nn);
        L.export_ptr.assign(lv->export_ptr, lv->export_ptr + nn + 1);
        L.import_ptr.assign(lv->import_ptr, lv->import_ptr + nn + 1);
        REQUIRE(L.import_ptr[nn] == L.n_nodes - L.n_owned, "import lists do not cover the halo range");
        REQUIRE(L.export_ptr[nn] == 0 || lv->export_idx, "null export list");
        L.export_idx.assign(lv->export_idx, lv->export_idx + L.export_ptr[nn]);
        for (int q : L.nbr_rank) REQUIRE(q >= 0 && q < ctx->n_ranks && q != ctx->rank, "neighbour rank out of range");
        for (int v : L.export_idx) REQUIRE(v >= 0 && v < L.n_owned, "export index is not an owned node");
    } else {
        REQUIRE(L.n_owned == L.n_nodes, "halo nodes without import lists");
    }
    return MGCFD_OK;
}

// boundary entries grouped by unique internal node, ascending file order inside a node
static int upload_bnd(mgcfd_ctx *ctx, int level)
{
    LevelHost &L = ctx->H[level];
    LevelDev &D = ctx->D[level];
    std::vector<int> idx;
    for (int i = 0; i < L.n_bnd; i++)
        if (L.new_of_old[L.b2n[i]] < L.n_owned) idx.push_back(i);
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return L.new_of_old[L.b2n[x]] < L.new_of_old[L.b2n[y]]; });
    std::vector<int> bu_node, bu_ptr, group(idx.size());
    std::vector<double> wt(idx.size() * 3);
    for (size_t k = 0; k < idx.size(); k++) {
        int i = idx[k], node = L.new_of_old[L.b2n[i]];
        if (bu_node.empty() || bu_node.back() != node) { bu_node.push_back(node); bu_ptr.push_back((int)k); }
        group[k] = L.bgroup[i];
        for (int d = 0; d < 3; d++) wt[k * 3 + d] = L.bwt[(size_t)i * 3 + d];
    }
    bu_ptr.push_back((int)idx.size());
    D.n_bnd_unique = (int)bu_node.size();
    std::vector<int> node_ptr(L.n_owned + 1, 0);      // the same grouping indexed by owned node (fused stage)
    for (size_t t = 0; t < bu_node.size(); t++) node_ptr[bu_node[t] + 1] = bu_ptr[t + 1] - bu_ptr[t];
    for (int i = 0; i < L.n_owned; i++) node_ptr[i + 1] += node_ptr[i];
    {
        int rc0 = dev_upload(ctx, &D.bnd_ptr, node_ptr);
        if (rc0) return rc0;
    }
    L.bnd_node_ptr = node_ptr;
    L.bnd_group_sorted = group;
    L.bnd_wt_sorted = wt;
    D.owner.valid = D.emit.valid = false;              // chunk descriptors carry a has-boundary flag
    cycle_drop_graphs(ctx);                            // captured graphs hold the pointers that are reallocated below
    int rc;
    if ((rc = dev_upload(ctx, &D.bu_node, bu_node))) return rc;
    if ((rc = dev_upload(ctx, &D.bu_ptr, bu_ptr))) return rc;
    if ((rc = dev_upload(ctx, &D.b_group, group))) return rc;
    if ((rc = dev_upload(ctx, &D.b_wt, wt))) return rc;
    return MGCFD_OK;
}

// ---- host <-> device transfers of node dats in FILE order -------------------------------------------
// The copy itself is one DMA between (pinned) host memory and a device staging buffer; the permutation
// between file order and internal order runs on the device.  Caller buffers that are already pinned
// (mgcfd_host_alloc / cudaHostRegister) are used directly, pageable ones go through a pinned bounce buffer.
static bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

static int ensure_staging(mgcfd_ctx *ctx, size_t bytes, bool need_pinned)
{
    if (ctx->d_stage_bytes < bytes) {
        if (ctx->d_stage) cudaFree(ctx->d_stage);
    for (double *p : ctx->io_stage)
        if (p) cudaFree(p);
        ctx->d_stage = nullptr; ctx->d_stage_bytes = 0;
        CK(cudaMalloc(&ctx->d_stage, bytes));
        ctx->d_stage_bytes = bytes;
    }
    if (need_pinned && ctx->h_stage_bytes < bytes) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr; ctx->h_stage_bytes = 0;
        CK(cudaMallocHost(&ctx->h_stage, bytes));
        ctx->h_stage_bytes = bytes;
    }
    return MGCFD_OK;
}

static int upload_node_dat(mgcfd_ctx *ctx, int level, double *dst, const double *src_file_order, int dim)
{
    LevelHost &L = ctx->H[level];
    const size_t bytes = (size_t)L.n_nodes * dim * sizeof(double);
    if (bytes == 0) return MGCFD_OK;
    const bool pinned = is_pinned(src_file_order);
    int rc = ensure_staging(ctx, bytes, !pinned);
    if (rc) return rc;
    const void *src = src_file_order;
    if (!pinned) { memcpy(ctx->h_stage, src_file_order, bytes); src = ctx->h_stage; }
    CK(cudaMemcpyAsync(ctx->d_stage, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->launches += k_permute_rows(ctx->stream, L.n_nodes, dim, static_cast<const double *>(ctx->d_stage), ctx->D[level].perm, dst, true);
    CK(cudaStreamSynchronize(ctx->stream));
    return api_check_launch(ctx, "permute_rows");
}

static int download_node_dat(mgcfd_ctx *ctx, int level, const double *src, double *dst_file_order, int dim)
{
    LevelHost &L = ctx->H[level];
    const size_t bytes = (size_t)L.n_nodes * dim * sizeof(double);
    if (bytes == 0) return MGCFD_OK;
    const bool pinned = is_pinned(dst_file_order);
    int rc = ensure_staging(ctx, bytes, !pinned);
    if (rc) return rc;
    ctx->launches += k_permute_rows(ctx->stream, L.n_nodes, dim, src, ctx->D[level].perm, static_cast<double *>(ctx->d_stage), false);
    CK(cudaMemcpyAsync(pinned ? (void *)dst_file_order : ctx->h_stage, ctx->d_stage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!pinned) memcpy(dst_file_order, ctx->h_stage, bytes);
    return api_check_launch(ctx, "permute_rows");
}

int mgcfd_plan(mgcfd_ctx *ctx)
{
    REQUIRE(ctx, "null ctx");
    REQUIRE(ctx->have_consts, "mgcfd_decl_consts() must precede mgcfd_plan()");
    for (int l = 0; l < ctx->n_levels; l++) {
        LevelHost &L = ctx->H[l];
        if (l + 1 < ctx->n_levels) REQUIRE((int)L.mg.size() == L.n_nodes, "node-->mg_node missing on a non-coarsest level");
        plan_renumber(L, ctx->opt.renumber != 0);
    }
    // node-->mg_node points into the next level (declared by now): range check in file order, also on planning-only contexts
    for (int l = 0; l + 1 < ctx->n_levels; l++) {
        const LevelHost &L = ctx->H[l];
        const int nc = ctx->H[l + 1].n_nodes;
        for (int f = 0; f < L.n_nodes; f++) {
            const int p = L.mg[f];
            if (p == -1 && f >= L.n_owned) continue;                       // halo node whose parent lives on another rank
            REQUIRE(p >= 0 && p < nc, "node-->mg_node entry out of range (check base_array_index)");
        }
    }
    if (ctx->n_ranks > 1)
        for (int l = 0; l < ctx->n_levels; l++) {
            // the export lists per owned node (internal numbering): (destination slot, row in that destination's import
            // range); destination slots number the neighbours that receive rows, in neighbour order.  The stage and node
            // kernels push exported rows with these; plan_query exposes them for the CPU tests.
            LevelHost &L = ctx->H[l];
            HaloLevel &Hd = ctx->halo[l];
            std::vector<std::vector<int>> ent_of(L.n_owned);
            int slot = 0;
            for (size_t k = 0; k < L.nbr_rank.size(); k++) {
                if (L.export_ptr[k + 1] == L.export_ptr[k]) continue;
                for (int j = L.export_ptr[k]; j < L.export_ptr[k + 1]; j++) {
                    std::vector<int> &e = ent_of[L.new_of_old[L.export_idx[j]]];
                    e.push_back(slot);
                    e.push_back(j - L.export_ptr[k]);
                }
                slot++;
            }
            Hd.xn_ptr.assign(L.n_owned + 1, 0);
            Hd.xn_ent.clear();
            for (int v = 0; v < L.n_owned; v++) {
                Hd.xn_ptr[v] = (int)Hd.xn_ent.size() / 2;
                Hd.xn_ent.insert(Hd.xn_ent.end(), ent_of[v].begin(), ent_of[v].end());
            }
            Hd.xn_ptr[L.n_owned] = (int)Hd.xn_ent.size() / 2;
        }
    if (ctx->device < 0) { ctx->planned = true; return MGCFD_OK; }
    CK(cudaSetDevice(ctx->device));
    if (ctx->n_ranks > 1) {
        // partitioned context: variables (both buffers) and residuals of every level, the exchange flags and the
        // min_dt mailboxes live in one arena that peers can map and store into (p2p transport)
        REQUIRE(ctx->n_levels <= P2P_MAX_LEVELS && ctx->n_ranks <= P2P_MAX_RANKS, "too many levels / ranks for the p2p arena");
        P2PState &P = ctx->p2p;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        memset(&P.me, 0, sizeof(P.me));
        P.me.n_levels = ctx->n_levels;
        P.me.rank = ctx->rank;
        for (int l = 0; l < ctx->n_levels; l++) {
            size_t bytes = (size_t)std::max(ctx->H[l].n_nodes, 1) * 40;
            P.me.off_var[0][l] = (long long)take(bytes);
            P.me.off_var[1][l] = (long long)take(bytes);
            P.me.off_res[l] = (long long)take(bytes);
            P.me.n_owned[l] = ctx->H[l].n_owned;
            for (int r = 0; r < P2P_MAX_RANKS; r++) { P.me.import_off[l][r] = -1; P.me.import_cnt[l][r] = 0; }
            for (size_t k = 0; k < ctx->H[l].nbr_rank.size(); k++) {
                int cnt = ctx->H[l].import_ptr[k + 1] - ctx->H[l].import_ptr[k];
                if (cnt > 0) { P.me.import_off[l][ctx->H[l].nbr_rank[k]] = ctx->H[l].import_ptr[k]; P.me.import_cnt[l][ctx->H[l].nbr_rank[k]] = cnt; }
            }
        }
        P.me.off_flags = (long long)take(sizeof(unsigned long long) * P2P_MAX_RANKS * (3 + 2 * P2P_MAX_LEVELS));
        P.arena_bytes = off;
        CK(cudaMalloc((void **)&P.arena, P.arena_bytes));
        CK(cudaMemsetAsync(P.arena, 0, P.arena_bytes, ctx->stream));
        P.arena_owner = true;
        int rcc = dev_alloc(ctx, &P.d_counters, (size_t)P2P_MAX_RANKS * 4);
        if (rcc) return rcc;
        for (int l = 0; l < ctx->n_levels; l++) {
            LevelDev &D = ctx->D[l];
            D.var = reinterpret_cast<double *>(P.arena + P.me.off_var[0][l]);
            D.var_alt = reinterpret_cast<double *>(P.arena + P.me.off_var[1][l]);
            D.res = reinterpret_cast<double *>(P.arena + P.me.off_res[l]);
            D.in_arena = true;
        }
    }
    for (int l = 0; l < ctx->n_levels; l++) {
        LevelHost &L = ctx->H[l];
        LevelDev &D = ctx->D[l];
        const size_t n = L.n_nodes;
        int rc;
        if (!D.in_arena) {
            if ((rc = dev_alloc(ctx, &D.var, n * 5))) return rc;
            if ((rc = dev_alloc(ctx, &D.var_alt, n * 5))) return rc;
            if ((rc = dev_alloc(ctx, &D.res, n * 5))) return rc;
        }
        if ((rc = dev_alloc(ctx, &D.old, n * 5))) return rc;
        if ((rc = dev_alloc(ctx, &D.flux, n * 5))) return rc;
        if ((rc = dev_alloc(ctx, &D.vol, n))) return rc;
        if ((rc = dev_alloc(ctx, &D.cbrt_vol, n))) return rc;
        if ((rc = dev_alloc(ctx, &D.sf, n))) return rc;
        if ((rc = dev_alloc(ctx, &D.coords, n * 3))) return rc;
        if ((rc = dev_alloc(ctx, &D.up_count, n))) return rc;
        D.flux_is_zero = true;
        if ((rc = dev_upload(ctx, &D.perm, L.new_of_old))) return rc;
        if ((rc = upload_node_dat(ctx, l, D.coords, L.coords.data(), 3))) return rc;
        if ((rc = upload_bnd(ctx, l))) return rc;
        if (l + 1 < ctx->n_levels) {
            LevelHost &Cs = ctx->H[l + 1];
            std::vector<int> mg(n);
            for (size_t i = 0; i < n; i++) {
                int f = L.old_of_new[i], p = L.mg[f];
                if (p == -1 && f >= L.n_owned) { mg[i] = -1; continue; }   // halo node whose parent lives elsewhere
                REQUIRE(p >= 0 && p < Cs.n_nodes, "node-->mg_node entry out of range");
                mg[i] = Cs.new_of_old[p];
            }
            if ((rc = dev_upload(ctx, &D.mg, mg))) return rc;
            // restrict gather lists of the OWNED coarse nodes: children (owned or halo) in ascending file order of
            // the undecomposed mesh, which is the order OP2-seq applies the increments in
            std::vector<int> forder(L.n_nodes);
            std::iota(forder.begin(), forder.end(), 0);
            if (!L.global_node.empty())
                std::stable_sort(forder.begin(), forder.end(), [&](int x, int y) { return L.global_node[x] < L.global_node[y]; });
            auto parent_of = [&](int f) {
                int p = L.mg[f];
                if (p < 0) return -1;
                p = Cs.new_of_old[p];
                return p < Cs.n_owned ? p : -1;
            };
            std::vector<int> cptr(Cs.n_nodes + 1, 0), cidx;
            for (int f : forder) {
                int p = parent_of(f);
                if (p >= 0) cptr[p + 1]++;
            }
            for (int p = 0; p < Cs.n_nodes; p++) cptr[p + 1] += cptr[p];
            cidx.resize(cptr[Cs.n_nodes]);
            std::vector<int> fill(cptr.begin(), cptr.end() - 1);
            for (int f : forder) {
                int p = parent_of(f);
                if (p >= 0) cidx[fill[p]++] = L.new_of_old[f];
            }
            LevelDev &DC = ctx->D[l + 1];
            if ((rc = dev_upload(ctx, &DC.child_ptr, cptr))) return rc;
            if ((rc = dev_upload(ctx, &DC.child_idx, cidx))) return rc;
        }
    }
    for (int l = 0; l < ctx->n_levels; l++) {
        LevelHost &L = ctx->H[l];
        HaloLevel &Hd = ctx->halo[l];
        Hd.nbr_rank = L.nbr_rank;
        Hd.exp_ptr = L.export_ptr;
        Hd.imp_ptr = L.import_ptr;
        Hd.n_export = (int)L.export_idx.size();
        if (Hd.n_export) {
            std::vector<int> idx(L.export_idx.size());
            for (size_t i = 0; i < idx.size(); i++) idx[i] = L.new_of_old[L.export_idx[i]];
            int rc;
            if ((rc = dev_upload(ctx, &Hd.d_export_idx, idx))) return rc;
            if ((rc = dev_alloc(ctx, &Hd.sendbuf, (size_t)Hd.n_export * 5))) return rc;
        }
        if (ctx->n_ranks > 1) {
            // per-owned-node export entries (built on the host before the device part of the plan): (slot, row) pairs
            std::vector<int2> ent(Hd.xn_ent.size() / 2);
            for (size_t j = 0; j < ent.size(); j++) ent[j] = make_int2(Hd.xn_ent[2 * j], Hd.xn_ent[2 * j + 1]);
            int rc;
            if ((rc = dev_upload(ctx, &Hd.d_xn_ptr, Hd.xn_ptr))) return rc;
            if ((rc = dev_upload(ctx, &Hd.d_xn_ent, ent))) return rc;
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->planned = true;
    return MGCFD_OK;
}

// ------------------------------------------------------------------------------------------
// flux plans (built lazily per variant; weights re-packed when the host weights changed)
// ------------------------------------------------------------------------------------------
static inline void pack_weight(const mgcfd_ctx *ctx, const LevelHost &L, int e, double out[4])
{
    const double *w = &L.ewt[(size_t)e * 3];
    double ewt = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);                 // flux.h:48-50
    out[0] = w[0]; out[1] = w[1]; out[2] = w[2];
    out[3] = -ewt * ctx->consts.smoothing_coefficient * 0.5;                    // flux.h:139 prefix
}

static int ensure_atomic(mgcfd_ctx *ctx, int level)
{
    LevelHost &L = ctx->H[level];
    LevelDev &D = ctx->D[level];
    if (D.atomic.valid) return MGCFD_OK;
    if (!L.have_sorted) plan_sort_edges(L);
    std::vector<int2> nodes(L.n_edges);
    std::vector<double4> w(L.n_edges);
    for (int i = 0; i < L.n_edges; i++) {
        int e = L.sorted.order[i];
        nodes[i] = make_int2(L.new_of_old[L.e2n[2 * (size_t)e]], L.new_of_old[L.e2n[2 * (size_t)e + 1]]);
        double p[4];
        pack_weight(ctx, L, e, p);
        w[i] = make_double4(p[0], p[1], p[2], p[3]);
    }
    int rc;
    if ((rc = dev_upload(ctx, &D.atomic.nodes, nodes))) return rc;
    if ((rc = dev_upload(ctx, &D.atomic.w, w))) return rc;
    D.atomic.valid = true;
    cycle_drop_graphs(ctx);
    return MGCFD_OK;
}

static int ensure_colour(mgcfd_ctx *ctx, int level)
{
    LevelHost &L = ctx->H[level];
    LevelDev &D = ctx->D[level];
    if (D.colour.valid) return MGCFD_OK;
    if (!L.have_colour) plan_colour(L, ctx->opt.colour_block_edges);
    ColourPlanHost &C = L.colour;
    for (int k = 0; k < C.n_blocks; k++)
        if (C.block_ncol[k] > 63 || C.block_colour[k] > 62) {
            ctx->err = "colour plan needs more than 63 colours";
            return MGCFD_ERR_PLAN;
        }
    if (flux_colour_smem_bytes(C.max_nodes, true) > 227 * 1024 || flux_colour_smem_bytes(C.max_nodes, false) > 227 * 1024) {
        ctx->err = "colour plan block does not fit in shared memory";
        return MGCFD_ERR_PLAN;
    }
    std::vector<int> edge0(C.n_blocks + 1, 0), ncol(C.n_blocks);
    for (int s = 0; s < C.n_blocks; s++) {
        int k = C.exec_block[s];
        int lo = k * C.block_edges, hi = std::min(L.n_edges, lo + C.block_edges);
        edge0[s + 1] = edge0[s] + (hi - lo);
        ncol[s] = C.block_ncol[k];
    }
    std::vector<double4> w(L.n_edges);
    for (int i = 0; i < L.n_edges; i++) {
        double p[4];
        pack_weight(ctx, L, C.exec_edge[i], p);
        w[i] = make_double4(p[0], p[1], p[2], p[3]);
    }
    int rc;
    if ((rc = dev_upload(ctx, &D.colour.blk_edge0, edge0))) return rc;
    if ((rc = dev_upload(ctx, &D.colour.blk_node0, C.node_off))) return rc;
    if ((rc = dev_upload(ctx, &D.colour.blk_ncol, ncol))) return rc;
    if ((rc = dev_upload(ctx, &D.colour.node_gid, C.node_gid))) return rc;
    if ((rc = dev_upload(ctx, &D.colour.lab, C.lab))) return rc;
    if ((rc = dev_upload(ctx, &D.colour.ecol, C.ecol))) return rc;
    if ((rc = dev_upload(ctx, &D.colour.w, w))) return rc;
    D.colour.valid = true;
    cycle_drop_graphs(ctx);
    return MGCFD_OK;
}

static int build_owner_host(mgcfd_ctx *ctx, int level)
{
    LevelHost &L = ctx->H[level];
    if (L.have_owner) return MGCFD_OK;
    int nb = ctx->opt.owner_chunk_nodes;
    // caps on local nodes and edges bound the shared-memory footprint (DESIGN.md "owner chunk sizing")
    int max_loc = nb + (nb * 3) / 2 + 64, max_edges = nb * 6;      // 6 edges per owned node: coarse levels (4-4.6 edges/node) fill their chunks
    // tuning knobs for experiments (not part of the specified plan): override the local-node / edge caps
    if (const char *e = getenv("MGCFD_OWNER_MAX_LOC")) max_loc = atoi(e);
    if (const char *e = getenv("MGCFD_OWNER_MAX_EDGES")) max_edges = atoi(e);
    std::string err;
    if (!plan_owner(L, nb, max_loc, max_edges, err)) {
        ctx->err = err;
        return MGCFD_ERR_PLAN;
    }
    return MGCFD_OK;
}

// ---------------------------------------------------------------------------------------
// Bank-aware slotting of a chunk's edges (device packing only; the plan's index sets are untouched).
// The edge phase reads the two endpoint states of 16 consecutive edge slots per shared-memory wavefront and the node
// phase reads the flux vectors of the incidences 16 consecutive threads are at; both are 8-byte accesses, so a
// wavefront is conflict-free when its 16 addresses fall into 16 different 8-byte banks (or coincide).  With the AoS
// state tile and the SoA flux planes the bank is (local node mod 16) resp. (edge slot mod 16).  Two greedy passes:
//   1. groups of 16 edges whose `a` ends and whose `b` ends are pairwise different mod 16 (or the same node), scanning
//      a bounded window of the remaining edges in first-touch order;
//   2. inside each group (any lane order costs the edge phase the same) the lanes are assigned so that the incidences
//      a half-warp of node threads reads together sit in different banks.
// `split` = threads per owned node in the node phase (1: thread n reads incidence `it` of node n; 2: threads 2n, 2n+1
// read incidences 2*it, 2*it+1).  On the M6-shaped deck this takes the state gathers from 1.98 to 1.08 wavefronts per
// half-warp and the flux-vector gathers from 2.08 to 1.24 (ideal 1).  slot[i] = new position of plan edge i.
// ---------------------------------------------------------------------------------------
// node order of a chunk's node phase: owned nodes by descending degree (stable), so that the 16 nodes a warp handles (two
// threads per node) have about the same number of incidences and the warp's trip count is not set by one outlier
static void degree_order(int n_own, const uint16_t *rowptr, unsigned char *order)
{
    std::vector<int> idx(n_own);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return rowptr[x + 1] - rowptr[x] > rowptr[y + 1] - rowptr[y]; });
    for (int i = 0; i < n_own; i++) order[i] = (unsigned char)idx[i];
}

static void bank_aware_slots(int ne, const uint32_t *lab, int n_own, const uint16_t *rowptr, const uint16_t *csr, int split,
                             std::vector<int> &slot)
{
    // thread slot -> node of the node phase (split 2 = the fast build's stage2 kernel: degree order; else identity)
    std::vector<unsigned char> norder(std::max(n_own, 1));
    if (split == 2 && n_own <= 64) degree_order(n_own, rowptr, norder.data());      // (chunks of up to 64 nodes run the stage2 kernel)
    else for (int i = 0; i < n_own; i++) norder[i] = (unsigned char)i;
    slot.assign(ne, 0);
    std::vector<int> order;
    order.reserve(ne);
    std::vector<char> used(ne, 0);
    const int WINDOW = 160;
    int start = 0;
    while ((int)order.size() < ne) {
        int bank_a[16], bank_b[16], cnt = 0;
        std::fill(bank_a, bank_a + 16, -1);
        std::fill(bank_b, bank_b + 16, -1);
        const size_t g0 = order.size();
        for (int e = start, seen = 0; e < ne && cnt < 16 && seen < WINDOW; e++) {
            if (used[e]) continue;
            seen++;
            const int a = (int)(lab[e] & 0xffff), b = (int)(lab[e] >> 16);
            if ((bank_a[a & 15] == -1 || bank_a[a & 15] == a) && (bank_b[b & 15] == -1 || bank_b[b & 15] == b)) {
                bank_a[a & 15] = a; bank_b[b & 15] = b;
                used[e] = 1; order.push_back(e); cnt++;
            }
        }
        for (int e = start; e < ne && cnt < 16; e++)           // not enough compatible edges: fill up in order
            if (!used[e]) { used[e] = 1; order.push_back(e); cnt++; }
        (void)g0;
        while (start < ne && used[start]) start++;
    }
    // node-phase access sets: set id per (half-warp of node threads, iteration)
    std::vector<int> set_of[2];                                 // per edge: the (at most two) access sets it belongs to
    set_of[0].assign(ne, -1); set_of[1].assign(ne, -1);
    int n_sets = 0;
    const int nodes_per_hw = 16 / split;
    for (int n0 = 0; n0 < n_own; n0 += nodes_per_hw) {
        int max_it = 0;
        for (int q = n0; q < std::min(n_own, n0 + nodes_per_hw); q++) {
            const int n = n_own <= 256 ? norder[q] : q;
            max_it = std::max(max_it, (rowptr[n + 1] - rowptr[n] + split - 1) / split);
        }
        for (int q = n0; q < std::min(n_own, n0 + nodes_per_hw); q++) {
            const int n = n_own <= 256 ? norder[q] : q;
            for (int j = rowptr[n]; j < rowptr[n + 1]; j++) {
                const int e = csr[j] & 0x7fff, sid = n_sets + (j - rowptr[n]) / split;
                (set_of[0][e] == -1 ? set_of[0][e] : set_of[1][e]) = sid;
            }
        }
        n_sets += max_it;
    }
    std::vector<unsigned char> set_bank((size_t)std::max(n_sets, 1) * 16, 0);     // edges already placed per (set, bank)
    auto pressure = [&](int e) {
        int p = 0;
        for (int k = 0; k < 2; k++)
            if (set_of[k][e] >= 0)
                for (int b = 0; b < 16; b++) p += set_bank[(size_t)set_of[k][e] * 16 + b];
        return p;
    };
    std::vector<int> grp;
    for (int g0 = 0; g0 < ne; g0 += 16) {
        const int r = std::min(16, ne - g0);
        grp.assign(order.begin() + g0, order.begin() + g0 + r);
        std::stable_sort(grp.begin(), grp.end(), [&](int x, int y) { return pressure(x) > pressure(y); });   // most constrained first
        bool taken[16] = {false};
        for (int e : grp) {
            int best = -1, best_cost = 1 << 30;
            for (int b = 0; b < r; b++) {
                if (taken[b]) continue;
                int c = 0;
                for (int k = 0; k < 2; k++)
                    if (set_of[k][e] >= 0) c += set_bank[(size_t)set_of[k][e] * 16 + b];
                if (c < best_cost) { best_cost = c; best = b; }
            }
            taken[best] = true;
            slot[e] = g0 + best;
            for (int k = 0; k < 2; k++)
                if (set_of[k][e] >= 0) set_bank[(size_t)set_of[k][e] * 16 + best]++;
        }
    }
}

// slots of the chunks [k0, k1), concatenated like OwnerPlanHost::edge_file (index: edge_off[k] - edge_off[k0] + i);
// chunks are independent, so the greedy passes run on up to 16 host threads (identity slots when slotting is off)
static void slots_for_chunks(const OwnerPlanHost &O, int k0, int k1, int split, bool slotting, std::vector<int> &buf)
{
    buf.resize((size_t)(O.edge_off[k1] - O.edge_off[k0]));
    auto work = [&](int a, int b) {
        std::vector<int> slot;
        for (int k = a; k < b; k++) {
            int *dst = buf.data() + (O.edge_off[k] - O.edge_off[k0]);
            if (slotting) {
                bank_aware_slots(O.n_edges[k], &O.lab[O.edge_off[k]], O.node0[k + 1] - O.node0[k], &O.rowptr[O.rowptr_off[k]],
                                 O.csr.data() + O.csr_off[k], split, slot);
                std::copy(slot.begin(), slot.end(), dst);
            } else {
                std::iota(dst, dst + O.n_edges[k], 0);
            }
        }
    };
    const int n = k1 - k0;
    int nt = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    nt = std::max(1, std::min(nt, n / 64));                  // small plans: not worth a thread
    if (nt == 1) { work(k0, k1); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++) pool.emplace_back(work, k0 + (int)((long long)n * t / nt), k0 + (int)((long long)n * (t + 1) / nt));
    for (std::thread &t : pool) t.join();
}

// slots of one chunk at a time, computed a batch of chunks ahead (bounded host memory on 100M-node decks)
struct SlotBatches {
    const OwnerPlanHost &O;
    int split;
    bool slotting;
    int batch;
    std::vector<int> buf;
    int k0 = -1;
    const int *of(int k)
    {
        if (k0 < 0 || k < k0 || k >= k0 + batch) {
            k0 = k - k % batch;
            slots_for_chunks(O, k0, std::min(O.n_chunks, k0 + batch), split, slotting, buf);
        }
        return buf.data() + (O.edge_off[k] - O.edge_off[k0]);
    }
};

static int ensure_owner(mgcfd_ctx *ctx, int level)
{
    LevelHost &L = ctx->H[level];
    LevelDev &D = ctx->D[level];
    if (D.owner.valid) return MGCFD_OK;
    int rc0 = build_owner_host(ctx, level);
    if (rc0) return rc0;
    OwnerPlanHost &O = L.owner;
    if (O.max_own > 256 || O.max_loc > 768) {
        ctx->err = "owner variant needs owner_chunk_nodes <= 256";
        return MGCFD_ERR_PLAN;
    }
    // device packing: the plan's blob, then the boundary entries of the chunk's owned nodes (they are a contiguous
    // range of the entries sorted by internal node)
    auto pad16 = [](long long b) { return (b + 15) & ~15ll; };
    O.dev_blob_off.assign(O.n_chunks + 1, 0);
    O.dev_max_blob = 0;
    O.dev_max_tail = 0;
    for (int k = 0; k < O.n_chunks; k++) {
        long long nb = L.bnd_node_ptr[O.node0[k + 1]] - L.bnd_node_ptr[O.node0[k]];
        const long long n_own = O.node0[k + 1] - O.node0[k];
        // plan blob | node order of the node phase (u8 per owned node) | boundary entries
        long long bytes = (O.blob_off[k + 1] - O.blob_off[k]) + pad16(n_own) + (nb ? pad16(nb * (24 + 2) + ((n_own + 2) & ~1ll) * 2) : 0);
        O.dev_blob_off[k + 1] = O.dev_blob_off[k] + bytes;
        O.dev_max_blob = std::max(O.dev_max_blob, (int)bytes);
        O.dev_max_tail = std::max(O.dev_max_tail, (int)(O.blob_off[k + 1] - O.blob_off[k] + pad16(n_own)) - 32 * ((O.n_edges[k] + 3) & ~3));
    }
    if (flux_owner_smem_bytes(O.max_loc, O.max_edges, O.dev_max_blob, ctx->opt.exact_arith != 0) > 227 * 1024) {
        ctx->err = "owner chunk does not fit in shared memory; lower owner_chunk_nodes";
        return MGCFD_ERR_PLAN;
    }
    if (ctx->n_ranks > 1 && !ctx->halo[level].d_chunk_list) {
        // chunks that own exported nodes run first so that the halo exchange overlaps the remaining (interior) chunks
        HaloLevel &Hd = ctx->halo[level];
        std::vector<char> exported(L.n_owned, 0);
        for (int f : L.export_idx) exported[L.new_of_old[f]] = 1;
        std::vector<int> first, rest;
        for (int k = 0; k < O.n_chunks; k++) {
            bool b = false;
            for (int v = O.node0[k]; v < O.node0[k + 1] && !b; v++) b = exported[v];
            (b ? first : rest).push_back(k);
        }
        Hd.n_boundary_chunks = (int)first.size();
        Hd.n_chunks = O.n_chunks;
        // fused push: per chunk with exported nodes the row pointers of its owned nodes into the (destination slot, row)
        // entries; destination slots number the neighbours that receive rows from this rank, in neighbour order, and a
        // row is the node's position in the export list for that neighbour (= its row in the neighbour's import range)
        std::vector<std::vector<int2>> ent_of(L.n_owned);
        {
            int slot = 0;
            for (size_t k = 0; k < L.nbr_rank.size(); k++) {
                if (L.export_ptr[k + 1] == L.export_ptr[k]) continue;
                for (int j = L.export_ptr[k]; j < L.export_ptr[k + 1]; j++)
                    ent_of[L.new_of_old[L.export_idx[j]]].push_back(make_int2(slot, j - L.export_ptr[k]));
                slot++;
            }
        }
        Hd.xp_base.assign(O.n_chunks, -1);
        std::vector<int> xp_ptr;
        std::vector<int2> xp_ent;
        for (int k : first) {
            Hd.xp_base[k] = (int)xp_ptr.size();
            for (int v = O.node0[k]; v < O.node0[k + 1]; v++) {
                xp_ptr.push_back((int)xp_ent.size());
                xp_ent.insert(xp_ent.end(), ent_of[v].begin(), ent_of[v].end());
            }
            xp_ptr.push_back((int)xp_ent.size());
        }
        first.insert(first.end(), rest.begin(), rest.end());
        Hd.launch_order = first;
        int rcl = dev_upload(ctx, &Hd.d_chunk_list, first);
        if (rcl) return rcl;
        if ((rcl = dev_upload(ctx, &Hd.d_xp_base, Hd.xp_base))) return rcl;
        if ((rcl = dev_upload(ctx, &Hd.d_xp_ptr, xp_ptr))) return rcl;
        if ((rcl = dev_upload(ctx, &Hd.d_xp_ent, xp_ent))) return rcl;
    }
    // edge slots inside a chunk are chosen against shared-memory bank conflicts (MGCFD_OWNER_SLOTTING=0: plan order);
    // the fused stage's node phase uses two threads per owned node in the fast build (CTAs have 128 threads for up to
    // 64 owned nodes, 256 beyond), one in the exact build
    const char *slot_s = getenv("MGCFD_OWNER_SLOTTING"), *split_s = getenv("MGCFD_OWNER_SLOT_SPLIT");
    const bool slotting = !(slot_s && atoi(slot_s) == 0);
    int node_split = (!ctx->opt.exact_arith && O.max_own <= 128) ? 2 : 1;
    if (split_s && (atoi(split_s) == 1 || atoi(split_s) == 2)) node_split = atoi(split_s);
    SlotBatches slot_batches{O, node_split, slotting, 8192, {}, -1};
    std::vector<OwnerChunkDesc> desc(O.n_chunks);
    std::vector<unsigned char> blob((size_t)O.dev_blob_off[O.n_chunks], 0);
    for (int k = 0; k < O.n_chunks; k++) {
        const int *slot = slot_batches.of(k);
        OwnerChunkDesc &d = desc[k];
        d.node0 = O.node0[k];
        d.n_own = O.node0[k + 1] - O.node0[k];
        d.n_halo = O.halo_off[k + 1] - O.halo_off[k];
        d.halo_off = O.halo_off[k];
        d.n_edges = O.n_edges[k];
        d.e_pad = (d.n_edges + 3) & ~3;
        d.n_inc = O.n_inc[k];
        d.blob_off = O.dev_blob_off[k];
        d.blob_bytes = (int)(O.dev_blob_off[k + 1] - O.dev_blob_off[k]);
        const int b_first = L.bnd_node_ptr[O.node0[k]];
        d.has_bnd = L.bnd_node_ptr[O.node0[k + 1]] - b_first;
        const int order_off = (int)(O.blob_off[k + 1] - O.blob_off[k]);
        d.bnd_off = order_off + (int)pad16(d.n_own);
        unsigned char *base = blob.data() + d.blob_off;
        if (O.max_own <= 64 && node_split == 2) degree_order(d.n_own, &O.rowptr[O.rowptr_off[k]], base + order_off);
        else for (int i = 0; i < d.n_own && i < 256; i++) base[order_off + i] = (unsigned char)i;
        if (d.has_bnd) {
            double *bw = reinterpret_cast<double *>(base + d.bnd_off);
            uint16_t *bptr = reinterpret_cast<uint16_t *>(bw + (size_t)d.has_bnd * 3);       // [n_own+1] entry ranges
            int16_t *bgrp = reinterpret_cast<int16_t *>(bptr + (((d.n_own + 1) + 1) & ~1));
            int i = 0;
            for (int v = O.node0[k]; v < O.node0[k + 1]; v++) {
                bptr[v - O.node0[k]] = (uint16_t)i;
                for (int j = L.bnd_node_ptr[v]; j < L.bnd_node_ptr[v + 1]; j++, i++) {
                    for (int c = 0; c < 3; c++) bw[(size_t)i * 3 + c] = L.bnd_wt_sorted[(size_t)j * 3 + c];
                    bgrp[i] = (int16_t)std::max(-32768, std::min(32767, L.bnd_group_sorted[j]));   // only <=2 / 3..7 matter
                }
            }
            bptr[d.n_own] = (uint16_t)i;
        }
        double *w0 = reinterpret_cast<double *>(base), *w1 = w0 + d.e_pad, *w2 = w1 + d.e_pad, *g = w2 + d.e_pad;
        uint32_t *lab = reinterpret_cast<uint32_t *>(g + d.e_pad);
        uint16_t *rowptr = reinterpret_cast<uint16_t *>(lab + d.e_pad);
        uint16_t *csr = rowptr + (((d.n_own + 1) + 7) & ~7);
        for (int i = 0; i < d.n_edges; i++) {
            double p[4];
            pack_weight(ctx, L, O.edge_file[O.edge_off[k] + i], p);
            const int t = slot[i];
            w0[t] = p[0]; w1[t] = p[1]; w2[t] = p[2]; g[t] = p[3];
            lab[t] = O.lab[O.edge_off[k] + i];
        }
        memcpy(rowptr, &O.rowptr[O.rowptr_off[k]], sizeof(uint16_t) * (d.n_own + 1));
        for (int j = 0; j < d.n_inc; j++) {
            const uint16_t c = O.csr[O.csr_off[k] + j];
            csr[j] = (uint16_t)(slot[c & 0x7fff] | (c & 0x8000));
        }
    }
    int rc;
    if (O.max_own <= 64) {
        // stage2 / lean kernels: descriptor and halo ids of a chunk in one fixed-stride record
        int hs = 0;
        for (int k = 0; k < O.n_chunks; k++) hs = std::max(hs, O.halo_off[k + 1] - O.halo_off[k]);
        hs = (hs + 3) & ~3;
        const int xs = 12 + hs;
        std::vector<int> xtab((size_t)O.n_chunks * xs, -1);
        static_assert(sizeof(OwnerChunkDesc) == 48, "record head = descriptor");
        // records are stored in LAUNCH order (multi-GPU: chunks that own exported nodes first), so that a CTA finds its
        // record at blockIdx.x without an indirection
        const std::vector<int> &order = ctx->halo[level].launch_order;
        for (int i = 0; i < O.n_chunks; i++) {
            const int k = order.empty() ? i : order[i];
            memcpy(&xtab[(size_t)i * xs], &desc[k], sizeof(OwnerChunkDesc));
            // word 3 (halo_off in the descriptor; the record carries the ids itself) = base of the chunk's export row pointers
            xtab[(size_t)i * xs + 3] = ctx->n_ranks > 1 && !ctx->halo[level].xp_base.empty() ? ctx->halo[level].xp_base[k] : -1;
            std::copy(O.halo_gid.begin() + O.halo_off[k], O.halo_gid.begin() + O.halo_off[k + 1], xtab.begin() + (size_t)i * xs + 12);
        }
        if (D.owner.xtab) { cudaFree(D.owner.xtab); D.owner.xtab = nullptr; }
        if ((rc = dev_upload(ctx, &D.owner.xtab, xtab))) return rc;
        D.owner.xs = xs; D.owner.hs = hs;
    }
    if ((rc = dev_upload(ctx, &D.owner.desc, desc))) return rc;
    if ((rc = dev_upload(ctx, &D.owner.halo_gid, O.halo_gid))) return rc;
    if ((rc = dev_upload(ctx, &D.owner.blob, blob))) return rc;
    D.owner.blob_bytes = (long long)blob.size();
    D.owner.valid = true;
    if (getenv("MGCFD_DEBUG"))
        fprintf(stderr, "[mgcfd] level %d owner plan: %d chunks, max_own %d max_loc %d max_edges %d max_inc %d tail %d blob %d, stage2 %s\n", level,
                O.n_chunks, O.max_own, O.max_loc, O.max_edges, O.max_inc, O.dev_max_tail, O.dev_max_blob,
                flux_owner_uses_stage2(D.owner, O, ctx->opt.exact_arith != 0) ? "yes" : "no");
    cycle_drop_graphs(ctx);
    return MGCFD_OK;
}

// sliced-ELL re-layout of the owner chunks for the node-gather variant
static int ensure_gather(mgcfd_ctx *ctx, int level)
{
    LevelHost &L = ctx->H[level];
    LevelDev &D = ctx->D[level];
    if (D.gather.valid) return MGCFD_OK;
    int rc0 = build_owner_host(ctx, level);
    if (rc0) return rc0;
    OwnerPlanHost &O = L.owner;
    if (O.max_own > 256) { ctx->err = "gather variant needs owner_chunk_nodes <= 256"; return MGCFD_ERR_PLAN; }
    if (flux_gather_smem_bytes(O.max_loc, ctx->opt.exact_arith != 0) > 227 * 1024) {
        ctx->err = "gather chunk does not fit in shared memory";
        return MGCFD_ERR_PLAN;
    }
    const bool exact = ctx->opt.exact_arith != 0;
    std::vector<GatherChunkDesc> desc(O.n_chunks);
    std::vector<uint16_t> row_node((size_t)O.n_chunks * 256, 0xffff), row_deg((size_t)O.n_chunks * 256, 0);
    std::vector<uint32_t> ent;
    std::vector<double> w0, w1, w2, g;
    std::vector<int> order;
    for (int k = 0; k < O.n_chunks; k++) {
        GatherChunkDesc &d = desc[k];
        d.node0 = O.node0[k];
        d.n_own = O.node0[k + 1] - O.node0[k];
        d.n_halo = O.halo_off[k + 1] - O.halo_off[k];
        d.halo_off = O.halo_off[k];
        d.ent_off = (long long)ent.size();
        const uint16_t *rowptr = &O.rowptr[O.rowptr_off[k]];
        const uint16_t *csr = O.csr.data() + O.csr_off[k];
        order.resize(d.n_own);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
            return (rowptr[x + 1] - rowptr[x]) > (rowptr[y + 1] - rowptr[y]);
        });
        for (int s = 0; s < 8; s++) {
            int lo = s * 32, hi = std::min(d.n_own, lo + 32), len = 0;
            for (int t = lo; t < hi; t++) len = std::max(len, rowptr[order[t] + 1] - rowptr[order[t]]);
            d.slice_len[s] = (unsigned short)len;
            size_t base = ent.size();
            ent.resize(base + (size_t)len * 32, 0);
            w0.resize(ent.size(), 0.0); w1.resize(ent.size(), 0.0); w2.resize(ent.size(), 0.0); g.resize(ent.size(), 0.0);
            for (int t = lo; t < hi; t++) {
                int node = order[t], deg = rowptr[node + 1] - rowptr[node];
                row_node[(size_t)k * 256 + t] = (uint16_t)node;
                row_deg[(size_t)k * 256 + t] = (uint16_t)deg;
                for (int j = 0; j < len; j++) {
                    size_t idx = base + (size_t)j * 32 + (t - lo);
                    if (j >= deg) { ent[idx] = (uint32_t)node; continue; }   // padding: neighbour = self, zero weights
                    uint16_t c = csr[rowptr[node] + j];
                    int e = c & 0x7fff;
                    bool is_b = (c & 0x8000) != 0;
                    uint32_t lab = O.lab[O.edge_off[k] + e];
                    uint32_t nb = is_b ? (lab & 0xffff) : (lab >> 16);
                    double p[4];
                    pack_weight(ctx, L, O.edge_file[O.edge_off[k] + e], p);
                    // fast build: weights pre-signed so that this node is the edge's end "a" (flipping an edge
                    // negates its weight vector and leaves |w| unchanged); exact build keeps the file orientation
                    double sgn = (!exact && is_b) ? -1.0 : 1.0;
                    ent[idx] = nb | (is_b ? 0x10000u : 0u);
                    w0[idx] = sgn * p[0]; w1[idx] = sgn * p[1]; w2[idx] = sgn * p[2]; g[idx] = p[3];
                }
            }
        }
    }
    int rc;
    if ((rc = dev_upload(ctx, &D.gather.desc, desc))) return rc;
    if ((rc = dev_upload(ctx, &D.gather.halo_gid, O.halo_gid))) return rc;
    if ((rc = dev_upload(ctx, &D.gather.row_node, row_node))) return rc;
    if ((rc = dev_upload(ctx, &D.gather.row_deg, row_deg))) return rc;
    if ((rc = dev_upload(ctx, &D.gather.ent, ent))) return rc;
    if ((rc = dev_upload(ctx, &D.gather.w0, w0))) return rc;
    if ((rc = dev_upload(ctx, &D.gather.w1, w1))) return rc;
    if ((rc = dev_upload(ctx, &D.gather.w2, w2))) return rc;
    if ((rc = dev_upload(ctx, &D.gather.g, g))) return rc;
    D.gather.n_ent = (long long)ent.size();
    D.gather.valid = true;
    cycle_drop_graphs(ctx);
    return MGCFD_OK;
}

// emit variant: per chunk, the edges each owned node emits as two sliced-ELL half-rows per node + the incidence
// lists of the non-emitter ends, packed into one blob per chunk
static int ensure_emit(mgcfd_ctx *ctx, int level)
{
    LevelHost &L = ctx->H[level];
    LevelDev &D = ctx->D[level];
    if (D.emit.valid) return MGCFD_OK;
    if (ctx->opt.exact_arith) { ctx->err = "the emit variant is fast-arithmetic only (its sums are not in file order)"; return MGCFD_ERR_ARG; }
    int rc0 = ensure_owner(ctx, level);          // chunks, halo lists, chunk launch order
    if (rc0) return rc0;
    OwnerPlanHost &O = L.owner;
    if (O.max_own > 128) { ctx->err = "emit variant needs owner_chunk_nodes <= 128 (two threads per owned node)"; return MGCFD_ERR_PLAN; }
    std::vector<EmitChunkDesc> desc(O.n_chunks);
    std::vector<uint16_t> row_node((size_t)O.n_chunks * 256, 0xffff), row_cnt((size_t)O.n_chunks * 256, 0);
    std::vector<unsigned char> blob;
    std::vector<uint32_t> ent;
    std::vector<double> w0, w1, w2, g;
    std::vector<int> order, emitter, slot_of_edge;
    std::vector<std::vector<int>> emitted;
    int max_ent = 0, max_blob = 0;
    for (int k = 0; k < O.n_chunks; k++) {
        EmitChunkDesc &d = desc[k];
        d.node0 = O.node0[k];
        d.n_own = O.node0[k + 1] - O.node0[k];
        d.n_halo = O.halo_off[k + 1] - O.halo_off[k];
        d.halo_off = O.halo_off[k];
        d.has_bnd = L.bnd_node_ptr[O.node0[k + 1]] > L.bnd_node_ptr[O.node0[k]] ? 1 : 0;
        const int ne = O.n_edges[k];
        const uint32_t *lab = &O.lab[O.edge_off[k]];
        emitter.assign(ne, 0);
        slot_of_edge.assign(ne, -1);
        emitted.assign(d.n_own, std::vector<int>());
        for (int e = 0; e < ne; e++) {
            int la = lab[e] & 0xffff, lb = lab[e] >> 16;
            int em = (la < d.n_own && lb < d.n_own) ? std::min(la, lb) : (la < d.n_own ? la : lb);
            emitter[e] = em;
            emitted[em].push_back(e);
        }
        order.resize(d.n_own);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return emitted[x].size() > emitted[y].size(); });
        ent.clear(); w0.clear(); w1.clear(); w2.clear(); g.clear();
        const int n_threads = 2 * d.n_own;
        for (int s = 0; s < 8; s++) {
            int lo = s * 32, hi = std::min(n_threads, lo + 32), len = 0;
            for (int t = lo; t < hi; t++) {
                int cnt = (int)emitted[order[t / 2]].size();
                len = std::max(len, (cnt + 1 - (t & 1)) / 2);
            }
            d.slice_len[s] = (unsigned short)len;
            size_t base = ent.size();
            ent.resize(base + (size_t)len * 32, 0);
            w0.resize(ent.size(), 0.0); w1.resize(ent.size(), 0.0); w2.resize(ent.size(), 0.0); g.resize(ent.size(), 0.0);
            for (int t = lo; t < hi; t++) {
                int node = order[t / 2], half = t & 1, cnt = (int)emitted[node].size(), hcnt = (cnt + 1 - half) / 2;
                row_node[(size_t)k * 256 + t] = (uint16_t)node;
                row_cnt[(size_t)k * 256 + t] = (uint16_t)hcnt;
                for (int kk = 0; kk < len; kk++) {
                    size_t idx = base + (size_t)kk * 32 + (t - lo);
                    if (kk >= hcnt) { ent[idx] = (uint32_t)node; continue; }      // padding: neighbour = self, zero weights
                    int e = emitted[node][2 * kk + half];
                    int la = lab[e] & 0xffff, lb = lab[e] >> 16;
                    bool is_b = node == lb;
                    int other = is_b ? la : lb;
                    double p[4];
                    pack_weight(ctx, L, O.edge_file[O.edge_off[k] + e], p);
                    double sgn = is_b ? -1.0 : 1.0;      // flipping an edge negates its weight vector; |w| is unchanged
                    ent[idx] = (uint32_t)other | (other < d.n_own ? 0x10000u : 0u);
                    w0[idx] = sgn * p[0]; w1[idx] = sgn * p[1]; w2[idx] = sgn * p[2]; g[idx] = p[3];
                    slot_of_edge[e] = (int)idx;
                }
            }
        }
        d.n_ent = (int)ent.size();
        if (d.n_ent > 65535) { ctx->err = "emit chunk has more than 65535 row slots"; return MGCFD_ERR_PLAN; }
        max_ent = std::max(max_ent, d.n_ent);
        // incidence lists of the non-emitter ends: rowptr2[n_own+1] | csr2, both u16
        const uint16_t *rowptr = &O.rowptr[O.rowptr_off[k]];
        const uint16_t *csr = O.csr.data() + O.csr_off[k];
        std::vector<uint16_t> rp2(d.n_own + 1, 0), c2;
        for (int m = 0; m < d.n_own; m++) {
            rp2[m] = (uint16_t)c2.size();
            for (int j = rowptr[m]; j < rowptr[m + 1]; j++) {
                int e = csr[j] & 0x7fff;
                if (emitter[e] != m) c2.push_back((uint16_t)slot_of_edge[e]);
            }
        }
        rp2[d.n_own] = (uint16_t)c2.size();
        while (rp2.size() % 8) rp2.push_back(0);
        while (c2.size() % 8) c2.push_back(0);
        d.rowptr_pad = (int)rp2.size();
        d.blob_off = (long long)blob.size();
        size_t bytes = (size_t)d.n_ent * 36 + (rp2.size() + c2.size()) * 2;
        d.blob_bytes = (int)bytes;
        max_blob = std::max(max_blob, d.blob_bytes);
        blob.resize(blob.size() + bytes);
        unsigned char *p = blob.data() + d.blob_off;
        memcpy(p, w0.data(), (size_t)d.n_ent * 8); p += (size_t)d.n_ent * 8;
        memcpy(p, w1.data(), (size_t)d.n_ent * 8); p += (size_t)d.n_ent * 8;
        memcpy(p, w2.data(), (size_t)d.n_ent * 8); p += (size_t)d.n_ent * 8;
        memcpy(p, g.data(), (size_t)d.n_ent * 8); p += (size_t)d.n_ent * 8;
        memcpy(p, ent.data(), (size_t)d.n_ent * 4); p += (size_t)d.n_ent * 4;
        memcpy(p, rp2.data(), rp2.size() * 2); p += rp2.size() * 2;
        memcpy(p, c2.data(), c2.size() * 2);
    }
    if (flux_emit_smem_bytes(O.max_loc, max_ent, max_blob, O.max_own) > 227 * 1024) {
        ctx->err = "emit chunk does not fit in shared memory";
        return MGCFD_ERR_PLAN;
    }
    int rc;
    if ((rc = dev_upload(ctx, &D.emit.desc, desc))) return rc;
    if ((rc = dev_upload(ctx, &D.emit.halo_gid, O.halo_gid))) return rc;
    if ((rc = dev_upload(ctx, &D.emit.row_node, row_node))) return rc;
    if ((rc = dev_upload(ctx, &D.emit.row_cnt, row_cnt))) return rc;
    if ((rc = dev_upload(ctx, &D.emit.blob, blob))) return rc;
    D.emit.n_chunks = O.n_chunks; D.emit.max_loc = O.max_loc; D.emit.max_ent = max_ent; D.emit.max_blob = max_blob;
    D.emit.max_own = O.max_own;
    D.emit.valid = true;
    cycle_drop_graphs(ctx);
    return MGCFD_OK;
}

extern "C++" int mgcfd::api_ensure_flux_plan(mgcfd_ctx *ctx, int level)
{
    switch (ctx->opt.flux_variant) {
    case MGCFD_FLUX_EMIT: return ensure_emit(ctx, level);
    case MGCFD_FLUX_GATHER: return ensure_gather(ctx, level);
    case MGCFD_FLUX_ATOMIC: return ensure_atomic(ctx, level);
    case MGCFD_FLUX_COLOUR: return ensure_colour(ctx, level);
    case MGCFD_FLUX_OWNER: return ensure_owner(ctx, level);
    }
    ctx->err = "unknown flux variant";
    return MGCFD_ERR_ARG;
}

int mgcfd_set_flux_variant(mgcfd_ctx *ctx, int variant)
{
    REQUIRE(ctx && variant >= 0 && variant < MGCFD_FLUX_NVARIANTS, "unknown flux variant");
    ctx->opt.flux_variant = variant;
    cycle_drop_graphs(ctx);
    return MGCFD_OK;
}

// ------------------------------------------------------------------------------------------
// initialisation loops (euler3d.cpp:413-441)
// ------------------------------------------------------------------------------------------
int mgcfd_loop_initialize_variables(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    LoopScope t(ctx, "initialize_variables", level, ctx->H[level].n_nodes);
    ctx->launches += k_init_vars(ctx->stream, ctx->H[level].n_nodes, ctx->D[level].var, api_dev_consts(ctx));
    return api_check_launch(ctx, "initialize_variables_kernel");
}

int mgcfd_loop_zero_fluxes(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    ctx->launches += k_fill(ctx->stream, (long long)ctx->H[level].n_nodes * 5, ctx->D[level].flux, 0.0);
    ctx->D[level].flux_is_zero = true;
    return api_check_launch(ctx, "zero_5d_array_kernel");
}

static int upload_volumes(mgcfd_ctx *ctx, int level, const std::vector<double> &vol_file_order)
{
    LevelHost &L = ctx->H[level];
    std::vector<double> cb(vol_file_order.size());
    for (size_t i = 0; i < cb.size(); i++) cb[i] = cbrt(vol_file_order[i]);   // time_stepping_kernels.h:31 evaluates cbrt(volume)
    int rc;
    if ((rc = upload_node_dat(ctx, level, ctx->D[level].vol, vol_file_order.data(), 1))) return rc;
    if ((rc = upload_node_dat(ctx, level, ctx->D[level].cbrt_vol, cb.data(), 1))) return rc;
    (void)L;
    return MGCFD_OK;
}

int mgcfd_loop_zero_volumes(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    std::vector<double> vol(ctx->H[level].n_nodes, 0.0);
    return upload_volumes(ctx, level, vol);
}

int mgcfd_loop_calculate_cell_volumes(mgcfd_ctx *ctx, int level)
{
    // misc.h:40-76.  Run once, on the host, in file order: the volumes (an OP_INC reduction over edges) and the
    // rewritten edge weights are then bit-identical to OP2-seq.  The volumes on the device are read back first
    // so that the OP_INC semantics (vol += ...) hold for whatever they contained.
    CHECK_LEVEL(level); CHECK_PLANNED();
    LevelHost &L = ctx->H[level];
    std::vector<double> vol(L.n_nodes);
    int rc = mgcfd_fetch_dat(ctx, level, "volumes", vol.data());
    if (rc) return rc;
    for (int e = 0; e < L.n_edges; e++) {
        int a = L.e2n[2 * (size_t)e], b = L.e2n[2 * (size_t)e + 1];
        const double *c1 = &L.coords[(size_t)a * 3], *c2 = &L.coords[(size_t)b * 3];
        double *w = &L.ewt[(size_t)e * 3];
        double d[3], dist = 0.0, area = 0.0;
        for (int i = 0; i < 3; i++) { d[i] = c2[i] - c1[i]; dist += d[i] * d[i]; }
        dist = sqrt(dist);
        for (int i = 0; i < 3; i++) area += w[i] * w[i];
        area = sqrt(area);
        double tet = (1.0 / 3.0) * 0.5 * dist * area;
        vol[a] += tet;
        vol[b] += tet;
        for (int i = 0; i < 3; i++) w[i] = (d[i] / dist) * area;
        for (int i = 0; i < 3; i++) w[i] /= dist;
    }
    ctx->D[level].atomic.valid = ctx->D[level].colour.valid = ctx->D[level].owner.valid = ctx->D[level].gather.valid = ctx->D[level].emit.valid = false;
    return upload_volumes(ctx, level, vol);
}

int mgcfd_loop_dampen_ewt_edges(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    for (double &w : ctx->H[level].ewt) w *= 1e-7;                               // misc.h:78-84
    ctx->D[level].atomic.valid = ctx->D[level].colour.valid = ctx->D[level].owner.valid = ctx->D[level].gather.valid = ctx->D[level].emit.valid = false;
    return MGCFD_OK;
}

int mgcfd_loop_dampen_ewt_bnd(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    for (double &w : ctx->H[level].bwt) w *= 1e-7;
    return upload_bnd(ctx, level);
}

// ------------------------------------------------------------------------------------------
// cycle loops
// ------------------------------------------------------------------------------------------
int mgcfd_loop_copy_double(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    LoopScope t(ctx, "copy_double", level, ctx->H[level].n_owned);
    ctx->launches += k_copy(ctx->stream, ctx->H[level].n_owned, ctx->D[level].var, ctx->D[level].old);
    return api_check_launch(ctx, "copy_double_kernel");
}

int mgcfd_loop_calculate_dt(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    LoopScope t(ctx, "calculate_dt", level, ctx->H[level].n_owned);
    ctx->launches += k_calculate_dt(ctx->stream, ctx->H[level].n_owned, ctx->D[level].var, ctx->D[level].cbrt_vol,
                                    ctx->D[level].sf);
    return api_check_launch(ctx, "calculate_dt_kernel");
}

int mgcfd_loop_get_min_dt(mgcfd_ctx *ctx, int level, double *min_dt)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    REQUIRE(min_dt, "null min_dt");
    {
        LoopScope t(ctx, "get_min_dt", level, ctx->H[level].n_owned);
        ctx->h_pinned[0] = *min_dt;
        CK(cudaMemcpyAsync(&ctx->d_min_dt[level], &ctx->h_pinned[0], sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->launches += k_min_dt(ctx->stream, ctx->H[level].n_owned, ctx->D[level].sf, &ctx->d_min_dt[level], ctx->d_flags);
        CK(cudaMemcpyAsync(&ctx->h_pinned[1], &ctx->d_min_dt[level], sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    *min_dt = ctx->h_pinned[1];
    return api_check_launch(ctx, "get_min_dt_kernel");
}

int mgcfd_loop_compute_step_factor(mgcfd_ctx *ctx, int level, const double *min_dt)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    REQUIRE(min_dt, "null min_dt");
    LoopScope t(ctx, "compute_step_factor", level, ctx->H[level].n_owned);
    ctx->h_pinned[2] = *min_dt;
    CK(cudaMemcpyAsync(&ctx->d_min_dt[level], &ctx->h_pinned[2], sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // h_pinned[2] is reused by the next call
    ctx->launches += k_step_factor(ctx->stream, ctx->H[level].n_owned, ctx->D[level].vol, &ctx->d_min_dt[level], ctx->D[level].sf);
    return api_check_launch(ctx, "compute_step_factor_kernel");
}

extern "C++" int mgcfd::api_run_flux(mgcfd_ctx *ctx, int level, bool stream_kernel)
{
    int rc = api_ensure_flux_plan(ctx, level);
    if (rc) return rc;
    LevelHost &L = ctx->H[level];
    LevelDev &D = ctx->D[level];
    FluxArgs a;
    a.n_edges = L.n_edges; a.n_owned = L.n_owned; a.n_nodes = L.n_nodes;
    a.var = D.var;
    a.stream_kernel = stream_kernel;
    if (stream_kernel) {
        if (!D.dummy_flux) {
            if ((rc = dev_alloc(ctx, &D.dummy_flux, (size_t)L.n_nodes * 5))) return rc;
        }
        a.flux = D.dummy_flux;
        a.overwrite = false;
    } else {
        a.flux = D.flux;
        a.overwrite = D.flux_is_zero;
    }
    const bool exact = ctx->opt.exact_arith != 0;
    LoopScope t(ctx, stream_kernel ? "unstructured_stream" : "compute_flux_edge", level, L.n_edges);
    switch (ctx->opt.flux_variant) {
    case MGCFD_FLUX_ATOMIC: ctx->launches += flux_atomic(ctx->stream, a, D.atomic, exact); break;
    case MGCFD_FLUX_COLOUR: ctx->launches += flux_colour(ctx->stream, a, D.colour, L.colour, exact); break;
    case MGCFD_FLUX_OWNER: ctx->launches += flux_owner(ctx->stream, a, D.owner, L.owner, exact); break;
    case MGCFD_FLUX_GATHER: ctx->launches += flux_gather(ctx->stream, a, D.gather, L.owner.n_chunks, L.owner.max_loc, exact); break;
    case MGCFD_FLUX_EMIT:
        if (stream_kernel) {          // unstructured_stream_kernel has no emit form: the owner chunks run it
            int rco = ensure_owner(ctx, level);
            if (rco) return rco;
            ctx->launches += flux_owner(ctx->stream, a, D.owner, L.owner, exact);
        } else {
            ctx->launches += flux_emit(ctx->stream, a, D.emit);
        }
        break;
    }
    if (!stream_kernel) D.flux_is_zero = false;
    return api_check_launch(ctx, "compute_flux_edge_kernel");
}

extern "C++" int mgcfd::api_ensure_dummy_flux(mgcfd_ctx *ctx)
{
    for (int l = 0; l < ctx->n_levels; l++)
        if (!ctx->D[l].dummy_flux) {
            int rc = dev_alloc(ctx, &ctx->D[l].dummy_flux, (size_t)ctx->H[l].n_nodes * 5);
            if (rc) return rc;
        }
    return MGCFD_OK;
}

int mgcfd_loop_compute_flux_edge(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    return api_run_flux(ctx, level, false);
}

int mgcfd_loop_unstructured_stream(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    return api_run_flux(ctx, level, true);
}

int mgcfd_loop_compute_bnd_node_flux(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    LevelDev &D = ctx->D[level];
    LoopScope t(ctx, "compute_bnd_node_flux", level, ctx->H[level].n_bnd);
    ctx->launches += k_bnd_flux(ctx->stream, D.n_bnd_unique, D.bu_node, D.bu_ptr, D.b_group, D.b_wt, D.var, D.flux,
                                api_dev_consts(ctx), ctx->opt.exact_arith != 0);
    D.flux_is_zero = false;
    return api_check_launch(ctx, "compute_bnd_node_flux_kernel");
}

int mgcfd_loop_time_step(mgcfd_ctx *ctx, int level, const int *rk)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    REQUIRE(rk && *rk >= 0 && *rk < MGCFD_RK, "rkCycle out of range");
    LevelDev &D = ctx->D[level];
    LoopScope t(ctx, "time_step", level, ctx->H[level].n_owned);
    ctx->launches += k_time_step(ctx->stream, ctx->H[level].n_owned, *rk, D.sf, D.flux, D.old, D.var);
    D.flux_is_zero = true;
    return api_check_launch(ctx, "time_step_kernel");
}

int mgcfd_loop_residual(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    LevelDev &D = ctx->D[level];
    LoopScope t(ctx, "residual", level, ctx->H[level].n_owned);
    ctx->launches += k_residual(ctx->stream, ctx->H[level].n_owned, D.old, D.var, D.res);
    return api_check_launch(ctx, "residual_kernel");
}

int mgcfd_loop_calc_rms(mgcfd_ctx *ctx, int level, double *rms)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    REQUIRE(rms, "null rms");
    {
        LoopScope t(ctx, "calc_rms", level, ctx->H[level].n_owned);
        ctx->h_pinned[3] = *rms;
        CK(cudaMemcpyAsync(ctx->d_rms, &ctx->h_pinned[3], sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->launches += k_rms(ctx->stream, ctx->H[level].n_owned, ctx->D[level].res, ctx->d_rms);
        CK(cudaMemcpyAsync(&ctx->h_pinned[4], ctx->d_rms, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    *rms = ctx->h_pinned[4];
    return api_check_launch(ctx, "calc_rms_kernel");
}

int mgcfd_loop_count_bad_vals(mgcfd_ctx *ctx, int level, int *count)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    REQUIRE(count, "null count");
    int *hp = reinterpret_cast<int *>(&ctx->h_pinned[5]);
    {
        LoopScope t(ctx, "count_bad_vals", level, ctx->H[level].n_owned);
        hp[0] = *count;
        CK(cudaMemcpyAsync(&ctx->d_flags[0], &hp[0], sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        ctx->launches += k_bad_vals(ctx->stream, ctx->H[level].n_owned, ctx->D[level].var, &ctx->d_flags[0]);
        CK(cudaMemcpyAsync(&hp[1], &ctx->d_flags[0], sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    *count = hp[1];
    return api_check_launch(ctx, "count_bad_vals");
}

int mgcfd_loop_up_pre(mgcfd_ctx *ctx, int la)
{
    CHECK_LEVEL(la); CHECK_PLANNED();
    REQUIRE(la >= 1, "up_pre needs a finer level below");
    LoopScope t(ctx, "up_pre", la, ctx->H[la - 1].n_owned);
    ctx->launches += k_up_pre(ctx->stream, ctx->H[la - 1].n_owned, ctx->D[la - 1].mg, ctx->D[la].var, ctx->D[la].up_count);
    return api_check_launch(ctx, "up_pre_kernel");
}

int mgcfd_loop_up(mgcfd_ctx *ctx, int la)
{
    CHECK_LEVEL(la); CHECK_PLANNED();
    REQUIRE(la >= 1, "up needs a finer level below");
    LoopScope t(ctx, "up", la, ctx->H[la - 1].n_owned);
    ctx->launches += k_up(ctx->stream, ctx->H[la].n_nodes, ctx->D[la].child_ptr, ctx->D[la].child_idx, ctx->D[la - 1].var,
                          ctx->D[la].var, ctx->D[la].up_count);
    return api_check_launch(ctx, "up_kernel");
}

int mgcfd_loop_up_post(mgcfd_ctx *ctx, int la)
{
    CHECK_LEVEL(la); CHECK_PLANNED();
    REQUIRE(la >= 1, "up_post needs a finer level below");
    LoopScope t(ctx, "up_post", la, ctx->H[la].n_owned);
    ctx->launches += k_up_post(ctx->stream, ctx->H[la].n_owned, ctx->D[la].var, ctx->D[la].up_count);
    return api_check_launch(ctx, "up_post_kernel");
}

int mgcfd_loop_down(mgcfd_ctx *ctx, int level)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    REQUIRE(level + 1 < ctx->n_levels, "down needs a coarser level above");
    LevelDev &D = ctx->D[level], &A = ctx->D[level + 1];
    LoopScope t(ctx, "down", level, ctx->H[level].n_owned);
    ctx->launches += k_down(ctx->stream, ctx->H[level].n_owned, D.mg, D.var, D.res, D.coords, A.res, A.coords);
    return api_check_launch(ctx, "down_kernel");
}

// ------------------------------------------------------------------------------------------
// fetch / set / validate / introspection
// ------------------------------------------------------------------------------------------
struct DatRef { double *ptr; int dim; };
static bool find_node_dat(mgcfd_ctx *ctx, int level, const char *name, DatRef &r)
{
    LevelDev &D = ctx->D[level];
    std::string s(name);
    if (s == "variables") r = {D.var, 5};
    else if (s == "old_variables") r = {D.old, 5};
    else if (s == "residuals") r = {D.res, 5};
    else if (s == "fluxes") r = {D.flux, 5};
    else if (s == "dummy_fluxes") r = {D.dummy_flux, 5};
    else if (s == "volumes") r = {D.vol, 1};
    else if (s == "step_factors") r = {D.sf, 1};
    else if (s == "node_coordinates") r = {D.coords, 3};
    else return false;
    return true;
}

int mgcfd_sync(mgcfd_ctx *ctx)
{
    REQUIRE(ctx, "null ctx");
    CK(cudaStreamSynchronize(ctx->stream));
    return MGCFD_OK;
}

int mgcfd_fetch_dat(mgcfd_ctx *ctx, int level, const char *name, void *host_out)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    REQUIRE(name && host_out, "null argument");
    LevelHost &L = ctx->H[level];
    std::string s(name);
    if (s == "edge_weights") { memcpy(host_out, L.ewt.data(), L.ewt.size() * sizeof(double)); return MGCFD_OK; }
    if (s == "bnd_node_weights") { memcpy(host_out, L.bwt.data(), L.bwt.size() * sizeof(double)); return MGCFD_OK; }
    if (s == "up_scratch") {
        std::vector<int> tmp(L.n_nodes);
        CK(cudaMemcpyAsync(tmp.data(), ctx->D[level].up_count, tmp.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        int *out = static_cast<int *>(host_out);
        for (int i = 0; i < L.n_nodes; i++) out[i] = tmp[L.new_of_old[i]];
        return MGCFD_OK;
    }
    DatRef r;
    REQUIRE(find_node_dat(ctx, level, name, r), std::string("unknown dat '") + name + "'");
    REQUIRE(r.ptr, std::string("dat '") + name + "' has not been allocated");
    return download_node_dat(ctx, level, r.ptr, static_cast<double *>(host_out), r.dim);
}

int mgcfd_set_dat(mgcfd_ctx *ctx, int level, const char *name, const void *host_in)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    REQUIRE(name && host_in, "null argument");
    LevelHost &L = ctx->H[level];
    LevelDev &D = ctx->D[level];
    std::string s(name);
    if (s == "edge_weights") {
        memcpy(L.ewt.data(), host_in, L.ewt.size() * sizeof(double));
        D.atomic.valid = D.colour.valid = D.owner.valid = D.gather.valid = D.emit.valid = false;
        return MGCFD_OK;
    }
    if (s == "bnd_node_weights") {
        memcpy(L.bwt.data(), host_in, L.bwt.size() * sizeof(double));
        return upload_bnd(ctx, level);
    }
    if (s == "up_scratch") {
        std::vector<int> tmp(L.n_nodes);
        const int *in = static_cast<const int *>(host_in);
        for (int i = 0; i < L.n_nodes; i++) tmp[L.new_of_old[i]] = in[i];
        CK(cudaMemcpyAsync(D.up_count, tmp.data(), tmp.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return MGCFD_OK;
    }
    if (s == "volumes") {
        const double *in = static_cast<const double *>(host_in);
        return upload_volumes(ctx, level, std::vector<double>(in, in + L.n_nodes));
    }
    if (s == "dummy_fluxes" && !D.dummy_flux) {
        int rc = dev_alloc(ctx, &D.dummy_flux, (size_t)L.n_nodes * 5);
        if (rc) return rc;
    }
    DatRef r;
    REQUIRE(find_node_dat(ctx, level, name, r), std::string("unknown dat '") + name + "'");
    if (s == "fluxes") D.flux_is_zero = false;
    return upload_node_dat(ctx, level, r.ptr, static_cast<const double *>(host_in), r.dim);
}

int mgcfd_run_cycles_host(mgcfd_ctx *ctx, int n_cycles, const double *const *vin, double *const *vout)
{
    if (!ctx) return MGCFD_ERR_ARG;
    CHECK_PLANNED();
    REQUIRE(n_cycles >= 0, "negative cycle count");
    const int nl = ctx->n_levels;
    bool pinned = true;
    for (int l = 0; l < nl; l++) {
        if (vin && vin[l] && !is_pinned(vin[l])) pinned = false;
        if (vout && vout[l] && !is_pinned(vout[l])) pinned = false;
    }
    const bool fused_single = ctx->n_ranks == 1 && !ctx->nccl_comm && !ctx->p2p.ipc && n_cycles >= 1 && pinned;
    if (!fused_single) {
        // decomposed contexts, pageable buffers: the three calls one after the other
        int rc;
        for (int l = 0; l < nl; l++)
            if (vin && vin[l] && (rc = mgcfd_set_dat(ctx, l, "variables", vin[l]))) return rc;
        if ((rc = mgcfd_run_cycles(ctx, n_cycles))) return rc;
        for (int l = 0; l < nl; l++)
            if (vout && vout[l] && (rc = mgcfd_fetch_dat(ctx, l, "variables", vout[l]))) return rc;
        return MGCFD_OK;
    }
    CK(cudaSetDevice(ctx->device));
    // per-level device staging (file order <-> internal order is a device-side permutation)
    if (ctx->io_stage.size() != (size_t)nl) ctx->io_stage.assign(nl, nullptr);
    for (int l = 0; l < nl; l++)
        if (!ctx->io_stage[l] && ((vin && vin[l]) || (vout && vout[l])))
            CK(cudaMalloc((void **)&ctx->io_stage[l], (size_t)std::max(ctx->H[l].n_nodes, 1) * 5 * sizeof(double)));
    mgcfd_ctx::IoHooks io;
    io.uploaded.resize(nl); io.final_.resize(nl);
    io.wait_upload.assign(nl, 0); io.want_final.assign(nl, 0);
    io.last_cycle = n_cycles - 1;
    for (int l = 0; l < nl; l++) {
        CK(cudaEventCreateWithFlags(&io.uploaded[l], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&io.final_[l], cudaEventDisableTiming));
    }
    cudaStream_t cs = ctx->comm_stream;
    auto cleanup = [&]() {
        for (int l = 0; l < nl; l++) { cudaEventDestroy(io.uploaded[l]); cudaEventDestroy(io.final_[l]); }
        ctx->io = nullptr;
    };
    // uploads: level 0 on the compute stream (the cycle starts with it), the coarser levels on the copy stream
    for (int l = 0; l < nl; l++) {
        if (!(vin && vin[l])) continue;
        cudaStream_t st = l == 0 ? ctx->stream : cs;
        const size_t bytes = (size_t)ctx->H[l].n_nodes * 5 * sizeof(double);
        if (cudaMemcpyAsync(ctx->io_stage[l], vin[l], bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) { cleanup(); ctx->err = "upload failed"; return MGCFD_ERR_CUDA; }
        ctx->launches += k_permute_rows(st, ctx->H[l].n_nodes, 5, ctx->io_stage[l], ctx->D[l].perm, ctx->D[l].var, true);
        if (l > 0) { cudaEventRecord(io.uploaded[l], cs); io.wait_upload[l] = 1; }
    }
    for (int l = 0; l < nl; l++) io.want_final[l] = (vout && vout[l] && l > 0) ? 1 : 0;
    ctx->io = &io;
    int rc = cycle_enqueue_single_nograph(ctx, n_cycles);
    ctx->io = nullptr;
    if (rc) { cleanup(); return rc; }
    // a level the cycle never restricted into (single-level decks have none) still orders its upload before the end
    for (int l = 1; l < nl; l++)
        if (io.wait_upload[l]) cudaStreamWaitEvent(ctx->stream, io.uploaded[l], 0);
    // downloads: coarse levels on the copy stream as soon as they are final, level 0 behind the cycle
    for (int l = nl - 1; l >= 0; l--) {
        if (!(vout && vout[l])) continue;
        cudaStream_t st = l == 0 ? ctx->stream : cs;
        if (l > 0) cudaStreamWaitEvent(cs, io.final_[l], 0);
        const size_t bytes = (size_t)ctx->H[l].n_nodes * 5 * sizeof(double);
        ctx->launches += k_permute_rows(st, ctx->H[l].n_nodes, 5, ctx->D[l].var, ctx->D[l].perm, ctx->io_stage[l], false);
        if (cudaMemcpyAsync(vout[l], ctx->io_stage[l], bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) { cleanup(); ctx->err = "download failed"; return MGCFD_ERR_CUDA; }
    }
    cudaError_t e = cudaStreamSynchronize(cs);
    rc = cycle_finish_run(ctx);                      // synchronises the compute stream, reads the deferred error flags
    cleanup();
    if (e != cudaSuccess) { ctx->err = std::string("copy stream: ") + cudaGetErrorString(e); return MGCFD_ERR_CUDA; }
    return rc;
}

int mgcfd_validate_level(mgcfd_ctx *ctx, int level, const double *master, int *n_diff)
{
    CHECK_LEVEL(level); CHECK_PLANNED();
    REQUIRE(master && n_diff, "null argument");
    LevelHost &L = ctx->H[level];
    double *d_master = nullptr;
    int rc = dev_alloc(ctx, &d_master, (size_t)L.n_nodes * 5, false);
    if (rc) return rc;
    if ((rc = upload_node_dat(ctx, level, d_master, master, 5))) { cudaFree(d_master); return rc; }
    CK(cudaMemsetAsync(&ctx->d_flags[2], 0, sizeof(int), ctx->stream));
    ctx->launches += k_validate(ctx->stream, L.n_owned, ctx->D[level].var, d_master, &ctx->d_flags[2]);
    int *hp = reinterpret_cast<int *>(&ctx->h_pinned[7]);
    CK(cudaMemcpyAsync(hp, &ctx->d_flags[2], sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_master);
    *n_diff = hp[0];
    return api_check_launch(ctx, "identify_differences");
}

static long long emit(const std::vector<int> &v, int *out, long long cap)
{
    if (out) {
        if (cap < (long long)v.size()) return MGCFD_ERR_ARG;
        std::copy(v.begin(), v.end(), out);
    }
    return (long long)v.size();
}

long long mgcfd_plan_query(mgcfd_ctx *ctx, int level, const char *what, int *out, long long cap)
{
    if (!ctx || level < 0 || level >= ctx->n_levels || !what || !ctx->planned) return MGCFD_ERR_ARG;
    LevelHost &L = ctx->H[level];
    std::string s(what);
    if (s == "node_perm") return emit(L.new_of_old, out, cap);
    if (s == "edge_order") {
        if (!L.have_sorted) plan_sort_edges(L);
        return emit(L.sorted.order, out, cap);
    }
    if (s == "edge_thread_colour" || s == "edge_block_colour" || s == "n_block_colours" || s == "colour_exec_edge") {
        if (!L.have_colour) plan_colour(L, ctx->opt.colour_block_edges);
        ColourPlanHost &C = L.colour;
        if (s == "n_block_colours") return emit(std::vector<int>{C.n_block_colours}, out, cap);
        if (s == "colour_exec_edge") return emit(C.exec_edge, out, cap);
        std::vector<int> v(L.n_edges);
        for (int i = 0; i < L.n_edges; i++)
            v[L.sorted.order[i]] = (s == "edge_thread_colour") ? C.thread_colour[i] : C.block_colour[i / C.block_edges];
        return emit(v, out, cap);
    }
    if (s == "export_node_ptr") return emit(ctx->halo[level].xn_ptr, out, cap);
    if (s == "export_node_ent") return emit(ctx->halo[level].xn_ent, out, cap);
    if (s.rfind("owner_", 0) == 0) {
        if (build_owner_host(ctx, level)) return MGCFD_ERR_PLAN;
        OwnerPlanHost &O = L.owner;
        if (s == "owner_chunk_start") return emit(O.node0, out, cap);
        if (s == "owner_halo_off") return emit(O.halo_off, out, cap);
        if (s == "owner_halo_gid") return emit(O.halo_gid, out, cap);
        if (s == "owner_edge_off") return emit(O.edge_off, out, cap);
        if (s == "owner_edge_file") return emit(O.edge_file, out, cap);
        if (s == "owner_lab") return emit(std::vector<int>(O.lab.begin(), O.lab.end()), out, cap);
        if (s == "owner_slots_split2_batch7") {
            // the same slots through the batched accessor ensure_owner packs with (a small batch, for tests)
            SlotBatches sb{O, 2, true, 7, {}, -1};
            std::vector<int> all;
            for (int k = 0; k < O.n_chunks; k++) {
                const int *p = sb.of(k);
                all.insert(all.end(), p, p + O.n_edges[k]);
            }
            return emit(all, out, cap);
        }
        if (s == "owner_slots_split1" || s == "owner_slots_split2") {
            // device packing: slot of every plan edge inside its chunk (bank_aware_slots), concatenated like owner_edge_file
            std::vector<int> all;
            slots_for_chunks(O, 0, O.n_chunks, s.back() == '2' ? 2 : 1, true, all);
            return emit(all, out, cap);
        }
        if (s == "owner_stats") return emit(std::vector<int>{O.n_chunks, O.max_loc, O.max_edges, O.max_own, O.max_inc,
                                                             (int)std::min<long long>(O.total_edges, 0x7fffffff)}, out, cap);
    }
    return MGCFD_ERR_ARG;
}

// ------------------------------------------------------------------------------------------
// measurement hooks
// ------------------------------------------------------------------------------------------
int mgcfd_timers_enable(mgcfd_ctx *ctx, int on)
{
    REQUIRE(ctx, "null ctx");
    if (!on && ctx->timers_on) timers_collect(ctx);
    ctx->timers_on = on;
    return MGCFD_OK;
}

int mgcfd_timers_reset(mgcfd_ctx *ctx)
{
    REQUIRE(ctx, "null ctx");
    timers_collect(ctx);
    for (auto &kv : ctx->timers) { kv.second.ms = 0.0; kv.second.calls = 0; kv.second.elements = 0; }
    return MGCFD_OK;
}

int mgcfd_timers_get(mgcfd_ctx *ctx, const char *loop_name, int level, double *ms, long long *calls, long long *elements)
{
    REQUIRE(ctx && loop_name, "null argument");
    timers_collect(ctx);
    double m = 0.0;
    long long c = 0, el = 0;
    for (auto &kv : ctx->timers) {
        size_t h = kv.first.rfind('#');
        if (kv.first.substr(0, h) != loop_name) continue;
        if (level >= 0 && std::stoi(kv.first.substr(h + 1)) != level) continue;
        m += kv.second.ms; c += kv.second.calls; el += kv.second.elements;
    }
    if (ms) *ms = m;
    if (calls) *calls = c;
    if (elements) *elements = el;
    return MGCFD_OK;
}

int mgcfd_host_alloc(void **out, size_t bytes)
{
    if (!out) return MGCFD_ERR_ARG;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { g_create_error = std::string("cudaMallocHost: ") + cudaGetErrorString(e); return MGCFD_ERR_CUDA; }
    return MGCFD_OK;
}

void mgcfd_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

long long mgcfd_kernel_launches(const mgcfd_ctx *ctx) { return ctx ? ctx->launches : 0; }
void *mgcfd_stream(mgcfd_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

void *mgcfd_device_ptr(mgcfd_ctx *ctx, int level, const char *name)
{
    if (!ctx || level < 0 || level >= ctx->n_levels || !name || !ctx->planned) return nullptr;
    DatRef r;
    if (!find_node_dat(ctx, level, name, r)) return nullptr;
    return r.ptr;
}

}  // extern "C"
