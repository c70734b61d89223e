// cycle.cu -- whole multigrid V-cycles on the device (the schedule of euler3d.cpp:458-641 with the per-visit host
// checks deferred), on one context or on a set of ranks with halo exchange.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <dlfcn.h>

#include <nccl.h>

#include "internal.h"

using namespace mgcfd;

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
            return MGCFD_ERR_CUDA;                                                                       \
        }                                                                                                \
    } while (0)

#define REQUIRE(cond, msg)                                                                               \
    do {                                                                                                 \
        if (!(cond)) {                                                                                   \
            ctx->err = (msg);                                                                            \
            return MGCFD_ERR_ARG;                                                                        \
        }                                                                                                \
    } while (0)

// programmatic dependent launch of the cycle's kernels (internal.h): on for the enqueues of the device-driven schedule
// unless the per-call-site timers are recording (event nodes between the kernels would serialise them anyway)
thread_local int mgcfd::tl_pdl = 0;
bool mgcfd::pdl_wanted()
{
    const char *e = getenv("MGCFD_PDL");
    return e ? atoi(e) != 0 : MGCFD_PDL_DEFAULT != 0;
}
namespace {
struct PdlScope {
    int saved;
    explicit PdlScope(const mgcfd_ctx *ctx, bool allowed = true) : saved(tl_pdl) { tl_pdl = allowed && pdl_wanted() && !ctx->timers_on; }
    ~PdlScope() { tl_pdl = saved; }
};
}  // namespace

// ------------------------------------------------------------------------------------------
// one context, no communication
// ------------------------------------------------------------------------------------------
// CUDA-graph replay of the enqueue functions below.  A graph holds ONE cycle.  A cycle flips the variables double
// buffer and the min_dt slot parity of the levels it visits an odd number of times, so graphs are keyed by that
// parity state, and each graph remembers the state it leaves behind (applied to the host bookkeeping on replay).
namespace {
unsigned parity_state(mgcfd_ctx *ctx)
{
    unsigned s = 0;
    for (int l = 0; l < ctx->n_levels; l++) s |= ((unsigned)ctx->D[l].visit_parity | ((unsigned)ctx->D[l].var_flip << 1)) << (2 * l);
    return s;
}

template <typename Enqueue>
int run_with_graph(mgcfd_ctx *ctx, int n_cycles, Enqueue enqueue)
{
    const bool timed = ctx->timers_on == 3;
    bool usable = !ctx->opt.no_graphs && (!ctx->timers_on || timed) && n_cycles >= 1 && ctx->n_levels <= 15;
    for (int l = 0; l < ctx->n_levels && usable; l++) usable = ctx->D[l].flux_is_zero;
    if (!usable) return enqueue(n_cycles);
    for (int i = 0; i < n_cycles; i++) {
        const unsigned key = parity_state(ctx) | (timed ? 0x80000000u : 0u);      // timed graphs carry event-record nodes
        auto it = ctx->graphs.find(key);
        if (it == ctx->graphs.end()) {
            GraphEntry g;
            long long l0 = ctx->launches, h0 = ctx->halo_bytes;
            ctx->capture_spans.clear();
            CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue(1);
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) { ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); return MGCFD_ERR_CUDA; }
            e = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return MGCFD_ERR_CUDA; }
            g.launches = ctx->launches - l0;
            g.halo_bytes = ctx->halo_bytes - h0;
            g.spans.swap(ctx->capture_spans);
            ctx->launches = l0;              // capturing enqueues nothing; the replay below does
            ctx->halo_bytes = h0;
            for (int l = 0; l < ctx->n_levels; l++)
                g.after.push_back({ctx->D[l].var, ctx->D[l].var_alt, ctx->D[l].visit_parity, ctx->D[l].var_flip});
            it = ctx->graphs.emplace(key, g).first;
        }
        CK(cudaGraphLaunch(it->second.exec, ctx->stream));
        if (timed && !it->second.spans.empty()) {
            // one replay at a time: read the spans of this cycle before the next replay re-records the events
            CK(cudaStreamSynchronize(ctx->stream));
            for (const GraphEntry::TimedSpan &sp : it->second.spans) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) != cudaSuccess) { cudaGetLastError(); continue; }
                LoopTimer &t = ctx->timers[sp.key];
                t.ms += ms; t.calls++; t.elements += sp.elems;
            }
        }
        ctx->launches += it->second.launches;
        ctx->halo_bytes += it->second.halo_bytes;
        for (int l = 0; l < ctx->n_levels; l++) {
            const GraphEntry::LevelState &st = it->second.after[l];
            ctx->D[l].var = st.var; ctx->D[l].var_alt = st.var_alt;
            ctx->D[l].visit_parity = st.visit_parity; ctx->D[l].var_flip = st.var_flip;
        }
    }
    return MGCFD_OK;
}

int finish_run(mgcfd_ctx *ctx)
{
    int rc = api_check_launch(ctx, "mgcfd_run_cycles");
    if (rc) return rc;
    int *hp = reinterpret_cast<int *>(&ctx->h_pinned[8]);
    CK(cudaMemcpyAsync(hp, ctx->d_flags, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->comm_stream) CK(cudaStreamSynchronize(ctx->comm_stream));
    if (hp[3]) { ctx->err = "a peer rank did not answer a halo / min_dt exchange in time"; return MGCFD_ERR_COMM; }
    if (hp[1]) { ctx->err = "Fatal error during 'step factor' calculation, min_dt < 0"; return MGCFD_ERR_MIN_DT; }
    if (hp[0] > 0) { ctx->err = "Bad variable values detected"; return MGCFD_ERR_BAD_VALS; }
    return MGCFD_OK;
}
}  // namespace

static int enqueue_single(mgcfd_ctx *ctx, int n_cycles);

int mgcfd::cycle_run_single(mgcfd_ctx *ctx, int n_cycles)
{
    for (int l = 0; l < ctx->n_levels; l++) {
        int rc = api_ensure_flux_plan(ctx, l);
        if (rc) return rc;
    }
    CK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int) * 4, ctx->stream));
    if (ctx->opt.measure_mem_bound) {
        int rcm = api_ensure_dummy_flux(ctx);       // p_dummy_fluxes (euler3d.cpp:397-400): allocated outside graph capture
        if (rcm) return rcm;
    }
    int rc = run_with_graph(ctx, n_cycles, [&](int k) { return enqueue_single(ctx, k); });
    if (rc) return rc;
    return finish_run(ctx);
}

// the schedule of euler3d.cpp:458-641 for n_cycles cycles: enqueue only, no host synchronisation
static int enqueue_single(mgcfd_ctx *ctx, int n_cycles)
{
    const int nl = ctx->n_levels;
    cudaStream_t s = ctx->stream;
    const bool exact = ctx->opt.exact_arith != 0;
    const DevConsts dc = api_dev_consts(ctx);
    PdlScope pdl(ctx);
    // the owner variant runs the fused schedule: one kernel per Runge-Kutta stage, fused visit prologue and restrict
    const bool fused = (ctx->opt.flux_variant == MGCFD_FLUX_OWNER || ctx->opt.flux_variant == MGCFD_FLUX_EMIT) && !ctx->opt.no_fusion;
    int level = 0, dir = 0, i = 0;
    while (i < n_cycles) {
        LevelHost &L = ctx->H[level];
        LevelDev &D = ctx->D[level];
        const int no = L.n_owned;
        if (fused) {
            unsigned long long *slot = &ctx->d_min_enc[2 * level + D.visit_parity];
            unsigned long long *next = &ctx->d_min_enc[2 * level + (D.visit_parity ^ 1)];
            D.visit_parity ^= 1;
            // compute_step_factor rides in the first stage when that stage runs the stage2 kernel (fast build)
            const bool fold = ctx->opt.flux_variant == MGCFD_FLUX_OWNER && D.flux_is_zero && flux_owner_uses_stage2(D.owner, L.owner, exact);
            {
                LoopScope t(ctx, "visit_begin", level, no);
                ctx->launches += k_visit_begin(s, no, D.var, D.cbrt_vol, D.old, D.sf, slot, nullptr, level == 0 ? ctx->d_rms : nullptr);
            }
            if (!fold) { LoopScope t(ctx, "compute_step_factor", level, no); ctx->launches += k_step_factor_fused(s, no, D.vol, slot, next, D.sf, &ctx->d_min_dt[level], ctx->d_flags); }
            for (int rk = 0; rk < MGCFD_RK; rk++) {
                if (!D.flux_is_zero) {      // only after a caller poked the fluxes: unfused stage keeps OP_INC semantics
                    int rc = api_run_flux(ctx, level, false);
                    if (rc) return rc;
                    ctx->launches += k_bnd_flux(s, D.n_bnd_unique, D.bu_node, D.bu_ptr, D.b_group, D.b_wt, D.var, D.flux, dc, exact);
                    ctx->launches += k_time_step(s, no, rk, D.sf, D.flux, D.old, D.var);
                    D.flux_is_zero = true;
                    if (rk == MGCFD_RK - 1) {
                        ctx->launches += k_residual(s, no, D.old, D.var, D.res);
                        if (level == 0) { ctx->launches += k_rms(s, no, D.res, ctx->d_rms); ctx->launches += k_bad_vals(s, no, D.var, &ctx->d_flags[0]); }
                    }
                    continue;
                }
                RkStageArgs ra{};
                ra.old = D.old; ra.sf = D.sf; ra.var_out = D.var_alt; ra.res = D.res;
                ra.d_rms = level == 0 ? ctx->d_rms : nullptr;
                ra.d_bad = &ctx->d_flags[0];
                ra.bnd_ptr = D.bnd_ptr; ra.b_group = D.b_group; ra.b_wt = D.b_wt;
                ra.rk = rk; ra.last = rk == MGCFD_RK - 1; ra.c = dc;
                if (fold && rk == 0) {
                    ra.fold.on = 1; ra.fold.n_slots = 1; ra.fold.slot[0] = slot; ra.fold.n_wait = 0;
                    ra.fold.next_slot = next; ra.fold.d_min_out = &ctx->d_min_dt[level]; ra.fold.d_flags = ctx->d_flags;
                    ra.fold.vol = D.vol; ra.fold.sf_out = D.sf;
                }
                FluxArgs a;
                a.n_edges = L.n_edges; a.n_owned = no; a.n_nodes = L.n_nodes;
                a.var = D.var; a.flux = D.flux; a.overwrite = true; a.rk = &ra;
                {
                    LoopScope t(ctx, "rk_stage", level, L.n_edges);
                    ctx->launches += ctx->opt.flux_variant == MGCFD_FLUX_EMIT ? flux_emit(s, a, D.emit) : flux_owner(s, a, D.owner, L.owner, exact);
                }
                if (L.n_nodes > no)   // halo entries of the new buffer are refreshed by the exchange; keep them defined
                    CK(cudaMemcpyAsync(D.var_alt + (size_t)no * 5, D.var + (size_t)no * 5, (size_t)(L.n_nodes - no) * 40,
                                       cudaMemcpyDeviceToDevice, s));
                std::swap(D.var, D.var_alt);
                D.var_flip ^= 1;
                if (ctx->opt.measure_mem_bound) { int rcm = api_run_flux(ctx, level, true); if (rcm) return rcm; }      // -b, euler3d.cpp:518-525
            }
        } else {
            { LoopScope t(ctx, "copy_double", level, no); ctx->launches += k_copy(s, no, D.var, D.old); }
            { LoopScope t(ctx, "calculate_dt", level, no); ctx->launches += k_calculate_dt(s, no, D.var, D.cbrt_vol, D.sf); }
            {
                LoopScope t(ctx, "get_min_dt", level, no);
                ctx->launches += k_fill(s, 1, &ctx->d_min_dt[level], DBL_MAX);
                ctx->launches += k_min_dt(s, no, D.sf, &ctx->d_min_dt[level], ctx->d_flags);
            }
            { LoopScope t(ctx, "compute_step_factor", level, no); ctx->launches += k_step_factor(s, no, D.vol, &ctx->d_min_dt[level], D.sf); }
            for (int rk = 0; rk < MGCFD_RK; rk++) {
                int rc = api_run_flux(ctx, level, false);
                if (rc) return rc;
                {
                    LoopScope t(ctx, "compute_bnd_node_flux", level, L.n_bnd);
                    ctx->launches += k_bnd_flux(s, D.n_bnd_unique, D.bu_node, D.bu_ptr, D.b_group, D.b_wt, D.var, D.flux, dc, exact);
                }
                { LoopScope t(ctx, "time_step", level, no); ctx->launches += k_time_step(s, no, rk, D.sf, D.flux, D.old, D.var); }
                D.flux_is_zero = true;
                if (ctx->opt.measure_mem_bound) { int rcm = api_run_flux(ctx, level, true); if (rcm) return rcm; }      // -b, euler3d.cpp:518-525
            }
            { LoopScope t(ctx, "residual", level, no); ctx->launches += k_residual(s, no, D.old, D.var, D.res); }
            if (level == 0) {
                { LoopScope t(ctx, "calc_rms", level, no); ctx->launches += k_fill(s, 1, ctx->d_rms, 0.0); ctx->launches += k_rms(s, no, D.res, ctx->d_rms); }
                { LoopScope t(ctx, "count_bad_vals", level, no); ctx->launches += k_bad_vals(s, no, D.var, &ctx->d_flags[0]); }
            }
        }
        // mgcfd_run_cycles_host: in the run's last cycle the level just visited is final when no further visit of it follows
        // (the coarsest level, every level on the way down, level 0 once the prolongation into it has run)
        auto mark_final = [&](int l) {
            if (ctx->io && i == ctx->io->last_cycle && ctx->io->want_final[l]) cudaEventRecord(ctx->io->final_[l], s);
        };
        if (nl <= 1) {
            mark_final(0);
            i++;
        } else if (dir == 0) {
            level++;
            if (ctx->io && ctx->io->wait_upload[level]) {
                // the level's uploaded variables must have landed before the restrict writes into them
                cudaStreamWaitEvent(s, ctx->io->uploaded[level], 0);
                ctx->io->wait_upload[level] = 0;
            }
            LevelDev &A = ctx->D[level], &F = ctx->D[level - 1];
            const int nf = ctx->H[level - 1].n_owned;
            if (fused) {
                LoopScope t(ctx, "restrict", level, nf);
                ctx->launches += k_restrict_fused(s, ctx->H[level].n_nodes, A.child_ptr, A.child_idx, F.var, A.var, A.up_count);
            } else {
                { LoopScope t(ctx, "up_pre", level, nf); ctx->launches += k_up_pre(s, nf, F.mg, A.var, A.up_count); }
                { LoopScope t(ctx, "up", level, nf); ctx->launches += k_up(s, ctx->H[level].n_nodes, A.child_ptr, A.child_idx, F.var, A.var, A.up_count); }
                { LoopScope t(ctx, "up_post", level, ctx->H[level].n_owned); ctx->launches += k_up_post(s, ctx->H[level].n_owned, A.var, A.up_count); }
            }
            if (level == nl - 1) dir = 1;
        } else {
            // the level just visited (on the way down, or the coarsest) gets no further visit in this cycle
            mark_final(level);
            level--;
            LevelDev &F = ctx->D[level], &A = ctx->D[level + 1];
            { LoopScope t(ctx, "down", level, ctx->H[level].n_owned); ctx->launches += k_down(s, ctx->H[level].n_owned, F.mg, F.var, F.res, F.coords, A.res, A.coords); }
            if (level == 0) { mark_final(0); dir = 0; i++; }
        }
    }
    return MGCFD_OK;
}

int mgcfd::cycle_enqueue_single_nograph(mgcfd_ctx *ctx, int n_cycles)
{
    for (int l = 0; l < ctx->n_levels; l++) {
        int rc = api_ensure_flux_plan(ctx, l);
        if (rc) return rc;
    }
    CK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int) * 4, ctx->stream));
    if (ctx->opt.measure_mem_bound) {
        int rcm = api_ensure_dummy_flux(ctx);
        if (rcm) return rcm;
    }
    return enqueue_single(ctx, n_cycles);
}

int mgcfd::cycle_finish_run(mgcfd_ctx *ctx) { return finish_run(ctx); }

// ------------------------------------------------------------------------------------------
// NCCL, loaded lazily from whatever libnccl.so.2 the process already has (torch's when launched by torchrun)
// ------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load()
    {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define SYM(f, name) f = reinterpret_cast<decltype(f)>(dlsym(handle, name)); if (!f) { error = std::string("dlsym ") + name; return false; }
        SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
        SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
        SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
        return true;
    }
} g_nccl;

#define NCK(call)                                                                                        \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess) {                                                                         \
            ctx->err = std::string(#call) + ": " + g_nccl.GetErrorString(r_);                            \
            return MGCFD_ERR_CUDA;                                                                       \
        }                                                                                                \
    } while (0)

enum { DAT_VAR = 0, DAT_RES = 1 };

// bound on every device-side wait for a peer (MGCFD_COMM_TIMEOUT_MS, default 60 s)
long long comm_timeout_ns()
{
    const char *e = getenv("MGCFD_COMM_TIMEOUT_MS");
    const long long ms = e && atoll(e) > 0 ? atoll(e) : 60000;
    return ms * 1000000ll;
}

inline double *dat_ptr(mgcfd_ctx *c, int level, int which) { return which == DAT_VAR ? c->D[level].var : c->D[level].res; }

// ---- halo exchange.  Every transfer runs on the rank's communication stream:
//   producer kernel (main stream) -> ev_prod -> [comm stream: pack exports, move, land in the halo range] -> ev_ready
// and the main stream waits for ev_ready only before the next kernel that reads halo data, so whatever is queued on
// the main stream in between (the interior chunks of the same Runge-Kutta stage) overlaps the exchange.

int pack_exports(mgcfd_ctx *ctx, int level, int which)
{
    HaloLevel &H = ctx->halo[level];
    if (H.n_export == 0) return MGCFD_OK;
    ctx->launches += k_pack_rows(ctx->comm_stream, H.n_export, H.d_export_idx, dat_ptr(ctx, level, which), H.sendbuf);
    ctx->halo_bytes += (long long)H.n_export * 40;
    return api_check_launch(ctx, "pack_rows");
}

// single-process group: every rank packs, then pulls its imports from the neighbours' send buffers
int exchange_group(mgcfd_ctx **R, int n, int level, int which)
{
    for (int r = 0; r < n; r++) {
        mgcfd_ctx *ctx = R[r];
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_prod, 0));
        int rc = pack_exports(ctx, level, which);
        if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_pack, ctx->comm_stream));
    }
    for (int q = 0; q < n; q++) {
        mgcfd_ctx *ctx = R[q];
        CK(cudaSetDevice(ctx->device));
        HaloLevel &H = ctx->halo[level];
        const int no = ctx->H[level].n_owned;
        for (size_t k = 0; k < H.nbr_rank.size(); k++) {
            int cnt = H.imp_ptr[k + 1] - H.imp_ptr[k];
            if (cnt == 0) continue;
            mgcfd_ctx *src = R[H.nbr_rank[k]];
            HaloLevel &S = src->halo[level];
            size_t kq = std::find(S.nbr_rank.begin(), S.nbr_rank.end(), ctx->rank) - S.nbr_rank.begin();
            if (kq == S.nbr_rank.size() || S.exp_ptr[kq + 1] - S.exp_ptr[kq] != cnt) {
                ctx->err = "halo lists of the ranks do not match";
                return MGCFD_ERR_ARG;
            }
            CK(cudaStreamWaitEvent(ctx->comm_stream, src->ev_pack, 0));
            CK(cudaMemcpyAsync(dat_ptr(ctx, level, which) + (size_t)(no + H.imp_ptr[k]) * 5, S.sendbuf + (size_t)S.exp_ptr[kq] * 5,
                               (size_t)cnt * 40, cudaMemcpyDefault, ctx->comm_stream));
        }
        CK(cudaEventRecord(ctx->ev_done, ctx->comm_stream));
    }
    for (int r = 0; r < n; r++) {       // a send buffer may only be repacked after every neighbour has pulled from it
        mgcfd_ctx *ctx = R[r];
        CK(cudaSetDevice(ctx->device));
        for (int q : ctx->halo[level].nbr_rank) CK(cudaStreamWaitEvent(ctx->comm_stream, R[q]->ev_done, 0));
        CK(cudaEventRecord(ctx->ev_ready, ctx->comm_stream));
    }
    return MGCFD_OK;
}

// one process per GPU: grouped ncclSend / ncclRecv per neighbour, receiving straight into the halo range
int exchange_nccl(mgcfd_ctx *ctx, int level, int which)
{
    HaloLevel &H = ctx->halo[level];
    CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_prod, 0));
    if (!H.nbr_rank.empty()) {
        int rc = pack_exports(ctx, level, which);
        if (rc) return rc;
        ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
        const int no = ctx->H[level].n_owned;
        NCK(g_nccl.GroupStart());
        for (size_t k = 0; k < H.nbr_rank.size(); k++) {
            int ns = H.exp_ptr[k + 1] - H.exp_ptr[k], nr = H.imp_ptr[k + 1] - H.imp_ptr[k];
            if (ns) NCK(g_nccl.Send(H.sendbuf + (size_t)H.exp_ptr[k] * 5, (size_t)ns * 5, ncclDouble, H.nbr_rank[k], comm, ctx->comm_stream));
            if (nr) NCK(g_nccl.Recv(dat_ptr(ctx, level, which) + (size_t)(no + H.imp_ptr[k]) * 5, (size_t)nr * 5, ncclDouble,
                                    H.nbr_rank[k], comm, ctx->comm_stream));
        }
        NCK(g_nccl.GroupEnd());
    }
    CK(cudaEventRecord(ctx->ev_ready, ctx->comm_stream));
    return MGCFD_OK;
}

// direct peer stores: the pack kernel writes my exported rows straight into every neighbour's halo range, then one
// warp publishes an epoch to each neighbour's flag word and waits for the neighbours' epochs
int exchange_p2p(mgcfd_ctx *ctx, int level, int which)
{
    HaloLevel &H = ctx->halo[level];
    P2PState &P = ctx->p2p;
    CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_prod, 0));
    if (!H.nbr_rank.empty()) {
        PushTable t;
        memset(&t, 0, sizeof(t));
        t.err_flag = &ctx->d_flags[3]; t.timeout_ns = comm_timeout_ns();
        const int me = ctx->rank;
        // which of the two variables buffers is "var" right now (identical on every rank: they swap in lock step)
        const int buf = ctx->D[level].var == reinterpret_cast<double *>(P.arena + P.me.off_var[0][level]) ? 0 : 1;
        unsigned long long *cnt = P.d_counters;
        for (size_t k = 0; k < H.nbr_rank.size(); k++) {
            const int q = H.nbr_rank[k];
            const int ns = H.exp_ptr[k + 1] - H.exp_ptr[k], nr = H.imp_ptr[k + 1] - H.imp_ptr[k];
            const P2PInfo &Q = P.peer[q];
            if (ns) {
                if (Q.import_cnt[level][me] != ns) { ctx->err = "halo lists of the ranks do not match"; return MGCFD_ERR_ARG; }
                long long off = which == DAT_VAR ? Q.off_var[buf][level] : Q.off_res[level];
                int d = t.n_dst++;
                t.exp_ptr[d] = H.exp_ptr[k];
                t.exp_ptr[d + 1] = H.exp_ptr[k + 1];
                t.dst[d] = reinterpret_cast<double *>(P.peer_base[q] + off) + (size_t)(Q.n_owned[level] + Q.import_off[level][me]) * 5;
                t.dst_flag[d] = reinterpret_cast<unsigned long long *>(P.peer_base[q] + Q.off_flags) + me;
                t.sent[d] = cnt + q;
            }
            if (nr) {
                int s = t.n_src++;
                t.src_flag[s] = reinterpret_cast<const unsigned long long *>(P.arena + P.me.off_flags) + q;
                t.expected[s] = cnt + P2P_MAX_RANKS + q;
            }
        }
        // export lists are grouped by neighbour in ascending order, and neighbours without exports have empty ranges:
        // the destinations' row ranges are contiguous in the export list
        LoopScope ts(ctx, "halo_exchange", level, H.n_export, ctx->comm_stream);
        ctx->launches += k_push_rows(ctx->comm_stream, H.n_export, H.d_export_idx, dat_ptr(ctx, level, which), t);
        ctx->launches += k_signal_wait(ctx->comm_stream, t);
        ctx->halo_bytes += (long long)H.n_export * 40;
        int rc = api_check_launch(ctx, "p2p exchange");
        if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->ev_ready, ctx->comm_stream));
    return MGCFD_OK;
}

// all-reduce(MIN) through the peers' mailboxes (p2p transport); leaves ev_ready recorded on the communication stream
int min_exchange_p2p(mgcfd_ctx *ctx, int level, const unsigned long long *slot, int parity)
{
    P2PState &P = ctx->p2p;
    MinTable t;
    memset(&t, 0, sizeof(t));
    t.me = ctx->rank;
    t.parity = parity;
    unsigned long long *cnt = P.d_counters;
    for (int q = 0; q < ctx->n_ranks; q++) {
        if (q == ctx->rank) continue;
        int d = t.n_peers++;
        unsigned long long *qflags = reinterpret_cast<unsigned long long *>(P.peer_base[q] + P.peer[q].off_flags);
        t.dst_box[d] = qflags + 2 * P2P_MAX_RANKS + 2 * (ctx->rank * P2P_MAX_LEVELS + level) + parity;   // one box per level and parity
        t.dst_flag[d] = qflags + P2P_MAX_RANKS + ctx->rank;
        t.sent[d] = cnt + 2 * P2P_MAX_RANKS + q;
        t.src_flag[d] = reinterpret_cast<const unsigned long long *>(P.arena + P.me.off_flags) + P2P_MAX_RANKS + q;
        t.expected[d] = cnt + 3 * P2P_MAX_RANKS + q;
    }
    t.err_flag = &ctx->d_flags[3]; t.timeout_ns = comm_timeout_ns();
    CK(cudaEventRecord(ctx->ev_prod, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_prod, 0));
    {
        LoopScope ts(ctx, "min_exchange", level, 1, ctx->comm_stream);
        ctx->launches += k_min_exchange(ctx->comm_stream, slot, t);
    }
    CK(cudaEventRecord(ctx->ev_ready, ctx->comm_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_ready, 0));
    return api_check_launch(ctx, "p2p min exchange");
}

// device-resident StagePush tables of a p2p context: [level][output buffer][without / with residuals] (built once, after
// the flux plans -- they hold the export row pointers)
int build_push_tables(mgcfd_ctx *ctx)
{
    P2PState &P = ctx->p2p;
    if (!P.h_push.empty()) return MGCFD_OK;
    const int nl = ctx->n_levels, me = ctx->rank;
    std::vector<StagePush> tab((size_t)nl * 4);
    CK(cudaMalloc((void **)&P.d_done, 2 * sizeof(unsigned int)));
    CK(cudaMemset(P.d_done, 0, 2 * sizeof(unsigned int)));
    unsigned long long *cnt = P.d_counters;
    for (int l = 0; l < nl; l++) {
        HaloLevel &H = ctx->halo[l];
        for (int ob = 0; ob < 2; ob++)
            for (int wr = 0; wr < 2; wr++) {
                StagePush &t = tab[(size_t)(l * 2 + ob) * 2 + wr];
                memset(&t, 0, sizeof(t));
                t.xp_base = H.d_xp_base; t.xp_ptr = H.d_xp_ptr; t.xp_ent = H.d_xp_ent;
                t.n_boundary = H.n_boundary_chunks;
                t.done = P.d_done;
                t.err_flag = &ctx->d_flags[3]; t.timeout_ns = comm_timeout_ns();
                for (size_t k = 0; k < H.nbr_rank.size(); k++) {
                    const int q = H.nbr_rank[k];
                    const int ns = H.exp_ptr[k + 1] - H.exp_ptr[k], nr = H.imp_ptr[k + 1] - H.imp_ptr[k];
                    const P2PInfo &Q = P.peer[q];
                    if (ns) {
                        if (Q.import_cnt[l][me] != ns) { ctx->err = "halo lists of the ranks do not match"; return MGCFD_ERR_ARG; }
                        const int d = t.n_dst++;
                        const size_t row0 = (size_t)(Q.n_owned[l] + Q.import_off[l][me]) * 5;
                        t.var_dst[d] = reinterpret_cast<double *>(P.peer_base[q] + Q.off_var[ob][l]) + row0;
                        t.res_dst[d] = wr ? reinterpret_cast<double *>(P.peer_base[q] + Q.off_res[l]) + row0 : nullptr;
                        t.dst_flag[d] = reinterpret_cast<unsigned long long *>(P.peer_base[q] + Q.off_flags) + me;
                        t.sent[d] = cnt + q;
                    }
                    if (nr) {
                        const int sidx = t.n_src++;
                        t.src_flag[sidx] = reinterpret_cast<const unsigned long long *>(P.arena + P.me.off_flags) + q;
                        t.expected[sidx] = cnt + P2P_MAX_RANKS + q;
                    }
                }
            }
    }
    P.h_push.swap(tab);
    return MGCFD_OK;
}

// the sources of a level as a stand-alone wait (flag >= expected, no increment)
WaitTable wait_table(mgcfd_ctx *ctx, int level)
{
    WaitTable t;
    memset(&t, 0, sizeof(t));
    HaloLevel &H = ctx->halo[level];
    P2PState &P = ctx->p2p;
    for (size_t k = 0; k < H.nbr_rank.size(); k++)
        if (H.imp_ptr[k + 1] > H.imp_ptr[k]) {
            const int q = H.nbr_rank[k], sidx = t.n_src++;
            t.src_flag[sidx] = reinterpret_cast<const unsigned long long *>(P.arena + P.me.off_flags) + q;
            t.expected[sidx] = P.d_counters + P2P_MAX_RANKS + q;
        }
    t.err_flag = &ctx->d_flags[3]; t.timeout_ns = comm_timeout_ns();
    return t;
}

// node kernels of a fused-push cycle: push the `which` dat of `push_level` (current buffer), wait for the sources of
// `wait_level` (the level whose halo rows the kernel reads; -1: none)
int node_push_table(mgcfd_ctx *ctx, int push_level, int which, int wait_level, NodePush &t)
{
    memset(&t, 0, sizeof(t));
    P2PState &P = ctx->p2p;
    HaloLevel &H = ctx->halo[push_level];
    const int me = ctx->rank;
    unsigned long long *cnt = P.d_counters;
    t.on = 1;
    t.xn_ptr = H.d_xn_ptr; t.xn_ent = H.d_xn_ent;
    t.done = P.d_done; t.err_flag = &ctx->d_flags[3]; t.timeout_ns = comm_timeout_ns();
    const int buf = ctx->D[push_level].var == reinterpret_cast<double *>(P.arena + P.me.off_var[0][push_level]) ? 0 : 1;
    for (size_t k = 0; k < H.nbr_rank.size(); k++) {
        const int q = H.nbr_rank[k];
        const int ns = H.exp_ptr[k + 1] - H.exp_ptr[k], nr = H.imp_ptr[k + 1] - H.imp_ptr[k];
        const P2PInfo &Q = P.peer[q];
        if (ns) {
            if (Q.import_cnt[push_level][me] != ns) { ctx->err = "halo lists of the ranks do not match"; return MGCFD_ERR_ARG; }
            const int d = t.n_dst++;
            const long long off = which == DAT_VAR ? Q.off_var[buf][push_level] : Q.off_res[push_level];
            t.dst[d] = reinterpret_cast<double *>(P.peer_base[q] + off) + (size_t)(Q.n_owned[push_level] + Q.import_off[push_level][me]) * 5;
            t.dst_flag[d] = reinterpret_cast<unsigned long long *>(P.peer_base[q] + Q.off_flags) + me;
            t.sent[d] = cnt + q;
        }
        if (nr) {
            const int sidx = t.n_src++;
            t.src_flag[sidx] = reinterpret_cast<const unsigned long long *>(P.arena + P.me.off_flags) + q;
            t.expected[sidx] = cnt + P2P_MAX_RANKS + q;
        }
    }
    if (wait_level >= 0) {
        const WaitTable w = wait_table(ctx, wait_level);
        t.n_wait = w.n_src;
        for (int i = 0; i < w.n_src; i++) { t.wait_flag[i] = w.src_flag[i]; t.wait_expected[i] = w.expected[i]; }
    }
    return MGCFD_OK;
}

// min_dt all-reduce folded into the visit prologue / step-factor kernels (mailboxes per level and visit parity)
void min_push_table(mgcfd_ctx *ctx, int level, int parity, MinPush &t)
{
    memset(&t, 0, sizeof(t));
    P2PState &P = ctx->p2p;
    unsigned long long *cnt = P.d_counters;
    t.on = 1;
    for (int q = 0; q < ctx->n_ranks; q++) {
        if (q == ctx->rank) continue;
        const int d = t.n_peers++;
        unsigned long long *qflags = reinterpret_cast<unsigned long long *>(P.peer_base[q] + P.peer[q].off_flags);
        t.dst_box[d] = qflags + 2 * P2P_MAX_RANKS + 2 * (ctx->rank * P2P_MAX_LEVELS + level) + parity;
        t.dst_flag[d] = qflags + P2P_MAX_RANKS + ctx->rank;
        t.sent[d] = cnt + 2 * P2P_MAX_RANKS + q;
        t.src_flag[d] = reinterpret_cast<const unsigned long long *>(P.arena + P.me.off_flags) + P2P_MAX_RANKS + q;
        t.expected[d] = cnt + 3 * P2P_MAX_RANKS + q;
    }
    t.done = P.d_done; t.err_flag = &ctx->d_flags[3]; t.timeout_ns = comm_timeout_ns();
}

// end of a multi-rank run, one process per GPU, p2p: every rank's deferred error flags (bad values, min_dt < 0, a wait
// that ran out) reach every rank through the status mailboxes, so that all ranks return the same code
int status_exchange_p2p(mgcfd_ctx *ctx)
{
    P2PState &P = ctx->p2p;
    MinTable t;
    memset(&t, 0, sizeof(t));
    t.me = ctx->rank;
    unsigned long long *cnt = P.d_counters;
    const size_t status0 = 2 * P2P_MAX_RANKS + 2 * (size_t)P2P_MAX_RANKS * P2P_MAX_LEVELS;
    for (int q = 0; q < ctx->n_ranks; q++) {
        if (q == ctx->rank) continue;
        int d = t.n_peers++;
        unsigned long long *qflags = reinterpret_cast<unsigned long long *>(P.peer_base[q] + P.peer[q].off_flags);
        t.dst_box[d] = qflags + status0 + ctx->rank;
        t.dst_flag[d] = qflags + P2P_MAX_RANKS + ctx->rank;
        t.sent[d] = cnt + 2 * P2P_MAX_RANKS + q;
        t.src_flag[d] = reinterpret_cast<const unsigned long long *>(P.arena + P.me.off_flags) + P2P_MAX_RANKS + q;
        t.expected[d] = cnt + 3 * P2P_MAX_RANKS + q;
    }
    t.err_flag = &ctx->d_flags[3]; t.timeout_ns = comm_timeout_ns();
    const unsigned long long *boxes = reinterpret_cast<const unsigned long long *>(P.arena + P.me.off_flags) + status0;
    ctx->launches += k_status_exchange(ctx->stream, ctx->d_flags, boxes, ctx->n_ranks, t);
    return api_check_launch(ctx, "p2p status exchange");
}

// mark "the producers of this exchange are queued" on every rank's main stream
void mark_produced(mgcfd_ctx **R, int n)
{
    for (int r = 0; r < n; r++) {
        cudaSetDevice(R[r]->device);
        cudaEventRecord(R[r]->ev_prod, R[r]->stream);
    }
}

// start the exchange of `which` on the communication streams (after mark_produced)
int exchange_start(mgcfd_ctx **R, int n, int level, int which)
{
    if (R[0]->p2p.enabled) {
        for (int r = 0; r < n; r++) {
            cudaSetDevice(R[r]->device);
            int rc = exchange_p2p(R[r], level, which);
            if (rc) { R[0]->err = R[r]->err; return rc; }
        }
        return MGCFD_OK;
    }
    if (n == 1 && R[0]->nccl_comm) return exchange_nccl(R[0], level, which);
    if (n == 1) return MGCFD_OK;
    return exchange_group(R, n, level, which);
}

// main streams wait for the last started exchange
void exchange_wait(mgcfd_ctx **R, int n)
{
    if (n == 1 && !R[0]->nccl_comm && !R[0]->p2p.enabled) return;
    for (int r = 0; r < n; r++) {
        cudaSetDevice(R[r]->device);
        cudaStreamWaitEvent(R[r]->stream, R[r]->ev_ready, 0);
    }
}

int exchange(mgcfd_ctx **R, int n, int level, int which)      // not overlapped: produce -> exchange -> wait
{
    mark_produced(R, n);
    int rc = exchange_start(R, n, level, which);
    if (rc) return rc;
    exchange_wait(R, n);
    return MGCFD_OK;
}

int enqueue_ranks(mgcfd_ctx **R, int n, int n_cycles);

// the fused schedule over a set of ranks in lock step
int run_ranks(mgcfd_ctx **R, int n, int n_cycles)
{
    mgcfd_ctx *ctx = R[0];
    const int nl = ctx->n_levels;
    const bool nccl = n == 1 && ctx->nccl_comm;
    const bool remote = n == 1 && (ctx->nccl_comm || ctx->p2p.ipc);      // one process per GPU
    for (int r = 0; r < n; r++) {
        mgcfd_ctx *c = R[r];
        if (!c->planned || c->device < 0 || c->n_levels != nl) { ctx->err = "ranks are not planned alike"; return MGCFD_ERR_ARG; }
        if ((c->opt.flux_variant != MGCFD_FLUX_OWNER && c->opt.flux_variant != MGCFD_FLUX_EMIT) || c->opt.no_fusion) {
            ctx->err = "multi-GPU runs use the fused schedule (flux_variant owner or emit, no_fusion 0)";
            return MGCFD_ERR_ARG;
        }
        cudaSetDevice(c->device);
        for (int l = 0; l < nl; l++) {
            int rc = api_ensure_flux_plan(c, l);
            if (rc) { ctx->err = c->err; return rc; }
            if (c->opt.measure_mem_bound && (rc = api_ensure_dummy_flux(c))) { ctx->err = c->err; return rc; }
            if (!c->D[l].flux_is_zero) { ctx->err = "fluxes must be zero before a multi-GPU run"; return MGCFD_ERR_ARG; }
        }
        if (cudaMemsetAsync(c->d_flags, 0, sizeof(int) * 4, c->stream) != cudaSuccess) return MGCFD_ERR_CUDA;
    }
    int rc;
    // halos of the start state (a caller may have set variables on the owned nodes only)
    for (int l = 0; l < nl; l++)
        if ((rc = exchange(R, n, l, DAT_VAR))) return rc;
    // fused push: the stage kernels store their exported rows into the neighbours and hand-shake themselves
    for (int r = 0; r < n; r++) {
        mgcfd_ctx *c = R[r];
        if (!c->p2p.enabled || !c->p2p.fused_push) continue;
        cudaSetDevice(c->device);
        if ((rc = build_push_tables(c))) { ctx->err = c->err; return rc; }
    }
    // one process per GPU: the whole multi-stream schedule, NCCL calls included, replays as a CUDA graph
    if (remote) rc = run_with_graph(ctx, n_cycles, [&](int k) { return enqueue_ranks(R, n, k); });
    else rc = enqueue_ranks(R, n, n_cycles);
    if (rc) return rc;
    // deferred host checks (euler3d.cpp:480, :544) and exchange time-outs: every rank returns the same code
    if (remote && ctx->p2p.enabled && ctx->n_ranks > 1) {
        if ((rc = status_exchange_p2p(ctx))) return rc;
    } else if (remote && nccl && ctx->n_ranks > 1) {
        CK(cudaEventRecord(ctx->ev_prod, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_prod, 0));
        NCK(g_nccl.AllReduce(ctx->d_flags, ctx->d_flags, 4, ncclInt, ncclMax, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->comm_stream));
        CK(cudaEventRecord(ctx->ev_ready, ctx->comm_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_ready, 0));
    }
    int worst = MGCFD_OK;
    std::string worst_msg;
    for (int r = 0; r < n; r++) {
        mgcfd_ctx *c = R[r];
        cudaSetDevice(c->device);
        int rcr = finish_run(c);
        if (rcr && !worst) { worst = rcr; worst_msg = c->err; }
    }
    if (worst) { ctx->err = worst_msg; for (int r = 0; r < n; r++) R[r]->err = worst_msg; }
    return worst;
}

int enqueue_ranks(mgcfd_ctx **R, int n, int n_cycles)
{
    mgcfd_ctx *ctx = R[0];
    const int nl = ctx->n_levels;
    const bool nccl = n == 1 && ctx->nccl_comm;
    PdlScope pdl(ctx, ctx->p2p.enabled && ctx->p2p.fused_push);      // the exchange schedules with pack kernels / NCCL stay as they are
    int level = 0, dir = 0, i = 0, rc;
    while (i < n_cycles) {
        // ---- visit prologue: copy, dt, local min (euler3d.cpp:467-479)
        for (int r = 0; r < n; r++) {
            mgcfd_ctx *c = R[r];
            cudaSetDevice(c->device);
            LevelDev &D = c->D[level];
            unsigned long long *slot = &c->d_min_enc[2 * level + D.visit_parity];
            LoopScope t(c, "visit_begin", level, c->H[level].n_owned);
            if (c->p2p.enabled && c->p2p.fused_push) {
                MinPush mp;                       // the last block sends the rank's minimum to every peer's mailbox
                min_push_table(c, level, D.visit_parity, mp);
                c->launches += k_visit_begin(c->stream, c->H[level].n_owned, D.var, D.cbrt_vol, D.old, D.sf, slot, &mp, level == 0 ? c->d_rms : nullptr);
            } else {
                c->launches += k_visit_begin(c->stream, c->H[level].n_owned, D.var, D.cbrt_vol, D.old, D.sf, slot, nullptr, level == 0 ? c->d_rms : nullptr);
            }
            if (n > 1 && !c->p2p.enabled) cudaEventRecord(c->ev_k1, c->stream);
        }
        // ---- global minimum + step factor (euler3d.cpp:477-489)
        for (int r = 0; r < n; r++) {
            mgcfd_ctx *c = R[r];
            cudaSetDevice(c->device);
            LevelDev &D = c->D[level];
            unsigned long long *slot = &c->d_min_enc[2 * level + D.visit_parity];
            unsigned long long *next = &c->d_min_enc[2 * level + (D.visit_parity ^ 1)];
            const int no = c->H[level].n_owned;
            // fused push + stage2: the step factor (and the wait for the peers' minima) rides in the first stage
            if (c->p2p.enabled && c->p2p.fused_push && c->H[level].n_owned > 0 &&
                flux_owner_uses_stage2(D.owner, c->H[level].owner, c->opt.exact_arith != 0))
                continue;
            LoopScope t(c, "compute_step_factor", level, no);
            if (c->p2p.enabled) {
                // mailboxes: my minimum goes to every peer, theirs arrive in my arena; K2 reads my slot + the boxes
                const bool folded = c->p2p.fused_push;
                if (!folded && (rc = min_exchange_p2p(c, level, slot, D.visit_parity))) { ctx->err = c->err; return rc; }
                MinPush mp;
                min_push_table(c, level, D.visit_parity, mp);
                MinSlots ms;
                ms.n = 0;
                const unsigned long long *boxes = reinterpret_cast<const unsigned long long *>(c->p2p.arena + c->p2p.me.off_flags) + 2 * P2P_MAX_RANKS;
                for (int q = 0; q < c->n_ranks; q++) ms.p[ms.n++] = q == c->rank ? slot : boxes + 2 * (q * P2P_MAX_LEVELS + level) + D.visit_parity;
                c->launches += k_step_factor_group(c->stream, no, D.vol, ms, next, D.sf, &c->d_min_dt[level], c->d_flags, folded ? &mp : nullptr);
            } else if (nccl) {
                ctx = c;
                // every NCCL call of this rank goes through the communication stream, in one program order
                CK(cudaEventRecord(c->ev_prod, c->stream));
                CK(cudaStreamWaitEvent(c->comm_stream, c->ev_prod, 0));
                NCK(g_nccl.AllReduce(slot, slot, 1, ncclUint64, ncclMin, static_cast<ncclComm_t>(c->nccl_comm), c->comm_stream));
                CK(cudaEventRecord(c->ev_ready, c->comm_stream));
                CK(cudaStreamWaitEvent(c->stream, c->ev_ready, 0));
                c->launches += k_step_factor_fused(c->stream, no, D.vol, slot, next, D.sf, &c->d_min_dt[level], c->d_flags);
            } else {
                MinSlots ms;
                ms.n = n;
                for (int q = 0; q < n; q++) {
                    ms.p[q] = &R[q]->d_min_enc[2 * level + R[q]->D[level].visit_parity];
                    if (q != r) cudaStreamWaitEvent(c->stream, R[q]->ev_k1, 0);
                }
                c->launches += k_step_factor_group(c->stream, no, D.vol, ms, next, D.sf, &c->d_min_dt[level], c->d_flags);
            }
        }
        for (int r = 0; r < n; r++) R[r]->D[level].visit_parity ^= 1;
        // ---- three fused Runge-Kutta stages (euler3d.cpp:492-531).  Per stage: chunks owning exported nodes first,
        //      then the halo exchange of the new variables starts and the interior chunks run underneath it
        const bool fp = ctx->p2p.enabled && ctx->p2p.fused_push;
        for (int rk = 0; fp && rk < MGCFD_RK; rk++) {
            // fused push: ONE launch per stage over all chunks (those that own exported nodes first); the kernel waits
            // for its sources, pushes var_new (+ residuals after the last stage of a level >= 1) and publishes the epoch
            const bool last = rk == MGCFD_RK - 1;
            for (int r = 0; r < n; r++) {
                mgcfd_ctx *c = R[r];
                cudaSetDevice(c->device);
                LevelHost &L = c->H[level];
                LevelDev &D = c->D[level];
                HaloLevel &Hd = c->halo[level];
                RkStageArgs ra{};
                ra.old = D.old; ra.sf = D.sf; ra.var_out = D.var_alt; ra.res = D.res;
                ra.d_rms = level == 0 ? c->d_rms : nullptr;
                ra.d_bad = &c->d_flags[0];
                ra.bnd_ptr = D.bnd_ptr; ra.b_group = D.b_group; ra.b_wt = D.b_wt;
                ra.rk = rk; ra.last = last; ra.c = api_dev_consts(c);
                if (rk == 0 && L.n_owned > 0 && flux_owner_uses_stage2(D.owner, L.owner, c->opt.exact_arith != 0)) {
                    // (visit_parity was flipped after the prologue: the slot of this visit is the other one)
                    const int par = D.visit_parity ^ 1;
                    MinPush mp;
                    min_push_table(c, level, par, mp);
                    StepFold &f = ra.fold;
                    f.on = 1;
                    const unsigned long long *boxes = reinterpret_cast<const unsigned long long *>(c->p2p.arena + c->p2p.me.off_flags) + 2 * P2P_MAX_RANKS;
                    for (int q = 0; q < c->n_ranks; q++)
                        f.slot[f.n_slots++] = q == c->rank ? &c->d_min_enc[2 * level + par] : boxes + 2 * (q * P2P_MAX_LEVELS + level) + par;
                    f.n_wait = mp.n_peers;
                    for (int q = 0; q < mp.n_peers; q++) { f.wait_flag[q] = mp.src_flag[q]; f.wait_expected[q] = mp.expected[q]; }
                    f.next_slot = &c->d_min_enc[2 * level + (par ^ 1)];
                    f.d_min_out = &c->d_min_dt[level]; f.d_flags = c->d_flags;
                    f.vol = D.vol; f.sf_out = D.sf; f.timeout_ns = comm_timeout_ns();
                }
                const int ob = D.var_alt == reinterpret_cast<double *>(c->p2p.arena + c->p2p.me.off_var[0][level]) ? 0 : 1;
                const int wr = (last && level >= 1) ? 1 : 0;
                ra.push_on = Hd.n_boundary_chunks > 0 ? 1 : 0;
                if (ra.push_on) ra.push = c->p2p.h_push[(size_t)(level * 2 + ob) * 2 + wr];
                FluxArgs a;
                a.n_edges = L.n_edges; a.n_owned = L.n_owned; a.n_nodes = L.n_nodes;
                a.var = D.var; a.flux = D.flux; a.overwrite = true; a.rk = &ra;
                a.chunk_list = Hd.d_chunk_list; a.n_list = Hd.n_chunks; a.list_offset = 0;
                {
                    LoopScope t(c, "rk_stage", level, L.n_edges);
                    c->launches += flux_owner(c->stream, a, D.owner, L.owner, c->opt.exact_arith != 0);
                }
                c->halo_bytes += (long long)Hd.n_export * 40 * (1 + wr);
                if (Hd.n_boundary_chunks == 0) {
                    // nothing to export on this level, yet rows may arrive (restrict / prolong halos): keep the epoch book
                    PushTable t;
                    memset(&t, 0, sizeof(t));
                    t.err_flag = &c->d_flags[3]; t.timeout_ns = comm_timeout_ns();
                    WaitTable w = wait_table(c, level);
                    for (int q = 0; q < w.n_src; q++) { t.src_flag[q] = w.src_flag[q]; t.expected[q] = const_cast<unsigned long long *>(w.expected[q]); }
                    t.n_src = w.n_src;
                    c->launches += k_signal_wait(c->stream, t);
                }
                if ((rc = api_check_launch(c, "rk_stage"))) { ctx->err = c->err; return rc; }
            }
            for (int r = 0; r < n; r++) { std::swap(R[r]->D[level].var, R[r]->D[level].var_alt); R[r]->D[level].var_flip ^= 1; }
            for (int r = 0; r < n; r++)
                if (R[r]->opt.measure_mem_bound) {      // -b, euler3d.cpp:518-525 (owned + recomputed cut edges of the rank)
                    cudaSetDevice(R[r]->device);
                    if ((rc = api_run_flux(R[r], level, true))) { ctx->err = R[r]->err; return rc; }
                }
            // (no stand-alone wait after the last stage: the next kernel that reads halo rows -- restrict, prolong or the
            //  next stage -- waits for its sources itself)
        }
        for (int rk = 0; !fp && rk < MGCFD_RK; rk++) {
            const bool last = rk == MGCFD_RK - 1;
            for (int part = 0; part < 2; part++) {
                for (int r = 0; r < n; r++) {
                    mgcfd_ctx *c = R[r];
                    cudaSetDevice(c->device);
                    LevelHost &L = c->H[level];
                    LevelDev &D = c->D[level];
                    HaloLevel &Hd = c->halo[level];
                    RkStageArgs ra{};
                    ra.old = D.old; ra.sf = D.sf; ra.var_out = D.var_alt; ra.res = D.res;
                    ra.d_rms = level == 0 ? c->d_rms : nullptr;
                    ra.d_bad = &c->d_flags[0];
                    ra.bnd_ptr = D.bnd_ptr; ra.b_group = D.b_group; ra.b_wt = D.b_wt;
                    ra.rk = rk; ra.last = last; ra.c = api_dev_consts(c);
                    FluxArgs a;
                    a.n_edges = L.n_edges; a.n_owned = L.n_owned; a.n_nodes = L.n_nodes;
                    a.var = D.var; a.flux = D.flux; a.overwrite = true; a.rk = &ra;
                    a.chunk_list = Hd.d_chunk_list + (part == 0 ? 0 : Hd.n_boundary_chunks);
                    a.list_offset = part == 0 ? 0 : Hd.n_boundary_chunks;
                    a.n_list = part == 0 ? Hd.n_boundary_chunks : Hd.n_chunks - Hd.n_boundary_chunks;
                    {
                        LoopScope t(c, "rk_stage", level, part == 0 ? 0 : L.n_edges);
                        c->launches += c->opt.flux_variant == MGCFD_FLUX_EMIT ? flux_emit(c->stream, a, D.emit)
                                                                               : flux_owner(c->stream, a, D.owner, L.owner, c->opt.exact_arith != 0);
                    }
                    if ((rc = api_check_launch(c, "rk_stage"))) { ctx->err = c->err; return rc; }
                }
                if (part == 0) {
                    // exported values exist now: swap so that "var" names the new buffer, then start the exchange(s)
                    for (int r = 0; r < n; r++) { std::swap(R[r]->D[level].var, R[r]->D[level].var_alt); R[r]->D[level].var_flip ^= 1; }
                    mark_produced(R, n);
                    if ((rc = exchange_start(R, n, level, DAT_VAR))) return rc;
                    if (last && level >= 1 && (rc = exchange_start(R, n, level, DAT_RES))) return rc;   // prolong of level-1 reads res[level]
                    for (int r = 0; r < n; r++) { std::swap(R[r]->D[level].var, R[r]->D[level].var_alt); R[r]->D[level].var_flip ^= 1; }
                }
            }
            for (int r = 0; r < n; r++) { std::swap(R[r]->D[level].var, R[r]->D[level].var_alt); R[r]->D[level].var_flip ^= 1; }
            exchange_wait(R, n);
            for (int r = 0; r < n; r++)
                if (R[r]->opt.measure_mem_bound) {      // -b, euler3d.cpp:518-525
                    cudaSetDevice(R[r]->device);
                    if ((rc = api_run_flux(R[r], level, true))) { ctx->err = R[r]->err; return rc; }
                }
        }
        if (nl <= 1) {
            i++;
        } else if (dir == 0) {
            level++;
            for (int r = 0; r < n; r++) {
                mgcfd_ctx *c = R[r];
                cudaSetDevice(c->device);
                LevelDev &A = c->D[level], &F = c->D[level - 1];
                LoopScope t(c, "restrict", level, c->H[level - 1].n_owned);
                if (fp) {
                    NodePush np;                  // reads the fine level's halo children, pushes the coarse level's variables
                    if ((rc = node_push_table(c, level, DAT_VAR, level - 1, np))) { ctx->err = c->err; return rc; }
                    c->launches += k_restrict_fused(c->stream, c->H[level].n_nodes, A.child_ptr, A.child_idx, F.var, A.var, A.up_count, &np);
                    c->halo_bytes += (long long)c->halo[level].n_export * 40;
                } else {
                    c->launches += k_restrict_fused(c->stream, c->H[level].n_nodes, A.child_ptr, A.child_idx, F.var, A.var, A.up_count);
                }
            }
            if (!fp && (rc = exchange(R, n, level, DAT_VAR))) return rc;
            if (level == nl - 1) dir = 1;
        } else {
            level--;
            for (int r = 0; r < n; r++) {
                mgcfd_ctx *c = R[r];
                cudaSetDevice(c->device);
                LevelDev &F = c->D[level], &A = c->D[level + 1];
                LoopScope t(c, "down", level, c->H[level].n_owned);
                if (fp) {
                    NodePush np;                  // reads the coarse level's halo residuals, pushes this level's variables
                    if ((rc = node_push_table(c, level, DAT_VAR, level + 1, np))) { ctx->err = c->err; return rc; }
                    c->launches += k_down(c->stream, c->H[level].n_owned, F.mg, F.var, F.res, F.coords, A.res, A.coords, &np);
                    c->halo_bytes += (long long)c->halo[level].n_export * 40;
                } else {
                    c->launches += k_down(c->stream, c->H[level].n_owned, F.mg, F.var, F.res, F.coords, A.res, A.coords);
                }
            }
            if (!fp && (rc = exchange(R, n, level, DAT_VAR))) return rc;
            if (level == 0) { dir = 0; i++; }
        }
    }
    return MGCFD_OK;
}

}  // namespace

// the stage kernels push and hand-shake themselves unless MGCFD_FUSED_PUSH=0 (the exchange then runs as separate pack /
// signal kernels on the communication stream, overlapped with the interior chunks); owner variant only
static bool fused_push_wanted(const mgcfd_ctx *ctx)
{
    const char *e = getenv("MGCFD_FUSED_PUSH");
    return !(e && atoi(e) == 0) && ctx->opt.flux_variant == MGCFD_FLUX_OWNER;
}

void mgcfd::cycle_drop_graphs(mgcfd_ctx *ctx)
{
    for (auto &kv : ctx->graphs) {
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        for (auto &sp : kv.second.spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
    }
    ctx->graphs.clear();
}

extern "C" {

int mgcfd_run_cycles(mgcfd_ctx *ctx, int n_cycles)
{
    if (!ctx) return MGCFD_ERR_ARG;
    REQUIRE(ctx->planned, "mgcfd_plan() has not been called");
    if (ctx->device < 0) { ctx->err = "planning-only context (device -1): compute entry points need a CUDA device"; return MGCFD_ERR_NODEVICE; }
    REQUIRE(n_cycles >= 0, "negative cycle count");
    CK(cudaSetDevice(ctx->device));
    if (ctx->n_ranks > 1 && (ctx->nccl_comm || ctx->p2p.ipc)) return run_ranks(&ctx, 1, n_cycles);
    REQUIRE(ctx->n_ranks == 1, "a partitioned context needs mgcfd_comm_init_nccl(), mgcfd_comm_init_ipc() or mgcfd_group_run_cycles()");
    return cycle_run_single(ctx, n_cycles);
}

int mgcfd_group_run_cycles(mgcfd_ctx **ranks, int n_ranks, int n_cycles)
{
    if (!ranks || n_ranks < 1 || n_ranks > 16 || !ranks[0]) return MGCFD_ERR_ARG;
    mgcfd_ctx *ctx = ranks[0];
    REQUIRE(n_cycles >= 0, "negative cycle count");
    for (int r = 0; r < n_ranks; r++) {
        REQUIRE(ranks[r] && ranks[r]->rank == r && ranks[r]->n_ranks == n_ranks, "ranks[r] must be the context of rank r of n_ranks");
        REQUIRE(!ranks[r]->nccl_comm, "group runs and NCCL contexts do not mix");
    }
    if (n_ranks == 1) return mgcfd_run_cycles(ctx, n_cycles);
    // peer access between the devices of the group (same-device groups need none)
    for (int a = 0; a < n_ranks; a++)
        for (int b = 0; b < n_ranks; b++) {
            if (ranks[a]->device == ranks[b]->device) continue;
            cudaSetDevice(ranks[a]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(ranks[b]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
                return MGCFD_ERR_CUDA;
            }
            cudaGetLastError();
        }
    return run_ranks(ranks, n_ranks, n_cycles);
}

int mgcfd_nccl_unique_id(void *id_out_128)
{
    if (!id_out_128 || !g_nccl.load()) return MGCFD_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return MGCFD_ERR_CUDA;
    memcpy(id_out_128, &id, sizeof(id));
    return MGCFD_OK;
}

int mgcfd_comm_init_nccl(mgcfd_ctx *ctx, int n_ranks, int rank, const void *unique_id_128)
{
    if (!ctx) return MGCFD_ERR_ARG;
    REQUIRE(unique_id_128 && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad arguments");
    REQUIRE(ctx->rank == rank && ctx->n_ranks == n_ranks, "context was not created for this rank");
    if (!g_nccl.load()) { ctx->err = g_nccl.error; return MGCFD_ERR_CUDA; }
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, unique_id_128, sizeof(id));
    ncclComm_t comm;
    NCK(g_nccl.CommInitRank(&comm, n_ranks, id, rank));
    ctx->nccl_comm = comm;
    return MGCFD_OK;
}

int mgcfd_group_enable_p2p(mgcfd_ctx **ranks, int n_ranks)
{
    if (!ranks || n_ranks < 2 || n_ranks > P2P_MAX_RANKS || !ranks[0]) return MGCFD_ERR_ARG;
    mgcfd_ctx *ctx = ranks[0];
    for (int r = 0; r < n_ranks; r++)
        REQUIRE(ranks[r] && ranks[r]->planned && ranks[r]->rank == r && ranks[r]->n_ranks == n_ranks && ranks[r]->p2p.arena,
                "ranks[r] must be the planned context of rank r of n_ranks");
    for (int a = 0; a < n_ranks; a++) {
        for (int b = 0; b < n_ranks; b++) {
            ranks[a]->p2p.peer[b] = ranks[b]->p2p.me;
            ranks[a]->p2p.peer_base[b] = ranks[b]->p2p.arena;
            if (ranks[a]->device != ranks[b]->device) {
                cudaSetDevice(ranks[a]->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(ranks[b]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
                    return MGCFD_ERR_CUDA;
                }
                cudaGetLastError();
            }
        }
        ranks[a]->p2p.enabled = true;
        ranks[a]->p2p.fused_push = fused_push_wanted(ranks[a]);
        cycle_drop_graphs(ranks[a]);
    }
    return MGCFD_OK;
}

int mgcfd_ipc_export(mgcfd_ctx *ctx, void *blob_out_4096)
{
    if (!ctx) return MGCFD_ERR_ARG;
    REQUIRE(blob_out_4096 && ctx->planned && ctx->p2p.arena, "mgcfd_ipc_export needs a planned, partitioned context");
    static_assert(sizeof(P2PInfo) <= 4096 && sizeof(cudaIpcMemHandle_t) == 64, "blob layout");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->p2p.arena));
    memcpy(ctx->p2p.me.ipc_handle, &h, 64);
    memset(blob_out_4096, 0, 4096);
    memcpy(blob_out_4096, &ctx->p2p.me, sizeof(P2PInfo));
    return MGCFD_OK;
}

int mgcfd_comm_init_ipc(mgcfd_ctx *ctx, const void *blobs)
{
    if (!ctx) return MGCFD_ERR_ARG;
    REQUIRE(blobs && ctx->planned && ctx->p2p.arena && ctx->n_ranks <= P2P_MAX_RANKS, "mgcfd_comm_init_ipc needs a planned, partitioned context");
    CK(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->n_ranks; r++) {
        memcpy(&ctx->p2p.peer[r], static_cast<const unsigned char *>(blobs) + (size_t)r * 4096, sizeof(P2PInfo));
        REQUIRE(ctx->p2p.peer[r].rank == r && ctx->p2p.peer[r].n_levels == ctx->n_levels, "blob r is not rank r's export");
        if (r == ctx->rank) { ctx->p2p.peer_base[r] = ctx->p2p.arena; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, ctx->p2p.peer[r].ipc_handle, 64);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->p2p.peer_base[r] = static_cast<unsigned char *>(p);
    }
    ctx->p2p.enabled = ctx->p2p.ipc = true;
    ctx->p2p.fused_push = fused_push_wanted(ctx);
    cycle_drop_graphs(ctx);
    return MGCFD_OK;
}

long long mgcfd_halo_bytes_sent(const mgcfd_ctx *ctx) { return ctx ? ctx->halo_bytes : 0; }

}  // extern "C"
