// internal.h -- shared declarations of libmgcfd_b200 (not part of the public C-ABI).
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "mgcfd_b200.h"

namespace mgcfd {

// default owner-kernel mode (MGCFD_OWNER_PIPE overrides): 0 one CTA per chunk, 1 persistent CTAs with one
// shared-memory stage and L2 prefetch of the next chunk, 2 persistent CTAs with two shared-memory stages
#ifndef MGCFD_OWNER_PIPE_DEFAULT
#define MGCFD_OWNER_PIPE_DEFAULT 0
#endif

constexpr int NVAR = MGCFD_NVAR;
constexpr int NDIM = MGCFD_NDIM;

// ------------------------------------------------------------------ programmatic dependent launch (MGCFD_PDL)
// The four kernels of the device-driven cycle (visit_begin, rk_stage2, restrict_fused, down) can be launched with the
// programmatic-stream-serialization attribute: a kernel's CTAs then become resident while the previous kernel's last
// wave drains, do everything that touches only static plan data (chunk record, mbarrier set-up, the bulk copies of the
// chunk's weight blob, index loads) and block in griddepcontrol.wait until the previous grid has completed and its
// memory is visible.  Every such kernel executes the wait in every CTA before its first access to mutable data, so
// completion stays transitive along the stream.  Without the launch attribute both instructions are no-ops.
// tl_pdl is set by the cycle drivers (cycle.cu: PdlScope) for the enqueue they are doing; the per-call-site API never sets it.
extern thread_local int tl_pdl;
#ifndef MGCFD_PDL_DEFAULT
#define MGCFD_PDL_DEFAULT 0
#endif
bool pdl_wanted();                       // MGCFD_PDL=0|1 (default MGCFD_PDL_DEFAULT)

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KA, typename... A>
inline cudaError_t launch_k(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A &&... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = tl_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}
#endif

// ------------------------------------------------------------------ host-side plans
// Sorted edge list shared by the atomic and colour variants: edges ordered by
// (min(internal a, internal b), max(..), file index).
struct SortedEdges {
    std::vector<int> order;      // order[i] = file edge at sorted position i
};

// OP2-style two-level colouring over blocks of consecutive sorted edges (SURVEY.md 8c "index-set oracle")
struct ColourPlanHost {
    int block_edges = 0;
    int n_blocks = 0, n_block_colours = 0;
    std::vector<int> thread_colour;   // [E] by sorted position
    std::vector<int> block_colour;    // [n_blocks] by block (sorted-position order)
    std::vector<int> block_ncol;      // [n_blocks] number of thread colours in the block
    // execution layout: blocks grouped by block colour
    std::vector<int> exec_block;      // [n_blocks] block id at execution slot
    std::vector<int> colour_start;    // [n_block_colours+1] into exec_block
    std::vector<int> node_off;        // [n_blocks+1] into node_gid (by execution slot)
    std::vector<int> node_gid;        // internal node ids of each block, ascending
    std::vector<uint32_t> lab;        // [E] by execution position: local a | local b << 16
    std::vector<int> exec_edge;       // [E] file edge at execution position
    std::vector<unsigned char> ecol;  // [E] thread colour at execution position
    int max_nodes = 0;
};

// Owner-compute chunks of consecutive internal nodes
struct OwnerPlanHost {
    int n_chunks = 0;
    std::vector<int> node0;           // [n_chunks+1] first owned internal node of each chunk
    std::vector<int> halo_off;        // [n_chunks+1] into halo_gid
    std::vector<int> halo_gid;        // internal ids of halo nodes per chunk, ascending
    std::vector<int> n_edges;         // [n_chunks]
    std::vector<int> n_inc;           // [n_chunks] owned incidences (CSR length)
    std::vector<long long> blob_off;  // [n_chunks] byte offset of the chunk's edge blob (16B aligned)
    std::vector<int> edge_file;       // file edge ids per chunk, concatenated in blob order
    std::vector<int> edge_off;        // [n_chunks+1] into edge_file
    std::vector<uint32_t> lab;        // per edge_file entry: local a | local b << 16
    std::vector<uint16_t> rowptr;     // per chunk (n_own+1), concatenated
    std::vector<int> rowptr_off;      // [n_chunks+1]
    std::vector<uint16_t> csr;        // per chunk n_inc entries: local edge | (is_b << 15)
    std::vector<int> csr_off;         // [n_chunks+1]
    int max_loc = 0, max_edges = 0, max_own = 0, max_inc = 0, max_blob = 0;
    long long total_edges = 0;
    // device packing (ensure_owner): the plan's blob followed by the chunk's boundary entries
    std::vector<long long> dev_blob_off;   // [n_chunks+1]
    int dev_max_blob = 0;
    int dev_max_tail = 0;                  // largest blob minus its four weight planes (stage2 kernel's runtime-sized part)
};

struct LevelHost {
    int n_nodes = 0, n_edges = 0, n_bnd = 0, n_owned = 0;
    std::vector<double> coords, ewt, bwt;     // file order
    std::vector<int> e2n, b2n, bgroup, mg;    // 0-based, file order (mg empty on the coarsest)
    std::vector<int> new_of_old, old_of_new;  // node renumbering
    std::vector<int> bnd_node_ptr;            // [n_owned+1] boundary entries per owned internal node
    std::vector<int> bnd_group_sorted;        // boundary entries sorted by (internal node, file index): group ...
    std::vector<double> bnd_wt_sorted;        // ... and weights [3]
    std::vector<int> global_node;             // partition: file index in the undecomposed mesh (else empty)
    std::vector<int> nbr_rank, export_ptr, export_idx, import_ptr;   // partition: halo lists in local file numbering
    SortedEdges sorted;
    ColourPlanHost colour;
    OwnerPlanHost owner;
    bool have_sorted = false, have_colour = false, have_owner = false;
};

// planner entry points (plan.cpp)
void plan_renumber(LevelHost &L, bool renumber);
void plan_sort_edges(LevelHost &L);
void plan_colour(LevelHost &L, int block_edges);
// returns false if a chunk cannot satisfy the local-index limits
bool plan_owner(LevelHost &L, int max_own, int max_loc, int max_edges, std::string &err);

// ------------------------------------------------------------------ device-side level
struct DevConsts {
    double smoothing;
    double ff_variable[NVAR];
    double ff_fc[NVAR][NDIM];   // rows 1..3 momentum x,y,z; row 4 density-energy; row 0 unused
};

struct AtomicPlanDev {
    int2 *nodes = nullptr;      // [E] internal (a, b)
    double4 *w = nullptr;       // [E] (wx, wy, wz, g) with g = -|w|*smoothing*0.5
    bool valid = false;
};

struct ColourPlanDev {
    int *blk_edge0 = nullptr;   // [n_blocks+1] by execution slot
    int *blk_node0 = nullptr;   // [n_blocks+1]
    int *blk_ncol = nullptr;    // [n_blocks]
    int *node_gid = nullptr;
    uint32_t *lab = nullptr;
    unsigned char *ecol = nullptr;
    double4 *w = nullptr;       // [E] by execution position
    bool valid = false;
};

struct OwnerChunkDesc {         // one per chunk, read by the kernel
    int node0, n_own, n_halo, halo_off;
    int n_edges, e_pad, n_inc, blob_bytes;
    long long blob_off;
    int has_bnd, bnd_off;        // boundary entries of the owned nodes (count) and where they sit in the blob:
                                 // bw [has_bnd][3] double | entry range per owned node [n_own+1] u16 (padded to even) |
                                 // group [has_bnd] i16
};

struct OwnerPlanDev {
    OwnerChunkDesc *desc = nullptr;
    int *halo_gid = nullptr;
    unsigned char *blob = nullptr;   // per chunk: w0[e_pad] w1 w2 g (double) | lab[e_pad] (u32) | rowptr | csr (u16) | boundary entries
    long long blob_bytes = 0;
    int *xtab = nullptr;             // lean kernel: per chunk [descriptor (12 ints) | halo ids, -1 padded (hs)] at stride xs
    int xs = 0, hs = 0;
    bool valid = false;
};

struct GatherChunkDesc {
    int node0, n_own, n_halo, halo_off;
    long long ent_off;              // first entry of the chunk in the sliced-ELL planes (multiple of 32)
    unsigned short slice_len[8];    // padded row length of each 32-row slice
};

struct GatherPlanDev {
    GatherChunkDesc *desc = nullptr;
    int *halo_gid = nullptr;
    uint16_t *row_node = nullptr, *row_deg = nullptr;   // [n_chunks*256] thread slot -> local owned node / its degree
    uint32_t *ent = nullptr;                             // neighbour local id | (this node is edge end b) << 16
    double *w0 = nullptr, *w1 = nullptr, *w2 = nullptr, *g = nullptr;
    long long n_ent = 0;
    bool valid = false;
};

struct EmitChunkDesc {
    int node0, n_own, n_halo, halo_off;
    long long blob_off;             // chunk blob: w0 w1 w2 g [n_ent] doubles | ent [n_ent] u32 | rowptr2 | csr2 (u16)
    int blob_bytes, n_ent;          // n_ent: padded sliced-ELL slots (multiple of 32)
    unsigned short slice_len[8];    // padded half-row length of each 32-thread slice
    int rowptr_pad, has_bnd;        // u16 entries before csr2 starts
};

struct EmitPlanDev {
    EmitChunkDesc *desc = nullptr;
    int *halo_gid = nullptr;
    uint16_t *row_node = nullptr, *row_cnt = nullptr;   // [n_chunks*256] thread slot -> local owned node / half-row length
    unsigned char *blob = nullptr;
    int n_chunks = 0, max_loc = 0, max_ent = 0, max_blob = 0, max_own = 0;
    bool valid = false;
};

struct LevelDev {
    double *var = nullptr, *old = nullptr, *res = nullptr, *flux = nullptr, *dummy_flux = nullptr;
    bool in_arena = false;         // var / var_alt / res live in the p2p arena (not freed individually)
    double *var_alt = nullptr;     // second variables buffer: the fused stage writes var_new here, then the two swap
    int *bnd_ptr = nullptr;        // [n_owned+1] boundary entry range per owned node (fused stage)
    int visit_parity = 0;          // which min_dt slot the next visit reduces into
    int var_flip = 0;              // whether var / var_alt are currently swapped (CUDA-graph cache key)
    double *vol = nullptr, *sf = nullptr, *coords = nullptr;
    int *up_count = nullptr;       // p_up_scratch payload
    int *mg = nullptr;             // internal fine node -> internal coarse node (level+1)
    int *child_ptr = nullptr;      // restrict gather CSR over this level's nodes: children on level-1
    int *child_idx = nullptr;
    // boundary entries grouped by unique node (ascending file order within a node)
    int n_bnd_unique = 0;
    int *bu_node = nullptr, *bu_ptr = nullptr, *b_group = nullptr;
    double *b_wt = nullptr;
    int *perm = nullptr;           // new_of_old: internal index of file node i
    double *cbrt_vol = nullptr;    // cbrt(volume) per node, evaluated once on the host (volumes are static after init)
    AtomicPlanDev atomic;
    ColourPlanDev colour;
    OwnerPlanDev owner;
    GatherPlanDev gather;
    EmitPlanDev emit;
    bool flux_is_zero = false;     // tracked so that the owner variant may overwrite instead of accumulate
};

// halo exchange lists of one level (internal node numbering; halo nodes keep their file positions)
struct HaloLevel {
    std::vector<int> nbr_rank, exp_ptr, imp_ptr;
    int *d_export_idx = nullptr;     // internal indices of exported owned nodes, concatenated per neighbour
    double *sendbuf = nullptr;       // [n_export][5]
    int n_export = 0;
    int *d_chunk_list = nullptr;     // owner chunks that own exported nodes first (n_boundary_chunks), then the rest
    int n_boundary_chunks = 0, n_chunks = 0;
    // fused push (the stage kernel stores exported rows straight into the neighbours' halo ranges): per chunk the base
    // of its (n_own + 1) row pointers in d_xp_ptr or -1, and per exported node its (destination slot, row) entries
    std::vector<int> xp_base;
    std::vector<int> launch_order;   // host copy of d_chunk_list
    int *d_xp_base = nullptr, *d_xp_ptr = nullptr;
    int2 *d_xp_ent = nullptr;
    std::vector<int> xn_ptr, xn_ent; // host copies (plan_query "export_node_ptr" / "export_node_ent": slot, row pairs)
    int *d_xn_ptr = nullptr;         // the same entries per owned node of the level (node kernels: restrict, prolong)
    int2 *d_xn_ent = nullptr;
};

struct GraphEntry {
    struct LevelState { double *var, *var_alt; int visit_parity, var_flip; };
    cudaGraphExec_t exec = nullptr;
    long long launches = 0, halo_bytes = 0;   // per replay (one cycle)
    std::vector<LevelState> after;            // host bookkeeping the cycle leaves behind
    // timers mode 3: event-record nodes captured around every call site (name#level, start, end, elements)
    struct TimedSpan { std::string key; cudaEvent_t e0, e1; long long elems; };
    std::vector<TimedSpan> spans;
};

// ---- direct peer-store halo exchange ("p2p" transport): every rank keeps variables (both buffers) and residuals of
// all levels, its flag words and the min_dt mailboxes in ONE arena that its peers map (CUDA IPC between processes,
// plain pointers inside one process) and write into.
constexpr int P2P_MAX_RANKS = 16, P2P_MAX_LEVELS = 16;
// Everything a stage kernel needs to push its exported rows and to hand-shake with the neighbours (device-resident, one
// per level x output buffer x with/without residuals; RkStageArgs::push points at the one a launch uses).
//   consumer side: a chunk that owns exported nodes (the only chunks that read rank-halo nodes) first waits until every
//                  source's flag has reached expected[s] -- the neighbours' rows of the previous exchange have landed and
//                  the neighbours have finished reading the halo ranges this stage is about to overwrite;
//   producer side: after its node phase the chunk stores var_new (and the residuals after the last stage) of its exported
//                  nodes into the destinations' halo ranges, fences, and counts itself in `done`; the last such chunk of the
//                  launch publishes the next epoch to every destination and advances `expected` for the next consumer.
struct StagePush {
    int n_dst, n_src, n_boundary, pad_;
    const int *xp_base, *xp_ptr;
    const int2 *xp_ent;
    double *var_dst[P2P_MAX_RANKS];                  // where my rows start in destination d's halo range (output buffer)
    double *res_dst[P2P_MAX_RANKS];                  // the same in its residual array, or null
    unsigned long long *dst_flag[P2P_MAX_RANKS];     // the destination's halo_flag[my rank]
    unsigned long long *sent[P2P_MAX_RANKS];         // my epoch counter towards that destination
    const unsigned long long *src_flag[P2P_MAX_RANKS];   // my halo_flag[source rank]
    unsigned long long *expected[P2P_MAX_RANKS];     // epoch the next consumer needs from that source
    unsigned int *done;                              // chunks with exports finished in this launch
    int *err_flag;                                   // d_flags[3]: a bounded wait ran out
    long long timeout_ns;
};
// The same hand-shake for the node kernels of a multi-rank cycle (visit prologue, step factor, restrict, prolong), passed by
// value: wait for the sources of the level whose halo rows the kernel READS, push the rows of the exported nodes the kernel
// WRITES, and let the last block publish the epoch (and arm the next consumer of the pushed level).
struct NodePush {
    int on, n_dst, n_src, n_wait;
    const int *xn_ptr;                               // [n_owned + 1] of the pushed level: entries per owned node
    const int2 *xn_ent;                              // (destination slot, row)
    double *dst[P2P_MAX_RANKS];                      // where my rows start in destination d's halo range of the pushed dat
    unsigned long long *dst_flag[P2P_MAX_RANKS], *sent[P2P_MAX_RANKS];
    const unsigned long long *src_flag[P2P_MAX_RANKS];
    unsigned long long *expected[P2P_MAX_RANKS];
    const unsigned long long *wait_flag[P2P_MAX_RANKS];
    const unsigned long long *wait_expected[P2P_MAX_RANKS];
    unsigned int *done;
    int *err_flag;
    long long timeout_ns;
};
// min_dt all-reduce folded into the visit prologue (the last block stores the rank's minimum into every peer's mailbox and
// publishes) and the step-factor kernel (every block waits for the peers' flags, then reads the mailboxes)
struct MinPush {
    int on, n_peers;
    unsigned long long *dst_box[P2P_MAX_RANKS], *dst_flag[P2P_MAX_RANKS], *sent[P2P_MAX_RANKS];
    const unsigned long long *src_flag[P2P_MAX_RANKS];
    unsigned long long *expected[P2P_MAX_RANKS];
    unsigned int *done;
    int *err_flag;
    long long timeout_ns;
};

struct WaitTable {                       // the sources of a level: (flag, expected epoch) pairs
    int n_src;
    const unsigned long long *src_flag[P2P_MAX_RANKS];
    const unsigned long long *expected[P2P_MAX_RANKS];
    int *err_flag;
    long long timeout_ns;
};

struct P2PInfo {                         // what a peer must know about a rank's arena (exchanged as an opaque blob)
    unsigned char ipc_handle[64];
    int n_levels, rank;
    long long off_var[2][P2P_MAX_LEVELS], off_res[P2P_MAX_LEVELS];
    long long off_flags;                 // u64 halo_flag[P2P_MAX_RANKS] | u64 min_flag[P2P_MAX_RANKS] |
                                         // u64 min_box[P2P_MAX_RANKS][P2P_MAX_LEVELS][2] | u64 status_box[P2P_MAX_RANKS]
    int n_owned[P2P_MAX_LEVELS];
    int import_off[P2P_MAX_LEVELS][P2P_MAX_RANKS];   // where rows from rank r land in the halo range (nodes), -1: none
    int import_cnt[P2P_MAX_LEVELS][P2P_MAX_RANKS];
};
struct P2PState {
    bool arena_owner = false, enabled = false, ipc = false;
    unsigned char *arena = nullptr;
    size_t arena_bytes = 0;
    P2PInfo me{};
    P2PInfo peer[P2P_MAX_RANKS];
    unsigned char *peer_base[P2P_MAX_RANKS] = {};
    unsigned long long *d_counters = nullptr;   // sent_halo[16] | expected_halo[16] | sent_min[16] | expected_min[16]
    std::vector<StagePush> h_push;              // [n_levels][2 output buffers][without / with residuals]
    unsigned int *d_done = nullptr;             // [0] chunks / blocks finished, [1] pusher CTAs finished
    bool fused_push = false;                    // the stage kernels push and hand-shake themselves (default with p2p)
};

struct PushTable {                       // kernel parameter of one halo push
    int n_dst, n_src;
    int exp_ptr[P2P_MAX_RANKS + 1];
    double *dst[P2P_MAX_RANKS];                      // where my rows land in each destination's arena
    unsigned long long *dst_flag[P2P_MAX_RANKS];     // the destination's halo_flag[my rank]
    unsigned long long *sent[P2P_MAX_RANKS];         // my epoch counter towards that destination
    const unsigned long long *src_flag[P2P_MAX_RANKS];   // my halo_flag[source rank]
    unsigned long long *expected[P2P_MAX_RANKS];     // my epoch counter for that source
    int *err_flag;
    long long timeout_ns;
};
struct MinTable {
    int n_peers, me, parity;
    unsigned long long *dst_box[P2P_MAX_RANKS];      // peer's min_box[me][parity]
    unsigned long long *dst_flag[P2P_MAX_RANKS];     // peer's min_flag[me]
    unsigned long long *sent[P2P_MAX_RANKS];
    const unsigned long long *src_flag[P2P_MAX_RANKS];
    unsigned long long *expected[P2P_MAX_RANKS];
    int *err_flag;
    long long timeout_ns;
};

struct LoopTimer {
    double ms = 0.0;
    long long calls = 0, elements = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    std::vector<long long> pending_elems;
};

}  // namespace mgcfd

struct mgcfd_ctx {
    int device = 0, n_levels = 0;
    mgcfd_options opt{};
    mgcfd_consts consts{};
    bool have_consts = false, planned = false;
    std::vector<mgcfd::LevelHost> H;
    std::vector<mgcfd::LevelDev> D;
    cudaStream_t stream = nullptr;
    std::string err;
    // device scalars
    double *d_min_dt = nullptr;      // [n_levels] scratch for reductions
    unsigned long long *d_min_enc = nullptr;   // [n_levels][2] order-encoded min_dt slots of the fused path
    double *d_rms = nullptr;
    int *d_flags = nullptr;          // [0]=bad value count, [1]=min_dt<0 flag, [2]=validate count
    double *h_pinned = nullptr;      // pinned host scratch (16 doubles)
    void *d_stage = nullptr, *h_stage = nullptr;   // device / pinned-host staging for file-order transfers
    size_t d_stage_bytes = 0, h_stage_bytes = 0;
    long long launches = 0;
    int timers_on = 0;                // 0 off, 1 every call site, 2 compute_flux_edge only (both: launch by launch, no graphs),
                                      // 3 every call site INSIDE CUDA-graph replay (event-record nodes in the captured graph)
    std::vector<mgcfd::GraphEntry::TimedSpan> capture_spans;   // spans recorded by the capture in progress (mode 3)
    std::map<std::string, mgcfd::LoopTimer> timers;
    std::vector<cudaEvent_t> event_pool;
    std::map<unsigned, mgcfd::GraphEntry> graphs;   // captured one-cycle graphs by parity state
    // mgcfd_run_cycles_host in flight: per level the event its upload records (waited for before the level's first use)
    // and, in the run's last cycle, the event recorded when the level's variables are final (its download waits for it)
    struct IoHooks { std::vector<cudaEvent_t> uploaded, final_; std::vector<char> wait_upload, want_final; int last_cycle = 0; };
    IoHooks *io = nullptr;
    std::vector<double *> io_stage;   // per-level device staging of mgcfd_run_cycles_host
    // multi-GPU
    std::vector<mgcfd::HaloLevel> halo;
    int rank = 0, n_ranks = 1;
    void *nccl_comm = nullptr;
    cudaEvent_t ev_pack = nullptr, ev_done = nullptr, ev_k1 = nullptr, ev_prod = nullptr, ev_ready = nullptr;
    cudaStream_t comm_stream = nullptr;   // halo exchanges run here, overlapped with interior chunks on `stream`
    long long halo_bytes = 0;
    mgcfd::P2PState p2p;
};

namespace mgcfd {

// helpers of api.cu used by the cycle driver (cycle.cu)
int api_check_launch(mgcfd_ctx *ctx, const char *what);
DevConsts api_dev_consts(const mgcfd_ctx *ctx);
int api_ensure_flux_plan(mgcfd_ctx *ctx, int level);
int api_run_flux(mgcfd_ctx *ctx, int level, bool stream_kernel);
int api_ensure_dummy_flux(mgcfd_ctx *ctx);
int cycle_enqueue_single_nograph(mgcfd_ctx *ctx, int n_cycles);
int cycle_finish_run(mgcfd_ctx *ctx);
void timers_collect(mgcfd_ctx *ctx);
int cycle_run_single(mgcfd_ctx *ctx, int n_cycles);
void cycle_drop_graphs(mgcfd_ctx *ctx);

// per-call-site device timer (CUDA events on the context's stream)
struct LoopScope {
    mgcfd_ctx *ctx;
    LoopTimer *t = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    long long elems;
    static cudaEvent_t get_event(mgcfd_ctx *ctx)
    {
        if (!ctx->event_pool.empty()) {
            cudaEvent_t e = ctx->event_pool.back();
            ctx->event_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    cudaStream_t stream = nullptr;
    std::string key;
    bool in_graph = false;
    LoopScope(mgcfd_ctx *c, const char *name, int level, long long elements, cudaStream_t on = nullptr) : ctx(c), elems(elements)
    {
        if (!ctx->timers_on) return;
        if (ctx->timers_on == 2 && strcmp(name, "compute_flux_edge") != 0 && strcmp(name, "rk_stage") != 0)
            return;   // flux-edge launches (stand-alone or as the fused Runge-Kutta stage) only
        stream = on ? on : ctx->stream;
        key = std::string(name) + "#" + std::to_string(level);
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cs);
        in_graph = cs == cudaStreamCaptureStatusActive;
        if (ctx->timers_on == 3 && !in_graph) return;       // mode 3 times graph replays only
        t = &ctx->timers[key];
        if (in_graph) {
            // event-record NODES in the captured graph: the events are re-recorded by every replay
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecordWithFlags(e0, stream, cudaEventRecordExternal);
        } else {
            e0 = get_event(ctx);
            e1 = get_event(ctx);
            cudaEventRecord(e0, stream);
        }
    }
    ~LoopScope()
    {
        if (!t) return;
        if (in_graph) {
            cudaEventRecordWithFlags(e1, stream, cudaEventRecordExternal);
            ctx->capture_spans.push_back({key, e0, e1, elems});
            return;
        }
        cudaEventRecord(e1, stream);
        t->pending.push_back({e0, e1});
        t->pending_elems.push_back(elems);
    }
};


// kernels.cu / flux_*.cu launchers.  All enqueue on `s` and return the number of kernels launched.
int k_copy(cudaStream_t s, int n, const double *var, double *old);
int k_calculate_dt(cudaStream_t s, int n, const double *var, const double *cbrt_vol, double *sf);
int k_min_dt(cudaStream_t s, int n, const double *sf, double *d_min, int *d_flags);  // d_min must hold the start value
int k_step_factor(cudaStream_t s, int n, const double *vol, const double *d_min, double *sf);
int k_time_step(cudaStream_t s, int n, int rk, const double *sf, double *flux, const double *old, double *var);
int k_residual(cudaStream_t s, int n, const double *old, const double *var, double *res);
int k_rms(cudaStream_t s, int n, const double *res, double *d_rms);
int k_bad_vals(cudaStream_t s, int n, const double *var, int *d_count);
int k_up_pre(cudaStream_t s, int n_fine, const int *mg, double *var_above, int *count_above);
int k_up(cudaStream_t s, int n_coarse, const int *child_ptr, const int *child_idx, const double *var,
         double *var_above, int *count_above);
int k_up_post(cudaStream_t s, int n_coarse, double *var, const int *count);
int k_down(cudaStream_t s, int n_fine, const int *mg, double *var, const double *res, const double *coords,
           const double *res_above, const double *coords_above, const NodePush *np = nullptr);
int k_bnd_flux(cudaStream_t s, int n_unique, const int *bu_node, const int *bu_ptr, const int *b_group,
               const double *b_wt, const double *var, double *flux, const DevConsts &c, bool exact);
int k_validate(cudaStream_t s, int n, const double *test, const double *master, int *d_count);
int k_fill(cudaStream_t s, long long n, double *a, double v);
// to_internal: dst[perm[i]] = src[i] (file -> internal); else dst[i] = src[perm[i]] (internal -> file)
int k_permute_rows(cudaStream_t s, int n, int dim, const double *src, const int *perm, double *dst, bool to_internal);
int k_init_vars(cudaStream_t s, int n, double *var, const DevConsts &c);
// fused node kernels of mgcfd_run_cycles
int k_visit_begin(cudaStream_t s, int n, const double *var, const double *cbrt_vol, double *old, double *dt,
                  unsigned long long *min_slot, const MinPush *mp = nullptr, double *zero_me = nullptr);
int k_step_factor_fused(cudaStream_t s, int n, const double *vol, unsigned long long *min_slot, unsigned long long *next_slot,
                        double *sf, double *d_min_out, int *d_flags);
int k_restrict_fused(cudaStream_t s, int n_coarse, const int *child_ptr, const int *child_idx, const double *var,
                     double *var_above, int *count_above, const NodePush *np = nullptr);
int k_reset_min_slots(cudaStream_t s, int n, unsigned long long *slots);
struct MinSlots { const unsigned long long *p[16]; int n; };
int k_step_factor_group(cudaStream_t s, int n, const double *vol, MinSlots slots, unsigned long long *next_slot, double *sf,
                        double *d_min_out, int *d_flags, const MinPush *mp = nullptr);
int k_pack_rows(cudaStream_t s, int n, const int *idx, const double *src, double *dst);
// p2p transport: rows straight into the peers' halo ranges, then epoch flags; returns kernels launched
int k_push_rows(cudaStream_t s, int n_rows, const int *idx, const double *src, const PushTable &t);
int k_signal_wait(cudaStream_t s, const PushTable &t);
int k_min_exchange(cudaStream_t s, const unsigned long long *my_slot, const MinTable &t);
int k_status_exchange(cudaStream_t s, int *flags, const unsigned long long *boxes, int n_ranks, const MinTable &t);

// compute_step_factor_kernel folded into the first Runge-Kutta stage of a visit (stage2 kernel): every chunk takes the minimum
// over the reduced slots (its own and, multi-rank, the peers' mailboxes -- after their flags have arrived), divides by its
// nodes' volumes, uses the result and stores it as `step_factors` for the later stages
struct StepFold {
    int on, n_slots, n_wait, pad_;
    const unsigned long long *slot[P2P_MAX_RANKS];
    const unsigned long long *wait_flag[P2P_MAX_RANKS];
    const unsigned long long *wait_expected[P2P_MAX_RANKS];
    unsigned long long *next_slot;
    double *d_min_out;
    int *d_flags;
    const double *vol;
    double *sf_out;
    long long timeout_ns;
};

// extra arguments of the fused Runge-Kutta stage (flux + boundary flux + time_step [+ residual, rms, bad values])
struct RkStageArgs {
    const double *old, *sf;
    double *var_out, *res;
    double *d_rms;               // non-null on level 0: sum of squared residuals / bad-value count after the last stage
    int *d_bad;
    const int *bnd_ptr;          // per owned node: range into the boundary entry arrays
    const int *b_group;
    const double *b_wt;
    int rk, last;
    int push_on, pad2_;          // multi-GPU: push exported rows from the node phase and hand-shake (fused push)
    StagePush push;
    StepFold fold;               // first stage of a visit: step factors computed in the kernel (stage2 only)
    int max_own, pad_;           // filled by the launcher: tile sizes of the prefetched update operands
    double inv_denom;            // filled by the launcher: 1 / (RK + 1 - rk) (stage2 kernel)
    DevConsts c;
};

struct FluxArgs {
    int n_edges = 0, n_owned = 0, n_nodes = 0;
    const double *var = nullptr;
    double *flux = nullptr;
    bool stream_kernel = false;     // unstructured_stream_kernel body instead of the flux body
    bool overwrite = false;         // owner variant: flux known to be zero -> plain store, no read
    const RkStageArgs *rk = nullptr; // owner variant: fuse the rest of the Runge-Kutta stage into the kernel
    const int *chunk_list = nullptr; // owner variant: launch only these chunks (device array of n_list chunk ids)
    int n_list = 0;
    int list_offset = 0;             // position of chunk_list[0] in the level's launch order (the xtab records are stored in that order)
};
int flux_atomic(cudaStream_t s, const FluxArgs &a, const AtomicPlanDev &p, bool exact);
int flux_colour(cudaStream_t s, const FluxArgs &a, const ColourPlanDev &p, const ColourPlanHost &h, bool exact);
int flux_owner(cudaStream_t s, const FluxArgs &a, const OwnerPlanDev &p, const OwnerPlanHost &h, bool exact);
// whether a fused stage on this plan runs the stage2 kernel (which can compute the step factors itself)
bool flux_owner_uses_stage2(const OwnerPlanDev &p, const OwnerPlanHost &h, bool exact);
int flux_gather(cudaStream_t s, const FluxArgs &a, const GatherPlanDev &p, int n_chunks, int max_loc, bool exact);
int fast_flux_gather(cudaStream_t s, const FluxArgs &a, const GatherPlanDev &p, int n_chunks, int max_loc);
size_t flux_gather_smem_bytes(int max_loc, bool exact);
int flux_emit(cudaStream_t s, const FluxArgs &a, const EmitPlanDev &p);      // fast arithmetic only
size_t flux_emit_smem_bytes(int max_loc, int max_ent, int max_blob, int max_own);
// one-time kernel attribute setup (dynamic shared memory opt-in); returns "" or an error text
std::string flux_configure();
// the fast-math translation unit (flux_fast.cu)
int fast_flux_atomic(cudaStream_t s, const FluxArgs &a, const AtomicPlanDev &p);
int fast_flux_colour(cudaStream_t s, const FluxArgs &a, const ColourPlanDev &p, const ColourPlanHost &h);
int fast_flux_owner(cudaStream_t s, const FluxArgs &a, const OwnerPlanDev &p, const OwnerPlanHost &h);
int fast_bnd_flux(cudaStream_t s, int n_unique, const int *bu_node, const int *bu_ptr, const int *b_group,
                  const double *b_wt, const double *var, double *flux, const DevConsts &c);
std::string fast_configure();
size_t flux_owner_smem_bytes(int max_loc, int max_edges, int max_blob, bool exact);
size_t flux_colour_smem_bytes(int max_nodes, bool exact);

}  // namespace mgcfd
