/*
 * mgcfd_b200.h -- C-ABI of the B200-native MG-CFD hot path (libmgcfd_b200.so).
 *
 * Drop-in boundary = the reference's op_par_loop call sites in euler3d.cpp.
 * OP2's code generator turns every op_par_loop into one host stub per kernel; this header
 * exports one entry point per such call site, plus the set/map/dat declaration, constant
 * declaration, fetch and teardown calls the driver makes around them.  Each declaration
 * cites the reference interface it replaces.  Plain pointers and sizes only; no C++ or
 * torch types cross this boundary.
 *
 * Conventions
 *   - every call returns MGCFD_OK (0) or a negative error code; mgcfd_last_error() gives the
 *     message.  No exceptions cross the ABI.  (Reference: print + op_exit() + return 1,
 *     euler3d.cpp:480-484, 544-548.)
 *   - host arrays are borrowed for the duration of the call only; the context owns all
 *     device storage (OP2 likewise owns dat storage after op_decl_*).
 *   - node/edge/boundary arrays are passed and fetched in FILE order; the planner's
 *     renumbering, blocking and partitioning are invisible to the caller.
 *   - loops are enqueued on the context's CUDA stream and are stream-ordered; only calls that
 *     return a value to the host (get_min_dt, calc_rms, count_bad_vals, fetch, sync) block.
 *   - there is no CPU fallback: without a CUDA device mgcfd_create() fails.
 */
#ifndef MGCFD_B200_H
#define MGCFD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGCFD_NVAR 5   /* const.h:41 */
#define MGCFD_NDIM 3   /* const.h:29 */
#define MGCFD_RK   3   /* const.h:31 */

enum {
    MGCFD_OK = 0,
    MGCFD_ERR_ARG = -1,        /* bad argument / call order */
    MGCFD_ERR_CUDA = -2,       /* CUDA runtime error (message has the details) */
    MGCFD_ERR_NODEVICE = -3,   /* no usable CUDA device: there is no CPU fallback */
    MGCFD_ERR_MIN_DT = -4,     /* min_dt < 0 (euler3d.cpp:480-484) */
    MGCFD_ERR_BAD_VALS = -5,   /* NaN/Inf in variables (euler3d.cpp:544-548) */
    MGCFD_ERR_PLAN = -6,       /* planner limit exceeded */
    MGCFD_ERR_COMM = -7        /* a peer rank did not answer a halo / min_dt exchange in time (bounded device-side wait ran out) */
};

/* flux-edge implementations (north_star: colouring scheme vs atomics, choice by measurement) */
enum {
    MGCFD_FLUX_ATOMIC = 0,     /* thread per edge, register accumulators, warp-aggregated fp64 RED */
    MGCFD_FLUX_COLOUR = 1,     /* OP2-style hierarchical colouring: edge blocks staged in shared memory,
                                  thread-colour-ordered shared accumulation, block-colour-ordered launches */
    MGCFD_FLUX_OWNER = 2,      /* owner-compute node chunks: cut edges recomputed, per-edge fluxes staged in
                                  shared memory, gathered per owned node, plain coalesced stores (default) */
    MGCFD_FLUX_GATHER = 3,     /* node gather over the same owner chunks: one thread per owned node evaluates its
                                  incident edges from its own side (interior edges twice), rows in sliced-ELL
                                  layout streamed coalesced, neighbour states gathered from shared memory */
    MGCFD_FLUX_EMIT = 4,       /* owner chunks, one thread per owned node: every edge evaluated once by its lowest owned
                                  endpoint, which keeps its own sum in registers; only the other owned end reads the
                                  edge's flux vector from shared memory.  Fast arithmetic only (sums not in file order) */
    MGCFD_FLUX_NVARIANTS = 5
};

typedef struct mgcfd_ctx mgcfd_ctx;

/* op_decl_const set, euler3d.cpp:232-238 (values computed at :47, :157-189) */
typedef struct {
    double smoothing_coefficient;
    double ff_variable[MGCFD_NVAR];
    double ff_flux_contribution_momentum_x[MGCFD_NDIM];
    double ff_flux_contribution_momentum_y[MGCFD_NDIM];
    double ff_flux_contribution_momentum_z[MGCFD_NDIM];
    double ff_flux_contribution_density_energy[MGCFD_NDIM];
    int mesh_name;
    int pad_;
} mgcfd_consts;

/* One multigrid level = the sets, maps and dats euler3d.cpp:248-312 declares from one level
 * file; field names are the HDF5 dataset names with "-->" spelled "_to_". */
typedef struct {
    int n_nodes;                       /* op_decl_set_hdf5_infer_size(.., "node_coordinates")   :251 */
    int n_edges;                       /* op_decl_set_hdf5_infer_size(.., "edge-->node")        :258 */
    int n_bnd_nodes;                   /* op_decl_set_hdf5_infer_size(.., "bnd_node-->node")    :265 */
    int n_owned_nodes;                 /* == n_nodes on one GPU; on a partition: nodes [0,n_owned) are owned,
                                          the rest are import halo (OP2 keeps this inside op_set) */
    const double *node_coordinates;    /* [n_nodes*3]   op_decl_dat_hdf5 :311 */
    const int    *edge_to_node;        /* [n_edges*2]   op_decl_map_hdf5 :273, base_array_index-based */
    const double *edge_weights;        /* [n_edges*3]   op_decl_dat_hdf5 :304 */
    const int    *bnd_node_to_node;    /* [n_bnd_nodes] op_decl_map_hdf5 :274, base_array_index-based */
    const int    *bnd_node_to_group;   /* [n_bnd_nodes] op_decl_dat_hdf5 :276 */
    const double *bnd_node_weights;    /* [n_bnd_nodes*3] op_decl_dat_hdf5 :306 */
    const int    *node_to_mg_node;     /* [n_nodes] map into level+1 (op_decl_map_hdf5 :283), NULL on the coarsest;
                                          on a partition a HALO node whose parent is not on this rank holds base-1 */
    /* ---- partition only (all NULL / 0 on one GPU); OP2 keeps these inside op_set / op_map after op_partition ---- */
    const int    *global_node_id;      /* [n_nodes] file index of each local node in the undecomposed mesh; fixes the
                                          summation order of restrict to the undecomposed one */
    int           n_neighbours;        /* ranks this rank exchanges halos with on this level */
    int           pad_;
    const int    *neighbour_rank;      /* [n_neighbours] ascending */
    const int    *export_ptr;          /* [n_neighbours+1] into export_idx */
    const int    *export_idx;          /* local (file) indices of owned nodes sent to each neighbour, in the
                                          neighbour's import order */
    const int    *import_ptr;          /* [n_neighbours+1] offsets, in nodes, into the halo range [n_owned, n_nodes) */
} mgcfd_level_host;

typedef struct {
    int flux_variant;      /* MGCFD_FLUX_*; default MGCFD_FLUX_OWNER */
    int renumber;          /* 1 (default): Hilbert-curve locality renumbering of nodes; 0: keep file order */
    int owner_chunk_nodes; /* owner/gather variants: max owned nodes per chunk (even, <= 256; default 64) */
    int colour_block_edges;/* colour variant: edges per block (default 256) */
    int exact_arith;       /* 1: reference operation order, IEEE div/sqrt, no FMA contraction in the flux kernels */
    int no_fusion;         /* 1: mgcfd_run_cycles launches one kernel per call site instead of the fused schedule
                              (fused Runge-Kutta stage, visit prologue and restrict; owner variant only) */
    int rank, n_ranks;     /* position of this context in a multi-GPU run (default 0 of 1) */
    int no_graphs;         /* 1: mgcfd_run_cycles enqueues every launch; 0 (default): each cycle replays as a
                              captured CUDA graph (also with NCCL: halo exchanges and the all-reduce are captured) */
    int measure_mem_bound; /* 1: mgcfd_run_cycles also runs unstructured_stream_kernel after every Runge-Kutta stage, as the
                              reference does with -b (euler3d.cpp:518-525); it writes p_dummy_fluxes only */
    int reserved[6];
} mgcfd_options;

/* ---- lifetime (op_init / op_exit, euler3d.cpp:126, :824) ---- */
void mgcfd_default_options(mgcfd_options *opt);
int  mgcfd_create(mgcfd_ctx **out, int device, int n_levels, const mgcfd_options *opt);
void mgcfd_destroy(mgcfd_ctx *ctx);
const char *mgcfd_last_error(const mgcfd_ctx *ctx);   /* ctx may be NULL: error of the last failed create */
const char *mgcfd_version(void);

/* ---- declarations ---- */
/* far-field constants exactly as euler3d.cpp:47,157-189 evaluates them */
void mgcfd_compute_farfield_consts(mgcfd_consts *out);
int  mgcfd_decl_consts(mgcfd_ctx *ctx, const mgcfd_consts *c);                        /* op_decl_const x7 :232-238 */
int  mgcfd_decl_level(mgcfd_ctx *ctx, int level, const mgcfd_level_host *lv, int base_array_index); /* :248-312 */
/* op_partition/op_renumber/op_decl_dat_temp_char point (:340-409): renumber, build the edge plans
 * (chunks / blocks / colours), allocate the temp dats variables, old_variables, residuals, volumes,
 * step_factors, fluxes, up_scratch zero-initialised */
int  mgcfd_plan(mgcfd_ctx *ctx);

/* ---- initialisation loops, euler3d.cpp:413-441 (run once; see DESIGN.md "init on host") ---- */
int mgcfd_loop_initialize_variables(mgcfd_ctx *ctx, int level);     /* :414 initialize_variables_kernel */
int mgcfd_loop_zero_fluxes(mgcfd_ctx *ctx, int level);              /* :416 zero_5d_array_kernel(p_fluxes) */
int mgcfd_loop_zero_volumes(mgcfd_ctx *ctx, int level);             /* :424 zero_1d_array_kernel(p_volumes) */
int mgcfd_loop_calculate_cell_volumes(mgcfd_ctx *ctx, int level);   /* :426-431 */
int mgcfd_loop_dampen_ewt_edges(mgcfd_ctx *ctx, int level);         /* :437 dampen_ewt(p_edge_weights) */
int mgcfd_loop_dampen_ewt_bnd(mgcfd_ctx *ctx, int level);           /* :439 dampen_ewt(p_bnd_node_weights) */

/* ---- the 14 live op_par_loop call sites of the cycle loop, euler3d.cpp:467-631 ---- */
int mgcfd_loop_copy_double(mgcfd_ctx *ctx, int level);                          /* :467-469 */
int mgcfd_loop_calculate_dt(mgcfd_ctx *ctx, int level);                         /* :472-475 */
int mgcfd_loop_get_min_dt(mgcfd_ctx *ctx, int level, double *min_dt);           /* :477-479 op_arg_gbl OP_MIN (in/out) */
int mgcfd_loop_compute_step_factor(mgcfd_ctx *ctx, int level, const double *min_dt); /* :485-489 op_arg_gbl OP_READ */
int mgcfd_loop_compute_flux_edge(mgcfd_ctx *ctx, int level);                    /* :498-503 */
int mgcfd_loop_compute_bnd_node_flux(mgcfd_ctx *ctx, int level);                /* :505-509 */
int mgcfd_loop_time_step(mgcfd_ctx *ctx, int level, const int *rkCycle);        /* :511-516 op_arg_gbl OP_READ */
int mgcfd_loop_unstructured_stream(mgcfd_ctx *ctx, int level);                  /* :518-525 (-b) into p_dummy_fluxes */
int mgcfd_loop_residual(mgcfd_ctx *ctx, int level);                             /* :528-531 */
int mgcfd_loop_calc_rms(mgcfd_ctx *ctx, int level, double *rms);                /* :534-536 op_arg_gbl OP_INC */
int mgcfd_loop_count_bad_vals(mgcfd_ctx *ctx, int level, int *count);           /* :540-542 op_arg_gbl OP_INC */
int mgcfd_loop_up_pre(mgcfd_ctx *ctx, int level_above);                         /* :581-583 set = nodes[level_above-1] */
int mgcfd_loop_up(mgcfd_ctx *ctx, int level_above);                             /* :585-588 */
int mgcfd_loop_up_post(mgcfd_ctx *ctx, int level_above);                        /* :590-592 */
int mgcfd_loop_down(mgcfd_ctx *ctx, int level);                                 /* :626-631 prolong level+1 -> level */

/* ---- domain decomposition (op_partition, euler3d.cpp:340-375; OP2 derives the halo lists internally) ----
 * Host-side and deterministic; every rank calls these on the undecomposed mesh and builds its own local mesh. */
typedef struct mgcfd_local_mesh mgcfd_local_mesh;
/* recursive coordinate bisection of level-0 nodes (the reference offers "INERTIAL"/"GEOM" methods, :373-375) */
int mgcfd_partition_rcb(int n_nodes, const double *node_coordinates, int n_parts, int *part_out);
/* op_partition's library / method selection (euler3d.cpp:340-375, config.h:203-240) on level 0: method "geom" /
 * "inertial" = mgcfd_partition_rcb; "kway" / "parmetis" / "ptscotch" / "geomkway" = recursive graph bisection refined by
 * Fiduccia-Mattheyses passes (smaller edge cut than the coordinate split); "block" = contiguous index ranges;
 * "random" = equal shares of a hashed node order.  Deterministic; MGCFD_ERR_ARG for an unknown method. */
int mgcfd_partition_graph(int n_nodes, const double *node_coordinates, int n_edges, const int *edge_to_node,
                          int base_array_index, int n_parts, const char *method, int *part_out);
/* coarse node -> owner of its lowest-numbered child; childless -> owner of the nearest edge neighbour with children */
int mgcfd_partition_coarse(int n_fine, const int *fine_part, const int *fine_to_coarse, int base_array_index, int n_coarse,
                           int n_coarse_edges, const int *coarse_edge_to_node, const double *coarse_coordinates,
                           int *coarse_part_out);
/* this rank's share of every level: [owned | import halo] nodes, edges with an owned endpoint, boundary entries of
 * owned nodes, 0-based local maps, export/import lists.  part[l][n] = owner of node n of level l. */
int mgcfd_local_mesh_build(int n_levels, const mgcfd_level_host *global_levels, int base_array_index,
                           const int *const *part, int rank, int n_ranks, mgcfd_local_mesh **out);
const mgcfd_level_host *mgcfd_local_mesh_level(const mgcfd_local_mesh *m, int level);   /* pass to mgcfd_decl_level(.., base 0) */
/* what: "global_node" "global_edge" "global_bnd" "neighbour_rank" "export_ptr" "export_idx" "import_ptr"
 * "edge_to_node" "node_to_mg_node"; returns the element count (out may be NULL) */
long long mgcfd_local_mesh_query(const mgcfd_local_mesh *m, int level, const char *what, int *out, long long capacity);
void mgcfd_local_mesh_free(mgcfd_local_mesh *m);

/* ---- multi-GPU execution.  Halo exchange of `variables` after every Runge-Kutta stage / restrict / prolong and
 * of `residuals` after every visit (OP2: dirty-bit halo exchanges inside op_par_loop), min_dt all-reduce(MIN).
 * Three transports:
 *   group  one process drives several contexts (one per GPU, or several on one GPU for tests): packed export
 *          buffers are pulled by the neighbour with peer copies, ordered by CUDA events;
 *   NCCL   one process per GPU (torchrun): grouped ncclSend/ncclRecv per neighbour + ncclAllReduce(min);
 *          the 128-byte unique id is created on rank 0 and distributed by the launcher. ---- */
int mgcfd_group_run_cycles(mgcfd_ctx **ranks, int n_ranks, int n_cycles);
int mgcfd_nccl_unique_id(void *id_out_128);
int mgcfd_comm_init_nccl(mgcfd_ctx *ctx, int n_ranks, int rank, const void *unique_id_128);
/*   p2p    direct peer stores: the pack kernel writes the exported rows straight into the neighbours' halo ranges
 *          (peer-mapped memory over NVLink), one warp publishes an epoch flag per neighbour and waits for theirs;
 *          min_dt travels through per-rank mailboxes the same way.  No library call on the data path.
 *          - ranks of one process: mgcfd_group_enable_p2p(ranks, n) then mgcfd_group_run_cycles;
 *          - one process per GPU: every rank calls mgcfd_ipc_export (a 4096-byte blob holding a CUDA IPC handle and
 *            the layout of its exchange arena), the launcher all-gathers the blobs, every rank calls
 *            mgcfd_comm_init_ipc(all blobs in rank order), then mgcfd_run_cycles. */
int mgcfd_group_enable_p2p(mgcfd_ctx **ranks, int n_ranks);
int mgcfd_ipc_export(mgcfd_ctx *ctx, void *blob_out_4096);
int mgcfd_comm_init_ipc(mgcfd_ctx *ctx, const void *blobs_n_ranks_x_4096);
/* bytes this rank has sent in halo exchanges so far */
long long mgcfd_halo_bytes_sent(const mgcfd_ctx *ctx);

/* ---- whole V-cycles on the device: the schedule of euler3d.cpp:458-641 with the per-visit host
 * checks (:480, :544) deferred to one flag read per call.  Results equal the loop-by-loop path. ---- */
int mgcfd_run_cycles(mgcfd_ctx *ctx, int n_cycles);

/* ---- op_fetch_data / test hooks ---- */
/* dat names: "variables" "old_variables" "residuals" "fluxes" "dummy_fluxes" (dim 5), "volumes"
 * "step_factors" (dim 1), "node_coordinates" (dim 3) -> double, n_nodes*dim, file node order;
 * "edge_weights" (double, n_edges*3, file edge order), "bnd_node_weights" (double, n_bnd*3),
 * "up_scratch" (int, n_nodes).  op_fetch_data_hdf5_file: euler3d.cpp:564,740-770 */
int mgcfd_fetch_dat(mgcfd_ctx *ctx, int level, const char *name, void *host_out);
int mgcfd_set_dat(mgcfd_ctx *ctx, int level, const char *name, const void *host_in);   /* tests / restart */
int mgcfd_sync(mgcfd_ctx *ctx);                                                        /* cudaStreamSynchronize */
/* page-locked host buffers: fetch/set move them with one DMA (pageable buffers bounce through an internal
 * pinned buffer).  OP2 has no counterpart; euler3d.cpp never touches dat storage directly. */
int  mgcfd_host_alloc(void **out, size_t bytes);
void mgcfd_host_free(void *p);

/* ---- -v validation path, euler3d.cpp:662-716 (identify_differences + count_non_zeros on device) ---- */
int mgcfd_validate_level(mgcfd_ctx *ctx, int level, const double *master_variables, int *n_differences);

/* ---- plan introspection (bit-exactness tests of renumbering / colouring / chunking) ---- */
/* what: "node_perm" (int[n_nodes]: internal index of file node i), "edge_order" (int[n_edges]: file edge at
 * sorted position i), "edge_block_colour" / "edge_thread_colour" (int[n_edges], colour variant, file edge
 * order), "n_block_colours" (int[1]), "owner_chunk_start" (int[n_chunks+1]), ...  Returns the number of
 * elements written, or a negative error; out may be NULL to query the count. */
long long mgcfd_plan_query(mgcfd_ctx *ctx, int level, const char *what, int *out, long long capacity);

/* ---- end-to-end call with HOST buffers ----
 * variables_in[l] / variables_out[l]: the flow state of level l in FILE order ([n_nodes][5], as op_decl_dat_hdf5 /
 * op_fetch_data_hdf5_file hold it, euler3d.cpp:379, :740-770); either array or single entries may be NULL (level not
 * uploaded / not fetched).  One call = mgcfd_set_dat("variables") for the given levels + mgcfd_run_cycles(n_cycles) +
 * mgcfd_fetch_dat("variables"), pipelined on one GPU when the buffers are page-locked (mgcfd_host_alloc): level 0 is
 * uploaded first and the first level visit starts as soon as it has landed, the coarser levels arrive underneath it on a
 * copy stream (each is first touched by the restrict into it); on the way down every coarse level is fetched as soon as
 * its last visit of the run has finished, underneath the remaining visits.  Same results and error codes as the three
 * separate calls. */
int  mgcfd_run_cycles_host(mgcfd_ctx *ctx, int n_cycles, const double *const *variables_in, double *const *variables_out);

/* ---- measurement hooks ---- */
/* per-call-site device timers (CUDA events on the context's stream): 0 off (default), 1 every call site,
 * 2 compute_flux_edge launches only (both: mgcfd_run_cycles then enqueues launch by launch instead of replaying graphs),
 * 3 every call site of mgcfd_run_cycles INSIDE CUDA-graph replay: the one-cycle graphs are captured with event-record
 *   nodes around each call site and read back after every replay */
int  mgcfd_timers_enable(mgcfd_ctx *ctx, int on);
int  mgcfd_timers_reset(mgcfd_ctx *ctx);
/* accumulated milliseconds / launch count / element count of a call site ("compute_flux_edge", ...) */
int  mgcfd_timers_get(mgcfd_ctx *ctx, const char *loop_name, int level, double *ms, long long *calls,
                      long long *elements);
long long mgcfd_kernel_launches(const mgcfd_ctx *ctx);      /* kernels launched by this library so far */
int  mgcfd_set_flux_variant(mgcfd_ctx *ctx, int variant);   /* switch implementation (plans built lazily) */
void *mgcfd_stream(mgcfd_ctx *ctx);                         /* cudaStream_t the loops are enqueued on */
void *mgcfd_device_ptr(mgcfd_ctx *ctx, int level, const char *name); /* raw device pointer of a dat (internal order) */

#ifdef __cplusplus
}
#endif
#endif /* MGCFD_B200_H */
