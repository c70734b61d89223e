/*
 * oracle_api.h -- C API shared by the two CPU checkers of the MG-CFD hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load these libraries, and only as the checker or the CPU baseline.
 *
 * Two libraries export exactly this API:
 *   oracle/_ref/libmgcfd_ref.so   ref_shim.cpp: the reference's OWN elemental kernel
 *                                 headers, #included in place from /root/reference
 *                                 (never copied), driven by op2_seq.inc.
 *   oracle/libmgcfd_oracle.so     mgcfd_oracle.c: a plain-C restatement ("port") of the
 *                                 same arithmetic, driven by the same op2_seq.inc.
 * tests/test_oracle_pin.py checks the two agree bit-for-bit on every loop and on full
 * multigrid runs; golden fixtures under tests/golden/ are produced by the _ref build.
 *
 * Loop executors restate OP2 *seq* semantics: elements visited
 * 0..n-1 in file order, indirect arguments resolved through 0-based maps, OP_INC
 * applied in place immediately, op_arg_gbl reduced in iteration order.
 */
#ifndef MGCFD_ORACLE_API_H
#define MGCFD_ORACLE_API_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NVAR 5
#define ORC_NDIM 3

/* One multigrid level: the datasets of one reference level file
 * (euler3d.cpp:248-312) plus the temp dats of euler3d.cpp:379-409.
 * All arrays are caller-owned (numpy).  Maps are 0-based. */
typedef struct {
    int n_nodes, n_edges, n_bnd;
    int pad_;
    double *coords;      /* node_coordinates  [n_nodes*3]                     */
    int    *e2n;         /* edge-->node       [n_edges*2]                     */
    double *ewt;         /* edge_weights      [n_edges*3]  (rewritten by init) */
    int    *b2n;         /* bnd_node-->node   [n_bnd]                         */
    int    *bgroup;      /* bnd_node-->group  [n_bnd]                         */
    double *bwt;         /* bnd_node_weights  [n_bnd*3]    (scaled by init)   */
    int    *mg;          /* node-->mg_node    [n_nodes] into the next coarser level; NULL on the coarsest */
    double *var;         /* p_variables       [n_nodes*5] */
    double *old;         /* p_old_variables   [n_nodes*5] */
    double *res;         /* p_residuals       [n_nodes*5] */
    double *flux;        /* p_fluxes          [n_nodes*5] */
    double *vol;         /* p_volumes         [n_nodes]   */
    double *sf;          /* p_step_factors    [n_nodes]   */
    int    *up_scratch;  /* p_up_scratch      [n_nodes*2]: int payload at 8-byte stride (euler3d.cpp:405) */
} orc_level;

typedef struct {
    double wall_total;        /* seconds around the cycle loop (euler3d.cpp:450,645-647)   */
    double wall_flux_edge;    /* seconds inside compute_flux_edge_kernel loops             */
    long long flux_edges;     /* edges processed by compute_flux_edge_kernel loops         */
    long long cycles;         /* completed multigrid cycles                                */
    double last_min_dt;
    double last_rms;
} orc_stats;

const char *orc_name(void);                 /* "reference" | "port" */
void orc_set_threads(int nthreads);         /* 1 = OP2-seq order (parity); >1 = OpenMP block-coloured (baseline timing only) */
int  orc_get_threads(void);

/* euler3d.cpp:157-189; out = {smoothing, ff_variable[5], ff_fc_mx[3], ff_fc_my[3], ff_fc_mz[3], ff_fc_de[3]} (18 doubles) */
void orc_set_farfield(double *out18);

/* individual op_par_loop call sites (cites are to euler3d.cpp) */
void orc_initialize_variables(int n, double *var);                                            /* :414 */
void orc_zero(int n_values, double *a);                                                       /* :416,:424 */
void orc_calculate_cell_volumes(int E, const int *e2n, const double *coords, double *ewt, double *vol); /* :426-431 */
void orc_dampen_ewt(int n, double *w);                                                        /* :436-441 */
void orc_copy_double(int n, const double *var, double *old);                                  /* :467-469 */
void orc_calculate_dt(int n, const double *var, const double *vol, double *sf);               /* :472-475 */
void orc_get_min_dt(int n, const double *sf, double *min_dt);                                 /* :477-479 */
void orc_compute_step_factor(int n, const double *var, const double *vol, const double *min_dt, double *sf); /* :485-489 */
void orc_compute_flux_edge(int E, const int *e2n, const double *var, const double *ewt, double *flux);      /* :498-503 */
void orc_compute_bnd_node_flux(int B, const int *bgroup, const double *bwt, const int *b2n,
                               const double *var, double *flux);                              /* :505-509 */
void orc_time_step(int n, int rk, const double *sf, double *flux, const double *old, double *var); /* :511-516 */
void orc_unstructured_stream(int E, const int *e2n, const double *var, const double *ewt, double *flux); /* :518-525 */
void orc_residual(int n, const double *old, const double *var, double *res);                  /* :528-531 */
void orc_calc_rms(int n, const double *res, double *rms);                                     /* :534-536 */
void orc_count_bad_vals(int n, const double *var, int *count);                                /* :540-542 */
void orc_up_pre(int n_fine, const int *mg, double *var_above, int *scratch_above);            /* :581-583 */
void orc_up(int n_fine, const int *mg, const double *var, double *var_above, int *scratch_above); /* :585-588 */
void orc_up_post(int n_coarse, double *var, const int *scratch);                              /* :590-592 */
void orc_down(int n_fine, const int *mg, double *var, const double *res, const double *coords,
              const double *res_above, const double *coords_above);                           /* :626-631 */
/* -v path, euler3d.cpp:662-716: returns number of differing values (count_non_zeros of identify_differences) */
int  orc_validate_count(int n, const double *test, const double *master);

/* euler3d.cpp:413-441 */
void orc_init_levels(orc_level *L, int n_levels);
/* euler3d.cpp:458-641.  returns 0, 1 (min_dt < 0, :480-484) or 2 (bad values, :544-548) */
int  orc_run_cycles(orc_level *L, int n_levels, int n_cycles, orc_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
