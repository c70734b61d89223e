/*
 * mgcfd_oracle.c -- plain-C restatement ("port") of the MG-CFD hot-path arithmetic.
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h): the CPU checker for the CUDA path.
 *
 * Pinning: tests/test_oracle_pin.py checks every loop and full multigrid runs of this
 * file bit-for-bit against oracle/_ref/libmgcfd_ref.so (the reference's own headers
 * compiled in place) and against the committed fixtures in tests/golden/ that the
 * _ref build produced (oracle/gen_golden.py).  The reference tree itself holds no
 * golden vectors or unit tests.
 *
 * Every function cites the reference lines it follows.  Floating-point operation
 * ORDER is part of the contract (the comparison is bit-exact), so expressions keep
 * the reference's association; build with -ffp-contract=off.
 */
#include <math.h>
#include "oracle_api.h"

#define GAMMA_ 1.4            /* const.h:27 */
enum { RHO = 0, MX = 1, MY = 2, MZ = 3, ENE = 4 };   /* const.h:38-41 */

/* kernel-visible constants (global.h:6-11; set at euler3d.cpp:47,157-189) */
static double k_smoothing;
static double k_ff_var[5];
static double k_ff_fc[5][3];  /* [variable][direction]; row RHO unused */

/* Per-node quantities every flux kernel derives from the 5 conserved variables. */
typedef struct {
    double rho, m[3], ene;
    double v[3];          /* inlined_funcs.h:120-125 */
    double speed_sqd;     /* inlined_funcs.h:98-101  */
    double p;             /* inlined_funcs.h:103-106 */
    double c;             /* inlined_funcs.h:126-129 */
    double fc[5][3];      /* inlined_funcs.h:70-96: fc[MX..MZ][d] momentum rows, fc[ENE][d] energy row */
} node_state;

/* inlined_funcs.h:70-96 */
static void flux_contribution(const double m[3], double ene, double p, const double v[3], double fc[5][3])
{
    fc[MX][0] = v[0] * m[0] + p;
    fc[MX][1] = v[0] * m[1];
    fc[MX][2] = v[0] * m[2];
    fc[MY][0] = fc[MX][1];
    fc[MY][1] = v[1] * m[1] + p;
    fc[MY][2] = v[1] * m[2];
    fc[MZ][0] = fc[MX][2];
    fc[MZ][1] = fc[MY][2];
    fc[MZ][2] = v[2] * m[2] + p;
    {
        double ep = ene + p;
        fc[ENE][0] = v[0] * ep;
        fc[ENE][1] = v[1] * ep;
        fc[ENE][2] = v[2] * ep;
    }
}

/* the block flux.h:52-91 (side b), :101-136 (side a), flux_boundary.elem_func:9-47,
 * flux_wall.elem_func:8-46 and time_stepping_kernels.h:18-29 all evaluate */
static void derive_state(const double *u, node_state *s, int want_fc)
{
    int d;
    s->rho = u[RHO];
    s->m[0] = u[MX]; s->m[1] = u[MY]; s->m[2] = u[MZ];
    s->ene = u[ENE];
    for (d = 0; d < 3; d++) s->v[d] = s->m[d] / s->rho;
    s->speed_sqd = s->v[0] * s->v[0] + s->v[1] * s->v[1] + s->v[2] * s->v[2];
    s->p = (GAMMA_ - 1.0) * (s->ene - 0.5 * s->rho * s->speed_sqd);
    s->c = sqrt(GAMMA_ * s->p / s->rho);
    if (want_fc) flux_contribution(s->m, s->ene, s->p, s->v, s->fc);
}

/* euler3d.cpp:157-189 */
static void lib_set_farfield(double *out18)
{
    const double pi = 3.1415926535897931;                      /* const.h:43 */
    const double aoa = (pi / 180.0) * 0.0;                     /* const.h:33 */
    double ff_p = 1.0, ff_c, ff_speed, ff_v[3], ff_m[3];
    int i;
    k_smoothing = (double)0.2f;                                /* euler3d.cpp:47 */
    k_ff_var[RHO] = 1.4;
    ff_c = sqrt(GAMMA_ * ff_p / k_ff_var[RHO]);
    ff_speed = 1.2 * ff_c;                                     /* const.h:32 */
    ff_v[0] = ff_speed * cos(aoa);
    ff_v[1] = ff_speed * sin(aoa);
    ff_v[2] = 0.0;
    for (i = 0; i < 3; i++) k_ff_var[MX + i] = k_ff_var[RHO] * ff_v[i];
    k_ff_var[ENE] = k_ff_var[RHO] * (0.5 * (ff_speed * ff_speed)) + (ff_p / (GAMMA_ - 1.0));
    for (i = 0; i < 3; i++) ff_m[i] = k_ff_var[MX + i];
    flux_contribution(ff_m, k_ff_var[ENE], ff_p, ff_v, k_ff_fc);
    if (out18) {
        out18[0] = k_smoothing;
        for (i = 0; i < 5; i++) out18[1 + i] = k_ff_var[i];
        for (i = 0; i < 3; i++) {
            out18[6 + i] = k_ff_fc[MX][i];
            out18[9 + i] = k_ff_fc[MY][i];
            out18[12 + i] = k_ff_fc[MZ][i];
            out18[15 + i] = k_ff_fc[ENE][i];
        }
    }
}

/* ---- misc.h ---- */
static void orc_k_initialize_variables_kernel(double *u)           /* misc.h:10-16 */
{
    int j;
    for (j = 0; j < 5; j++) u[j] = k_ff_var[j];
}
static void orc_k_zero_1d_array_kernel(double *a) { *a = 0.0; }     /* misc.h:34-38 */

static void orc_k_calculate_cell_volumes(const double *c1, const double *c2, double *w,
                                         double *vol1, double *vol2)   /* misc.h:40-76 */
{
    double d[3], dist = 0.0, area = 0.0, tet;
    int i;
    for (i = 0; i < 3; i++) { d[i] = c2[i] - c1[i]; dist += d[i] * d[i]; }
    dist = sqrt(dist);
    for (i = 0; i < 3; i++) area += w[i] * w[i];
    area = sqrt(area);
    tet = (1.0 / 3.0) * 0.5 * dist * area;
    *vol1 += tet;
    *vol2 += tet;
    for (i = 0; i < 3; i++) w[i] = (d[i] / dist) * area;   /* re-aim along the edge ... */
    for (i = 0; i < 3; i++) w[i] /= dist;                   /* ... magnitude area/dist   */
}
static void orc_k_dampen_ewt(double *w) { w[0] *= 1e-7; w[1] *= 1e-7; w[2] *= 1e-7; }  /* misc.h:78-84 */

/* ---- copy_double_kernel.h:6-13 ---- */
static void orc_k_copy_double_kernel(const double *u, double *old)
{
    int i;
    for (i = 0; i < 5; i++) old[i] = u[i];
}

/* ---- time_stepping_kernels.h ---- */
static void orc_k_calculate_dt_kernel(const double *u, const double *vol, double *dt)   /* :13-32 */
{
    node_state s;
    derive_state(u, &s, 0);
    *dt = 0.5 * (cbrt(*vol) / (sqrt(s.speed_sqd) + s.c));
}
static void orc_k_get_min_dt_kernel(const double *dt, double *min_dt)                  /* :34-41 */
{
    if (*dt < *min_dt) *min_dt = *dt;
}
static void orc_k_compute_step_factor_kernel(const double *u, const double *vol, const double *min_dt,
                                             double *sf)                                /* :43-64 */
{
    (void)u;                       /* :49-60 recompute unused quantities */
    *sf = (*min_dt) / (*vol);
}
static void orc_k_time_step_kernel(const int *rk, const double *sf, double *flux, const double *old,
                                   double *u)                                           /* :66-86 */
{
    double factor = (*sf) / (double)(3 + 1 - (*rk));
    int i;
    for (i = 0; i < 5; i++) u[i] = old[i] + factor * flux[i];
    for (i = 0; i < 5; i++) flux[i] = 0.0;
}

/* ---- flux.h:41-208 ---- */
static void orc_k_compute_flux_edge_kernel(const double *ua, const double *ub, const double *w,
                                           double *fa, double *fb)
{
    node_state A, B;
    double ewt, factor_a, factor_b, f[3], speed_a, speed_b;
    int i;
    ewt = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);          /* :48-50 */
    derive_state(ub, &B, 1);                                       /* :52-91 */
    speed_b = sqrt(B.speed_sqd);
    derive_state(ua, &A, 1);                                       /* :101-136 */
    speed_a = sqrt(A.speed_sqd);
    factor_a = -ewt * k_smoothing * 0.5 * (speed_a + sqrt(B.speed_sqd) + A.c + B.c);   /* :139-141 */
    factor_b = -ewt * k_smoothing * 0.5 * (speed_b + sqrt(A.speed_sqd) + B.c + A.c);   /* :143-145 */
    for (i = 0; i < 3; i++) f[i] = -0.5 * w[i];                    /* :147 */

    /* :149-177 */
    fa[RHO] += factor_a * (A.rho - B.rho)
             + f[0] * (A.m[0] + B.m[0]) + f[1] * (A.m[1] + B.m[1]) + f[2] * (A.m[2] + B.m[2]);
    fa[ENE] += factor_a * (A.ene - B.ene)
             + f[0] * (A.fc[ENE][0] + B.fc[ENE][0]) + f[1] * (A.fc[ENE][1] + B.fc[ENE][1])
             + f[2] * (A.fc[ENE][2] + B.fc[ENE][2]);
    for (i = 0; i < 3; i++)
        fa[MX + i] += factor_a * (A.m[i] - B.m[i])
                    + f[0] * (A.fc[MX + i][0] + B.fc[MX + i][0]) + f[1] * (A.fc[MX + i][1] + B.fc[MX + i][1])
                    + f[2] * (A.fc[MX + i][2] + B.fc[MX + i][2]);
    /* :179-207 */
    fb[RHO] += factor_b * (B.rho - A.rho)
             - f[0] * (A.m[0] + B.m[0]) - f[1] * (A.m[1] + B.m[1]) - f[2] * (A.m[2] + B.m[2]);
    fb[ENE] += factor_b * (B.ene - A.ene)
             - f[0] * (A.fc[ENE][0] + B.fc[ENE][0]) - f[1] * (A.fc[ENE][1] + B.fc[ENE][1])
             - f[2] * (A.fc[ENE][2] + B.fc[ENE][2]);
    for (i = 0; i < 3; i++)
        fb[MX + i] += factor_b * (B.m[i] - A.m[i])
                    - f[0] * (A.fc[MX + i][0] + B.fc[MX + i][0]) - f[1] * (A.fc[MX + i][1] + B.fc[MX + i][1])
                    - f[2] * (A.fc[MX + i][2] + B.fc[MX + i][2]);
}

/* ---- flux.h:14-39 with flux_boundary.elem_func / flux_wall.elem_func ---- */
static void orc_k_compute_bnd_node_flux_kernel(const int *g, const double *w, const double *u, double *fl)
{
    node_state S;
    int i;
    if (*g <= 2) {
        /* flux_boundary.elem_func:9-54 -- despite the file name: pressure-only wall */
        derive_state(u, &S, 0);
        fl[RHO] += 0;
        for (i = 0; i < 3; i++) fl[MX + i] += w[i] * S.p;
        fl[ENE] += 0;
    } else if (*g == 3 || (*g >= 4 && *g <= 7)) {
        /* flux_wall.elem_func:8-76 -- despite the file name: far field */
        double f[3];
        derive_state(u, &S, 1);
        for (i = 0; i < 3; i++) f[i] = 0.5 * w[i];
        fl[RHO] += f[0] * (k_ff_var[MX + 0] + S.m[0]) + f[1] * (k_ff_var[MX + 1] + S.m[1])
                 + f[2] * (k_ff_var[MX + 2] + S.m[2]);
        fl[ENE] += f[0] * (k_ff_fc[ENE][0] + S.fc[ENE][0]) + f[1] * (k_ff_fc[ENE][1] + S.fc[ENE][1])
                 + f[2] * (k_ff_fc[ENE][2] + S.fc[ENE][2]);
        for (i = 0; i < 3; i++)
            fl[MX + i] += f[0] * (k_ff_fc[MX + i][0] + S.fc[MX + i][0]) + f[1] * (k_ff_fc[MX + i][1] + S.fc[MX + i][1])
                        + f[2] * (k_ff_fc[MX + i][2] + S.fc[MX + i][2]);
    }
}

/* ---- unstructured_stream.h:7-57 ---- */
static void orc_k_unstructured_stream_kernel(const double *ua, const double *ub, const double *w,
                                             double *fa, double *fb)
{
    fa[RHO] += ub[RHO] + w[0];
    fa[MX] += ub[MX] + w[2];
    fa[MY] += ub[MY];
    fa[MZ] += ub[MZ];
    fa[ENE] += ub[ENE] + w[1];
    fb[RHO] += ua[RHO];
    fb[MX] += ua[MX];
    fb[MY] += ua[MY];
    fb[MZ] += ua[MZ];
    fb[ENE] += ua[ENE];
}

/* ---- validation.h ---- */
static void orc_k_residual_kernel(const double *old, const double *u, double *res)      /* :27-35 */
{
    int v;
    for (v = 0; v < 5; v++) res[v] = u[v] - old[v];
}
static void orc_k_calc_rms_kernel(const double *res, double *rms)                       /* :37-44 */
{
    int i;
    for (i = 0; i < 5; i++) *rms += res[i] * res[i];
}
static void orc_k_identify_differences(const double *test, const double *master, double *diff) /* :46-89 */
{
    int v;
    for (v = 0; v < 5; v++) {
        double tol = master[v] * 10.0e-8, d = test[v] - master[v];
        if (tol < 0.0) tol *= -1.0;
        if (tol < 3.0e-19) tol = 3.0e-19;
        if (d < 0.0) d *= -1.0;
        diff[v] = (d > tol) ? d : 0.0;
    }
}
static void orc_k_count_non_zeros(const double *value, int *count)                      /* :91-100 */
{
    int v;
    for (v = 0; v < 5; v++) if (value[v] > 0.0) (*count)++;
}
static void orc_k_count_bad_vals(const double *value, int *count)                       /* :102-115 */
{
    int v;
    for (v = 0; v < 5; v++) if (isnan(value[v]) || isinf(value[v])) *count += 1;
}

/* ---- mg.h ---- */
static void orc_k_up_pre_kernel(double *u, int *scratch)                                /* :29-39 */
{
    int i;
    for (i = 0; i < 5; i++) u[i] = 0.0;
    *scratch = 0;
}
static void orc_k_up_kernel(const double *u, double *above, int *scratch)               /* :41-52 */
{
    int i;
    for (i = 0; i < 5; i++) above[i] += u[i];
    *scratch += 1;
}
static void orc_k_up_post_kernel(double *u, const int *scratch)                         /* :54-64 */
{
    double avg = (*scratch) == 0 ? 1.0 : 1.0 / (double)(*scratch);
    int i;
    for (i = 0; i < 5; i++) u[i] *= avg;
}
static void orc_k_down_kernel(double *u, const double *res, const double *xyz, const double *res_above,
                              const double *xyz_above)                                  /* :66-88 */
{
    double d[3], dm;
    int i;
    for (i = 0; i < 3; i++) d[i] = fabs(xyz[i] - xyz_above[i]);
    dm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    u[RHO] -= dm * (res_above[RHO] - res[RHO]);
    for (i = 0; i < 3; i++) u[MX + i] -= d[i] * (res_above[MX + i] - res[MX + i]);   /* Q10: per-axis distance */
    u[ENE] -= dm * (res_above[ENE] - res[ENE]);
}

#define EK(name) orc_k_##name
#define ORC_LIB_NAME "port"
#define G_FF_VARIABLE k_ff_var
#include "op2_seq.inc"
