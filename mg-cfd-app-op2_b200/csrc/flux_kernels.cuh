// flux_kernels.cuh -- the three compute_flux_edge_kernel implementations, the boundary flux and
// (as a body switch) unstructured_stream_kernel.  Included by two translation units:
//   kernels.cu    with MGCFD_EXACT defined, compiled -fmad=false: reference operation order
//                 (flux.h:41-208), IEEE div/sqrt, no contraction -> per-edge increments are
//                 bit-identical to the CPU reference;
//   flux_fast.cu  default flags: per-node derived quantities (p, |v|+c, 1/rho) computed once
//                 per staged node, one antisymmetric 5-vector per edge, FMA contraction.
#pragma once
#include <cstdlib>
#include <map>
#include <tuple>
#include <type_traits>

#include <cuda_runtime.h>

#include "internal.h"

namespace mgcfd {

#ifdef MGCFD_EXACT
#define FLUX_NS exact
#else
#define FLUX_NS fast
#endif

namespace FLUX_NS {

constexpr double GAMMA = 1.4;   // const.h:27

// ------------------------------------------------------------------------------------------
// per-node state
// ------------------------------------------------------------------------------------------
#ifdef MGCFD_EXACT
constexpr int NF = 5;           // staged fields per node: the conserved variables only
struct Side {                   // everything flux.h:52-91 / :101-136 derives for one end
    double rho, m[3], E, v[3], q2, speed, p, c;
};
__device__ __forceinline__ Side derive(const double u[5])
{
    Side s;
    s.rho = u[0]; s.m[0] = u[1]; s.m[1] = u[2]; s.m[2] = u[3]; s.E = u[4];
    s.v[0] = s.m[0] / s.rho; s.v[1] = s.m[1] / s.rho; s.v[2] = s.m[2] / s.rho;     // inlined_funcs.h:120-125
    s.q2 = s.v[0] * s.v[0] + s.v[1] * s.v[1] + s.v[2] * s.v[2];                     // :98-101
    s.speed = sqrt(s.q2);
    s.p = (GAMMA - 1.0) * (s.E - 0.5 * s.rho * s.q2);                               // :103-106
    s.c = sqrt(GAMMA * s.p / s.rho);                                                // :126-129
    return s;
}
// fc[i][d]: flux contribution of momentum component i (0..2) / energy (3) in direction d; inlined_funcs.h:70-96
__device__ __forceinline__ void contributions(const Side &s, double fc[4][3])
{
    fc[0][0] = s.v[0] * s.m[0] + s.p; fc[0][1] = s.v[0] * s.m[1];       fc[0][2] = s.v[0] * s.m[2];
    fc[1][0] = fc[0][1];              fc[1][1] = s.v[1] * s.m[1] + s.p; fc[1][2] = s.v[1] * s.m[2];
    fc[2][0] = fc[0][2];              fc[2][1] = fc[1][2];              fc[2][2] = s.v[2] * s.m[2] + s.p;
    double ep = s.E + s.p;
    fc[3][0] = s.v[0] * ep; fc[3][1] = s.v[1] * ep; fc[3][2] = s.v[2] * ep;
}
// both increments in the reference's association order, flux.h:139-207.  g = ((-|w|)*smoothing)*0.5 (host)
__device__ __forceinline__ void edge_flux(const double ua[5], const double ub[5], double w0, double w1, double w2,
                                          double g, double fa[5], double fb[5])
{
    Side B = derive(ub), A = derive(ua);
    double fcA[4][3], fcB[4][3];
    contributions(B, fcB);
    contributions(A, fcA);
    double factor_a = g * (A.speed + B.speed + A.c + B.c);
    double factor_b = g * (B.speed + A.speed + B.c + A.c);
    double fx = -0.5 * w0, fy = -0.5 * w1, fz = -0.5 * w2;
    fa[0] = factor_a * (A.rho - B.rho) + fx * (A.m[0] + B.m[0]) + fy * (A.m[1] + B.m[1]) + fz * (A.m[2] + B.m[2]);
    fa[4] = factor_a * (A.E - B.E) + fx * (fcA[3][0] + fcB[3][0]) + fy * (fcA[3][1] + fcB[3][1]) + fz * (fcA[3][2] + fcB[3][2]);
    fb[0] = factor_b * (B.rho - A.rho) - fx * (A.m[0] + B.m[0]) - fy * (A.m[1] + B.m[1]) - fz * (A.m[2] + B.m[2]);
    fb[4] = factor_b * (B.E - A.E) - fx * (fcA[3][0] + fcB[3][0]) - fy * (fcA[3][1] + fcB[3][1]) - fz * (fcA[3][2] + fcB[3][2]);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        fa[1 + i] = factor_a * (A.m[i] - B.m[i]) + fx * (fcA[i][0] + fcB[i][0]) + fy * (fcA[i][1] + fcB[i][1]) + fz * (fcA[i][2] + fcB[i][2]);
        fb[1 + i] = factor_b * (B.m[i] - A.m[i]) - fx * (fcA[i][0] + fcB[i][0]) - fy * (fcA[i][1] + fcB[i][1]) - fz * (fcA[i][2] + fcB[i][2]);
    }
}
constexpr int NFLUX = 10;       // per-edge values kept for the gather: fa[5], fb[5]
#else
constexpr int NF = 8;           // staged fields per node: rho, mx, my, mz, E, p, s = |v|+c, 1/rho
__device__ __forceinline__ void derive(const double u[5], double r[8])
{
    double rinv = 1.0 / u[0];
    double vx = u[1] * rinv, vy = u[2] * rinv, vz = u[3] * rinv;
    double q2 = vx * vx + vy * vy + vz * vz;
    double p = (GAMMA - 1.0) * (u[4] - 0.5 * u[0] * q2);
    r[0] = u[0]; r[1] = u[1]; r[2] = u[2]; r[3] = u[3]; r[4] = u[4];
    r[5] = p;
    r[6] = sqrt(q2) + sqrt(GAMMA * p * rinv);
    r[7] = rinv;
}
// one antisymmetric increment F: flux_a += F, flux_b -= F.  Same mathematics as flux.h:139-207 with
// sum_d w_d*fc_i[d] rewritten as v_i*(w.m) + p*w_i and fc_E[d] as v_d*(E+p).
__device__ __forceinline__ void edge_flux(const double a[8], const double b[8], double w0, double w1, double w2,
                                          double g, double F[5])
{
    double factor = g * (a[6] + b[6]);
    double wma = w0 * a[1] + w1 * a[2] + w2 * a[3];
    double wmb = w0 * b[1] + w1 * b[2] + w2 * b[3];
    double wva = wma * a[7], wvb = wmb * b[7];       // w . v
    double psum = a[5] + b[5];
    F[0] = factor * (a[0] - b[0]) - 0.5 * (wma + wmb);
    F[1] = factor * (a[1] - b[1]) - 0.5 * (wva * a[1] + wvb * b[1] + w0 * psum);
    F[2] = factor * (a[2] - b[2]) - 0.5 * (wva * a[2] + wvb * b[2] + w1 * psum);
    F[3] = factor * (a[3] - b[3]) - 0.5 * (wva * a[3] + wvb * b[3] + w2 * psum);
    F[4] = factor * (a[4] - b[4]) - 0.5 * (wva * (a[4] + a[5]) + wvb * (b[4] + b[5]));
}
constexpr int NFLUX = 5;
#endif

// unstructured_stream.h:7-57 on raw variables
__device__ __forceinline__ void stream_flux(const double ua[5], const double ub[5], double w0, double w1, double w2,
                                            double fa[5], double fb[5])
{
    fa[0] = ub[0] + w0; fa[1] = ub[1] + w2; fa[2] = ub[2]; fa[3] = ub[3]; fa[4] = ub[4] + w1;
    fb[0] = ua[0]; fb[1] = ua[1]; fb[2] = ua[2]; fb[3] = ua[3]; fb[4] = ua[4];
}

__device__ __forceinline__ void load5(const double *__restrict__ p, double u[5])
{
#pragma unroll
    for (int v = 0; v < 5; v++) u[v] = __ldg(p + v);
}

// ------------------------------------------------------------------------------------------
// compute_bnd_node_flux_kernel body (flux.h:14-39, flux_boundary.elem_func, flux_wall.elem_func) for ONE node:
// applies the node's boundary entries [j0, j1) (ascending file order) to fl[5].  Shared by the stand-alone
// boundary kernel and the fused Runge-Kutta stage.
// ------------------------------------------------------------------------------------------
template <typename GroupT>
__device__ __forceinline__ void bnd_apply(const double u[5], double fl[5], int j0, int j1,
                                          const GroupT *__restrict__ b_group, const double *__restrict__ b_wt,
                                          const DevConsts &c)
{
    double rho = u[0], m[3] = {u[1], u[2], u[3]}, E = u[4];
    double v[3] = {m[0] / rho, m[1] / rho, m[2] / rho};
    double q2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double p = (GAMMA - 1.0) * (E - 0.5 * rho * q2);
    double fc[4][3];
    fc[0][0] = v[0] * m[0] + p; fc[0][1] = v[0] * m[1];     fc[0][2] = v[0] * m[2];
    fc[1][0] = fc[0][1];        fc[1][1] = v[1] * m[1] + p; fc[1][2] = v[1] * m[2];
    fc[2][0] = fc[0][2];        fc[2][1] = fc[1][2];        fc[2][2] = v[2] * m[2] + p;
    double ep = E + p;
    fc[3][0] = v[0] * ep; fc[3][1] = v[1] * ep; fc[3][2] = v[2] * ep;
    for (int i = j0; i < j1; i++) {
        int g = b_group[i];
        double w[3] = {b_wt[(size_t)i * 3], b_wt[(size_t)i * 3 + 1], b_wt[(size_t)i * 3 + 2]};
        if (g <= 2) {
            // pressure-only wall (flux_boundary.elem_func:50-54); "+= 0" only matters for -0.0
            fl[0] += 0.0;
            fl[1] += w[0] * p;
            fl[2] += w[1] * p;
            fl[3] += w[2] * p;
            fl[4] += 0.0;
        } else if (g == 3 || (g >= 4 && g <= 7)) {
            // far field (flux_wall.elem_func:49-76)
            double fx = 0.5 * w[0], fy = 0.5 * w[1], fz = 0.5 * w[2];
            fl[0] += fx * (c.ff_variable[1] + m[0]) + fy * (c.ff_variable[2] + m[1]) + fz * (c.ff_variable[3] + m[2]);
            fl[4] += fx * (c.ff_fc[4][0] + fc[3][0]) + fy * (c.ff_fc[4][1] + fc[3][1]) + fz * (c.ff_fc[4][2] + fc[3][2]);
#pragma unroll
            for (int k = 0; k < 3; k++)
                fl[1 + k] += fx * (c.ff_fc[1 + k][0] + fc[k][0]) + fy * (c.ff_fc[1 + k][1] + fc[k][1]) + fz * (c.ff_fc[1 + k][2] + fc[k][2]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// variant 0: thread per edge, fp64 RED with warp aggregation on the sorted (lower) endpoint
// ------------------------------------------------------------------------------------------
template <bool STREAM>
__global__ void __launch_bounds__(256)
flux_atomic_kernel(int E, int n_owned, const int2 *__restrict__ nodes, const double4 *__restrict__ wts,
                   const double *__restrict__ var, double *__restrict__ flux)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int lo = -1 - lane, hi = -1;
    double flo[5] = {0, 0, 0, 0, 0}, fhi[5] = {0, 0, 0, 0, 0};
    if (e < E) {
        int2 ab = nodes[e];
        double4 w = wts[e];
        double ua[5], ub[5], fa[5], fb[5];
        load5(var + (size_t)ab.x * 5, ua);
        load5(var + (size_t)ab.y * 5, ub);
        if (STREAM) {
            stream_flux(ua, ub, w.x, w.y, w.z, fa, fb);
        } else {
#ifdef MGCFD_EXACT
            edge_flux(ua, ub, w.x, w.y, w.z, w.w, fa, fb);
#else
            double ra[8], rb[8];
            derive(ua, ra);
            derive(ub, rb);
            edge_flux(ra, rb, w.x, w.y, w.z, w.w, fa);
#pragma unroll
            for (int v = 0; v < 5; v++) fb[v] = -fa[v];
#endif
        }
        bool a_low = ab.x <= ab.y;
        lo = a_low ? ab.x : ab.y;
        hi = a_low ? ab.y : ab.x;
#pragma unroll
        for (int v = 0; v < 5; v++) {
            flo[v] = a_low ? fa[v] : fb[v];
            fhi[v] = a_low ? fb[v] : fa[v];
        }
    }
    // segmented inclusive scan over runs of equal `lo` (runs are contiguous: edges are sorted by lo)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int key = __shfl_up_sync(0xffffffffu, lo, d);
        bool take = lane >= d && key == lo;
#pragma unroll
        for (int v = 0; v < 5; v++) {
            double t = __shfl_up_sync(0xffffffffu, flo[v], d);
            if (take) flo[v] += t;
        }
    }
    int next = __shfl_down_sync(0xffffffffu, lo, 1);
    bool tail = lane == 31 || next != lo;
    if (e < E) {
        if (tail && lo < n_owned) {
#pragma unroll
            for (int v = 0; v < 5; v++) atomicAdd(flux + (size_t)lo * 5 + v, flo[v]);
        }
        if (hi < n_owned) {
#pragma unroll
            for (int v = 0; v < 5; v++) atomicAdd(flux + (size_t)hi * 5 + v, fhi[v]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// variant 1: OP2-style hierarchical colouring.  One launch per block colour; a CTA stages the
// block's nodes in shared memory, computes one edge per thread, accumulates into shared memory
// thread colour by thread colour, then adds the block's sums to HBM without atomics.
// shared: rec[NF][max_nodes] | acc[5][max_nodes]
// ------------------------------------------------------------------------------------------
template <bool STREAM>
__global__ void __launch_bounds__(256)
flux_colour_kernel(int slot0, int max_nodes, int n_owned, const int *__restrict__ blk_edge0,
                   const int *__restrict__ blk_node0, const int *__restrict__ blk_ncol,
                   const int *__restrict__ node_gid, const uint32_t *__restrict__ lab,
                   const unsigned char *__restrict__ ecol, const double4 *__restrict__ wts,
                   const double *__restrict__ var, double *__restrict__ flux)
{
    extern __shared__ double sm[];
    constexpr int NREC = STREAM ? 5 : NF;
    double *rec = sm;
    double *acc = sm + (size_t)NREC * max_nodes;
    const int s = slot0 + blockIdx.x, tid = threadIdx.x;
    const int e0 = blk_edge0[s], ne = blk_edge0[s + 1] - e0;
    const int n0 = blk_node0[s], nn = blk_node0[s + 1] - n0;

    for (int i = tid; i < nn; i += blockDim.x) {
        double u[5];
        load5(var + (size_t)node_gid[n0 + i] * 5, u);
#ifdef MGCFD_EXACT
#pragma unroll
        for (int f = 0; f < 5; f++) rec[f * max_nodes + i] = u[f];
#else
        if (STREAM) {
#pragma unroll
            for (int f = 0; f < 5; f++) rec[f * max_nodes + i] = u[f];
        } else {
            double r[8];
            derive(u, r);
#pragma unroll
            for (int f = 0; f < 8; f++) rec[f * max_nodes + i] = r[f];
        }
#endif
#pragma unroll
        for (int v = 0; v < 5; v++) acc[v * max_nodes + i] = 0.0;
    }
    __syncthreads();

    double fa[5], fb[5];
    int la = 0, lb = 0, col = -1;
    if (tid < ne) {
        uint32_t l = lab[e0 + tid];
        la = l & 0xffff;
        lb = l >> 16;
        col = ecol[e0 + tid];
        double4 w = wts[e0 + tid];
        double a[NREC], b[NREC];
#pragma unroll
        for (int f = 0; f < NREC; f++) { a[f] = rec[f * max_nodes + la]; b[f] = rec[f * max_nodes + lb]; }
        if (STREAM) {
            stream_flux(a, b, w.x, w.y, w.z, fa, fb);
        } else {
#ifdef MGCFD_EXACT
            edge_flux(a, b, w.x, w.y, w.z, w.w, fa, fb);
#else
            edge_flux(a, b, w.x, w.y, w.z, w.w, fa);
#pragma unroll
            for (int v = 0; v < 5; v++) fb[v] = -fa[v];
#endif
        }
    }
    const int ncol = blk_ncol[s];
    for (int c = 0; c < ncol; c++) {
        if (col == c) {
#pragma unroll
            for (int v = 0; v < 5; v++) {
                acc[v * max_nodes + la] += fa[v];
                acc[v * max_nodes + lb] += fb[v];
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < nn; i += blockDim.x) {
        int gid = node_gid[n0 + i];
        if (gid < n_owned) {
#pragma unroll
            for (int v = 0; v < 5; v++) flux[(size_t)gid * 5 + v] += acc[v * max_nodes + i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// variant 2: owner-compute chunks.  A CTA owns a run of consecutive nodes.
//   1. one thread arms an mbarrier and issues ONE bulk async copy (cp.async.bulk, the TMA 1-D path)
//      of the chunk's contiguous blob -- edge weights, block-local endpoints, incidence CSR -- into
//      shared memory; meanwhile all threads stage owned + halo node states (SoA, derived quantities
//      once per node);
//   2. one thread per edge (cut edges are recomputed by the neighbouring chunk) turns its weights
//      in place into the edge's flux vector;
//   3. one thread per owned node sums its incident edges in ascending file order and stores the
//      result: no atomics, no colours, deterministic, and in the exact build bit-identical to OP2-seq.
// shared: mbarrier | blob (w0 w1 w2 g [e_pad] doubles, lab [e_pad] u32, rowptr, csr u16) |
//         extra flux planes [NFL-4][max_edges] | raw[max_loc][5] (AoS state tile) | der[3][max_loc] (fast build)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

__device__ __forceinline__ double flip_sign(double x, int sign_mask)
{
    return __hiloint2double(__double2hiint(x) ^ sign_mask, __double2loint(x));
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// order-preserving map double -> uint64 so that atomicMin implements OP_MIN on doubles
__device__ __forceinline__ unsigned long long enc_min(double d)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_min(unsigned long long u)
{
    unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)b);
}

// ---- multi-GPU (p2p transport): system-scope flags and the fused halo push of the stage kernels (StagePush, internal.h)
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// wait until *flag >= e, at most timeout_ns of wall clock (%globaltimer): a peer that never arrives raises *err
// (MGCFD_ERR_COMM) instead of hanging the GPU for ever.  MGCFD_COMM_TIMEOUT_MS sets the bound (default 60 s).
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// ACQUIRE = true: ld.acquire.sys -- on this hardware a strong load followed by an invalidation of the SM's whole L1.  The
// stage kernels need it: their halo gathers allocate in L1, and the line that holds the first rank-halo rows also holds
// the last owned rows, which a chunk without exports may have pulled in before the neighbours' rows landed.
// ACQUIRE = false: a strong (relaxed, system-scope) load only -- for kernels in which EVERY block waits before it reads
// anything a peer writes (node kernels, min_dt mailboxes): L1 starts empty at kernel launch, so nothing stale can be
// cached, and the 4-16 K blocks of such a grid do not flush each other's L1 4-16 K times.
template <bool ACQUIRE = true>
__device__ __forceinline__ void bounded_wait(const unsigned long long *flag, unsigned long long e, int *err, long long timeout_ns)
{
    if ((ACQUIRE ? ld_acquire_sys(flag) : ld_relaxed_sys(flag)) >= e) return;
    const unsigned long long t0 = global_ns();
    for (unsigned it = 0;; it++) {
        if ((ACQUIRE ? ld_acquire_sys(flag) : ld_relaxed_sys(flag)) >= e) return;
        __nanosleep(it < 64 ? 32 : 512);
        if ((it & 1023u) == 1023u && (long long)(global_ns() - t0) > timeout_ns) break;
    }
    atomicExch(err, 1);
}
// consumer side of a chunk that owns exported nodes: all sources' rows of the previous exchange have landed (and the
// neighbours are done reading what this stage overwrites).  Call with all threads of the CTA.
__device__ __forceinline__ void push_wait_sources(const StagePush &P, int tid)
{
    if (tid < P.n_src) bounded_wait(P.src_flag[tid], *P.expected[tid], P.err_flag, P.timeout_ns);
    __syncthreads();
}
// producer side, per owned node: store component v of var_new (and of the residual) into every destination that holds
// the node in its halo; [j0, j1) = the node's entries
__device__ __forceinline__ void push_component(const StagePush &P, int j0, int j1, int v, double vn, double r, bool with_res)
{
    for (int j = j0; j < j1; j++) {
        const int2 t = __ldg(P.xp_ent + j);
        P.var_dst[t.x][(size_t)t.y * 5 + v] = vn;
        if (with_res && P.res_dst[t.x]) P.res_dst[t.x][(size_t)t.y * 5 + v] = r;
    }
}
// after the chunk's pushes: count the chunk; the last one of the launch publishes the epoch and arms the next consumer.
// Every chunk orders its threads' peer stores before its count with a CTA barrier and ONE device-scope fence; the last
// chunk -- which has observed all counts -- issues the only system-scope fence of the launch before the flags go out
// (causality is cumulative across the two scopes; a system-scope fence in every chunk measured 15-20 us per launch).
__device__ __forceinline__ void push_publish(const StagePush &P, int tid)
{
    __syncthreads();
    if (tid < 32) {
        int last = 0;
        if (tid == 0) {
            __threadfence();
            const unsigned int prev = atomicAdd(P.done, 1u);
            last = prev + 1u == (unsigned int)P.n_boundary;
            if (last) { atomicExch(P.done, 0u); __threadfence_system(); }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            // one lane per destination: the release stores travel in parallel, not one NVLink round trip after the other
            if (tid < P.n_dst) {
                const unsigned long long e = *P.sent[tid] + 1;
                *P.sent[tid] = e;
                __threadfence_system();
                st_release_sys(P.dst_flag[tid], e);
            }
            if (tid < P.n_src) *P.expected[tid] += 1;
        }
    }
}

// state of local node i for the edge body: conserved variables from the AoS tile, derived quantities from 3 planes
template <int NREC>
__device__ __forceinline__ void load_state(const double *raw, const double *der, int stride, int i, double r[NREC])
{
#pragma unroll
    for (int v = 0; v < 5; v++) r[v] = raw[i * 5 + v];
    if (NREC == 8) {
        r[5] = der[i]; r[6] = der[stride + i]; r[7] = der[2 * stride + i];
    }
}

// Stage the conserved variables of a chunk's local nodes into the AoS tile `raw`:
//   owned nodes: one bulk async copy (they are a contiguous run starting on an even node index, hence 16-byte
//                aligned; a trailing odd node's last 8 bytes are copied by hand), completion on `bar`;
//   halo nodes:  8-byte cp.async requests, four per thread in flight at a time.
// Call with all threads; returns after this thread's cp.async requests have completed.  The caller must
// __syncthreads() and mbar_wait(bar, 0) before reading the tile.
__device__ __forceinline__ uint32_t owned_bulk_bytes(int n_own) { return ((uint32_t)n_own * 40u) & ~15u; }
__device__ __forceinline__ void stage_tile(double *raw, uint64_t *bar, int node0, int n_own,
                                           int n_halo, const int *__restrict__ hg, const double *__restrict__ var,
                                           int tid, int nthreads)
{
    const uint32_t own_bytes = (uint32_t)n_own * 40u, bulk_bytes = own_bytes & ~15u;
    if (tid == 0 && bulk_bytes) bulk_g2s(raw, var + (size_t)node0 * 5, bulk_bytes, bar);   // expect_tx armed by the caller
    if (tid == 32 && bulk_bytes != own_bytes) raw[n_own * 5 - 1] = __ldg(var + (size_t)(node0 + n_own) * 5 - 1);
    const int nh5 = n_halo * 5;
    double *hraw = raw + (size_t)n_own * 5;
    for (int base = 0; base < nh5; base += 4 * nthreads) {
        int gid[4], f[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            f[k] = base + k * nthreads + tid;
            gid[k] = f[k] < nh5 ? __ldg(hg + f[k] / 5) : -1;
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (gid[k] >= 0) cp_async8(hraw + f[k], var + (size_t)gid[k] * 5 + (f[k] % 5));
    }
    cp_async_wait_all();
}

// Last step of a chunk: the tile `raw` holds the owned nodes' flux sums (AoS).  Plain variant: one contiguous,
// coalesced store to `flux`.  FUSE: time_stepping_kernels.h:66-86 applied on the spot (+ validation.h:27-44,102-115
// after the last stage) with the prefetched old_variables / step_factor tiles.
template <bool FUSE>
__device__ __forceinline__ void finish_chunk(const double *raw, int node0, int n_own, double *__restrict__ flux,
                                             const RkStageArgs &rk, const double *told, const double *tsf,
                                             uint32_t old_bulk, uint32_t sf_bulk)
{
    const int tid = threadIdx.x;
    if (!FUSE) {
        double *out = flux + (size_t)node0 * 5;
        for (int f = tid; f < n_own * 5; f += blockDim.x) out[f] = raw[f];
    } else {
        const size_t g0 = (size_t)node0 * 5;
        const double denom = (double)(MGCFD_RK + 1 - rk.rk);
        double sq = 0.0;
        int bad = 0;
        const int old_n = (int)(old_bulk >> 3), sf_n = (int)(sf_bulk >> 3);    // an odd last chunk has one tail element
        for (int f = tid; f < n_own * 5; f += blockDim.x) {
            int node = f / 5;
            double factor = (node < sf_n ? tsf[node] : rk.sf[node0 + node]) / denom;
            double o = f < old_n ? told[f] : rk.old[g0 + f];
            double vn = __dadd_rn(o, __dmul_rn(factor, raw[f]));
            rk.var_out[g0 + f] = vn;
            if (rk.last) {
                double r = vn - o;
                rk.res[g0 + f] = r;
                sq += r * r;
                bad += (isnan(vn) || isinf(vn)) ? 1 : 0;
            }
        }
        if (rk.last && rk.d_rms) {
            for (int o = 16; o > 0; o >>= 1) {
                sq += __shfl_xor_sync(0xffffffffu, sq, o);
                bad += __shfl_xor_sync(0xffffffffu, bad, o);
            }
            if ((tid & 31) == 0) {
                atomicAdd(rk.d_rms, sq);
                if (bad) atomicAdd(rk.d_bad, bad);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Node phase shared by the owner kernels (after the edge phase has parked every edge's flux vector in shared
// memory): owned nodes sum their incident edges in ascending file order -- the fast build splits a node's incidences
// between 2 or 4 threads when the CTA has them --, add the chunk's boundary entries (they travel in the blob, sorted by
// owned node), and finish from registers: plain store of the flux sums, or FUSE: time_stepping_kernels.h:66-86
// (+ validation.h:27-44,102-115 after the last stage) with the prefetched old_variables / step_factor tiles, bit for
// bit the arithmetic of time_step_kernel.  No shared-memory staging of the result and no barrier after it.
// ------------------------------------------------------------------------------------------
template <bool OVERWRITE, bool FUSE, int LG_SPLIT, bool LEAN = false>
__device__ __forceinline__ void node_phase(int tid, const OwnerChunkDesc &d, const unsigned char *sblob,
                                           const double *raw, const double *w0, const double *w1, const double *w2,
                                           const double *gg, const double *Fx, const uint16_t *rowptr, const uint16_t *csr,
                                           int max_edges, const double *told, const double *tsf, double *__restrict__ flux,
                                           const RkStageArgs &rk, int xb = -1)
{
    constexpr int lg_split = LG_SPLIT, step = 1 << LG_SPLIT;      // threads per owned node (fast build: 1, 2 or 4)
    const int n = tid >> lg_split, part = tid & (step - 1);
    const bool active = n < d.n_own;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (active) {
        if (!OVERWRITE && part == 0) {
#pragma unroll
            for (int v = 0; v < 5; v++) acc[v] = flux[(size_t)(d.node0 + n) * 5 + v];
        }
        // the flux vector an incidence contributes (end b of an edge receives the negated / the fb vector)
        auto fetch = [&](uint16_t c, double f[5]) {
            const int e = c & 0x7fff;
#ifdef MGCFD_EXACT
            if (c & 0x8000) {
#pragma unroll
                for (int v = 0; v < 5; v++) f[v] = Fx[(1 + v) * max_edges + e];
            } else {
                f[0] = w0[e]; f[1] = w1[e]; f[2] = w2[e]; f[3] = gg[e]; f[4] = Fx[e];
            }
#else
            const int sgn = (c & 0x8000) ? (int)0x80000000 : 0;
            f[0] = flip_sign(w0[e], sgn); f[1] = flip_sign(w1[e], sgn); f[2] = flip_sign(w2[e], sgn);
            f[3] = flip_sign(gg[e], sgn); f[4] = flip_sign(Fx[e], sgn);
#endif
        };
        if (!LEAN) {
            const int j1 = rowptr[n + 1];
            for (int jj = rowptr[n] + part; jj < j1; jj += step) {
                double f[5];
                fetch(csr[jj], f);
#pragma unroll
                for (int v = 0; v < 5; v++) acc[v] += f[v];
            }
        }
    }
#ifndef MGCFD_EXACT
    if (LEAN) {
        // lean sum loop (fast build): a warp-uniform trip count (no divergent-loop bookkeeping), the sign of an
        // incidence applied by fma(+-1, F, acc) (exactly acc +- F), plane addresses by adds from one shifted index
        const int j0 = active ? rowptr[n] + part : 0, j1 = active ? rowptr[n + 1] : 0;
        const int mine = j1 > j0 ? (j1 - j0 + step - 1) >> lg_split : 0;
        const int trips = __reduce_max_sync(0xffffffffu, mine);
        const uint32_t pad8 = (uint32_t)(w1 - w0) * 8u;
        const char *b0 = reinterpret_cast<const char *>(w0), *bx = reinterpret_cast<const char *>(Fx);
        for (int it = 0, jj = j0; it < trips; it++, jj += step) {
            const bool valid = jj < j1;
            const uint32_t c = csr[valid ? jj : 0];
            const uint32_t e8 = (c & 0x7fffu) << 3;
            const double sg = (c & 0x8000u) ? -1.0 : 1.0;
            const char *a0 = b0 + e8;
            const double f0 = *reinterpret_cast<const double *>(a0), f1 = *reinterpret_cast<const double *>(a0 + pad8),
                         f2 = *reinterpret_cast<const double *>(a0 + 2 * pad8), f3 = *reinterpret_cast<const double *>(a0 + 3 * pad8),
                         f4 = *reinterpret_cast<const double *>(bx + e8);
            if (valid) {
                acc[0] = fma(sg, f0, acc[0]); acc[1] = fma(sg, f1, acc[1]); acc[2] = fma(sg, f2, acc[2]);
                acc[3] = fma(sg, f3, acc[3]); acc[4] = fma(sg, f4, acc[4]);
            }
        }
    }
    if (lg_split >= 1) {
#pragma unroll
        for (int v = 0; v < 5; v++) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], 1);      // every thread of a node
    }                                                                                       // ends up with the sum
    if (lg_split == 2) {
#pragma unroll
        for (int v = 0; v < 5; v++) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], 2);
    }
#endif
    double sq = 0.0;
    int bad = 0;
    if (active) {
        if (FUSE && d.has_bnd) {
            // this node's range of the chunk's boundary entries
            const double *bw = reinterpret_cast<const double *>(sblob + d.bnd_off);
            const uint16_t *bptr = reinterpret_cast<const uint16_t *>(bw + (size_t)d.has_bnd * 3);
            const int16_t *bgrp = reinterpret_cast<const int16_t *>(bptr + (((d.n_own + 1) + 1) & ~1));
            const int b0 = bptr[n], b1 = bptr[n + 1];
            if (b1 > b0) {
                double u[5];
#pragma unroll
                for (int v = 0; v < 5; v++) u[v] = raw[n * 5 + v];
                bnd_apply(u, acc, b0, b1, bgrp, bw, rk.c);                 // a split node's threads all do the same
            }
        }
        const size_t g0 = (size_t)d.node0 * 5;
        // The node's thread `part` finishes components part, part+step, ...: at most NK of them, picked out of acc with
        // static indices (all threads of a warp then run the same NK update bodies).
        constexpr int NK = (5 + step - 1) / step;
        double mine[NK];
#pragma unroll
        for (int k = 0; k < NK; k++) {
            mine[k] = acc[k * step];
#pragma unroll
            for (int q = 1; q < step; q++)
                if (k * step + q < 5 && part == q) mine[k] = acc[k * step + q];
        }
        if (!FUSE) {
            double *out = flux + g0 + n * 5;
#pragma unroll
            for (int k = 0; k < NK; k++)
                if (part + k * step < 5) out[part + k * step] = mine[k];
        } else {
            // An odd owned run (the last chunk of a level only) has one old_variables element and one step factor
            // outside the bulk-copied tiles; the test is uniform, so that every other chunk runs without it.
            int xj0 = 0, xj1 = 0;
            if (xb >= 0) { xj0 = __ldg(rk.push.xp_ptr + xb + n); xj1 = __ldg(rk.push.xp_ptr + xb + n + 1); }
            auto finish = [&](auto tail_c) {
                constexpr bool TAIL = decltype(tail_c)::value;
                const int old_n = (int)(owned_bulk_bytes(d.n_own) >> 3), sf_n = (int)((((uint32_t)d.n_own * 8u) & ~15u) >> 3);
                const double factor = ((!TAIL || n < sf_n) ? tsf[n] : rk.sf[d.node0 + n]) / (double)(MGCFD_RK + 1 - rk.rk);
#pragma unroll
                for (int k = 0; k < NK; k++) {
                    const int v = part + k * step;
                    if (v < 5) {
                        const int f = n * 5 + v;
                        const double o = (!TAIL || f < old_n) ? told[f] : rk.old[g0 + f];
                        const double vn = __dadd_rn(o, __dmul_rn(factor, mine[k]));
                        rk.var_out[g0 + f] = vn;
                        const double r = vn - o;
                        if (rk.last) {
                            rk.res[g0 + f] = r;
                            sq += r * r;
                            bad += (isnan(vn) || isinf(vn)) ? 1 : 0;
                        }
                        if (xb >= 0) push_component(rk.push, xj0, xj1, v, vn, r, rk.last != 0);
                    }
                }
            };
            if (d.n_own & 1)
                finish(std::true_type{});
            else
                finish(std::false_type{});
        }
    }
    if (FUSE && rk.last && rk.d_rms) {
        for (int o = 16; o > 0; o >>= 1) {
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
            bad += __shfl_xor_sync(0xffffffffu, bad, o);
        }
        if ((tid & 31) == 0) {
            atomicAdd(rk.d_rms, sq);
            if (bad) atomicAdd(rk.d_bad, bad);
        }
    }
}

// FUSE: the Runge-Kutta stage in one kernel -- compute_flux_edge + compute_bnd_node_flux + time_step (+ residual,
// calc_rms, count_bad_vals after the last stage).  The owner of a node holds the node's complete edge-flux sum in
// registers, so it adds the boundary entries and applies var_new = old + step_factor/(RK+1-rk) * flux on the spot;
// the fluxes never travel to HBM (they are zero before and after every stage).  var_new goes to the
// alternate variables buffer because neighbouring chunks still read this stage's input.
#ifndef MGCFD_OWNER_MINB
#define MGCFD_OWNER_MINB 3      // resident CTAs per SM the register allocation aims for (4 forces spills and measured slower)
#endif
template <bool STREAM, bool OVERWRITE, bool FUSE, bool REGEPI = false>
__global__ void __launch_bounds__(256, MGCFD_OWNER_MINB)
flux_owner_kernel(int max_loc, int max_edges, int max_blob, const OwnerChunkDesc *__restrict__ descs,
                  const int *__restrict__ chunk_list, const int *__restrict__ halo_gid,
                  const unsigned char *__restrict__ blob, const double *__restrict__ var, double *__restrict__ flux,
                  const __grid_constant__ RkStageArgs rk)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int NREC = STREAM ? 5 : NF;
    constexpr int NFL = STREAM ? 10 : NFLUX;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smraw);
    unsigned char *sblob = smraw + 16;
    double *Fx = reinterpret_cast<double *>(sblob + max_blob);          // flux planes 4..NFL-1
    double *raw = Fx + (size_t)(NFL - 4) * max_edges;                   // conserved variables, AoS [max_loc][5]
    double *der = raw + (size_t)max_loc * 5;                            // p, |v|+c, 1/rho planes (fast build only)
    // FUSE: old_variables tile [max_own][5] and step factors [max_own], both 16-byte aligned bulk-copy targets
    double *told = raw + (((size_t)NREC * max_loc + 1) & ~(size_t)1);
    double *tsf = told + (((size_t)rk.max_own * 5 + 1) & ~(size_t)1);
    // chunk_list (multi-GPU): the launch covers a subset of the chunks, e.g. those that own exported nodes
    const OwnerChunkDesc d = descs[chunk_list ? chunk_list[blockIdx.x] : blockIdx.x];
    const int tid = threadIdx.x, nloc = d.n_own + d.n_halo;
    const uint32_t old_bulk = FUSE ? owned_bulk_bytes(d.n_own) : 0u, sf_bulk = FUSE ? (((uint32_t)d.n_own * 8u) & ~15u) : 0u;
    // multi-GPU, fused push: a chunk that owns exported nodes waits for its sources before it reads halo rows
    int xb = -1;
    if (FUSE && REGEPI && rk.push_on) {
        xb = __ldg(rk.push.xp_base + (chunk_list ? chunk_list[blockIdx.x] : blockIdx.x));
        if (xb >= 0) push_wait_sources(rk.push, tid);
    }

    // 1. bulk async copies (TMA 1-D) of the chunk's blob and of the owned nodes' conserved variables; 2. halo nodes
    //    by 8-byte async copies straight into the tile
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)d.blob_bytes + owned_bulk_bytes(d.n_own) + old_bulk + sf_bulk);
        bulk_g2s(sblob, blob + d.blob_off, (uint32_t)d.blob_bytes, bar);
        if (FUSE) {
            // the update's operands are fetched now and land while the fluxes are being computed
            if (old_bulk) bulk_g2s(told, rk.old + (size_t)d.node0 * 5, old_bulk, bar);
            if (sf_bulk) bulk_g2s(tsf, rk.sf + d.node0, sf_bulk, bar);
        }
    }
    stage_tile(raw, bar, d.node0, d.n_own, d.n_halo, halo_gid + d.halo_off, var, tid, blockDim.x);
    __syncthreads();          // halo tile complete; mbarrier initialisation visible to all threads
    mbar_wait(bar, 0);        // blob and owned tile landed
#ifndef MGCFD_EXACT
    // 3. derived quantities once per staged node
    if (!STREAM) {
        for (int i = tid; i < nloc; i += blockDim.x) {
            double u[5], r[8];
#pragma unroll
            for (int v = 0; v < 5; v++) u[v] = raw[i * 5 + v];
            derive(u, r);
            der[i] = r[5]; der[max_loc + i] = r[6]; der[2 * max_loc + i] = r[7];
        }
        __syncthreads();
    }
#endif

    double *w0 = reinterpret_cast<double *>(sblob);
    double *w1 = w0 + d.e_pad, *w2 = w1 + d.e_pad, *gg = w2 + d.e_pad;
    const uint32_t *lab = reinterpret_cast<const uint32_t *>(gg + d.e_pad);
    const uint16_t *rowptr = reinterpret_cast<const uint16_t *>(lab + d.e_pad);
    const uint16_t *csr = rowptr + (((d.n_own + 1) + 7) & ~7);

    // 4. one thread per edge: the edge's weights are replaced in place by its flux vector
    for (int e = tid; e < d.n_edges; e += blockDim.x) {
        uint32_t l = lab[e];
        int la = l & 0xffff, lb = l >> 16;
        double x = w0[e], y = w1[e], z = w2[e], g = gg[e];
        double a[NREC], b[NREC];
        load_state<NREC>(raw, der, max_loc, la, a);
        load_state<NREC>(raw, der, max_loc, lb, b);
        double fa[5], fb[5];
        if (STREAM) {
            stream_flux(a, b, x, y, z, fa, fb);
        } else {
#ifdef MGCFD_EXACT
            edge_flux(a, b, x, y, z, g, fa, fb);
#else
            edge_flux(a, b, x, y, z, g, fa);
#endif
        }
        w0[e] = fa[0]; w1[e] = fa[1]; w2[e] = fa[2]; gg[e] = fa[3];      // slot e is private to this thread
        Fx[e] = fa[4];
        if (NFL == 10) {
#pragma unroll
            for (int v = 0; v < 5; v++) Fx[(1 + v) * max_edges + e] = fb[v];
        }
    }
    __syncthreads();

    // 5. (REGEPI) node sums, boundary entries and the update straight from registers: see node_phase
    if constexpr (REGEPI) {
#ifdef MGCFD_EXACT
        node_phase<OVERWRITE, FUSE, 0>(tid, d, sblob, raw, w0, w1, w2, gg, Fx, rowptr, csr, max_edges, told, tsf, flux, rk, xb);
#else
        if (2 * rk.max_own <= (int)blockDim.x)
            node_phase<OVERWRITE, FUSE, 1>(tid, d, sblob, raw, w0, w1, w2, gg, Fx, rowptr, csr, max_edges, told, tsf, flux, rk, xb);
        else
            node_phase<OVERWRITE, FUSE, 0>(tid, d, sblob, raw, w0, w1, w2, gg, Fx, rowptr, csr, max_edges, told, tsf, flux, rk, xb);
#endif
        if (xb >= 0) push_publish(rk.push, tid);
        return;
    }
    // 5. one thread per owned node (chunks never own more than blockDim nodes) sums its incident edges in ascending
    //    file order; the sums go to HBM through shared memory so that the store is one contiguous, coalesced run
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int n = tid;
    if (n < d.n_own) {
        if (!OVERWRITE) {
#pragma unroll
            for (int v = 0; v < 5; v++) acc[v] = flux[(size_t)(d.node0 + n) * 5 + v];
        }
        int j0 = rowptr[n], j1 = rowptr[n + 1];
        for (int j = j0; j < j1; j++) {
            uint16_t c = csr[j];
            int e = c & 0x7fff;
            bool is_b = c & 0x8000;
            if (NFL == 10) {
                if (is_b) {
#pragma unroll
                    for (int v = 0; v < 5; v++) acc[v] += Fx[(1 + v) * max_edges + e];
                } else {
                    acc[0] += w0[e]; acc[1] += w1[e]; acc[2] += w2[e]; acc[3] += gg[e]; acc[4] += Fx[e];
                }
            } else {
                // end b receives the negated vector: flip the sign bit instead of selecting between f and -f
                const int sgn = is_b ? (int)0x80000000 : 0;
                acc[0] += flip_sign(w0[e], sgn);
                acc[1] += flip_sign(w1[e], sgn);
                acc[2] += flip_sign(w2[e], sgn);
                acc[3] += flip_sign(gg[e], sgn);
                acc[4] += flip_sign(Fx[e], sgn);
            }
        }
        if (FUSE && d.has_bnd) {
            int j0 = rk.bnd_ptr[d.node0 + n], j1 = rk.bnd_ptr[d.node0 + n + 1];
            if (j1 > j0) {
                double u[5];
#pragma unroll
                for (int v = 0; v < 5; v++) u[v] = raw[n * 5 + v];
                bnd_apply(u, acc, j0, j1, rk.b_group, rk.b_wt, rk.c);
            }
        }
    }
    if (n < d.n_own) {
#pragma unroll
        for (int v = 0; v < 5; v++) raw[n * 5 + v] = acc[v];      // the state tile is dead: reuse it for the output
    }
    __syncthreads();
    finish_chunk<FUSE>(raw, d.node0, d.n_own, flux, rk, told, tsf, old_bulk, sf_bulk);
}

// ------------------------------------------------------------------------------------------
// variant 2l ("lean", MGCFD_OWNER_LEAN=1; fast build, fused stage, chunks of up to 64 owned nodes, 128 threads): the
// one-CTA-per-chunk kernel with the two instruction-heavy phases of profiles/README.md section 5 rewritten --
//   staging: the chunk's descriptor and halo ids sit in ONE fixed-stride record (xtab[chunk][12 + hs]), so the ids are
//            requested together with the descriptor (two dependent L2 latencies instead of three); halo rows are copied by
//            groups of 8 lanes (lane = component 0..4 of the row: no divisions, one id load per row);
//   node phase: node_phase<..., LEAN> (warp-uniform trip count, fma sign, additive plane addresses).
// Everything else is flux_owner_kernel<false, true, true, true>.
// ------------------------------------------------------------------------------------------
#ifndef MGCFD_EXACT
constexpr int LEAN_ROWS = 8;            // halo rows per 8-lane group requested with the descriptor: 16 groups x 8 = 128
                                        // halo nodes; longer lists finish in a loop once the descriptor is known
__global__ void __launch_bounds__(128, 6)
flux_owner_lean_kernel(int max_loc, int max_edges, int max_blob, const int *__restrict__ xtab, int xs, int hs,
                       int rec0, const unsigned char *__restrict__ blob,
                       const double *__restrict__ var, RkStageArgs rk)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smraw);
    unsigned char *sblob = smraw + 16;
    double *Fx = reinterpret_cast<double *>(sblob + max_blob);
    double *raw = Fx + (size_t)(NFLUX - 4) * max_edges;
    double *der = raw + (size_t)max_loc * 5;
    double *told = raw + (((size_t)NF * max_loc + 1) & ~(size_t)1);
    double *tsf = told + (((size_t)rk.max_own * 5 + 1) & ~(size_t)1);
    const int tid = threadIdx.x, grp = tid >> 3, comp = tid & 7;
    const int chunk = rec0 + (int)blockIdx.x;                // the records are stored in launch order
    const int *rec = xtab + (size_t)chunk * xs;
    // halo ids of this thread's rows and the descriptor: independent loads, one round trip
    int hgv[LEAN_ROWS];
#pragma unroll
    for (int k = 0; k < LEAN_ROWS; k++) {
        const int row = grp + 16 * k;
        hgv[k] = row < hs ? __ldg(rec + 12 + row) : -1;
    }
    const OwnerChunkDesc d = *reinterpret_cast<const OwnerChunkDesc *>(rec);
    const int nloc = d.n_own + d.n_halo;
    const uint32_t old_bulk = owned_bulk_bytes(d.n_own), sf_bulk = ((uint32_t)d.n_own * 8u) & ~15u;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)d.blob_bytes + owned_bulk_bytes(d.n_own) + old_bulk + sf_bulk);
        bulk_g2s(sblob, blob + d.blob_off, (uint32_t)d.blob_bytes, bar);
        if (owned_bulk_bytes(d.n_own)) bulk_g2s(raw, var + (size_t)d.node0 * 5, owned_bulk_bytes(d.n_own), bar);
        if (old_bulk) bulk_g2s(told, rk.old + (size_t)d.node0 * 5, old_bulk, bar);
        if (sf_bulk) bulk_g2s(tsf, rk.sf + d.node0, sf_bulk, bar);
    }
    if (tid == 32 && (d.n_own & 1)) raw[d.n_own * 5 - 1] = __ldg(var + (size_t)(d.node0 + d.n_own) * 5 - 1);
    {
        double *hraw = raw + (size_t)d.n_own * 5;
#pragma unroll
        for (int k = 0; k < LEAN_ROWS; k++)
            if (hgv[k] >= 0 && comp < 5) cp_async8(hraw + (grp + 16 * k) * 5 + comp, var + (size_t)hgv[k] * 5 + comp);
        for (int row = grp + 16 * LEAN_ROWS; row < d.n_halo; row += 16)
            if (comp < 5) cp_async8(hraw + row * 5 + comp, var + (size_t)__ldg(rec + 12 + row) * 5 + comp);
    }
    cp_async_wait_all();
    __syncthreads();
    mbar_wait(bar, 0);
    for (int i = tid; i < nloc; i += 128) {
        double u[5], r[8];
#pragma unroll
        for (int v = 0; v < 5; v++) u[v] = raw[i * 5 + v];
        derive(u, r);
        der[i] = r[5]; der[max_loc + i] = r[6]; der[2 * max_loc + i] = r[7];
    }
    __syncthreads();
    double *w0 = reinterpret_cast<double *>(sblob);
    double *w1 = w0 + d.e_pad, *w2 = w1 + d.e_pad, *gg = w2 + d.e_pad;
    const uint32_t *lab = reinterpret_cast<const uint32_t *>(gg + d.e_pad);
    const uint16_t *rowptr = reinterpret_cast<const uint16_t *>(lab + d.e_pad);
    const uint16_t *csr = rowptr + (((d.n_own + 1) + 7) & ~7);
    for (int e = tid; e < d.n_edges; e += 128) {
        uint32_t l = lab[e];
        double x = w0[e], y = w1[e], z = w2[e], g = gg[e];
        double a[NF], b[NF], fa[5];
        load_state<NF>(raw, der, max_loc, (int)(l & 0xffff), a);
        load_state<NF>(raw, der, max_loc, (int)(l >> 16), b);
        edge_flux(a, b, x, y, z, g, fa);
        w0[e] = fa[0]; w1[e] = fa[1]; w2[e] = fa[2]; gg[e] = fa[3];
        Fx[e] = fa[4];
    }
    __syncthreads();
    node_phase<true, true, 1, true>(tid, d, sblob, raw, w0, w1, w2, gg, Fx, rowptr, csr, max_edges, told, tsf, nullptr, rk);
}
#endif

// ------------------------------------------------------------------------------------------
// variant 2p: the owner kernel as a persistent, double-buffered pipeline.  The grid is sized to the number of CTAs
// that are resident at once; CTA b works through chunks b, b+G, b+2G, ... of the launch.  While chunk j is being
// computed out of stage j&1, everything chunk j+1 needs is already on its way into the other stage:
//   * blob and owned conserved variables: bulk async copies (TMA 1-D) issued by one thread at the top of iteration j;
//   * halo node states: 8-byte cp.async requests issued by all threads at the top of iteration j from halo ids that
//     were loaded into registers one iteration earlier (so neither the id list nor the gather waits on HBM);
//   * both complete on the stage's mbarrier (expect_tx for the bulk copies, cp.async.mbarrier.arrive.noinc for every
//     thread's gather requests; phase parity (j>>1)&1), so a stage is consumed after ONE wait and no CTA barrier;
//   * chunk descriptors: a 4-entry ring in shared memory, loaded three chunks ahead through registers.
// Only the prologue (first chunk of a CTA) sees memory latency.  Per chunk there are two CTA barriers:
//   D. one thread per edge turns the edge's weights in place into its flux vector            -> barrier
//   E. owned nodes sum their incident edges (ascending file order; the fast build splits a node between two
//      threads), add the boundary entries, apply the Runge-Kutta update and store to HBM straight from registers;
//      a warp that is done waits for the next stage and computes the next chunk's derived quantities (p, |v|+c,
//      1/rho per staged node) in the shadow of the warps still summing                        -> barrier
// The update operands (old_variables, step_factors; FUSE) are single-buffered: they are requested at the top of their
// own chunk and only read by phase E.
// shared: bar[3] | desc ring [4] | stage 0 | stage 1 | told [max_own][5] | tsf [max_own] | Fx [NFL-4][max_edges] |
//         der [3][max_loc] (fast build);   stage = blob [max_blob] | raw [max_loc][5]
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

constexpr int PIPE_HG = 8;          // halo ids a thread keeps in registers (covers 5*n_halo <= 8*blockDim)
constexpr int PIPE_HEAD = 32 + 4 * (int)sizeof(OwnerChunkDesc);   // 3 mbarriers + descriptor ring
static_assert(sizeof(OwnerChunkDesc) == 48, "descriptor ring moves three 16-byte pieces per chunk");

struct PipeLayout {
    size_t raw_d, told_d, tsf_d, stage_bytes, total;
};
__host__ __device__ inline PipeLayout pipe_layout(int max_loc, int max_edges, int max_blob, int max_own, bool fuse, int nrec, int nfl,
                                                  int stages)
{
    PipeLayout l;
    l.raw_d = ((size_t)max_loc * 5 + 1) & ~(size_t)1;
    l.told_d = fuse ? (((size_t)max_own * 5 + 1) & ~(size_t)1) : 0;
    l.tsf_d = fuse ? (((size_t)max_own + 1) & ~(size_t)1) : 0;
    l.stage_bytes = (size_t)max_blob + 8 * l.raw_d;
    l.total = PIPE_HEAD + (size_t)stages * l.stage_bytes + 8 * (l.told_d + l.tsf_d + (size_t)(nfl - 4) * max_edges + (size_t)(nrec - 5) * max_loc);
    return l;
}

template <int THREADS, int MINB, int STAGES, bool OVERWRITE, bool FUSE>
__global__ void __launch_bounds__(THREADS, MINB)
flux_owner_pipe_kernel(int max_loc, int max_edges, int max_blob, int n_list, const OwnerChunkDesc *__restrict__ descs,
                       const int *__restrict__ chunk_list, const int *__restrict__ halo_gid,
                       const unsigned char *__restrict__ blob, const double *__restrict__ var, double *__restrict__ flux,
                       RkStageArgs rk)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int NREC = NF, NFL = NFLUX;
    constexpr bool DB = STAGES == 2;       // double-buffered stages; otherwise one stage, the next chunk is prefetched into L2
    const PipeLayout lay = pipe_layout(max_loc, max_edges, max_blob, rk.max_own, FUSE, NREC, NFL, STAGES);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smraw);
    unsigned char *dring = smraw + 32;
    unsigned char *stage0 = smraw + PIPE_HEAD;
    double *told = reinterpret_cast<double *>(stage0 + (size_t)STAGES * lay.stage_bytes);
    double *tsf = told + lay.told_d;
    double *Fx = tsf + lay.tsf_d;
    double *der = Fx + (size_t)(NFL - 4) * max_edges;
    constexpr int nthreads = THREADS;
    const int tid = threadIdx.x, G = gridDim.x;
    const int first = blockIdx.x;

    auto stage_raw = [&](int st) { return reinterpret_cast<double *>(stage0 + (size_t)st * lay.stage_bytes + max_blob); };
    // bulk copies of chunk `c` into stage `st` (one thread)
    auto issue_bulk = [&](const OwnerChunkDesc &c, int st) {
        const uint32_t own_bulk = owned_bulk_bytes(c.n_own);
        fence_proxy_async();          // earlier generic-proxy accesses to this stage are ordered before the async writes
        mbar_expect_tx(bar + st, (uint32_t)c.blob_bytes + own_bulk);
        bulk_g2s(stage0 + (size_t)st * lay.stage_bytes, blob + c.blob_off, (uint32_t)c.blob_bytes, bar + st);
        if (own_bulk) bulk_g2s(stage_raw(st), var + (size_t)c.node0 * 5, own_bulk, bar + st);
    };
    // old_variables / step_factor tiles of chunk `c` (one thread)
    auto issue_update_operands = [&](const OwnerChunkDesc &c) {
        const uint32_t old_b = owned_bulk_bytes(c.n_own), sf_b = ((uint32_t)c.n_own * 8u) & ~15u;
        fence_proxy_async();
        mbar_expect_tx(bar + 2, old_b + sf_b);
        if (old_b) bulk_g2s(told, rk.old + (size_t)c.node0 * 5, old_b, bar + 2);
        if (sf_b) bulk_g2s(tsf, rk.sf + c.node0, sf_b, bar + 2);
    };
    // halo ids of chunk `c`, element-consecutive mapping: this thread's k-th element is f = k*nthreads + tid
    auto load_hg = [&](const OwnerChunkDesc &c, int hg[PIPE_HG]) {
        const int nh5 = c.n_halo * 5;
#pragma unroll
        for (int k = 0; k < PIPE_HG; k++) {
            int f = k * nthreads + tid;
            hg[k] = f < nh5 ? __ldg(halo_gid + c.halo_off + f / 5) : -1;
        }
    };
    // halo gather of chunk `c` into stage `st`: every thread issues its requests and lets the stage's mbarrier count them
    auto issue_gather = [&](const OwnerChunkDesc &c, int st, const int hg[PIPE_HG]) {
        double *sraw = stage_raw(st);
        double *hraw = sraw + (size_t)c.n_own * 5;
        const int nh5 = c.n_halo * 5;
#pragma unroll
        for (int k = 0; k < PIPE_HG; k++) {
            int f = k * nthreads + tid;
            if (hg[k] >= 0) cp_async8(hraw + f, var + (size_t)hg[k] * 5 + (f % 5));
        }
        for (int f = PIPE_HG * nthreads + tid; f < nh5; f += nthreads)      // chunks with very long halo lists
            cp_async8(hraw + f, var + (size_t)__ldg(halo_gid + c.halo_off + f / 5) * 5 + (f % 5));
        cp_async_arrive_noinc(bar + st);
        // an odd owned run (last chunk of a level only): its last 8 bytes are not part of the bulk copy
        if (tid == 32 && (c.n_own & 1)) sraw[c.n_own * 5 - 1] = __ldg(var + (size_t)(c.node0 + c.n_own) * 5 - 1);
    };
    // derived quantities once per staged node of chunk `c` (stage `st` must be complete)
    auto derive_stage = [&](const OwnerChunkDesc &c, int st) {
#ifndef MGCFD_EXACT
        const double *sraw = stage_raw(st);
        const int nloc = c.n_own + c.n_halo;
        for (int i = tid; i < nloc; i += nthreads) {
            double u[5], r[8];
#pragma unroll
            for (int v = 0; v < 5; v++) u[v] = sraw[i * 5 + v];
            derive(u, r);
            der[i] = r[5]; der[max_loc + i] = r[6]; der[2 * max_loc + i] = r[7];
        }
#endif
    };
    auto chunk_id = [&](int pos) { return chunk_list ? __ldg(chunk_list + pos) : pos; };
    auto ring = [&](int j) { return reinterpret_cast<OwnerChunkDesc *>(dring + (size_t)(j & 3) * sizeof(OwnerChunkDesc)); };
    auto desc_piece = [&](int cid) { return __ldg(reinterpret_cast<const uint4 *>(descs + cid) + tid); };     // tid < 3

    // ---- prologue: descriptors of the first three chunks, mbarriers, first chunk's stage and derived quantities
    int cid_ahead = 0;                 // threads 0..2: id of the chunk four positions ahead, loaded one iteration early
    if (tid < 3) {
        for (int j = 0; j < 3; j++) {
            int pos = first + j * G;
            if (pos < n_list) reinterpret_cast<uint4 *>(ring(j))[tid] = desc_piece(chunk_id(pos));
        }
        int pos3 = first + 3 * G;
        cid_ahead = pos3 < n_list ? chunk_id(pos3) : 0;
    }
    if (tid == 0) {
        mbar_init(bar, 1 + THREADS);
        mbar_init(bar + 1, 1 + THREADS);
        mbar_init(bar + 2, 1);
    }
    __syncthreads();
    int hg[PIPE_HG];
    if (DB) {
        const OwnerChunkDesc c0 = *ring(0);
        if (tid == 0) issue_bulk(c0, 0);
        load_hg(c0, hg);
        issue_gather(c0, 0, hg);
        if (first + G < n_list) load_hg(*ring(1), hg);
        mbar_wait(bar, 0);
        __syncthreads();              // the hand-copied tail element of an odd owned run
        derive_stage(c0, 0);
        __syncthreads();
    } else {
        load_hg(*ring(0), hg);
    }

    for (int j = 0, pos = first; pos < n_list; j++, pos += G) {
        const int st = DB ? (j & 1) : 0;
        const OwnerChunkDesc d = *ring(j);
        const bool has_next = pos + G < n_list;
        if (FUSE && tid == 0) issue_update_operands(d);
        if (DB) {
            // ---- top: everything the next chunk needs is requested now
            if (has_next) {
                const OwnerChunkDesc dn = *ring(j + 1);
                if (tid == 0) issue_bulk(dn, st ^ 1);
                issue_gather(dn, st ^ 1, hg);
            }
        } else {
            // ---- top: this chunk's stage is requested (its halo ids are in registers, its lines were prefetched into
            //      L2 during the previous chunk), then the next chunk's contiguous inputs are prefetched into L2
            if (tid == 0) issue_bulk(d, 0);
            issue_gather(d, 0, hg);
            if (has_next) {
                const OwnerChunkDesc dn = *ring(j + 1);
                load_hg(dn, hg);                                      // consumed by the L2 prefetch in phase E and the next gather
                if (tid == 0) {
                    const uint32_t own_b = owned_bulk_bytes(dn.n_own), sf_b = ((uint32_t)dn.n_own * 8u) & ~15u;
                    bulk_prefetch_l2(blob + dn.blob_off, (uint32_t)dn.blob_bytes);
                    if (own_b) bulk_prefetch_l2(var + (size_t)dn.node0 * 5, own_b);
                    if (FUSE && own_b) bulk_prefetch_l2(rk.old + (size_t)dn.node0 * 5, own_b);
                    if (FUSE && sf_b) bulk_prefetch_l2(rk.sf + dn.node0, sf_b);
                }
            }
            mbar_wait(bar, (uint32_t)j & 1u);
            __syncthreads();          // the hand-copied tail element of an odd owned run
            derive_stage(d, 0);
            __syncthreads();
        }

        unsigned char *sblob = stage0 + (size_t)st * lay.stage_bytes;
        const double *raw = stage_raw(st);
        double *w0 = reinterpret_cast<double *>(sblob);
        double *w1 = w0 + d.e_pad, *w2 = w1 + d.e_pad, *gg = w2 + d.e_pad;
        const uint32_t *lab = reinterpret_cast<const uint32_t *>(gg + d.e_pad);
        const uint16_t *rowptr = reinterpret_cast<const uint16_t *>(lab + d.e_pad);
        const uint16_t *csr = rowptr + (((d.n_own + 1) + 7) & ~7);

        // ---- D. one thread per edge: the edge's weights are replaced in place by its flux vector
        auto one_edge = [&](int e) {
            uint32_t l = lab[e];
            int la = l & 0xffff, lb = l >> 16;
            double x = w0[e], y = w1[e], z = w2[e], g = gg[e];
            double a[NREC], b[NREC];
            load_state<NREC>(raw, der, max_loc, la, a);
            load_state<NREC>(raw, der, max_loc, lb, b);
            double fa[5];
#ifdef MGCFD_EXACT
            double fb[5];
            edge_flux(a, b, x, y, z, g, fa, fb);
#pragma unroll
            for (int v = 0; v < 5; v++) Fx[(1 + v) * max_edges + e] = fb[v];
#else
            edge_flux(a, b, x, y, z, g, fa);
#endif
            w0[e] = fa[0]; w1[e] = fa[1]; w2[e] = fa[2]; gg[e] = fa[3];
            Fx[e] = fa[4];
        };
        for (int e = tid; e < d.n_edges; e += nthreads) one_edge(e);
        __syncthreads();

        // requests whose results are needed one iteration from now: halo ids of the chunk after next, descriptor of
        // the chunk three ahead (issued here so that the registers are not live across the edge phase)
        if (DB && pos + 2 * G < n_list) load_hg(*ring(j + 2), hg);
        uint4 dreg = make_uint4(0, 0, 0, 0);
        const bool ring_fill = pos + 3 * G < n_list;
        if (tid < 3) {
            if (ring_fill) dreg = desc_piece(cid_ahead);              // stored into the ring at the end of the iteration
            int pos4 = pos + 4 * G;
            cid_ahead = pos4 < n_list ? chunk_id(pos4) : 0;
        }

        // ---- E. owned nodes: incident-edge sums in ascending file order (the fast build splits a node's incidences
        //         between two threads when the CTA has them), boundary entries, Runge-Kutta update, store
        if (FUSE) mbar_wait(bar + 2, (uint32_t)j & 1u);               // old_variables / step_factor tiles
#ifdef MGCFD_EXACT
        node_phase<OVERWRITE, FUSE, 0>(tid, d, sblob, raw, w0, w1, w2, gg, Fx, rowptr, csr, max_edges, told, tsf, flux, rk);
#else
        if (4 * rk.max_own <= nthreads)
            node_phase<OVERWRITE, FUSE, 2>(tid, d, sblob, raw, w0, w1, w2, gg, Fx, rowptr, csr, max_edges, told, tsf, flux, rk);
        else if (2 * rk.max_own <= nthreads)
            node_phase<OVERWRITE, FUSE, 1>(tid, d, sblob, raw, w0, w1, w2, gg, Fx, rowptr, csr, max_edges, told, tsf, flux, rk);
        else
            node_phase<OVERWRITE, FUSE, 0>(tid, d, sblob, raw, w0, w1, w2, gg, Fx, rowptr, csr, max_edges, told, tsf, flux, rk);
#endif
        // ---- next chunk's derived quantities, in the shadow of the warps still in E
        if (DB) {
            if (has_next) {
                mbar_wait(bar + (st ^ 1), (uint32_t)((j + 1) >> 1) & 1u);
                derive_stage(*ring(j + 1), st ^ 1);
            }
        } else if (has_next) {
            // the next chunk's halo rows into L2 (first and last element of a row cover the sectors it touches)
#pragma unroll
            for (int k = 0; k < PIPE_HG; k++) {
                const int c5 = (k * nthreads + tid) % 5;
                if (hg[k] >= 0 && (c5 == 0 || c5 == 4)) prefetch_l2(var + (size_t)hg[k] * 5 + c5);
            }
        }
        if (tid < 3 && ring_fill) reinterpret_cast<uint4 *>(ring(j + 3))[tid] = dreg;
        __syncthreads();              // this stage, the operand tiles and Fx are free; der and the ring are published
    }
}

#ifndef MGCFD_EXACT
// ------------------------------------------------------------------------------------------
// variant 4: emit.  Same owner chunks, TWO threads per owned node.  Every edge of the chunk is evaluated ONCE, by
// the thread pair of its lowest-numbered owned endpoint (its "emitter"), which keeps its own state and its own flux
// sum in registers; only when the other endpoint is owned too is the edge's flux vector parked in shared memory
// (in place of the edge's weights) for that node to subtract afterwards.  Compared with the owner kernel this
// halves the state reads (one neighbour per edge instead of two endpoints), parks fewer flux vectors and reads each
// of them once instead of twice.  A node's emitted edges are split between its two threads (even / odd entries);
// the half-rows are stored sliced-ELL (sorted by length, slices of 32 threads padded, column-major) with pre-signed
// weights inside the chunk's blob, which one bulk async copy brings into shared memory.  Fast arithmetic only:
// sums are not in file order.
// shared: mbarrier | blob (w0 w1 w2 g [n_ent] doubles, ent [n_ent] u32, rowptr2 | csr2 u16) | Fx[max_ent] |
//         raw[max_loc][5] | der[3][max_loc] | told | tsf
// ------------------------------------------------------------------------------------------
template <bool OVERWRITE, bool FUSE>
__global__ void __launch_bounds__(256, 3)
flux_emit_kernel(int max_loc, int max_ent, int max_blob, const EmitChunkDesc *__restrict__ descs,
                 const int *__restrict__ chunk_list, const int *__restrict__ halo_gid,
                 const uint16_t *__restrict__ row_node, const uint16_t *__restrict__ row_cnt,
                 const unsigned char *__restrict__ blob, const double *__restrict__ var, double *__restrict__ flux,
                 RkStageArgs rk)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smraw);
    unsigned char *sblob = smraw + 16;
    double *Fx = reinterpret_cast<double *>(sblob + max_blob);
    double *raw = Fx + max_ent;
    double *der = raw + (size_t)max_loc * 5;
    double *told = raw + (((size_t)max_loc * 8 + 1) & ~(size_t)1);
    double *tsf = told + (((size_t)rk.max_own * 5 + 1) & ~(size_t)1);
    const int chunk = chunk_list ? chunk_list[blockIdx.x] : blockIdx.x;
    const EmitChunkDesc d = descs[chunk];
    const int tid = threadIdx.x, lane = tid & 31, slice = tid >> 5, half = tid & 1;
    const int nloc = d.n_own + d.n_halo;
    const uint32_t old_bulk = FUSE ? owned_bulk_bytes(d.n_own) : 0u, sf_bulk = FUSE ? (((uint32_t)d.n_own * 8u) & ~15u) : 0u;

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)d.blob_bytes + owned_bulk_bytes(d.n_own) + old_bulk + sf_bulk);
        bulk_g2s(sblob, blob + d.blob_off, (uint32_t)d.blob_bytes, bar);
        if (FUSE) {
            if (old_bulk) bulk_g2s(told, rk.old + (size_t)d.node0 * 5, old_bulk, bar);
            if (sf_bulk) bulk_g2s(tsf, rk.sf + d.node0, sf_bulk, bar);
        }
    }
    const int me = row_node[(size_t)chunk * 256 + tid];               // local owned index or 0xffff
    const int cnt = row_cnt[(size_t)chunk * 256 + tid];               // entries of this thread's half-row
    int len = 0, base = 0;
#pragma unroll
    for (int s = 0; s < 8; s++) {
        int L = d.slice_len[s];
        if (s < slice) base += L * 32;
        if (s == slice) len = L;
    }
    stage_tile(raw, bar, d.node0, d.n_own, d.n_halo, halo_gid + d.halo_off, var, tid, blockDim.x);
    __syncthreads();
    mbar_wait(bar, 0);
    for (int i = tid; i < nloc; i += blockDim.x) {
        double u[5], r[8];
#pragma unroll
        for (int v = 0; v < 5; v++) u[v] = raw[i * 5 + v];
        derive(u, r);
        der[i] = r[5]; der[max_loc + i] = r[6]; der[2 * max_loc + i] = r[7];
    }
    __syncthreads();

    double *w0 = reinterpret_cast<double *>(sblob);
    double *w1 = w0 + d.n_ent, *w2 = w1 + d.n_ent, *gg = w2 + d.n_ent;
    const uint32_t *ent = reinterpret_cast<const uint32_t *>(gg + d.n_ent);
    const uint16_t *rowptr2 = reinterpret_cast<const uint16_t *>(ent + d.n_ent);
    const uint16_t *csr2 = rowptr2 + d.rowptr_pad;

    const bool valid = me != 0xffff;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (!OVERWRITE && valid && half == 0) {
#pragma unroll
        for (int v = 0; v < 5; v++) acc[v] = flux[(size_t)(d.node0 + me) * 5 + v];
    }
    const int self = valid ? me : 0;
    double a[8];
    load_state<8>(raw, der, max_loc, self, a);
    for (int k = 0; k < len; k++) {
        const int idx = base + k * 32 + lane;
        const uint32_t en = ent[idx];
        const double x = w0[idx], y = w1[idx], z = w2[idx], g = gg[idx];
        double b[8], F[5];
        load_state<8>(raw, der, max_loc, (int)(en & 0xffff), b);
        edge_flux(a, b, x, y, z, g, F);                               // weights pre-signed: the emitter is end "a"
        if (k < cnt) {
#pragma unroll
            for (int v = 0; v < 5; v++) acc[v] += F[v];
            if (en & 0x10000u) {                                       // the other end is owned: it subtracts this vector
                w0[idx] = F[0]; w1[idx] = F[1]; w2[idx] = F[2]; gg[idx] = F[3];      // slot idx is private to this thread
                Fx[idx] = F[4];
            }
        }
    }
    __syncthreads();
    if (valid) {
        for (int j = rowptr2[me] + half; j < rowptr2[me + 1]; j += 2) {
            const int slot = csr2[j];
            acc[0] -= w0[slot]; acc[1] -= w1[slot]; acc[2] -= w2[slot]; acc[3] -= gg[slot]; acc[4] -= Fx[slot];
        }
    }
#pragma unroll
    for (int v = 0; v < 5; v++) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], 1);       // the node's two half-rows
    if (valid && half == 0 && FUSE && d.has_bnd) {
        int j0 = rk.bnd_ptr[d.node0 + me], j1 = rk.bnd_ptr[d.node0 + me + 1];
        if (j1 > j0) {
            double u[5];
#pragma unroll
            for (int v = 0; v < 5; v++) u[v] = raw[me * 5 + v];
            bnd_apply(u, acc, j0, j1, rk.b_group, rk.b_wt, rk.c);
        }
    }
    __syncthreads();                                                  // every read of the state tile is done
    if (valid && half == 0) {
#pragma unroll
        for (int v = 0; v < 5; v++) raw[me * 5 + v] = acc[v];
    }
    __syncthreads();
    finish_chunk<FUSE>(raw, d.node0, d.n_own, flux, rk, told, tsf, old_bulk, sf_bulk);
}

inline size_t emit_smem(int max_loc, int max_ent, int max_blob, int max_own)
{
    return 16 + (size_t)max_blob + ((size_t)max_ent + (size_t)max_loc * 8 + (size_t)max_own * 6 + 4) * 8;
}

inline int launch_emit(cudaStream_t s, const FluxArgs &a, const EmitPlanDev &p)
{
    const int grid = a.chunk_list ? a.n_list : p.n_chunks;
    if (grid == 0) return 0;
    size_t smem = emit_smem(p.max_loc, p.max_ent, p.max_blob, p.max_own);
    RkStageArgs ra{};
    if (a.rk) ra = *a.rk;
    ra.max_own = p.max_own;
    const int threads = p.max_own <= 64 ? 128 : 256;                  // two threads per owned node
#define EMIT_ARGS p.max_loc, p.max_ent, p.max_blob, p.desc, a.chunk_list, p.halo_gid, p.row_node, p.row_cnt, p.blob, a.var, a.flux, ra
    if (a.rk)
        flux_emit_kernel<true, true><<<grid, threads, smem, s>>>(EMIT_ARGS);
    else if (a.overwrite)
        flux_emit_kernel<true, false><<<grid, threads, smem, s>>>(EMIT_ARGS);
    else
        flux_emit_kernel<false, false><<<grid, threads, smem, s>>>(EMIT_ARGS);
#undef EMIT_ARGS
    return 1;
}
#include "stage_kernel.cuh"
#endif  // !MGCFD_EXACT

// ------------------------------------------------------------------------------------------
// variant 3: node gather ("pull").  Same owner chunks, but one thread per owned node walks the node's
// incident edges and evaluates each edge from its own side only (every interior edge is evaluated
// twice, once per endpoint).  Rows are stored sliced-ELL: the chunk's nodes are sorted by degree,
// each warp-slice is padded to its longest row and stored column-major, so neighbour ids and the
// (pre-signed) weights stream from HBM fully coalesced; only the neighbour's state is gathered, from
// shared memory.  No per-edge staging, no scatter, sums in ascending file order.
// shared: mbarrier | raw[max_loc][5] (AoS state tile, reused for the output) | der[3][max_loc] (fast build)
// ------------------------------------------------------------------------------------------
template <bool STREAM, bool OVERWRITE>
__global__ void __launch_bounds__(256, 3)
flux_gather_kernel(int max_loc, const GatherChunkDesc *__restrict__ descs, const int *__restrict__ halo_gid,
                   const uint16_t *__restrict__ row_node, const uint16_t *__restrict__ row_deg,
                   const uint32_t *__restrict__ ent, const double *__restrict__ pw0, const double *__restrict__ pw1,
                   const double *__restrict__ pw2, const double *__restrict__ pg, const double *__restrict__ var,
                   double *__restrict__ flux)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int NREC = STREAM ? 5 : NF;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smraw);
    double *raw = reinterpret_cast<double *>(smraw + 16);               // conserved variables, AoS [max_loc][5]
    double *der = raw + (size_t)max_loc * 5;                            // p, |v|+c, 1/rho planes (fast build only)
    const GatherChunkDesc d = descs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, slice = tid >> 5;
    const int nloc = d.n_own + d.n_halo;

    // 1. state tile: owned run by one bulk async copy, halo nodes by 8-byte async copies
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, owned_bulk_bytes(d.n_own));
    }
    stage_tile(raw, bar, d.node0, d.n_own, d.n_halo, halo_gid + d.halo_off, var, tid, blockDim.x);
    __syncthreads();
    mbar_wait(bar, 0);
#ifndef MGCFD_EXACT
    // 2. derived quantities once per staged node
    if (!STREAM) {
        for (int i = tid; i < nloc; i += blockDim.x) {
            double u[5], r[8];
#pragma unroll
            for (int v = 0; v < 5; v++) u[v] = raw[i * 5 + v];
            derive(u, r);
            der[i] = r[5]; der[max_loc + i] = r[6]; der[2 * max_loc + i] = r[7];
        }
        __syncthreads();
    }
#endif

    // 3. one thread per owned node (degree-sorted order), rows in sliced-ELL layout
    const int me = row_node[(size_t)blockIdx.x * 256 + tid];          // local owned index or 0xffff
    const int deg = row_deg[(size_t)blockIdx.x * 256 + tid];
    int len = 0;
    long long base = d.ent_off;
#pragma unroll
    for (int s = 0; s < 8; s++) {
        int L = d.slice_len[s];
        if (s < slice) base += (long long)L * 32;
        if (s == slice) len = L;
    }
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (me != 0xffff) {
        if (!OVERWRITE) {
#pragma unroll
            for (int v = 0; v < 5; v++) acc[v] = flux[(size_t)(d.node0 + me) * 5 + v];
        }
    }
    const int self = me != 0xffff ? me : 0;
    double a[NREC];
    load_state<NREC>(raw, der, max_loc, self, a);
    for (int j = 0; j < len; j++) {
        long long idx = base + (long long)j * 32 + lane;
        uint32_t en = __ldg(ent + idx);
        double x = __ldg(pw0 + idx), y = __ldg(pw1 + idx), z = __ldg(pw2 + idx), g = __ldg(pg + idx);
        int nb = en & 0xffff;
        double b[NREC];
        load_state<NREC>(raw, der, max_loc, nb, b);
        double fa[5], fb[5];
        if (STREAM) {
            // the row stores the edge as seen from this node; bit 16 says whether this node is the edge's end b
            if (en & 0x10000u) { stream_flux(b, a, x, y, z, fb, fa); } else { stream_flux(a, b, x, y, z, fa, fb); }
        } else {
#ifdef MGCFD_EXACT
            // reference orientation: weights are stored unsigned here, evaluate (a_edge, b_edge) and take my side
            if (en & 0x10000u) { edge_flux(b, a, x, y, z, g, fb, fa); } else { edge_flux(a, b, x, y, z, g, fa, fb); }
#else
            edge_flux(a, b, x, y, z, g, fa);      // weights pre-signed: this node is always "a"
#endif
        }
        if (j < deg) {
#pragma unroll
            for (int v = 0; v < 5; v++) acc[v] += fa[v];
        }
    }
    // 4. coalesced store through shared memory (the state tile is dead after the barrier)
    __syncthreads();
    if (me != 0xffff) {
#pragma unroll
        for (int v = 0; v < 5; v++) raw[me * 5 + v] = acc[v];
    }
    __syncthreads();
    double *out = flux + (size_t)d.node0 * 5;
    for (int f = tid; f < d.n_own * 5; f += blockDim.x) out[f] = raw[f];
}

// ------------------------------------------------------------------------------------------
// compute_bnd_node_flux_kernel, flux.h:14-39 (+ flux_boundary.elem_func, flux_wall.elem_func).
// One thread per UNIQUE boundary-touched node; its boundary entries are applied in ascending
// file order, so several entries on one node need no atomics and the sum order is OP2-seq's.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
bnd_flux_kernel(int n_unique, const int *__restrict__ bu_node, const int *__restrict__ bu_ptr,
                const int *__restrict__ b_group, const double *__restrict__ b_wt,
                const double *__restrict__ var, double *__restrict__ flux, DevConsts c)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_unique) return;
    int node = bu_node[t];
    double u[5], fl[5];
    load5(var + (size_t)node * 5, u);
    double *out = flux + (size_t)node * 5;
#pragma unroll
    for (int k = 0; k < 5; k++) fl[k] = out[k];
    bnd_apply(u, fl, bu_ptr[t], bu_ptr[t + 1], b_group, b_wt, c);
#pragma unroll
    for (int k = 0; k < 5; k++) out[k] = fl[k];
}

inline int launch_bnd(cudaStream_t s, int n_unique, const int *bu_node, const int *bu_ptr, const int *b_group,
                      const double *b_wt, const double *var, double *flux, const DevConsts &c)
{
    if (n_unique == 0) return 0;
    bnd_flux_kernel<<<(n_unique + 127) / 128, 128, 0, s>>>(n_unique, bu_node, bu_ptr, b_group, b_wt, var, flux, c);
    return 1;
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
inline size_t colour_smem(int max_nodes, bool stream) { return (size_t)((stream ? 5 : NF) + 5) * max_nodes * sizeof(double); }
inline size_t owner_smem(int max_loc, int max_edges, int max_blob, bool stream, int fuse_max_own = 0)
{
    return 16 + (size_t)max_blob +
           ((size_t)(stream ? 5 : NF) * max_loc + (size_t)((stream ? 10 : NFLUX) - 4) * max_edges + (size_t)fuse_max_own * 6 + 4) *
               sizeof(double);
}

inline int launch_atomic(cudaStream_t s, const FluxArgs &a, const AtomicPlanDev &p)
{
    if (a.n_edges == 0) return 0;
    int grid = (a.n_edges + 255) / 256;
    if (a.stream_kernel)
        flux_atomic_kernel<true><<<grid, 256, 0, s>>>(a.n_edges, a.n_owned, p.nodes, p.w, a.var, a.flux);
    else
        flux_atomic_kernel<false><<<grid, 256, 0, s>>>(a.n_edges, a.n_owned, p.nodes, p.w, a.var, a.flux);
    return 1;
}

inline int launch_colour(cudaStream_t s, const FluxArgs &a, const ColourPlanDev &p, const ColourPlanHost &h)
{
    int launches = 0;
    size_t smem = colour_smem(h.max_nodes, a.stream_kernel);
    for (int c = 0; c < h.n_block_colours; c++) {
        int slot0 = h.colour_start[c], nb = h.colour_start[c + 1] - slot0;
        if (nb == 0) continue;
        if (a.stream_kernel)
            flux_colour_kernel<true><<<nb, h.block_edges, smem, s>>>(slot0, h.max_nodes, a.n_owned, p.blk_edge0, p.blk_node0,
                                                                    p.blk_ncol, p.node_gid, p.lab, p.ecol, p.w, a.var, a.flux);
        else
            flux_colour_kernel<false><<<nb, h.block_edges, smem, s>>>(slot0, h.max_nodes, a.n_owned, p.blk_edge0, p.blk_node0,
                                                                     p.blk_ncol, p.node_gid, p.lab, p.ecol, p.w, a.var, a.flux);
        launches++;
    }
    return launches;
}

inline size_t gather_smem(int max_loc, bool stream) { return 16 + (size_t)(stream ? 5 : NF) * max_loc * sizeof(double); }

inline int launch_gather(cudaStream_t s, const FluxArgs &a, const GatherPlanDev &p, int n_chunks, int max_loc)
{
    if (n_chunks == 0) return 0;
    size_t smem = gather_smem(max_loc, a.stream_kernel);
#define GATHER_ARGS max_loc, p.desc, p.halo_gid, p.row_node, p.row_deg, p.ent, p.w0, p.w1, p.w2, p.g, a.var, a.flux
    if (a.stream_kernel)
        flux_gather_kernel<true, false><<<n_chunks, 256, smem, s>>>(GATHER_ARGS);
    else if (a.overwrite)
        flux_gather_kernel<false, true><<<n_chunks, 256, smem, s>>>(GATHER_ARGS);
    else
        flux_gather_kernel<false, false><<<n_chunks, 256, smem, s>>>(GATHER_ARGS);
#undef GATHER_ARGS
    return 1;
}

// persistent pipelined owner kernel: grid = resident CTAs (balanced over the waves the launch needs).
// Returns -1 when the two stages do not fit in shared memory (the caller falls back to one CTA per chunk).
inline int launch_owner_pipe(cudaStream_t s, const FluxArgs &a, const OwnerPlanDev &p, const OwnerPlanHost &h, int threads, int stages)
{
    const int n_list = a.chunk_list ? a.n_list : h.n_chunks;
    const bool fuse = a.rk != nullptr;
    const PipeLayout lay = pipe_layout(h.max_loc, h.max_edges, h.dev_max_blob, h.max_own, fuse, NF, NFLUX, stages);
    if (lay.total > 227 * 1024) return -1;
    RkStageArgs ra{};
    if (a.rk) ra = *a.rk;
    ra.max_own = h.max_own;
    const int which = fuse ? 0 : (a.overwrite ? 1 : 2);
    static std::map<std::tuple<int, int, int, size_t, int>, int> slots_cache;   // (device, kernel, threads, smem, cap) -> resident CTAs
    int dev = 0;
    cudaGetDevice(&dev);
    const char *cap_s = getenv("MGCFD_OWNER_PIPE_CTAS");      // experiments: cap on resident CTAs per SM
    const int env_occ = cap_s ? atoi(cap_s) : 0;
#define PIPE_ROW(T, B, S) {(const void *)flux_owner_pipe_kernel<T, B, S, true, true>, (const void *)flux_owner_pipe_kernel<T, B, S, true, false>, \
                           (const void *)flux_owner_pipe_kernel<T, B, S, false, false>}
    // rows: two stages with 128 threads x 4 CTAs/SM and 256 threads x 2; one stage with 128 x 6 and 256 x 3
    static const void *const kernels[4][3] = {PIPE_ROW(128, 4, 2), PIPE_ROW(256, 2, 2), PIPE_ROW(128, 6, 1), PIPE_ROW(256, 3, 1)};
#undef PIPE_ROW
    if ((threads != 128 && threads != 256) || (stages != 1 && stages != 2)) return -1;
    const int row = (stages == 2 ? 0 : 2) + (threads == 128 ? 0 : 1);
    const void *kernel = kernels[row][which];
    auto key = std::make_tuple(dev, which, row, lay.total, env_occ);
    auto it = slots_cache.find(key);
    if (it == slots_cache.end()) {
        int occ = 0, sms = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, lay.total);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (env_occ > 0 && env_occ < occ) occ = env_occ;
        if (occ < 1 || sms < 1) return -1;
        it = slots_cache.emplace(key, occ * sms).first;
    }
    const int slots = it->second;
    const int waves = (n_list + slots - 1) / slots;
    const int grid = (n_list + waves - 1) / waves;
    int max_loc = h.max_loc, max_edges = h.max_edges, max_blob = h.dev_max_blob, n = n_list;
    const OwnerChunkDesc *desc = p.desc;
    const int *list = a.chunk_list, *hgid = p.halo_gid;
    const unsigned char *blob = p.blob;
    const double *var = a.var;
    double *flux = a.flux;
    void *args[] = {&max_loc, &max_edges, &max_blob, &n, &desc, &list, &hgid, &blob, &var, &flux, &ra};
    if (cudaLaunchKernel(kernel, dim3(grid), dim3(threads), args, lay.total, s) != cudaSuccess) return -1;
    return 1;
}

inline int launch_owner(cudaStream_t s, const FluxArgs &a, const OwnerPlanDev &p, const OwnerPlanHost &h)
{
    if (h.n_chunks == 0) return 0;
    size_t smem = owner_smem(h.max_loc, h.max_edges, h.dev_max_blob, a.stream_kernel);
    RkStageArgs none{};
    const int grid = a.chunk_list ? a.n_list : h.n_chunks;
    if (grid == 0) return 0;
    // CTA size: 128 threads for chunks of up to 64 owned nodes (the default; more, smaller CTAs hide the staging
    // latency better: 0.49 -> 0.53 of the HBM roofline on an 18.75M-node deck), 256 threads for larger chunks.
    // MGCFD_OWNER_THREADS overrides for experiments (it must cover the largest chunk: one thread per owned node).
    const char *thr_s = getenv("MGCFD_OWNER_THREADS");
    const int env_threads = thr_s ? atoi(thr_s) : 0;
    int threads = h.max_own <= 64 ? 128 : 256;
    if ((env_threads == 64 || env_threads == 128 || env_threads == 256) && h.max_own <= env_threads) threads = env_threads;
    // MGCFD_OWNER_PIPE: 0 = one CTA per chunk (flux_owner_kernel); 1 = persistent CTAs, one shared-memory stage, next
    // chunk prefetched into L2; 2 = persistent CTAs, two shared-memory stages
    const char *pipe_s = getenv("MGCFD_OWNER_PIPE");
    const int pipe = pipe_s ? atoi(pipe_s) : MGCFD_OWNER_PIPE_DEFAULT;
#ifndef MGCFD_EXACT
    // default: the second-generation stage kernel (stage_kernel.cuh) whenever the plan fits it
    if (a.rk && !a.stream_kernel && stage2_applies(p, h) && launch_stage2(s, a, p, h, grid)) return 1;
#endif
    if (pipe > 0 && !a.stream_kernel && threads >= 128 && !(a.rk && a.rk->push_on)) {      // (the pipelined variants have no fused push)
        int rc = launch_owner_pipe(s, a, p, h, threads, pipe >= 2 ? 2 : 1);
        if (rc >= 0) return rc;
    }
#define OWNER_ARGS h.max_loc, h.max_edges, h.dev_max_blob, p.desc, a.chunk_list, p.halo_gid, p.blob, a.var, a.flux
    if (a.stream_kernel)
        flux_owner_kernel<true, false, false><<<grid, threads, smem, s>>>(OWNER_ARGS, none);
    else if (a.rk) {
        RkStageArgs ra = *a.rk;
        ra.max_own = h.max_own;
        size_t fsmem = owner_smem(h.max_loc, h.max_edges, h.dev_max_blob, false, h.max_own);
#ifndef MGCFD_EXACT
        // MGCFD_OWNER_LEAN=1: the lean kernel (needs the fixed-stride descriptor + halo-id table of ensure_owner)
        const char *lean_s = getenv("MGCFD_OWNER_LEAN");
        if (lean_s && atoi(lean_s) == 1 && p.xtab && threads == 128 && h.max_own <= 64 && !ra.push_on) {
            flux_owner_lean_kernel<<<grid, 128, fsmem, s>>>(h.max_loc, h.max_edges, h.dev_max_blob, p.xtab, p.xs, p.hs, a.chunk_list ? a.list_offset : 0,
                                                            p.blob, a.var, ra);
            return 1;
        }
#endif
        // MGCFD_OWNER_EPILOGUE=0: node sums staged through shared memory and a separate coalesced update pass
        const char *epi_s = getenv("MGCFD_OWNER_EPILOGUE");
        if (epi_s && atoi(epi_s) == 0 && !ra.push_on)
            flux_owner_kernel<false, true, true, false><<<grid, threads, fsmem, s>>>(OWNER_ARGS, ra);
        else
            flux_owner_kernel<false, true, true, true><<<grid, threads, fsmem, s>>>(OWNER_ARGS, ra);
    }
    else if (a.overwrite)
        flux_owner_kernel<false, true, false><<<grid, threads, smem, s>>>(OWNER_ARGS, none);
    else
        flux_owner_kernel<false, false, false><<<grid, threads, smem, s>>>(OWNER_ARGS, none);
#undef OWNER_ARGS
    return 1;
}

inline std::string configure()
{
    const int max_smem = 227 * 1024;
    cudaError_t e;
#define OPT_IN(k)                                                                                  \
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);            \
    if (e != cudaSuccess) return std::string("cudaFuncSetAttribute(" #k "): ") + cudaGetErrorString(e);
    OPT_IN((flux_colour_kernel<true>));
    OPT_IN((flux_colour_kernel<false>));
    OPT_IN((flux_owner_kernel<true, false, false>));
    OPT_IN((flux_owner_kernel<false, true, false>));
    OPT_IN((flux_owner_kernel<false, false, false>));
    OPT_IN((flux_owner_kernel<false, true, true, false>));
    OPT_IN((flux_owner_kernel<false, true, true, true>));
#ifndef MGCFD_EXACT
    OPT_IN(flux_owner_lean_kernel);
    OPT_IN((rk_stage2_kernel<336, 240, true, 6>));
    OPT_IN((rk_stage2_kernel<336, 240, false, 6>));
    OPT_IN((rk_stage2_kernel<336, 240, false, 7>));
    OPT_IN((rk_stage2_kernel<400, 240, true, 6>));
    OPT_IN((rk_stage2_kernel<400, 240, false, 6>));
#endif
#define OPT_IN_PIPE(T, B, S)                                    \
    OPT_IN((flux_owner_pipe_kernel<T, B, S, true, true>));     \
    OPT_IN((flux_owner_pipe_kernel<T, B, S, true, false>));    \
    OPT_IN((flux_owner_pipe_kernel<T, B, S, false, false>));
    OPT_IN_PIPE(128, 4, 2)
    OPT_IN_PIPE(256, 2, 2)
    OPT_IN_PIPE(128, 6, 1)
    OPT_IN_PIPE(256, 3, 1)
#undef OPT_IN_PIPE
    OPT_IN((flux_gather_kernel<true, false>));
    OPT_IN((flux_gather_kernel<false, true>));
    OPT_IN((flux_gather_kernel<false, false>));
#ifndef MGCFD_EXACT
    OPT_IN((flux_emit_kernel<true, true>));
    OPT_IN((flux_emit_kernel<true, false>));
    OPT_IN((flux_emit_kernel<false, false>));
#endif
#undef OPT_IN
    return "";
}

}  // namespace FLUX_NS
}  // namespace mgcfd
