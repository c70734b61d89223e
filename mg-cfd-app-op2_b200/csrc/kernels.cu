// kernels.cu -- node-set kernels of the MG-CFD cycle and the reference-order ("exact") build of the
// flux kernels.  This translation unit is compiled with -fmad=false and keeps the reference's
// operation order, so every kernel here is bit-identical to the CPU reference element by element
// (double division and sqrt are IEEE-rounded on the device).  They are all bandwidth-bound
// streaming kernels; contraction would not make them faster.
#define MGCFD_EXACT 1
#include <cfloat>
#include <cstring>

#include "flux_kernels.cuh"

namespace mgcfd {

size_t fast_owner_smem(int max_loc, int max_edges, int max_blob);
bool fast_owner_uses_stage2(const OwnerPlanDev &p, const OwnerPlanHost &h);
size_t fast_colour_smem(int max_nodes);
size_t fast_gather_smem(int max_loc);

namespace {

constexpr int TPB = 256;
inline int blocks_for(long long n, int tpb = TPB) { return (int)((n + tpb - 1) / tpb); }

using exact::enc_min;
using exact::dec_min;

// copy_double_kernel.h:6-13 over the flattened [n*5] array
__global__ void copy_kernel(long long n, const double *__restrict__ src, double *__restrict__ dst)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

__global__ void fill_kernel(long long n, double *__restrict__ a, double v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

__global__ void permute_rows_kernel(long long total, int dim, const double *__restrict__ src, const int *__restrict__ perm,
                                    double *__restrict__ dst, bool to_internal)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    long long row = i / dim;
    int d = (int)(i - row * dim);
    long long other = (long long)perm[row] * dim + d;
    if (to_internal) dst[other] = src[i];
    else dst[i] = src[other];
}

// misc.h:10-16
__global__ void init_vars_kernel(int n, double *__restrict__ var, DevConsts c)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
#pragma unroll
        for (int v = 0; v < 5; v++) var[(size_t)i * 5 + v] = c.ff_variable[v];
    }
}

// time_stepping_kernels.h:13-32; cbrt(volume) is precomputed on the host (volumes never change after init)
__global__ void calculate_dt_kernel(int n, const double *__restrict__ var, const double *__restrict__ cbrt_vol,
                                    double *__restrict__ dt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *u = var + (size_t)i * 5;
    double rho = u[0];
    double vx = u[1] / rho, vy = u[2] / rho, vz = u[3] / rho;
    double q2 = vx * vx + vy * vy + vz * vz;
    double p = (1.4 - 1.0) * (u[4] - 0.5 * rho * q2);
    double c = sqrt(1.4 * p / rho);
    dt[i] = 0.5 * (cbrt_vol[i] / (sqrt(q2) + c));
}

// time_stepping_kernels.h:34-41: OP_MIN reduction; NaN never wins the comparison `dt < min_dt`
__global__ void min_dt_kernel(int n, const double *__restrict__ dt, unsigned long long *__restrict__ d_min)
{
    __shared__ unsigned long long smin[TPB / 32];
    unsigned long long m = ~0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double d = dt[i];
        if (d == d) {
            unsigned long long k = enc_min(d);
            if (k < m) m = k;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
        if (t < m) m = t;
    }
    if ((threadIdx.x & 31) == 0) smin[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < TPB / 32; w++)
            if (smin[w] < m) m = smin[w];
        atomicMin(d_min, m);
    }
}

// decode the reduced minimum in place and raise the deferred "min_dt < 0" flag (euler3d.cpp:480)
__global__ void finish_min_kernel(unsigned long long *d_min, int *d_flags)
{
    double m = dec_min(*d_min);
    *reinterpret_cast<double *>(d_min) = m;
    if (m < 0.0f) d_flags[1] = 1;
}

__global__ void encode_min_kernel(unsigned long long *d_min)
{
    *d_min = enc_min(*reinterpret_cast<double *>(d_min));
}

using exact::st_release_sys;
using exact::ld_acquire_sys;
using exact::bounded_wait;
using exact::ld_relaxed_sys;

// multi-rank node kernels (NodePush): every block first waits for the sources of the halo rows it reads ...
__device__ __forceinline__ void node_wait(const NodePush &P)
{
    if ((int)threadIdx.x < P.n_wait) bounded_wait<false>(P.wait_flag[threadIdx.x], *P.wait_expected[threadIdx.x], P.err_flag, P.timeout_ns);
    __syncthreads();
}
// ... stores the 5-vector of an exported node into the destinations' halo ranges ...
__device__ __forceinline__ int node_push(const NodePush &P, int node, const double v[5])
{
    const int j0 = __ldg(P.xn_ptr + node), j1 = __ldg(P.xn_ptr + node + 1);
    for (int j = j0; j < j1; j++) {
        const int2 t = __ldg(P.xn_ent + j);
        double *d = P.dst[t.x] + (size_t)t.y * 5;
#pragma unroll
        for (int k = 0; k < 5; k++) d[k] = v[k];
    }
    return j1 > j0;
}
// ... and the last block of the grid publishes the epoch to every destination and arms the next consumer (a block that
// stored into a peer orders those stores before its count with a device-scope fence; the last block issues the only
// system-scope fence before the flags go out)
__device__ __forceinline__ void node_publish(const NodePush &P, int pushed)
{
    const int any = __syncthreads_or(pushed);
    if (threadIdx.x < 32) {
        int last = 0;
        if (threadIdx.x == 0) {
            if (any) __threadfence();
            const unsigned int prev = atomicAdd(P.done, 1u);
            last = prev + 1u == gridDim.x;
            if (last) { atomicExch(P.done, 0u); __threadfence_system(); }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if ((int)threadIdx.x < P.n_dst) {        // one lane per destination: the release stores travel in parallel
                const unsigned long long e = *P.sent[threadIdx.x] + 1;
                *P.sent[threadIdx.x] = e;
                __threadfence_system();
                st_release_sys(P.dst_flag[threadIdx.x], e);
            }
            if ((int)threadIdx.x < P.n_src) *P.expected[threadIdx.x] += 1;
        }
    }
}

// fused start of a level visit: copy_double_kernel + calculate_dt_kernel + get_min_dt_kernel (euler3d.cpp:467-479)
__global__ void visit_begin_kernel(int n, const double *__restrict__ var, const double *__restrict__ cbrt_vol,
                                   double *__restrict__ old, double *__restrict__ dt, unsigned long long *__restrict__ min_slot,
                                   const __grid_constant__ MinPush mp, double *__restrict__ zero_me)
{
    __shared__ unsigned long long smin[TPB / 32];
    __shared__ int s_last;
    unsigned long long m = ~0ull;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    const double cv = i < n ? __ldg(cbrt_vol + i) : 0.0;      // static: requested before the previous kernel has finished
    pdl_wait();                                     // programmatic dependent launch (internal.h): nothing mutable is touched above
    if (i == 0 && zero_me) *zero_me = 0.0;          // level 0: the rms accumulator of this visit (euler3d.cpp:534)
    if (i < n) {
        double u[5];
#pragma unroll
        for (int v = 0; v < 5; v++) { u[v] = var[(size_t)i * 5 + v]; old[(size_t)i * 5 + v] = u[v]; }
        double rho = u[0];
        double vx = u[1] / rho, vy = u[2] / rho, vz = u[3] / rho;
        double q2 = vx * vx + vy * vy + vz * vz;
        double p = (1.4 - 1.0) * (u[4] - 0.5 * rho * q2);
        double c = sqrt(1.4 * p / rho);
        double d = 0.5 * (cv / (sqrt(q2) + c));
        dt[i] = d;
        if (d == d) m = enc_min(d);
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
        if (t < m) m = t;
    }
    if ((threadIdx.x & 31) == 0) smin[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < TPB / 32; w++)
            if (smin[w] < m) m = smin[w];
        atomicMin(min_slot, m);
        s_last = 0;
        if (mp.on) {
            // multi-rank: the last block of the grid holds the rank's minimum and sends it to every peer's mailbox
            __threadfence();
            const unsigned int prev = atomicAdd(mp.done, 1u);
            if (prev + 1u == gridDim.x) { atomicExch(mp.done, 0u); s_last = 1; }
        }
    }
    if (!mp.on) return;
    __syncthreads();
    if (!s_last) return;
    const int lane = threadIdx.x;
    if (lane < mp.n_peers) {
        const unsigned long long mine = atomicMin(min_slot, ~0ull);       // the finished reduction (atomic read)
        *mp.dst_box[lane] = mine;
        const unsigned long long e = *mp.sent[lane] + 1;
        *mp.sent[lane] = e;
        __threadfence_system();
        st_release_sys(mp.dst_flag[lane], e);
        *mp.expected[lane] += 1;                                           // what the step-factor kernel waits for
    }
}

// compute_step_factor_kernel reading the reduced minimum from its slot; also re-arms the other slot for the next
// visit of this level, publishes min_dt and raises the deferred min_dt < 0 flag (euler3d.cpp:480)
__global__ void step_factor_fused_kernel(int n, const double *__restrict__ vol, const unsigned long long *__restrict__ min_slot,
                                         unsigned long long *__restrict__ next_slot, double *__restrict__ sf,
                                         double *__restrict__ d_min_out, int *__restrict__ d_flags)
{
    const double m = dec_min(*min_slot);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        *next_slot = enc_min(DBL_MAX);
        *d_min_out = m;
        if (m < 0.0f) d_flags[1] = 1;
    }
    if (i < n) sf[i] = m / vol[i];
}

// the same with the minimum taken over the slots of all ranks (peer-mapped pointers / mailboxes); mp.on: every block
// first waits until all peers' minima of this visit have arrived (flags published by their visit prologues)
__global__ void step_factor_group_kernel(int n, const double *__restrict__ vol, MinSlots slots,
                                         unsigned long long *__restrict__ next_slot, double *__restrict__ sf,
                                         double *__restrict__ d_min_out, int *__restrict__ d_flags, const __grid_constant__ MinPush mp)
{
    if (mp.on) {
        if ((int)threadIdx.x < mp.n_peers) bounded_wait<false>(mp.src_flag[threadIdx.x], *mp.expected[threadIdx.x], mp.err_flag, mp.timeout_ns);
        __syncthreads();
    }
    unsigned long long u = ~0ull;
    for (int r = 0; r < slots.n; r++) {
        unsigned long long t = *reinterpret_cast<const volatile unsigned long long *>(slots.p[r]);
        if (t < u) u = t;
    }
    const double m = dec_min(u);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        *next_slot = enc_min(DBL_MAX);
        *d_min_out = m;
        if (m < 0.0f) d_flags[1] = 1;
    }
    if (i < n) sf[i] = m / vol[i];
}

// halo export: gather the rows of the exported owned nodes into a contiguous send buffer
__global__ void pack_rows_kernel(int n5, const int *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n5) dst[i] = src[(size_t)idx[i / 5] * 5 + i % 5];
}

// ---- p2p transport ---------------------------------------------------------------------------------
// gather the exported rows and store them straight into each destination rank's halo range (peer-mapped pointers)
__global__ void push_rows_kernel(int n5, const int *__restrict__ idx, const double *__restrict__ src, PushTable t)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n5) return;
    int row = i / 5, c = i - row * 5, k = 0;
    while (k + 1 < t.n_dst && row >= t.exp_ptr[k + 1]) k++;
    t.dst[k][(size_t)(row - t.exp_ptr[k]) * 5 + c] = src[(size_t)idx[row] * 5 + c];
}

// one warp: tell every destination that my rows of this exchange have landed (epoch counters live on the device so
// that the kernel can be replayed from a CUDA graph), then wait until every source has told me the same
__global__ void signal_wait_kernel(PushTable t)
{
    const int lane = threadIdx.x;
    if (lane < t.n_dst) {
        unsigned long long e = *t.sent[lane] + 1;
        *t.sent[lane] = e;
        __threadfence_system();
        st_release_sys(t.dst_flag[lane], e);
    }
    __syncwarp();
    if (lane < t.n_src) {
        unsigned long long e = *t.expected[lane] + 1;
        *t.expected[lane] = e;
        bounded_wait(t.src_flag[lane], e, t.err_flag, t.timeout_ns);
    }
}

// all-reduce(MIN) of min_dt by mailboxes: my encoded minimum goes into every peer's box, then the same flag handshake
__global__ void min_exchange_kernel(const unsigned long long *my_slot, MinTable t)
{
    const int lane = threadIdx.x;
    if (lane < t.n_peers) {
        *t.dst_box[lane] = *my_slot;
        unsigned long long e = *t.sent[lane] + 1;
        *t.sent[lane] = e;
        __threadfence_system();
        st_release_sys(t.dst_flag[lane], e);
    }
    __syncwarp();
    if (lane < t.n_peers) {
        unsigned long long e = *t.expected[lane] + 1;
        *t.expected[lane] = e;
        bounded_wait(t.src_flag[lane], e, t.err_flag, t.timeout_ns);
    }
}

// end of a multi-rank run: my deferred error flags go into every peer's status box (same mailbox handshake as the min_dt
// exchange), theirs are folded into mine
__global__ void status_exchange_kernel(int *flags, const unsigned long long *boxes, int n_ranks, MinTable t)
{
    const int lane = threadIdx.x;
    const unsigned long long mine = (flags[0] > 0 ? 1ull : 0ull) | (flags[1] ? 2ull : 0ull) | (flags[3] ? 4ull : 0ull);
    if (lane < t.n_peers) {
        *t.dst_box[lane] = mine;
        unsigned long long e = *t.sent[lane] + 1;
        *t.sent[lane] = e;
        __threadfence_system();
        st_release_sys(t.dst_flag[lane], e);
    }
    __syncwarp();
    if (lane < t.n_peers) {
        unsigned long long e = *t.expected[lane] + 1;
        *t.expected[lane] = e;
        bounded_wait(t.src_flag[lane], e, t.err_flag, t.timeout_ns);
    }
    __syncwarp();
    if (lane == 0)
        for (int q = 0; q < n_ranks; q++) {
            if (q == t.me) continue;
            const unsigned long long u = *reinterpret_cast<const volatile unsigned long long *>(boxes + q);
            if ((u & 1ull) && flags[0] == 0) flags[0] = 1;
            if (u & 2ull) flags[1] = 1;
            if (u & 4ull) flags[3] = 1;
        }
}

__global__ void reset_min_slots_kernel(int n, unsigned long long *slots)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slots[i] = enc_min(DBL_MAX);
}

// up_pre + up + up_post (mg.h:29-64) in one gather: a coarse node with children becomes their average, summed from
// 0.0 in ascending file order exactly as the three loops do; a childless coarse node is left untouched (x * 1.0 == x)
__global__ void restrict_fused_kernel(int n_coarse, const int *__restrict__ child_ptr, const int *__restrict__ child_idx,
                                      const double *__restrict__ var, double *__restrict__ var_above,
                                      int *__restrict__ count_above, const __grid_constant__ NodePush np)
{
    pdl_launch_dependents();
    int pushed = 0;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int j0 = 0, j1 = 0;
    if (p < n_coarse) { j0 = __ldg(child_ptr + p); j1 = __ldg(child_ptr + p + 1); }      // static index data
    pdl_wait();                                     // programmatic dependent launch (internal.h): nothing mutable is touched above
    if (np.on) node_wait(np);                       // the children's rows on other ranks (pushed by their last stage) are in
    if (p < n_coarse) {
        if (j0 != j1) {
            double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            for (int j = j0; j < j1; j++) {
                const double *u = var + (size_t)child_idx[j] * 5;
#pragma unroll
                for (int v = 0; v < 5; v++) acc[v] += u[v];
            }
            double avg = 1.0 / (double)(j1 - j0);
#pragma unroll
            for (int v = 0; v < 5; v++) { acc[v] = acc[v] * avg; var_above[(size_t)p * 5 + v] = acc[v]; }
            count_above[p] = j1 - j0;
            if (np.on) pushed = node_push(np, p, acc);       // only owned coarse nodes have children here
        }
        // (a childless owned coarse node keeps its value, Q8: the neighbours already hold it -- pushed by the last stage
        // of the level's previous visit into this same buffer)
    }
    if (np.on) node_publish(np, pushed);
}

// time_stepping_kernels.h:43-64 (only line :63 has an effect)
__global__ void step_factor_kernel(int n, const double *__restrict__ vol, const double *__restrict__ d_min,
                                   double *__restrict__ sf)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sf[i] = (*d_min) / vol[i];
}

// time_stepping_kernels.h:66-86
__global__ void time_step_kernel(int n, int rk, const double *__restrict__ sf, double *__restrict__ flux,
                                 const double *__restrict__ old, double *__restrict__ var)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * 5) return;
    double factor = sf[i / 5] / (double)(MGCFD_RK + 1 - rk);
    var[i] = old[i] + factor * flux[i];
    flux[i] = 0.0;
}

// validation.h:27-35
__global__ void residual_kernel(long long n5, const double *__restrict__ old, const double *__restrict__ var,
                                double *__restrict__ res)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n5) res[i] = var[i] - old[i];
}

// validation.h:37-44 (OP_INC into a global; summation order is not the sequential one)
__global__ void rms_kernel(long long n5, const double *__restrict__ res, double *__restrict__ d_rms)
{
    __shared__ double ssum[TPB / 32];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n5; i += (long long)gridDim.x * blockDim.x)
        s += res[i] * res[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ssum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < TPB / 32; w++) s += ssum[w];
        atomicAdd(d_rms, s);
    }
}

// validation.h:102-115
__global__ void bad_vals_kernel(long long n5, const double *__restrict__ var, int *__restrict__ d_count)
{
    int c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n5; i += (long long)gridDim.x * blockDim.x) {
        double v = var[i];
        if (isnan(v) || isinf(v)) c++;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(d_count, c);
}

// validation.h:46-100: identify_differences + count_non_zeros fused
__global__ void validate_kernel(long long n5, const double *__restrict__ test, const double *__restrict__ master,
                                int *__restrict__ d_count)
{
    int c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n5; i += (long long)gridDim.x * blockDim.x) {
        double tol = master[i] * 10.0e-8;
        if (tol < 0.0) tol *= -1.0;
        if (tol < 3.0e-19) tol = 3.0e-19;
        double d = test[i] - master[i];
        if (d < 0.0) d *= -1.0;
        if (d > tol) c++;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(d_count, c);
}

// mg.h:29-39 through node-->mg_node: only coarse nodes with a child are zeroed (all writers store 0)
__global__ void up_pre_kernel(int n_fine, const int *__restrict__ mg, double *__restrict__ var_above,
                              int *__restrict__ count_above)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_fine) return;
    int p = mg[i];
#pragma unroll
    for (int v = 0; v < 5; v++) var_above[(size_t)p * 5 + v] = 0.0;
    count_above[p] = 0;
}

// mg.h:41-52 as a gather: one thread per coarse node adds its children in ascending FILE order,
// which is the order OP2-seq applies the increments in
__global__ void up_gather_kernel(int n_coarse, const int *__restrict__ child_ptr, const int *__restrict__ child_idx,
                                 const double *__restrict__ var, double *__restrict__ var_above,
                                 int *__restrict__ count_above)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_coarse) return;
    int j0 = child_ptr[p], j1 = child_ptr[p + 1];
    if (j0 == j1) return;
    double acc[5];
#pragma unroll
    for (int v = 0; v < 5; v++) acc[v] = var_above[(size_t)p * 5 + v];
    for (int j = j0; j < j1; j++) {
        const double *u = var + (size_t)child_idx[j] * 5;
#pragma unroll
        for (int v = 0; v < 5; v++) acc[v] += u[v];
    }
#pragma unroll
    for (int v = 0; v < 5; v++) var_above[(size_t)p * 5 + v] = acc[v];
    count_above[p] += j1 - j0;
}

// mg.h:54-64
__global__ void up_post_kernel(int n_coarse, double *__restrict__ var, const int *__restrict__ count)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_coarse) return;
    int k = count[p];
    double avg = k == 0 ? 1.0 : 1.0 / (double)k;
#pragma unroll
    for (int v = 0; v < 5; v++) var[(size_t)p * 5 + v] *= avg;
}

// mg.h:66-88
__global__ void down_kernel(int n_fine, const int *__restrict__ mg, double *__restrict__ var,
                            const double *__restrict__ res, const double *__restrict__ xyz,
                            const double *__restrict__ res_above, const double *__restrict__ xyz_above,
                            const __grid_constant__ NodePush np)
{
    pdl_launch_dependents();
    int pushed = 0;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int p = 0;
    double dx = 0.0, dy = 0.0, dz = 0.0;
    if (i < n_fine) {                               // static: the map and the coordinates
        p = mg[i];
        dx = fabs(xyz[(size_t)i * 3] - xyz_above[(size_t)p * 3]);
        dy = fabs(xyz[(size_t)i * 3 + 1] - xyz_above[(size_t)p * 3 + 1]);
        dz = fabs(xyz[(size_t)i * 3 + 2] - xyz_above[(size_t)p * 3 + 2]);
    }
    pdl_wait();                                     // programmatic dependent launch (internal.h): nothing mutable is touched above
    if (np.on) node_wait(np);                       // the parents' residuals on other ranks (pushed by their last stage) are in
    if (i < n_fine) {
        double dm = sqrt(dx * dx + dy * dy + dz * dz);
        double *u = var + (size_t)i * 5;
        const double *r = res + (size_t)i * 5, *ra = res_above + (size_t)p * 5;
        double w[5];
        w[0] = u[0] - dm * (ra[0] - r[0]);
        w[1] = u[1] - dx * (ra[1] - r[1]);
        w[2] = u[2] - dy * (ra[2] - r[2]);
        w[3] = u[3] - dz * (ra[3] - r[3]);
        w[4] = u[4] - dm * (ra[4] - r[4]);
#pragma unroll
        for (int v = 0; v < 5; v++) u[v] = w[v];
        if (np.on) pushed = node_push(np, i, w);
    }
    if (np.on) node_publish(np, pushed);
}

}  // namespace

int k_copy(cudaStream_t s, int n, const double *var, double *old)
{
    if (n == 0) return 0;
    long long n5 = (long long)n * 5;
    copy_kernel<<<blocks_for(n5), TPB, 0, s>>>(n5, var, old);
    return 1;
}
int k_fill(cudaStream_t s, long long n, double *a, double v)
{
    if (n == 0) return 0;
    fill_kernel<<<blocks_for(n), TPB, 0, s>>>(n, a, v);
    return 1;
}
int k_permute_rows(cudaStream_t s, int n, int dim, const double *src, const int *perm, double *dst, bool to_internal)
{
    if (n == 0) return 0;
    long long total = (long long)n * dim;
    permute_rows_kernel<<<blocks_for(total), TPB, 0, s>>>(total, dim, src, perm, dst, to_internal);
    return 1;
}
int k_init_vars(cudaStream_t s, int n, double *var, const DevConsts &c)
{
    if (n == 0) return 0;
    init_vars_kernel<<<blocks_for(n), TPB, 0, s>>>(n, var, c);
    return 1;
}
int k_calculate_dt(cudaStream_t s, int n, const double *var, const double *cbrt_vol, double *sf)
{
    if (n == 0) return 0;
    calculate_dt_kernel<<<blocks_for(n), TPB, 0, s>>>(n, var, cbrt_vol, sf);
    return 1;
}
int k_min_dt(cudaStream_t s, int n, const double *sf, double *d_min, int *d_flags)
{
    // d_min holds the start value as a double; encode, reduce, decode in place
    unsigned long long *u = reinterpret_cast<unsigned long long *>(d_min);
    encode_min_kernel<<<1, 1, 0, s>>>(u);
    int launches = 2;
    if (n > 0) {
        int grid = blocks_for(n);
        if (grid > 148 * 8) grid = 148 * 8;
        min_dt_kernel<<<grid, TPB, 0, s>>>(n, sf, u);
        launches++;
    }
    finish_min_kernel<<<1, 1, 0, s>>>(u, d_flags);
    return launches;
}
int k_visit_begin(cudaStream_t s, int n, const double *var, const double *cbrt_vol, double *old, double *dt,
                  unsigned long long *min_slot, const MinPush *mp, double *zero_me)
{
    MinPush none;
    memset(&none, 0, sizeof(none));
    if (n == 0 && !(mp && mp->on)) return 0;
    launch_k(visit_begin_kernel, dim3(n > 0 ? blocks_for(n) : 1), dim3(TPB), 0, s, n, var, cbrt_vol, old, dt, min_slot, mp ? *mp : none, zero_me);
    return 1;
}
int k_step_factor_fused(cudaStream_t s, int n, const double *vol, unsigned long long *min_slot, unsigned long long *next_slot,
                        double *sf, double *d_min_out, int *d_flags)
{
    step_factor_fused_kernel<<<n > 0 ? blocks_for(n) : 1, TPB, 0, s>>>(n, vol, min_slot, next_slot, sf, d_min_out, d_flags);
    return 1;
}
int k_step_factor_group(cudaStream_t s, int n, const double *vol, MinSlots slots, unsigned long long *next_slot, double *sf,
                        double *d_min_out, int *d_flags, const MinPush *mp)
{
    MinPush none;
    memset(&none, 0, sizeof(none));
    step_factor_group_kernel<<<n > 0 ? blocks_for(n) : 1, TPB, 0, s>>>(n, vol, slots, next_slot, sf, d_min_out, d_flags, mp ? *mp : none);
    return 1;
}
int k_pack_rows(cudaStream_t s, int n, const int *idx, const double *src, double *dst)
{
    if (n == 0) return 0;
    pack_rows_kernel<<<blocks_for((long long)n * 5), TPB, 0, s>>>(n * 5, idx, src, dst);
    return 1;
}
int k_push_rows(cudaStream_t s, int n_rows, const int *idx, const double *src, const PushTable &t)
{
    if (n_rows == 0) return 0;
    push_rows_kernel<<<blocks_for((long long)n_rows * 5), TPB, 0, s>>>(n_rows * 5, idx, src, t);
    return 1;
}
int k_signal_wait(cudaStream_t s, const PushTable &t)
{
    if (t.n_dst == 0 && t.n_src == 0) return 0;
    signal_wait_kernel<<<1, 32, 0, s>>>(t);
    return 1;
}
int k_min_exchange(cudaStream_t s, const unsigned long long *my_slot, const MinTable &t)
{
    min_exchange_kernel<<<1, 32, 0, s>>>(my_slot, t);
    return 1;
}
int k_status_exchange(cudaStream_t s, int *flags, const unsigned long long *boxes, int n_ranks, const MinTable &t)
{
    status_exchange_kernel<<<1, 32, 0, s>>>(flags, boxes, n_ranks, t);
    return 1;
}
int k_reset_min_slots(cudaStream_t s, int n, unsigned long long *slots)
{
    reset_min_slots_kernel<<<blocks_for(n), TPB, 0, s>>>(n, slots);
    return 1;
}
int k_restrict_fused(cudaStream_t s, int n_coarse, const int *child_ptr, const int *child_idx, const double *var,
                     double *var_above, int *count_above, const NodePush *np)
{
    NodePush none;
    memset(&none, 0, sizeof(none));
    if (n_coarse == 0 && !(np && np->on)) return 0;
    launch_k(restrict_fused_kernel, dim3(n_coarse > 0 ? blocks_for(n_coarse) : 1), dim3(TPB), 0, s, n_coarse, child_ptr, child_idx, var, var_above,
             count_above, np ? *np : none);
    return 1;
}
int k_step_factor(cudaStream_t s, int n, const double *vol, const double *d_min, double *sf)
{
    if (n == 0) return 0;
    step_factor_kernel<<<blocks_for(n), TPB, 0, s>>>(n, vol, d_min, sf);
    return 1;
}
int k_time_step(cudaStream_t s, int n, int rk, const double *sf, double *flux, const double *old, double *var)
{
    if (n == 0) return 0;
    time_step_kernel<<<blocks_for((long long)n * 5), TPB, 0, s>>>(n, rk, sf, flux, old, var);
    return 1;
}
int k_residual(cudaStream_t s, int n, const double *old, const double *var, double *res)
{
    if (n == 0) return 0;
    long long n5 = (long long)n * 5;
    residual_kernel<<<blocks_for(n5), TPB, 0, s>>>(n5, old, var, res);
    return 1;
}
static int reduce_grid(long long n)
{
    int grid = blocks_for(n);
    return grid > 148 * 8 ? 148 * 8 : grid;
}
int k_rms(cudaStream_t s, int n, const double *res, double *d_rms)
{
    if (n == 0) return 0;
    long long n5 = (long long)n * 5;
    rms_kernel<<<reduce_grid(n5), TPB, 0, s>>>(n5, res, d_rms);
    return 1;
}
int k_bad_vals(cudaStream_t s, int n, const double *var, int *d_count)
{
    if (n == 0) return 0;
    long long n5 = (long long)n * 5;
    bad_vals_kernel<<<reduce_grid(n5), TPB, 0, s>>>(n5, var, d_count);
    return 1;
}
int k_validate(cudaStream_t s, int n, const double *test, const double *master, int *d_count)
{
    if (n == 0) return 0;
    long long n5 = (long long)n * 5;
    validate_kernel<<<reduce_grid(n5), TPB, 0, s>>>(n5, test, master, d_count);
    return 1;
}
int k_up_pre(cudaStream_t s, int n_fine, const int *mg, double *var_above, int *count_above)
{
    if (n_fine == 0) return 0;
    up_pre_kernel<<<blocks_for(n_fine), TPB, 0, s>>>(n_fine, mg, var_above, count_above);
    return 1;
}
int k_up(cudaStream_t s, int n_coarse, const int *child_ptr, const int *child_idx, const double *var,
         double *var_above, int *count_above)
{
    if (n_coarse == 0) return 0;
    up_gather_kernel<<<blocks_for(n_coarse), TPB, 0, s>>>(n_coarse, child_ptr, child_idx, var, var_above, count_above);
    return 1;
}
int k_up_post(cudaStream_t s, int n_coarse, double *var, const int *count)
{
    if (n_coarse == 0) return 0;
    up_post_kernel<<<blocks_for(n_coarse), TPB, 0, s>>>(n_coarse, var, count);
    return 1;
}
int k_down(cudaStream_t s, int n_fine, const int *mg, double *var, const double *res, const double *coords,
           const double *res_above, const double *coords_above, const NodePush *np)
{
    NodePush none;
    memset(&none, 0, sizeof(none));
    if (n_fine == 0 && !(np && np->on)) return 0;
    launch_k(down_kernel, dim3(n_fine > 0 ? blocks_for(n_fine) : 1), dim3(TPB), 0, s, n_fine, mg, var, res, coords, res_above, coords_above, np ? *np : none);
    return 1;
}

// ---- flux dispatch: exact build lives here, fast build in flux_fast.cu
int k_bnd_flux(cudaStream_t s, int n_unique, const int *bu_node, const int *bu_ptr, const int *b_group,
               const double *b_wt, const double *var, double *flux, const DevConsts &c, bool exact_mode)
{
    return exact_mode ? exact::launch_bnd(s, n_unique, bu_node, bu_ptr, b_group, b_wt, var, flux, c)
                      : fast_bnd_flux(s, n_unique, bu_node, bu_ptr, b_group, b_wt, var, flux, c);
}
int flux_atomic(cudaStream_t s, const FluxArgs &a, const AtomicPlanDev &p, bool exact_mode)
{
    return exact_mode ? exact::launch_atomic(s, a, p) : fast_flux_atomic(s, a, p);
}
int flux_colour(cudaStream_t s, const FluxArgs &a, const ColourPlanDev &p, const ColourPlanHost &h, bool exact_mode)
{
    return exact_mode ? exact::launch_colour(s, a, p, h) : fast_flux_colour(s, a, p, h);
}
int flux_owner(cudaStream_t s, const FluxArgs &a, const OwnerPlanDev &p, const OwnerPlanHost &h, bool exact_mode)
{
    return exact_mode ? exact::launch_owner(s, a, p, h) : fast_flux_owner(s, a, p, h);
}
int flux_gather(cudaStream_t s, const FluxArgs &a, const GatherPlanDev &p, int n_chunks, int max_loc, bool exact_mode)
{
    return exact_mode ? exact::launch_gather(s, a, p, n_chunks, max_loc) : fast_flux_gather(s, a, p, n_chunks, max_loc);
}
size_t flux_gather_smem_bytes(int max_loc, bool exact_mode)
{
    return exact_mode ? exact::gather_smem(max_loc, false) : fast_gather_smem(max_loc);
}
bool flux_owner_uses_stage2(const OwnerPlanDev &p, const OwnerPlanHost &h, bool exact_mode)
{
    return !exact_mode && fast_owner_uses_stage2(p, h);
}
std::string flux_configure()
{
    std::string e = exact::configure();
    if (!e.empty()) return e;
    return fast_configure();
}
size_t flux_owner_smem_bytes(int max_loc, int max_edges, int max_blob, bool exact_mode)
{
    size_t stream = exact::owner_smem(max_loc, max_edges, max_blob, true);
    size_t body = (exact_mode ? exact::owner_smem(max_loc, max_edges, max_blob, false) : fast_owner_smem(max_loc, max_edges, max_blob)) +
                  (size_t)256 * 6 * sizeof(double);      // + the fused stage's old_variables / step_factor tiles
    return stream > body ? stream : body;
}
size_t flux_colour_smem_bytes(int max_nodes, bool exact_mode)
{
    return exact_mode ? exact::colour_smem(max_nodes, false) : fast_colour_smem(max_nodes);
}

}  // namespace mgcfd
