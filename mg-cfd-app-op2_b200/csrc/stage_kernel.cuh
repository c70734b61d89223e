// stage_kernel.cuh -- the fused Runge-Kutta stage, second generation ("stage2").  Fast build only; included by
// flux_kernels.cuh inside namespace mgcfd::fast.  One launch = compute_flux_edge_kernel (flux.h:41-208) +
// compute_bnd_node_flux_kernel (flux.h:14-39) + time_step_kernel (time_stepping_kernels.h:66-86) and, after the last
// stage, residual_kernel / calc_rms_kernel / count_bad_vals (validation.h:27-44,102-115) over one level.
//
// Same owner-compute chunks, same blobs and the same five phases as flux_owner_kernel<FUSE, REGEPI>, rewritten around
// what round 1's ncu captures showed (profiles/README.md section 5: the kernel is bound by issued instructions and
// resident warps, not by HBM):
//   * the shared-memory layout is a compile-time constant (plane strides MAXE / MAXL), so every plane address is
//     `register + immediate`; the blob's weight planes arrive by one bulk async copy each;
//   * a chunk's descriptor and halo ids are ONE fixed-stride record (xtab): the ids are requested together with the
//     descriptor, so staging is two dependent memory latencies (record -> {bulk copies, halo gather}), not three;
//   * halo rows are gathered by groups of five lanes (lane = component), six rows per warp pass: no divisions, one id
//     per row, the requests of all passes in flight at once;
//   * derived quantities (p, |v|+c, 1/rho) by Newton iterations from the hardware reciprocal / reciprocal-square-root
//     seeds (rcp.approx.f64, rsqrt.approx.f64) instead of IEEE division and square roots (|rel. error| ~1e-16);
//   * node phase: incidence sign applied by fma(+-1, F, acc), byte offsets straight from the csr entry, the update
//     factor by a multiplication with the host-computed 1/(RK+1-rk), optionally without old_variables / step_factor
//     tiles in shared memory (TILES = false: 3 KB less per CTA -> one more resident CTA per SM).
#pragma once

template <int MAXE, int MAXL, bool TILES>
struct Stage2Layout {
    static constexpr int C = 64;                              // owned nodes per chunk (at most)
    static constexpr int BAR = 0;
    static constexpr int W0 = 16, W1 = W0 + MAXE * 8, W2 = W1 + MAXE * 8, G = W2 + MAXE * 8, FX = G + MAXE * 8;
    static constexpr int RAW = FX + MAXE * 8;                 // conserved variables, AoS [MAXL][5]
    static constexpr int DER = RAW + MAXL * 40;               // p | |v|+c | 1/rho planes [MAXL]
    static constexpr int OLD = DER + MAXL * 24;               // TILES: old_variables [C][5], step_factors [C]
    static constexpr int SF = OLD + (TILES ? C * 40 : 0);
    static constexpr int TAIL = SF + (TILES ? C * 8 : 0);     // lab | rowptr | csr (the blob's tail without its boundary entries)
    static_assert(MAXE % 2 == 0 && MAXL % 2 == 0, "16-byte aligned bulk-copy targets");
};

// 1/x and 1/sqrt(x) from the hardware seeds (about 20 bits) by two Newton steps each
__device__ __forceinline__ double rcp_nr(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double sqrt_nr(double x)
{
    const double xs = x + 1.0e-290;                           // x == 0 would give 0 * inf; any normal x is unchanged by the add
    double y;                                                 // negative or NaN arguments come back as NaN (counted as bad values)
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));
    const double h = 0.5 * xs;
    double e = fma(-h, y * y, 0.5);
    y = fma(y, e, y);
    e = fma(-h, y * y, 0.5);
    y = fma(y, e, y);
    return xs * y;
}
__device__ __forceinline__ void derive_nr(const double u[5], double &p, double &s, double &rinv)
{
    rinv = rcp_nr(u[0]);
    const double mm = u[1] * u[1] + u[2] * u[2] + u[3] * u[3];
    const double q2 = mm * rinv * rinv;
    p = (GAMMA - 1.0) * (u[4] - 0.5 * mm * rinv);
    s = sqrt_nr(q2) + sqrt_nr(GAMMA * p * rinv);
}

template <int MAXE, int MAXL, bool TILES, int MINB>
__global__ void __launch_bounds__(128, MINB)
rk_stage2_kernel(const int *__restrict__ xtab, int xs, int hs, int rec0, int pf_chunk,
                 const unsigned char *__restrict__ blob, const double *__restrict__ var, const __grid_constant__ RkStageArgs rk)
{
    using L = Stage2Layout<MAXE, MAXL, TILES>;
    extern __shared__ __align__(128) unsigned char sm2[];
    unsigned char *const sm = sm2;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + L::BAR);
    double *raw = reinterpret_cast<double *>(sm + L::RAW);
    double *der = reinterpret_cast<double *>(sm + L::DER);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = rec0 + (int)blockIdx.x;               // the records are stored in launch order
    pdl_launch_dependents();                                // programmatic dependent launch (internal.h); a no-op otherwise
    const int *rec = xtab + (size_t)chunk * xs;

    // ---- 1. the chunk's record: halo ids of this thread's rows and the descriptor, independent loads (one round trip)
    constexpr int KH = 7;                                   // 7 passes x 24 rows = 168 halo rows without a dependent id load
    const int rsub = (lane * 13) >> 6;                      // lane / 5
    const int comp = lane - 5 * rsub;
    const int row0 = warp * 6 + rsub;
    const bool lane_ok = lane < 30;
    int hg[KH];
#pragma unroll
    for (int k = 0; k < KH; k++) {
        const int r = row0 + 24 * k;
        hg[k] = (lane_ok && r < hs) ? __ldg(rec + 12 + r) : -1;
    }
    const int4 q0 = __ldg(reinterpret_cast<const int4 *>(rec));
    const int4 q1 = __ldg(reinterpret_cast<const int4 *>(rec) + 1);
    const int4 q2 = __ldg(reinterpret_cast<const int4 *>(rec) + 2);
    const int node0 = q0.x, n_own = q0.y, n_halo = q0.z;
    const int n_edges = q1.x, e_pad = q1.y, n_inc = q1.z;
    const long long blob_off = (long long)(unsigned)q2.x | ((long long)q2.y << 32);
    const int has_bnd = q2.z, bnd_off = q2.w;
    // multi-GPU, fused push: word 3 of the record is the chunk's base into the export row pointers (-1: no exported node);
    // such a chunk -- the only kind that reads rank-halo rows -- first waits for its sources
    const int xb = rk.push_on ? q0.w : -1;
    // experiment (MGCFD_STAGE2_PF=distance): thread 64 fetches the descriptor of the chunk `distance` launches ahead and
    // later prefetches that chunk's contiguous inputs into L2, so that its CTA finds them there
    int4 p0 = make_int4(0, 0, 0, 0), p1 = p0, p2 = p0;
    const bool pf = pf_chunk > 0 && tid == 64 && chunk + pf_chunk < (int)gridDim.x;
    if (pf) {
        const int4 *r2 = reinterpret_cast<const int4 *>(xtab + (size_t)(chunk + pf_chunk) * xs);
        p0 = __ldg(r2); p1 = __ldg(r2 + 1); p2 = __ldg(r2 + 2);
    }
    const uint32_t pb = (uint32_t)e_pad * 8u;               // bytes of one weight plane in the blob
    const uint32_t own_b = ((uint32_t)n_own * 40u) & ~15u, sf_b = ((uint32_t)n_own * 8u) & ~15u;

    // ---- 2. bulk async copies (TMA 1-D): four weight planes, the blob's tail, the owned nodes' variables
    //         (+ old_variables and step factors: they land while the fluxes are computed).  The blob is static plan data:
    //         under a programmatic dependent launch it is requested while the previous kernel is still draining; everything
    //         the previous kernels (or the peers) write is touched only behind pdl_wait()
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)bnd_off + own_b + (TILES ? own_b + sf_b : 0u));
        const unsigned char *src = blob + blob_off;
        if (pb) {
            bulk_g2s(sm + L::W0, src, pb, bar);
            bulk_g2s(sm + L::W1, src + pb, pb, bar);
            bulk_g2s(sm + L::W2, src + 2 * pb, pb, bar);
            bulk_g2s(sm + L::G, src + 3 * pb, pb, bar);
        }
        bulk_g2s(sm + L::TAIL, src + 4 * pb, (uint32_t)bnd_off - 4 * pb, bar);      // lab | rowptr | csr (not the boundary entries)
    }
    pdl_wait();
    if (xb >= 0) push_wait_sources(rk.push, tid);
    if (tid == 0) {
        if (own_b) bulk_g2s(raw, var + (size_t)node0 * 5, own_b, bar);
        if (TILES) {
            if (own_b) bulk_g2s(sm + L::OLD, rk.old + (size_t)node0 * 5, own_b, bar);
            if (sf_b) bulk_g2s(sm + L::SF, rk.sf + node0, sf_b, bar);
        }
    }
    // ---- 3. halo rows: 8-byte async copies, five lanes per row
    {
        double *hraw = raw + n_own * 5 + comp;
        const double *vsrc = var + comp;
#pragma unroll
        for (int k = 0; k < KH; k++)
            if (hg[k] >= 0) cp_async8(hraw + (row0 + 24 * k) * 5, vsrc + (size_t)hg[k] * 5);
        if (lane_ok)
            for (int r = row0 + 24 * KH; r < n_halo; r += 24) cp_async8(hraw + r * 5, vsrc + (size_t)__ldg(rec + 12 + r) * 5);
        if (tid == 32 && (n_own & 1)) raw[n_own * 5 - 1] = __ldg(var + (size_t)(node0 + n_own) * 5 - 1);   // odd last chunk
    }
    cp_async_wait_all();
    __syncthreads();          // halo tile complete; mbarrier initialisation visible
    mbar_wait(bar, 0);        // bulk copies landed

    // ---- 4. derived quantities once per staged node
    const int nloc = n_own + n_halo;
    for (int i = tid; i < nloc; i += 128) {
        double u[5];
#pragma unroll
        for (int v = 0; v < 5; v++) u[v] = raw[i * 5 + v];
        double p, s, rinv;
        derive_nr(u, p, s, rinv);
        der[i] = p; der[MAXL + i] = s; der[2 * MAXL + i] = rinv;
    }
    __syncthreads();

    if (pf) {
        const long long boff = (long long)(unsigned)p2.x | ((long long)p2.y << 32);
        const uint32_t ob = ((uint32_t)p0.y * 40u) & ~15u;
        bulk_prefetch_l2(blob + boff, (uint32_t)p2.w);                              // weights | lab | rowptr | csr
        bulk_prefetch_l2(xtab + (size_t)(chunk + pf_chunk) * xs, (uint32_t)(48 + ((p0.z * 4 + 15) & ~15)));
        if (ob) {
            bulk_prefetch_l2(var + (size_t)p0.x * 5, ob);
            bulk_prefetch_l2(rk.old + (size_t)p0.x * 5, ob);
        }
    }
    // ---- 5. one thread per edge: the edge's weights are replaced in place by its flux vector
    const uint32_t *lab = reinterpret_cast<const uint32_t *>(sm + L::TAIL);
    {
        double *w0 = reinterpret_cast<double *>(sm + L::W0);
        for (int e = tid; e < n_edges; e += 128) {
            const uint32_t l = lab[e];
            const int la = l & 0xffff, lb = l >> 16;
            double a[8], b[8], F[5];
#pragma unroll
            for (int v = 0; v < 5; v++) { a[v] = raw[la * 5 + v]; b[v] = raw[lb * 5 + v]; }
#pragma unroll
            for (int v = 0; v < 3; v++) { a[5 + v] = der[v * MAXL + la]; b[5 + v] = der[v * MAXL + lb]; }
            edge_flux(a, b, w0[e], w0[MAXE + e], w0[2 * MAXE + e], w0[3 * MAXE + e], F);
#pragma unroll
            for (int v = 0; v < 5; v++) w0[v * MAXE + e] = F[v];      // slot e is private to this thread
        }
    }
    // first stage of a visit with the step factor folded in (compute_step_factor_kernel, time_stepping_kernels.h:43-64):
    // warp 0 takes the minimum over the reduced slots -- multi-rank: once the peers' minima have arrived, which the
    // chunk's staging and edge phase have given them time for -- and leaves it in shared memory for the node phase
    if (!TILES && rk.fold.on && warp == 0) {
        if (lane < rk.fold.n_wait) bounded_wait<false>(rk.fold.wait_flag[lane], *rk.fold.wait_expected[lane], rk.d_bad + 3, rk.fold.timeout_ns);
        __syncwarp();
        unsigned long long u = ~0ull;
        if (lane < rk.fold.n_slots) u = *reinterpret_cast<const volatile unsigned long long *>(rk.fold.slot[lane]);
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, u, off);
            if (t < u) u = t;
        }
        if (lane == 0) {
            const double m = dec_min(u);
            *reinterpret_cast<double *>(sm + L::BAR + 8) = m;
            if (blockIdx.x == 0) {
                *rk.fold.next_slot = enc_min(1.7976931348623157e308);       // re-arm the level's other slot (DBL_MAX)
                *rk.fold.d_min_out = m;
                if (m < 0.0f) rk.fold.d_flags[1] = 1;                       // euler3d.cpp:480, checked at the end of the run
            }
        }
    }
    __syncthreads();

    // ---- 6. node phase: two threads per owned node (even / odd incidences), boundary entries, update, stores
    const uint16_t *rowptr = reinterpret_cast<const uint16_t *>(lab + e_pad);
    const uint16_t *csr = rowptr + (((n_own + 1) + 7) & ~7);
    // thread slot -> node: the chunk's owned nodes by descending degree (u8 table behind the csr), so that the 16 nodes of a
    // warp have about the same number of incidences and no outlier sets the warp's trip count
    const unsigned char *order = reinterpret_cast<const unsigned char *>(csr) + (((uint32_t)n_inc * 2u + 15u) & ~15u);
    const int part = tid & 1;
    const bool active = (tid >> 1) < n_own;
    const int n = active ? order[tid >> 1] : 0;
    const size_t g0 = (size_t)node0 * 5;
    // boundary chunks: the node's boundary entries (straight from the level's arrays, sorted by owned node) start the
    // sum of the node's first thread -- here, where nothing else is live, the divisions of bnd_apply cost no spills
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (has_bnd && active && part == 0) {
        const int b0 = __ldg(rk.bnd_ptr + node0 + n), b1 = __ldg(rk.bnd_ptr + node0 + n + 1);
        if (b1 > b0) {
            double u[5];
#pragma unroll
            for (int v = 0; v < 5; v++) u[v] = raw[n * 5 + v];
            bnd_apply(u, acc, b0, b1, rk.b_group, rk.b_wt, rk.c);
        }
    }
    // this thread finishes components part, part+2, part+4 of its node
    double o[3] = {0.0, 0.0, 0.0}, sfn = 0.0;
    int xj0 = 0, xj1 = 0;                                   // the node's export entries (requested now, used after the sums)
    if (xb >= 0 && active) { xj0 = __ldg(rk.push.xp_ptr + xb + n); xj1 = __ldg(rk.push.xp_ptr + xb + n + 1); }
    if (active) {
        if (TILES) {
            const double *told = reinterpret_cast<const double *>(sm + L::OLD), *tsf = reinterpret_cast<const double *>(sm + L::SF);
            const bool tail = (n_own & 1) && n == n_own - 1;          // odd last chunk: outside the bulk-copied tiles
            sfn = tail ? rk.sf[node0 + n] : tsf[n];
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (part + 2 * k < 5) o[k] = (tail && n * 5 + part + 2 * k >= (int)(own_b >> 3)) ? rk.old[g0 + n * 5 + part + 2 * k] : told[n * 5 + part + 2 * k];
        } else {
            sfn = rk.fold.on ? __ldg(rk.fold.vol + node0 + n) : __ldg(rk.sf + node0 + n);      // fold: the volume for now
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (part + 2 * k < 5) o[k] = __ldg(rk.old + g0 + n * 5 + part + 2 * k);
        }
    }
    {
        const int j1 = active ? rowptr[n + 1] : 0;
        const unsigned char *pl = sm + L::W0;
        for (int jj = active ? rowptr[n] + part : 0; jj < j1; jj += 2) {
            const uint32_t c = csr[jj];
            const double sg = __hiloint2double((int)(0x3ff00000u | ((c & 0x8000u) << 16)), 0);     // +1.0 / -1.0 (end b)
            const unsigned char *pe = pl + ((c & 0x7fffu) << 3);
#pragma unroll
            for (int v = 0; v < 5; v++) acc[v] = fma(sg, *reinterpret_cast<const double *>(pe + v * MAXE * 8), acc[v]);
        }
    }
#pragma unroll
    for (int v = 0; v < 5; v++) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], 1);       // both threads hold the node's sum
    double sq = 0.0;
    int bad = 0;
    if (active) {
        if (!TILES && rk.fold.on) {
            sfn = *reinterpret_cast<const double *>(sm + L::BAR + 8) / sfn;          // min_dt / volume
            if (part == 0) rk.fold.sf_out[node0 + n] = sfn;
        }
        const double factor = sfn * rk.inv_denom;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int v = part + 2 * k;
            if (v < 5) {
                const double f = (k == 0) ? (part ? acc[1] : acc[0]) : (k == 1) ? (part ? acc[3] : acc[2]) : acc[4];
                const double vn = fma(factor, f, o[k]);
                rk.var_out[g0 + n * 5 + v] = vn;
                const double r = vn - o[k];
                if (rk.last) {
                    rk.res[g0 + n * 5 + v] = r;
                    sq = fma(r, r, sq);
                    bad += (isnan(vn) || isinf(vn)) ? 1 : 0;
                }
                if (xb >= 0) push_component(rk.push, xj0, xj1, v, vn, r, rk.last != 0);
            }
        }
    }
    if (xb >= 0) push_publish(rk.push, tid);
    if (rk.last && rk.d_rms) {
        for (int off = 16; off > 0; off >>= 1) {
            sq += __shfl_xor_sync(0xffffffffu, sq, off);
            bad += __shfl_xor_sync(0xffffffffu, bad, off);
        }
        if (lane == 0) {
            atomicAdd(rk.d_rms, sq);
            if (bad) atomicAdd(rk.d_bad, bad);
        }
    }
}

template <int MAXE, int MAXL, bool TILES, int MINB>
inline void stage2_launch_one(cudaStream_t s, int grid, size_t tail, const OwnerPlanDev &p, const FluxArgs &a, const RkStageArgs &ra)
{
    const size_t smem = Stage2Layout<MAXE, MAXL, TILES>::TAIL + tail;     // dynamic shared memory opt-in: configure()
    const char *pf = getenv("MGCFD_STAGE2_PF");
    launch_k(rk_stage2_kernel<MAXE, MAXL, TILES, MINB>, dim3(grid), dim3(128), smem, s, p.xtab, p.xs, p.hs, a.chunk_list ? a.list_offset : 0,
             pf ? atoi(pf) : 0, p.blob, a.var, ra);
}

// whether a fused stage on this plan runs the stage2 kernel: the plan fits a compiled configuration and no experiment knob
// selects another kernel (MGCFD_STAGE2=0, MGCFD_OWNER_PIPE>0, MGCFD_OWNER_THREADS=256)
inline bool stage2_applies(const OwnerPlanDev &p, const OwnerPlanHost &h)
{
    if (!p.xtab || h.max_own > 64 || h.max_loc > 240 || h.max_edges > 400) return false;
    const char *off = getenv("MGCFD_STAGE2"), *pipe = getenv("MGCFD_OWNER_PIPE"), *thr = getenv("MGCFD_OWNER_THREADS");
    if (off && atoi(off) == 0) return false;
    if ((pipe ? atoi(pipe) : MGCFD_OWNER_PIPE_DEFAULT) > 0) return false;
    if (thr && atoi(thr) == 256) return false;
    return true;
}

// returns 1 when the launch was made, 0 when the plan does not fit a compiled configuration (caller falls back)
inline int launch_stage2(cudaStream_t s, const FluxArgs &a, const OwnerPlanDev &p, const OwnerPlanHost &h, int grid)
{
    if (!a.rk || !stage2_applies(p, h)) return 0;
    RkStageArgs ra = *a.rk;
    ra.max_own = h.max_own;
    ra.inv_denom = 1.0 / (double)(MGCFD_RK + 1 - ra.rk);
    const char *tl = getenv("MGCFD_STAGE2_TILES"), *mb = getenv("MGCFD_STAGE2_MINB");
    const int tiles = (tl && !ra.fold.on) ? atoi(tl) : 0;        // the folded step factor lives in the tile-less instances
    const int minb = mb ? atoi(mb) : 7;
    const size_t tail = (size_t)h.dev_max_tail;
    if (h.max_edges <= 336) {
        if (tiles) stage2_launch_one<336, 240, true, 6>(s, grid, tail, p, a, ra);
        else if (minb >= 7) stage2_launch_one<336, 240, false, 7>(s, grid, tail, p, a, ra);
        else stage2_launch_one<336, 240, false, 6>(s, grid, tail, p, a, ra);
    } else {
        // chunks of dense levels (4-5 edges per node fill the 6C edge cap, which is soft by one node): 6 CTAs per SM
        if (tiles) stage2_launch_one<400, 240, true, 6>(s, grid, tail, p, a, ra);
        else stage2_launch_one<400, 240, false, 6>(s, grid, tail, p, a, ra);
    }
    return 1;
}
