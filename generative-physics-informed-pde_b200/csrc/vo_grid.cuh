// Structured-grid virtual-observable residual kernel (sm_100a).  Included by vo.cu.
//
// When the fine mesh is the reference's own (pixels of two P1 triangles on an nx x ny grid,
// factories/model.py:130-133; Dirichlet data on the left/right edges, LinearEllipticFactories.py:173-179,
// 239-281) and the conductivity comes per pixel (bottleneck/utils.py:41-98), K_fom(a) is a 5-point
// operator: the coupling of two horizontally adjacent nodes is chs * (a[pixel below] + a[pixel above]),
// of two vertically adjacent ones cvs * (a[pixel left] + a[pixel right]); the hypotenuse couplings of
// right-angled P1 triangles vanish.  gpde_vo_plan_create verifies this against the element data it is
// given and otherwise leaves the generic kernels (vo_fused.cuh / version 1) in charge.
//
//   r[b,:] = V^T (K_fom(a_b) u~_b - f)_free        (VirtualObservables.py:61-69, 662, 990)
//
// Work decomposition (FP64 pipe and HBM are co-limiting at m = 25):
//   * a CTA of 16 warps owns S = 8*groups samples and marches over the node rows bottom to top;
//   * R node rows of y, R pixel rows of a (all S samples) and R rows of the fragment-packed V form one
//     pipeline stage, brought into shared memory NS-1 stages ahead: a / y rows by 16-byte cp.async shared
//     out over all warps, the packed V rows by one bulk copy (cp.async.bulk); both complete on the
//     stage's mbarrier.  Every input byte crosses HBM once, nothing waits on a global load;
//   * warp (group, strip): 8 samples x 16 node columns; lane (s = lane/4, k = lane%4) owns the 4
//     columns 16*strip + 4k .. +3 of sample s and keeps the row below (u, conductivities, vertical
//     fluxes) in registers, so a step reads 6 + 5 doubles from shared memory for 4 nodes;
//   * flux form: S_i = rh (Fh_right - Fh_left) + (Fv_up - Fv_down), rho_i = cvs * S_i - f_i;
//   * the 4 values a lane produces ARE its A fragments of four mma.sync.m8n8k4.f64 k-steps
//     (M = 8 samples, K = 4 lanes' columns, N = 8 columns of V); the B fragments come from the packed
//     V row with conflict-free 8-byte loads; accumulators stay in registers for the whole pass;
//   * exp() of the log-field: 2^(k/16) table in shared memory + degree-6 polynomial (11 FP64 ops).
// Algorithmic HBM bytes per sample: 8 * (n_pixels + d + n_bc + m)   (SURVEY.md 8d).
#pragma once

namespace gpde {

struct GridDev {
    int ok;
    int nx, ny;          // pixels per row, pixel rows; nodes are (nx+1) x (ny+1)
    int ncol;            // free node columns per row (nx - 1)
    int cols;            // node columns per lane (4 or 8); a strip = 4 * cols columns
    int nstrips;         // strips, rounded up to a power of two (<= 16)
    int groups;          // sample groups (8 samples each) per CTA = warps / nstrips
    long long in0, sy;   // conductivity entry of pixel (cx, cy) = in0 + cy * sy + cx
    double rh, scale;    // rh = chs / cvs, scale = cvs
    int has_load;
    const double *f_over;   // [d] f_i / cvs
    int a_stride, y_stride;   // doubles per sample inside a stage
    int a_off, y_off, v_off;  // byte offsets inside a stage
};

constexpr int kGridWarpsMax = 16;   // warps per CTA: 16 (one CTA per SM) or 8 (two CTAs per SM)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16_u32(unsigned smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)   // (a suspend-time hint made wake-ups ~10x slower on B200: measured, not used)
            : "memory");
        if (!done) __nanosleep(40);
    } while (!done);
}
// global -> shared bulk copy (TMA engine, 1-D); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__constant__ double kExp16Tab[16] = {1.0,
                                     1.0442737824274138,
                                     1.0905077326652577,
                                     1.1387886347566916,
                                     1.189207115002721,
                                     1.241857812073484,
                                     1.2968395546510096,
                                     1.3542555469368927,
                                     1.4142135623730951,
                                     1.4768261459394993,
                                     1.5422108254079407,
                                     1.6104903319492543,
                                     1.681792830507429,
                                     1.7562521603732995,
                                     1.8340080864093424,
                                     1.9152065613971474};

// exp(x) = 2^e * T[j] * P6(r),  x = (16 e + j) ln2/16 + r,  |r| <= ln2/32: ~5e-16 relative for |x| <= 700;
// the caller routes anything else (huge, inf, NaN) to libm exp().  tab = the 16-entry table in shared memory
// (one entry per 8-byte bank: lanes with different j never conflict).
__device__ __forceinline__ double exp_tab16(double x, const double *tab) {
    const double t = fma(x, 23.083120654223414, 6755399441055744.0);   // 1.5*2^52: low word = rint(16 x / ln2)
    const int ki = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -0.04332169877307024, x);
    r = fma(kd, -1.1926343307941173e-11, r);
    double p = 1.38888888888888888889e-03;
    p = fma(p, r, 8.33333333333333333333e-03);
    p = fma(p, r, 4.16666666666666666667e-02);
    p = fma(p, r, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = p * tab[ki & 15];
    return __hiloint2double(__double2hiint(v) + ((ki >> 4) << 20), __double2loint(v));
}
// |x| <= 700 (false for NaN / inf) from the high word alone: integer pipe, no FP64 compare
__device__ __forceinline__ int exp_arg_hi(double x) { return __double2hiint(x) & 0x7fffffff; }
constexpr int kExpHiMax = 0x4085e000;   // high word of 700.0

// V[d,m] row-major -> fragment order.  Vp[row t][strip q][k-step jj][n-tile tt][lane]:
//   lane = 4 n + kk  holds  V[t*ncol + 4C q + C kk + jj][8 tt + n]   (0 outside the matrix; C = columns per lane)
__global__ void vo_grid_pack_kernel(GridDev G, const double *__restrict__ V, int m, int NT, double *__restrict__ Vp) {
    const int C = G.cols;
    const int per_row = G.nstrips * C * NT * 32;
    const long long total = (long long)(G.ny + 1) * per_row;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(idx / per_row);
        int rem = (int)(idx - (long long)t * per_row);
        const int lane = rem & 31;
        rem >>= 5;
        const int tt = rem % NT;
        rem /= NT;
        const int jj = rem % C, q = rem / C;
        const int c = 4 * C * q + C * (lane & 3) + jj, col = 8 * tt + (lane >> 2);
        Vp[idx] = (c < G.ncol && col < m) ? V[((long long)t * G.ncol + c) * m + col] : 0.0;
    }
}

// RHO = false: contraction with V inside the kernel (NT n-tiles, m <= 32), output r[B,m].
// RHO = true : no contraction; the fine residual rho[b, i] = cvs * S_i - f_i goes to r (row pitch = m doubles,
//              K padding [d, m) zeroed) for the tensor-core GEMM of vo_gemm.cuh (m > 32).  Vp is unused.
// R = node rows per pipeline stage: stage ts holds y rows [R ts, R ts + R), pixel rows and packed V rows
//     [R ts - 1, R ts + R - 1) (clipped to the mesh); barrier traffic and staging overhead are per stage.
// W = warps per CTA: 16 with one CTA per SM, or 8 with two CTAs per SM (two independent rings per SM break the
//     per-stage lockstep of a single ring).
// C = node columns per lane: 4 (16 warps of 120 registers) or 8 (8 warps with twice the work per step: half
//     the per-step overhead per node, 12.5 % instead of 25 % redundant exp()).
template <int NT, bool RHO, int R, int W, int C>
__global__ void __launch_bounds__(W * 32, (W == 16 || C == 8) ? 1 : 2)
vo_grid_kernel(GridDev G, const double *__restrict__ a, long long a_stride, int a_is_log,
               const double *__restrict__ y, const double *__restrict__ g, long long g_stride,
               const double *__restrict__ Vp, int m, double *__restrict__ r, long long B, int NS,
               int stage_bytes, int dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *stages = smem_raw;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)NS * stage_bytes);
    unsigned long long *empty = full + NS;
    double *tab = reinterpret_cast<double *>(empty + NS);

    constexpr int kGridWarps = W, kGridThreads = W * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp % G.nstrips, grp = warp / G.nstrips;
    const int s = lane >> 2, k = lane & 3;
    const int S = 8 * G.groups;
    const int sl = grp * 8 + s;                       // sample slot inside the CTA
    const long long cta_b0 = (long long)blockIdx.x * S;
    long long b = cta_b0 + sl;
    const bool b_valid = b < B;
    if (b >= B) b = B - 1;                            // duplicates the last sample; never stored
    const int ncol = G.ncol, nx = G.nx, ny = G.ny;
    const long long d = (long long)ncol * (ny + 1);
    const int c0 = 4 * C * q + C * k;
    const int n_steps = ny + 2;                       // node rows 0..ny are produced at steps 1..ny+1
    const int n_stages = (n_steps + R - 1) / R;
    const int v_row_bytes = G.nstrips * C * NT * 32 * 8;
    const int row_bytes = ncol * 8, prow_bytes = nx * 8;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(full + i, kGridThreads + 1);
            mbar_init(empty + i, kGridWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 16) tab[threadIdx.x] = kExp16Tab[threadIdx.x];
    __syncthreads();

    // ---- staging, NS stages in a ring, issued NS-1 stages ahead of their use.  The a / y rows are 0.5 KB
    // pieces per sample: too small for the bulk-copy engine (measured: ~130 cycles per request) and one warp can
    // keep only a few cp.async in flight (measured: a single staging warp delivers 4 GB/s), so EVERY warp stages
    // the rows of its samples w, w+16, ... with 16-byte cp.async (LDGSTS, L2 -> shared, no registers) whose
    // completion arrives on the stage's mbarrier; the packed V rows are ONE bulk copy by thread 0.  The R rows of
    // a sample are contiguous in global memory and are copied as one block; y rows start on 8-byte boundaries:
    // the block is copied from the enclosing 16-byte boundary and the consumer adds the same shift.
    const unsigned stages_u32 = smem_u32(stages);
    // this warp stages samples cta_b0 + warp + 16 i (i < n_mine); samples past the batch are not staged (their
    // lanes compute on stale shared memory and are never stored)
    const int n_mine = (int)max(0ll, min((long long)(S - warp + kGridWarps - 1) / kGridWarps,
                                         (B - cta_b0 - warp + kGridWarps - 1) / kGridWarps));
    const double *a_w = a + (cta_b0 + warp) * a_stride + G.in0;                      // pixel (0,0) of the first sample
    const char *y_w = reinterpret_cast<const char *>(y + (cta_b0 + warp) * d);        // node row 0
    const long long a_step = (long long)kGridWarps * a_stride, y_step = (long long)kGridWarps * d * 8;
    const unsigned dst_a0 = G.a_off + warp * (G.a_stride * 8);
    const unsigned dst_y0 = G.y_off + (warp * G.y_stride + 2 * ((warp >> 1) & 1)) * 8;
    const bool y_end_odd = (((unsigned long long)(y + B * d)) & 15ull) != 0;
    auto issue_stage = [&](int ts, int slot) {
        unsigned char *st = stages + (size_t)slot * stage_bytes;
        unsigned long long *bar = full + slot;
        const int t0 = R * ts;
        if (!(dbg & 2)) {
            const unsigned sbase = stages_u32 + slot * stage_bytes;
            const int plo = max(0, t0 - 1), phi = min(ny, t0 - 1 + R);           // pixel rows [plo, phi)
            if (phi > plo) {
                const int first = G.sy > 0 ? plo : phi - 1;                        // lowest address
                const int pos = G.sy > 0 ? plo - (t0 - 1) : R - 1 - (phi - 1 - (t0 - 1));
                const int bytes = (phi - plo) * prow_bytes;
                const char *src = reinterpret_cast<const char *>(a_w + (long long)first * G.sy);
                unsigned dst = sbase + dst_a0 + pos * prow_bytes;
                for (int i = 0; i < n_mine; ++i, src += a_step * 8, dst += kGridWarps * G.a_stride * 8)
                    for (int o = 16 * lane; o < bytes; o += 512) cp_async16_u32(dst + o, src + o);
            }
            const int yhi = min(ny + 1, t0 + R);                                   // node rows [t0, yhi)
            if (yhi > t0) {
                const int bytes = (yhi - t0) * row_bytes;
                const char *src = y_w + (long long)t0 * row_bytes;
                unsigned dst = sbase + dst_y0;
                for (int i = 0; i < n_mine; ++i, src += y_step, dst += kGridWarps * G.y_stride * 8) {
                    const int shift = (int)((unsigned long long)src & 15ull);
                    // 16-byte pieces at offsets o = 16*lane - shift, + 512, ... from the block start; the last piece
                    // of the whole tensor may stick out past its end (odd element count): left to the tail lanes
                    const bool last = y_end_odd && yhi == ny + 1 && cta_b0 + warp + kGridWarps * i == B - 1;
                    for (int o = 16 * lane - shift; o < bytes; o += 512)
                        if (!last || o + 24 <= bytes) cp_async16_u32(dst + o + shift, src + o);
                }
            }
        }
        // (one arrival per THREAD; the alternative -- cp.async groups + wait_group + one arrival per warp -- halves the
        // empty ring's cost (42 -> 23 us) but makes the full kernel slower (153 -> 159 us): it adds a second CTA-wide
        // rendezvous per stage.  Measured on B200, round 1.)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        if (threadIdx.x == 0) {
            const int vlo = max(0, t0 - 1), vhi = min(ny + 1, t0 - 1 + R);       // packed V rows [vlo, vhi)
            if (!RHO && vhi > vlo && !(dbg & 4)) {
                const unsigned bytes = (unsigned)(vhi - vlo) * v_row_bytes;
                mbar_arrive_expect_tx(bar, bytes);
                bulk_g2s(st + G.v_off + (vlo - (t0 - 1)) * v_row_bytes,
                         reinterpret_cast<const char *>(Vp) + (size_t)vlo * v_row_bytes, bytes, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    };
    int i_slot = 0;                  // slot of the next stage to issue
    unsigned e_par = 0;              // per-slot parity of the next wait on empty[]
    for (int ts = 0; ts < NS - 1 && ts < n_stages; ++ts) {
        issue_stage(ts, i_slot);
        if (++i_slot == NS) i_slot = 0;
    }

    // ---- per-lane constants
    // category of columns c0-1 .. c0+4: 0 shared memory, 1 left Dirichlet value, 2 right one, 3 zero
    int code = 0;
#pragma unroll
    for (int p = -1; p <= C; ++p) {
        const int c = c0 + p;
        const int cat = (c == -1) ? 1 : (c < ncol ? 0 : (c == ncol ? 2 : 3));
        code |= cat << (2 * (p + 1));
    }
    int pixmask = 0, nodemask = 0;
#pragma unroll
    for (int j = 0; j <= C; ++j) pixmask |= (c0 + j < nx) ? (1 << j) : 0;
#pragma unroll
    for (int j = 0; j < C; ++j) nodemask |= (c0 + j < ncol) ? (1 << j) : 0;
    asm volatile("" : "+r"(code), "+r"(pixmask), "+r"(nodemask));   // opaque: keep them live instead of recomputing per step
    const bool edge_lane = code != 0;
    const bool need_gl = (code & 3) == 1;
    bool need_gr = false;
#pragma unroll
    for (int p = 0; p <= C + 1; ++p) need_gr |= ((code >> (2 * p)) & 3) == 2;
    const double *yb = y + b * d;
    const double *gb = g ? g + b * g_stride : nullptr;
    // the very last element of y cannot be copied in a 16-byte piece when the tensor ends off a 16-byte boundary
    const int tail_p = ncol - 1 - c0;   // window position (-1..C) of the last free column, if inside
    const bool tail_lane = y_end_odd && b == B - 1 && tail_p >= -1 && tail_p <= C;
    // byte offsets of this lane's first column inside a stage; the 8-byte phase of a stage's first y row
    // alternates from stage to stage when R * ncol is odd
    const int y_lane_off = G.y_off + (sl * G.y_stride + 2 * ((sl >> 1) & 1) + c0) * 8;
    const int a_lane_off = G.a_off + (sl * G.a_stride + c0) * 8;
    const int v_lane_off = G.v_off + (q * C * NT * 32 + lane) * 8;
    int y_shift = (int)(((unsigned long long)yb) & 15ull);
    const int y_shift_step = ((R * ncol) & 1) * 8;

    double acc[NT][2];
#pragma unroll
    for (int tt = 0; tt < NT; ++tt) acc[tt][0] = acc[tt][1] = 0.0;
    constexpr int kPixAll = (1 << (C + 1)) - 1, kNodeAll = (1 << C) - 1;
    double uc[C], ulc = 0.0, urc = 0.0, ap[C + 1], fvp[C];
#pragma unroll
    for (int j = 0; j < C; ++j) uc[j] = fvp[j] = 0.0;
#pragma unroll
    for (int j = 0; j <= C; ++j) ap[j] = 0.0;
    double gl_next = (need_gl && gb) ? gb[0] : 0.0, gr_next = (need_gr && gb) ? gb[1] : 0.0;

    int c_slot = 0;                  // slot of the stage being consumed
    unsigned f_par = 0;              // per-slot parity of the next wait on full[]
    for (int ts = 0; ts < n_stages; ++ts) {
        {
            const int tn = ts + NS - 1;
            if (tn < n_stages) {
                if (ts >= 1) {       // the slot held stage ts-1: wait until every warp has released it
                    mbar_wait(empty + i_slot, (e_par >> i_slot) & 1);
                    e_par ^= 1u << i_slot;
                }
                issue_stage(tn, i_slot);
                if (++i_slot == NS) i_slot = 0;
            }
        }
        const unsigned char *st = stages + (size_t)c_slot * stage_bytes;
        mbar_wait(full + c_slot, (f_par >> c_slot) & 1);
        f_par ^= 1u << c_slot;

#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int t = R * ts + rr;
            if (t >= n_steps || (dbg & 1)) break;
            // ---- new node row t and pixel row t-1
            double un[C], unl = 0.0, unr = 0.0, an[C + 1];
#pragma unroll
            for (int j = 0; j < C; ++j) un[j] = 0.0;
#pragma unroll
            for (int j = 0; j <= C; ++j) an[j] = 0.0;
            if (t <= ny) {
                const double gl = gl_next, gr = gr_next;
                if (t < ny && gb) {
                    if (need_gl) gl_next = gb[2 * (t + 1)];
                    if (need_gr) gr_next = gb[2 * (t + 1) + 1];
                }
                const double *yr = reinterpret_cast<const double *>(st + y_lane_off + y_shift + rr * row_bytes);
                unl = yr[-1];
#pragma unroll
                for (int j = 0; j < C; ++j) un[j] = yr[j];
                unr = yr[C];
                if (tail_lane && t == ny) {
                    const double v = __ldg(yb + d - 1);
                    if (tail_p == -1) unl = v;
#pragma unroll
                    for (int j = 0; j < C; ++j)
                        if (tail_p == j) un[j] = v;
                    if (tail_p == C) unr = v;
                }
                if (edge_lane) {
                    auto pick = [&](int p, double v) {
                        const int cat = (code >> (2 * p)) & 3;
                        return cat == 0 ? v : (cat == 1 ? gl : (cat == 2 ? gr : 0.0));
                    };
                    unl = pick(0, unl);
#pragma unroll
                    for (int j = 0; j < C; ++j) un[j] = pick(j + 1, un[j]);
                    unr = pick(C + 1, unr);
                }
            }
            if (t >= 1 && t <= ny) {
                const int pos = G.sy > 0 ? rr : R - 1 - rr;
                const double *ar = reinterpret_cast<const double *>(st + a_lane_off + pos * prow_bytes);
#pragma unroll
                for (int j = 0; j < C; j += 2) {
                    const double2 pp = *reinterpret_cast<const double2 *>(ar + j);
                    an[j] = pp.x; an[j + 1] = pp.y;
                }
                an[C] = ar[C];
                if (pixmask != kPixAll) {   // columns past the last pixel hold stale shared memory
#pragma unroll
                    for (int j = 0; j <= C; ++j) an[j] = ((pixmask >> j) & 1) ? an[j] : 0.0;
                }
                if (a_is_log) {
                    int hmax = exp_arg_hi(an[C]);
#pragma unroll
                    for (int j = 0; j < C; ++j) hmax = max(hmax, exp_arg_hi(an[j]));
                    if (hmax <= kExpHiMax) {
#pragma unroll
                        for (int j = 0; j <= C; ++j) an[j] = exp_tab16(an[j], tab);
                    } else {
#pragma unroll
                        for (int j = 0; j <= C; ++j) an[j] = exp(an[j]);
                    }
                }
                if (pixmask != kPixAll) {
#pragma unroll
                    for (int j = 0; j <= C; ++j) an[j] = ((pixmask >> j) & 1) ? an[j] : 0.0;
                }
            }

            // ---- node row t-1: fluxes -> S -> tensor-core contraction with packed V row t-1 (or rho output)
            if (t >= 1) {
                double fh[C + 1];
                fh[0] = (ap[0] + an[0]) * (uc[0] - ulc);
#pragma unroll
                for (int j = 1; j < C; ++j) fh[j] = (ap[j] + an[j]) * (uc[j] - uc[j - 1]);
                fh[C] = (ap[C] + an[C]) * (urc - uc[C - 1]);
                double Sv[C];
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    const double fv = (an[j] + an[j + 1]) * (un[j] - uc[j]);
                    Sv[j] = fma(G.rh, fh[j + 1] - fh[j], fv - fvp[j]);
                    fvp[j] = fv;
                }
                if (G.has_load) {
#pragma unroll
                    for (int j = 0; j < C; ++j)
                        if ((nodemask >> j) & 1) Sv[j] -= __ldg(G.f_over + (long long)(t - 1) * ncol + c0 + j);
                }
                if (nodemask != kNodeAll) {
#pragma unroll
                    for (int j = 0; j < C; ++j) Sv[j] = ((nodemask >> j) & 1) ? Sv[j] : 0.0;
                }
                if constexpr (RHO) {
                    if (b_valid) {
                        double *dst = r + b * (long long)m + (long long)(t - 1) * ncol + c0;
#pragma unroll
                        for (int j = 0; j < C; ++j)
                            if ((nodemask >> j) & 1) dst[j] = G.scale * Sv[j];
                    }
                } else {
                    const double *vs = reinterpret_cast<const double *>(st + v_lane_off + rr * v_row_bytes);
#pragma unroll
                    for (int j0 = 0; j0 < C; j0 += 4) {   // B fragments four k-steps at a time
                        double bf[4][NT];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                            for (int tt = 0; tt < NT; ++tt) bf[jj][tt] = vs[((j0 + jj) * NT + tt) * 32];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                            for (int tt = 0; tt < NT; ++tt) dmma884(acc[tt][0], acc[tt][1], Sv[j0 + jj], bf[jj][tt]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < C; ++j) uc[j] = un[j];
            ulc = unl; urc = unr;
#pragma unroll
            for (int j = 0; j <= C; ++j) ap[j] = an[j];
        }
        y_shift = (y_shift + y_shift_step) & 15;
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + c_slot);
        if (++c_slot == NS) c_slot = 0;
    }

    if constexpr (RHO) {
        // zero the K padding [d, m) of this CTA's rows
        const int pad = m - (int)d;
        for (int idx = threadIdx.x; idx < S * pad; idx += kGridThreads) {
            const int si = idx / pad, c = idx - si * pad;
            if (cta_b0 + si < B) r[(cta_b0 + si) * (long long)m + d + c] = 0.0;
        }
        return;
    }
    // ---- sum the strips' partial tiles and store r = cvs * sum   (stage memory is free now)
    __syncthreads();
    double *red = reinterpret_cast<double *>(stages);   // [groups][nstrips][8 samples][NT*8]
    {
        double *dst = red + (((size_t)grp * G.nstrips + q) * 8 + s) * (NT * 8) + 2 * k;
#pragma unroll
        for (int tt = 0; tt < NT; ++tt) {
            dst[tt * 8] = acc[tt][0];
            dst[tt * 8 + 1] = acc[tt][1];
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < S * NT * 8; idx += kGridThreads) {
        const int si = idx / (NT * 8), col = idx - si * (NT * 8);
        const long long bs = cta_b0 + si;
        if (col < m && bs < B) {
            const int gi = si >> 3, ss = si & 7;
            double v = 0.0;
            for (int qq = 0; qq < G.nstrips; ++qq) v += red[(((size_t)gi * G.nstrips + qq) * 8 + ss) * (NT * 8) + col];
            r[bs * m + col] = G.scale * v;
        }
    }
}

}  // namespace gpde
