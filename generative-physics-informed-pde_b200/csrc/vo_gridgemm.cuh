// Structured-grid virtual-observable residual for MANY weighting functions (m > 64), one kernel (sm_100a).
// Included by vo.cu after vo_grid2.cuh and vo_gemm.cuh.
//
//   r[b,:] = V^T (K_fom(a_b) u~_b - f)_free        (VirtualObservables.py:61-69, 662, 990)
//
// vo_grid2_kernel<rho> + vo_gemm_kernel move the fine residual rho [B,d] through HBM (2.1 GB written and read at BASELINE
// config 3: ~1.1 of 5.5 ms).  Here the CTA that contracts a tile of 64 samples with V also PRODUCES the tile's rho values,
// 16 nodes of one node row at a time, straight into the A operand of the contraction in shared memory:
//   * the contraction index runs strip-major: strip q (16 node columns), node rows bottom to top inside the strip; a chunk
//     = the 16 nodes (q, t) -- V is re-packed in this order, one chunk = one contiguous stage image [16][BN + 4] with the
//     chunk's tile mask in the padding, fetched by ONE bulk copy (TMA engine) two chunks ahead (ring of 4);
//   * four producer warps, one per scheduler (32 samples x 8 node columns each, lane = sample), march up the strip with
//     their own window of node rows of y and pixel rows of a in shared memory (cp.async one step ahead, 8 lanes per sample
//     row segment; Dirichlet columns come from g the same way), form the flux-form residual of the row (same expressions
//     and order as vo_grid2.cuh) and write their quarter of the 64 x 16 tile K-major;
//   * eight consumer warps (2 x 4, warp tile 32 x BN/4, n-tiles interleaved over the warps) run DMMA.8x8x4 out of the
//     two rings; an n-tile whose 16 x 8 block of V is all zero in this chunk is skipped (exact: it drops products with 0
//     -- the coarse-grained-residual columns W, VirtualObservables.py:297-321, are P1 hat functions: 1-3 of their 10
//     n-tiles are non-zero in a chunk);
//   * roles get their registers by setmaxnreg (consumers 200, producers / loader 96: 384 x 168 registers are allocated at launch); full / empty mbarriers per stage.
// The contraction length is cut over gridDim.z by whole strips (partial tiles + vo_gemm_reduce_kernel, deterministic).
#pragma once

namespace gpde {

constexpr int kGGBM = 64;                 // samples per CTA
constexpr int kGGKC = 16;                 // nodes per chunk
// 8 consumer warps + 2 * (16 / NW) producer warps, NW = node columns per producer warp (8 or 4)
static inline __host__ __device__ constexpr int gg_threads(int nw) { return 256 + 64 * (16 / nw); }
constexpr int kGGStagesA = 3;
constexpr int kGGLdA = kGGBM + 4;         // doubles per k row of the rho tile (= 4 mod 16: conflict-free fragment loads)
// a producer warp's window: 3 slots of NW + 2 y columns (its nodes + halo), 3 slots of NW + 1 pixel columns; one slot of
// zeros (the rows outside the mesh) is shared by the warps; a column holds the warp's 32 samples at a pitch that makes both
// the copies (NW lanes per sample row segment) and the reads (a lane per sample) conflict-free
static inline __host__ __device__ constexpr int gg_win_cols(int nw) { return 3 * (nw + 2) + 3 * (nw + 1); }
static inline __host__ __device__ constexpr int gg_win_pitch(int elem, int nw) { return 32 + (elem == 8 ? 16 : 32) / nw; }

struct GGDev {
    int nx, ny, ncol, nstrips;
    long long in0, sy;       // conductivity entry of pixel (cx, cy) = in0 + cy * sy + cx
    double rh, scale;
    int ctiles;              // column tiles of BN
    int stages_b;            // V stages in the ring
};

static inline size_t gg_smem_bytes(int bn, int stages_b, int elem, int nw) {
    return (size_t)stages_b * sizeof(double) * kGGKC * (bn + 4) + (size_t)kGGStagesA * sizeof(double) * kGGKC * kGGLdA +
           (size_t)elem * gg_win_pitch(elem, nw) * (2 * (16 / nw) * gg_win_cols(nw) + nw + 2) + 16 +
           sizeof(unsigned long long) * (2 * kGGStagesA + 2 * 8);
}
static inline size_t gg_packed_bytes(int bn, int ctiles, int chunks) {
    return (size_t)ctiles * chunks * sizeof(double) * kGGKC * (bn + 4);
}

// V[d,m] row-major -> [column tile][chunk (q, t)][k < 16][BN + 4]; row k of chunk (q, t) = node row t, free column 16 q + k
// (zero past the last free column and past column m); the 4 padding doubles of row 0 hold the chunk's tile mask
// (bit j <=> the 16 x 8 block of columns 8 j .. 8 j + 7 has a non-zero entry), the other padding is zero.
template <int BN, typename TV>
__global__ void vo_gridgemm_pack_kernel(GGDev G, const TV *__restrict__ V, int m, double *__restrict__ Vp) {
    constexpr int LdB = BN + 4, SEG = BN / 32;
    const int chunks = G.nstrips * (G.ny + 1);
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = warp; w < (long long)G.ctiles * chunks; w += nwarps) {
        const int ct = (int)(w / chunks), ch = (int)(w - (long long)ct * chunks);
        const int q = ch / (G.ny + 1), t = ch - q * (G.ny + 1);
        double *dst = Vp + (size_t)w * kGGKC * LdB;
        unsigned nz[SEG];
#pragma unroll
        for (int sg = 0; sg < SEG; ++sg) nz[sg] = 0;
        for (int k = 0; k < kGGKC; ++k) {
            const int c = 16 * q + k;
            const TV *src = V + ((long long)t * G.ncol + c) * m;
#pragma unroll
            for (int sg = 0; sg < SEG; ++sg) {
                const int col = ct * BN + sg * 32 + lane;
                const double v = (c < G.ncol && col < m) ? (double)src[col] : 0.0;
                dst[k * LdB + sg * 32 + lane] = v;
                nz[sg] |= (v != 0.0) ? 1u : 0u;
            }
            if (lane < 4 && k > 0) dst[k * LdB + BN + lane] = 0.0;
        }
        unsigned mask = 0;
#pragma unroll
        for (int sg = 0; sg < SEG; ++sg) {
            const unsigned bal = __ballot_sync(0xffffffffu, nz[sg] != 0);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((bal >> (8 * u)) & 0xffu) mask |= 1u << (sg * 4 + u);
        }
        if (lane == 0) {
            unsigned *mw = reinterpret_cast<unsigned *>(dst + BN);
            mw[0] = mask;
            for (int i = 1; i < 8; ++i) mw[i] = 0;
        }
    }
}

template <typename T> __device__ __forceinline__ void gg_cp_async_elem(unsigned dst, const T *src);
template <> __device__ __forceinline__ void gg_cp_async_elem<double>(unsigned dst, const double *src) { cp_async8_u32(dst, src); }
template <> __device__ __forceinline__ void gg_cp_async_elem<float>(unsigned dst, const float *src) { cp_async4_u32(dst, src); }
template <typename T> __device__ __forceinline__ void gg_sts_elem(unsigned dst, double v);
template <> __device__ __forceinline__ void gg_sts_elem<double>(unsigned dst, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(dst), "d"(v) : "memory");
}
template <> __device__ __forceinline__ void gg_sts_elem<float>(unsigned dst, double v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst), "f"((float)v) : "memory");
}

// Log-field input (np.exp(self._x) of VirtualObservables.py:57): the conductivities are formed by one streaming pass before the
// contraction kernel (inside it, 17 exp() per producer step tripled the FP64 instructions that compete with the DMMAs for the
// pipe: 5.62 ms at config 3 against 4.26 ms + this pass).  rows = 1 for a field shared by the batch.
template <typename T>
__global__ void vo_exp_rows_kernel(const T *__restrict__ x, long long n, T *__restrict__ out) {
    __shared__ double tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = kExp256Tab[i];
    __syncthreads();
    const unsigned tab32 = smem_u32(tab);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        out[i] = (T)(exp256_in_range(v) ? exp_tab256c(v, tab32) : exp(v));
    }
}

// grid: (sample tiles of 64, column tiles of BN, parts of the contraction length = groups of whole strips)
template <int BN, int NW, typename T, typename TR>
__global__ void __launch_bounds__(gg_threads(NW), 1)
vo_gridgemm_kernel(GGDev G, const T *__restrict__ a, long long a_stride, const T *__restrict__ y, const T *__restrict__ g,
                   long long g_stride, const double *__restrict__ Vp, TR *__restrict__ R, int m, int ldp, long long B,
                   double *__restrict__ partial) {
    constexpr int LdB = BN + 4, NT = BN / 32;     // n-tiles per consumer warp (4 warps along N, interleaved)
    constexpr int A_STAGE = kGGKC * kGGLdA, B_STAGE = kGGKC * LdB;
    constexpr int E = (int)sizeof(T), P = gg_win_pitch(E, NW);
    constexpr int CG = 16 / NW, NPROD = 2 * CG;                 // column groups of a strip; producer warps
    constexpr int YC = NW + 2, AC = NW + 1;                     // window columns per slot
    constexpr int kThreads = gg_threads(NW);
    constexpr int WIN_BYTES = gg_win_cols(NW) * P * E;          // one producer warp's window
    extern __shared__ __align__(128) unsigned char gg_smem[];
    const int SB = G.stages_b;
    double *Bs = reinterpret_cast<double *>(gg_smem);
    double *As = Bs + (size_t)SB * B_STAGE;
    unsigned char *win = reinterpret_cast<unsigned char *>(As + kGGStagesA * A_STAGE);
    constexpr int WIN_ALL = NPROD * WIN_BYTES + YC * P * E;      // + the shared slot of zeros
    unsigned long long *bars = reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(win + WIN_ALL) + 15) & ~(uintptr_t)15);
    unsigned long long *fullA = bars, *emptyA = bars + kGGStagesA, *fullB = bars + 2 * kGGStagesA, *emptyB = fullB + 8;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row0 = (long long)blockIdx.x * kGGBM;
    const int rows_per_strip = G.ny + 1;
    const int q_lo = (int)((long long)G.nstrips * blockIdx.z / gridDim.z), q_hi = (int)((long long)G.nstrips * (blockIdx.z + 1) / gridDim.z);
    const int n_chunks = (q_hi - q_lo) * rows_per_strip;
    const long long chunk0 = (long long)blockIdx.y * G.nstrips * rows_per_strip + (long long)q_lo * rows_per_strip;

    // ---- setup: windows zeroed (rows outside the mesh and the slots of samples past the batch read as 0), barriers, exp table
    {
        unsigned *w32 = reinterpret_cast<unsigned *>(win);
        for (int i = tid; i < WIN_ALL >> 2; i += kThreads) w32[i] = 0u;
    }
    if (tid == 0) {
        for (int i = 0; i < kGGStagesA; ++i) { mbar_init(fullA + i, 32 * NPROD); mbar_init(emptyA + i, 8); }
        for (int i = 0; i < SB; ++i) { mbar_init(fullB + i, 1); mbar_init(emptyB + i, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < 8) {
        // =================================================== consumers: DMMA out of the two rings
        // registers: 384 x 168 (NW = 8) / 512 x 128 (NW = 4) are allocated at launch and re-divided between the roles
        if constexpr (NW == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
        const int wm = warp & 1, wn = warp >> 1;
        const int gq = lane >> 2, tq = lane & 3;
        double acc[4][NT][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        const unsigned as0 = smem_u32(As) + 8u * (tq * kGGLdA + wm * 32 + gq);      // + stage, + ks * 4 * LdA, + mt * 8
        const unsigned bs0 = smem_u32(Bs) + 8u * (tq * LdB + wn * 8 + gq);          // + stage, + ks * 4 * LdB, + nt * 32
        const unsigned mk0 = smem_u32(Bs) + 8u * BN;
        int sa = 0, sb = 0;
        unsigned pa = 0, pb = 0;
        // packed V: one bulk copy per chunk, issued by warp 0 two chunks ahead into the slot of the chunk every warp left two
        // iterations ago (SB = 4: no waiting in practice; the producers never touch this ring)
        const unsigned vbytes = (unsigned)(B_STAGE * 8);
        const int LA = SB - 2;           // chunks ahead: the slot of chunk kc + LA last held chunk kc - 2
        int vs = LA, vwrap = 0;          // slot of the next chunk to fetch; how often the ring has wrapped
        if (warp == 0 && lane == 0) {
            for (int v = 0; v < LA && v < n_chunks; ++v) {
                mbar_arrive_expect_tx(fullB + v, vbytes);
                bulk_g2s(Bs + (size_t)v * B_STAGE, Vp + (size_t)(chunk0 + v) * B_STAGE, vbytes, fullB + v);
            }
        }
        for (int kc = 0; kc < n_chunks; ++kc) {
            if (warp == 0) {
                if (kc + LA < n_chunks) {
                    if (vwrap > 0) mbar_wait(emptyB + vs, (unsigned)(vwrap - 1) & 1u);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(fullB + vs, vbytes);
                        bulk_g2s(Bs + (size_t)vs * B_STAGE, Vp + (size_t)(chunk0 + kc + LA) * B_STAGE, vbytes, fullB + vs);
                    }
                }
                if (++vs == SB) { vs = 0; ++vwrap; }
            }
            mbar_wait(fullB + sb, pb);
            mbar_wait(fullA + sa, pa);
            const unsigned ab = as0 + sa * (A_STAGE * 8), bb = bs0 + sb * (B_STAGE * 8);
            unsigned mask;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(mask) : "r"(mk0 + sb * (B_STAGE * 8)));
            double af[4][4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) af[ks][mt] = lds64(ab + 8u * (ks * 4 * kGGLdA + mt * 8));
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                if ((mask >> (wn + 4 * nt)) & 1u) {
                    double bf[4];
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) bf[ks] = lds64(bb + 8u * (ks * 4 * LdB + nt * 32));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[ks][mt], bf[ks]);
                }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(emptyA + sa); mbar_arrive(emptyB + sb); }
            if (++sa == kGGStagesA) { sa = 0; pa ^= 1; }
            if (++sb == SB) { sb = 0; pb ^= 1; }
        }
        // epilogue: lane holds C[row = lane / 4][cols 2 (lane % 4), + 1] of every 8 x 8 tile
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const long long row = row0 + wm * 32 + mt * 8 + gq;
            if (row >= B) continue;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = blockIdx.y * BN + (wn + 4 * nt) * 8 + 2 * tq;
                if (partial) {
                    *reinterpret_cast<double2 *>(partial + ((long long)blockIdx.z * B + row) * ldp + col) =
                        make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                } else {
                    if (col < m) R[row * m + col] = (TR)acc[mt][nt][0];
                    if (col + 1 < m) R[row * m + col + 1] = (TR)acc[mt][nt][1];
                }
            }
        }
        return;
    }

    // ======================================================= producers: rho of 32 samples x 8 nodes per warp and step
    if constexpr (NW == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int pw = warp - 8, sh = pw / CG, ch = pw % CG;    // sample half, column group of the strip
    const int ncol = G.ncol, ny = G.ny, nx = G.nx;
    const long long d = (long long)ncol * (ny + 1);
    const int nvw = (int)max(0ll, min(32ll, B - row0 - sh * 32));    // valid samples of this warp
    const unsigned wb = smem_u32(win) + pw * WIN_BYTES;
    const unsigned zero_b = smem_u32(win) + NPROD * WIN_BYTES + lane * E;
    // copies: NW lanes per sample: lane group lq takes samples SPI i + lq (i < NW), lane lc of it the window column lc (< NW);
    // window columns NW (NW + 1) of sample ``lane``, and the Dirichlet values g[2 t], g[2 t + 1] of the mesh's first / last
    // column, by that lane.
    // Sources are 32-bit byte offsets from the first sample of this warp (the launcher checks that they fit).
    constexpr int SPI = 32 / NW;        // samples per copy instruction
    const int lq = lane / NW, lc = lane % NW;
    const char *y_w = reinterpret_cast<const char *>(y + (row0 + sh * 32) * d);
    const char *a_w = reinterpret_cast<const char *>(a + (row0 + sh * 32) * a_stride + G.in0);
    const char *g_w = reinterpret_cast<const char *>(g ? g + (row0 + sh * 32) * g_stride : nullptr);
    const unsigned y_smp = (unsigned)(SPI * d * E), a_smp = (unsigned)(SPI * a_stride * E);
    const unsigned y_adv = (unsigned)(ncol * E);
    const int a_adv = (int)(G.sy * E);
    unsigned oy = 0, ox = 0, og = 0;    // next node row: main copy of this lane, column 8 of the lane's sample, Dirichlet pair
    int oa = 0, oax = 0;                // next pixel row
    bool left = false, right = false;   // the warp's first / last node column is a Dirichlet column
    auto set_strip = [&](int q) {       // sources of the strip's rows 0
        left = q == 0 && ch == 0;
        right = ch == CG - 1 && 16 * q + 16 >= nx;
        const int c = 16 * q + NW * ch - 1;                   // free column behind window column 0
        oy = (unsigned)((lq * d + c + lc) * E);
        ox = (unsigned)((lane * d + c + NW) * E);
        og = (unsigned)(lane * g_stride * E);
        oa = (int)((lq * a_stride + 16 * q + NW * ch + lc) * E);
        oax = (int)((lane * a_stride + 16 * q + NW * ch + NW) * E);
    };
    auto load_y_row = [&](int slot) {   // next node row of the strip -> y slot
        const unsigned dst = wb + (unsigned)((slot * YC + lc) * P + lq) * E;
        if (!(left && lc == 0)) {
            // (rolled on purpose: unrolled, the eight per-sample base addresses become loop invariants that do not fit the
            // producers' register budget)
            const char *src = y_w + oy;
#pragma unroll 1
            for (int i = 0; i < 32; i += SPI, src += y_smp)
                if (i + lq < nvw) gg_cp_async_elem<T>(dst + i * E, reinterpret_cast<const T *>(src));
        }
        const unsigned dx = wb + (unsigned)((slot * YC + NW) * P + lane) * E;
        if (lane < nvw) {
            if (!right) {
                gg_cp_async_elem<T>(dx, reinterpret_cast<const T *>(y_w + ox));
                gg_cp_async_elem<T>(dx + P * E, reinterpret_cast<const T *>(y_w + ox + E));
            } else if (g_w) {
                gg_cp_async_elem<T>(dx, reinterpret_cast<const T *>(g_w + og + E));
            } else {
                gg_sts_elem<T>(dx, 0.0);
            }
            if (left && g_w) gg_cp_async_elem<T>(wb + (unsigned)((slot * YC) * P + lane) * E, reinterpret_cast<const T *>(g_w + og));
        }
        oy += y_adv; ox += y_adv; og += 2 * E;
    };
    auto load_a_row = [&](int slot) {   // next pixel row of the strip -> a slot
        const unsigned dst = wb + (unsigned)((3 * YC + slot * AC + lc) * P + lq) * E;
        const char *src = a_w + oa;
#pragma unroll 1
        for (int i = 0; i < 32; i += SPI, src += a_smp)
            if (i + lq < nvw) gg_cp_async_elem<T>(dst + i * E, reinterpret_cast<const T *>(src));
        if (!right && lane < nvw)
            gg_cp_async_elem<T>(wb + (unsigned)((3 * YC + slot * AC + NW) * P + lane) * E, reinterpret_cast<const T *>(a_w + oax));
        oa += a_adv; oax += a_adv;
    };
    auto ld = [&](unsigned base, int col) -> double { return lds_elem<T>(base + (unsigned)(col * P * E)); };

    // slots: at a step, ycs holds the node row t, ycs + 1 the row above, ycs + 2 takes row t + 2; aas holds pixel row t,
    // aas - 1 the row below, aas + 1 takes row t + 1 (all mod 3, advancing by one per step -- across strip boundaries too)
    int ycs = 0, aas = 0;
    set_strip(q_lo);
    load_y_row(0);
    load_y_row(1);
    load_a_row(0);
    cp_async_commit();
    double fvp[NW];
    int sa = 0, kcl = 0;
    unsigned pe = 1;                    // waiting on the "previous" phase of a fresh barrier returns at once
    const double rh = G.rh, scale = G.scale;
    for (int q = q_lo; q < q_hi; ++q) {
#pragma unroll
        for (int j = 0; j < NW; ++j) fvp[j] = 0.0;
        for (int t = 0; t <= ny; ++t, ++kcl) {
            const int y1 = ycs == 2 ? 0 : ycs + 1, y2 = ycs == 0 ? 2 : ycs - 1;      // ycs + 1, ycs + 2 (mod 3)
            const int a1 = aas == 2 ? 0 : aas + 1, am = aas == 0 ? 2 : aas - 1;      // aas + 1, aas - 1 (mod 3)
            // ---- next step's rows (every lane has left the slots they replace: they were last read a step ago)
            __syncwarp();
            if (t < ny) {
                if (t + 2 <= ny) load_y_row(y2);
                if (t + 1 < ny) load_a_row(a1);
            } else if (q + 1 < q_hi) {
                set_strip(q + 1);
                load_y_row(y1);
                load_y_row(y2);
                load_a_row(a1);
            }
            cp_async_commit();
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            mbar_wait(emptyA + sa, pe);
            // ---- this step: node row t between pixel rows t - 1 (below) and t (above)
            const bool has_b = t > 0, has_a = t < ny;
            const bool was_right = ch == CG - 1 && 16 * q + 16 >= nx;
            const unsigned yc_b = wb + (unsigned)(ycs * YC * P + lane) * E;
            const unsigned ya_b = has_a ? wb + (unsigned)(y1 * YC * P + lane) * E : zero_b;
            const unsigned ab_b = has_b ? wb + (unsigned)((3 * YC + am * AC) * P + lane) * E : zero_b;
            const unsigned aa_b = has_a ? wb + (unsigned)((3 * YC + aas * AC) * P + lane) * E : zero_b;
            const unsigned a_dst = smem_u32(As) + 8u * (sa * A_STAGE + (NW * ch) * kGGLdA + sh * 32 + lane);
            double ul = ld(yc_b, 0), uc = ld(yc_b, 1);
            double aBj = ld(ab_b, 0), aAj = ld(aa_b, 0);
            double fh_l = (aBj + aAj) * (uc - ul);
#pragma unroll
            for (int j = 0; j < NW; ++j) {
                const double ur = ld(yc_b, j + 2), aBn = ld(ab_b, j + 1), un = ld(ya_b, j + 1);
                const double aAn = ld(aa_b, j + 1);
                const double fh_r = (aBn + aAn) * (ur - uc);
                const double fv = (aAj + aAn) * (un - uc);
                double Sv = fma(rh, fh_r - fh_l, fv - fvp[j]);
                fvp[j] = fv;
                if (j == NW - 1 && was_right) Sv = 0.0;
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(a_dst + 8u * (j * kGGLdA)), "d"(scale * Sv) : "memory");
                fh_l = fh_r; aBj = aBn; aAj = aAn; ul = uc; uc = ur;
            }
            mbar_arrive(fullA + sa);
            if (++sa == kGGStagesA) { sa = 0; pe ^= 1; }
            ycs = y1; aas = a1;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace gpde
