// Fused virtual-observable residual kernel.  Included by vo.cu.
//
// One launch computes r[b,:] = V^T (K_fom(a_b) u~_b - f)_free for all samples:
//   * a CTA owns 8 samples and walks over tiles of 128 consecutive free rows; 256 threads =
//     (row, sample half): every thread handles one row for 4 samples;
//   * the slices of y and of the conductivity field that a tile's rows touch live in shared-memory
//     ring buffers.  Only the part not resident from the previous tile is fetched, ONE TILE AHEAD, with
//     cp.async straight from global memory into its ring slot (no registers, no exposed DRAM latency);
//     a log-field is exponentiated in place once per value.  Every input byte is read from HBM once;
//   * ring entries are sample-interleaved ([index][8 samples + 2 pad] doubles): a thread's 4 samples
//     of one node / pixel are two conflict-free LDS.128 at "plan byte offset + immediate";
//   * the stiffness is applied in "edge form": rho_i = sum_j (s0 a[c0] + s1 a[c1]) (u_j - u_i) - f_i,
//     valid because every P1 diffusion element matrix has zero row sums (checked at plan creation;
//     meshes that fail the check use the version-1 kernels).  The per-row edge records of the next tile
//     are also brought into shared memory by cp.async while the current tile's MMAs run, so the
//     matvec itself issues no global load;
//   * the tile's rho[8][128] goes through shared memory into FP64 tensor-core MMAs
//     (mma.sync.m8n8k4.f64: M = 8 samples, K = rows, N = columns of V) whose accumulators stay in
//     registers across all tiles; rho never touches HBM (unless the caller asks for it).
// Algorithmic HBM bytes per sample: s*(n_inputs + d + n_bc + m)  (SURVEY.md 8d).
#pragma once

namespace gpde {

constexpr int kFR = 128;             // rows per tile
constexpr int kFT = 256;             // threads per CTA: (row, half) -> 4 samples per thread
constexpr int kFS = 8;               // samples per CTA == MMA M
constexpr int kFH = 4;               // samples per thread
constexpr int kPitch = 10;           // doubles per ring entry (8 samples + 2 pad -> conflict-free LDS.128)
constexpr int kPitchB = kPitch * 8;  // bytes
constexpr int kRhoPitch = kFR + 4;   // (pitch % 16 == 4) -> conflict-free A-fragment loads

// One neighbour of one row is a pair (int4 offsets, double2 stiffness contributions):
//   off.x = byte offset (from the shared-memory base) of the neighbour's u~ entry -- u ring for a free
//           node, Dirichlet area for a constrained one; off.y / off.z = the two adjacent cells'
//           conductivities in the a ring; off.w = byte offset of the row's own u~ entry;
//   coef  = unit-conductivity stiffness contributions of the two cells.
// Per tile the records are stored contiguously ("tile block") so that they can be copied into shared
// memory with 16-byte cp.async:  [nnb][kFR] int4 | [nnb][kFR] double2 | [kFR] double f
struct TileMeta {            // [n_tiles], copied to shared memory when the kernel starts
    int u_first, u_count, u_off;   // entries of the u ring to fetch for this tile (free indices; offset in entries)
    int a_first, a_count, a_off;   // entries of the a ring to fetch for this tile (input ids)
    int pad0, pad1;
};

// Shared-memory map (bytes from the base):
//   [u ring: y by free index][Dirichlet area: g][a ring][rho tile][record block][tile metadata]
struct VoTiles {
    int ok;            // 0 -> fused path unavailable for this mesh
    int async_ok;      // rings advance monotonically -> next tile can be fetched while this one is used
    int n_tiles, nnb, ring_u, ring_a, n_bc;
    int g_base, a_base, rs_base, rec_base, meta_base;
    int block_bytes;                        // bytes of one tile block (multiple of 16)
    const TileMeta *meta;                   // [n_tiles]
    const char *blocks;                     // [n_tiles][block_bytes]
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__constant__ double kExpC[13] = {
    2.08767569878680989792e-09, 2.50521083854417187751e-08, 2.75573192239858906526e-07,
    2.75573192239858906526e-06, 2.48015873015873015873e-05, 1.98412698412698412698e-04,
    1.38888888888888888889e-03, 8.33333333333333333333e-03, 4.16666666666666666667e-02,
    1.66666666666666666667e-01, 0.5, 1.0, 1.0};

// exp(x) to ~4e-16 relative for |x| <= 700: 2^k * P12(r), r = x - k ln2 in [-0.35, 0.35]; anything else
// (huge, inf, NaN) takes the libm path.  Half the instructions of the libm exp; the parity tolerance on
// this path is 1e-10.
__device__ __forceinline__ double fast_exp(double x) {
    if (!(fabs(x) <= 700.0)) return exp(x);
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);   // 1.5*2^52: low word = rint(x log2 e)
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = kExpC[0];
#pragma unroll
    for (int c = 1; c < 13; ++c) p = fma(p, r, kExpC[c]);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

__device__ __forceinline__ void lds4(const char *p, double (&v)[kFH]) {
    const double2 q0 = reinterpret_cast<const double2 *>(p)[0], q1 = reinterpret_cast<const double2 *>(p)[1];
    v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y;
}

__device__ __forceinline__ void sts4(char *p, const double (&v)[kFH]) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(v[0], v[1]);
    reinterpret_cast<double2 *>(p)[1] = make_double2(v[2], v[3]);
}

__device__ __forceinline__ int ring_wrap(int off, int ring) {
    while (off >= ring) off -= ring;
    return off;
}

// Copy tile t's record block into shared memory (asynchronously).
__device__ __forceinline__ void fetch_records(const VoTiles &Tl, int t, char *dst) {
    const char *src = Tl.blocks + (size_t)t * Tl.block_bytes;
    for (int o = threadIdx.x * 16; o < Tl.block_bytes; o += kFT * 16) cp_async16(dst + o, src + o);
}

// Asynchronous fetch of tile t's new ring entries for this thread's 4 samples (FP64 inputs only: the
// bytes land in their final place; a log-field is exponentiated later by convert_inputs()).
__device__ __forceinline__ void stage_async(const VoTiles &Tl, const TileMeta &mt, int k0,
                                            const double *const (&ap)[kFH], const double *const (&yp)[kFH],
                                            bool have_y, char *smh) {
    for (int k = k0; k < mt.a_count; k += kFR) {
        char *dst = smh + Tl.a_base + ring_wrap(mt.a_off + k, Tl.ring_a) * kPitchB;
#pragma unroll
        for (int s = 0; s < kFH; ++s) cp_async8(dst + 8 * s, ap[s] + mt.a_first + (k - k0));
    }
    for (int k = k0; k < mt.u_count; k += kFR) {
        char *dst = smh + ring_wrap(mt.u_off + k, Tl.ring_u) * kPitchB;
        if (have_y) {
#pragma unroll
            for (int s = 0; s < kFH; ++s) cp_async8(dst + 8 * s, yp[s] + mt.u_first + (k - k0));
        } else {
            const double z[kFH] = {0.0, 0.0, 0.0, 0.0};
            sts4(dst, z);
        }
    }
}

// exp() in place on the a-ring entries this thread fetched for tile t (after its cp.async completed).
__device__ __forceinline__ void convert_inputs(const VoTiles &Tl, const TileMeta &mt, int k0, char *smh) {
    for (int k = k0; k < mt.a_count; k += kFR) {
        char *p = smh + Tl.a_base + ring_wrap(mt.a_off + k, Tl.ring_a) * kPitchB;
        double v[kFH];
        lds4(p, v);
#pragma unroll
        for (int s = 0; s < kFH; ++s) v[s] = fast_exp(v[s]);
        sts4(p, v);
    }
}

// Synchronous staging of tile t (FP32 inputs, or meshes whose rings do not advance monotonically).
template <typename T>
__device__ __forceinline__ void stage_sync(const VoTiles &Tl, const TileMeta &mt, int k0, int a_is_log,
                                           const T *const (&ap)[kFH], const T *const (&yp)[kFH], bool have_y,
                                           char *smh) {
    for (int k = k0; k < mt.a_count; k += kFR) {
        double v[kFH];
#pragma unroll
        for (int s = 0; s < kFH; ++s) v[s] = ldd(ap[s] + mt.a_first + (k - k0));
        if (a_is_log) {
#pragma unroll
            for (int s = 0; s < kFH; ++s) v[s] = fast_exp(v[s]);
        }
        sts4(smh + Tl.a_base + ring_wrap(mt.a_off + k, Tl.ring_a) * kPitchB, v);
    }
    for (int k = k0; k < mt.u_count; k += kFR) {
        double v[kFH];
#pragma unroll
        for (int s = 0; s < kFH; ++s) v[s] = have_y ? ldd(yp[s] + mt.u_first + (k - k0)) : 0.0;
        sts4(smh + ring_wrap(mt.u_off + k, Tl.ring_u) * kPitchB, v);
    }
}

// rho for this thread's row (k0 within the tile) and 4 samples, edge form, records read from shared memory.
// smh = shared base + this thread's sample-half offset; rec = this tile's record block.
__device__ __forceinline__ void apply_row(const VoTiles &Tl, int k0, const char *smh, const char *rec, int sub_f,
                                          double (&acc)[kFH]) {
    const int4 *ro = reinterpret_cast<const int4 *>(rec) + k0;                                   // [nnb][kFR]
    const double2 *rc = reinterpret_cast<const double2 *>(rec + (size_t)Tl.nnb * kFR * 16) + k0;  // [nnb][kFR]
    const double f = sub_f ? reinterpret_cast<const double *>(rec + (size_t)Tl.nnb * kFR * 32)[k0] : 0.0;
    int4 o = ro[0];
    double2 sc = rc[0];
    double ui[kFH];
    lds4(smh + o.w, ui);
#pragma unroll
    for (int s = 0; s < kFH; ++s) acc[s] = -f;
    for (int nb = 0; nb < Tl.nnb; ++nb) {
        double uj[kFH], a0[kFH], a1[kFH];
        lds4(smh + o.x, uj);
        lds4(smh + o.y, a0);
        lds4(smh + o.z, a1);
        const double s0 = sc.x, s1 = sc.y;
        if (nb + 1 < Tl.nnb) {
            o = ro[(nb + 1) * kFR];
            sc = rc[(nb + 1) * kFR];
        }
#pragma unroll
        for (int s = 0; s < kFH; ++s) {
            const double kij = fma(s1, a1[s], s0 * a0[s]);
            acc[s] = fma(kij, uj[s] - ui[s], acc[s]);
        }
    }
}

// Sample rows of this thread (4 samples), offset by k0; samples past the end of the batch alias the last
// valid one (their results are never stored).
template <typename T>
__device__ __forceinline__ void sample_rows(const T *base, long long stride, long long b0, long long B, int h,
                                            int k0, const T *(&out)[kFH]) {
#pragma unroll
    for (int s = 0; s < kFH; ++s) {
        long long b = b0 + h * kFH + s;
        if (b >= B) b = B - 1;
        out[s] = base ? base + b * stride + k0 : nullptr;
    }
}

// Dirichlet values of the CTA's samples -> [n_bc][kPitch] area, tile metadata -> shared (once per CTA).
template <typename T>
__device__ __forceinline__ void stage_constants(const VoTiles &Tl, const T *g, long long g_stride, long long b0,
                                                long long B, char *sm) {
    for (int idx = threadIdx.x; idx < Tl.n_bc * kFS; idx += kFT) {
        const int c = idx / kFS, s = idx - c * kFS;
        long long b = b0 + s;
        if (b >= B) b = B - 1;
        reinterpret_cast<double *>(sm + Tl.g_base + c * kPitchB)[s] = g ? ldd(g + b * g_stride + c) : 0.0;
    }
    int *meta = reinterpret_cast<int *>(sm + Tl.meta_base);
    for (int idx = threadIdx.x; idx < Tl.n_tiles * (int)(sizeof(TileMeta) / 4); idx += kFT)
        meta[idx] = reinterpret_cast<const int *>(Tl.meta)[idx];
}

template <typename T, bool ASYNC>
struct Stager;
template <>
struct Stager<double, true> {
    static __device__ __forceinline__ void ahead(const VoTiles &Tl, const TileMeta &mt, int k0,
                                                 const double *const (&ap)[kFH], const double *const (&yp)[kFH],
                                                 bool have_y, char *smh) {
        stage_async(Tl, mt, k0, ap, yp, have_y, smh);
    }
    static __device__ __forceinline__ void now(const VoTiles &Tl, const TileMeta &mt, int k0, int a_is_log,
                                               const double *const (&)[kFH], const double *const (&)[kFH], bool,
                                               char *smh) {
        if (a_is_log) convert_inputs(Tl, mt, k0, smh);
    }
};
template <typename T>
struct Stager<T, false> {
    static __device__ __forceinline__ void ahead(const VoTiles &, const TileMeta &, int, const T *const (&)[kFH],
                                                 const T *const (&)[kFH], bool, char *) {}
    static __device__ __forceinline__ void now(const VoTiles &Tl, const TileMeta &mt, int k0, int a_is_log,
                                               const T *const (&ap)[kFH], const T *const (&yp)[kFH], bool have_y,
                                               char *smh) {
        stage_sync<T>(Tl, mt, k0, a_is_log, ap, yp, have_y, smh);
    }
};

// WN = number of 8-column tiles of V handled side by side (1, 2 or 4 -> m <= 8, 16, 32);
// the 8 warps form a (8/WN) x WN grid over (row slices of the tile) x (column tiles).
// ASYNC: next tile's ring entries are fetched with cp.async while this tile is processed.
template <typename T, int WN, bool ASYNC>
__global__ void __launch_bounds__(kFT, 2)
vo_fused_kernel(VoDev P, VoTiles Tl, const T *__restrict__ a, long long a_stride, int a_is_log,
                const T *__restrict__ y, const T *__restrict__ g, long long g_stride, const T *__restrict__ V,
                int m, T *__restrict__ r, T *__restrict__ rho_out, int sub_f, long long B) {
    extern __shared__ double sm[];
    constexpr int WK = 8 / WN;            // warps along the rows of a tile
    constexpr int KS = (kFR / 4) / WK;    // k-steps (4 rows each) per warp per tile
    char *smb = reinterpret_cast<char *>(sm);
    double *rs = reinterpret_cast<double *>(smb + Tl.rs_base);   // [8][kRhoPitch]
    const TileMeta *meta = reinterpret_cast<const TileMeta *>(smb + Tl.meta_base);
    char *recbuf = smb + Tl.rec_base;
    const int d = P.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wn = warp % WN, wk = warp / WN;
    const int k0 = threadIdx.x & (kFR - 1), h = threadIdx.x >> 7;
    char *smh = smb + h * (kFH * 8);               // this thread's sample half inside every ring entry
    const long long b0 = (long long)blockIdx.x * kFS;
    const int col = wn * 8 + (lane >> 2);          // column of V this lane feeds into the B fragment
    const bool col_ok = col < m;

    const T *ap[kFH], *yp[kFH];
    sample_rows<T>(a, a_stride, b0, B, h, k0, ap);
    sample_rows<T>(y, d, b0, B, h, k0, yp);
    const bool have_y = y != nullptr;

    // B-fragment source of this lane: V[(row0 + (wk*KS + j)*4 + (lane&3)) * m + col]
    const int rlane = wk * KS * 4 + (lane & 3);
    const T *vptr = V + (long long)rlane * m + (col_ok ? col : 0);
    const long long vstep = 4LL * m;
    const double *rs_lane = rs + (lane >> 2) * kRhoPitch + rlane;

    fetch_records(Tl, 0, recbuf);
    stage_constants<T>(Tl, g, g_stride, b0, B, smb);
    __syncthreads();
    Stager<T, ASYNC>::ahead(Tl, meta[0], k0, ap, yp, have_y, smh);
    cp_async_commit();

    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;   // two C fragments (even / odd k-steps): shorter MMA chains
    for (int t = 0; t < Tl.n_tiles; ++t, vptr += (long long)kFR * m) {
        const int row0 = t * kFR;
        // (A) tile t's ring entries and records (fetched one tile ago) have landed; finish them
        cp_async_wait_all();
        Stager<T, ASYNC>::now(Tl, meta[t], k0, a_is_log, ap, yp, have_y, smh);
        __syncthreads();
        // (B) start fetching tile t+1's ring entries, then one row per thread pair, 4 samples per thread
        if (t + 1 < Tl.n_tiles) Stager<T, ASYNC>::ahead(Tl, meta[t + 1], k0, ap, yp, have_y, smh);
        cp_async_commit();
        {
            const int i = row0 + k0;
            double acc[kFH];
            if (i < d) {
                apply_row(Tl, k0, smh, recbuf, sub_f, acc);
            } else {
#pragma unroll
                for (int s = 0; s < kFH; ++s) acc[s] = 0.0;
            }
#pragma unroll
            for (int s = 0; s < kFH; ++s) rs[(h * kFH + s) * kRhoPitch + k0] = acc[s];
            if (rho_out && i < d) {
#pragma unroll
                for (int s = 0; s < kFH; ++s)
                    if (b0 + h * kFH + s < B) rho_out[(b0 + h * kFH + s) * d + i] = (T)acc[s];
            }
        }
        // (C) this tile's B fragments (their latency overlaps the barrier)
        double breg[KS];
        if (col_ok && row0 + kFR <= d) {
            const T *q = vptr;
#pragma unroll
            for (int j = 0; j < KS; ++j, q += vstep) breg[j] = ldd(q);
        } else {
#pragma unroll
            for (int j = 0; j < KS; ++j)
                breg[j] = (col_ok && row0 + rlane + j * 4 < d) ? ldd(vptr + j * vstep) : 0.0;
        }
        __syncthreads();
        // (D) records of tile t+1 may now replace tile t's; C[8 x 8] += rho[8 x 4] * V[4 x 8] per k-step
        if (t + 1 < Tl.n_tiles) fetch_records(Tl, t + 1, recbuf);
        cp_async_commit();
#pragma unroll
        for (int j = 0; j < KS; j += 2) {
            dmma884(c0, c1, rs_lane[j * 4], breg[j]);
            dmma884(e0, e1, rs_lane[j * 4 + 4], breg[j + 1]);
        }
    }
    c0 += e0;
    c1 += e1;
    // reduce the WK partial C tiles through shared memory (ring space is free now)
    cp_async_wait_all();
    __syncthreads();
    double *red = sm;    // [WK][8 samples][WN*8 cols]
    {
        const int s = lane >> 2, cc = wn * 8 + 2 * (lane & 3);
        red[(wk * 8 + s) * (WN * 8) + cc] = c0;
        red[(wk * 8 + s) * (WN * 8) + cc + 1] = c1;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 8 * WN * 8; idx += kFT) {
        const int s = idx / (WN * 8), cc = idx - s * (WN * 8);
        if (cc < m && b0 + s < B) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < WK; ++w) v += red[(w * 8 + s) * (WN * 8) + cc];
            r[(b0 + s) * m + cc] = (T)v;
        }
    }
}

// q[b,:] = K_ff(a_b) (V s_b): same rings + edge-form matvec, u~ replaced by V s (0 on Dirichlet nodes),
// no contraction.  ss = the 8 coefficient vectors, [m][8] in shared memory in place of the rho tile.
template <typename T>
__global__ void __launch_bounds__(kFT, 2)
vo_fused_T_kernel(VoDev P, VoTiles Tl, const T *__restrict__ a, long long a_stride, int a_is_log,
                  const T *__restrict__ V, int m, const T *__restrict__ svec, T *__restrict__ q, long long B) {
    extern __shared__ double sm[];
    char *smb = reinterpret_cast<char *>(sm);
    double *ss = reinterpret_cast<double *>(smb + Tl.rs_base);   // [m][8]
    const TileMeta *meta = reinterpret_cast<const TileMeta *>(smb + Tl.meta_base);
    char *recbuf = smb + Tl.rec_base;
    const int d = P.d;
    const int k0 = threadIdx.x & (kFR - 1), h = threadIdx.x >> 7;
    char *smh = smb + h * (kFH * 8);
    const long long b0 = (long long)blockIdx.x * kFS;
    const T *ap[kFH];
    sample_rows<T>(a, a_stride, b0, B, h, k0, ap);
    for (int idx = threadIdx.x; idx < kFS * m; idx += kFT) {
        const int qq = idx / kFS, s = idx - qq * kFS;
        const long long b = (b0 + s < B) ? b0 + s : B - 1;
        ss[idx] = ldd(svec + b * m + qq);
    }
    fetch_records(Tl, 0, recbuf);
    cp_async_commit();
    stage_constants<T>(Tl, (const T *)nullptr, 0, b0, B, smb);
    __syncthreads();
    for (int t = 0; t < Tl.n_tiles; ++t) {
        const TileMeta mt = meta[t];
        for (int k = k0; k < mt.a_count; k += kFR) {   // conductivities
            double v[kFH];
#pragma unroll
            for (int s = 0; s < kFH; ++s) {
                v[s] = ldd(ap[s] + mt.a_first + (k - k0));
                if (a_is_log) v[s] = fast_exp(v[s]);
            }
            sts4(smh + Tl.a_base + ring_wrap(mt.a_off + k, Tl.ring_a) * kPitchB, v);
        }
        for (int k = k0; k < mt.u_count; k += kFR) {   // u~ = V s on the free nodes
            double acc[kFH] = {0.0, 0.0, 0.0, 0.0};
            const T *vrow = V + (long long)(mt.u_first + k) * m;
            for (int qq = 0; qq < m; ++qq) {
                const double v = ldd(vrow + qq);
#pragma unroll
                for (int s = 0; s < kFH; ++s) acc[s] = fma(v, ss[qq * kFS + h * kFH + s], acc[s]);
            }
            sts4(smh + ring_wrap(mt.u_off + k, Tl.ring_u) * kPitchB, acc);
        }
        cp_async_wait_all();
        __syncthreads();
        const int i = t * kFR + k0;
        if (i < d) {
            double acc[kFH];
            apply_row(Tl, k0, smh, recbuf, 0, acc);
#pragma unroll
            for (int s = 0; s < kFH; ++s)
                if (b0 + h * kFH + s < B) q[(b0 + h * kFH + s) * d + i] = (T)acc[s];
        }
        __syncthreads();
        if (t + 1 < Tl.n_tiles) fetch_records(Tl, t + 1, recbuf);
        cp_async_commit();
    }
}

// ------------------------------------------------------------------------------------------- host side
struct RangeRing {
    int ring = 0;
    std::vector<int> first, count, off;   // per tile
    int mode = 0;                          // 0 full restage, +1 increasing, -1 decreasing
    std::vector<int> lo;
    int offset_of(int t, int id) const { return mode == 0 ? id - lo[t] : id % ring; }
};

// Ring able to hold the entries of two consecutive tiles (tile t in use while tile t+1 is fetched).
static RangeRing build_ring(const std::vector<int> &lo, const std::vector<int> &hi) {
    RangeRing R;
    const int nt = (int)lo.size();
    R.lo = lo;
    bool inc = true, dec = true;
    int maxlen = 0;
    for (int t = 0; t < nt; ++t) {
        maxlen = std::max(maxlen, hi[t] - lo[t]);
        if (t > 0) {
            if (lo[t] < lo[t - 1] || hi[t] < hi[t - 1]) inc = false;
            if (lo[t] > lo[t - 1] || hi[t] > hi[t - 1]) dec = false;
            maxlen = std::max(maxlen, std::max(hi[t], hi[t - 1]) - std::min(lo[t], lo[t - 1]));
        }
    }
    R.ring = std::max(maxlen, 1);
    R.mode = inc ? 1 : (dec ? -1 : 0);
    R.first.resize(nt); R.count.resize(nt); R.off.resize(nt);
    for (int t = 0; t < nt; ++t) {
        int f = lo[t], e = hi[t];
        if (t > 0 && R.mode == 1) f = std::max(lo[t], hi[t - 1]);
        if (t > 0 && R.mode == -1) e = std::min(hi[t], lo[t - 1]);
        R.first[t] = f;
        R.count[t] = std::max(0, e - f);
        R.off[t] = R.mode == 0 ? 0 : f % R.ring;
    }
    return R;
}

}  // namespace gpde
