// Fused virtual-observable residual kernel (version 2).  Included by vo.cu.
//
// One launch computes r[b,:] = V^T (K_fom(a_b) u~_b - f)_free for all samples:
//   * a CTA owns S=8 samples and walks over tiles of R=256 consecutive free rows;
//   * per tile, the slice of u~ (y | Dirichlet values) and of the conductivity field that the tile's
//     rows touch is staged ONCE per sample into shared-memory ring buffers (exp() applied while
//     staging when the input is a log-field) -- only the part not already resident from the
//     previous tile is (re)loaded, so every input byte is read from HBM once and exp() is
//     evaluated once per input value;
//   * the stiffness is applied in "edge form": rho_i = sum_j (s0 a[c0] + s1 a[c1]) (u_j - u_i) - f_i,
//     valid because every P1 diffusion element matrix has zero row sums (checked at plan creation;
//     meshes that fail the check use the version-1 kernels); the per-row edge data is read once per
//     tile and reused for the S samples;
//   * the tile's rho[S][R] goes through shared memory into FP64 tensor-core MMAs
//     (mma.sync.m8n8k4.f64: M = 8 samples, K = rows, N = columns of V) whose accumulators stay in
//     registers across all tiles; rho never touches HBM (unless the caller asks for it).
// Algorithmic HBM bytes per sample: s*(n_inputs + d + n_bc + m)  (SURVEY.md 8d).
#pragma once

namespace gpde {

constexpr int kFR = 256;        // rows per tile == threads per CTA
constexpr int kFS = 8;          // samples per CTA == MMA M
constexpr int kRhoPitch = kFR + 4;   // (pitch % 16 == 4) -> conflict-free A-fragment loads

struct VoTiles {
    int ok;            // 0 -> fused path unavailable for this mesh
    int n_tiles, nnb, ring_u, ring_a;
    const int *u_first, *u_count, *u_off;   // [n_tiles] staging lists (node ids)
    const int *a_first, *a_count, *a_off;   // [n_tiles] staging lists (input ids)
    const int *node_src;                    // [n_nodes] >=0: y index, <0: -(g index)-1
    const int *row_u;                       // [d]        ring offset of the row's own value
    const int *nb_u, *nb_a0, *nb_a1;        // [nnb*d]    ring offsets (slot-major)
    const double *nb_s0, *nb_s1;            // [nnb*d]
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Stage the new part of the conductivity ring for the S samples of this CTA.
template <typename T>
__device__ __forceinline__ void stage_inputs(const VoTiles &Tl, int t, const T *__restrict__ a, long long a_stride,
                                             int a_is_log, long long b0, long long B, double *as) {
    const int first = Tl.a_first[t], count = Tl.a_count[t], off0 = Tl.a_off[t], ring = Tl.ring_a;
    for (int k = threadIdx.x; k < count; k += kFR) {
        int off = off0 + k;
        if (off >= ring) off -= ring;
#pragma unroll
        for (int s = 0; s < kFS; ++s) {
            const long long b = min(b0 + s, B - 1);
            double v = ldd(a + b * a_stride + first + k);
            as[s * ring + off] = a_is_log ? exp(v) : v;
        }
    }
}

// Stage the new part of the u~ ring (y on free nodes, g on Dirichlet nodes).
template <typename T>
__device__ __forceinline__ void stage_solution(const VoTiles &Tl, int t, const T *__restrict__ y,
                                               const T *__restrict__ g, long long g_stride, int d, long long b0,
                                               long long B, double *us) {
    const int first = Tl.u_first[t], count = Tl.u_count[t], off0 = Tl.u_off[t], ring = Tl.ring_u;
    for (int k = threadIdx.x; k < count; k += kFR) {
        int off = off0 + k;
        if (off >= ring) off -= ring;
        const int src = Tl.node_src[first + k];
#pragma unroll
        for (int s = 0; s < kFS; ++s) {
            const long long b = min(b0 + s, B - 1);
            double v = 0.0;
            if (src >= 0) {
                if (y) v = ldd(y + b * d + src);
            } else if (g) {
                v = ldd(g + b * g_stride + (-src - 1));
            }
            us[s * ring + off] = v;
        }
    }
}

// Transposed flavour: u~ = V s_b on free nodes, 0 on Dirichlet nodes.  ss = s vectors [S][m] in smem.
template <typename T>
__device__ __forceinline__ void stage_expanded(const VoTiles &Tl, int t, const T *__restrict__ V, int m,
                                               const double *ss, double *us) {
    const int first = Tl.u_first[t], count = Tl.u_count[t], off0 = Tl.u_off[t], ring = Tl.ring_u;
    for (int k = threadIdx.x; k < count; k += kFR) {
        int off = off0 + k;
        if (off >= ring) off -= ring;
        const int src = Tl.node_src[first + k];
        double acc[kFS];
#pragma unroll
        for (int s = 0; s < kFS; ++s) acc[s] = 0.0;
        if (src >= 0) {
            const T *vrow = V + (long long)src * m;
            for (int q = 0; q < m; ++q) {
                const double v = ldd(vrow + q);
#pragma unroll
                for (int s = 0; s < kFS; ++s) acc[s] = fma(v, ss[s * m + q], acc[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < kFS; ++s) us[s * ring + off] = acc[s];
    }
}

// rho for row i of the S samples (edge form).  Returns through acc[].
__device__ __forceinline__ void apply_row(const VoDev &P, const VoTiles &Tl, int i, const double *us,
                                          const double *as, int sub_f, double (&acc)[kFS]) {
    const int d = P.d;
    const int ou = Tl.row_u[i];
    double ui[kFS];
    const double f = sub_f ? P.f_free[i] : 0.0;
#pragma unroll
    for (int s = 0; s < kFS; ++s) {
        ui[s] = us[s * Tl.ring_u + ou];
        acc[s] = -f;
    }
    for (int nb = 0; nb < Tl.nnb; ++nb) {
        const int k = nb * d + i;
        const int oj = Tl.nb_u[k], o0 = Tl.nb_a0[k], o1 = Tl.nb_a1[k];
        const double s0 = Tl.nb_s0[k], s1 = Tl.nb_s1[k];
#pragma unroll
        for (int s = 0; s < kFS; ++s) {
            const double kij = fma(s1, as[s * Tl.ring_a + o1], s0 * as[s * Tl.ring_a + o0]);
            acc[s] = fma(kij, us[s * Tl.ring_u + oj] - ui[s], acc[s]);
        }
    }
}

// WN = number of 8-column tiles of V handled side by side (1, 2 or 4 -> m <= 8, 16, 32);
// the 8 warps form a (8/WN) x WN grid over (row slices of the tile) x (column tiles).
template <typename T, int WN>
__global__ void __launch_bounds__(kFR, 2)
vo_fused_kernel(VoDev P, VoTiles Tl, const T *__restrict__ a, long long a_stride, int a_is_log,
                const T *__restrict__ y, const T *__restrict__ g, long long g_stride, const T *__restrict__ V,
                int m, T *__restrict__ r, T *__restrict__ rho_out, int sub_f, long long B) {
    extern __shared__ double sm[];
    constexpr int WK = 8 / WN;            // warps along the rows of a tile
    constexpr int KS = (kFR / 4) / WK;    // k-steps (4 rows each) per warp per tile
    double *us = sm;                               // [S][ring_u]
    double *as = us + kFS * Tl.ring_u;             // [S][ring_a]
    double *rs = as + kFS * Tl.ring_a;             // [S][kRhoPitch]
    const int d = P.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wn = warp % WN, wk = warp / WN;
    const long long b0 = (long long)blockIdx.x * kFS;
    const int col = wn * 8 + (lane >> 2);          // column of V this lane feeds into the B fragment

    double c0 = 0.0, c1 = 0.0;                     // C fragment: C[sample = lane>>2][2*(lane&3) + {0,1}]
    for (int t = 0; t < Tl.n_tiles; ++t) {
        const int row0 = t * kFR;
        // (1) B fragments of this tile straight from global/L2 into registers; consumed in (4)
        double breg[KS];
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            const int row = row0 + (wk * KS + j) * 4 + (lane & 3);
            breg[j] = (row < d && col < m) ? ldd(V + (long long)row * m + col) : 0.0;
        }
        // (2) stage what this tile needs and is not resident yet
        stage_inputs<T>(Tl, t, a, a_stride, a_is_log, b0, B, as);
        stage_solution<T>(Tl, t, y, g, g_stride, d, b0, B, us);
        __syncthreads();
        // (3) one row per thread, S samples
        {
            const int i = row0 + threadIdx.x;
            double acc[kFS];
            if (i < d) {
                apply_row(P, Tl, i, us, as, sub_f, acc);
            } else {
#pragma unroll
                for (int s = 0; s < kFS; ++s) acc[s] = 0.0;
            }
#pragma unroll
            for (int s = 0; s < kFS; ++s) rs[s * kRhoPitch + threadIdx.x] = acc[s];
            if (rho_out && i < d) {
#pragma unroll
                for (int s = 0; s < kFS; ++s)
                    if (b0 + s < B) rho_out[(b0 + s) * d + i] = (T)acc[s];
            }
        }
        __syncthreads();
        // (4) C[8 samples x 8 cols] += rho[8 x 4] * V[4 x 8] over this warp's row slice
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            const double av = rs[(lane >> 2) * kRhoPitch + (wk * KS + j) * 4 + (lane & 3)];
            dmma884(c0, c1, av, breg[j]);
        }
    }
    // reduce the WK partial C tiles through shared memory (ring space is free now)
    __syncthreads();
    double *red = sm;    // [WK][8 samples][WN*8 cols]
    {
        const int s = lane >> 2, cc = wn * 8 + 2 * (lane & 3);
        red[(wk * 8 + s) * (WN * 8) + cc] = c0;
        red[(wk * 8 + s) * (WN * 8) + cc + 1] = c1;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 8 * WN * 8; idx += kFR) {
        const int s = idx / (WN * 8), cc = idx - s * (WN * 8);
        if (cc < m && b0 + s < B) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < WK; ++w) v += red[(w * 8 + s) * (WN * 8) + cc];
            r[(b0 + s) * m + cc] = (T)v;
        }
    }
}

// q[b,:] = K_ff(a_b) (V s_b): same staging + edge-form matvec, u~ replaced by V s, no contraction.
template <typename T>
__global__ void __launch_bounds__(kFR, 2)
vo_fused_T_kernel(VoDev P, VoTiles Tl, const T *__restrict__ a, long long a_stride, int a_is_log,
                  const T *__restrict__ V, int m, const T *__restrict__ svec, T *__restrict__ q, long long B) {
    extern __shared__ double sm[];
    double *us = sm;
    double *as = us + kFS * Tl.ring_u;
    double *ss = as + kFS * Tl.ring_a;             // [S][m]
    const int d = P.d;
    const long long b0 = (long long)blockIdx.x * kFS;
    for (int idx = threadIdx.x; idx < kFS * m; idx += kFR) {
        const int s = idx / m, qq = idx - s * m;
        ss[idx] = ldd(svec + min(b0 + s, B - 1) * m + qq);
    }
    __syncthreads();
    for (int t = 0; t < Tl.n_tiles; ++t) {
        stage_inputs<T>(Tl, t, a, a_stride, a_is_log, b0, B, as);
        stage_expanded<T>(Tl, t, V, m, ss, us);
        __syncthreads();
        const int i = t * kFR + threadIdx.x;
        if (i < d) {
            double acc[kFS];
            apply_row(P, Tl, i, us, as, 0, acc);
#pragma unroll
            for (int s = 0; s < kFS; ++s)
                if (b0 + s < B) q[(b0 + s) * d + i] = (T)acc[s];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------- host side
struct RangeRing {
    int ring = 0;
    std::vector<int> first, count, off;   // per tile
    // offset of id at tile t
    int mode = 0;                          // 0 full restage, +1 increasing, -1 decreasing
    std::vector<int> lo;
    int offset_of(int t, int id) const { return mode == 0 ? id - lo[t] : id % ring; }
};

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
