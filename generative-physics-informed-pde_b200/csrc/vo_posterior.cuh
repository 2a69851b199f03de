// Batched Gaussian conditioning of the virtual observables, matrix-free (sm_100a).  Included by vo.cu.
//
// Reference (per data point, in Python loops over dense Gamma = V^T K on the CPU/GPU):
//   VirtualObservable.update                 bottleneck/VirtualObservables.py:642-669
//       Lambda = Gamma C Gamma^T + diag(sigma^2),  C = diag(1 / prec)
//       mean   = g - C Gamma^T Lambda^-1 (Gamma g - alpha)
//       vars   = diag(C) - diag(C Gamma^T Lambda^-1 Gamma C)
//   VirtualObservablesEnsemble.update_vo_precision   :971-998
//       beta = 1/2 sum_n [(Gamma_n mean_n - alpha_n)^2 + Gamma_n^2 vars_n] + beta_0
//
// Here Gamma is never stored.  Row i of Gamma^T = K_ff(a) V is  q_i = sum_{j ~ i} K_ij(a) V[j,:]  (<= 7 neighbours on a P1
// mesh), with K_ij(a) = sum_t a[in_t] coef_t from the plan's CSR term lists.  One CTA owns one data point and streams over
// tiles of 32 free rows:
//   pass A   Q tile -> shared memory;  Lambda += sum_r c_r q_r q_r^T   (register blocks of the m x m matrix, 256 threads)
//            resid += sum_r V[r,:] rho_r   with rho = K_fom(a) g~ - f the fine residual of the prior mean (one rho-only
//            launch of gpde_vo_residual before this one): resid = V^T rho = Gamma g - alpha
//   middle   Lambda += diag(sigma^2); Cholesky, explicit inverse and sol = Lambda^-1 resid in shared memory
//   pass B   Q tile again (7 m FMAs per row, cheaper than keeping d x m doubles per data point anywhere);
//            mean_r = g_r - c_r q_r . sol,   vars_r = c_r - c_r^2 q_r^T Lambda^-1 q_r
// moments:   r[n,:] = V_n^T rho_n (= Gamma_n mean_n - alpha_n for rho = rho(mean_n)),  s2[n,j] = sum_i q_i[j]^2 v_i
//            (the two terms of the precision hyper-update) with the same tiles.
// V is shared by the data points (v_stride = 0: V = W of the coarse-grained-residual sampler) or one matrix per data point.
#pragma once

namespace gpde {

struct VoCsrDev {
    int d, nnz, max_row;        // free rows, stored entries of K_ff, most entries in one row
    const int *row_ptr;         // [d+1]
    const int *col;             // [nnz] free index of the neighbour (own row included)
    const int *term_ptr;        // [nnz+1]
    const int *term_in;         // conductivity input of the term's cell
    const double *term_coef;    // element-matrix entry
};

constexpr int kPostThreads = 256, kPostRows = 32;

// K_ij(a) of the tile's rows -> vals_s (CSR order, relative to the tile's first entry)
__device__ __forceinline__ void post_tile_values(const VoCsrDev &C, const double *__restrict__ a, int row0, int rows,
                                                 double *vals_s) {
    const int k0 = C.row_ptr[row0], k1 = C.row_ptr[row0 + rows];
    for (int k = k0 + threadIdx.x; k < k1; k += kPostThreads) {
        double v = 0.0;
        const int t1 = C.term_ptr[k + 1];
        for (int t = C.term_ptr[k]; t < t1; ++t) v = fma(__ldg(a + C.term_in[t]), C.term_coef[t], v);
        vals_s[k - k0] = v;
    }
}

// q_r[c] for the thread's row r = tid / 8 and columns c = tid % 8 + 8 j  (j < MP / 8)
template <int MP>
__device__ __forceinline__ void post_tile_rows(const VoCsrDev &C, const double *__restrict__ V, int m, int row0, int rows,
                                               const double *vals_s, double (&q)[MP / 8]) {
    const int r = threadIdx.x >> 3, cg = threadIdx.x & 7;
#pragma unroll
    for (int j = 0; j < MP / 8; ++j) q[j] = 0.0;
    if (r < rows) {
        const int kbase = C.row_ptr[row0];
        const int k1 = C.row_ptr[row0 + r + 1];
        for (int k = C.row_ptr[row0 + r]; k < k1; ++k) {
            const double v = vals_s[k - kbase];
            const double *vr = V + (long long)C.col[k] * m + cg;
#pragma unroll
            for (int j = 0; j < MP / 8; ++j)
                if (cg + 8 * j < m) q[j] = fma(v, __ldg(vr + 8 * j), q[j]);
        }
    }
}

// shared memory (doubles): Lam [MP][MP+1] | Linv [MP][MP+1] | Qs [32][MP+2] | Qc [32][MP+2] | vals [32 * max_row] | sol [MP] | red
template <int MP>
__global__ void __launch_bounds__(kPostThreads)
vo_posterior_kernel(VoCsrDev C, const double *__restrict__ a, long long a_stride, const double *__restrict__ V,
                    long long v_stride, int m, const double *__restrict__ rho, const double *__restrict__ noise_var,
                    const double *__restrict__ g, const double *__restrict__ prec, double *__restrict__ mean,
                    double *__restrict__ vars, int *info) {
    extern __shared__ __align__(16) double post_smem[];
    constexpr int LP = MP + 1, QP = MP + 2, BS = MP / 16;
    double *Lam = post_smem;
    double *Linv = Lam + MP * LP;
    double *Qs = Linv + MP * LP;
    double *Qc = Qs + kPostRows * QP;
    double *vals_s = Qc + kPostRows * QP;
    double *sol = vals_s + kPostRows * C.max_row;
    const int tid = threadIdx.x;
    const long long n = blockIdx.x;
    const double *an = a + n * a_stride, *Vn = V + n * v_stride;
    const double *gn = g + n * C.d, *pn = prec + n * C.d, *rn = rho + n * C.d;
    const int d = C.d;
    const int r = tid >> 3, cg = tid & 7;
    const int bi = tid >> 4, bj = tid & 15;

    // ---- pass A: Lambda = sum_i c_i q_i q_i^T
    double acc[BS][BS];
#pragma unroll
    for (int x = 0; x < BS; ++x)
#pragma unroll
        for (int y = 0; y < BS; ++y) acc[x][y] = 0.0;
    double rp[MP / 8];          // partial sums of resid = V^T rho over the thread's rows
#pragma unroll
    for (int j = 0; j < MP / 8; ++j) rp[j] = 0.0;
    for (int row0 = 0; row0 < d; row0 += kPostRows) {
        const int rows = min(kPostRows, d - row0);
        __syncthreads();
        post_tile_values(C, an, row0, rows, vals_s);
        __syncthreads();
        double q[MP / 8];
        post_tile_rows<MP>(C, Vn, m, row0, rows, vals_s, q);
        const double cr = r < rows ? 1.0 / pn[row0 + r] : 0.0;
        if (r < rows) {
            const double rr = rn[row0 + r];
            const double *vr = Vn + (long long)(row0 + r) * m + cg;
#pragma unroll
            for (int j = 0; j < MP / 8; ++j)
                if (cg + 8 * j < m) rp[j] = fma(__ldg(vr + 8 * j), rr, rp[j]);
        }
#pragma unroll
        for (int j = 0; j < MP / 8; ++j) {
            Qs[r * QP + cg + 8 * j] = q[j];
            Qc[r * QP + cg + 8 * j] = cr * q[j];
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < kPostRows; ++rr) {
            double qa[BS], qb[BS];
#pragma unroll
            for (int x = 0; x < BS; ++x) {
                qa[x] = Qc[rr * QP + BS * bi + x];
                qb[x] = Qs[rr * QP + BS * bj + x];
            }
#pragma unroll
            for (int x = 0; x < BS; ++x)
#pragma unroll
                for (int y = 0; y < BS; ++y) acc[x][y] = fma(qa[x], qb[y], acc[x][y]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int x = 0; x < BS; ++x)
#pragma unroll
        for (int y = 0; y < BS; ++y) {
            const int i = BS * bi + x, j = BS * bj + y;
            double v = acc[x][y];
            if (i == j) v += i < m ? noise_var[i] : 1.0;     // padding rows / columns: identity
            Lam[i * LP + j] = v;
        }
#pragma unroll
    for (int j = 0; j < MP / 8; ++j) Qs[r * QP + cg + 8 * j] = rp[j];     // Qs is free: reduce resid over the 32 row slots
    __syncthreads();
    double *resid_s = Qc;
    for (int c = tid; c < MP; c += kPostThreads) {
        double sum = 0.0;
        for (int rr = 0; rr < kPostRows; ++rr) sum += Qs[rr * QP + c];
        resid_s[c] = sum;
    }
    __syncthreads();

    // ---- Cholesky (lower, in place), inverse of L, Lambda^-1 = L^-T L^-1 and sol = Lambda^-1 resid: warp 0, m <= 64
    if (tid < 32) {
        const int lane = tid;
        int bad = 0;
        for (int k = 0; k < MP; ++k) {
            const double dkk = Lam[k * LP + k];
            if (!(dkk > 0.0)) bad = 1;
            const double lkk = sqrt(dkk);
            __syncwarp();
            for (int i = k + lane; i < MP; i += 32) Lam[i * LP + k] = (i == k) ? lkk : Lam[i * LP + k] / lkk;
            __syncwarp();
            for (int j = k + 1 + lane; j < MP; j += 32) {       // column j of the trailing block (rows >= j)
                const double ljk = Lam[j * LP + k];
                for (int i = j; i < MP; ++i) Lam[i * LP + j] -= Lam[i * LP + k] * ljk;
            }
            __syncwarp();
        }
        if (bad && info) atomicOr(info, GPDE_INFO_NOT_SPD);
        // Linv = L^-1 (lower): column c solved by lane c (forward substitution)
        for (int c = lane; c < MP; c += 32) {
            for (int i = 0; i < MP; ++i) {
                double s = (i == c) ? 1.0 : 0.0;
                for (int k = c; k < i; ++k) s -= Lam[i * LP + k] * Linv[k * LP + c];
                Linv[i * LP + c] = i < c ? 0.0 : s / Lam[i * LP + i];
            }
        }
        __syncwarp();
        // Lam <- Lambda^-1 = Linv^T Linv (symmetric, full)
        for (int idx = lane; idx < MP * MP; idx += 32) {
            const int i = idx / MP, j = idx - i * MP;
            double s = 0.0;
            for (int k = max(i, j); k < MP; ++k) s = fma(Linv[k * LP + i], Linv[k * LP + j], s);
            Lam[i * LP + j] = s;
        }
        __syncwarp();
        for (int i = lane; i < MP; i += 32) {
            double s = 0.0;
            for (int j = 0; j < m; ++j) s = fma(Lam[i * LP + j], resid_s[j], s);
            sol[i] = i < m ? s : 0.0;
        }
    }
    __syncthreads();

    // ---- pass B: posterior mean and variances, row by row
    for (int row0 = 0; row0 < d; row0 += kPostRows) {
        const int rows = min(kPostRows, d - row0);
        __syncthreads();
        post_tile_values(C, an, row0, rows, vals_s);
        __syncthreads();
        double q[MP / 8];
        post_tile_rows<MP>(C, Vn, m, row0, rows, vals_s, q);
#pragma unroll
        for (int j = 0; j < MP / 8; ++j) Qs[r * QP + cg + 8 * j] = q[j];
        __syncwarp();     // the 8 threads of a row sit in one warp
        // t = (Lambda^-1 q)[own columns]; quad = q . t; lin = q . sol
        double quad = 0.0, lin = 0.0;
#pragma unroll
        for (int j = 0; j < MP / 8; ++j) {
            const int c = cg + 8 * j;
            double t = 0.0;
            for (int k = 0; k < m; ++k) t = fma(Lam[c * LP + k], Qs[r * QP + k], t);
            quad = fma(q[j], t, quad);
            lin = fma(q[j], sol[c], lin);
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            quad += __shfl_xor_sync(0xffffffffu, quad, o);
            lin += __shfl_xor_sync(0xffffffffu, lin, o);
        }
        if (cg == 0 && r < rows) {
            const double cr = 1.0 / pn[row0 + r];
            mean[n * d + row0 + r] = gn[row0 + r] - cr * lin;
            vars[n * d + row0 + r] = cr - cr * cr * quad;
        }
    }
}

// r[n, j] = sum_i V_n[i, j] rho[n, i],   s2[n, j] = sum_i q_i[j]^2 v[n, i]
template <int MP>
__global__ void __launch_bounds__(kPostThreads)
vo_moments_kernel(VoCsrDev C, const double *__restrict__ a, long long a_stride, const double *__restrict__ V,
                  long long v_stride, int m, const double *__restrict__ rho, const double *__restrict__ v,
                  double *__restrict__ out_r, double *__restrict__ out_s2) {
    extern __shared__ __align__(16) double post_smem[];
    double *vals_s = post_smem;
    double *red = vals_s + kPostRows * C.max_row;      // [2][32][MP]
    const int tid = threadIdx.x, r = tid >> 3, cg = tid & 7;
    const long long n = blockIdx.x;
    const double *an = a + n * a_stride, *Vn = V + n * v_stride, *vn = v + n * C.d, *rn = rho + n * C.d;
    double part[MP / 8], rp[MP / 8];
#pragma unroll
    for (int j = 0; j < MP / 8; ++j) part[j] = rp[j] = 0.0;
    for (int row0 = 0; row0 < C.d; row0 += kPostRows) {
        const int rows = min(kPostRows, C.d - row0);
        __syncthreads();
        post_tile_values(C, an, row0, rows, vals_s);
        __syncthreads();
        double q[MP / 8];
        post_tile_rows<MP>(C, Vn, m, row0, rows, vals_s, q);
        if (r < rows) {
            const double w = vn[row0 + r], rr = rn[row0 + r];
            const double *vr = Vn + (long long)(row0 + r) * m + cg;
#pragma unroll
            for (int j = 0; j < MP / 8; ++j) {
                part[j] = fma(q[j] * q[j], w, part[j]);
                if (cg + 8 * j < m) rp[j] = fma(__ldg(vr + 8 * j), rr, rp[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < MP / 8; ++j) {
        red[r * MP + cg + 8 * j] = part[j];
        red[(kPostRows + r) * MP + cg + 8 * j] = rp[j];
    }
    __syncthreads();
    for (int c = tid; c < 2 * m; c += kPostThreads) {
        const int which = c >= m, cc = c - which * m;
        double s = 0.0;
        for (int rr = 0; rr < kPostRows; ++rr) s += red[(which * kPostRows + rr) * MP + cc];
        (which ? out_r : out_s2)[n * m + cc] = s;
    }
}

template <int MP>
static inline size_t post_smem_bytes(int max_row) {
    return sizeof(double) * ((size_t)2 * MP * (MP + 1) + 2 * kPostRows * (MP + 2) + (size_t)kPostRows * max_row + MP);
}
template <int MP>
static inline size_t moments_smem_bytes(int max_row) {
    return sizeof(double) * ((size_t)kPostRows * max_row + (size_t)2 * kPostRows * MP);
}

}  // namespace gpde
