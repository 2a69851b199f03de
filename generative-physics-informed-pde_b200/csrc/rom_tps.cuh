// Coarse-grained model, one THREAD per sample (small systems).  Included by rom.cu.
// EXPERIMENT, opt-in with GPDE_ROM_PATH=tps: measured slower than the cooperative kernels of rom.cu on B200
// (see gpde_rom_plan_create); kept for A/B runs and as the starting point of a register-resident variant.
//
// For the reference's 4x4 coarse mesh (n_free = 15, half bandwidth 3-4: presets highres32 and BASELINE configs
// 1, 2, 4) the whole banded LDL^T fits in the registers of one thread when the loops are unrolled at compile
// time (template <NF, HBW>): no cross-lane cooperation, no barrier per elimination step, no idle lanes -- the
// cooperative kernels of rom.cu spend > 90 % of their instructions on that.  Same maths, same tables
// (bottleneck/ROM.py:59-100 and its autograd):
//   * the CTA's 128 samples are loaded with coalesced reads and transposed through shared memory into
//     per-thread columns (pitch 129 doubles: conflict-free both ways); exp(X)+1e-8 is applied on the way in;
//   * assembly walks the plan's (uniform) contribution tables from shared memory: Ab[p] += coef * x[elem];
//   * factorisation, forward and backward substitution are fully unrolled over the band held in registers;
//   * the factor is stashed TRANSPOSED, factor[p][B] (diagonal slots hold 1/d_k), so that both kernels read /
//     write it coalesced (the stash is opaque to callers: gpde_rom_factor_bytes);
//   * the adjoint reuses it, then forms dL/dX[e] = -sum_t coef * lam[i_t] * u[j_t] (* exp(X)) from
//     shared-memory columns of lam and u.
#pragma once

namespace gpde {

constexpr int kTpsThreads = 128;
constexpr int kTpsPitch = kTpsThreads + 1;   // doubles between consecutive rows of a per-thread column array

static __constant__ double kRomExpTab[16] = {1.0,
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

// exp(x) = 2^e * T[j] * P6(r), |r| <= ln2/32 (~5e-16 relative for |x| <= 700); libm outside that range
__device__ __forceinline__ double rom_exp(double x, const double *tab) {
    if ((__double2hiint(x) & 0x7fffffff) > 0x4085e000) return exp(x);
    const double t = fma(x, 23.083120654223414, 6755399441055744.0);
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

// Coalesced load of rows [b0, b0+128) x [0, width) of a row-major [B, width] array into per-thread columns
// col[j * pitch + thread].  WITH_EXP: x = exp(v) + 1e-8 (components.py:298), dcol receives exp(v) (chain rule);
// returns GPDE_INFO_NONPOSITIVE_X if any loaded conductivity is <= 1e-12 (ROM.py:74-76).
template <typename T>
__device__ __forceinline__ int tps_load(const T *__restrict__ src, int width, long long b0, long long B, double *col,
                                        double *dcol, bool conductivity, int x_is_log, const double *tab) {
    int bad = 0;
    const long long base = b0 * width;
    const long long limit = min((long long)kTpsThreads, B - b0) * width;
    for (long long i = threadIdx.x; i < limit; i += kTpsThreads) {
        const int row = (int)(i / width), j = (int)(i - (long long)row * width);
        double v = ld_as_double(src + base + i);
        if (conductivity) {
            double dv = 1.0;
            if (x_is_log) {
                dv = rom_exp(v, tab);
                v = dv + 1e-8;
            }
            if (!(v > 1e-12)) bad = GPDE_INFO_NONPOSITIVE_X;
            if (dcol) dcol[j * kTpsPitch + row] = dv;
        }
        col[j * kTpsPitch + row] = v;
    }
    return bad;
}

template <typename T>
__device__ __forceinline__ void tps_store(T *__restrict__ dst, int width, long long b0, long long B, const double *col) {
    const long long base = b0 * width;
    const long long limit = min((long long)kTpsThreads, B - b0) * width;
    for (long long i = threadIdx.x; i < limit; i += kTpsThreads) {
        const int row = (int)(i / width), j = (int)(i - (long long)row * width);
        dst[base + i] = (T)col[j * kTpsPitch + row];
    }
}

// Ab[i][s] = A[i][i-s] of K_ff(x) for this thread's sample (xs = its conductivity column)
template <int NF, int HBW>
__device__ __forceinline__ void tps_assemble(const RomDev &P, const double *xs, double (&Ab)[NF][HBW + 1]) {
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int s = 0; s <= HBW; ++s) {
            const int p = i * (HBW + 1) + s;
            double acc = 0.0;
            const int t1 = P.band_ptr[p + 1];
            for (int t = P.band_ptr[p]; t < t1; ++t) acc = fma(P.band_coef[t], xs[P.band_elem[t] * kTpsPitch], acc);
            Ab[i][s] = acc;
        }
}

// In-place LDL^T (column entries stay unscaled); the diagonal slot Ab[k][0] ends up holding 1/d_k;
// z (if WITH_RHS) becomes w = D^-1 L^-1 z
template <int NF, int HBW, bool WITH_RHS>
__device__ __forceinline__ int tps_factor(double (&Ab)[NF][HBW + 1], double (&z)[NF]) {
    int bad = 0;
#pragma unroll
    for (int k = 0; k < NF; ++k) {
        const double d = Ab[k][0];
        if (!(d > 0.0)) bad = GPDE_INFO_NOT_SPD;
        const double invd = fast_rcp(d);
        Ab[k][0] = invd;
#pragma unroll
        for (int si = 1; si <= HBW; ++si) {
            if (k + si < NF) {
                const double ci = Ab[k + si][si] * invd;
#pragma unroll
                for (int sj = 1; sj <= si; ++sj) Ab[k + si][si - sj] = fma(-ci, Ab[k + sj][sj], Ab[k + si][si - sj]);
            }
        }
        if (WITH_RHS) {
            const double wk = z[k] * invd;
#pragma unroll
            for (int s = 1; s <= HBW; ++s)
                if (k + s < NF) z[k + s] = fma(-Ab[k + s][s], wk, z[k + s]);
            z[k] = wk;
        }
    }
    return bad;
}

// w = D^-1 L^-1 z with a stored factor (in place)
template <int NF, int HBW>
__device__ __forceinline__ void tps_forward_subst(const double (&Ab)[NF][HBW + 1], double (&z)[NF]) {
#pragma unroll
    for (int k = 0; k < NF; ++k) {
        const double wk = z[k] * Ab[k][0];
#pragma unroll
        for (int s = 1; s <= HBW; ++s)
            if (k + s < NF) z[k + s] = fma(-Ab[k + s][s], wk, z[k + s]);
        z[k] = wk;
    }
}

// L^T sol = w (in place): sol_k = w_k - dinv_k * sum_s Ab[k+s][s] sol_{k+s}
template <int NF, int HBW>
__device__ __forceinline__ void tps_backward_subst(const double (&Ab)[NF][HBW + 1], double (&z)[NF]) {
#pragma unroll
    for (int k = NF - 1; k >= 0; --k) {
        double acc = 0.0;
#pragma unroll
        for (int s = 1; s <= HBW; ++s)
            if (k + s < NF) acc = fma(Ab[k + s][s], z[k + s], acc);
        z[k] = fma(-Ab[k][0], acc, z[k]);
    }
}

// shared memory: [table arena][exp table 16][columns ...]
template <typename T, int NF, int HBW>
__global__ void __launch_bounds__(kTpsThreads)
rom_tps_forward_kernel(RomDev P0, const T *__restrict__ X, int x_is_log, const T *__restrict__ F, T *__restrict__ u,
                       double *__restrict__ factor, int *info, long long B) {
    extern __shared__ __align__(16) double smem_all[];
    double *smem;
    const RomDev P = stage_tables(P0, smem_all, &smem);
    double *tab = smem;                       // [16]
    double *xs = tab + 16;                    // [E][pitch]
    double *Fs = xs + P.E * kTpsPitch;        // [n][pitch]  F, then u
    if (threadIdx.x < 16) tab[threadIdx.x] = kRomExpTab[threadIdx.x];
    __syncthreads();
    const long long b0 = (long long)blockIdx.x * kTpsThreads;
    int bad = tps_load<T>(X, P.E, b0, B, xs, nullptr, true, x_is_log, tab);
    tps_load<T>(F, P.n, b0, B, Fs, nullptr, false, 0, tab);
    __syncthreads();
    const long long b = b0 + threadIdx.x;
    if (b < B) {
        const double *x = xs + threadIdx.x;
        double *f = Fs + threadIdx.x;
        double Ab[NF][HBW + 1], z[NF];
        tps_assemble<NF, HBW>(P, x, Ab);
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            double acc = f[P.free_dof[i] * kTpsPitch];
            const int t1 = P.rhs_ptr[i + 1];
            for (int t = P.rhs_ptr[i]; t < t1; ++t)
                acc = fma(-P.rhs_coef[t] * x[P.rhs_elem[t] * kTpsPitch], f[P.rhs_dof[t] * kTpsPitch], acc);
            z[i] = acc;
        }
        bad |= tps_factor<NF, HBW, true>(Ab, z);
        tps_backward_subst<NF, HBW>(Ab, z);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[P.free_dof[i] * kTpsPitch] = z[i];
        if (factor) {   // transposed stash: entry p of sample b at factor[p * B + b]; diagonal slots hold 1/d_k
#pragma unroll
            for (int i = 0; i < NF; ++i)
#pragma unroll
                for (int s = 0; s <= HBW; ++s) factor[(long long)(i * (HBW + 1) + s) * B + b] = Ab[i][s];
        }
    }
    if (bad && info) atomicOr(info, bad);
    __syncthreads();
    tps_store<T>(u, P.n, b0, B, Fs);
}

template <typename T, int NF, int HBW>
__global__ void __launch_bounds__(kTpsThreads)
rom_tps_adjoint_kernel(RomDev P0, const T *__restrict__ X, int x_is_log, const T *__restrict__ u,
                       const double *__restrict__ factor, const T *__restrict__ gbar, T *__restrict__ gradX,
                       T *__restrict__ gradF, long long B) {
    extern __shared__ __align__(16) double smem_all[];
    double *smem;
    const RomDev P = stage_tables(P0, smem_all, &smem);
    double *tab = smem;                        // [16]
    double *xs = tab + 16;                     // [E][pitch]  x, then dL/dX
    double *us = xs + P.E * kTpsPitch;         // [n][pitch]
    double *gs = us + P.n * kTpsPitch;         // [n][pitch]  gbar, then dL/dF
    double *ls = gs + P.n * kTpsPitch;         // [NF][pitch] lambda on the free dofs
    if (threadIdx.x < 16) tab[threadIdx.x] = kRomExpTab[threadIdx.x];
    __syncthreads();
    const long long b0 = (long long)blockIdx.x * kTpsThreads;
    tps_load<T>(X, P.E, b0, B, xs, nullptr, true, x_is_log, tab);
    tps_load<T>(u, P.n, b0, B, us, nullptr, false, 0, tab);
    tps_load<T>(gbar, P.n, b0, B, gs, nullptr, false, 0, tab);
    __syncthreads();
    const long long b = b0 + threadIdx.x;
    if (b < B) {
        double *x = xs + threadIdx.x, *uu = us + threadIdx.x, *gg = gs + threadIdx.x;
        double *lam = ls + threadIdx.x;
        double Ab[NF][HBW + 1], z[NF];
        if (factor) {
#pragma unroll
            for (int i = 0; i < NF; ++i)
#pragma unroll
                for (int s = 0; s <= HBW; ++s) Ab[i][s] = factor[(long long)(i * (HBW + 1) + s) * B + b];
        } else {
            tps_assemble<NF, HBW>(P, x, Ab);
            tps_factor<NF, HBW, false>(Ab, z);
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) z[i] = gg[P.free_dof[i] * kTpsPitch];
        tps_forward_subst<NF, HBW>(Ab, z);
        tps_backward_subst<NF, HBW>(Ab, z);
#pragma unroll
        for (int i = 0; i < NF; ++i) lam[i * kTpsPitch] = z[i];
        if (gradF) {   // lambda on constrained rows first (needs gbar there), then the free rows overwrite gs
            for (int c = 0; c < P.n_bc; ++c) {
                double acc = gg[P.bc_dof[c] * kTpsPitch];
                const int t1 = P.cf_ptr[c + 1];
                for (int t = P.cf_ptr[c]; t < t1; ++t)
                    acc = fma(-P.cf_coef[t] * x[P.cf_elem[t] * kTpsPitch], lam[P.cf_free[t] * kTpsPitch], acc);
                gg[P.bc_dof[c] * kTpsPitch] = acc;
            }
#pragma unroll
            for (int i = 0; i < NF; ++i) gg[P.free_dof[i] * kTpsPitch] = z[i];
        }
        for (int e = 0; e < P.E; ++e) {
            double acc = 0.0;
            const int t1 = P.grad_ptr[e + 1];
            for (int t = P.grad_ptr[e]; t < t1; ++t)
                acc = fma(P.grad_coef[t] * lam[P.grad_i[t] * kTpsPitch], uu[P.grad_j[t] * kTpsPitch], acc);
            // chain rule through x = exp(X) + 1e-8: dx/dX = exp(X) = x - 1e-8 (exact to an ulp of x)
            const double dx = x_is_log ? x[e * kTpsPitch] - 1e-8 : 1.0;
            x[e * kTpsPitch] = -acc * dx;
        }
    }
    __syncthreads();
    tps_store<T>(gradX, P.E, b0, B, xs);
    if (gradF) tps_store<T>(gradF, P.n, b0, B, gs);
}

static inline size_t tps_smem_forward(const RomDev &D) {
    return (size_t)D.arena_bytes + sizeof(double) * (16 + (size_t)(D.E + D.n) * kTpsPitch);
}
static inline size_t tps_smem_adjoint(const RomDev &D) {
    return (size_t)D.arena_bytes + sizeof(double) * (16 + (size_t)(D.E + 2 * D.n + D.n_free) * kTpsPitch);
}

}  // namespace gpde
