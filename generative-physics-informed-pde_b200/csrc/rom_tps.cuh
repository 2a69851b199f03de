// Coarse-grained model, one THREAD per sample, structure fixed at compile time.  Included by rom.cu.
//
// For the reference's 4x4 coarse mesh (presets highres32 and BASELINE configs 1, 2, 4: 15 free dofs, half bandwidth 3,
// 32 cells, 25 dofs) the banded LDL^T of K_ff fits in the registers of one thread, so one sample needs no cross-lane
// cooperation, no barrier per elimination step and no index table on its critical path:
//   * everything that depends on the mesh is a SHAPE (template <TpsShape>): loop bounds, band positions and the number
//     of terms per band entry are compile-time constants, every loop is fully unrolled and the band lives in registers;
//   * everything that depends on M (coefficients, which element feeds which entry) sits in a table passed BY VALUE as a
//     __grid_constant__ kernel parameter: the coefficients are constant-bank operands of the DFMAs and the element
//     offsets uniform-register offsets of the shared-memory loads -- no dependent table walk (the first
//     thread-per-sample kernels walked band_ptr -> band_elem -> x through shared memory: a ~100-cycle dependent chain per
//     term, 51 + 39 us for 4096 samples);
//   * the CTA's 128 samples are read with coalesced loads and transposed through shared memory into per-thread columns
//     (pitch 129 doubles: conflict-free both ways), exp(X)+1e-8 (components.py:298) applied on the way in;
//   * no factor stash: the adjoint re-assembles and re-factorises (about 500 instructions per sample) instead of
//     writing and re-reading 480 bytes per sample through HBM;
//   * dL/dX[e] = -sum_t coef_t lam[i_t] u[j_t] (* exp(X)) from shared-memory columns of lam and u, terms per element
//     padded to the shape's TG.
// Same maths as the cooperative kernels of rom.cu (bottleneck/ROM.py:59-100 and its autograd).
// gpde_rom_plan_create checks that M fits the shape (sizes, bandwidth, terms per entry) and otherwise keeps the plan on
// the cooperative kernels.
#pragma once
#include "exp256.cuh"

namespace gpde {

constexpr int kTpsThreads = 128;
constexpr int kTpsPitch = kTpsThreads + 1;   // doubles between consecutive rows of a per-thread column array

// NF free dofs, half bandwidth HBW, E cells, N dofs; terms per off-diagonal band entry / diagonal entry / right-hand-side
// row (couplings to Dirichlet dofs) / element of the gradient
template <int NF_, int HBW_, int E_, int N_, int TO_, int TD_, int TR_, int TG_>
struct TpsShape {
    static constexpr int NF = NF_, HBW = HBW_, E = E_, N = N_, TO = TO_, TD = TD_, TR = TR_, TG = TG_;
};
using TpsShape4x4 = TpsShape<15, 3, 32, 25, 2, 6, 2, 7>;

// element / dof entries are column offsets (index * kTpsPitch); padding terms have coefficient 0 and offset 0
template <class S>
struct TpsFwdTab {
    double off_coef[S::NF * S::HBW * S::TO];    // entry (i, s >= 1) = A[i][i-s], term t: [(i * HBW + s - 1) * TO + t]
    double diag_coef[S::NF * S::TD];
    double rhs_coef[S::NF * S::TR];             // z_i = F[free_i] - sum_t rhs_coef * x[rhs_elem] * F[rhs_dof]
    unsigned short off_elem[S::NF * S::HBW * S::TO];
    unsigned short diag_elem[S::NF * S::TD];
    unsigned short rhs_elem[S::NF * S::TR], rhs_dof[S::NF * S::TR];
    unsigned short free_dof[S::NF];
};
template <class S>
struct TpsAdjTab {
    TpsFwdTab<S> f;
    double grad_coef[S::E * S::TG];             // dL/dx_e = -sum_t grad_coef * lam[grad_i] * u[grad_j]
    unsigned short grad_i[S::E * S::TG], grad_j[S::E * S::TG];   // dof columns (grad_i is a free dof)
};

// Coalesced load of rows [b0, b0+rows) x [0, W) of a row-major [B, W] array: element threadIdx.x + 128 k of the CTA's block
// into raw[k].  All W loads are issued before anything consumes them (one DRAM latency per array, not one per few elements:
// with one or two warps per scheduler nothing else hides it).  Rows past the batch get ``fill``.
template <typename T, int W>
__device__ __forceinline__ void tps_fetch(const T *__restrict__ src, long long b0, int rows, double fill, T (&raw)[W]) {
    const T *p = src + b0 * W;
    const int limit = rows * W;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        const int i = threadIdx.x + k * kTpsThreads;
        raw[k] = i < limit ? p[i] : (T)fill;
    }
}
// ... and their transposition into per-thread columns col[j * pitch + row] (values as loaded).
template <typename T, int W>
__device__ __forceinline__ void tps_scatter(const T (&raw)[W], double *col) {
#pragma unroll
    for (int k = 0; k < W; ++k) {
        const int i = threadIdx.x + k * kTpsThreads;
        const int row = i / W, j = i - row * W;
        col[j * kTpsPitch + row] = (double)raw[k];
    }
}
// Conductivities of this thread's sample, in place in its column: x = exp(v) + 1e-8 when x_is_log (components.py:298), dcol
// receives dx/dv (exp(v), or 1 for conductivity input).  A ROLLED loop: the fully unrolled version of this and of the
// gradient loop made the kernels 60 KB of straight-line code whose instruction fetch stalled the few resident warps
// (no_instruction 1.5 stalled warps per issue, r2b profile).  Returns GPDE_INFO_NONPOSITIVE_X if a conductivity is
// <= 1e-12 (ROM.py:74-76).
template <int E>
__device__ __forceinline__ int tps_conductivities(double *x, double *dx, int x_is_log, unsigned etab) {
    int bad = 0;
#pragma unroll 2
    for (int e = 0; e < E; ++e) {
        double v = x[e * kTpsPitch], dv = 1.0;
        if (x_is_log) {
            dv = exp256_in_range(v) ? exp_tab256c(v, etab) : exp(v);
            v = dv + 1e-8;
            x[e * kTpsPitch] = v;
        }
        if (!(v > 1e-12)) bad = GPDE_INFO_NONPOSITIVE_X;
        if (dx) dx[e * kTpsPitch] = dv;
    }
    return bad;
}

template <typename T, int W>
__device__ __forceinline__ void tps_stage_out(T *__restrict__ dst, long long b0, int rows, const double *col) {
    T *p = dst + b0 * W;
    const int limit = rows * W;
#pragma unroll 2
    for (int i = threadIdx.x; i < limit; i += kTpsThreads) {
        const int row = i / W, j = i - row * W;
        p[i] = (T)col[j * kTpsPitch + row];
    }
}

// Ab[i][s] = A[i][i-s] of K_ff(x) for this thread's sample (x = its conductivity column)
template <class S>
__device__ __forceinline__ void tps_assemble(const TpsFwdTab<S> &tab, const double *x, double (&Ab)[S::NF][S::HBW + 1]) {
#pragma unroll
    for (int i = 0; i < S::NF; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < S::TD; ++t) acc = fma(tab.diag_coef[i * S::TD + t], x[tab.diag_elem[i * S::TD + t]], acc);
        Ab[i][0] = acc;
#pragma unroll
        for (int s = 1; s <= S::HBW; ++s) {
            double off = 0.0;
            if (i - s >= 0) {
#pragma unroll
                for (int t = 0; t < S::TO; ++t) {
                    const int k = (i * S::HBW + s - 1) * S::TO + t;
                    off = fma(tab.off_coef[k], x[tab.off_elem[k]], off);
                }
            }
            Ab[i][s] = off;
        }
    }
}

// In-place LDL^T (column entries stay unscaled); the diagonal slot Ab[k][0] ends up holding 1/d_k;
// z (if WITH_RHS) becomes w = D^-1 L^-1 z
template <class S, bool WITH_RHS>
__device__ __forceinline__ int tps_factor(double (&Ab)[S::NF][S::HBW + 1], double (&z)[S::NF]) {
    constexpr int NF = S::NF, HBW = S::HBW;
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
template <class S>
__device__ __forceinline__ void tps_forward_subst(const double (&Ab)[S::NF][S::HBW + 1], double (&z)[S::NF]) {
#pragma unroll
    for (int k = 0; k < S::NF; ++k) {
        const double wk = z[k] * Ab[k][0];
#pragma unroll
        for (int s = 1; s <= S::HBW; ++s)
            if (k + s < S::NF) z[k + s] = fma(-Ab[k + s][s], wk, z[k + s]);
        z[k] = wk;
    }
}

// L^T sol = w (in place): sol_k = w_k - dinv_k * sum_s Ab[k+s][s] sol_{k+s}
template <class S>
__device__ __forceinline__ void tps_backward_subst(const double (&Ab)[S::NF][S::HBW + 1], double (&z)[S::NF]) {
#pragma unroll
    for (int k = S::NF - 1; k >= 0; --k) {
        double acc = 0.0;
#pragma unroll
        for (int s = 1; s <= S::HBW; ++s)
            if (k + s < S::NF) acc = fma(Ab[k + s][s], z[k + s], acc);
        z[k] = fma(-Ab[k][0], acc, z[k]);
    }
}

// shared memory: [exp table 256][x: E columns][F -> u: N columns]
template <typename T, class S>
// (three CTAs per SM: 166 registers without spills, 12 warps per SM instead of 8 -- 37.8 -> 35.2 us at B = 131072, A/B; the
//  adjoint spills at that budget and stays at two)
__global__ void __launch_bounds__(kTpsThreads, 3)
rom_tps_forward_kernel(const __grid_constant__ TpsFwdTab<S> tab, const T *__restrict__ X, int x_is_log,
                       const T *__restrict__ F, T *__restrict__ u, int *info, long long B) {
    extern __shared__ __align__(16) double tps_smem[];
    double *etab = tps_smem;
    double *xs = etab + 256;
    double *fs = xs + S::E * kTpsPitch;
    const long long b0 = (long long)blockIdx.x * kTpsThreads;
    const int rows = (int)min((long long)kTpsThreads, B - b0);
    {
        T xr[S::E], fr[S::N];
        tps_fetch<T, S::E>(X, b0, rows, x_is_log ? 0.0 : 1.0, xr);
        tps_fetch<T, S::N>(F, b0, rows, 0.0, fr);
        for (int i = threadIdx.x; i < 256; i += kTpsThreads) etab[i] = kExp256Tab[i];
        tps_scatter<T, S::E>(xr, xs);
        tps_scatter<T, S::N>(fr, fs);
    }
    __syncthreads();
    int bad = tps_conductivities<S::E>(xs + threadIdx.x, nullptr, x_is_log, smem_u32_of(etab));
    if ((int)threadIdx.x >= rows) bad = 0;
    {
        const double *x = xs + threadIdx.x;
        double *f = fs + threadIdx.x;
        double Ab[S::NF][S::HBW + 1], z[S::NF];
        tps_assemble<S>(tab, x, Ab);
#pragma unroll
        for (int i = 0; i < S::NF; ++i) {
            double acc = f[tab.free_dof[i]];
#pragma unroll
            for (int t = 0; t < S::TR; ++t) {
                const int k = i * S::TR + t;
                acc = fma(-tab.rhs_coef[k] * x[tab.rhs_elem[k]], f[tab.rhs_dof[k]], acc);
            }
            z[i] = acc;
        }
        const int fbad = tps_factor<S, true>(Ab, z);
        if ((int)threadIdx.x < rows) bad |= fbad;
        tps_backward_subst<S>(Ab, z);
#pragma unroll
        for (int i = 0; i < S::NF; ++i) f[tab.free_dof[i]] = z[i];
    }
    if (bad && info) atomicOr(info, bad);
    __syncthreads();
    tps_stage_out<T, S::N>(u, b0, rows, fs);
}

// shared memory: [exp table 256][x: E columns (then u when !GRADF)][exp(X) -> dL/dX: E columns][gbar -> lambda: N columns]
//                ([u: N columns] when GRADF: x stays live until the constrained rows of lambda are formed)
template <typename T, class S, bool GRADF>
__global__ void __launch_bounds__(kTpsThreads, GRADF ? 1 : 2)
rom_tps_adjoint_kernel(const __grid_constant__ TpsAdjTab<S> tab, const T *__restrict__ X, int x_is_log,
                       const T *__restrict__ u, const T *__restrict__ gbar, T *__restrict__ gradX,
                       T *__restrict__ gradF, long long B) {
    extern __shared__ __align__(16) double tps_smem[];
    double *etab = tps_smem;
    double *xs = etab + 256;
    double *dxs = xs + S::E * kTpsPitch;
    double *gs = dxs + S::E * kTpsPitch;
    double *us = GRADF ? gs + S::N * kTpsPitch : xs;
    const long long b0 = (long long)blockIdx.x * kTpsThreads;
    const int rows = (int)min((long long)kTpsThreads, B - b0);
    T ur[S::N];   // u is fetched with the other inputs; without GRADF it waits in registers until the columns of x are free
    {
        T xr[S::E], gr[S::N];
        tps_fetch<T, S::E>(X, b0, rows, x_is_log ? 0.0 : 1.0, xr);
        tps_fetch<T, S::N>(gbar, b0, rows, 0.0, gr);
        tps_fetch<T, S::N>(u, b0, rows, 0.0, ur);
        for (int i = threadIdx.x; i < 256; i += kTpsThreads) etab[i] = kExp256Tab[i];
        tps_scatter<T, S::E>(xr, xs);
        tps_scatter<T, S::N>(gr, gs);
        if (GRADF) tps_scatter<T, S::N>(ur, us);
    }
    __syncthreads();
    tps_conductivities<S::E>(xs + threadIdx.x, dxs + threadIdx.x, x_is_log, smem_u32_of(etab));
    const double *x = xs + threadIdx.x;
    double *g = gs + threadIdx.x;
    {
        double Ab[S::NF][S::HBW + 1], z[S::NF];
        tps_assemble<S>(tab.f, x, Ab);
        tps_factor<S, false>(Ab, z);
#pragma unroll
        for (int i = 0; i < S::NF; ++i) z[i] = g[tab.f.free_dof[i]];
        tps_forward_subst<S>(Ab, z);
        tps_backward_subst<S>(Ab, z);
        if (GRADF) {   // lambda on the constrained rows: gbar_c - sum_f K_fc lambda_f  (same couplings as the forward rhs)
#pragma unroll
            for (int i = 0; i < S::NF; ++i)
#pragma unroll
                for (int t = 0; t < S::TR; ++t) {
                    const int k = i * S::TR + t;
                    g[tab.f.rhs_dof[k]] = fma(-tab.f.rhs_coef[k] * x[tab.f.rhs_elem[k]], z[i], g[tab.f.rhs_dof[k]]);
                }
        }
#pragma unroll
        for (int i = 0; i < S::NF; ++i) g[tab.f.free_dof[i]] = z[i];
    }
    if (!GRADF) {   // x is dead: its columns take u
        __syncthreads();
        tps_scatter<T, S::N>(ur, us);
        __syncthreads();
    }
    {
        const double *uu = us + threadIdx.x;
        double *dx = dxs + threadIdx.x;
#pragma unroll 1
        for (int e = 0; e < S::E; ++e) {      // rolled: the tables are read with a run-time index from the constant bank
            double acc = 0.0;
#pragma unroll
            for (int t = 0; t < S::TG; ++t) {
                const int k = e * S::TG + t;
                acc = fma(tab.grad_coef[k] * g[tab.grad_i[k]], uu[tab.grad_j[k]], acc);
            }
            dx[e * kTpsPitch] = -acc * dx[e * kTpsPitch];   // chain rule through x = exp(X) + 1e-8 (dx = 1 for conductivity input)
        }
    }
    __syncthreads();
    tps_stage_out<T, S::E>(gradX, b0, rows, dxs);
    if (GRADF) tps_stage_out<T, S::N>(gradF, b0, rows, gs);
}

template <class S>
constexpr size_t tps_smem_forward() {
    return sizeof(double) * (256 + (size_t)(S::E + S::N) * kTpsPitch);
}
template <class S>
constexpr size_t tps_smem_adjoint(bool gradF) {
    return sizeof(double) * (256 + (size_t)(2 * S::E + S::N + (gradF ? S::N : 0)) * kTpsPitch);
}

}  // namespace gpde
