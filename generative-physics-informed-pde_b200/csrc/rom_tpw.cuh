// Coarse-grained model, one THREAD per sample with a SLIDING REGISTER WINDOW over the band.  Included by rom.cu.
//
// For coarse meshes whose banded LDL^T no longer fits the registers of one thread (the reference's 8x8 mesh of BASELINE
// config 3: 63 free dofs, half bandwidth 7 -> 504 band entries), the thread-per-sample idea of rom_tps.cuh is kept and the
// band is streamed through a window: at pivot k only rows k .. k+HBW of the band are live, W = HBW + 1 rows of W entries.
//   * the window R[row mod W][s] = A[row][row - s] lives in registers; the pivot loop is rolled over blocks of W pivots and
//     unrolled inside a block, so every register index is a compile-time constant and a row's slot is reused W pivots later;
//   * row k + HBW is assembled from the sample's conductivity column (shared memory) just before pivot k touches it: the
//     5-point structure of K_ff on the reference's meshes (the hypotenuse of a right triangle does not couple) leaves three
//     structural entries per row -- diagonal, s = 1 and s = HBW -- everything in between is fill-in that starts at zero;
//   * the forward substitution rides along in a window of W right-hand-side values; finished values go to the sample's
//     column of F in shared memory (their slots are dead by then);
//   * column k of the factor (1/d_k and the HBW unscaled entries below it) leaves the registers for the factor stash as soon
//     as pivot k is done.  The stash is sample-interleaved ([block of 128 samples][pivot][entry][sample]): every store and
//     every later load of a warp is 256 contiguous bytes.  The back substitution of the same kernel and both substitutions
//     of the adjoint kernel stream it back (L2-resident within the kernel: 516 KB per CTA);
//   * rows past the end of the matrix are phantom rows of zeros (tables padded with zero coefficients), pivots past the end
//     are skipped by a uniform predicate: no per-entry bounds in the inner loops.
// The cooperative kernels of rom.cu spent ~1000 warp-instructions per sample on table walks, a __syncwarp per pivot and
// shared-memory round trips of the band (0.32 + 0.25 ms for 16384 samples of the 8x8 mesh); here a sample costs ~90
// instructions per pivot of straight-line FP64 code.  Same maths: bottleneck/ROM.py:59-100 and its autograd.
#pragma once
#include <utility>

#include "exp256.cuh"

namespace gpde {

constexpr int kTpwThreads = 128;
constexpr int kTpwPitch = kTpwThreads + 1;   // doubles between consecutive rows of a per-thread column array
constexpr int kTpwGradChunk = 32;            // gradient entries per round of the adjoint's output staging

template <int NF_, int HBW_, int E_, int N_, int TO_, int TD_, int TR_, int TG_>
struct TpwShape {
    static constexpr int NF = NF_, HBW = HBW_, E = E_, N = N_, TO = TO_, TD = TD_, TR = TR_, TG = TG_;
    static constexpr int W = HBW_ + 1;                 // window rows = entries per row
    static constexpr int NB = (NF_ + W - 1) / W;       // blocks of W pivots
    static constexpr int NFP = NB * W + HBW_;          // table rows incl. phantom rows
    static_assert(HBW_ >= 2, "s = 1 and s = HBW must be different slots");
};
using TpwShape8x8 = TpwShape<63, 7, 128, 81, 2, 6, 2, 7>;

// One record per band row, everything the row needs in one place (the pivot loop indexes the table with a run-time row:
// as separate arrays in the constant bank every row entry cost ~10 cold constant-cache misses -- the tables are copied to
// SHARED memory once per CTA instead and read there with uniform addresses).
// element / dof entries are column offsets (index * kTpwPitch); padding terms have coefficient 0 and offset 0
template <class S>
struct TpwRow {
    double diag_coef[S::TD];
    double s1_coef[S::TO];      // A[i][i-1]
    double sh_coef[S::TO];      // A[i][i-HBW]
    double rhs_coef[S::TR];     // z_i = F[free_i] - sum_t rhs_coef * x[rhs_elem] * F[rhs_dof]
    unsigned short diag_elem[S::TD], s1_elem[S::TO], sh_elem[S::TO], rhs_elem[S::TR], rhs_dof[S::TR];
    unsigned short free_dof, pad[(8 - (S::TD + 2 * S::TO + 2 * S::TR + 1) % 8) % 8];
};
template <class S>
struct TpwFwdTab {
    TpwRow<S> row[S::NFP];
    __device__ __forceinline__ int free_of(int k) const { return row[k].free_dof; }
};
// Gradient in EDGE form: a symmetric element matrix with zero row sums is K_e = sum over its edges (a,b) of
// w_ab (e_a - e_b)(e_a - e_b)^T with w_ab = -K_e[a][b], so
//     dL/dx_e = - lam~^T K_e u = - sum_edges w_ab (lam~_a - lam~_b)(u_a - u_b),   lam~ = lam on free dofs, 0 on constrained ones
// (rows of constrained dofs were overwritten and contribute nothing).  2 edges of a right triangle carry a
// weight (the hypotenuse does not couple): 4 values, 2 subtractions, a multiplication and an FMA per edge instead of the 7
// coefficient terms with 14 values of the entry-by-entry form (the gradient loop was 42 % of the adjoint kernel's samples).
// lam offsets of constrained dofs point to a column of zeros.
constexpr int kTpwEdges = 3;
template <class S>
struct TpwAdjTab {
    double edge_w[S::E * kTpwEdges];
    double rhs_coef[S::NF * S::TR];                              // couplings K_fc (gradient w.r.t. F only)
    unsigned short lam_a[S::E * kTpwEdges], lam_b[S::E * kTpwEdges];   // lambda columns of the edge's ends (zero column if constrained)
    unsigned short u_a[S::E * kTpwEdges], u_b[S::E * kTpwEdges];       // u columns of the edge's ends
    unsigned short rhs_elem[(S::NF * S::TR + 7) / 8 * 8];        // element INDEX (x is read from global memory there)
    unsigned short rhs_dof[(S::NF * S::TR + 7) / 8 * 8];         // dof column
    unsigned short free_dof[(S::NFP + 7) / 8 * 8];
    __device__ __forceinline__ int free_of(int k) const { return free_dof[k]; }
};

template <class S>
__host__ __device__ constexpr size_t tpw_stash_doubles_per_block() {
    return (size_t)S::NF * S::W * kTpwThreads;
}

// Rows [b0, b0+rows) x [0, WD) of a row-major [B, WD] array into per-thread columns col[j * pitch + row], CH coalesced
// loads in flight per thread; rows past the batch get ``fill``.
// COND: the values are conductivities -- x = exp(v) + 1e-8 (components.py:298) applied on the way in when x_is_log, and
// GPDE_INFO_NONPOSITIVE_X returned if some x <= 1e-12 (ROM.py:74-76); CH independent exp chains per thread.
template <typename T, int WD, int CH, bool COND = false>
__device__ __forceinline__ int tpw_stage_in(const T *__restrict__ src, long long b0, int rows, double fill, double *col,
                                            int x_is_log = 0, unsigned etab = 0) {
    const T *p = src + b0 * WD;
    const int limit = rows * WD;
    int bad = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < WD; k0 += CH) {
        T raw[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int i = threadIdx.x + (k0 + c) * kTpwThreads;
            raw[c] = (k0 + c < WD && i < limit) ? p[i] : (T)fill;
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (k0 + c < WD) {
                const int i = threadIdx.x + (k0 + c) * kTpwThreads;
                const int row = i / WD, j = i - row * WD;
                double v = (double)raw[c];
                if (COND) {
                    if (x_is_log) v = (exp256_in_range(v) ? exp_tab256c(v, etab) : exp(v)) + 1e-8;
                    if (!(v > 1e-12)) bad = GPDE_INFO_NONPOSITIVE_X;
                }
                col[j * kTpwPitch + row] = v;
            }
        }
    }
    return bad;
}

template <typename T, int WD>
__device__ __forceinline__ void tpw_stage_out(T *__restrict__ dst, long long b0, int rows, const double *col) {
    T *p = dst + b0 * WD;
    const int limit = rows * WD;
#pragma unroll 4
    for (int i = threadIdx.x; i < limit; i += kTpwThreads) {
        const int row = i / WD, j = i - row * WD;
        p[i] = (T)col[j * kTpwPitch + row];
    }
}

// x = exp(v) + 1e-8 in place (components.py:298) for this thread's column; GPDE_INFO_NONPOSITIVE_X if some x <= 1e-12
template <int E>
__device__ __forceinline__ int tpw_conductivities(double *x, int x_is_log, unsigned etab) {
    int bad = 0;
#pragma unroll 8
    for (int e = 0; e < E; ++e) {      // 8 independent exp chains in flight (one warp per scheduler: nothing else hides them)
        double v = x[e * kTpwPitch];
        if (x_is_log) {
            v = (exp256_in_range(v) ? exp_tab256c(v, etab) : exp(v)) + 1e-8;
            x[e * kTpwPitch] = v;
        }
        if (!(v > 1e-12)) bad = GPDE_INFO_NONPOSITIVE_X;
    }
    return bad;
}

// Row r of K_ff(x) into window slot SLOT (entries s = 0 .. HBW) and its right-hand side into zw[SLOT]
template <class S, int SLOT>
__device__ __forceinline__ void tpw_enter_row(const TpwFwdTab<S> &tab, int r, const double *x, const double *f,
                                              double (&R)[S::W][S::W], double (&zw)[S::W]) {
    const TpwRow<S> &row = tab.row[r];
    double d = 0.0, o1 = 0.0, oh = 0.0;
#pragma unroll
    for (int t = 0; t < S::TD; ++t) d = fma(row.diag_coef[t], x[row.diag_elem[t]], d);
#pragma unroll
    for (int t = 0; t < S::TO; ++t) {
        o1 = fma(row.s1_coef[t], x[row.s1_elem[t]], o1);
        oh = fma(row.sh_coef[t], x[row.sh_elem[t]], oh);
    }
    R[SLOT][0] = d;
    R[SLOT][1] = o1;
#pragma unroll
    for (int s = 2; s < S::HBW; ++s) R[SLOT][s] = 0.0;
    R[SLOT][S::HBW] = oh;
    double acc = f[row.free_dof];
#pragma unroll
    for (int t = 0; t < S::TR; ++t) acc = fma(-row.rhs_coef[t] * x[row.rhs_elem[t]], f[row.rhs_dof[t]], acc);
    zw[SLOT] = acc;
}

// rows 0 .. HBW-1 of the window (row k + HBW enters at pivot k)
template <class S, int... Js>
__device__ __forceinline__ void tpw_prologue(std::integer_sequence<int, Js...>, const TpwFwdTab<S> &tab, const double *x,
                                             const double *f, double (&R)[S::W][S::W], double (&zw)[S::W]) {
    (tpw_enter_row<S, Js>(tab, Js, x, f, R, zw), ...);
}

// Pivot k = kb * W + J: bring in row k + HBW, eliminate column k from the window, advance the forward substitution,
// flush column k of the factor to the stash.  J (= k mod W) fixes every register index at compile time.
template <class S, int J>
__device__ __forceinline__ void tpw_pivot(const TpwFwdTab<S> &tab, int kb, const double *x, double *f, double (&R)[S::W][S::W],
                                          double (&zw)[S::W], double *__restrict__ st, int &bad) {
    constexpr int W = S::W, HBW = S::HBW;
    const int k = kb * W + J;
    tpw_enter_row<S, (J + HBW) % W>(tab, k + HBW, x, f, R, zw);
    if (k < S::NF) {
        const double d = R[J][0];
        if (!(d > 0.0)) bad |= GPDE_INFO_NOT_SPD;
        const double invd = fast_rcp(d);
        double *col = st + (size_t)k * W * kTpwThreads;
        col[0] = invd;
#pragma unroll
        for (int si = 1; si <= HBW; ++si) {
            const double c = R[(J + si) % W][si];
            col[si * kTpwThreads] = c;
            const double l = c * invd;
#pragma unroll
            for (int sj = 1; sj <= si; ++sj)
                R[(J + si) % W][si - sj] = fma(-l, R[(J + sj) % W][sj], R[(J + si) % W][si - sj]);
        }
        const double wk = zw[J] * invd;
#pragma unroll
        for (int s = 1; s <= HBW; ++s) zw[(J + s) % W] = fma(-R[(J + s) % W][s], wk, zw[(J + s) % W]);
        f[tab.row[k].free_dof] = wk;
    }
}
template <class S, int... Js>
__device__ __forceinline__ void tpw_pivot_block(std::integer_sequence<int, Js...>, const TpwFwdTab<S> &tab, int kb, const double *x,
                                                double *f, double (&R)[S::W][S::W], double (&zw)[S::W], double *__restrict__ st,
                                                int &bad) {
    (tpw_pivot<S, Js>(tab, kb, x, f, R, zw, st, bad), ...);
}

// Column k of the stashed factor (W entries) into registers; columns past the matrix read as zeros
template <class S>
__device__ __forceinline__ void tpw_load_column(const double *__restrict__ st, int k, double (&c)[S::W]) {
    const double *col = st + (size_t)k * S::W * kTpwThreads;
    const bool in = k >= 0 && k < S::NF;
#pragma unroll
    for (int s = 0; s < S::W; ++s) c[s] = in ? col[s * kTpwThreads] : 0.0;
}

// L^T sol = w with the stashed factor; w and then sol live in column ``v`` at the free dofs.  A rotating register buffer
// holds the next W factor columns: as soon as a column has been used, the one W pivots further on is fetched into its slot,
// so every load is W pivots of dependent arithmetic ahead of its use: with one warp per scheduler nothing else hides the
// L2 / HBM latency (36 % of the forward kernel's samples sat on these loads before)
template <class S, class Tab>
__device__ __forceinline__ void tpw_backward_subst(const double *__restrict__ st, const Tab &tab, double *v) {
    constexpr int W = S::W, HBW = S::HBW, NF = S::NF, NB = S::NB;
    double sw[W], cur[W][W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        sw[j] = 0.0;
        tpw_load_column<S>(st, (NB - 1) * W + j, cur[j]);
    }
#pragma unroll 1
    for (int kb = NB - 1; kb >= 0; --kb) {
#pragma unroll
        for (int j = W - 1; j >= 0; --j) {
            const int k = kb * W + j;
            if (k < NF) {
                double acc = 0.0;
#pragma unroll
                for (int s = 1; s <= HBW; ++s) acc = fma(cur[j][s], sw[(j + s) % W], acc);
                const int at = tab.free_of(k);
                const double sol = fma(-cur[j][0], acc, v[at]);
                sw[j] = sol;
                v[at] = sol;
            }
            tpw_load_column<S>(st, k - W, cur[j]);
        }
    }
}

// 16-byte copy of a plan table into shared memory (all threads)
template <class Tab>
__device__ __forceinline__ void tpw_table_to_smem(const Tab *__restrict__ src, Tab *dst) {
    static_assert(sizeof(Tab) % 16 == 0, "tables are copied in 16-byte pieces");
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(Tab) / 16); i += kTpwThreads) d4[i] = s4[i];
}

// shared memory: [row table][exp table 256][x: E columns][F -> u: N columns]
template <typename T, class S>
__global__ void __launch_bounds__(kTpwThreads, 1)
rom_tpw_forward_kernel(const TpwFwdTab<S> *__restrict__ tab_g, const T *__restrict__ X, int x_is_log,
                       const T *__restrict__ F, T *__restrict__ u, double *__restrict__ stash, int *info, long long B) {
    extern __shared__ __align__(16) double tpw_smem[];
    constexpr int W = S::W, HBW = S::HBW, NB = S::NB;
    TpwFwdTab<S> &tab = *reinterpret_cast<TpwFwdTab<S> *>(tpw_smem);
    double *etab = tpw_smem + sizeof(TpwFwdTab<S>) / sizeof(double);
    double *xs = etab + 256;
    double *fs = xs + S::E * kTpwPitch;
    tpw_table_to_smem(tab_g, &tab);
    const long long b0 = (long long)blockIdx.x * kTpwThreads;
    const int rows = (int)min((long long)kTpwThreads, B - b0);
    for (int i = threadIdx.x; i < 256; i += kTpwThreads) etab[i] = kExp256Tab[i];
    __syncthreads();      // the exp table is used while staging
    // (32 coalesced loads in flight per thread: the staging is a chain of DRAM round trips, one per chunk; the conductivities
    // get their exp() on the way in -- 32 independent chains per thread instead of a second pass over shared memory)
    int bad = tpw_stage_in<T, S::E, 32, true>(X, b0, rows, x_is_log ? 0.0 : 1.0, xs, x_is_log, smem_u32_of(etab));
    tpw_stage_in<T, S::N, 41>(F, b0, rows, 0.0, fs);
    __syncthreads();
    double *x = xs + threadIdx.x, *f = fs + threadIdx.x;
    double *st = stash + (size_t)blockIdx.x * tpw_stash_doubles_per_block<S>() + threadIdx.x;
    {
        double R[W][W], zw[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            zw[j] = 0.0;
#pragma unroll
            for (int s = 0; s < W; ++s) R[j][s] = 0.0;
        }
        tpw_prologue<S>(std::make_integer_sequence<int, HBW>{}, tab, x, f, R, zw);
#pragma unroll 1
        for (int kb = 0; kb < NB; ++kb) tpw_pivot_block<S>(std::make_integer_sequence<int, W>{}, tab, kb, x, f, R, zw, st, bad);
    }
    tpw_backward_subst<S>(st, tab, f);
    // (the staging flags belong to whatever rows a thread copied, the factor flag to its own sample: rows past the batch
    // were filled with harmless values, so every flag raised is a real one)
    if (bad && info) atomicOr(info, bad);
    __syncthreads();
    tpw_stage_out<T, S::N>(u, b0, rows, fs);
}

// shared memory: [tables][exp table 256][gbar -> lambda: N columns][u: N columns][gradient chunk: kTpwGradChunk columns]
template <typename T, class S, bool GRADF>
__global__ void __launch_bounds__(kTpwThreads, 1)
rom_tpw_adjoint_kernel(const TpwAdjTab<S> *__restrict__ tab_g, const T *__restrict__ X, int x_is_log,
                       const T *__restrict__ u, const double *__restrict__ stash, const T *__restrict__ gbar,
                       T *__restrict__ gradX, T *__restrict__ gradF, long long B) {
    extern __shared__ __align__(16) double tpw_smem[];
    constexpr int W = S::W, HBW = S::HBW, NF = S::NF, NB = S::NB, EC = kTpwGradChunk;
    static_assert(S::E % EC == 0, "gradient chunks");
    TpwAdjTab<S> &tab = *reinterpret_cast<TpwAdjTab<S> *>(tpw_smem);
    double *etab = tpw_smem + sizeof(TpwAdjTab<S>) / sizeof(double);
    double *gs = etab + 256;                        // N columns + one column of zeros (lambda~ of constrained dofs)
    double *us = gs + (S::N + 1) * kTpwPitch;
    gs[S::N * kTpwPitch + threadIdx.x] = 0.0;
    tpw_table_to_smem(tab_g, &tab);
    double *ds = us + S::N * kTpwPitch;
    const long long b0 = (long long)blockIdx.x * kTpwThreads;
    const int rows = (int)min((long long)kTpwThreads, B - b0);
    for (int i = threadIdx.x; i < 256; i += kTpwThreads) etab[i] = kExp256Tab[i];
    tpw_stage_in<T, S::N, 41>(gbar, b0, rows, 0.0, gs);
    tpw_stage_in<T, S::N, 41>(u, b0, rows, 0.0, us);
    __syncthreads();
    double *g = gs + threadIdx.x;
    const double *uu = us + threadIdx.x;
    const double *st = stash + (size_t)blockIdx.x * tpw_stash_doubles_per_block<S>() + threadIdx.x;
    {   // w = D^-1 L^-1 gbar_f, window of W values; the next block's factor columns are fetched ahead of the chain
        double zw[W], cur[W][W];
#pragma unroll
        for (int j = 0; j < HBW; ++j) zw[j] = g[tab.free_dof[j]];
        zw[HBW] = 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) tpw_load_column<S>(st, j, cur[j]);
#pragma unroll 1
        for (int kb = 0; kb < NB; ++kb) {
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int k = kb * W + j;
                zw[(j + HBW) % W] = g[tab.free_dof[k + HBW]];     // phantom rows read column 0: never used (their L entries are 0)
                if (k < NF) {
                    const double wk = zw[j] * cur[j][0];
#pragma unroll
                    for (int s = 1; s <= HBW; ++s) zw[(j + s) % W] = fma(-cur[j][s], wk, zw[(j + s) % W]);
                    g[tab.free_dof[k]] = wk;
                }
                tpw_load_column<S>(st, k + W, cur[j]);
            }
        }
    }
    tpw_backward_subst<S>(st, tab, g);
    if (GRADF && (int)threadIdx.x < rows) {
        // lambda on the constrained rows: gbar_c - sum_f K_cf lambda_f (x of the few boundary elements from global memory)
        const T *Xb = X + (b0 + threadIdx.x) * S::E;
#pragma unroll 1
        for (int i = 0; i < NF; ++i) {
            const double li = g[tab.free_dof[i]];
#pragma unroll
            for (int t = 0; t < S::TR; ++t) {
                const int k = i * S::TR + t;
                const double c = tab.rhs_coef[k];
                if (c != 0.0) {
                    double xv = (double)Xb[tab.rhs_elem[k]];
                    if (x_is_log) xv = (exp256_in_range(xv) ? exp_tab256c(xv, smem_u32_of(etab)) : exp(xv)) + 1e-8;
                    g[tab.rhs_dof[k]] = fma(-c * xv, li, g[tab.rhs_dof[k]]);
                }
            }
        }
    }
    // dL/dX in rounds of EC elements: per-thread columns -> coalesced rows, chain rule through x = exp(X) + 1e-8 on the way out
    double *dcol = ds + threadIdx.x;
#pragma unroll 1
    for (int c0 = 0; c0 < S::E; c0 += EC) {
        // the conductivity inputs this thread will need for the round's output (entries threadIdx.x + 128 it of the CTA's
        // [rows][EC] block) are requested first: they arrive while the round's gradient entries are computed
        T xv[EC];
#pragma unroll
        for (int it = 0; it < EC; ++it) {
            const int i = threadIdx.x + it * kTpwThreads;
            const int row = i / EC, j = i - row * EC;
            xv[it] = (x_is_log && i < rows * EC) ? X[(b0 + row) * S::E + c0 + j] : (T)0;
        }
#pragma unroll 2
        for (int e = 0; e < EC; ++e) {
            double acc = 0.0;
#pragma unroll
            for (int t = 0; t < kTpwEdges; ++t) {
                const int k = (c0 + e) * kTpwEdges + t;
                acc = fma(tab.edge_w[k] * (g[tab.lam_a[k]] - g[tab.lam_b[k]]), uu[tab.u_a[k]] - uu[tab.u_b[k]], acc);
            }
            dcol[e * kTpwPitch] = -acc;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < EC; ++it) {
            const int i = threadIdx.x + it * kTpwThreads;
            if (i < rows * EC) {
                const int row = i / EC, j = i - row * EC;
                double v = ds[j * kTpwPitch + row];
                if (x_is_log) {
                    const double xd = (double)xv[it];
                    v *= exp256_in_range(xd) ? exp_tab256c(xd, smem_u32_of(etab)) : exp(xd);
                }
                gradX[(b0 + row) * S::E + c0 + j] = (T)v;
            }
        }
        __syncthreads();
    }
    if (GRADF) tpw_stage_out<T, S::N>(gradF, b0, rows, gs);
}

template <class S>
constexpr size_t tpw_smem_forward() {
    return sizeof(TpwFwdTab<S>) + sizeof(double) * (256 + (size_t)(S::E + S::N) * kTpwPitch);
}
template <class S>
constexpr size_t tpw_smem_adjoint() {
    return sizeof(TpwAdjTab<S>) + sizeof(double) * (256 + (size_t)(2 * S::N + 1 + kTpwGradChunk) * kTpwPitch);
}

}  // namespace gpde
