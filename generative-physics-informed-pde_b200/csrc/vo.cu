// Virtual-observable residuals on sm_100a, matrix-free.
//   r_b = V^T (K_fom(a_b) u~_b - f)_free          (= Gamma_b y_b - alpha_b,  VirtualObservables.py:61-69, 662)
//   q_b = K_ff(a_b) (V s_b)                        (= Gamma_b^T s_b,          VirtualObservables.py:663)
// The reference materialises K_fom as scipy CSR per data point with FEniCS and forms the dense
// Gamma = V^T K on the CPU; here K_fom(a) = sum_c a_c K_c is never formed: each free row i gathers its
// incident cells ("row-cell ELL", slot-major so that consecutive rows read consecutive words):
//   rho_i = sum_slots a[in] * (c0 * u_i + c1 * u~[j1] + c2 * u~[j2]) - f_i
//
// Kernel families, chosen per call (gpde_vo_plan_kernel_path):
//   vo_grid.cuh   structured pixel grid (verified at plan creation), FP64 I/O: marching flux-form kernel with the
//                 contraction on the FP64 tensor pipe inside the kernel (m <= 32), or producing rho for ...
//   vo_gemm.cuh   ... the FP64 tensor-core contraction rho[B,d] V[d,m] for many weighting functions (m > 32);
//   vo_fused.cuh  any P1 diffusion mesh, m <= 32: edge-form matvec + ring-staged tiles + DMMA contraction;
//   version 1     (below) any mesh, any m, FP32/FP64: row-cell ELL matvec, then vo_gemm.cuh.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace gpde {

struct VoDev {
    int n_nodes, n_cells, n_inputs, d, n_bc, nslots;
    const int *ell_in;    // [nslots*d] conductivity input index of the slot's cell
    const int *ell_j1;    // [nslots*d] source of the 2nd vertex: >=0 -> y index, <0 -> -(g index)-1
    const int *ell_j2;    // [nslots*d]
    const double *ell_c0, *ell_c1, *ell_c2;   // [nslots*d]
    const double *f_free;  // [d]
};

}  // namespace gpde

namespace gpde {

__device__ __forceinline__ double ldd(const double *p) { return *p; }
__device__ __forceinline__ double ldd(const float *p) { return (double)*p; }

constexpr int kVoThreads = 256;

}  // namespace gpde

#include "vo_fused.cuh"
#include "vo_grid.cuh"
#include "vo_grid2.cuh"
#include "vo_expand.cuh"
#include "vo_gemm.cuh"
#include "vo_gridgemm.cuh"
#include "vo_posterior.cuh"

namespace gpde {
// Experiment switches (A/B runs of the kernel families).  Read from the environment ONCE per plan, in gpde_vo_plan_create,
// never on a launch path.
struct VoEnv {
    int grid_r = 2;           // GPDE_GRID_R: node rows per stage of the general grid kernel (1 or 2)
    bool path_v1 = false;     // GPDE_VO_PATH=v1: version-1 kernels only
    bool path_fused = false;  // GPDE_VO_PATH=fused: no structured-grid kernels
    bool no_grid2 = false;    // GPDE_GRID_V=1: general grid kernel instead of the lean one
    int grid2_nvs = 0;        // GPDE_GRID2_NVS: V stages of the lean kernel (2 or 3; 0 = automatic)
    int grid2_flags = 0;      // GPDE_GRID2_FLAGS
    bool grid2_pdl = true;    // GPDE_GRID2_PDL=0: no programmatic dependent launch
    int grid_debug = 0;       // GPDE_GRID_DEBUG
    bool sync_staging = false;   // GPDE_VO_SYNC_STAGING
    bool expand_gemm = false;    // GPDE_VO_EXPAND=gemm
    int gemm_splits = 0;         // GPDE_GEMM_SPLITS: parts of the contraction length (0 = automatic)
    int grid2_split = 1;         // GPDE_GRID2_SPLIT: 0 = never cut the node rows over a cluster, 1 = automatic, n > 1 = force n
    bool gridgemm = true;        // GPDE_VO_GRIDGEMM=0: rho through HBM + vo_gemm_kernel instead of the one-kernel route (m > 32)
    bool fused_t = true;         // GPDE_VO_FUSED_T=0: residual_T as expansion kernel + marching kernel (w [B,d] through HBM)
    int grid2_spc = 0;           // GPDE_GRID2_SPC: samples per CTA of the lean grid kernel (0 = automatic, see grid2_samples_per_cta)
    VoEnv() {
        const char *e;
        if ((e = getenv("GPDE_GRID_R"))) grid_r = atoi(e);
        if ((e = getenv("GPDE_VO_PATH"))) { path_v1 = strcmp(e, "v1") == 0; path_fused = strcmp(e, "fused") == 0; }
        if ((e = getenv("GPDE_GRID_V"))) no_grid2 = atoi(e) == 1;
        if ((e = getenv("GPDE_GRID2_NVS"))) grid2_nvs = atoi(e);
        if ((e = getenv("GPDE_GRID2_FLAGS"))) grid2_flags = atoi(e);
        if ((e = getenv("GPDE_GRID2_PDL"))) grid2_pdl = atoi(e) != 0;
        if ((e = getenv("GPDE_GRID_DEBUG"))) grid_debug = atoi(e);
        sync_staging = getenv("GPDE_VO_SYNC_STAGING") != nullptr;
        if ((e = getenv("GPDE_VO_EXPAND"))) expand_gemm = strcmp(e, "gemm") == 0;
        if ((e = getenv("GPDE_GEMM_SPLITS"))) gemm_splits = std::max(0, std::min(8, atoi(e)));
        if ((e = getenv("GPDE_GRID2_SPLIT"))) grid2_split = std::max(0, std::min(8, atoi(e)));
        if ((e = getenv("GPDE_VO_GRIDGEMM"))) gridgemm = atoi(e) != 0;
        if ((e = getenv("GPDE_GRID2_SPC"))) grid2_spc = std::max(0, atoi(e));
        if ((e = getenv("GPDE_VO_FUSED_T"))) fused_t = atoi(e) != 0;
    }
};
}  // namespace gpde

struct gpde_vo_plan {
    gpde::VoDev dev;
    gpde::VoTiles tiles;
    gpde::GridDev grid;
    gpde::VoCsrDev csr;     // K_ff as CSR term lists (vo_posterior.cuh)
    size_t fused_smem;
    int device;
    int n_sm;
    gpde::VoEnv env;
    std::vector<void *> allocs;
};

namespace gpde {

// rho[b,i] for S samples per CTA.  ea (shared) = conductivities of the S samples, exp() applied.
template <typename Ta, typename Ty, typename To, int S>
__global__ void __launch_bounds__(kVoThreads)
vo_matvec_kernel(VoDev P, const Ta *__restrict__ a, long long a_stride, int a_is_log,
                 const Ty *__restrict__ y, const Ta *__restrict__ g, long long g_stride, int sub_f,
                 double *__restrict__ rho_ws, int ws_stride, To *__restrict__ rho_out, long long B, int stage_a) {
    extern __shared__ double ea[];   // [S][n_inputs] when stage_a
    const int d = P.d;
    for (long long b0 = (long long)blockIdx.x * S; b0 < B; b0 += (long long)gridDim.x * S) {
        if (stage_a) {
            __syncthreads();
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const long long b = min(b0 + s, B - 1);
                for (int p = threadIdx.x; p < P.n_inputs; p += kVoThreads) {
                    double v = ldd(a + b * a_stride + p);
                    ea[s * P.n_inputs + p] = a_is_log ? exp(v) : v;
                }
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < d; i += kVoThreads) {
            double acc[S], yi[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const long long b = min(b0 + s, B - 1);
                yi[s] = y ? ldd(y + b * d + i) : 0.0;
                acc[s] = sub_f ? -P.f_free[i] : 0.0;
            }
            for (int t = 0; t < P.nslots; ++t) {
                const int k = t * d + i;
                const int in = P.ell_in[k], j1 = P.ell_j1[k], j2 = P.ell_j2[k];
                const double c0 = P.ell_c0[k], c1 = P.ell_c1[k], c2 = P.ell_c2[k];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const long long b = min(b0 + s, B - 1);
                    double av;
                    if (stage_a) av = ea[s * P.n_inputs + in];
                    else {
                        av = ldd(a + b * a_stride + in);
                        if (a_is_log) av = exp(av);
                    }
                    double u1, u2;
                    if (j1 >= 0) u1 = y ? ldd(y + b * d + j1) : 0.0;
                    else u1 = g ? ldd(g + b * g_stride + (-j1 - 1)) : 0.0;
                    if (j2 >= 0) u2 = y ? ldd(y + b * d + j2) : 0.0;
                    else u2 = g ? ldd(g + b * g_stride + (-j2 - 1)) : 0.0;
                    acc[s] = fma(av, fma(c0, yi[s], fma(c1, u1, c2 * u2)), acc[s]);
                }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const long long b = b0 + s;
                if (b < B) {
                    if (rho_ws) rho_ws[b * ws_stride + i] = acc[s];
                    if (rho_out) rho_out[b * d + i] = (To)acc[s];
                }
            }
        }
        if (rho_ws)   // zero the K padding the tensor-core contraction reads
            for (int i = d + threadIdx.x; i < ws_stride; i += kVoThreads)
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (b0 + s < B) rho_ws[(b0 + s) * ws_stride + i] = 0.0;
    }
}

// w[B,d] = s[B,m] V^T
template <typename Tv>
__global__ void vo_expand_kernel(const Tv *__restrict__ s, const Tv *__restrict__ V, double *__restrict__ w,
                                 long long B, int d, int m) {
    const long long total = B * (long long)d;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / d;
        const int i = (int)(idx - b * d);
        double acc = 0.0;
        for (int q = 0; q < m; ++q) acc = fma(ldd(s + b * m + q), ldd(V + (long long)i * m + q), acc);
        w[idx] = acc;
    }
}

template <typename T>
static cudaError_t track_vo(gpde_vo_plan *pl, const T **dst, const std::vector<T> &src) {
    T *p = nullptr;
    cudaError_t e = upload(&p, src);
    if (p) pl->allocs.push_back((void *)p);
    *dst = p;
    return e;
}

// Edge form + tile staging lists for the fused kernels.  Leaves pl->tiles.ok = 0 when the mesh does not
// qualify (element matrices without zero row sums, an edge shared by more than two cells, or rings that
// would not fit in shared memory); the version-1 kernels then serve every call.
static cudaError_t build_fused_tiles(gpde_vo_plan *pl, int n_nodes, int n_cells, const int32_t *cell_dofs,
                                     const double *Ke, const int32_t *cell_to_input, int d,
                                     const std::vector<int> &src, const std::vector<double> &pl_f_free) {
    VoTiles &Tl = pl->tiles;
    memset(&Tl, 0, sizeof(Tl));
    pl->fused_smem = 0;
    double kmax = 0.0;
    for (size_t k = 0; k < (size_t)n_cells * 9; ++k) kmax = std::max(kmax, fabs(Ke[k]));
    for (int c = 0; c < n_cells; ++c)
        for (int l = 0; l < 3; ++l)
            if (fabs(Ke[9 * c + 3 * l] + Ke[9 * c + 3 * l + 1] + Ke[9 * c + 3 * l + 2]) > 1e-13 * kmax)
                return cudaSuccess;   // not a pure diffusion operator
    struct Nb { int node, in0, in1, ncell; double s0, s1; };
    std::vector<std::vector<Nb>> nbs(d);
    for (int c = 0; c < n_cells; ++c)
        for (int l = 0; l < 3; ++l) {
            const int v = cell_dofs[3 * c + l];
            if (src[v] < 0) continue;
            std::vector<Nb> &list = nbs[src[v]];
            for (int l2 = 0; l2 < 3; ++l2) {
                if (l2 == l) continue;
                const double coef = Ke[9 * c + 3 * l + l2];
                if (coef == 0.0) continue;
                const int w = cell_dofs[3 * c + l2];
                Nb *hit = nullptr;
                for (Nb &nb : list)
                    if (nb.node == w) hit = &nb;
                if (!hit) {
                    list.push_back(Nb{w, cell_to_input[c], cell_to_input[c], 1, coef, 0.0});
                } else if (hit->ncell == 1) {
                    hit->in1 = cell_to_input[c];
                    hit->s1 = coef;
                    hit->ncell = 2;
                } else {
                    return cudaSuccess;   // edge shared by > 2 cells
                }
            }
        }
    int nnb = 1;
    for (int i = 0; i < d; ++i) nnb = std::max(nnb, (int)nbs[i].size());
    const int nt = (d + kFR - 1) / kFR;
    // u ring holds y by free index; a ring holds the conductivity inputs
    std::vector<int> ulo(nt, INT32_MAX), uhi(nt, -1), alo(nt, INT32_MAX), ahi(nt, -1);
    for (int i = 0; i < d; ++i) {
        const int t = i / kFR;
        ulo[t] = std::min(ulo[t], i);
        uhi[t] = std::max(uhi[t], i + 1);
        for (const Nb &nb : nbs[i]) {
            if (src[nb.node] >= 0) {
                ulo[t] = std::min(ulo[t], src[nb.node]);
                uhi[t] = std::max(uhi[t], src[nb.node] + 1);
            }
            alo[t] = std::min(alo[t], std::min(nb.in0, nb.in1));
            ahi[t] = std::max(ahi[t], std::max(nb.in0, nb.in1) + 1);
        }
    }
    for (int t = 0; t < nt; ++t)
        if (ahi[t] < 0) { alo[t] = 0; ahi[t] = 1; }   // a tile of isolated rows
    RangeRing ru = build_ring(ulo, uhi), ra = build_ring(alo, ahi);
    const int n_bc = pl->dev.n_bc;
    const int block_bytes = nnb * kFR * 32 + kFR * 8;
    const int g_base = ru.ring * kPitchB;
    const int a_base = g_base + n_bc * kPitchB;
    const int rs_base = a_base + ra.ring * kPitchB;
    const int rec_base = rs_base + (int)sizeof(double) * kFS * kRhoPitch;
    const int meta_base = rec_base + block_bytes;
    const size_t smem = (size_t)meta_base + (size_t)nt * sizeof(TileMeta);
    if (smem > 110 * 1024) return cudaSuccess;   // keep two CTAs per SM; otherwise fall back
    std::vector<TileMeta> meta(nt);
    std::vector<char> blocks((size_t)nt * block_bytes, 0);
    for (int t = 0; t < nt; ++t) {
        meta[t] = TileMeta{ru.first[t], ru.count[t], ru.off[t], ra.first[t], ra.count[t], ra.off[t], 0, 0};
        char *blk = blocks.data() + (size_t)t * block_bytes;
        int4 *ro = reinterpret_cast<int4 *>(blk);
        double2 *rc = reinterpret_cast<double2 *>(blk + (size_t)nnb * kFR * 16);
        double *fo = reinterpret_cast<double *>(blk + (size_t)nnb * kFR * 32);
        for (int k = 0; k < kFR; ++k) {
            const int i = t * kFR + k;
            if (i >= d) {   // rows past the end: harmless offsets, zero coefficients
                for (int s = 0; s < nnb; ++s) {
                    ro[s * kFR + k] = make_int4(0, a_base, a_base, 0);
                    rc[s * kFR + k] = make_double2(0.0, 0.0);
                }
                fo[k] = 0.0;
                continue;
            }
            const int own = ru.offset_of(t, i) * kPitchB;
            const int pad_in = nbs[i].empty() ? alo[t] : nbs[i][0].in0;
            for (int s = 0; s < nnb; ++s) {
                if (s < (int)nbs[i].size()) {
                    const Nb &nb = nbs[i][s];
                    const int sj = src[nb.node];
                    const int ou = sj >= 0 ? ru.offset_of(t, sj) * kPitchB : g_base + (-sj - 1) * kPitchB;
                    ro[s * kFR + k] = make_int4(ou, a_base + ra.offset_of(t, nb.in0) * kPitchB,
                                                a_base + ra.offset_of(t, nb.in1) * kPitchB, own);
                    rc[s * kFR + k] = make_double2(nb.s0, nb.s1);
                } else {   // padding: zero conductance to itself
                    const int oa = a_base + ra.offset_of(t, pad_in) * kPitchB;
                    ro[s * kFR + k] = make_int4(own, oa, oa, own);
                    rc[s * kFR + k] = make_double2(0.0, 0.0);
                }
            }
            fo[k] = pl_f_free[i];
        }
    }
    Tl.n_tiles = nt; Tl.nnb = nnb; Tl.ring_u = ru.ring; Tl.ring_a = ra.ring; Tl.n_bc = n_bc;
    Tl.g_base = g_base; Tl.a_base = a_base; Tl.rs_base = rs_base;
    Tl.rec_base = rec_base; Tl.meta_base = meta_base; Tl.block_bytes = block_bytes;
    Tl.async_ok = (ru.mode != 0 && ra.mode != 0) ? 1 : 0;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = track_vo(pl, &Tl.meta, meta);
    if (e == cudaSuccess) e = track_vo(pl, &Tl.blocks, blocks);
    if (e == cudaSuccess) {
        pl->fused_smem = smem;
        Tl.ok = 1;
    }
    return e;
}


// K_ff(a) as CSR over the free dofs with, per stored entry, the list of (conductivity input, element-matrix entry) terms:
// K_ij(a) = sum_t a[term_in[t]] * term_coef[t].  Used by the batched posterior update (vo_posterior.cuh).
static cudaError_t build_csr_terms(gpde_vo_plan *pl, int n_cells, const int32_t *cell_dofs, const double *Ke,
                                   const int32_t *cell_to_input, int d, const std::vector<int> &src) {
    struct Term { int in; double coef; };
    std::vector<std::vector<std::pair<int, std::vector<Term>>>> rows(d);
    for (int c = 0; c < n_cells; ++c)
        for (int l = 0; l < 3; ++l) {
            const int i = src[cell_dofs[3 * c + l]];
            if (i < 0) continue;
            for (int l2 = 0; l2 < 3; ++l2) {
                const int j = src[cell_dofs[3 * c + l2]];
                const double v = Ke[9 * c + 3 * l + l2];
                if (j < 0 || v == 0.0) continue;
                std::vector<Term> *hit = nullptr;
                for (auto &e : rows[i])
                    if (e.first == j) hit = &e.second;
                if (!hit) {
                    rows[i].push_back({j, {}});
                    hit = &rows[i].back().second;
                }
                hit->push_back(Term{cell_to_input[c], v});
            }
        }
    std::vector<int> row_ptr(d + 1, 0), col, term_ptr(1, 0), term_in;
    std::vector<double> term_coef;
    int max_row = 1;
    for (int i = 0; i < d; ++i) {
        std::sort(rows[i].begin(), rows[i].end(), [](const auto &x, const auto &y) { return x.first < y.first; });
        for (auto &e : rows[i]) {
            col.push_back(e.first);
            for (const Term &t : e.second) {
                term_in.push_back(t.in);
                term_coef.push_back(t.coef);
            }
            term_ptr.push_back((int)term_in.size());
        }
        row_ptr[i + 1] = (int)col.size();
        max_row = std::max(max_row, row_ptr[i + 1] - row_ptr[i]);
    }
    VoCsrDev &C = pl->csr;
    C.d = d; C.nnz = (int)col.size(); C.max_row = max_row;
    cudaError_t e = track_vo(pl, &C.row_ptr, row_ptr);
    if (e == cudaSuccess) e = track_vo(pl, &C.col, col);
    if (e == cudaSuccess) e = track_vo(pl, &C.term_ptr, term_ptr);
    if (e == cudaSuccess) e = track_vo(pl, &C.term_in, term_in);
    if (e == cudaSuccess) e = track_vo(pl, &C.term_coef, term_coef);
    return e;
}

// Structured-grid detection for vo_grid.cuh.  Leaves pl->grid.ok = 0 unless the element data is exactly the
// 5-point pixel operator the grid kernel evaluates (every check below is against the arrays the caller
// passed, nothing is assumed from names):
//   nodes (nx+1) x (ny+1) numbered x-fastest; Dirichlet dofs = the columns ix = 0 and ix = nx, listed row by
//   row (left, right); free dofs = all other nodes ascending; two cells per pixel sharing one input entry
//   in0 + cy*sy + cx; horizontal / vertical couplings uniform (chs / cvs), hypotenuse couplings zero.
static cudaError_t build_grid_plan(gpde_vo_plan *pl, int n_nodes, int n_cells, const int32_t *cell_dofs,
                                   const double *Ke, const int32_t *cell_to_input, int n_inputs,
                                   const int64_t *free_dofs, int d, const int64_t *bc_dofs, int n_bc,
                                   const std::vector<double> &f_free) {
    GridDev &G = pl->grid;
    memset(&G, 0, sizeof(G));
    if (n_bc < 4 || (n_bc & 1) || bc_dofs[0] != 0) return cudaSuccess;
    const long long nxn = bc_dofs[1] + 1;               // nodes per row
    if (nxn < 3 || n_nodes % nxn != 0) return cudaSuccess;
    const int nx = (int)nxn - 1, nyn = (int)(n_nodes / nxn), ny = nyn - 1, ncol = nx - 1;
    if (ny < 1 || n_bc != 2 * nyn || d != ncol * nyn || n_cells != 2 * nx * ny || n_inputs != nx * ny) return cudaSuccess;
    if ((nx & 1) || ncol > 256) return cudaSuccess;       // pixel rows must be whole 16-byte units
    for (int r = 0; r < nyn; ++r)
        if (bc_dofs[2 * r] != (long long)r * nxn || bc_dofs[2 * r + 1] != (long long)r * nxn + nx) return cudaSuccess;
    for (int i = 0; i < d; ++i)
        if (free_dofs[i] != (long long)(i / ncol) * nxn + 1 + i % ncol) return cudaSuccess;
    double kmax = 0.0;
    for (size_t k = 0; k < (size_t)n_cells * 9; ++k) kmax = std::max(kmax, fabs(Ke[k]));
    if (!(kmax > 0.0)) return cudaSuccess;
    const double tol = 1e-13 * kmax;
    std::vector<int> sq_input((size_t)nx * ny, -1), sq_count((size_t)nx * ny, 0);
    // side coverage per square: bottom, top (horizontal), left, right (vertical)
    std::vector<unsigned char> side((size_t)nx * ny * 4, 0);
    double chs = 0.0, cvs = 0.0;
    bool have_h = false, have_v = false;
    for (int c = 0; c < n_cells; ++c) {
        int ix[3], iy[3];
        for (int l = 0; l < 3; ++l) {
            ix[l] = cell_dofs[3 * c + l] % (int)nxn;
            iy[l] = cell_dofs[3 * c + l] / (int)nxn;
        }
        const int cx = std::min(ix[0], std::min(ix[1], ix[2])), cy = std::min(iy[0], std::min(iy[1], iy[2]));
        if (cx >= nx || cy >= ny) return cudaSuccess;
        for (int l = 0; l < 3; ++l)
            if (ix[l] - cx > 1 || iy[l] - cy > 1) return cudaSuccess;
        const size_t sq = (size_t)cy * nx + cx;
        if (sq_count[sq] == 0) sq_input[sq] = cell_to_input[c];
        else if (sq_input[sq] != cell_to_input[c]) return cudaSuccess;   // per-cell input: generic kernels
        if (++sq_count[sq] > 2) return cudaSuccess;
        for (int l = 0; l < 3; ++l) {
            double rs = 0.0;
            for (int l2 = 0; l2 < 3; ++l2) rs += Ke[9 * c + 3 * l + l2];
            if (fabs(rs) > tol) return cudaSuccess;                     // not a pure diffusion operator
            for (int l2 = 0; l2 < 3; ++l2) {
                if (l2 == l) continue;
                const double v = Ke[9 * c + 3 * l + l2];
                if (fabs(v - Ke[9 * c + 3 * l2 + l]) > tol) return cudaSuccess;
                const int dx = ix[l2] - ix[l], dy = iy[l2] - iy[l];
                if (dx != 0 && dy != 0) {
                    if (fabs(v) > tol) return cudaSuccess;               // hypotenuse must not couple
                } else if (dy == 0) {
                    if (!have_h) { chs = v; have_h = true; }
                    if (fabs(v - chs) > tol) return cudaSuccess;
                    if (l < l2) side[sq * 4 + (iy[l] == cy ? 0 : 1)]++;
                } else {
                    if (!have_v) { cvs = v; have_v = true; }
                    if (fabs(v - cvs) > tol) return cudaSuccess;
                    if (l < l2) side[sq * 4 + 2 + (ix[l] == cx ? 0 : 1)]++;
                }
            }
        }
    }
    if (!have_h || !have_v || chs == 0.0 || cvs == 0.0) return cudaSuccess;
    for (size_t k = 0; k < side.size(); ++k)
        if (side[k] != 1) return cudaSuccess;                          // every pixel side carried by exactly one cell
    for (size_t sq = 0; sq < sq_count.size(); ++sq)
        if (sq_count[sq] != 2) return cudaSuccess;
    const long long in0 = sq_input[0];
    const long long sy = ny > 1 ? (long long)sq_input[nx] - in0 : nx;
    if (sy != nx && sy != -nx) return cudaSuccess;
    for (int cy = 0; cy < ny; ++cy)
        for (int cx = 0; cx < nx; ++cx)
            if (sq_input[(size_t)cy * nx + cx] != in0 + cy * sy + cx) return cudaSuccess;
    if ((in0 & 1) || (sy & 1)) return cudaSuccess;

    G.nx = nx; G.ny = ny; G.ncol = ncol; G.cols = 4; G.nstrips = 0; G.groups = 0;   // decomposition: grid_pick()
    G.in0 = in0; G.sy = sy; G.rh = chs / cvs; G.scale = cvs;
    G.has_load = 0;
    std::vector<double> f_over(d);
    for (int i = 0; i < d; ++i) {
        f_over[i] = f_free[i] / cvs;
        if (f_free[i] != 0.0) G.has_load = 1;
    }
    cudaError_t e = track_vo(pl, &G.f_over, f_over);
    if (e == cudaSuccess) G.ok = 1;
    return e;
}

// Shared-memory layout of one pipeline stage holding R node rows (see vo_grid.cuh): per-sample regions of
// a_stride / y_stride doubles (pitches chosen so that the 8 samples of a warp hit distinct banks), then the
// packed V rows.  Returns the stage size in bytes for NT n-tiles (NT = 0: no V in the stage).
static inline size_t grid_layout(GridDev &G, int R, int NT) {
    const int S = 8 * G.groups;
    G.a_stride = std::max(R * G.nx, (R - 1) * G.nx + 4 * G.cols * G.nstrips) + 2;
    G.y_stride = ((R - 1) * G.ncol + 4 * G.cols * G.nstrips + 4 + 3) & ~3;
    G.a_off = 0;
    G.y_off = S * G.a_stride * 8;
    G.v_off = (G.y_off + S * G.y_stride * 8 + 127) & ~127;
    return (size_t)G.v_off + (size_t)R * G.nstrips * G.cols * NT * 32 * 8;
}
static inline int grid_nstrips(int ncol, int cols) {
    int n = 1;
    while (n * 4 * cols < ncol) n *= 2;
    return n;
}
static inline size_t grid_packed_bytes(const GridDev &G, int NT) {
    return (size_t)(G.ny + 1) * grid_nstrips(G.ncol, 4) * 4 * NT * 32 * 8;
}
// rows per stage and ring depth that fit the 227 KB of a CTA: two rows per stage halve the per-row barrier and
// staging overhead (GPDE_GRID_R forces 1 or 2 for experiments)
// Decomposition: columns per lane C, warps per CTA W, rows per stage R, ring depth NS.  The kernel is templated on
// all of them; measured on B200 at cfg 2 (round 1):
//     C = 4, W = 16, R = 2 : 156 us   <- instantiated (16 warps of 128 registers, one CTA per SM)
//     C = 4, W = 16, R = 1 : 186 us   <- instantiated (used when two 2-row stages do not fit)
//     C = 4, W =  8, R = 1 : 174 us   (8 warps per CTA, two CTAs per SM)
//     C = 8, W =  8, R = 2 : 176 us   (8 fat warps, half the per-step overhead per node)
// The kernel is latency-bound at 4 warps per scheduler: more resident warps win over less overhead per node.
// GPDE_GRID_R=1 forces one row per stage (A/B runs).
static inline bool grid_pick(const VoEnv &env, GridDev &G, int NT, int &R, int &NS, int &W, size_t &stage) {
    const int C = 4;
    W = 16;
    if (grid_nstrips(G.ncol, C) > W) return false;
    G.cols = C;
    G.nstrips = grid_nstrips(G.ncol, C);
    G.groups = W / G.nstrips;
    const size_t budget = 225 * 1024 - 512;
    for (R = env.grid_r; R >= 1; --R) {
        stage = grid_layout(G, R, NT);
        NS = (int)std::min<size_t>(R == 2 ? 3 : 4, budget / stage);
        if (NS >= 2) return true;
    }
    return false;
}

// GPDE_VO_PATH=v1 forces the unfused version-1 kernels (A/B testing, fallback check)
static bool use_fused(const gpde_vo_plan *pl) {
    return pl->tiles.ok && !pl->env.path_v1;
}

// GPDE_VO_PATH=fused keeps the structured-grid kernel out as well (generic fused kernel instead)
static bool use_grid(const gpde_vo_plan *pl) {
    return pl->grid.ok && !pl->env.path_v1 && !pl->env.path_fused;
}

// Lean structured-grid kernel (vo_grid2.cuh): nx in {16, 32, 64, 128}, even ny, no load vector.  Returns 1 if it
// served the call, 0 if vo_grid.cuh should.  rho_pitch > 0 selects the rho variant (V, m unused).
// GPDE_GRID_V=1 keeps it out (A/B runs against the general kernel).
// mesh-side conditions and the decomposition of the lean kernel for m weighting functions (rho: no V inside)
static bool grid2_setup(const gpde_vo_plan *pl, int m, bool rho, int sub_f, Grid2Dev &G, int &NT, int &NX, size_t &smem,
                        int ea = 8, int ey = 8) {   // ea / ey: bytes per element of a / y (8 or 4)
    const GridDev &G0 = pl->grid;
    if (pl->env.no_grid2) return false;
    const int nx = G0.nx, ny = G0.ny;
    if (!G0.ok || !(nx == 16 || nx == 32 || nx == 64 || nx == 128) || ny < 2 || (ny & 1)) return false;
    if (G0.has_load && sub_f) return false;
    NT = 1; NX = 0;
    if (!rho) {
        if (m < 1 || m > 32) return false;
        NT = m >> 3; NX = m & 7;
        if (NX > 1 || NT == 0) { NT += NX ? 1 : 0; NX = 0; }
    }
    memset(&G, 0, sizeof(G));
    G.nx = nx; G.ny = ny; G.ncol = G0.ncol;
    for (G.lognx = 0; (1 << G.lognx) < nx; ++G.lognx) {}
    G.nstrips = nx / 16;
    for (G.lognstrips = 0; (1 << G.lognstrips) < G.nstrips; ++G.lognstrips) {}
    G.groups = 16 / G.nstrips;
    G.in0 = G0.in0; G.sy = G0.sy; G.rh = G0.rh; G.scale = G0.scale;
    const int S = 8 * G.groups;
    // pitches in elements, chosen against bank conflicts of the warp's reads (vo_grid2.cuh): FP32 rows need a_pitch = 16 mod 32
    // (LDS.128 of two samples per phase) and y_pitch = 4 mod 8 (the samples' phases then spread the banks)
    G.a_pitch = ea == 8 ? 2 * nx + 2 : 2 * nx + 16;
    G.y_pitch = ey == 8 ? 2 * nx + 8 : 2 * nx + 12;
    G.y_off = S * G.a_pitch * ea;
    G.v_off = 0;
    G.v_row_bytes = rho ? 0 : G.nstrips * (4 * NT * 32 * 8 + NX * 128) + kGrid2MaskBytes;
    G.stage_bytes = (G.y_off + S * G.y_pitch * ey + 127) & ~127;
    const size_t fixed = 2 * (size_t)G.stage_bytes + (4 * 16 + 8) * sizeof(unsigned long long) + 258 * sizeof(double);
    // V ring: three 2-row stages when they fit (the sample groups of a CTA may then drift a stage apart), else two
    G.nvs = rho ? 0 : ((fixed + 3 * 2 * (size_t)G.v_row_bytes <= 227 * 1024) ? 3 : 2);
    if (!rho && (pl->env.grid2_nvs == 2 || pl->env.grid2_nvs == 3)) G.nvs = pl->env.grid2_nvs;
    smem = fixed + (size_t)G.nvs * 2 * G.v_row_bytes;
    if (smem > 227 * 1024) return false;
    G.flags = pl->env.grid2_flags;
    G.spc = S;
    return true;
}

// Samples per CTA of the lean grid kernel.  A CTA holds S = 8 * groups sample slots and owns a whole SM; with
// ceil(B / S) CTAs the last wave leaves SMs idle (4096 samples, S = 32: 128 CTAs on 148 SMs).  The kernel is bound by
// the bytes an SM streams, so the same number of waves is spread over all SMs instead: the smallest spc with
// ceil(B / spc) <= waves * SMs (4096 samples: 147 CTAs of 28; a slot left empty costs an idle row of the 8-row DMMA
// tile, nothing else).  `sm_reserve` SMs are left to kernels the caller runs beside this one on other streams.
static int grid2_samples_per_cta(const gpde_vo_plan *pl, int S, long long B, int sm_reserve) {
    if (pl->env.grid2_spc > 0) return std::max(S / 2, std::min(S, pl->env.grid2_spc));
    const long long sms = std::max(1, pl->n_sm - std::max(0, sm_reserve));
    const long long blocks = (B + S - 1) / S, waves = (blocks + sms - 1) / sms;
    const long long spc = (B + waves * sms - 1) / (waves * sms);
    return (int)std::max<long long>((3 * S + 3) / 4, std::min<long long>(S, spc));
}

template <typename TV>
static void grid2_pack(const Grid2Dev &G, const TV *V, int m, int NT, int NX, double *Vp, int n_sm, cudaStream_t st) {
    const long long warps = (long long)(G.ny + 1) * 8;   // one warp per (node row, strip slot)
    const unsigned grid = (unsigned)std::min<long long>((warps + 3) / 4, (long long)n_sm * 16);
    vo_grid2_pack_kernel<TV><<<grid, 128, 0, st>>>(G, V, m, NT, NX, Vp);
}

// TA: conductivities, Dirichlet values and V; TY: y; TR: result (r, or the rho rows of the rho variant)
template <typename TA, typename TY, typename TR>
static int launch_grid2(const gpde_vo_plan *pl, const TA *a, long long a_stride, int a_is_log, const TY *y,
                        const TA *g, long long g_stride, const TA *V, int m, TR *r, void *workspace,
                        int rho_pitch, int sub_f, bool prepacked, long long B, cudaStream_t st, long long y_stride = 0,
                        int sm_reserve = 0) {
    constexpr int EA = (int)sizeof(TA), EY = (int)sizeof(TY);
    if (y_stride == 0) y_stride = pl->dev.d;
    if (y_stride != pl->dev.d && rho_pitch <= 0) return 0;   // only the rho variant is built for strided y
    const bool rho = rho_pitch > 0;
    Grid2Dev G;
    int NT, NX;
    size_t smem;
    if (!grid2_setup(pl, m, rho, sub_f, G, NT, NX, smem, EA, EY)) return 0;
    // 16-byte pieces of a: base, sample stride and the plan's pixel offsets; y only needs its natural alignment
    if (((uintptr_t)a & 15) || ((a_stride * EA) & 15) || ((G.in0 * EA) & 15) || ((G.sy * EA) & 15) || ((uintptr_t)y & (EY - 1))) return 0;
    if (!rho && ((uintptr_t)workspace & 15)) return 0;
    const int S = 8 * G.groups;
    double *Vp = (double *)workspace;
    if (!rho && !prepacked) grid2_pack(G, V, m, NT, NX, Vp, pl->n_sm, st);
    unsigned blocks = (unsigned)((B + S - 1) / S);
    // small batches: the node rows of a sample block are cut over the CTAs of a thread-block cluster (SPLIT variant of the
    // kernel) when the blocks alone would leave more than half of the SMs idle; every range gets at least two stages.
    // The cluster size depends on the number of sample blocks only.  GPDE_GRID2_SPLIT=0 keeps one CTA per block.
    int csize = 1;
    if (!rho && std::is_same<TA, TY>::value && std::is_same<TA, TR>::value && pl->env.grid2_split) {
        const int n_stages = G.ny / 2;
        csize = std::min(std::min(8, n_stages / 2), pl->n_sm / (int)blocks);
        if (pl->env.grid2_split > 1) csize = std::min(pl->env.grid2_split, n_stages / 2);   // forced size (tests)
        if (csize < 2 || 2 * (size_t)G.stage_bytes < (size_t)(16 * 8 * 34 + S * 33) * sizeof(double)) csize = 1;
    }
    if (csize == 1) {   // whole samples per CTA: balance the last wave over the SMs
        G.spc = grid2_samples_per_cta(pl, S, B, sm_reserve);
        blocks = (unsigned)((B + G.spc - 1) / G.spc);
    }
    const unsigned grid = blocks * (unsigned)csize;
    // the residual kernel is launched as a programmatic dependent of the packing kernel (its prologue and first
    // a / y stages overlap the packing); GPDE_GRID2_PDL=0 keeps the plain stream order
    const bool pdl = !rho && !prepacked && pl->env.grid2_pdl && csize == 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (csize > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)csize;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
    } else {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = (pdl || csize > 1) ? 1 : 0;
    const int m_arg = rho ? rho_pitch : m;
    // a_is_log is a template parameter: with the exp() path compiled in or out, each variant gets its own register
    // allocation (as a run-time branch the two paths cost each other 4-8 %, measured A/B)
#define GPDE_LAUNCH_GRID2_YS(NTV, NXV, RHOV, YSV)                                                                \
    {                                                                                                            \
        auto kern = a_is_log ? vo_grid2_kernel<NTV, NXV, RHOV, YSV, true, TA, TY, TR>                           \
                             : vo_grid2_kernel<NTV, NXV, RHOV, YSV, false, TA, TY, TR>;                         \
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        GPDE_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, G, a, a_stride, a_is_log, y, y_stride, g, g_stride,         \
                                        (const double *)Vp, m_arg, r, B));                                       \
    }
#define GPDE_LAUNCH_GRID2_SPLIT(NTV, NXV)                                                                        \
    {                                                                                                            \
        auto kern = a_is_log ? vo_grid2_kernel<NTV, NXV, false, false, true, TA, TY, TR, true>                   \
                             : vo_grid2_kernel<NTV, NXV, false, false, false, TA, TY, TR, true>;                 \
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        GPDE_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, G, a, a_stride, a_is_log, y, y_stride, g, g_stride,         \
                                        (const double *)Vp, m_arg, r, B));                                       \
    }
#define GPDE_LAUNCH_GRID2(NTV, NXV, RHOV)                                                                        \
    {                                                                                                            \
        if (csize > 1) { if constexpr (!RHOV && same) GPDE_LAUNCH_GRID2_SPLIT(NTV, NXV) }                        \
        else GPDE_LAUNCH_GRID2_YS(NTV, NXV, RHOV, false)                                                         \
    }
    // instantiated type combinations: (T,T,T) every variant but the strided-y one for FP32; (float,float,double) the rho
    // variant only (rho rows for the FP64 contraction); (T,double,T) the strided-y rho variant only (residual_T)
    constexpr bool same = std::is_same<TA, TY>::value && std::is_same<TA, TR>::value;
    constexpr bool ys_ok = std::is_same<TY, double>::value && std::is_same<TA, TR>::value;
    constexpr bool rho_ok = std::is_same<TA, TY>::value;
    if (rho && y_stride != pl->dev.d) {
        if constexpr (ys_ok) GPDE_LAUNCH_GRID2_YS(1, 0, true, true)
        else return 0;
    } else if (rho) {
        if constexpr (rho_ok) GPDE_LAUNCH_GRID2(1, 0, true)
        else return 0;
    } else if constexpr (same) {
        if (NT == 1 && NX == 0) GPDE_LAUNCH_GRID2(1, 0, false)
        else if (NT == 1) GPDE_LAUNCH_GRID2(1, 1, false)
        else if (NT == 2 && NX == 0) GPDE_LAUNCH_GRID2(2, 0, false)
        else if (NT == 2) GPDE_LAUNCH_GRID2(2, 1, false)
        else if (NT == 3 && NX == 0) GPDE_LAUNCH_GRID2(3, 0, false)
        else if (NT == 3) GPDE_LAUNCH_GRID2(3, 1, false)
        else GPDE_LAUNCH_GRID2(4, 0, false)
    } else {
        return 0;
    }
#undef GPDE_LAUNCH_GRID2
#undef GPDE_LAUNCH_GRID2_SPLIT
#undef GPDE_LAUNCH_GRID2_YS
    GPDE_CUDA_OK(cudaGetLastError());
    return 1;
}

// Transposed application q = K_ff(a) (V s) on the reference's pixel meshes in ONE kernel (WT variant of vo_grid2_kernel):
// V^T packed in fragment order (vo_grid2_pack_t_kernel, 4 us), then the marching kernel produces the rows of w = s V^T
// in shared memory itself.  Returns 1 if it served the call, 0 if the two-kernel route should (mesh / alignment / shared
// memory), <0 on error.  s, V and a share the element type T; q too.
template <typename T>
static int launch_grid2_wt(const gpde_vo_plan *pl, const T *a, long long a_stride, int a_is_log, const T *V, int mw, const T *s,
                           T *q, void *workspace, long long B, cudaStream_t st) {
    constexpr int EA = (int)sizeof(T);
    if (!pl->env.fused_t || mw < 1 || mw > 32) return 0;
    Grid2Dev G;
    int NT, NX;
    size_t smem;
    if (!grid2_setup(pl, 0, true, 0, G, NT, NX, smem, EA, 8)) return 0;
    if (((uintptr_t)a & 15) || ((a_stride * EA) & 15) || ((G.in0 * EA) & 15) || ((G.sy * EA) & 15) || ((uintptr_t)workspace & 15)) return 0;
    const int KS = mw <= 16 ? 4 : (mw <= 28 ? 7 : 8);
    G.v_row_bytes = G.nstrips * 2 * KS * 256 + kGrid2MaskBytes;
    const size_t fixed = smem;                  // (rho variant: no V stages counted yet)
    G.nvs = (fixed + 3 * 2 * (size_t)G.v_row_bytes <= 227 * 1024) ? 3 : 2;
    if (pl->env.grid2_nvs == 2 || pl->env.grid2_nvs == 3) G.nvs = pl->env.grid2_nvs;
    smem = fixed + (size_t)G.nvs * 2 * G.v_row_bytes;
    if (smem > 227 * 1024) return 0;
    const int S = 8 * G.groups, d = pl->dev.d;
    double *Vp = (double *)workspace;
    {
        const long long warps = (long long)(G.ny + 1) * 8;   // one warp per (node row, strip slot)
        const unsigned grid = (unsigned)std::min<long long>((warps + 3) / 4, (long long)pl->n_sm * 16);
        vo_grid2_pack_t_kernel<T><<<grid, 128, 0, st>>>(G, V, mw, KS, Vp);
    }
    G.spc = grid2_samples_per_cta(pl, S, B, 0);
    const unsigned blocks = (unsigned)((B + G.spc - 1) / G.spc);
#define GPDE_LAUNCH_GRID2_WT(KSV)                                                                                \
    {                                                                                                            \
        auto kern = a_is_log ? vo_grid2_kernel<1, 0, true, false, true, T, double, T, false, KSV>                \
                             : vo_grid2_kernel<1, 0, true, false, false, T, double, T, false, KSV>;              \
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        kern<<<blocks, 512, smem, st>>>(G, a, a_stride, a_is_log, (const double *)nullptr, 0ll, s, (long long)mw, \
                                        (const double *)Vp, d, q, B);                                            \
    }
    if (KS == 4) GPDE_LAUNCH_GRID2_WT(4)
    else if (KS == 7) GPDE_LAUNCH_GRID2_WT(7)
    else GPDE_LAUNCH_GRID2_WT(8)
#undef GPDE_LAUNCH_GRID2_WT
    GPDE_CUDA_OK(cudaGetLastError());
    return 1;
}

// Structured-grid path.  Returns 1 if it served the call, 0 if the caller should use the generic kernels (alignment /
// size conditions not met), <0 on error.  FP32 I/O: the lean kernel only.
template <typename T>
static int launch_grid_lean(const gpde_vo_plan *pl, const T *a, long long a_stride, int a_is_log, const T *y,
                            const T *g, long long g_stride, const T *V, int m, T *r, void *workspace,
                            int sub_f, bool prepacked, long long B, cudaStream_t st, int sm_reserve = 0) {
    if (!y || m < 1 || m > 32) return 0;
    const int rc2 = launch_grid2<T, T, T>(pl, a, a_stride, a_is_log, y, g, g_stride, V, m, r, workspace, 0, sub_f, prepacked, B, st, 0,
                                          sm_reserve);
    if (rc2 != 0) return rc2;
    if (prepacked)
        return fail(GPDE_ERR_ARG, "vo_residual: packed weights (flags bit1) need the lean structured-grid kernel for this call "
                                  "(16-byte aligned a and sample stride, y given, no load vector)");
    return 0;
}
static int launch_grid(const gpde_vo_plan *pl, const double *a, long long a_stride, int a_is_log, const double *y,
                       const double *g, long long g_stride, const double *V, int m, double *r, void *workspace,
                       int sub_f, bool prepacked, long long B, cudaStream_t st, int sm_reserve = 0) {
    GridDev G = pl->grid;
    if (!y || m < 1 || m > 32) return 0;
    {
        const int rc2 = launch_grid_lean<double>(pl, a, a_stride, a_is_log, y, g, g_stride, V, m, r, workspace, sub_f, prepacked, B, st,
                                                 sm_reserve);
        if (rc2 != 0) return rc2;
    }
    if (((uintptr_t)a & 15) || ((uintptr_t)y & 15) || (a_stride & 1) || ((uintptr_t)workspace & 15)) return 0;
    const int NT = m <= 8 ? 1 : (m <= 16 ? 2 : 4);
    int R, NS, W;
    size_t stage;
    if (!grid_pick(pl->env, G, NT, R, NS, W, stage)) return 0;
    if (!sub_f) G.has_load = 0;
    double *Vp = (double *)workspace;
    {
        const long long total = (long long)grid_packed_bytes(G, NT) / 8;
        const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)pl->n_sm * 8);
        vo_grid_pack_kernel<<<grid, 256, 0, st>>>(G, V, m, NT, Vp);
    }
    const int S = 8 * G.groups;
    const int dbg = pl->env.grid_debug;   // timing experiments only
    const unsigned grid = (unsigned)((B + S - 1) / S);
    const size_t smem = (size_t)NS * stage + 2 * NS * sizeof(unsigned long long) + 16 * sizeof(double);
#define GPDE_LAUNCH_GRID(NTV, RV, WV, CV)                                                                        \
    {                                                                                                            \
        auto kern = vo_grid_kernel<NTV, false, RV, WV, CV>;                                                      \
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        kern<<<grid, WV * 32, smem, st>>>(G, a, a_stride, a_is_log, y, g, g_stride, Vp, m, r, B, NS, (int)stage, dbg); \
    }
#define GPDE_LAUNCH_GRID_NT(RV, WV, CV)                                                                          \
    {                                                                                                            \
        if (NT == 1) GPDE_LAUNCH_GRID(1, RV, WV, CV)                                                             \
        else if (NT == 2) GPDE_LAUNCH_GRID(2, RV, WV, CV)                                                        \
        else GPDE_LAUNCH_GRID(4, RV, WV, CV)                                                                     \
    }
    if (R == 2) GPDE_LAUNCH_GRID_NT(2, 16, 4)
    else GPDE_LAUNCH_GRID_NT(1, 16, 4)
#undef GPDE_LAUNCH_GRID_NT
#undef GPDE_LAUNCH_GRID
    GPDE_CUDA_OK(cudaGetLastError());
    return 1;
}

// Structured-grid fine residual rho[B][pitch] (pitch >= d, padding zeroed) for the tensor-core contraction.
// Returns 1 if it served the call, 0 if the generic matvec kernel should (alignment / size), <0 on error.
template <typename T, typename TR>
static int launch_grid_rho_lean(const gpde_vo_plan *pl, const T *a, long long a_stride, int a_is_log, const T *y,
                                const T *g, long long g_stride, TR *rho, int pitch, int sub_f, long long B, cudaStream_t st) {
    if (!y) return 0;
    return launch_grid2<T, T, TR>(pl, a, a_stride, a_is_log, y, g, g_stride, (const T *)nullptr, 0, rho, nullptr, pitch, sub_f,
                                  false, B, st);
}
static int launch_grid_rho(const gpde_vo_plan *pl, const double *a, long long a_stride, int a_is_log, const double *y,
                           const double *g, long long g_stride, double *rho, int pitch, int sub_f, long long B,
                           cudaStream_t st) {
    GridDev G = pl->grid;
    if (!y) return 0;
    {
        const int rc2 = launch_grid_rho_lean<double, double>(pl, a, a_stride, a_is_log, y, g, g_stride, rho, pitch, sub_f, B, st);
        if (rc2 != 0) return rc2;
    }
    if (((uintptr_t)a & 15) || ((uintptr_t)y & 15) || (a_stride & 1)) return 0;
    int R, NS, W;
    size_t stage;                                     // no V rows in the stage
    if (!grid_pick(pl->env, G, 0, R, NS, W, stage)) return 0;
    if (!sub_f) G.has_load = 0;
    const int S = 8 * G.groups;
    const unsigned grid = (unsigned)((B + S - 1) / S);
    const size_t smem = (size_t)NS * stage + 2 * NS * sizeof(unsigned long long) + 16 * sizeof(double);
#define GPDE_LAUNCH_RHO(RV, WV, CV)                                                                              \
    {                                                                                                            \
        auto kern = vo_grid_kernel<1, true, RV, WV, CV>;                                                         \
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        kern<<<grid, WV * 32, smem, st>>>(G, a, a_stride, a_is_log, y, g, g_stride, nullptr, pitch, rho, B, NS, (int)stage, 0); \
    }
    if (R == 2) GPDE_LAUNCH_RHO(2, 16, 4) else GPDE_LAUNCH_RHO(1, 16, 4)
#undef GPDE_LAUNCH_RHO
    GPDE_CUDA_OK(cudaGetLastError());
    return 1;
}

template <typename Ta, typename Ty, typename To>
static int launch_matvec(const gpde_vo_plan *pl, const Ta *a, long long a_stride, int a_is_log, const Ty *y,
                         const Ta *g, long long g_stride, int sub_f, double *rho_ws, int ws_stride, To *rho_out,
                         long long B, cudaStream_t st) {
    const VoDev &P = pl->dev;
    const size_t per = sizeof(double) * (size_t)P.n_inputs;
    const size_t cap = 200 * 1024;
    const int nsm = pl->n_sm;
    int S = 1;
    if (B >= 4LL * nsm * 2 && 4 * per <= cap) S = 4;
    else if (B >= 2LL * nsm * 2 && 2 * per <= cap) S = 2;
    const int stage = (S * per <= cap) ? 1 : 0;
    const size_t smem = stage ? S * per : 0;
    const long long ctas = (B + S - 1) / S;
    const unsigned grid = (unsigned)std::min<long long>(ctas, (long long)nsm * 8);
#define GPDE_LAUNCH_MV(SS)                                                                               \
    {                                                                                                    \
        auto kern = vo_matvec_kernel<Ta, Ty, To, SS>;                                                    \
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, kVoThreads, smem, st>>>(P, a, a_stride, a_is_log, y, g, g_stride, sub_f, rho_ws,    \
                                             ws_stride, rho_out, B, stage);                               \
    }
    if (S == 4) GPDE_LAUNCH_MV(4)
    else if (S == 2) GPDE_LAUNCH_MV(2)
    else GPDE_LAUNCH_MV(1)
#undef GPDE_LAUNCH_MV
    GPDE_CUDA_OK(cudaGetLastError());
    return GPDE_OK;
}

// Many weighting functions on the reference's pixel meshes: rho produced inside the contraction kernel (vo_gridgemm.cuh).
// Returns 1 if it served the call, 0 if the two-kernel route should, < 0 on error.
static bool gridgemm_setup(const gpde_vo_plan *pl, int m, int sub_f, int elem, GGDev &G, int &bn, int &nw, size_t &smem) {
    Grid2Dev G2;
    int nt2, nx2;
    size_t smem2;
    if (!pl->env.gridgemm || m <= 32 || !use_grid(pl) || !grid2_setup(pl, 0, true, sub_f, G2, nt2, nx2, smem2, elem, elem)) return false;
    memset(&G, 0, sizeof(G));
    G.nx = G2.nx; G.ny = G2.ny; G.ncol = G2.ncol; G.nstrips = G2.nstrips;
    G.in0 = G2.in0; G.sy = G2.sy; G.rh = G2.rh; G.scale = G2.scale;
    bn = m <= 128 ? 128 : 256;
    G.ctiles = (m + bn - 1) / bn;
    nw = 8;   // node columns per producer warp: 4 producer warps (8 warps of 4 columns measured slower: 4.46 vs 4.26 ms at config 3)
    for (G.stages_b = 4; G.stages_b >= 3; --G.stages_b)
        if ((smem = gg_smem_bytes(bn, G.stages_b, elem, nw)) <= 227 * 1024) return true;
    return false;
}
static size_t gridgemm_workspace_bytes(const gpde_vo_plan *pl, long long B, int m) {
    GGDev G;
    int bn, nw;
    size_t smem;
    if (!gridgemm_setup(pl, m, 0, 8, G, bn, nw, smem)) return 0;
    // packed V | partial tiles of the split contraction | conductivities of a log-field input
    return gg_packed_bytes(bn, G.ctiles, G.nstrips * (G.ny + 1)) + 32 + sizeof(double) * 8 * (size_t)B * G.ctiles * bn +
           sizeof(double) * (size_t)B * pl->dev.n_inputs;
}
template <typename T>
static int launch_gridgemm(const gpde_vo_plan *pl, const T *a, long long a_stride, int a_is_log, const T *y, const T *g,
                           long long g_stride, const T *V, int m, T *r, void *workspace, int sub_f, long long B,
                           cudaStream_t st) {
    GGDev G;
    int bn, nw;
    size_t smem;
    if (((uintptr_t)workspace & 15) || !gridgemm_setup(pl, m, sub_f, (int)sizeof(T), G, bn, nw, smem)) return 0;
    // the producers address a warp's 32 samples by 32-bit byte offsets
    const long long span = 32 * std::max<long long>(std::max<long long>(pl->dev.d, a_stride), g_stride) + 2 * (long long)pl->dev.d;
    if (span * (long long)sizeof(T) >= (1ll << 31)) return 0;
    const int chunks = G.nstrips * (G.ny + 1);
    double *Vp = (double *)workspace;
    const long long tiles = (B + kGGBM - 1) / kGGBM * G.ctiles;
    int splits = pl->env.gemm_splits > 0 ? pl->env.gemm_splits : gemm_splits(tiles, pl->n_sm, chunks);
    splits = std::max(1, std::min(splits, G.nstrips));
    const int ldp = G.ctiles * bn;
    double *part0 = (double *)(((uintptr_t)Vp + gg_packed_bytes(bn, G.ctiles, chunks) + 15) & ~(uintptr_t)15);
    double *partial = splits > 1 ? part0 : nullptr;
    if (a_is_log) {   // conductivities by one streaming pass (behind the partial tiles in the workspace)
        T *cond = (T *)(((uintptr_t)(part0 + (size_t)8 * B * ldp) + 15) & ~(uintptr_t)15);
        const long long n = a_stride ? B * a_stride : (long long)pl->dev.n_inputs;
        if (a_stride && a_stride != pl->dev.n_inputs) return 0;
        vo_exp_rows_kernel<T><<<(unsigned)std::min<long long>((n + 255) / 256, (long long)pl->n_sm * 16), 256, 0, st>>>(a, n, cond);
        a = cond;
    }
    const dim3 grid((unsigned)((B + kGGBM - 1) / kGGBM), (unsigned)G.ctiles, (unsigned)splits);
    const unsigned pgrid = (unsigned)std::min<long long>(((long long)G.ctiles * chunks + 7) / 8, (long long)pl->n_sm * 8);
#define GPDE_LAUNCH_GG(BNV, NWV)                                                                                  \
    {                                                                                                             \
        vo_gridgemm_pack_kernel<BNV, T><<<pgrid, 256, 0, st>>>(G, V, m, Vp);                                      \
        auto kern = vo_gridgemm_kernel<BNV, NWV, T, T>;                                                           \
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        kern<<<grid, gg_threads(NWV), smem, st>>>(G, a, a_stride, y, g, g_stride, Vp, r, m, ldp, B, partial);     \
    }
    if (bn == 128) GPDE_LAUNCH_GG(128, 8) else GPDE_LAUNCH_GG(256, 8)
#undef GPDE_LAUNCH_GG
    GPDE_CUDA_OK(cudaGetLastError());
    if (splits > 1) {
        const long long total = B * m;
        const unsigned rgrid = (unsigned)std::min<long long>((total + 255) / 256, (long long)pl->n_sm * 16);
        vo_gemm_reduce_kernel<T><<<rgrid, 256, 0, st>>>(partial, splits, ldp, r, m, B);
        GPDE_CUDA_OK(cudaGetLastError());
    }
    return 1;
}

template <typename T>
static int vo_residual(const gpde_vo_plan *pl, const T *a, int64_t a_stride, int a_is_log, const T *y, const T *g,
                       int64_t g_stride, const T *V, int m, T *r, T *rho, void *workspace, int flags, int64_t B,
                       gpde_stream_t stream) {
    if (!pl || B < 0 || m < 0) return fail(GPDE_ERR_ARG, "vo_residual: bad argument");
    if (B == 0) return GPDE_OK;
    if (!a) return fail(GPDE_ERR_ARG, "vo_residual: null conductivity field");
    if (m > 0 && (!V || !r)) return fail(GPDE_ERR_ARG, "vo_residual: V and r are required when m > 0");
    if (m > 0 && !workspace) return fail(GPDE_ERR_ARG, "vo_residual: workspace required");
    if (m == 0 && !rho) return fail(GPDE_ERR_ARG, "vo_residual: nothing to compute");
    if (B == 0) return GPDE_OK;
    DeviceGuard guard(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (m > 0 && !rho && use_grid(pl)) {
        int rc;
        if constexpr (sizeof(T) == 8)
            rc = launch_grid(pl, (const double *)a, (long long)a_stride, a_is_log, (const double *)y,
                             (const double *)g, (long long)g_stride, (const double *)V, m, (double *)r,
                             workspace, (flags & 1) ? 0 : 1, (flags & 2) != 0, (long long)B, st, (flags >> 8) & 0xff);
        else
            rc = launch_grid_lean<T>(pl, a, (long long)a_stride, a_is_log, y, g, (long long)g_stride, V, m, r, workspace,
                                     (flags & 1) ? 0 : 1, (flags & 2) != 0, (long long)B, st, (flags >> 8) & 0xff);
        if (rc != 0) return rc < 0 ? rc : GPDE_OK;
    }
    if (m == 0 && rho && use_grid(pl)) {   // the fine residual alone: rho rows straight into the caller's [B,d]
        const int rc = launch_grid_rho_lean<T, T>(pl, a, (long long)a_stride, a_is_log, y, g, (long long)g_stride, rho, pl->dev.d,
                                                  (flags & 1) ? 0 : 1, (long long)B, st);
        if (rc != 0) return rc < 0 ? rc : GPDE_OK;
    }
    if (flags & 2) return fail(GPDE_ERR_ARG, "vo_residual: packed weights (flags bit1) are not usable for this call");
    if (m > 0 && m <= 32 && use_fused(pl)) {
        const unsigned grid = (unsigned)((B + kFS - 1) / kFS);
        const int sub_f = (flags & 1) ? 0 : 1;
        constexpr bool kCanAsync = sizeof(T) == 8;
        const bool use_async = kCanAsync && pl->tiles.async_ok && !pl->env.sync_staging;
#define GPDE_LAUNCH_FUSED(WN)                                                                                  \
    {                                                                                                          \
        auto kern = use_async ? vo_fused_kernel<T, WN, kCanAsync> : vo_fused_kernel<T, WN, false>;             \
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->fused_smem)); \
        kern<<<grid, kFT, pl->fused_smem, st>>>(pl->dev, pl->tiles, a, (long long)a_stride, a_is_log, y, g,   \
                                                (long long)g_stride, V, m, r, rho, sub_f, (long long)B);       \
    }
        if (m <= 8) GPDE_LAUNCH_FUSED(1)
        else if (m <= 16) GPDE_LAUNCH_FUSED(2)
        else GPDE_LAUNCH_FUSED(4)
#undef GPDE_LAUNCH_FUSED
        GPDE_CUDA_OK(cudaGetLastError());
        return GPDE_OK;
    }
    if (m > 32 && !rho && use_grid(pl)) {   // many weighting functions on a pixel mesh: one kernel, rho never leaves the SM
        const int rc = launch_gridgemm<T>(pl, a, (long long)a_stride, a_is_log, y, g, (long long)g_stride, V, m, r, workspace,
                                          (flags & 1) ? 0 : 1, (long long)B, st);
        if (rc != 0) return rc < 0 ? rc : GPDE_OK;
    }
    // version 1: rho -> K-padded workspace, then the FP64 tensor-core contraction (vo_gemm.cuh)
    const int d = pl->dev.d, dp = gemm_dp(d);
    double *ws = m > 0 ? (double *)workspace : nullptr;
    int rc = 0;
    if (m > 0 && !rho && use_grid(pl)) {   // structured pixel grid: the marching kernel produces rho (no V inside)
        if constexpr (sizeof(T) == 8)
            rc = launch_grid_rho(pl, (const double *)a, (long long)a_stride, a_is_log, (const double *)y, (const double *)g,
                                 (long long)g_stride, ws, dp, (flags & 1) ? 0 : 1, (long long)B, st);
        else
            rc = launch_grid_rho_lean<T, double>(pl, a, (long long)a_stride, a_is_log, y, g, (long long)g_stride, ws, dp,
                                                 (flags & 1) ? 0 : 1, (long long)B, st);
        if (rc < 0) return rc;
    }
    if (rc == 0) {
        rc = launch_matvec<T, T, T>(pl, a, a_stride, a_is_log, y, g, g_stride, (flags & 1) ? 0 : 1, ws, dp, rho, B, st);
        if (rc != GPDE_OK) return rc;
    }
    if (m > 0) {
        const int bn = gemm_bn(m), ldb = gemm_ldb(m);
        double *Vp = ws + (size_t)B * dp;
        {
            const long long total = (long long)dp * ldb;
            const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)pl->n_sm * 8);
            vo_gemm_pack_kernel_t<T><<<grid, 256, 0, st>>>(V, d, m, Vp, dp, ldb);
        }
        dim3 grid((unsigned)((B + kGemmBM - 1) / kGemmBM), (unsigned)(ldb / bn));
        // tail of the tile grid on the SMs: parts of the contraction length with a deterministic reduction (vo_gemm.cuh)
        const int splits = pl->env.gemm_splits > 0 ? pl->env.gemm_splits
                                                   : gemm_splits((long long)grid.x * grid.y, pl->n_sm, dp / kGemmKC);
        double *partial = splits > 1 ? Vp + (size_t)dp * ldb : nullptr;
        grid.z = (unsigned)splits;
        const size_t smem = gemm_smem(bn);
        if (bn == 64) {
            auto kern = vo_gemm_kernel<64, T>;
            GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, kGemmThreads, smem, st>>>(ws, dp, Vp, ldb, r, m, (long long)B, partial);
        } else {
            auto kern = vo_gemm_kernel<128, T>;
            GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, kGemmThreads, smem, st>>>(ws, dp, Vp, ldb, r, m, (long long)B, partial);
        }
        GPDE_CUDA_OK(cudaGetLastError());
        if (splits > 1) {
            const long long total = (long long)B * m;
            const unsigned rgrid = (unsigned)std::min<long long>((total + 255) / 256, (long long)pl->n_sm * 16);
            vo_gemm_reduce_kernel<T><<<rgrid, 256, 0, st>>>(partial, splits, ldb, r, m, (long long)B);
            GPDE_CUDA_OK(cudaGetLastError());
        }
    }
    return GPDE_OK;
}

template <typename T>
static int vo_residual_T(const gpde_vo_plan *pl, const T *a, int64_t a_stride, int a_is_log, const T *V, int m,
                         const T *s, T *q, void *workspace, int64_t B, gpde_stream_t stream) {
    if (!pl || m <= 0 || B < 0) return fail(GPDE_ERR_ARG, "vo_residual_T: bad argument");
    if (B == 0) return GPDE_OK;
    if (!a || !V || !s || !q || !workspace) return fail(GPDE_ERR_ARG, "vo_residual_T: null argument");
    DeviceGuard guard(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (use_grid(pl) && m <= 32) {   // one kernel: w = s V^T never leaves the SM
        const int rc = launch_grid2_wt<T>(pl, a, (long long)a_stride, a_is_log, V, m, s, q, workspace, (long long)B, st);
        if (rc != 0) return rc < 0 ? rc : GPDE_OK;
    }
    if constexpr (sizeof(T) == 4) {
        // FP32 I/O on the reference's pixel meshes, m <= 32: the same two launches as FP64 (s, V converted on load by the
        // expansion kernel; w stays FP64 workspace; the marching kernel stages a as floats and stores q as floats)
        Grid2Dev G2;
        int nt2, nx2;
        size_t smem2;
        if (use_grid(pl) && m <= 32 && !((uintptr_t)a & 15) && !((a_stride * 4) & 15) && !((uintptr_t)workspace & 15) &&
            grid2_setup(pl, 0, true, 0, G2, nt2, nx2, smem2, 4, 8)) {
            const int d = pl->dev.d;
            double *w = (double *)workspace;
            const long long ldw = ((long long)d + 3) & ~3ll;
            const int n_tiles = (d + 7) / 8;
            const unsigned gx = (unsigned)((B + 8 * kExpandMT * kExpandWarps - 1) / (8 * kExpandMT * kExpandWarps));
            const int nch = std::max(1, std::min(n_tiles, (int)(4 * pl->n_sm / gx)));
            const int tpc = std::min((n_tiles + nch - 1) / nch, 64);
            const dim3 grid(gx, (unsigned)((n_tiles + tpc - 1) / tpc));
            const size_t smem = (size_t)tpc * 8 * kExpandVPitch * sizeof(double);
#define GPDE_LAUNCH_EXPAND_F(KSV)                                                                                   \
    {                                                                                                               \
        auto kern = vo_expand_dmma_kernel<KSV, float>;                                                              \
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        kern<<<grid, kExpandWarps * 32, smem, st>>>((const float *)s, (const float *)V, w, ldw, (long long)B, d, m, tpc); \
    }
            if (m <= 8) GPDE_LAUNCH_EXPAND_F(2)
            else if (m <= 16) GPDE_LAUNCH_EXPAND_F(4)
            else if (m <= 28) GPDE_LAUNCH_EXPAND_F(7)
            else GPDE_LAUNCH_EXPAND_F(8)
#undef GPDE_LAUNCH_EXPAND_F
            GPDE_CUDA_OK(cudaGetLastError());
            const int rc2 = launch_grid2<float, double, float>(pl, (const float *)a, (long long)a_stride, a_is_log, w, nullptr, 0,
                                                               nullptr, 0, (float *)q, nullptr, d, 0, false, (long long)B, st, ldw);
            if (rc2 != 0) return rc2 < 0 ? rc2 : GPDE_OK;
        }
    }
    if constexpr (sizeof(T) == 8) {
        // structured pixel grid: w = s V^T on the FP64 tensor pipe, then q = K_ff w with the marching kernel
        // (rho of u~ = (w, 0) without the load vector IS K_ff w)
        if (use_grid(pl) && !((uintptr_t)a & 15) && !(a_stride & 1) && !((uintptr_t)workspace & 15)) {
            const int d = pl->dev.d, mp = gemm_dp(m), bn = 128, ldb = (d + bn - 1) / bn * bn;
            double *Sp = (double *)workspace, *Vt = Sp + (size_t)B * mp, *w = Vt + (size_t)mp * ldb;
            if (m <= 32 && !pl->env.expand_gemm) {   // GPDE_VO_EXPAND=gemm: the general GEMM also for m <= 32 (A/B runs)
                // contraction length m <= 32: one launch, operands read in place (vo_expand.cuh)
                w = (double *)workspace;
                Grid2Dev G2;
                int nt2, nx2;
                size_t smem2;
                const bool lean = grid2_setup(pl, 0, true, 0, G2, nt2, nx2, smem2);
                const long long ldw = lean ? (((long long)d + 3) & ~3ll) : d;   // the general grid kernel wants y contiguous
                const int n_tiles = (d + 7) / 8;
                const unsigned gx = (unsigned)((B + 8 * kExpandMT * kExpandWarps - 1) / (8 * kExpandMT * kExpandWarps));
                const int nch = std::max(1, std::min(n_tiles, (int)(4 * pl->n_sm / gx)));
                const int tpc = std::min((n_tiles + nch - 1) / nch, 64);         // <= 64 n-tiles of V per CTA in shared memory
                const dim3 grid(gx, (unsigned)((n_tiles + tpc - 1) / tpc));
                const size_t smem = (size_t)tpc * 8 * kExpandVPitch * sizeof(double);
                const double *sd = (const double *)s, *Vd = (const double *)V;
#define GPDE_LAUNCH_EXPAND(KSV)                                                                                     \
    {                                                                                                               \
        auto kern = vo_expand_dmma_kernel<KSV>;                                                                     \
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        kern<<<grid, kExpandWarps * 32, smem, st>>>(sd, Vd, w, ldw, (long long)B, d, m, tpc);                       \
    }
                if (m <= 8) GPDE_LAUNCH_EXPAND(2)
                else if (m <= 16) GPDE_LAUNCH_EXPAND(4)
                else if (m <= 28) GPDE_LAUNCH_EXPAND(7)
                else GPDE_LAUNCH_EXPAND(8)
#undef GPDE_LAUNCH_EXPAND
                GPDE_CUDA_OK(cudaGetLastError());
                if (lean) {
                    const int rc2 = launch_grid2<double, double, double>(pl, (const double *)a, (long long)a_stride, a_is_log, w,
                                                                         nullptr, 0, nullptr, 0, (double *)q, nullptr, d, 0, false,
                                                                         (long long)B, st, ldw);
                    if (rc2 != 0) return rc2 < 0 ? rc2 : GPDE_OK;
                    return fail(GPDE_ERR_ARG, "vo_residual_T: lean grid kernel declined after setup");
                }
            } else {
            {
                const long long total = (long long)B * mp;
                const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)pl->n_sm * 8);
                vo_gemm_pad_rows_kernel<<<grid, 256, 0, st>>>((const double *)s, (long long)B, m, Sp, mp);
                vo_gemm_pack_transposed_kernel<<<dim3((unsigned)(ldb / 32), (unsigned)((mp + 31) / 32)), dim3(32, 8), 0, st>>>(
                    (const double *)V, d, m, Vt, mp, ldb);
            }
            {
                auto kern = vo_gemm_kernel<128, double>;
                const size_t smem = gemm_smem(bn);
                GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                dim3 grid((unsigned)((B + kGemmBM - 1) / kGemmBM), (unsigned)(ldb / bn));
                kern<<<grid, kGemmThreads, smem, st>>>(Sp, mp, Vt, ldb, w, d, (long long)B, (double *)nullptr);
                GPDE_CUDA_OK(cudaGetLastError());
            }
            }
            const int rc = launch_grid_rho(pl, (const double *)a, (long long)a_stride, a_is_log, w, nullptr, 0, (double *)q, d,
                                           0, (long long)B, st);
            if (rc != 0) return rc < 0 ? rc : GPDE_OK;
        }
    }
    if (use_fused(pl) && sizeof(double) * kFS * (size_t)m <= sizeof(double) * kFS * kRhoPitch) {
        const unsigned grid = (unsigned)((B + kFS - 1) / kFS);
        const size_t smem = pl->fused_smem;   // the coefficient vectors live in the (unused) rho tile
        auto kern = vo_fused_T_kernel<T>;
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kFT, smem, st>>>(pl->dev, pl->tiles, a, (long long)a_stride, a_is_log, V, m, s, q, (long long)B);
        GPDE_CUDA_OK(cudaGetLastError());
        return GPDE_OK;
    }
    double *w = (double *)workspace;
    const long long total = B * (long long)pl->dev.d;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)pl->n_sm * 16);
    vo_expand_kernel<T><<<grid, 256, 0, st>>>(s, V, w, B, pl->dev.d, m);
    GPDE_CUDA_OK(cudaGetLastError());
    return launch_matvec<T, double, T>(pl, a, a_stride, a_is_log, w, (const T *)nullptr, 0, 0, (double *)nullptr, 0,
                                       q, B, st);
}

}  // namespace gpde

using namespace gpde;

extern "C" {

int gpde_vo_plan_create(gpde_vo_plan **plan, int n_nodes, int n_cells, const int32_t *cell_dofs, const double *Ke,
                        const int32_t *cell_to_input, int n_inputs, const int64_t *free_dofs, int d,
                        const int64_t *bc_dofs, int n_bc, const double *f_full, int device) {
    if (!plan || !cell_dofs || !Ke || !cell_to_input || !free_dofs || n_nodes <= 0 || n_cells <= 0 ||
        n_inputs <= 0 || d <= 0 || n_bc < 0 || (n_bc > 0 && !bc_dofs))
        return fail(GPDE_ERR_ARG, "vo_plan_create: bad argument");
    std::vector<int> src(n_nodes, INT32_MIN);   // node -> encoded source
    for (int i = 0; i < d; ++i) {
        if (free_dofs[i] < 0 || free_dofs[i] >= n_nodes) return fail(GPDE_ERR_ARG, "vo_plan_create: free dof range");
        src[free_dofs[i]] = i;
    }
    for (int c = 0; c < n_bc; ++c) {
        if (bc_dofs[c] < 0 || bc_dofs[c] >= n_nodes) return fail(GPDE_ERR_ARG, "vo_plan_create: bc dof range");
        if (src[bc_dofs[c]] != INT32_MIN) return fail(GPDE_ERR_ARG, "vo_plan_create: dof both free and constrained");
        src[bc_dofs[c]] = -c - 1;
    }
    for (int v = 0; v < n_nodes; ++v)
        if (src[v] == INT32_MIN) return fail(GPDE_ERR_ARG, "vo_plan_create: node %d neither free nor constrained", v);
    // incident cells per free row
    std::vector<int> count(d, 0);
    for (int c = 0; c < n_cells; ++c) {
        if (cell_to_input[c] < 0 || cell_to_input[c] >= n_inputs)
            return fail(GPDE_ERR_ARG, "vo_plan_create: cell_to_input out of range");
        for (int l = 0; l < 3; ++l) {
            const int v = cell_dofs[3 * c + l];
            if (v < 0 || v >= n_nodes) return fail(GPDE_ERR_ARG, "vo_plan_create: cell dof out of range");
            if (src[v] >= 0) count[src[v]]++;
        }
    }
    int nslots = 0;
    for (int i = 0; i < d; ++i) nslots = std::max(nslots, count[i]);
    const size_t tot = (size_t)nslots * d;
    std::vector<int> ell_in(tot, 0), ell_j1(tot), ell_j2(tot);
    std::vector<double> c0(tot, 0.0), c1(tot, 0.0), c2(tot, 0.0);
    for (size_t k = 0; k < tot; ++k) ell_j1[k] = ell_j2[k] = (int)(k % d);   // padding: self, zero coefficients
    std::fill(count.begin(), count.end(), 0);
    for (int c = 0; c < n_cells; ++c)
        for (int l = 0; l < 3; ++l) {
            const int v = cell_dofs[3 * c + l];
            if (src[v] < 0) continue;
            const int i = src[v], t = count[i]++;
            const int l1 = (l + 1) % 3, l2 = (l + 2) % 3;
            const size_t k = (size_t)t * d + i;
            ell_in[k] = cell_to_input[c];
            ell_j1[k] = src[cell_dofs[3 * c + l1]];
            ell_j2[k] = src[cell_dofs[3 * c + l2]];
            c0[k] = Ke[9 * c + 3 * l + l];
            c1[k] = Ke[9 * c + 3 * l + l1];
            c2[k] = Ke[9 * c + 3 * l + l2];
        }
    std::vector<double> f_free(d, 0.0);
    if (f_full)
        for (int i = 0; i < d; ++i) f_free[i] = f_full[free_dofs[i]];

    gpde_vo_plan *pl = new gpde_vo_plan();   // (its VoEnv member reads the experiment switches here, once)
    pl->device = device;
    pl->n_sm = sm_count(device);
    DeviceGuard guard(device);
    VoDev &D = pl->dev;
    D.n_nodes = n_nodes; D.n_cells = n_cells; D.n_inputs = n_inputs; D.d = d; D.n_bc = n_bc; D.nslots = nslots;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = track_vo(pl, &D.ell_in, ell_in);
    if (e == cudaSuccess) e = track_vo(pl, &D.ell_j1, ell_j1);
    if (e == cudaSuccess) e = track_vo(pl, &D.ell_j2, ell_j2);
    if (e == cudaSuccess) e = track_vo(pl, &D.ell_c0, c0);
    if (e == cudaSuccess) e = track_vo(pl, &D.ell_c1, c1);
    if (e == cudaSuccess) e = track_vo(pl, &D.ell_c2, c2);
    if (e == cudaSuccess) e = track_vo(pl, &D.f_free, f_free);
    if (e == cudaSuccess) e = build_csr_terms(pl, n_cells, cell_dofs, Ke, cell_to_input, d, src);
    if (e == cudaSuccess) e = build_fused_tiles(pl, n_nodes, n_cells, cell_dofs, Ke, cell_to_input, d, src, f_free);
    if (e == cudaSuccess)
        e = build_grid_plan(pl, n_nodes, n_cells, cell_dofs, Ke, cell_to_input, n_inputs, free_dofs, d, bc_dofs, n_bc,
                            f_free);
    if (e != cudaSuccess) {
        gpde_vo_plan_destroy(pl);
        return fail(GPDE_ERR_CUDA, "vo_plan_create: upload failed: %s", cudaGetErrorString(e));
    }
    *plan = pl;
    return GPDE_OK;
}

int gpde_vo_plan_destroy(gpde_vo_plan *pl) {
    if (!pl) return GPDE_OK;
    DeviceGuard guard(pl->device);
    for (void *p : pl->allocs) cudaFree(p);
    delete pl;
    return GPDE_OK;
}

int gpde_vo_plan_info(const gpde_vo_plan *pl, int64_t out[8]) {
    if (!pl || !out) return fail(GPDE_ERR_ARG, "vo_plan_info: null");
    out[0] = pl->dev.n_nodes; out[1] = pl->dev.n_cells; out[2] = pl->dev.n_inputs; out[3] = pl->dev.d;
    out[4] = pl->dev.n_bc; out[5] = pl->dev.nslots; out[6] = pl->device;
    out[7] = pl->tiles.ok ? (int64_t)pl->fused_smem : 0;   // bytes of shared memory of the fused path (0 = unavailable)
    return GPDE_OK;
}

int gpde_vo_plan_kernel_path(const gpde_vo_plan *pl, int m, int elem_bytes) {
    if (!pl) return fail(GPDE_ERR_ARG, "vo_plan_kernel_path: null");
    if (elem_bytes == 8) {
        if (m > 0 && m <= 32 && use_grid(pl)) return 2;
        if (m > 32 && use_grid(pl)) return 3;
    } else if (use_grid(pl)) {   // FP32 I/O: the lean kernel only (the reference's own meshes)
        Grid2Dev G;
        int NT, NX;
        size_t smem;
        if (m > 0 && m <= 32 && grid2_setup(pl, m, false, 1, G, NT, NX, smem, 4, 4)) return 2;
        if (m > 32 && grid2_setup(pl, 0, true, 1, G, NT, NX, smem, 4, 4)) return 3;
    }
    if (m > 0 && m <= 32 && use_fused(pl)) return 1;
    return 0;
}

size_t gpde_vo_workspace_bytes(const gpde_vo_plan *pl, int64_t B, int m) {
    if (!pl || B < 0) return 0;
    // version-1 kernels: K-padded rho [B][dp] + padded V [dp][ldb] (residual), V s [B][d] (residual_T)
    const size_t dp = (size_t)gemm_dp(pl->dev.d);
    // (+ up to 8 partial result tiles [B][ldb] of the split contraction)
    size_t need = sizeof(double) * (dp * (size_t)B + dp * (size_t)gemm_ldb(std::max(m, 1)) +
                                    8 * (size_t)B * (size_t)gemm_ldb(std::max(m, 1)));
    if (pl->grid.ok && m > 0) {   // residual_T on the grid path: padded s [B][mp], V^T [mp][ldb], w [B][d]
        const size_t mp = (size_t)gemm_dp(m), ldb = ((size_t)pl->dev.d + 127) / 128 * 128;
        need = std::max(need, sizeof(double) * ((size_t)B * mp + mp * ldb + (size_t)B * pl->dev.d));
    }
    if (m > 32) need = std::max(need, gridgemm_workspace_bytes(pl, (long long)B, m));
    if (pl->grid.ok && m > 0 && m <= 32) {   // packed V (+ the per-row tile masks of the lean kernel)
        need = std::max(need, grid_packed_bytes(pl->grid, 4) + (size_t)(pl->grid.ny + 1) * kGrid2MaskBytes);
        // packed V^T rows of the one-kernel transposed application (launch_grid2_wt): <= 8 k-steps
        need = std::max(need, (size_t)(pl->grid.ny + 1) * ((size_t)((pl->grid.nx + 15) / 16) * 2 * 8 * 256 + kGrid2MaskBytes));
    }
    return need;
}

int gpde_vo_pack_weights_f64(const gpde_vo_plan *pl, const double *V, int m, int flags, void *workspace,
                             gpde_stream_t stream) {
    if (!pl || !V || !workspace) return fail(GPDE_ERR_ARG, "vo_pack_weights: null argument");
    if ((uintptr_t)workspace & 15) return fail(GPDE_ERR_ARG, "vo_pack_weights: workspace must be 16-byte aligned");
    Grid2Dev G;
    int NT, NX;
    size_t smem;
    if (!use_grid(pl) || !grid2_setup(pl, m, false, (flags & 1) ? 0 : 1, G, NT, NX, smem)) return 1;
    DeviceGuard guard(pl->device);
    grid2_pack(G, V, m, NT, NX, (double *)workspace, pl->n_sm, (cudaStream_t)stream);
    GPDE_CUDA_OK(cudaGetLastError());
    return GPDE_OK;
}

int gpde_vo_pack_weights_f32(const gpde_vo_plan *pl, const float *V, int m, int flags, void *workspace,
                             gpde_stream_t stream) {
    if (!pl || !V || !workspace) return fail(GPDE_ERR_ARG, "vo_pack_weights: null argument");
    if ((uintptr_t)workspace & 15) return fail(GPDE_ERR_ARG, "vo_pack_weights: workspace must be 16-byte aligned");
    Grid2Dev G;
    int NT, NX;
    size_t smem;
    if (!use_grid(pl) || !grid2_setup(pl, m, false, (flags & 1) ? 0 : 1, G, NT, NX, smem, 4, 4)) return 1;
    DeviceGuard guard(pl->device);
    grid2_pack<float>(G, V, m, NT, NX, (double *)workspace, pl->n_sm, (cudaStream_t)stream);
    GPDE_CUDA_OK(cudaGetLastError());
    return GPDE_OK;
}

int gpde_vo_posterior_f64(const gpde_vo_plan *pl, const double *a, int64_t a_stride, const double *V, int64_t v_stride,
                          int m, const double *rho, const double *noise_var, const double *g, const double *prec,
                          double *mean, double *vars, int *info, int64_t N, gpde_stream_t stream) {
    if (!pl || N < 0 || m <= 0) return fail(GPDE_ERR_ARG, "vo_posterior: bad argument");
    if (m > 64) return fail(GPDE_ERR_SIZE, "vo_posterior: m = %d > 64 weighting functions per data point", m);
    if (N == 0) return GPDE_OK;
    if (!a || !V || !rho || !noise_var || !g || !prec || !mean || !vars) return fail(GPDE_ERR_ARG, "vo_posterior: null argument");
    if (v_stride != 0 && v_stride != (int64_t)pl->dev.d * m) return fail(GPDE_ERR_ARG, "vo_posterior: v_stride must be 0 or d*m");
    DeviceGuard guard(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (m <= 32) {
        auto kern = vo_posterior_kernel<32>;
        const size_t smem = post_smem_bytes<32>(pl->csr.max_row);
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)N, kPostThreads, smem, st>>>(pl->csr, a, (long long)a_stride, V, (long long)v_stride, m, rho, noise_var, g,
                                                      prec, mean, vars, info);
    } else {
        auto kern = vo_posterior_kernel<64>;
        const size_t smem = post_smem_bytes<64>(pl->csr.max_row);
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)N, kPostThreads, smem, st>>>(pl->csr, a, (long long)a_stride, V, (long long)v_stride, m, rho, noise_var, g,
                                                      prec, mean, vars, info);
    }
    GPDE_CUDA_OK(cudaGetLastError());
    return GPDE_OK;
}

int gpde_vo_moments_f64(const gpde_vo_plan *pl, const double *a, int64_t a_stride, const double *V, int64_t v_stride, int m,
                        const double *rho, const double *v, double *out_r, double *out_s2, int64_t N, gpde_stream_t stream) {
    if (!pl || N < 0 || m <= 0) return fail(GPDE_ERR_ARG, "vo_moments: bad argument");
    if (m > 64) return fail(GPDE_ERR_SIZE, "vo_moments: m = %d > 64 weighting functions per data point", m);
    if (N == 0) return GPDE_OK;
    if (!a || !V || !rho || !v || !out_r || !out_s2) return fail(GPDE_ERR_ARG, "vo_moments: null argument");
    if (v_stride != 0 && v_stride != (int64_t)pl->dev.d * m) return fail(GPDE_ERR_ARG, "vo_moments: v_stride must be 0 or d*m");
    DeviceGuard guard(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (m <= 32) {
        auto kern = vo_moments_kernel<32>;
        const size_t smem = moments_smem_bytes<32>(pl->csr.max_row);
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)N, kPostThreads, smem, st>>>(pl->csr, a, (long long)a_stride, V, (long long)v_stride, m, rho, v, out_r, out_s2);
    } else {
        auto kern = vo_moments_kernel<64>;
        const size_t smem = moments_smem_bytes<64>(pl->csr.max_row);
        GPDE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)N, kPostThreads, smem, st>>>(pl->csr, a, (long long)a_stride, V, (long long)v_stride, m, rho, v, out_r, out_s2);
    }
    GPDE_CUDA_OK(cudaGetLastError());
    return GPDE_OK;
}

int gpde_vo_residual_f64(const gpde_vo_plan *pl, const double *a, int64_t a_stride, int a_is_log, const double *y,
                         const double *g, int64_t g_stride, const double *V, int m, double *r, double *rho,
                         void *workspace, int flags, int64_t B, gpde_stream_t stream) {
    return vo_residual<double>(pl, a, a_stride, a_is_log, y, g, g_stride, V, m, r, rho, workspace, flags, B, stream);
}
int gpde_vo_residual_f32(const gpde_vo_plan *pl, const float *a, int64_t a_stride, int a_is_log, const float *y,
                         const float *g, int64_t g_stride, const float *V, int m, float *r, float *rho,
                         void *workspace, int flags, int64_t B, gpde_stream_t stream) {
    return vo_residual<float>(pl, a, a_stride, a_is_log, y, g, g_stride, V, m, r, rho, workspace, flags, B, stream);
}
int gpde_vo_residual_T_f64(const gpde_vo_plan *pl, const double *a, int64_t a_stride, int a_is_log, const double *V,
                           int m, const double *s, double *q, void *workspace, int64_t B, gpde_stream_t stream) {
    return vo_residual_T<double>(pl, a, a_stride, a_is_log, V, m, s, q, workspace, B, stream);
}
int gpde_vo_residual_T_f32(const gpde_vo_plan *pl, const float *a, int64_t a_stride, int a_is_log, const float *V,
                           int m, const float *s, float *q, void *workspace, int64_t B, gpde_stream_t stream) {
    return vo_residual_T<float>(pl, a, a_stride, a_is_log, V, m, s, q, workspace, B, stream);
}

}  // extern "C"
