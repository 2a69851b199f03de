/*
 * gpde_b200.h -- C ABI of the B200-native physics layer (libgpde_b200.so).
 *
 * Drop-in boundary for the ONE hot path of pkmtum/generative-physics-informed-pde:
 *   - the coarse-grained model (assemble K(x)=sum_e x_e K_e, solve, adjoint):
 *       reference bottleneck/ROM.py:59-100  (+ autograd of it)
 *       reference bottleneck/components.py:296-311 (exp(X)+1e-8, prolongation y = W u)
 *   - the virtual-observable residuals  r = V^T (K_fom(a) u~ - f)_free  and the transposed
 *     application  q = K_ff(a) (V s):
 *       reference bottleneck/VirtualObservables.py:57-69, 642-669, 971-998
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types cross this boundary;
 *   - "host" pointers are read during the call and never retained;
 *   - "device" pointers must live on the plan's device; the library never allocates or frees
 *     user data -- scratch comes from the caller (query *_workspace_bytes first);
 *   - every launch goes to the caller's stream (a cudaStream_t passed as void*); no entry point
 *     synchronises the device except plan_create (uploads constants) and plan_destroy;
 *   - plans are immutable after creation => re-entrant across streams; one plan per device;
 *   - return value: 0 = ok, <0 = error (gpde_last_error() gives the text for this thread);
 *     numerical failures are reported LAPACK-style through a device info word, never by a trap;
 *   - matrices are row-major; batch is the leading axis (X[B,E], F[B,n], u[B,n], y[B,d] ...),
 *     exactly the torch layouts the reference passes around.
 *   - *_f64: double I/O.  *_f32: float I/O, double arithmetic inside (meets the 1e-5 tier).
 */
#ifndef GPDE_B200_H
#define GPDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPDE_OK 0
#define GPDE_ERR_ARG (-1)      /* bad argument */
#define GPDE_ERR_CUDA (-2)     /* CUDA runtime error (allocation, launch) */
#define GPDE_ERR_SIZE (-3)     /* problem too large for this build's kernels */

/* info word bits written by the ROM kernels (atomicOr into *info, device int32) */
#define GPDE_INFO_NONPOSITIVE_X 1      /* some conductivity <= 1e-12   (ROM.py:74-76 -> ValueError) */
#define GPDE_INFO_NOT_SPD 2            /* non-positive pivot in the factorisation              */

typedef struct gpde_rom_plan gpde_rom_plan;
typedef struct gpde_vo_plan gpde_vo_plan;
typedef void *gpde_stream_t; /* cudaStream_t */

int gpde_version(void);
const char *gpde_last_error(void);

/* ------------------------------------------------------------------ coarse-grained model */

/* Replaces ROM.__init__ (bottleneck/ROM.py:8-15): keeps the element tensor and Dirichlet map on
 * the device, in the sparse/banded form the kernels use.
 *   M_host   [n,n,E] row-major, M[:,:,e] = unit-conductivity stiffness of coarse cell e (ROM.py:46-53)
 *   bc_dofs  [n_bc]  constrained dofs (physics.constrained_dofs, ROM.py:14)
 * Free dofs are the remaining ones, ascending (optionally reordered to shrink the band). */
int gpde_rom_plan_create(gpde_rom_plan **plan, int n, int E, const double *M_host,
                         const int64_t *bc_dofs, int n_bc, int device);
int gpde_rom_plan_destroy(gpde_rom_plan *plan);

/* out[0]=n, [1]=E, [2]=n_free, [3]=half bandwidth, [4]=factor doubles per sample,
 * [5]=assembly contributions, [6]=lanes per sample (1 = thread-per-sample kernels without a stash, 2 = windowed
 * thread-per-sample kernels: pass the stash to the forward call too, 8/16/32 = cooperative kernels), [7]=device */
int gpde_rom_plan_info(const gpde_rom_plan *plan, int64_t out[8]);

/* bytes of factor stash needed for a batch of B samples (always doubles; the layout is the kernels' business: the
 * windowed kernels interleave blocks of 128 samples, so size the buffer with this call, not with out[4] * B) */
size_t gpde_rom_factor_bytes(const gpde_rom_plan *plan, int64_t B);

/* Replaces ROM.__call__ (bottleneck/ROM.py:65-88)  [x_is_log = 0: X are conductivities]
 * and ReducedOrderModelOperator's exp(X)+1e-8 -> rom (components.py:298) [x_is_log = 1].
 *   X [B,E], F [B,n] (Dirichlet values already written at bc_dofs, BoundaryConditions.py:132-147)
 *   u [B,n]  solution incl. Dirichlet dofs
 *   factor   [gpde_rom_factor_bytes] banded LDL^T factor kept for the adjoint (may be NULL; plans with lanes == 2 then
 *            fall back to the slower cooperative kernels)
 *   info     device int32, OR-ed with GPDE_INFO_* (may be NULL)                              */
int gpde_rom_forward_f64(const gpde_rom_plan *plan, const double *X, int x_is_log, const double *F,
                         double *u, double *factor, int *info, int64_t B, gpde_stream_t stream);
int gpde_rom_forward_f32(const gpde_rom_plan *plan, const float *X, int x_is_log, const float *F,
                         float *u, double *factor, int *info, int64_t B, gpde_stream_t stream);

/* Replaces autograd through matmul/index_put/solve: with lambda = A^-T gbar_u,
 *   gradX[b,e] = - sum_{i free} sum_j lambda_i K_e[i,j] u_j   (times exp(X) when x_is_log)
 *   gradF[b,:] = lambda                                            (may be NULL)
 * factor == NULL => the factorisation is recomputed from X.                                  */
int gpde_rom_adjoint_f64(const gpde_rom_plan *plan, const double *X, int x_is_log, const double *u,
                         const double *factor, const double *gbar_u, double *gradX, double *gradF,
                         int64_t B, gpde_stream_t stream);
int gpde_rom_adjoint_f32(const gpde_rom_plan *plan, const float *X, int x_is_log, const float *u,
                         const double *factor, const float *gbar_u, float *gradX, float *gradF,
                         int64_t B, gpde_stream_t stream);

/* GetStiffness (bottleneck/ROM.py:91-100): K[n,n,B] (batch LAST, as the reference returns it),
 * Dirichlet rows replaced by identity rows when dirichlet != 0.  X are conductivities.       */
int gpde_rom_stiffness_f64(const gpde_rom_plan *plan, const double *X, double *K, int dirichlet,
                           int64_t B, gpde_stream_t stream);

/* Prolongation y = W u and its transpose gbar_u = W^T gbar_y (components.py:298 einsum 'sk,nk->ns')
 * with W[d,n] held as CSR on the device (rows have <= 3 non-zeros for P1).                    */
typedef struct gpde_prolong_plan gpde_prolong_plan;
int gpde_prolong_plan_create(gpde_prolong_plan **plan, int d, int n, const double *W_host, int device);
int gpde_prolong_plan_destroy(gpde_prolong_plan *plan);
int gpde_prolong_apply_f64(const gpde_prolong_plan *plan, const double *u, double *y, int64_t B,
                           gpde_stream_t stream);                 /* y[B,d]  = u[B,n] W^T */
int gpde_prolong_apply_T_f64(const gpde_prolong_plan *plan, const double *gy, double *gu, int64_t B,
                             gpde_stream_t stream);               /* gu[B,n] = gy[B,d] W   */
int gpde_prolong_apply_f32(const gpde_prolong_plan *plan, const float *u, float *y, int64_t B,
                           gpde_stream_t stream);
int gpde_prolong_apply_T_f32(const gpde_prolong_plan *plan, const float *gy, float *gu, int64_t B,
                             gpde_stream_t stream);

/* Fused operator epilogue: the diagonal-Gaussian log-likelihood of the operator output
 * DiagonalGaussianLogLikelihood(Y, W u, 2 ls) (bottleneck/utils.py:231-241 on components.py:296-298, as evaluated at
 * generative.py:438-439) and its gradients, without writing mu_y = W u [B,d]:
 *     L  [B]   L_b   = -1/2 sum_i [2 ls_i + ((Y_bi - (W u_b)_i) / exp(ls_i))^2 + log 2 pi]        (always double)
 *     gu [B,n] dL_b/du_b = W^T ((Y_b - W u_b) exp(-2 ls))                          (may be NULL)
 *     gls[d]   d(sum_b L_b)/dls_i = sum_b (e_bi^2 - 1)                             (always double; may be NULL)
 * u [B,n], Y [B,d], ls [d] (log standard deviations, the operator's logsigmas_y).  L and gls are zeroed by the call. */
int gpde_prolong_loglik_f64(const gpde_prolong_plan *plan, const double *u, const double *Y, const double *ls,
                            double *L, double *gu, double *gls, int64_t B, gpde_stream_t stream);
int gpde_prolong_loglik_f32(const gpde_prolong_plan *plan, const float *u, const float *Y, const float *ls,
                            double *L, float *gu, double *gls, int64_t B, gpde_stream_t stream);

/* Monte-Carlo predictive moments of the operator output per data point (generative.py:198-207: propagate_samples, then
 * torch.mean / torch.std over the S samples) without the [N S, d] samples: with ubar / Cov_u the sample mean / unbiased
 * covariance of the S coarse solutions u[n,s,:],
 *     y_mean[n,i] = W_i . ubar_n,     y_std[n,i] = sqrt(W_i Cov_u,n W_i^T + exp(2 ls_i))
 * (the reference's estimator with the output noise integrated out).  u [N,S,n_coarse], ls [d], y_mean / y_std [N,d]. */
int gpde_prolong_moments_f64(const gpde_prolong_plan *plan, const double *u, const double *ls, double *y_mean,
                             double *y_std, int64_t N, int S, gpde_stream_t stream);
int gpde_prolong_moments_f32(const gpde_prolong_plan *plan, const float *u, const float *ls, float *y_mean,
                             float *y_std, int64_t N, int S, gpde_stream_t stream);

/* ------------------------------------------------------------------ virtual observables */

/* Replaces QuerryPoint._assemble_system / LinearEllipticPhysics.assemble_system
 * (VirtualObservables.py:57-59, physics/LinearElliptic.py:137-159): the fine operator is kept
 * matrix-free as element data; K_fom(a) = sum_c a[cell_to_input[c]] * Ke[c].
 *   cell_dofs     [n_cells,3]   P1 connectivity
 *   Ke            [n_cells,3,3] unit-conductivity element stiffness
 *   cell_to_input [n_cells]     index of the per-sample conductivity entry used by cell c
 *                               (identity for DG0 input; pixel id for image input)
 *   free_dofs [d], bc_dofs [n_bc]; f_full [n_nodes] load vector or NULL (zero)               */
int gpde_vo_plan_create(gpde_vo_plan **plan, int n_nodes, int n_cells, const int32_t *cell_dofs,
                        const double *Ke, const int32_t *cell_to_input, int n_inputs,
                        const int64_t *free_dofs, int d, const int64_t *bc_dofs, int n_bc,
                        const double *f_full, int device);
int gpde_vo_plan_destroy(gpde_vo_plan *plan);
/* out[0]=n_nodes, [1]=n_cells, [2]=n_inputs, [3]=d, [4]=n_bc, [5]=slots per row, [6]=device */
int gpde_vo_plan_info(const gpde_vo_plan *plan, int64_t out[8]);

/* Which kernels serve gpde_vo_residual_* for m weighting functions and elem_bytes (8 = f64, 4 = f32) I/O:
 * 3 = structured pixel grid, m > 32: V packing + ONE kernel that produces the fine residual inside the FP64 tensor-core
 *     contraction (vo_gridgemm.cuh) + the reduction of its partial tiles (three launches),
 * 2 = structured-grid kernel (V packing launch + one fused launch), 1 = generic fused kernel (one launch),
 * 0 = version-1 kernels (generic matvec + V padding + tensor-core contraction; three launches).  Informational (launch counting, tests); calls whose
 * pointers are not 16-byte aligned, that pass y = NULL or that ask for rho fall back from 2 to 1. */
int gpde_vo_plan_kernel_path(const gpde_vo_plan *plan, int m, int elem_bytes);

/* scratch bytes for residual / residual_T on B samples with m weighting functions */
size_t gpde_vo_workspace_bytes(const gpde_vo_plan *plan, int64_t B, int m);

/* r[B,m] = V^T (K_fom(a_b) u~_b - f)_free ,  u~ = y on free dofs, g on constrained dofs
 *        = Gamma_b y_b - alpha_b   (VirtualObservables.py:61-69, 662, 990).
 *   a   [B or 1, n_inputs]  (a_stride = n_inputs, or 0 to share one field across the batch)
 *   a_is_log != 0: a holds log-conductivities (QuerryPoint.x), exp() applied inside
 *   y   [B,d]   (NULL = zeros);  g [B or 1, n_bc] with g_stride = n_bc or 0 (NULL = zeros)
 *   V   [d,m] row-major weighting matrix (NULL with m=0: only rho is produced)
 *   rho [B,d] optional output of the fine residual itself (may be NULL)
 *   flags: bit0 = ignore the load vector f
 *          bit1 = `workspace` already holds V packed by gpde_vo_pack_weights_f64 (same plan, m, bit0):
 *                 the call skips its packing launch; it fails with GPDE_ERR_ARG instead of falling
 *                 back when the lean structured-grid kernel cannot serve it (V must still be passed)
 *          bits 8-15 = number of SMs to leave to kernels the caller runs beside this call on other streams
 *                 (the structured-grid kernel sizes its last wave for the remaining SMs; 0 = all SMs)  */
int gpde_vo_residual_f64(const gpde_vo_plan *plan, const double *a, int64_t a_stride, int a_is_log,
                         const double *y, const double *g, int64_t g_stride, const double *V, int m,
                         double *r, double *rho, void *workspace, int flags, int64_t B,
                         gpde_stream_t stream);
int gpde_vo_residual_f32(const gpde_vo_plan *plan, const float *a, int64_t a_stride, int a_is_log,
                         const float *y, const float *g, int64_t g_stride, const float *V, int m,
                         float *r, float *rho, void *workspace, int flags, int64_t B,
                         gpde_stream_t stream);

/* Packs V[d,m] once into `workspace` (>= gpde_vo_workspace_bytes(plan, B, m), 16-byte aligned) in the
 * fragment order of the structured-grid residual kernel, for callers whose weighting functions stay fixed
 * over many residual calls (the CoarseGrainedResidual sampler's V = W, VirtualObservables.py:297-321, changes
 * only at resample()).  Returns 0 when packed (pass flags bit1 to gpde_vo_residual_f64 with that workspace),
 * 1 when this plan / m has no packed layout (call residual without bit1), < 0 on error. */
int gpde_vo_pack_weights_f64(const gpde_vo_plan *plan, const double *V, int m, int flags, void *workspace,
                             gpde_stream_t stream);
/* the same for FP32 I/O calls (gpde_vo_residual_f32 with flags bit1); the packed copy holds doubles either way */
int gpde_vo_pack_weights_f32(const gpde_vo_plan *plan, const float *V, int m, int flags, void *workspace,
                             gpde_stream_t stream);

/* Batched Gaussian conditioning of all data points of a virtual-observable ensemble in ONE launch, matrix-free
 * (VirtualObservable.update, bottleneck/VirtualObservables.py:642-669, looped over the data points at :891-898):
 *     Lambda_n = Gamma_n C_n Gamma_n^T + diag(noise_var),  C_n = diag(1 / prec_n),  Gamma_n^T = K_ff(a_n) V_n
 *     mean_n   = g_n - C_n Gamma_n^T Lambda_n^-1 (Gamma_n g_n - alpha_n),   Gamma_n g_n - alpha_n = V_n^T rho_n
 *     vars_n   = diag(C_n) - diag(C_n Gamma_n^T Lambda_n^-1 Gamma_n C_n)
 * a [N, n_inputs] conductivities (NOT logs; a_stride = 0: one field shared by all data points); V [d,m] shared
 * (v_stride = 0) or [N,d,m] (v_stride = d*m); m <= 64; rho [N,d] = the fine residual K_fom(a_n) g~_n - f of the prior
 * mean (the rho output of gpde_vo_residual with y = g); noise_var [m]; g, prec, mean, vars [N,d].
 * info (device int32, may be NULL) receives GPDE_INFO_NOT_SPD if some Lambda_n has a non-positive pivot.           */
int gpde_vo_posterior_f64(const gpde_vo_plan *plan, const double *a, int64_t a_stride, const double *V,
                          int64_t v_stride, int m, const double *rho, const double *noise_var, const double *g,
                          const double *prec, double *mean, double *vars, int *info, int64_t N, gpde_stream_t stream);

/* The two per-data-point terms of the precision hyper-update (VirtualObservablesEnsemble.update_vo_precision,
 * bottleneck/VirtualObservables.py:985-990), all data points at once:
 *     out_r [n,j] = (V_n^T rho_n)_j          = (Gamma_n mean_n - alpha_n)_j  for rho_n = rho(mean_n)
 *     out_s2[n,j] = sum_i Gamma_n[j,i]^2 v[n,i]                                                                  */
int gpde_vo_moments_f64(const gpde_vo_plan *plan, const double *a, int64_t a_stride, const double *V, int64_t v_stride,
                        int m, const double *rho, const double *v, double *out_r, double *out_s2, int64_t N,
                        gpde_stream_t stream);

/* q[B,d] = K_ff(a_b) (V s_b) = Gamma_b^T s_b  (VirtualObservables.py:663; with s = P r it is the
 * gradient of 1/2 r^T P r w.r.t. y).  With B = m, s = I and a_stride = 0 it yields Gamma itself.
 * On the reference's pixel meshes with m <= 32 this is one kernel (after a small packing launch of V^T): the rows
 * of V s are produced inside the marching kernel and never reach global memory; `workspace` holds the packed V^T
 * (or, on the other routes, V s [B,d]): >= gpde_vo_workspace_bytes(plan, B, m), 16-byte aligned. */
int gpde_vo_residual_T_f64(const gpde_vo_plan *plan, const double *a, int64_t a_stride, int a_is_log,
                           const double *V, int m, const double *s, double *q, void *workspace,
                           int64_t B, gpde_stream_t stream);
int gpde_vo_residual_T_f32(const gpde_vo_plan *plan, const float *a, int64_t a_stride, int a_is_log,
                           const float *V, int m, const float *s, float *q, void *workspace,
                           int64_t B, gpde_stream_t stream);

/* ---- fine-mesh label solves: batched preconditioned CG (setup-time replacement of the per-sample FEniCS / spsolve calls
 * of physics/LinearElliptic.py:85-101, 120-133 in utils/data.py:96-99).  One iteration = one gpde_vo_residual_f64 call with
 * y = p (rho output, flags bit0, no Dirichlet data: rho = K_ff(a) p for the whole batch) + one gpde_cg_step_f64 call.
 * All arrays device, float64: rhs, Ax, x, r, p [B,d]; dinv [B or 1, d] (dinv_stride = d or 0) the inverse diagonal of
 * K_ff(a_b); rz, stop2, rnorm2 [B] per-sample scalars owned by the caller.
 *   init: r = rhs - Ax (Ax = NULL: x0 = 0), p = dinv r, rz = r.p, rnorm2 = r.r, stop2 = tol^2 rhs.rhs
 *   step: alpha = rz / p.Ap; x += alpha p; r -= alpha Ap; z = dinv r; beta = r.z / rz; p = z + beta p; rz, rnorm2 updated;
 *         samples with rnorm2 <= stop2 are frozen. */
int gpde_cg_init_f64(const double *rhs, const double *Ax, const double *dinv, int64_t dinv_stride, double *r,
                     double *p, double *rz, double *stop2, double *rnorm2, double tol, int d, int64_t B,
                     int device, gpde_stream_t stream);
int gpde_cg_step_f64(const double *Ap, const double *dinv, int64_t dinv_stride, double *x, double *r, double *p,
                     double *rz, const double *stop2, double *rnorm2, int d, int64_t B, int device,
                     gpde_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GPDE_B200_H */
