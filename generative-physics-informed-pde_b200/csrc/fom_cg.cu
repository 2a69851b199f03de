// Batched preconditioned conjugate gradients for the fine-mesh label solves (sm_100a): the vector half of one iteration.
//
//   K_ff(a_b) y_b = f_eff,b   for every sample b     (reference: physics/LinearElliptic.py:85-101 df.solve, :120-133
//                                                      spsolve, one sample at a time in utils/data.py:96-99)
//
// The matrix half of an iteration, A p = K_ff(a) p, is ONE launch of the matrix-free residual kernels over the whole batch
// (gpde_vo_residual with y = p, no Dirichlet data, no load: the rho output).  This kernel does everything else of the
// iteration for sample b in one CTA: the two inner products, the updates of x, r, p with the Jacobi preconditioner, and the
// convergence bookkeeping -- a sample whose residual has reached its tolerance is frozen (alpha = beta = 0), so a batch runs
// a fixed number of iterations between two host-side checks without 0/0 in the samples that finished early.
//   pAp = p.Ap;  alpha = rz / pAp;  x += alpha p;  r -= alpha Ap;  z = dinv r;  rz' = r.z;  beta = rz' / rz;  p = z + beta p
// Reductions: per-thread strided partial sums, shuffle tree, then the warps' partials in warp order -- the same order for a
// sample wherever it sits in the batch.
#include "common.cuh"

namespace gpde {

constexpr int kCgThreads = 256;

__device__ __forceinline__ double cg_block_sum(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                       // red may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kCgThreads / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(kCgThreads)
cg_step_kernel(const double *__restrict__ Ap, const double *__restrict__ dinv, long long dinv_stride, double *__restrict__ x,
               double *__restrict__ r, double *__restrict__ p, double *__restrict__ rz, const double *__restrict__ stop2,
               double *__restrict__ rnorm2, int d) {
    __shared__ double red[kCgThreads / 32];
    const long long b = blockIdx.x;
    const double *Apb = Ap + b * d, *dv = dinv + b * dinv_stride;
    double *xb = x + b * d, *rb = r + b * d, *pb = p + b * d;
    double s = 0.0;
    for (int i = threadIdx.x; i < d; i += kCgThreads) s = fma(pb[i], Apb[i], s);
    const double pAp = cg_block_sum(s, red);
    const double rz_old = rz[b];
    const bool active = rnorm2[b] > stop2[b] && pAp > 0.0 && rz_old > 0.0;
    const double alpha = active ? rz_old / pAp : 0.0;
    double s_rz = 0.0, s_rr = 0.0;
    for (int i = threadIdx.x; i < d; i += kCgThreads) {
        const double pi = pb[i];
        xb[i] = fma(alpha, pi, xb[i]);
        const double ri = fma(-alpha, Apb[i], rb[i]);
        rb[i] = ri;
        s_rz = fma(ri * dv[i], ri, s_rz);
        s_rr = fma(ri, ri, s_rr);
    }
    const double rz_new = cg_block_sum(s_rz, red);
    const double rr = cg_block_sum(s_rr, red);
    const double beta = active ? rz_new / rz_old : 0.0;
    if (active)
        for (int i = threadIdx.x; i < d; i += kCgThreads) pb[i] = fma(beta, pb[i], rb[i] * dv[i]);
    if (threadIdx.x == 0 && active) {
        rz[b] = rz_new;
        rnorm2[b] = rr;
    }
}

// r = b - (A x0 given as Ax), z = dinv r, p = z, rz = r.z, rnorm2 = r.r, stop2 = tol2 * max(b.b, tiny)
__global__ void __launch_bounds__(kCgThreads)
cg_init_kernel(const double *__restrict__ rhs, const double *__restrict__ Ax, const double *__restrict__ dinv, long long dinv_stride,
               double *__restrict__ r, double *__restrict__ p, double *__restrict__ rz, double *__restrict__ stop2,
               double *__restrict__ rnorm2, double tol2, int d) {
    __shared__ double red[kCgThreads / 32];
    const long long b = blockIdx.x;
    const double *dv = dinv + b * dinv_stride;
    double s_rz = 0.0, s_rr = 0.0, s_bb = 0.0;
    for (int i = threadIdx.x; i < d; i += kCgThreads) {
        const double bi = rhs[b * d + i];
        const double ri = Ax ? bi - Ax[b * d + i] : bi;
        r[b * d + i] = ri;
        const double zi = ri * dv[i];
        p[b * d + i] = zi;
        s_rz = fma(ri, zi, s_rz);
        s_rr = fma(ri, ri, s_rr);
        s_bb = fma(bi, bi, s_bb);
    }
    const double a0 = cg_block_sum(s_rz, red), a1 = cg_block_sum(s_rr, red), a2 = cg_block_sum(s_bb, red);
    if (threadIdx.x == 0) {
        rz[b] = a0;
        rnorm2[b] = a1;
        stop2[b] = tol2 * fmax(a2, 1e-300);
    }
}

}  // namespace gpde

using namespace gpde;

extern "C" {

int gpde_cg_init_f64(const double *rhs, const double *Ax, const double *dinv, int64_t dinv_stride, double *r, double *p,
                     double *rz, double *stop2, double *rnorm2, double tol, int d, int64_t B, int device,
                     gpde_stream_t stream) {
    if (!rhs || !dinv || !r || !p || !rz || !stop2 || !rnorm2 || d <= 0 || B < 0 || !(tol >= 0.0))
        return fail(GPDE_ERR_ARG, "cg_init: bad argument");
    if (B == 0) return GPDE_OK;
    DeviceGuard guard(device);
    cg_init_kernel<<<(unsigned)B, kCgThreads, 0, (cudaStream_t)stream>>>(rhs, Ax, dinv, (long long)dinv_stride, r, p, rz, stop2,
                                                                           rnorm2, tol * tol, d);
    GPDE_CUDA_OK(cudaGetLastError());
    return GPDE_OK;
}

int gpde_cg_step_f64(const double *Ap, const double *dinv, int64_t dinv_stride, double *x, double *r, double *p, double *rz,
                     const double *stop2, double *rnorm2, int d, int64_t B, int device, gpde_stream_t stream) {
    if (!Ap || !dinv || !x || !r || !p || !rz || !stop2 || !rnorm2 || d <= 0 || B < 0)
        return fail(GPDE_ERR_ARG, "cg_step: bad argument");
    if (B == 0) return GPDE_OK;
    DeviceGuard guard(device);
    cg_step_kernel<<<(unsigned)B, kCgThreads, 0, (cudaStream_t)stream>>>(Ap, dinv, (long long)dinv_stride, x, r, p, rz, stop2,
                                                                           rnorm2, d);
    GPDE_CUDA_OK(cudaGetLastError());
    return GPDE_OK;
}

}  // extern "C"
