// Prolongation y = W u and its transpose, W[d,n] = coarse P1 basis at the fine free nodes
// (bottleneck/components.py:298 einsum('sk,nk->ns', W, u); factories/model.py:140).
// The reference keeps W dense (d x n) and calls a GEMM; its rows have <= 3 non-zeros for P1, so it is
// held here as CSR (forward) + CSC (transpose) and applied as gathers -- the [B,d] output is written
// once, coalesced, and W itself stays in L1/L2.
#include "common.cuh"

#include <algorithm>

namespace gpde {
struct ProlongDev {
    int d, n;
    const int *row_ptr, *col;   // CSR
    const double *val;
    const int *col_ptr, *row;   // CSC
    const double *cval;
};
}  // namespace gpde

struct gpde_prolong_plan {
    gpde::ProlongDev dev;
    int device;
    std::vector<void *> allocs;
};

namespace gpde {

__device__ __forceinline__ double pld(const double *p) { return *p; }
__device__ __forceinline__ double pld(const float *p) { return (double)*p; }

template <typename T>
__global__ void prolong_apply_kernel(ProlongDev P, const T *__restrict__ u, T *__restrict__ y, long long B) {
    const long long total = B * (long long)P.d;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / P.d;
        const int i = (int)(idx - b * P.d);
        double acc = 0.0;
        const int t1 = P.row_ptr[i + 1];
        for (int t = P.row_ptr[i]; t < t1; ++t) acc = fma(P.val[t], pld(u + b * P.n + P.col[t]), acc);
        y[idx] = (T)acc;
    }
}

// one warp per (sample, coarse dof): reduce over the fine nodes in the hat function's support
template <typename T>
__global__ void prolong_apply_T_kernel(ProlongDev P, const T *__restrict__ gy, T *__restrict__ gu, long long B) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = B * (long long)P.n;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
        const long long b = w / P.n;
        const int k = (int)(w - b * P.n);
        double acc = 0.0;
        const int t1 = P.col_ptr[k + 1];
        for (int t = P.col_ptr[k] + lane; t < t1; t += 32) acc = fma(P.cval[t], pld(gy + b * P.d + P.row[t]), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) gu[w] = (T)acc;
    }
}

template <typename T>
static cudaError_t track_p(gpde_prolong_plan *pl, const T **dst, const std::vector<T> &src) {
    T *p = nullptr;
    cudaError_t e = upload(&p, src);
    if (p) pl->allocs.push_back((void *)p);
    *dst = p;
    return e;
}

template <typename T>
static int prolong_apply(const gpde_prolong_plan *pl, const T *u, T *y, int64_t B, gpde_stream_t stream, bool transpose) {
    if (!pl || B < 0) return fail(GPDE_ERR_ARG, "prolong_apply: bad argument");
    if (B == 0) return GPDE_OK;
    if (!u || !y) return fail(GPDE_ERR_ARG, "prolong_apply: null argument");
    DeviceGuard guard(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int nsm = sm_count(pl->device);
    if (!transpose) {
        const long long total = B * (long long)pl->dev.d;
        const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)nsm * 16);
        prolong_apply_kernel<T><<<grid, 256, 0, st>>>(pl->dev, u, y, B);
    } else {
        const long long total = B * (long long)pl->dev.n;   // warps
        const unsigned grid = (unsigned)std::min<long long>((total + 7) / 8, (long long)nsm * 16);
        prolong_apply_T_kernel<T><<<grid, 256, 0, st>>>(pl->dev, u, y, B);
    }
    GPDE_CUDA_OK(cudaGetLastError());
    return GPDE_OK;
}

}  // namespace gpde

using namespace gpde;

extern "C" {

int gpde_prolong_plan_create(gpde_prolong_plan **plan, int d, int n, const double *W, int device) {
    if (!plan || !W || d <= 0 || n <= 0) return fail(GPDE_ERR_ARG, "prolong_plan_create: bad argument");
    std::vector<int> row_ptr(d + 1, 0), col, col_ptr(n + 1, 0), row;
    std::vector<double> val, cval;
    for (int i = 0; i < d; ++i) {
        for (int k = 0; k < n; ++k) {
            const double v = W[(size_t)i * n + k];
            if (v != 0.0) {
                col.push_back(k);
                val.push_back(v);
            }
        }
        row_ptr[i + 1] = (int)col.size();
    }
    for (int k = 0; k < n; ++k) {
        for (int i = 0; i < d; ++i) {
            const double v = W[(size_t)i * n + k];
            if (v != 0.0) {
                row.push_back(i);
                cval.push_back(v);
            }
        }
        col_ptr[k + 1] = (int)row.size();
    }
    gpde_prolong_plan *pl = new gpde_prolong_plan();
    pl->device = device;
    DeviceGuard guard(device);
    ProlongDev &D = pl->dev;
    D.d = d; D.n = n;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = track_p(pl, &D.row_ptr, row_ptr);
    if (e == cudaSuccess) e = track_p(pl, &D.col, col);
    if (e == cudaSuccess) e = track_p(pl, &D.val, val);
    if (e == cudaSuccess) e = track_p(pl, &D.col_ptr, col_ptr);
    if (e == cudaSuccess) e = track_p(pl, &D.row, row);
    if (e == cudaSuccess) e = track_p(pl, &D.cval, cval);
    if (e != cudaSuccess) {
        gpde_prolong_plan_destroy(pl);
        return fail(GPDE_ERR_CUDA, "prolong_plan_create: upload failed: %s", cudaGetErrorString(e));
    }
    *plan = pl;
    return GPDE_OK;
}

int gpde_prolong_plan_destroy(gpde_prolong_plan *pl) {
    if (!pl) return GPDE_OK;
    DeviceGuard guard(pl->device);
    for (void *p : pl->allocs) cudaFree(p);
    delete pl;
    return GPDE_OK;
}

int gpde_prolong_apply_f64(const gpde_prolong_plan *pl, const double *u, double *y, int64_t B, gpde_stream_t s) {
    return prolong_apply<double>(pl, u, y, B, s, false);
}
int gpde_prolong_apply_T_f64(const gpde_prolong_plan *pl, const double *gy, double *gu, int64_t B, gpde_stream_t s) {
    return prolong_apply<double>(pl, gy, gu, B, s, true);
}
int gpde_prolong_apply_f32(const gpde_prolong_plan *pl, const float *u, float *y, int64_t B, gpde_stream_t s) {
    return prolong_apply<float>(pl, u, y, B, s, false);
}
int gpde_prolong_apply_T_f32(const gpde_prolong_plan *pl, const float *gy, float *gu, int64_t B, gpde_stream_t s) {
    return prolong_apply<float>(pl, gy, gu, B, s, true);
}

}  // extern "C"
