// Measures the FP64 pipe denominators used in DESIGN.md: DFMA and DMMA (mma.sync.m8n8k4.f64) peak rates,
// plus a plain copy (HBM) for reference.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak.bin fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dmma_kernel(double *out, int iters) {
    double c[8][2];
    for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.0;
    const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<<<sms * 2, warps * 16>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * (double)sms * 2 * warps * 16;
        printf("DFMA  warps/SM=%2d  %.2f TFLOP/s\n", warps, fl / ms * 1e-9);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dmma_kernel<<<sms * 2, warps * 16>>>(out, iters / 4);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 256 * 8 * (iters / 4) * (double)sms * 2 * warps / 2;
        printf("DMMA  warps/SM=%2d  %.2f TFLOP/s\n", warps, fl / ms * 1e-9);
    }
    size_t n = (size_t)1 << 27;   // 2 GiB in, 2 GiB out
    double2 *a, *b;
    cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16);
    cudaMemset(a, 0, n * 16);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        copy_kernel<<<sms * 16, 512>>>(a, b, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("copy  %.1f GB/s (read+write)\n", 2.0 * n * 16 / ms * 1e-6);
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
