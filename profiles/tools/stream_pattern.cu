// Microbenchmark: HBM bandwidth of the VO access pattern -- many concurrent sample streams, each advancing by a
// small contiguous piece per step (512 B pixel row / 504 B node row), streams 32 KB apart -- against piece
// size, steps in flight and CTA count.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_pattern.bin stream_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int U>
__global__ void __launch_bounds__(512) pattern_kernel(const double2 *__restrict__ base, long long stream_stride16,
                                                      int streams_per_cta, int piece16, int steps, double *out) {
    // stream s of this CTA: base + (cta*streams_per_cta + s) * stream_stride16 ; step t reads piece16 16-byte units at offset t*piece16
    const int ops = streams_per_cta * piece16;
    double acc = 0.0;
    for (int t0 = 0; t0 < steps; t0 += U) {
        double2 v[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int o = threadIdx.x + 512 * i;
                v[u][i] = make_double2(0.0, 0.0);
                if (o < ops) {
                    const int s = o / piece16, ch = o - s * piece16;
                    v[u][i] = __ldcs(base + ((long long)blockIdx.x * streams_per_cta + s) * stream_stride16 +
                                     (long long)(t0 + u) * piece16 + ch);
                }
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc += v[u][i].x + v[u][i].y;
    }
    if (acc == 123.456) out[0] = acc;
}

int main() {
    const long long total16 = (1ll << 30) / 16 * 4;   // 4 GiB buffer
    double2 *buf;
    cudaMalloc(&buf, total16 * 16);
    cudaMemset(buf, 0, total16 * 16);
    double *out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int ctas, spc, piece, steps; long long stride; };
    // 8192 streams of 32 KB (4096 samples x {a,y}) in the VO kernel: 128 CTAs x 64 streams x 512 B x 64 steps
    Cfg cfgs[] = {
        {128, 64, 512, 64, 32768}, {128, 64, 512, 64, 32760 + 8}, {148, 64, 512, 64, 32768},
        {128, 64, 1024, 32, 32768}, {128, 64, 2048, 16, 32768}, {256, 32, 512, 64, 32768}, {512, 16, 512, 64, 32768},
        {128, 64, 512, 512, 262144}, {148, 64, 512, 512, 262144}, {296, 32, 512, 512, 262144}, {148, 64, 2048, 128, 262144},
    };
    for (auto c : cfgs) {
        for (int U = 1; U <= 8; U *= 2) {
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                const long long s16 = c.stride / 16;
                if (U == 1) pattern_kernel<1><<<c.ctas, 512>>>(buf, s16, c.spc, c.piece / 16, c.steps, out);
                if (U == 2) pattern_kernel<2><<<c.ctas, 512>>>(buf, s16, c.spc, c.piece / 16, c.steps, out);
                if (U == 4) pattern_kernel<4><<<c.ctas, 512>>>(buf, s16, c.spc, c.piece / 16, c.steps, out);
                if (U == 8) pattern_kernel<8><<<c.ctas, 512>>>(buf, s16, c.spc, c.piece / 16, c.steps, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double bytes = (double)c.ctas * c.spc * c.piece * c.steps;
            printf("ctas %3d streams/cta %2d piece %4d B steps %3d stride %6lld  in-flight steps %d : %8.1f GB/s  (%.1f MB, %.1f us)\n",
                   c.ctas, c.spc, c.piece, c.steps, c.stride, U, bytes / ms * 1e-6, bytes * 1e-6, ms * 1e3);
        }
    }
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
