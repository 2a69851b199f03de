// w[B,d] = s[B,m] V[d,m]^T for m <= 32 on the FP64 tensor pipe (sm_100a).  Included by vo.cu.
//
// First half of the transposed application q = K_ff(a) (V s) = Gamma^T s (VirtualObservables.py:663) on pixel grids;
// the marching kernel (vo_grid2.cuh, rho variant) then applies K_ff to w.  The contraction length is only m, so the
// general GEMM of vo_gemm.cuh (K-tiled cp.async ring, padded and transposed operand copies: three launches, 92 us at
// cfg 2) is replaced by one launch without any operand preparation:
//   * a warp owns 32 samples (4 m-tiles of mma.sync.m8n8k4.f64) and a chunk of 8-column n-tiles; its A fragments --
//     s[sample][4 ks + lane % 4], KS = ceil(m / 4) k-steps -- stay in registers for the whole chunk;
//   * the CTA's 8 warps work on the same chunk (different samples): its rows of V are copied once into shared memory
//     (row pitch 36 doubles: conflict-free) and the B fragments V[8 nt + lane / 4][4 ks + lane % 4] come from there;
//   * D fragments (2 adjacent columns of one sample per lane) go straight to w as 16-byte pieces; w's row pitch is padded
//     to a multiple of 4 doubles so that a warp's store is 8 rows x 64 aligned bytes (whole sectors).
// (First version, measured: B fragments by global loads and 8-byte stores into rows of odd pitch -- 75 us, bound by
// the L1 request rate: 164 line requests per warp and n-tile.)
// Work: (B / 8)(d / 8) KS DMMAs = 25 us of tensor-pipe time at cfg 2; w is written once (134 MB at cfg 2).
#pragma once

namespace gpde {

constexpr int kExpandWarps = 8, kExpandMT = 4;   // warps per CTA; m-tiles (of 8 samples) per warp
constexpr int kExpandVPitch = 36;                // doubles per V row in shared memory (== 4 mod 16: conflict-free fragments)

// ldw = row pitch of w in doubles (>= d); even ldw with a 16-byte aligned w lets every lane store its two adjacent
// columns as one 16-byte piece (a warp's store = 8 rows x 64 contiguous bytes = whole sectors)
// TS: element type of s and V (FP32 I/O converts on load; w is FP64 workspace in both cases)
template <int KS, typename TS = double>
__global__ void __launch_bounds__(kExpandWarps * 32, 2)
vo_expand_dmma_kernel(const TS *__restrict__ s, const TS *__restrict__ V, double *__restrict__ w, long long ldw,
                      long long B, int d, int m, int tiles_per_chunk) {
    extern __shared__ __align__(16) double vsm[];   // [rows of the chunk][kExpandVPitch]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r8 = lane >> 2, k4 = lane & 3;
    const long long b0 = ((long long)blockIdx.x * kExpandWarps + warp) * (8 * kExpandMT);
    const int n_tiles = (d + 7) >> 3;
    const int nt0 = blockIdx.y * tiles_per_chunk, nt1 = min(n_tiles, nt0 + tiles_per_chunk);
    // the chunk's rows of V (zero past the matrix / past column m), shared by the 8 warps
    {
        const int rows = 8 * (nt1 - nt0), row0 = 8 * nt0;
        for (int idx = threadIdx.x; idx < rows * (4 * KS); idx += kExpandWarps * 32) {
            const int rr = idx / (4 * KS), c = idx - rr * (4 * KS);
            vsm[rr * kExpandVPitch + c] = (row0 + rr < d && c < m) ? (double)__ldg(V + (long long)(row0 + rr) * m + c) : 0.0;
        }
    }
    double af[kExpandMT][KS];
#pragma unroll
    for (int mt = 0; mt < kExpandMT; ++mt) {
        const long long b = b0 + 8 * mt + r8;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int c = 4 * ks + k4;
            af[mt][ks] = (b < B && c < m) ? (double)__ldg(s + b * m + c) : 0.0;
        }
    }
    __syncthreads();
    if (b0 >= B) return;
    const bool wide = ((ldw & 1) == 0) && ((reinterpret_cast<unsigned long long>(w) & 15ull) == 0);
    for (int nt = nt0; nt < nt1; ++nt) {
        const double *vr = vsm + ((nt - nt0) * 8 + r8) * kExpandVPitch + k4;
        double bf[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) bf[ks] = vr[4 * ks];
        double acc[kExpandMT][2];
#pragma unroll
        for (int mt = 0; mt < kExpandMT; ++mt) acc[mt][0] = acc[mt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int mt = 0; mt < kExpandMT; ++mt) dmma884(acc[mt][0], acc[mt][1], af[mt][ks], bf[ks]);
        const int col = 8 * nt + 2 * k4;
#pragma unroll
        for (int mt = 0; mt < kExpandMT; ++mt) {
            const long long b = b0 + 8 * mt + r8;
            if (b < B) {
                double *dst = w + b * ldw + col;
                if (wide && col + 1 < ldw) {
                    *reinterpret_cast<double2 *>(dst) = make_double2(acc[mt][0], acc[mt][1]);   // columns [d, ldw) get zeros
                } else {
                    if (col < d) dst[0] = acc[mt][0];
                    if (col + 1 < d) dst[1] = acc[mt][1];
                }
            }
        }
    }
}

}  // namespace gpde
