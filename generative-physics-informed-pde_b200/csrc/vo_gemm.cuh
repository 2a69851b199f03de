// FP64 tensor-core contraction  R[B,m] = rho[B,d] V[d,m]  for many weighting functions (m > 32), sm_100a.
// Included by vo.cu.
//
// This is the (Gamma y - alpha) step of VirtualObservables.py:61-69, 662, 990 once the fine residual rho is
// known; at BASELINE config 3 (d = 16383, m = 256, B = 16384) it is 137 GFLOP of FP64 and bounds the whole VO
// evaluation (FP64 pipe: 37 TFLOP/s measured => 3.7 ms floor; tcgen05 has no FP64 kind, so the tool is
// mma.sync.m8n8k4.f64 = DMMA.8x8x4, issued at the full FP64 rate).
//
// Layout: rho comes from the matvec kernel in a K-padded workspace [B][dp] (dp = d rounded up to 16, pad = 0);
// V is copied once per call into [dp][ldb] (ldb = m rounded up to the column tile, pad = 0), so every tile load
// is a full 16-byte cp.async with no bounds logic in the main loop.
// CTA tile 128 samples x BN columns, 8 warps as 4 (M) x 2 (N), warp tile 32 x BN/2 => 4 x BN/16 DMMA tiles
// whose accumulators stay in registers; K in chunks of 16 through a 4-stage cp.async ring.  Shared-memory row
// pitches are = 4 (mod 16) doubles, which makes both fragment loads conflict-free.
// (A 112-row tile on 7 warps, tried against the 256-tiles-on-148-SMs tail of config 3, is no faster: the FP64
// tensor pipe is per scheduler, and 7 warps leave one scheduler with half the work -- measured, round 1.)
// Tail: one CTA per SM, so T output tiles take ceil(T / #SM) tile-times (config 3: 256 tiles on 148 SMs = 2.0 for 1.73 of
// work).  The contraction length can be cut into S equal parts (gridDim.z = S, partial tiles to a workspace, summed in part
// order by vo_gemm_reduce_kernel: deterministic, the same split for every sample): ceil(S T / #SM) / S tile-times -- 1.75 at
// S = 4 for config 3.  gemm_splits() picks S.
#pragma once

namespace gpde {

constexpr int kGemmBM = 128, kGemmKC = 16, kGemmStages = 4, kGemmThreads = 256;
constexpr int kGemmLdA = kGemmKC + 4;   // doubles per A row in shared memory

__device__ __forceinline__ void cp_async_wait_n(int n) {
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    }
}

// Vp[dp][ldb] <- V[d][m], zero padded
__global__ void vo_gemm_pack_kernel(const double *__restrict__ V, int d, int m, double *__restrict__ Vp, int dp, int ldb) {
    const long long total = (long long)dp * ldb;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / ldb), c = (int)(idx - (long long)i * ldb);
        Vp[idx] = (i < d && c < m) ? V[(long long)i * m + c] : 0.0;
    }
}
template <typename T>
__global__ void vo_gemm_pack_kernel_t(const T *__restrict__ V, int d, int m, double *__restrict__ Vp, int dp, int ldb) {
    const long long total = (long long)dp * ldb;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / ldb), c = (int)(idx - (long long)i * ldb);
        Vp[idx] = (i < d && c < m) ? (double)V[(long long)i * m + c] : 0.0;
    }
}

// Operands of the transposed application  w[B,d] = s[B,m] V^T  on the same kernel:
//   Sp[B][mp] <- s[B][m] (K padded to 16 with zeros),  Vt[mp][ldb] <- V^T (rows k < m, columns n < d, zero padded)
__global__ void vo_gemm_pad_rows_kernel(const double *__restrict__ s, long long B, int m, double *__restrict__ Sp, int mp) {
    const long long total = B * mp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / mp;
        const int k = (int)(idx - b * mp);
        Sp[idx] = k < m ? s[b * m + k] : 0.0;
    }
}
__global__ void vo_gemm_pack_transposed_kernel(const double *__restrict__ V, int d, int m, double *__restrict__ Vt, int mp,
                                               int ldb) {
    // 32 x 32 tiles through shared memory: coalesced reads of V rows, coalesced writes of Vt rows
    __shared__ double tile[32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int n = n0 + r, k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (n < d && k < m) ? V[(long long)n * m + k] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, n = n0 + threadIdx.x;
        if (k < mp && n < ldb) Vt[(long long)k * ldb + n] = tile[threadIdx.x][r];
    }
}

// R[b][c] = sum_z P[z][b][c] (c < m), parts in ascending order
template <typename To>
__global__ void vo_gemm_reduce_kernel(const double *__restrict__ P, int splits, int ldp, To *__restrict__ R, int m, long long B) {
    const long long total = B * m;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / m;
        const int c = (int)(idx - b * m);
        double v = 0.0;
        for (int z = 0; z < splits; ++z) v += P[((long long)z * B + b) * ldp + c];
        R[idx] = (To)v;
    }
}

// ``partial`` != nullptr: gridDim.z parts of the contraction length; part z writes its [B][ldb] partial tile rows to
// partial + z * B * ldb (all ldb columns) and R is left to vo_gemm_reduce_kernel
template <int BN, typename To>
__global__ void __launch_bounds__(kGemmThreads, 1)
vo_gemm_kernel(const double *__restrict__ A, int dp, const double *__restrict__ Vp, int ldb, To *__restrict__ R, int m,
               long long B, double *__restrict__ partial = nullptr) {
    constexpr int LdB = BN + 4;                    // doubles per B row in shared memory
    constexpr int NT = BN / 16;                    // 8-column DMMA tiles per warp
    constexpr int A_STAGE = kGemmBM * kGemmLdA, B_STAGE = kGemmKC * LdB;
    extern __shared__ __align__(16) double gsm[];
    double *As = gsm, *Bs = gsm + kGemmStages * A_STAGE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp & 3, wn = warp >> 2;       // 4 x 2 warps
    const long long row0 = (long long)blockIdx.x * kGemmBM;
    const int col0 = blockIdx.y * BN;
    const int all_chunks = dp / kGemmKC;
    const int kc0 = (int)((long long)all_chunks * blockIdx.z / gridDim.z);            // this part's chunks [kc0, kc0 + nchunks)
    const int nchunks = (int)((long long)all_chunks * (blockIdx.z + 1) / gridDim.z) - kc0;

    // cp.async assignments: A chunk = 128 rows x 8 pieces, B chunk = 16 rows x BN/2 pieces (16 bytes each)
    auto load_chunk = [&](int kc, int stage) {
        double *as = As + stage * A_STAGE, *bs = Bs + stage * B_STAGE;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int p = threadIdx.x + i * kGemmThreads;      // 0..1023
            const int rr = p >> 3, pc = p & 7;
            long long row = row0 + rr;
            if (row >= B) row = B - 1;
            cp_async16(as + rr * kGemmLdA + 2 * pc, A + row * dp + (long long)(kc0 + kc) * kGemmKC + 2 * pc);
        }
#pragma unroll
        for (int i = 0; i < (kGemmKC * BN / 2) / kGemmThreads; ++i) {
            const int p = threadIdx.x + i * kGemmThreads;
            const int kr = p / (BN / 2), pc = p - kr * (BN / 2);
            cp_async16(bs + kr * LdB + 2 * pc, Vp + ((long long)(kc0 + kc) * kGemmKC + kr) * ldb + col0 + 2 * pc);
        }
    };

    double acc[4][NT][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int s = 0; s < kGemmStages - 1; ++s) {
        if (s < nchunks) load_chunk(s, s);
        cp_async_commit();
    }
    // fragment source of this lane inside a stage
    const int a_lane = (wm * 32 + (lane >> 2)) * kGemmLdA + (lane & 3);          // + mt*8*LdA + ks*4
    const int b_lane = (lane & 3) * LdB + wn * (BN / 2) + (lane >> 2);           // + ks*4*LdB + nt*8
    for (int kc = 0; kc < nchunks; ++kc) {
        cp_async_wait_n(kGemmStages - 2);
        __syncthreads();
        {   // refill the stage consumed in the previous iteration
            const int nk = kc + kGemmStages - 1;
            if (nk < nchunks) load_chunk(nk, nk % kGemmStages);
            cp_async_commit();
        }
        const double *as = As + (kc % kGemmStages) * A_STAGE + a_lane;
        const double *bs = Bs + (kc % kGemmStages) * B_STAGE + b_lane;
#pragma unroll
        for (int ks = 0; ks < kGemmKC / 4; ++ks) {
            double af[4], bf[NT];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) af[mt] = as[mt * 8 * kGemmLdA + ks * 4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bf[nt] = bs[ks * 4 * LdB + nt * 8];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    cp_async_wait_n(0);
    // epilogue: lane holds C[row = lane/4][cols 2*(lane%4), +1] of every 8x8 tile
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const long long row = row0 + wm * 32 + mt * 8 + (lane >> 2);
        if (row >= B) continue;
        if (partial) {
            double *prow = partial + ((long long)blockIdx.z * B + row) * ldb;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = col0 + wn * (BN / 2) + nt * 8 + 2 * (lane & 3);
                *reinterpret_cast<double2 *>(prow + col) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            }
            continue;
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int col = col0 + wn * (BN / 2) + nt * 8 + 2 * (lane & 3);
            if (col < m) R[row * m + col] = (To)acc[mt][nt][0];
            if (col + 1 < m) R[row * m + col + 1] = (To)acc[mt][nt][1];
        }
    }
}

static inline int gemm_bn(int m) { return m <= 64 ? 64 : 128; }
static inline int gemm_ldb(int m) { const int bn = gemm_bn(m); return (m + bn - 1) / bn * bn; }
static inline int gemm_dp(int d) { return (d + kGemmKC - 1) / kGemmKC * kGemmKC; }
// parts of the contraction length: the S in {1, 2, 4, 8} with the fewest tile-times ceil(S T / #SM) / S, ties to the smaller S;
// a split must save at least 6 % (its partial tiles cost a write, a read and one more launch)
static inline int gemm_splits(long long tiles, int n_sm, int nchunks) {
    int best = 1;
    double best_t = (double)((tiles + n_sm - 1) / n_sm);
    for (int s = 2; s <= 8; s *= 2) {
        if (nchunks < 64 * s) break;
        const double t = (double)((tiles * s + n_sm - 1) / n_sm) / s;
        if (t < 0.94 * best_t) { best = s; best_t = t; }
    }
    return best;
}
static inline size_t gemm_smem(int bn) {
    return sizeof(double) * (size_t)kGemmStages * ((size_t)kGemmBM * kGemmLdA + (size_t)kGemmKC * (bn + 4));
}

}  // namespace gpde
