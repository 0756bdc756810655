// SoftSort variant of the score sort (reference lib/groomed_nms.py:131-165, `sorting_method="soft"`; Prillo & Eisenschlos,
// ICML 2020), forward and backward, hand-written:
//   h = sort(s, descending)                                        rank by counting (N^2 compares, N <= 8192)
//   E[i][j] = exp(-|s_j - h_i| / T)                                 (the reference subtracts the row maximum first, which is
//                                                                   exactly 0: every h_i is one of the s_j, :149-153)
//   S[i] = sum_j E[i][j] + 1e-3                                     (:154)
//   P[i][j] = E[i][j] / S[j]                                        the reference divides by the [N] vector of ROW sums, which
//                                                                   broadcasts along the last axis: column j is divided by
//                                                                   the sum of row j (:155, reproduced as written)
//   soft_scores = P s (:159),  soft_matrix = P M (:164, rows only)  the one dense contraction on the whole path
// The N x N x N product runs on the fp32 FMA pipe (128 x 128 x 8 tiles, 8 x 8 per thread, double-buffered shared memory):
// results must match the reference's fp32 matmul to 1e-5, which rules out TF32 / BF16 tensor-core inputs (10 / 7 mantissa bits).
#include "common.cuh"
#include <atomic>

namespace gnms {

// ------------------------------------------------------------------------------------------------ sort by counting
__global__ void __launch_bounds__(256) softsort_rank_kernel(const float* __restrict__ s, int N, float* __restrict__ h, int32_t* __restrict__ perm) {
    extern __shared__ float sh[];
    for (int j = threadIdx.x; j < N; j += 256) sh[j] = s[j];
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const float v = sh[i];
    int r = 0;
    for (int j = 0; j < N; ++j) {
        const float o = sh[j];
        r += (o > v) || (o == v && j < i);                    // stable: ties keep the lower index first
    }
    h[r] = v;
    perm[r] = i;
}

// ------------------------------------------------------------------------------------------------ E row sums
// one warp per row i: S[i] = sum_j exp(-|s_j - h_i| / T) + 1e-3
__global__ void __launch_bounds__(256) softsort_rowsum_kernel(const float* __restrict__ s, const float* __restrict__ h, int N, float inv_t,
                                                              float* __restrict__ S) {
    const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= N) return;
    const float hi = h[i];
    float acc = 0.f;
    for (int j = lane; j < N; j += 32) acc += expf(-fabsf(s[j] - hi) * inv_t);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) S[i] = acc + 1e-3f;
}

// P[i][j] = E[i][j] / S[j]; soft_scores[i] = sum_j P[i][j] s_j.  One warp per row, coalesced 128-byte row segments.
__global__ void __launch_bounds__(256) softsort_perm_kernel(const float* __restrict__ s, const float* __restrict__ h, const float* __restrict__ S,
                                                            int N, float inv_t, float* __restrict__ P, float* __restrict__ soft_scores) {
    const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= N) return;
    const float hi = h[i];
    float acc = 0.f;
    for (int j = lane; j < N; j += 32) {
        const float sj = s[j];
        const float p = expf(-fabsf(sj - hi) * inv_t) / S[j];
        P[(size_t)i * N + j] = p;
        acc = fmaf(p, sj, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) soft_scores[i] = acc;
}

// ------------------------------------------------------------------------------------------------ fp32 GEMM on the FMA pipe
// C[M,N] (+)= op(A) op(B), row-major storage, leading dimensions in floats.
//   kTA = false: A is [M,K] (lda),  true: A is stored [K,M] (C = A^T B)
//   kTB = false: B is [K,N] (ldb),  true: B is stored [N,K] (C = A B^T)
// 128 x 128 x 8 tiles, 256 threads, 8 x 8 results per thread, two shared-memory stages.
constexpr int kGM = 128, kGN = 128, kGK = 8;

template <bool kTA, bool kTB>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                    float* __restrict__ C, int ldc, int accumulate) {
    __shared__ __align__(16) float sA[2][kGK][kGM];                     // [k][m]
    __shared__ __align__(16) float sB[2][kGK][kGN];                     // [k][n]
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * kGM, n0 = blockIdx.x * kGN;
    const int tx = tid & 15, ty = tid >> 4;                             // 16 x 16 threads, each 8 x 8 (two 4-wide halves 64 apart)
    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
    // each thread loads 4 elements of the A tile and 4 of the B tile per k-step (128 x 8 = 1024 elements each)
    auto load_tile = [&](int k0, float (&ra)[4], float (&rb)[4]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + q * 256;                                // 0 .. 1023
            int m, k;
            if (kTA) { m = e & 127; k = e >> 7; } else { k = e & 7; m = e >> 3; }      // contiguous along the stored inner dimension
            const int gm = m0 + m, gk = k0 + k;
            ra[q] = (gm < M && gk < K) ? (kTA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk]) : 0.f;
            int n, kb;
            if (kTB) { kb = e & 7; n = e >> 3; } else { n = e & 127; kb = e >> 7; }
            const int gn = n0 + n, gkb = k0 + kb;
            rb[q] = (gn < N && gkb < K) ? (kTB ? B[(size_t)gn * ldb + gkb] : B[(size_t)gkb * ldb + gn]) : 0.f;
        }
    };
    auto store_tile = [&](int buf, const float (&ra)[4], const float (&rb)[4]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + q * 256;
            int m, k;
            if (kTA) { m = e & 127; k = e >> 7; } else { k = e & 7; m = e >> 3; }
            sA[buf][k][m] = ra[q];
            int n, kb;
            if (kTB) { kb = e & 7; n = e >> 3; } else { n = e & 127; kb = e >> 7; }
            sB[buf][kb][n] = rb[q];
        }
    };
    float ra[4], rb[4];
    load_tile(0, ra, rb);
    store_tile(0, ra, rb);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += kGK) {
        const bool more = k0 + kGK < K;
        if (more) load_tile(k0 + kGK, ra, rb);
#pragma unroll
        for (int k = 0; k < kGK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sA[buf][k][4 * ty]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sA[buf][k][64 + 4 * ty]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sB[buf][k][4 * tx]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sB[buf][k][64 + 4 * tx]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        if (more) {
            store_tile(buf ^ 1, ra, rb);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int gm = m0 + (a < 4 ? 4 * ty + a : 64 + 4 * ty + (a - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int gn = n0 + (b < 4 ? 4 * tx + b : 64 + 4 * tx + (b - 4));
            if (gn < N) {
                float* c = C + (size_t)gm * ldc + gn;
                *c = accumulate ? *c + acc[a][b] : acc[a][b];
            }
        }
    }
}

template <bool kTA, bool kTB>
static int launch_sgemm(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int accumulate, cudaStream_t s) {
    if (M <= 0 || N <= 0) return 0;
    dim3 grid(gnms_div_up(N, kGN), gnms_div_up(M, kGM));
    sgemm_kernel<kTA, kTB><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate);
    GNMS_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ backward, elementwise parts
// dP += g_scores[i] * s[j]   (the matvec P s, :159); in place on dP (which already holds g_P + g_SM M^T)
// then, with E = P S (column-wise), colterm[j] = sum_i dP[i][j] E[i][j] (needed for dS) -- one block per 32 columns
__global__ void __launch_bounds__(256) softsort_bwd_col_kernel(const float* __restrict__ P, const float* __restrict__ S, float* __restrict__ dP,
                                                               const float* __restrict__ g_scores, const float* __restrict__ s, int N,
                                                               float* __restrict__ dS) {
    __shared__ float red[8][32];
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    float acc = 0.f;
    if (j < N) {
        const float sj = s[j], Sj = S[j];
        for (int i = w; i < N; i += 8) {
            const size_t o = (size_t)i * N + j;
            const float d = g_scores ? fmaf(g_scores[i], sj, dP[o]) : dP[o];
            dP[o] = d;
            acc = fmaf(d, P[o] * Sj, acc);                              // dP * E
        }
    }
    red[w][threadIdx.x & 31] = acc;
    __syncthreads();
    if (w == 0 && j < N) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x & 31];
        const float Sj = S[j];
        dS[j] = -t / (Sj * Sj);                                         // dL/dS[j]
    }
}
// dE[i][j] = dP[i][j] / S[j] + dS[i];  W = dE * E / T;  sg = sign(s_j - h_i)
// ds[j] -= sum_i W sg   (column sums, atomics-free: one warp per row accumulates dh, columns through a second pass)
// pass 1 (one warp per row): dh[i] = sum_j W sg;  also stores W sg into dP (reused as scratch) for the column pass
__global__ void __launch_bounds__(256) softsort_bwd_row_kernel(const float* __restrict__ P, const float* __restrict__ S, float* __restrict__ dP,
                                                               const float* __restrict__ dS, const float* __restrict__ s, const float* __restrict__ h,
                                                               int N, float inv_t, float* __restrict__ dh) {
    const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= N) return;
    const float hi = h[i], dSi = dS[i];
    float acc = 0.f;
    for (int j = lane; j < N; j += 32) {
        const size_t o = (size_t)i * N + j;
        const float Sj = S[j];
        const float e = P[o] * Sj;
        const float dE = dP[o] / Sj + dSi;
        const float d = s[j] - hi;
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        const float wsg = dE * e * inv_t * sg;
        dP[o] = wsg;
        acc += wsg;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dh[i] = acc;
}
// pass 2 (one block per 32 columns): ds[j] = ds_in[j] - sum_i wsg[i][j];  then thread j adds dh through the permutation
__global__ void __launch_bounds__(256) softsort_bwd_colsum_kernel(const float* __restrict__ wsg, int N, float* __restrict__ ds) {
    __shared__ float red[8][32];
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    float acc = 0.f;
    if (j < N)
        for (int i = w; i < N; i += 8) acc += wsg[(size_t)i * N + j];
    red[w][threadIdx.x & 31] = acc;
    __syncthreads();
    if (w == 0 && j < N) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x & 31];
        ds[j] -= t;
    }
}
__global__ void softsort_bwd_perm_kernel(const float* __restrict__ dh, const int32_t* __restrict__ perm, int N, float* __restrict__ ds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) ds[perm[i]] += dh[i];                                    // perm is a permutation: no two threads share a target
}

}  // namespace gnms

using namespace gnms;

extern "C" size_t gnms_soft_sort_workspace_bytes(int N) { return (size_t)(N > 0 ? N : 1) * 4 * 4; }      // h, S, dS, dh

// forward: s[N], M[N,N] (ld_m) or NULL -> soft_scores[N], P[N,N], soft_matrix[N,N] (if M), perm int32[N] (sorted position ->
// input index), S[N] row sums (saved for the backward)
extern "C" int gnms_soft_sort_forward_f32(const float* s, int N, float temperature, const float* M, int64_t ld_m, float* soft_scores,
                                          float* P, float* soft_matrix, int32_t* perm, float* hsorted, float* S, void* stream) {
    if (N < 0 || !(temperature > 0.f)) return GNMS_E_BADARG;
    if (N > GNMS_MAX_BOXES) return GNMS_E_TOOLARGE;
    if (N == 0) return 0;
    if (!s || !soft_scores || !P || !perm || !hsorted || !S || (M && (!soft_matrix || ld_m < N))) return GNMS_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    static std::atomic<bool> attr_done[64];
    int dev = 0;
    GNMS_CUDA_TRY(cudaGetDevice(&dev));
    if (!attr_done[dev & 63].load(std::memory_order_acquire)) {
        GNMS_CUDA_TRY(cudaFuncSetAttribute(softsort_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GNMS_MAX_BOXES * 4));
        attr_done[dev & 63].store(true, std::memory_order_release);
    }
    const float inv_t = 1.0f / temperature;
    softsort_rank_kernel<<<gnms_div_up(N, 256), 256, (size_t)N * 4, st>>>(s, N, hsorted, perm);
    GNMS_LAUNCH_CHECK();
    softsort_rowsum_kernel<<<gnms_div_up(N, 8), 256, 0, st>>>(s, hsorted, N, inv_t, S);
    GNMS_LAUNCH_CHECK();
    softsort_perm_kernel<<<gnms_div_up(N, 8), 256, 0, st>>>(s, hsorted, S, N, inv_t, P, soft_scores);
    GNMS_LAUNCH_CHECK();
    if (M) return launch_sgemm<false, false>(N, N, N, P, N, M, (int)ld_m, soft_matrix, N, 0, st);
    return 0;
}

// backward: g_scores[N] / g_P[N,N] / g_SM[N,N] (each may be NULL) -> grad_s[N] (written), grad_M[N,N] (written if M given).
// scratch: dP[N,N] caller-allocated (destroyed), workspace gnms_soft_sort_workspace_bytes.
extern "C" int gnms_soft_sort_backward_f32(const float* s, int N, float temperature, const float* M, int64_t ld_m, const float* P,
                                           const int32_t* perm, const float* hsorted, const float* S, const float* g_scores,
                                           const float* g_P, const float* g_SM, float* grad_s, float* grad_M, float* dP_scratch,
                                           void* workspace, void* stream) {
    if (N < 0 || !(temperature > 0.f)) return GNMS_E_BADARG;
    if (N == 0) return 0;
    if (!s || !P || !perm || !hsorted || !S || !grad_s || !dP_scratch || !workspace || (g_SM && !M)) return GNMS_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    float* dS = reinterpret_cast<float*>(workspace);
    float* dh = dS + N;
    const size_t nn = (size_t)N * N * 4;
    // dP = g_P + g_SM M^T
    if (g_P) GNMS_CUDA_TRY(cudaMemcpyAsync(dP_scratch, g_P, nn, cudaMemcpyDeviceToDevice, st));
    else GNMS_CUDA_TRY(cudaMemsetAsync(dP_scratch, 0, nn, st));
    if (g_SM) {
        int rc = launch_sgemm<false, true>(N, N, N, g_SM, N, M, (int)ld_m, dP_scratch, N, 1, st);
        if (rc) return rc;
        if (grad_M) { rc = launch_sgemm<true, false>(N, N, N, P, N, g_SM, N, grad_M, N, 0, st); if (rc) return rc; }
    } else if (grad_M) {
        GNMS_CUDA_TRY(cudaMemsetAsync(grad_M, 0, nn, st));
    }
    // grad_s (direct term of the matvec) = P^T g_scores
    if (g_scores) { int rc = launch_sgemm<true, false>(N, 1, N, P, N, g_scores, 1, grad_s, 1, 0, st); if (rc) return rc; }
    else GNMS_CUDA_TRY(cudaMemsetAsync(grad_s, 0, (size_t)N * 4, st));
    softsort_bwd_col_kernel<<<gnms_div_up(N, 32), 256, 0, st>>>(P, S, dP_scratch, g_scores, s, N, dS);
    GNMS_LAUNCH_CHECK();
    softsort_bwd_row_kernel<<<gnms_div_up(N, 8), 256, 0, st>>>(P, S, dP_scratch, dS, s, hsorted, N, 1.0f / temperature, dh);
    GNMS_LAUNCH_CHECK();
    softsort_bwd_colsum_kernel<<<gnms_div_up(N, 32), 256, 0, st>>>(dP_scratch, N, grad_s);
    GNMS_LAUNCH_CHECK();
    softsort_bwd_perm_kernel<<<gnms_div_up(N, 256), 256, 0, st>>>(dh, perm, N, grad_s);
    GNMS_LAUNCH_CHECK();
    return 0;
}
