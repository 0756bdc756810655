// Score head of the batched-image configuration (BASELINE.json configs[3]): the scores that enter
// GrooMeD-NMS come from a shared linear layer + sigmoid over per-box features, so that a real parameter gradient exists for
// the one NCCL all-reduce of the step.  It stands for the reference's acceptance-probability head, a 1x1 convolution (= a
// per-anchor linear map of the 512 proposal features) followed by a sigmoid
// (models/densenet121_3d_dilate_decomp_alpha.py:112-121 builds it, :230 applies it).
//   forward : scores[m] = sigmoid(dot(x[m, :K], wb[:K]) + wb[K])
//   backward: grad_wb[k] = sum_m g[m] s[m] (1 - s[m]) x[m, k]   (k < K),   grad_wb[K] = sum_m g[m] s[m] (1 - s[m])
// Both are HBM streams over x (4 K bytes per box); the reduction is two-stage and ordered, so the gradient is deterministic.
#include "common.cuh"
#include <string.h>

namespace gnms {

constexpr int kHeadBlocks = 148 * 2;

template <int K>
__global__ void __launch_bounds__(256) score_head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wb,
                                                             float* __restrict__ scores, int64_t M) {
    constexpr int kLanes = K / 4;                         // lanes per box, one float4 each (K = 64: 16 lanes, 2 boxes per warp)
    static_assert(kLanes >= 1 && kLanes <= 32 && (kLanes & (kLanes - 1)) == 0, "K must be 4 * a power of two <= 128");
    const int lane = threadIdx.x & 31, sub = lane % kLanes;
    const float4 w = __ldg(reinterpret_cast<const float4*>(wb) + sub);
    const float bias = __ldg(wb + K);
    const int64_t per_warp = 32 / kLanes;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m0 = warp0 * per_warp; m0 < M; m0 += nwarps * per_warp) {
        const int64_t m = m0 + lane / kLanes;
        float acc = 0.f;
        if (m < M) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(x + m * K) + sub);
            acc = v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
        }
#pragma unroll
        for (int o = kLanes / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (m < M && sub == 0) scores[m] = 1.0f / (1.0f + __expf(-(acc + bias)));
    }
}

// stage 1: block b sums its contiguous slice of boxes into part[b][K + 1]; 256 threads = (256 / K) row groups x K columns
template <int K>
__global__ void __launch_bounds__(256) score_head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ scores,
                                                             const float* __restrict__ g, int64_t M, float* __restrict__ part) {
    constexpr int kGroups = 256 / K;
    __shared__ float red[kGroups][K + 1];
    const int c = threadIdx.x % K, rg = threadIdx.x / K;
    const int64_t per = (M + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = lo + per < M ? lo + per : M;
    float acc = 0.f, accb = 0.f;
    for (int64_t m = lo + rg; m < hi; m += kGroups) {
        const float s = __ldg(scores + m);
        const float coef = __ldg(g + m) * s * (1.0f - s);
        acc += coef * __ldcs(x + m * K + c);
        accb += coef;
    }
    red[rg][c] = acc;
    if (c == 0) red[rg][K] = accb;
    __syncthreads();
    if (rg == 0) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < kGroups; ++q) t += red[q][c];
        part[(size_t)blockIdx.x * (K + 1) + c] = t;
        if (c == 0) {
            float tb = 0.f;
#pragma unroll
            for (int q = 0; q < kGroups; ++q) tb += red[q][K];
            part[(size_t)blockIdx.x * (K + 1) + K] = tb;
        }
    }
}
// stage 2: one warp per parameter; lane l adds the partials of blocks l, l + 32, ... in block order, then a fixed shuffle tree
// (the same association every run: deterministic, like stage 1)
__global__ void __launch_bounds__(1024) score_head_reduce_kernel(const float* __restrict__ part, int nblocks, int K1, float* __restrict__ grad_wb) {
    const int lane = threadIdx.x & 31;
    for (int k = threadIdx.x >> 5; k < K1; k += blockDim.x >> 5) {
        float t = 0.f;
        for (int b = lane; b < nblocks; b += 32) t += part[(size_t)b * K1 + k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == 0) grad_wb[k] = t;
    }
}

// stage 2 fused with the collective: the block partials are reduced (as above) into this rank's slot of a small exchange buffer
// that every peer has mapped (CUDA IPC over NVLink / NVSwitch), a flag per source rank is released into every peer's buffer, and
// once all flags of this step have arrived every rank sums the W slots in rank order (remote loads over NVLink: W x 65 floats) --
// the same association on every rank, so all ranks hold bitwise the same all-reduced gradient.  One launch instead of the
// reduce kernel + an NCCL all-reduce of 272 bytes (21 us of collective latency at 4 GPUs, a few us here).
// Exchange buffer of a rank (kPeerBytes): two 128-float slots (step parity: a rank may run one step ahead of a peer that is still
// reading, never two), 16 flags (the last step each source rank has published), the step counter.
constexpr int kPeerSlotFloats = 128, kPeerMaxRanks = 16, kPeerBytes = 4096;
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__global__ void __launch_bounds__(1024) score_head_reduce_exchange_kernel(const float* __restrict__ part, int nblocks, int K1, gnms_peers P,
                                                                          float* __restrict__ grad_wb, int32_t* __restrict__ status) {
    __shared__ uint32_t s_step;
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31;
    char* mine = reinterpret_cast<char*>(P.buf[P.rank]);
    uint32_t* my_flags = reinterpret_cast<uint32_t*>(mine + 2 * kPeerSlotFloats * 4);
    uint32_t* counter = my_flags + kPeerMaxRanks;
    if (tid == 0) { s_step = *counter + 1u; s_bad = 0; }
    __syncthreads();
    const uint32_t step = s_step;
    float* slot = reinterpret_cast<float*>(mine) + (step & 1u) * kPeerSlotFloats;
    for (int k = tid >> 5; k < K1; k += blockDim.x >> 5) {
        float t = 0.f;
        for (int b = lane; b < nblocks; b += 32) t += part[(size_t)b * K1 + k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == 0) slot[k] = t;
    }
    __syncthreads();
    if (tid < P.world) {
        __threadfence_system();                                       // this rank's slot before its flag, system wide
        st_release_sys(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(P.buf[tid]) + 2 * kPeerSlotFloats * 4) + P.rank, step);
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(my_flags + tid) - step) < 0) {           // (flags only grow; wrap-safe compare)
            if (clock64() - t0 > 4000000000ll) { s_bad = 1; break; }           // ~2 s: a peer is gone -- report, do not hang
        }
    }
    __syncthreads();
    if (tid < K1) {
        float t = 0.f;
        for (int r = 0; r < P.world; ++r)
            t += ld_relaxed_sys(reinterpret_cast<const float*>(P.buf[r]) + (step & 1u) * kPeerSlotFloats + tid);
        grad_wb[tid] = s_bad ? __int_as_float(0x7fc00000) : t;
    }
    if (tid == 0) { *counter = step; if (s_bad && status) *status = 1; }
}

}  // namespace gnms

using namespace gnms;

extern "C" size_t gnms_score_head_workspace_bytes(int K) { return (size_t)kHeadBlocks * (size_t)(K + 1) * sizeof(float); }

extern "C" int gnms_score_head_forward_f32(const float* x, int64_t M, int K, const float* wb, float* scores, void* stream) {
    if (M < 0 || K != 64) return M < 0 ? GNMS_E_BADARG : GNMS_E_UNSUPPORTED;
    if (M == 0) return 0;
    if (!x || !wb || !scores) return GNMS_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(x) & 15u) || (reinterpret_cast<uintptr_t>(wb) & 15u)) return GNMS_E_ALIGN;
    int64_t blocks = (M / 2 + 7) / 8;                                 // 8 warps per block, 2 boxes per warp and trip
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    score_head_fwd_kernel<64><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, wb, scores, M);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_score_head_backward_f32(const float* x, int64_t M, int K, const float* scores, const float* grad_scores,
                                            float* grad_wb, void* workspace, void* stream) {
    if (M < 0 || K != 64) return M < 0 ? GNMS_E_BADARG : GNMS_E_UNSUPPORTED;
    if (!grad_wb || !workspace) return GNMS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (M == 0) { GNMS_CUDA_TRY(cudaMemsetAsync(grad_wb, 0, (size_t)(K + 1) * 4, s)); return 0; }
    if (!x || !scores || !grad_scores) return GNMS_E_BADARG;
    float* part = reinterpret_cast<float*>(workspace);
    score_head_bwd_kernel<64><<<kHeadBlocks, 256, 0, s>>>(x, scores, grad_scores, M, part);
    GNMS_LAUNCH_CHECK();
    score_head_reduce_kernel<<<1, 1024, 0, s>>>(part, kHeadBlocks, K + 1, grad_wb);
    GNMS_LAUNCH_CHECK();
    return 0;
}

// ---- peer exchange buffers (CUDA IPC): the plumbing under gnms_score_head_backward_allreduce_f32
extern "C" int gnms_peer_buffer_create(void** dev_ptr, unsigned char handle_out[64]) {
    if (!dev_ptr || !handle_out) return GNMS_E_BADARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    GNMS_CUDA_TRY(cudaMalloc(&p, kPeerBytes));
    GNMS_CUDA_TRY(cudaMemset(p, 0, kPeerBytes));
    cudaIpcMemHandle_t h;
    GNMS_CUDA_TRY(cudaIpcGetMemHandle(&h, p));
    memcpy(handle_out, &h, 64);
    *dev_ptr = p;
    return 0;
}
extern "C" int gnms_peer_buffer_open(const unsigned char handle[64], void** dev_ptr) {
    if (!handle || !dev_ptr) return GNMS_E_BADARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    GNMS_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int gnms_peer_buffer_close(void* dev_ptr, int owned) {
    if (!dev_ptr) return 0;
    if (owned) GNMS_CUDA_TRY(cudaFree(dev_ptr));
    else GNMS_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

extern "C" int gnms_score_head_backward_allreduce_f32(const float* x, int64_t M, int K, const float* scores, const float* grad_scores,
                                                      float* grad_wb, void* workspace, const gnms_peers* peers, int32_t* status,
                                                      void* stream) {
    if (M < 0 || K != 64) return M < 0 ? GNMS_E_BADARG : GNMS_E_UNSUPPORTED;
    if (!grad_wb || !workspace || !peers || peers->world < 1 || peers->world > kPeerMaxRanks || peers->rank < 0 || peers->rank >= peers->world)
        return GNMS_E_BADARG;
    for (int r = 0; r < peers->world; ++r) if (!peers->buf[r]) return GNMS_E_BADARG;
    if (M > 0 && (!x || !scores || !grad_scores)) return GNMS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    float* part = reinterpret_cast<float*>(workspace);
    // (an empty shard still takes part in the exchange: its partials are zero)
    if (M == 0) GNMS_CUDA_TRY(cudaMemsetAsync(part, 0, (size_t)kHeadBlocks * (K + 1) * sizeof(float), s));
    else score_head_bwd_kernel<64><<<kHeadBlocks, 256, 0, s>>>(x, scores, grad_scores, M, part);
    GNMS_LAUNCH_CHECK();
    score_head_reduce_exchange_kernel<<<1, 1024, 0, s>>>(part, kHeadBlocks, K + 1, *peers, grad_wb, status);
    GNMS_LAUNCH_CHECK();
    return 0;
}
