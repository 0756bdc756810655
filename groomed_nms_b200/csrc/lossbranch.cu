// Front and back end of the GrooMeD branch of the detection loss, batched over the images of a step (SURVEY.md section 8(f)
// rank 1): what the reference does per image with a full sort, host round trips and six small launches
// (lib/loss/rpn_3d.py:731-744 and :801-825).
//   masked_topk_kernel   : the (at most) K highest-scoring foreground anchors of every image, in descending score order,
//                          ties by lower anchor index (:731-737: sort of the foreground scores, first 500)
//   best_box_per_gt_kernel: for every ground truth the candidate box maximising 0.5 (1 + GIoU3D) * IoU2D, marked as a positive
//                          target if that score exceeds best_target_box_beta (:813-822)
#include "common.cuh"

namespace gnms {

// ---------------------------------------------------------------------------------------------- masked top-K
// One CTA per image.  Keys are made unique (score bits, then inverted anchor index), so "the K largest keys" is a set and
// the result is the stable descending order.  Radix select, 8 bits per pass from the top, stops as soon as the bucket that
// holds the K-th key is wanted in full; then the selected keys are sorted in shared memory (bitonic, K_pad <= 1024).
constexpr int kTopkThreads = 1024;

__device__ __forceinline__ uint32_t float_order_key(float f) {         // larger float -> larger key; -0 < +0; NaN on top
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(kTopkThreads) masked_topk_kernel(const float* __restrict__ scores, const uint8_t* __restrict__ mask,
                                                                   int A, int K, int Kpad, int64_t* __restrict__ out_idx,
                                                                   int32_t* __restrict__ out_n) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);           // [Kpad]
    __shared__ unsigned hist[256];
    __shared__ unsigned s_count, s_pick, s_need;
    __shared__ unsigned long long s_prefix;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* sc = scores + (size_t)b * A;
    const uint8_t* mk = mask + (size_t)b * A;
    auto key_of = [&](int i) -> unsigned long long {
        return ((unsigned long long)float_order_key(sc[i]) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    };
    // number of foreground anchors
    if (tid == 0) s_count = 0;
    __syncthreads();
    unsigned local = 0;
    for (int i = tid; i < A; i += kTopkThreads) local += mk[i] ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((tid & 31) == 0 && local) atomicAdd(&s_count, local);
    __syncthreads();
    const unsigned nfg = s_count;
    const unsigned want = nfg < (unsigned)K ? nfg : (unsigned)K;
    // radix select of the `want`-th largest key: after the loop every key >= lo_bound is selected (exactly `want` keys)
    unsigned long long prefix = 0, lo_bound = 0;
    unsigned need = want;
    if (want < nfg) {
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            for (int i = tid; i < 256; i += kTopkThreads) hist[i] = 0;
            __syncthreads();
            const unsigned long long pmask = pass == 0 ? 0ull : (~0ull << (shift + 8));
            for (int i = tid; i < A; i += kTopkThreads) {
                if (!mk[i]) continue;
                const unsigned long long k = key_of(i);
                if ((k & pmask) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned acc = 0, pick = 0;
                int d = 255;
                for (; d >= 0; --d) {                                   // from the largest digit down
                    if (acc + hist[d] >= need) break;
                    acc += hist[d];
                }
                pick = (unsigned)d;
                s_pick = pick;
                s_need = need - acc;                                    // keys still wanted inside bucket `pick`
                s_prefix = prefix | ((unsigned long long)pick << shift);
            }
            __syncthreads();
            prefix = s_prefix;
            need = s_need;
            const unsigned in_bucket = hist[s_pick];
            __syncthreads();
            lo_bound = prefix;                                          // every key with these leading digits or larger
            if (need == in_bucket) break;                               // the whole bucket is wanted: done
        }
    }
    // collect the selected keys (unordered), pad, sort descending
    if (tid == 0) s_count = 0;
    for (int i = tid; i < Kpad; i += kTopkThreads) keys[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < A; i += kTopkThreads) {
        if (!mk[i]) continue;
        const unsigned long long k = key_of(i);
        if (want == nfg || k >= lo_bound) {
            const unsigned slot = atomicAdd(&s_count, 1u);
            if (slot < (unsigned)Kpad) keys[slot] = k;
        }
    }
    __syncthreads();
    for (int size = 2; size <= Kpad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < Kpad; i += kTopkThreads) {
                const int j = i ^ stride;
                if (j > i) {
                    const unsigned long long x = keys[i], y = keys[j];
                    const bool desc = (i & size) == 0;
                    if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[j] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < K; i += kTopkThreads)
        out_idx[(size_t)b * K + i] = (unsigned)i < want ? (int64_t)(0xffffffffu - (uint32_t)(keys[i] & 0xffffffffull)) : 0;
    if (tid == 0) out_n[b] = (int32_t)want;
}

// ---------------------------------------------------------------------------------------------- best box per ground truth
// One warp per ground truth.  Candidates: the image's first n boxes (records + 2D boxes, NMS order); score as the reference
// forms it, scores_with_gt = (0.5 * (1 + giou3d)) * iou2d with separately rounded fp32 ops (:817); torch.max semantics:
// first maximum, NaN is the maximum (and then fails the `> beta` test).
__global__ void __launch_bounds__(128) best_box_per_gt_kernel(const float* __restrict__ rec, const float* __restrict__ box2d,
                                                              const int32_t* __restrict__ n_img, int Kc,
                                                              const float* __restrict__ gt_rec, const float* __restrict__ gt_2d,
                                                              const int32_t* __restrict__ gt_image, int n_gt, float beta,
                                                              const int64_t* __restrict__ top, int A, float* __restrict__ targets,
                                                              int32_t* __restrict__ best_slot, float* __restrict__ best_score) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_gt) return;
    const int b = gt_image[g];
    const int n = min(n_img[b], Kc);
    const Rec3 gr = load_rec3(gt_rec + (size_t)g * 8);
    const Box2 g2 = make_box2(*reinterpret_cast<const float4*>(gt_2d + (size_t)g * 4));
    float best = -INFINITY;
    int arg = 0x7fffffff;
    bool has = false;
    for (int i = lane; i < n; i += 32) {
        const Rec3 r = load_rec3(rec + ((size_t)b * Kc + i) * 8);
        const Box2 q = make_box2(*reinterpret_cast<const float4*>(box2d + ((size_t)b * Kc + i) * 4));
        const float v3 = iou3_exact_slow<true, true>(r, gr);            // 0.5 * (1 + giou3d(box, gt))
        const float v = __fmul_rn(v3, iou2(q, g2));
        const bool better = !has || (v != v && best == best) || (best == best && v > best);
        if (better) { best = v; arg = i; has = true; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        const bool oh = __shfl_xor_sync(0xffffffffu, (int)has, o) != 0;
        if (!oh) continue;
        const bool o_nan = ob != ob, m_nan = best != best;
        bool take;
        if (!has) take = true;
        else if (o_nan != m_nan) take = o_nan;                           // NaN beats any number
        else if (o_nan) take = oa < arg;                                 // both NaN: first one
        else take = ob > best || (ob == best && oa < arg);               // first maximum
        if (take) { best = ob; arg = oa; has = true; }
    }
    if (lane == 0) {
        if (best_slot) best_slot[g] = has ? arg : -1;
        if (best_score) best_score[g] = has ? best : -INFINITY;
        if (has && best > beta) targets[(size_t)b * A + top[(size_t)b * Kc + arg]] = 1.0f;       // :820-825
    }
}

}  // namespace gnms

using namespace gnms;

extern "C" int gnms_masked_topk_f32(const float* scores, const uint8_t* mask, int A, int batch, int K, int64_t* out_idx,
                                    int32_t* out_n, void* stream) {
    if (A < 0 || batch < 0 || K <= 0 || K > 1024) return GNMS_E_BADARG;
    if (batch == 0) return 0;
    if (!scores || !mask || !out_idx || !out_n) return GNMS_E_BADARG;
    int Kpad = 2;
    while (Kpad < K) Kpad <<= 1;
    masked_topk_kernel<<<batch, kTopkThreads, (size_t)Kpad * 8, (cudaStream_t)stream>>>(scores, mask, A, K, Kpad, out_idx, out_n);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_best_box_per_gt_f32(const float* rec, const float* box2d, const int32_t* n_per_image, int K, const float* gt_rec,
                                        const float* gt_2d, const int32_t* gt_image, int n_gt, float beta, const int64_t* top,
                                        int A, float* targets, int32_t* best_slot, float* best_score, void* stream) {
    if (K < 0 || n_gt < 0 || A < 0) return GNMS_E_BADARG;
    if (n_gt == 0 || K == 0) return 0;
    if (!rec || !box2d || !n_per_image || !gt_rec || !gt_2d || !gt_image || !top || !targets) return GNMS_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(rec) & 15u) || (reinterpret_cast<uintptr_t>(box2d) & 15u) || (reinterpret_cast<uintptr_t>(gt_rec) & 15u) ||
        (reinterpret_cast<uintptr_t>(gt_2d) & 15u))
        return GNMS_E_ALIGN;
    best_box_per_gt_kernel<<<gnms_div_up(n_gt, 4), 128, 0, (cudaStream_t)stream>>>(rec, box2d, n_per_image, K, gt_rec, gt_2d, gt_image,
                                                                                  n_gt, beta, top, A, targets, best_slot, best_score);
    GNMS_LAUNCH_CHECK();
    return 0;
}
