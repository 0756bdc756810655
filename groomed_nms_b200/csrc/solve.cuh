// Triangular solves for the reference's "inverse" modes (included by gnms.cu).
//
//   GROUP_NOMASK (group_boxes=True, mask_group_boxes=False): per group g, pre_g = (I + Phi_g)^-1 s_g
//   NOGROUP      (group_boxes=False):                              pre   = (I + Phi)^-1   s
// with Phi = strictly-lower-triangular p(iou_sorted) (lib/groomed_nms.py:71-73,107,110).  The reference inverts
// with LU (torch.inverse) and multiplies; since I + Phi is unit lower triangular the same vector comes out of a
// forward substitution, and the backward pass is the transposed (backward) substitution
//   ds = (I + Phi)^-T g~ ,   dPhi[a][b] = -ds_a * pre_b  (a > b).
// Blocked by 32: the part of a block row that depends on already-solved blocks is accumulated by all warps (one
// row per warp, lanes over columns), the 32 x 32 diagonal block is solved by one warp with shuffles.  Entries of
// Phi are gathered on the fly from the overlap matrix (or re-evaluated from the box records): k^2/2 gathers per
// system of size k.  These modes are not on the shipped training path (default = GROUP_MASK); they are built for
// parity, not for the roofline.
#pragma once

namespace gnms {

struct OvSrc {
    int src;                  // BoxSrc
    const float* iou;         // kSrcMatrix: image base
    int64_t ld;
    const int32_t* order;     // sorted position -> input index
    const float* sbox;        // sorted box records (box sources)
    int generalized, affine;
    __device__ __forceinline__ float get(int pa, int pb) const {      // overlap[row = pa (later box), col = pb]
        if (src == kSrcMatrix) return iou[(int64_t)order[pa] * ld + order[pb]];
        const float4* A = reinterpret_cast<const float4*>(sbox + (size_t)pa * 8);
        const float4* B = reinterpret_cast<const float4*>(sbox + (size_t)pb * 8);
        const float4 a0 = A[0], a1 = A[1], c0 = B[0], c1 = B[1];
        if (src == kSrcBox3d) {
            Rec3 ra = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            Rec3 rb = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            const float ib = inter_bev3(ra, rb);
            return generalized ? (affine ? iou3<true, true>(ra, rb, ib) : iou3<true, false>(ra, rb, ib))
                               : (affine ? iou3<false, true>(ra, rb, ib) : iou3<false, false>(ra, rb, ib));
        }
        Box2 ra = {a0.x, a0.y, a0.z, a0.w, a1.x};
        Box2 rb = {c0.x, c0.y, c0.z, c0.w, c1.x};
        return iou2(ra, rb);
    }
};

struct SolveArgs {
    int N, batch, mode;
    const int32_t* n_per_image;
    char* ws;
    size_t ws_img_stride;
    OvSrc ov;                 // iou / order / sbox are image-0 bases; per-image strides below
    int64_t iou_img_stride;
    gnms_params p;
    const float* sorted_scores;   // [batch,N]
    int32_t* order;               // [batch,N]
    int32_t* lead;                // [batch,N]  (NOGROUP forward writes lead[pos] = pos)
    float* pre;                   // [batch,N]  forward: out; backward: in
    // backward only
    const float* grad_prob;
    const int32_t* slot;
    float* grad_scores;           // zero-filled by the launcher
    float* grad_iou;              // nullable, zero-filled by the caller
    int64_t ld_gi;
};

constexpr int kSolveThreads = 256;

// resolve the system this CTA owns: idx list (sorted positions, ascending) and its length
__device__ __forceinline__ bool solve_system(const SolveArgs& A, int b, int n, const int32_t*& idx, int& k) {
    if (A.mode == GNMS_MODE_NOGROUP) {
        idx = nullptr;            // identity
        k = n;
        return blockIdx.x == 0;
    }
    const WsLayout L = ws_layout(A.N);
    char* w = A.ws + (size_t)b * A.ws_img_stride;
    const int ng = *reinterpret_cast<const int32_t*>(w + L.ngroups);
    if ((int)blockIdx.x >= ng) return false;
    const int32_t* gbeg = reinterpret_cast<const int32_t*>(w + L.gbeg);
    idx = reinterpret_cast<const int32_t*>(w + L.members) + gbeg[blockIdx.x];
    k = gbeg[blockIdx.x + 1] - gbeg[blockIdx.x];
    return true;
}

__global__ void __launch_bounds__(kSolveThreads) solve_fwd_kernel(SolveArgs A) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.y, N = A.N;
    const int n = A.n_per_image ? min(A.n_per_image[b], N) : N;
    const int32_t* idx;
    int k;
    if (!solve_system(A, b, n, idx, k)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kSolveThreads / 32;
    OvSrc ov = A.ov;
    ov.iou = A.ov.iou ? A.ov.iou + (size_t)b * A.iou_img_stride : nullptr;
    ov.order = A.order + (size_t)b * N;
    ov.sbox = reinterpret_cast<const float*>(A.ws + (size_t)b * A.ws_img_stride + ws_layout(N).sbox);
    const float* ss = A.sorted_scores + (size_t)b * N;
    float* pre = A.pre + (size_t)b * N;
    float* x = sm;                 // [k] solved values
    float* acc = sm + ((k + 3) & ~3);   // [32]
    const gnms_params P = A.p;
    auto at = [&](int a) { return idx ? idx[a] : a; };
    for (int blk0 = 0; blk0 < k; blk0 += 32) {
        const int nb = min(32, k - blk0);
        // rows of this block against everything already solved: one row per warp, lanes over columns
        for (int a = warp; a < nb; a += nwarps) {
            const int pa = at(blk0 + a);
            float sum = 0.f;
            for (int c = lane; c < blk0; c += 32)
                sum = __fmaf_rn(prune(ov.get(pa, at(c)), P.pruning_method, P.nms_threshold, P.temperature), x[c], sum);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
            if (lane == 0) acc[a] = sum;
        }
        __syncthreads();
        if (warp == 0) {
            float phi[32];
            const int pa = lane < nb ? at(blk0 + lane) : 0;
#pragma unroll
            for (int c = 0; c < 32; ++c)
                phi[c] = (lane < nb && c < lane) ? prune(ov.get(pa, at(blk0 + c)), P.pruning_method, P.nms_threshold, P.temperature) : 0.f;
            float mine = lane < nb ? __fsub_rn(ss[pa], acc[lane]) : 0.f;      // s_a - (already solved part)
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float xc = __shfl_sync(0xffffffffu, mine, c);           // final once all c' < c were applied
                if (lane > c) mine = __fmaf_rn(-phi[c], xc, mine);
            }
            if (lane < nb) { x[blk0 + lane] = mine; pre[pa] = mine; }
        }
        __syncthreads();
    }
    if (A.mode == GNMS_MODE_NOGROUP) {
        int32_t* lead = A.lead + (size_t)b * N;
        for (int pos = tid; pos < N; pos += kSolveThreads) lead[pos] = pos < n ? pos : -1;
    }
    if (blockIdx.x == 0 && tid == 0) {       // remember the overlap source for the backward (box sources have no iou)
        int32_t* meta = reinterpret_cast<int32_t*>(A.ws + (size_t)b * A.ws_img_stride + ws_layout(N).ngroups) + 1;
        meta[0] = A.ov.src; meta[1] = A.ov.generalized; meta[2] = A.ov.affine;
    }
}

__global__ void __launch_bounds__(kSolveThreads) solve_bwd_kernel(SolveArgs A) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.y, N = A.N;
    const int n = A.n_per_image ? min(A.n_per_image[b], N) : N;
    const int32_t* idx;
    int k;
    if (!solve_system(A, b, n, idx, k)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kSolveThreads / 32;
    OvSrc ov = A.ov;
    ov.iou = A.ov.iou ? A.ov.iou + (size_t)b * A.iou_img_stride : nullptr;
    ov.order = A.order + (size_t)b * N;
    ov.sbox = reinterpret_cast<const float*>(A.ws + (size_t)b * A.ws_img_stride + ws_layout(N).sbox);
    if (!ov.iou) {                            // forward ran from boxes: source kind was left in the workspace
        const int32_t* meta = reinterpret_cast<const int32_t*>(A.ws + (size_t)b * A.ws_img_stride + ws_layout(N).ngroups) + 1;
        ov.src = meta[0]; ov.generalized = meta[1]; ov.affine = meta[2];
    }
    const size_t o = (size_t)b * N;
    const gnms_params P = A.p;
    const float vthr = P.valid_box_prob_threshold;
    float* y = sm;                          // [k]
    float* acc = sm + ((k + 3) & ~3);       // [32]
    auto at = [&](int a) { return idx ? idx[a] : a; };
    // g~ = g * 1[0 <= pre <= 1] (* 1[r >= valid_thr] when the returned vector is the thresholded / sorted one)
    for (int a = tid; a < k; a += kSolveThreads) {
        const int pa = at(a);
        const float pre = A.pre[o + pa];
        float g = P.sorted_output ? A.grad_prob[o + A.slot[o + pa]] : A.grad_prob[o + pa];
        bool pass = (pre >= 0.f) && (pre <= 1.f);
        if (P.sorted_output || P.thresholded_output) pass = pass && (fminf(fmaxf(pre, 0.f), 1.f) >= vthr);
        y[a] = pass ? g : 0.f;
    }
    __syncthreads();
    // backward substitution with (I + Phi)^T, blocks from the end
    const int nblk = (k + 31) / 32;
    for (int bi = nblk - 1; bi >= 0; --bi) {
        const int blk0 = bi * 32, nb = min(32, k - blk0), done0 = blk0 + nb;
        for (int a = warp; a < nb; a += nwarps) {              // column a against the rows solved so far
            const int pa = at(blk0 + a);
            float sum = 0.f;
            for (int r = done0 + lane; r < k; r += 32)
                sum = __fmaf_rn(prune(ov.get(at(r), pa), P.pruning_method, P.nms_threshold, P.temperature), y[r], sum);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
            if (lane == 0) acc[a] = sum;
        }
        __syncthreads();
        if (warp == 0) {
            float phi[32];                                       // phi[r] = Phi[blk0 + r][blk0 + lane], r > lane
            const int pa = lane < nb ? at(blk0 + lane) : 0;
#pragma unroll
            for (int r = 0; r < 32; ++r)
                phi[r] = (r < nb && lane < nb && r > lane) ? prune(ov.get(at(blk0 + r), pa), P.pruning_method, P.nms_threshold, P.temperature) : 0.f;
            float mine = lane < nb ? __fsub_rn(y[blk0 + lane], acc[lane]) : 0.f;
#pragma unroll
            for (int r = 31; r >= 0; --r) {
                const float yr = __shfl_sync(0xffffffffu, mine, r);
                if (lane < r) mine = __fmaf_rn(-phi[r], yr, mine);
            }
            if (lane < nb) y[blk0 + lane] = mine;
        }
        __syncthreads();
    }
    for (int a = tid; a < k; a += kSolveThreads) A.grad_scores[o + ov.order[at(a)]] = y[a];
    if (A.grad_iou) {
        // dPhi[a][c] = -ds_a * pre_c  (a > c, "pre", not the clamped value);  d iou = dPhi * p'(iou)
        float* gi = A.grad_iou + (size_t)b * N * A.ld_gi;
        const int64_t total = (int64_t)k * (k - 1) / 2;
        for (int64_t e = tid; e < total; e += kSolveThreads) {
            int a = (int)((1.0 + sqrt(1.0 + 8.0 * (double)e)) * 0.5);
            while ((int64_t)a * (a - 1) / 2 > e) --a;
            while ((int64_t)(a + 1) * a / 2 <= e) ++a;
            const int c = (int)(e - (int64_t)a * (a - 1) / 2);
            const int pa = at(a), pc = at(c);
            const float v = ov.get(pa, pc);
            const float pv = prune(v, P.pruning_method, P.nms_threshold, P.temperature);
            const float dpv = prune_grad(v, pv, P.pruning_method, P.temperature);
            gi[(int64_t)ov.order[pa] * A.ld_gi + ov.order[pc]] = __fmul_rn(-__fmul_rn(y[a], A.pre[o + pc]), dpv);
        }
    }
}

}  // namespace gnms
