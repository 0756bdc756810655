// GrooMeD-NMS forward / backward, get_groups and classical hard NMS for sm_100a.
//
// Reference semantics: lib/groomed_nms.py:10-129 (differentiable_nms), :208-270 (get_groups),
// lib/nms/nms_kernel.cu:24-144 (+ cpu_nms.pyx, py_cpu_nms.py, nms_others.py:119-150).
//
// Pipeline per image (all on one stream, no host sync, no allocation):
//   1. rank_kernel        rank by counting (multi-CTA): order / rank / sorted scores (+ box records gathered into
//                         sorted order); zeroes the bitmask and partials.
//   2. mask_*_kernel      whole grid: the "suppression bitmask" in SORTED space,
//                             mask[l][jw] bit r  <=>  box at sorted position j = 32*jw + r leaves the pool when
//                                                     the box at sorted position l < j is picked as a leader,
//                                                     i.e. !(overlap[j,l] <= thr)   (lib/groomed_nms.py:249-250)
//                         either by streaming the N x N overlap matrix once (coalesced 16 B loads, HBM bound:
//                         4*N^2 bytes) or by evaluating the overlaps on the fly from the box records (lower
//                         triangle only, ALU bound, nothing N^2 touches HBM).  Also emits, per row word, whether
//                         any box of the word has an earlier overlapper at all (-> the certain leaders L0).
//   3. chain_kernel       one CTA: greedy leader election.  The certain leaders L0 are applied in parallel;
//                         what is left unresolved is processed in windows of 64 candidates whose mask columns
//                         are fetched together and resolved exactly, in score order, by one warp.  Then first
//                         suppressor -> group, in-group rank (group_size cap), gather of iou[m,leader],
//                         pruning function, rescore, clamp, threshold, valid / invalid index lists.
//   4. backward_kernel    one CTA: analytic vector-Jacobian product.
#include "common.cuh"

#include <cuda.h>
#include <limits.h>
#include <atomic>

namespace gnms {

constexpr int kChainThreads = 1024;
constexpr int kWin = 64;                 // candidates resolved per window
constexpr int kMaskThreads = 256;

// ------------------------------------------------------------------------------------------ workspace
constexpr int kBlkWords = 9 * 64;        // one 64-box block of the blocked SoA copy: 8 fields + rank
constexpr int kBlkBytes = kBlkWords * 4;

// Store record `src` (nf floats; nullptr: a harmless unit box) and `rank` as element e of a 64-box block.
__device__ __forceinline__ void blk_store(float* blk, int e, const float* src, int nf, int rank) {
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        // dummy: (ymin,ymax,bx1,bx2,bz1,bz2,vol,abev) = (0,1,0,1,0,1,1,1) in 3D, (x1,y1,x2,y2) = (0,0,1,1) in 2D
        const float dummy = nf == 8 ? ((f & 1) || f >= 6 ? 1.f : 0.f) : (f >= 2 ? 1.f : 0.f);
        if (f < nf) blk[f * 64 + e] = src ? src[f] : dummy;
    }
    // 2D: field 4 carries the area (x2-x1)*(y2-y1) for the tile kernel's range check
    if (nf == 4) blk[4 * 64 + e] = src ? __fmul_rn(__fsub_rn(src[2], src[0]), __fsub_rn(src[3], src[1])) : 1.f;
    reinterpret_cast<int32_t*>(blk)[8 * 64 + e] = rank;
}

struct WsLayout {
    size_t rank, sbox, mask, has_earlier, gbeg, members, ngroups, blk, gtab, dleader, dfsup, dflag, total;   // byte offsets inside one image's slice
    int he_slots;                                   // partial has-earlier words per row word (one per column-chunk CTA)
};
__host__ __device__ inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
__host__ __device__ inline WsLayout ws_layout(int N) {
    WsLayout L;
    size_t nw = (size_t)((N + 31) / 32);
    size_t off = 0;
    L.rank = off;        off += align_up((size_t)N * 4);
    L.sbox = off;        off += align_up((size_t)N * 8 * 4);
    L.mask = off;        off += align_up(nw * (size_t)N * 4);
    // "row word has an earlier overlapper" partials: slot = column-chunk CTA of the mask kernels (no atomics, no
    // shared flags: same-address global traffic from thousands of CTAs serialises in L2 and cost 4x the kernel)
    L.he_slots = (N + 255) / 256;
    L.has_earlier = off; off += align_up(nw * (size_t)L.he_slots * 4);
    // group structure for the per-group solves of mode GROUP_NOMASK: gbeg[g]..gbeg[g+1] indexes members[]
    L.gbeg = off;        off += align_up(((size_t)N + 1) * 4);
    L.members = off;     off += align_up((size_t)N * 4);
    L.ngroups = off;     off += 256;
    // blocked structure-of-arrays copy of the boxes for the tile kernel: per run of 64 boxes, 8 field arrays + the rank
    // array of 64 words each (2304 contiguous bytes = one bulk copy); input order, or spatial order in the culled pass
    L.blk = off;         off += align_up((size_t)((N + 63) / 64) * kBlkWords * 4);
    // per 64-box group of the spatial order: AABB, max extents, bad flag (10 floats), written by the spatial kernel
    L.gtab = off;        off += align_up((size_t)((N + 63) / 64) * 10 * 4);
    // direct leader election (elect_kernel): leader bitset, first suppressor per sorted position, 1 = "gave up" flag
    L.dleader = off;     off += align_up(nw * 4);
    L.dfsup = off;       off += align_up((size_t)N * 4);
    L.dflag = off;       off += 256;
    L.total = off;
    return L;
}

enum BoxSrc { kSrcMatrix = 0, kSrcBox2d = 1, kSrcBox3d = 2, kSrcBoxShift = 3 };

// bits of word w that belong to boxes below n
__device__ __forceinline__ uint32_t valid_word(int w, int n) {
    int lo = w * 32;
    if (n >= lo + 32) return 0xffffffffu;
    if (n <= lo) return 0u;
    return (1u << (n - lo)) - 1u;
}


// ------------------------------------------------------------------------------------------ 1b. rank by counting
// Multi-CTA replacement of the one-CTA bitonic sort: rank[i] = #{j : score_j > score_i or (== and j < i)}.
// grid (ceil(N/64), batch), 256 threads: 4 threads share one element and each scans a quarter of the keys from
// shared memory with 16-byte broadcast loads.  N^2 compares, but spread over the whole chip (~1 us per image at
// N=4096) instead of ~45 us of barrier-bound bitonic stages on one SM.  Also zeroes the has-earlier partials and,
// if asked, the suppression bitmask (the tile kernel below sets its bits with atomics).
constexpr int kRankElems = 64;
__global__ void __launch_bounds__(256)
rank_kernel(const float* __restrict__ scores, int64_t sstride, int64_t score_img_stride, int N,
            const int32_t* __restrict__ n_per_image, int32_t* __restrict__ order_out, float* __restrict__ ss_out,
            char* __restrict__ ws, size_t ws_img_stride, const float* __restrict__ boxes, int box_src,
            int64_t box_img_stride, float shift, int presorted, int zero_mask, int32_t* tile_count) {
    extern __shared__ __align__(16) uint32_t skeys[];
    const int b = blockIdx.y;
    const int n = n_per_image ? min(n_per_image[b], N) : N;
    const WsLayout L = ws_layout(N);
    char* w = ws + (size_t)b * ws_img_stride;
    int32_t* rank = reinterpret_cast<int32_t*>(w + L.rank);
    float* sbox = reinterpret_cast<float*>(w + L.sbox);
    uint32_t* has_earlier = reinterpret_cast<uint32_t*>(w + L.has_earlier);
    uint32_t* mask = reinterpret_cast<uint32_t*>(w + L.mask);
    const float* sc = scores + (size_t)b * score_img_stride;
    int32_t* order = order_out + (size_t)b * N;
    float* ss = ss_out + (size_t)b * N;
    const int tid = threadIdx.x;
    const int npad = (N + 15) & ~15;
    if (presorted == 0)                                              // (the other modes do not count)
        for (int j = tid; j < npad; j += 256) skeys[j] = j < n ? desc_key(sc[(int64_t)j * sstride]) : 0xffffffffu;
    // chores, split over the CTAs of this image
    const int nw = (N + 31) / 32;
    const size_t ncta = gridDim.x, me = blockIdx.x;
    if (presorted != 2)                                              // (sort_kernel's helper CTAs did the zeroing)
        for (size_t i = me * 256 + tid; i < (size_t)nw * L.he_slots; i += ncta * 256) has_earlier[i] = 0u;
    if (zero_mask && presorted != 2) {
        uint4* m4 = reinterpret_cast<uint4*>(mask);
        const size_t tot4 = ((size_t)nw * N + 3) / 4;                  // the mask slice is 256-byte aligned and padded
        for (size_t i = me * 256 + tid; i < tot4; i += ncta * 256) m4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (blockIdx.x == 0 && presorted != 2) {
        for (int pos = n + tid; pos < N; pos += 256) { order[pos] = -1; ss[pos] = 0.f; rank[pos] = INT_MAX; }
    }
    if (blockIdx.x == 0) {
        if (tile_count && b == 0 && tid == 0) { tile_count[0] = 0; tile_count[1] = 0; tile_count[2] = 0; }   // culled-tile count, work counters
    }
    __syncthreads();
    const int e = tid >> 2, q = tid & 3;
    const int i = blockIdx.x * kRankElems + e;
    // this CTA's 64 elements are exactly one block of the blocked SoA copy the tile kernel stages with bulk copies
    const bool want_blk = zero_mask && boxes && (box_src == kSrcBox2d || box_src == kSrcBox3d);
    float* blk = reinterpret_cast<float*>(w + L.blk) + (size_t)blockIdx.x * kBlkWords;
    const int blk_nf = box_src == kSrcBox3d ? 8 : 4;
    if (i >= n) {                                                     // whole quads leave together
        // dead slots: boxes past the live count keep their record (rank INT_MAX: never emit), slots past N a unit box
        if (want_blk && q == 0)
            blk_store(blk, e, i < N ? boxes + (size_t)b * box_img_stride + (size_t)i * blk_nf : nullptr, blk_nf, INT_MAX);
        return;
    }
    int r;
    if (presorted == 2) {
        r = rank[i];                                                  // sort_kernel ran first
    } else if (presorted) {
        r = i;
    } else {
        // count keys that sort before (key_i, i): key_j < key_i, or key_j == key_i with j < i.  The four threads of
        // an element take interleaved 16-byte groups (conflict-free, broadcast across the 8 elements of a warp);
        // groups entirely below i use `<= key_i` (as `< key_i + 1`), groups above use `<`, i's own group is split.
        const uint32_t ki = skeys[i];
        const uint32_t kle = ki == 0xffffffffu ? ki : ki + 1u;
        const int i4 = i >> 2, n4 = npad >> 2;
        const uint4* k4 = reinterpret_cast<const uint4*>(skeys);
        // a < b as the borrow of a - b, folded into the accumulator by subtract-with-borrow: two integer adds per key
        // (compare + select + add costs three); four independent accumulators.  Accumulators count DOWN.
        auto acc_lt = [](uint32_t& c, uint32_t a, uint32_t b) {
            uint32_t t;
            asm("{\n\t"
                "sub.cc.u32 %1, %2, %3;\n\t"
                "subc.u32 %0, %0, 0;\n\t"
                "}" : "+r"(c), "=r"(t) : "r"(a), "r"(b));
        };
        uint32_t c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
        int j4 = q;
#pragma unroll 4
        for (; j4 < i4; j4 += 4) {
            const uint4 k = k4[j4];
            acc_lt(c0, k.x, kle); acc_lt(c1, k.y, kle); acc_lt(c2, k.z, kle); acc_lt(c3, k.w, kle);
        }
        if (j4 == i4) {
            const uint4 k = k4[j4];
            const int sub = i & 3;
            acc_lt(c0, k.x, sub > 0 ? kle : ki); acc_lt(c1, k.y, sub > 1 ? kle : ki); acc_lt(c2, k.z, sub > 2 ? kle : ki); acc_lt(c3, k.w, ki);
            j4 += 4;
        }
#pragma unroll 4
        for (; j4 < n4; j4 += 4) {
            const uint4 k = k4[j4];
            acc_lt(c0, k.x, ki); acc_lt(c1, k.y, ki); acc_lt(c2, k.z, ki); acc_lt(c3, k.w, ki);
        }
        int cnt = -(int)(c0 + c1 + c2 + c3);
        const unsigned quad_mask = 0xfu << ((tid & 31) & ~3);        // the 4 lanes of this element (they exit together)
        cnt += __shfl_xor_sync(quad_mask, cnt, 1);
        cnt += __shfl_xor_sync(quad_mask, cnt, 2);
        r = cnt;
    }
    if (q != 0) return;
    if (presorted != 2) {
        order[r] = i;
        rank[i] = r;
        ss[r] = sc[(int64_t)i * sstride];
    }
    const float* bx = boxes ? boxes + (size_t)b * box_img_stride : nullptr;
    float4* o = reinterpret_cast<float4*>(sbox + (size_t)r * 8);
    if (box_src == kSrcBox2d) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(bx) + i);
        o[0] = v;
        o[1] = make_float4(make_box2(v).area, 0.f, 0.f, 0.f);
    } else if (box_src == kSrcBox3d) {
        const float4* src = reinterpret_cast<const float4*>(bx + (size_t)i * 8);
        o[0] = __ldg(src);
        o[1] = __ldg(src + 1);
    }
    if (want_blk) blk_store(blk, e, bx + (size_t)i * blk_nf, blk_nf, r);
    if (box_src == kSrcBoxShift) {
        const float* d = bx + (size_t)i * 5;
        const BoxS qb = make_boxs(d[0], d[1], d[2], d[3], shift);
        o[0] = make_float4(qb.x1, qb.y1, qb.x2, qb.y2);
        o[1] = make_float4(qb.area, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------ 1c. rank by sorting
// One CTA per image: stable LSD radix sort (4 passes of 8 bits) of the (descending-order key, index) pairs in shared
// memory.  The counting kernel above does batch * N^2 compares spread over the whole chip (~2.8 us per image at
// N = 4096, ALU-pipe bound); this one takes ~10 us whatever the batch size because images sort side by side on
// different SMs, so it wins as soon as a launch holds more than a few images.  Writes order / rank / sorted scores
// (and the dead tail); rank_kernel then runs in gather-only mode (presorted = 2) for its other chores.
// Per pass: keys are read striped per warp (element = warp base + 32 k + lane), ranked inside the warp among the lanes
// holding the same digit (8 ballots), warp-private digit counters in shared memory, one CTA-wide exclusive scan over
// (digit, warp), then a scatter into the other buffer.  Stable: ties keep the lower input index first.
// CTAs past the first `batch` ones zero the suppression masks and has-earlier partials of all images meanwhile.
constexpr int kSortThreads = 1024;
__host__ __device__ inline int sort_npad(int N) { return (N + kSortThreads - 1) / kSortThreads * kSortThreads; }
static size_t sort_smem_bytes(int N) { return (size_t)sort_npad(N) * 12 + 32 * 256 * 2 + 256; }

__global__ void __launch_bounds__(kSortThreads)
sort_kernel(const float* __restrict__ scores, int64_t score_img_stride, int N, int batch, const int32_t* __restrict__ n_per_image,
            int32_t* __restrict__ order_out, float* __restrict__ ss_out, char* __restrict__ ws, size_t ws_img_stride,
            int zero_mask) {
    extern __shared__ __align__(16) unsigned char s_sort[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (b >= batch) {                                                // helper CTAs: zero the per-image scratch
        const WsLayout L = ws_layout(N);
        const size_t nw = (size_t)((N + 31) / 32);
        const size_t he4 = (nw * L.he_slots * 4 + 15) / 16, mk4 = zero_mask ? (nw * N * 4 + 15) / 16 : 0;   // slices are 256-byte padded
        const size_t per_img = he4 + mk4, total = per_img * batch;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (size_t i = (size_t)(b - batch) * kSortThreads + tid; i < total; i += (size_t)(gridDim.x - batch) * kSortThreads) {
            const size_t img = i / per_img, o = i - img * per_img;
            char* base = ws + img * ws_img_stride;
            uint4* dst = o < he4 ? reinterpret_cast<uint4*>(base + L.has_earlier) + o : reinterpret_cast<uint4*>(base + L.mask) + (o - he4);
            *dst = z;
        }
        return;
    }
    const int n = n_per_image ? min(n_per_image[b], N) : N;
    const int npad = sort_npad(N), kper = npad / kSortThreads;
    uint32_t* keyA = reinterpret_cast<uint32_t*>(s_sort);
    uint32_t* keyB = keyA + npad;
    uint16_t* idxA = reinterpret_cast<uint16_t*>(keyB + npad);
    uint16_t* idxB = idxA + npad;
    uint16_t* cnt = idxB + npad;                                     // [32 warps][256 digits]
    __shared__ int s_wsum[32];
    const float* sc = scores + (size_t)b * score_img_stride;
    const WsLayout L = ws_layout(N);
    int32_t* rank = reinterpret_cast<int32_t*>(ws + (size_t)b * ws_img_stride + L.rank);
    for (int j = tid; j < npad; j += kSortThreads) { keyA[j] = j < n ? desc_key(sc[j]) : 0xffffffffu; idxA[j] = (uint16_t)j; }
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        uint32_t* c32 = reinterpret_cast<uint32_t*>(cnt);
        for (int j = tid; j < 32 * 256 / 2; j += kSortThreads) c32[j] = 0u;
        __syncthreads();
        // rank of every key among the keys of its warp that carry the same digit, in (k, lane) order
        int myrank[8];
        uint16_t* wc = cnt + w * 256;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            myrank[k] = 0;
            if (k < kper) {
                const uint32_t d = (keyA[w * 32 * kper + k * 32 + lane] >> shift) & 255u;
                // lanes holding the same digit: 8 ballots (MATCH.ANY measured ~43 cycles per warp instruction per SM)
                unsigned peers = 0xffffffffu;
#pragma unroll
                for (int bit = 0; bit < 8; ++bit) {
                    const bool one = (d >> bit) & 1u;
                    const unsigned bal = __ballot_sync(0xffffffffu, one);
                    peers &= one ? bal : ~bal;
                }
                const int leader = __ffs(peers) - 1;
                int base = 0;
                if (lane == leader) { base = wc[d]; wc[d] = (uint16_t)(base + __popc(peers)); }
                base = __shfl_sync(0xffffffffu, base, leader);
                myrank[k] = base + __popc(peers & lt_mask);
                __syncwarp();
            }
        }
        __syncthreads();
        // exclusive scan of the counters in (digit, warp) order: thread t owns digit t / 4, warps 8 (t % 4) .. + 7
        {
            const int d = tid >> 2, w0 = (tid & 3) * 8;
            int loc[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { loc[q] = cnt[(w0 + q) * 256 + d]; sum += loc[q]; }
            int incl = sum;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const int t2 = __shfl_up_sync(0xffffffffu, incl, dd);
                if (lane >= dd) incl += t2;
            }
            if (lane == 31) s_wsum[w] = incl;
            __syncthreads();
            if (w == 0) {
                const int v = s_wsum[lane];
                int iv = v;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const int t2 = __shfl_up_sync(0xffffffffu, iv, dd);
                    if (lane >= dd) iv += t2;
                }
                s_wsum[lane] = iv - v;
            }
            __syncthreads();
            int run = s_wsum[w] + incl - sum;
#pragma unroll
            for (int q = 0; q < 8; ++q) { cnt[(w0 + q) * 256 + d] = (uint16_t)run; run += loc[q]; }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < kper) {
                const int e = w * 32 * kper + k * 32 + lane;
                const uint32_t key = keyA[e];
                const int pos = wc[(key >> shift) & 255u] + myrank[k];
                keyB[pos] = key;
                idxB[pos] = idxA[e];
            }
        }
        __syncthreads();
        uint32_t* tk = keyA; keyA = keyB; keyB = tk;
        uint16_t* ti = idxA; idxA = idxB; idxB = ti;
    }
    int32_t* order = order_out + (size_t)b * N;
    float* ss = ss_out + (size_t)b * N;
    for (int r = tid; r < N; r += kSortThreads) {
        if (r < n) {
            const int i = idxA[r];
            order[r] = i; rank[i] = r; ss[r] = sc[i];
        } else {
            order[r] = -1; ss[r] = 0.f; rank[r] = INT_MAX;          // dead tail: input index r >= n
        }
    }
}

// ------------------------------------------------------------------------------------------ 2a. mask from a matrix
// grid (ceil(N/1024), NW, batch), 256 threads; thread = 4 consecutive INPUT columns x the 32 rows of row word jw.
template <bool kVec>
__global__ void __launch_bounds__(kMaskThreads)
mask_matrix_kernel(const float* __restrict__ iou, int64_t ld, int64_t img_stride, int N,
                   const int32_t* __restrict__ n_per_image, const int32_t* __restrict__ order_all,
                   char* __restrict__ ws, size_t ws_img_stride, float thr) {
    __shared__ int32_t rows[32];
    const int b = blockIdx.z;
    const int n = n_per_image ? min(n_per_image[b], N) : N;
    const int jw = blockIdx.y;
    if (jw * 32 >= n) return;
    const int NWm = (N + 31) / 32;
    const WsLayout L = ws_layout(N);
    char* w = ws + (size_t)b * ws_img_stride;
    const int32_t* rank = reinterpret_cast<const int32_t*>(w + L.rank);
    uint32_t* mask = reinterpret_cast<uint32_t*>(w + L.mask);
    uint32_t* has_earlier = reinterpret_cast<uint32_t*>(w + L.has_earlier);
    __shared__ uint32_t s_any;
    if (threadIdx.x == 0) s_any = 0u;
    const int32_t* order = order_all + (size_t)b * N;
    const float* m = iou + (size_t)b * img_stride;
    if (threadIdx.x < 32) {
        int pos = jw * 32 + threadIdx.x;
        rows[threadIdx.x] = pos < n ? order[pos] : -1;
    }
    __syncthreads();
    const int c0 = (blockIdx.x * kMaskThreads + threadIdx.x) * 4;
    uint32_t wd[4] = {0u, 0u, 0u, 0u};
    int rk[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
    if (c0 + 4 <= n) {
        const int4 q = *reinterpret_cast<const int4*>(rank + c0);
        rk[0] = q.x; rk[1] = q.y; rk[2] = q.z; rk[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c0 + k < n) rk[k] = rank[c0 + k];
    }
    const int rkmin = min(min(rk[0], rk[1]), min(rk[2], rk[3]));
    const int pos_last = min(jw * 32 + 31, n - 1);
    const int nrows = pos_last - jw * 32 + 1;                 // live rows of this word (1..32)
    // a thread whose 4 columns all come later than every row of this word has nothing to contribute
    if (rkmin < pos_last) {
        if (kVec && c0 + 4 <= N && nrows == 32) {
            // fast path: 4 batches of 8 rows, all eight 16-byte streaming loads of a batch issued before use
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {
                float4 q[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) q[u] = ld_cs_f4(m + (int64_t)rows[r0 + u] * ld + c0);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int r = r0 + u, pos = jw * 32 + r;
                    wd[0] |= (uint32_t)(!(q[u].x <= thr) && rk[0] < pos) << r;
                    wd[1] |= (uint32_t)(!(q[u].y <= thr) && rk[1] < pos) << r;
                    wd[2] |= (uint32_t)(!(q[u].z <= thr) && rk[2] < pos) << r;
                    wd[3] |= (uint32_t)(!(q[u].w <= thr) && rk[3] < pos) << r;
                }
            }
        } else {
            for (int r = 0; r < nrows; ++r) {
                const int pos = jw * 32 + r;
                const float* src = m + (int64_t)rows[r] * ld + c0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (c0 + k < n) {
                        const float v = __ldcs(src + k);
                        wd[k] |= (uint32_t)(!(v <= thr) && rk[k] < pos) << r;
                    }
                }
            }
        }
    }
    uint32_t any = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (rk[k] != INT_MAX && (rk[k] >> 5) <= jw) {         // words above the diagonal are never read
            mask[(size_t)rk[k] * NWm + jw] = wd[k];      // column-major: a leader's column is contiguous
        }
        any |= wd[k];
    }
    any = __reduce_or_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0 && any) atomicOr(&s_any, any);          // shared memory
    __syncthreads();
    // one partial per (row word, column-chunk CTA): this kernel's chunk is 1024 columns = 4 slots of 256
    if (threadIdx.x == 0) has_earlier[(size_t)jw * L.he_slots + blockIdx.x * 4] = s_any;
}

// ------------------------------------------------------------------------------------------ 2b. mask from boxes
// grid (ceil(N/256), NW, batch), 256 threads; thread = one SORTED column l x the 32 rows of row word jw.  Only
// tiles that touch the lower triangle (some l < some j) do any work.
template <int kSrc, bool kGen, bool kAffine, int kCmp>
__global__ void __launch_bounds__(kMaskThreads)
mask_boxes_kernel(int N, const int32_t* __restrict__ n_per_image, char* __restrict__ ws, size_t ws_img_stride,
                  float thr, float shift) {
    __shared__ float4 rrec[32][2];
    const int b = blockIdx.z;
    const int n = n_per_image ? min(n_per_image[b], N) : N;
    const int jw = blockIdx.y;
    const int l0 = blockIdx.x * kMaskThreads;
    if (jw * 32 >= n || l0 > jw * 32 + 31 || l0 >= n) return;
    const WsLayout L = ws_layout(N);
    char* w = ws + (size_t)b * ws_img_stride;
    const float* sbox = reinterpret_cast<const float*>(w + L.sbox);
    uint32_t* mask = reinterpret_cast<uint32_t*>(w + L.mask);
    uint32_t* has_earlier = reinterpret_cast<uint32_t*>(w + L.has_earlier);
    __shared__ uint32_t s_any;
    if (threadIdx.x == 0) s_any = 0u;
    if (threadIdx.x < 64) {
        int r = threadIdx.x >> 1, h = threadIdx.x & 1;
        int pos = jw * 32 + r;
        rrec[r][h] = pos < n ? reinterpret_cast<const float4*>(sbox + (size_t)pos * 8)[h] : make_float4(0, 0, 0, 0);
    }
    __syncthreads();
    const int l = l0 + threadIdx.x;
    uint32_t wd = 0u;
    if (l < n && l < jw * 32 + 31) {
        const float4 c0 = reinterpret_cast<const float4*>(sbox + (size_t)l * 8)[0];
        const float4 c1 = reinterpret_cast<const float4*>(sbox + (size_t)l * 8)[1];
        const int rend = min(32, n - jw * 32);
#pragma unroll 4
        for (int r = 0; r < rend; ++r) {
            const float4 a0 = rrec[r][0], a1 = rrec[r][1];
            float v;
            if (kSrc == kSrcBox3d) {
                Rec3 ra = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                Rec3 rb = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                v = iou3<kGen, kAffine>(ra, rb, inter_bev3(ra, rb));
            } else if (kSrc == kSrcBox2d) {
                Box2 ra = {a0.x, a0.y, a0.z, a0.w, a1.x};
                Box2 rb = {c0.x, c0.y, c0.z, c0.w, c1.x};
                v = iou2(ra, rb);
            } else {
                BoxS ra = {a0.x, a0.y, a0.z, a0.w, a1.x};
                BoxS rb = {c0.x, c0.y, c0.z, c0.w, c1.x};
                v = iou_shift(rb, ra, shift);   // (kept box, candidate) order of lib/nms/py_cpu_nms.py:26-33
            }
            bool hit = (kCmp == GNMS_CMP_GT) ? (v > thr) : (kCmp == GNMS_CMP_GE) ? (v >= thr) : !(v <= thr);
            uint32_t bit = hit && (l < jw * 32 + r);
            wd |= bit << r;
        }
    }
    if (l < n && (l >> 5) <= jw) {
        mask[(size_t)l * ((N + 31) / 32) + jw] = wd;
    }
    uint32_t any = __reduce_or_sync(0xffffffffu, wd);
    if ((threadIdx.x & 31) == 0 && any) atomicOr(&s_any, any);          // shared memory
    __syncthreads();
    if (threadIdx.x == 0) has_earlier[(size_t)jw * L.he_slots + blockIdx.x] = s_any;
}

// ------------------------------------------------------------------------------------------ 2c. fused overlap + mask tiles
// The north-star kernel: every unordered pair of boxes of an image is evaluated ONCE (symmetric 64 x 64 tiles, values
// register-resident) and feeds two consumers:
//   * the API-visible overlap matrix (optional): the tile and its mirror image are streamed to HBM straight from
//     registers with 16-byte stores (algorithmic 4 N^2 bytes, never read back), and
//   * the grouping stage: for a pair with !(v <= thr) the later-ranked box gets its bit set in the earlier box's
//     mask column (sorted space) with a fire-and-forget atomic OR -- ~3 % of the pairs on clustered boxes.
// Design notes (each measured on B200, see profiles/ and tools/exp/):
//   * persistent CTAs walk the tile list; the two 64-box blocks of the next tile are prefetched with 16-byte cp.async
//     into a double buffer from the blocked structure-of-arrays copy the rank / spatial kernel wrote (2304 contiguous
//     bytes per block: 8 field arrays + the rank array), so a thread fetches the 4 rows and 4 columns of a sub-tile
//     with conflict-free 16-byte shared loads, one per field;
//   * a thread owns 4 CONSECUTIVE rows x 4 consecutive columns: both the direct row segment and the mirrored one are
//     16-byte register vectors, so the transposed tile needs no shared-memory round trip and no second barrier
//     (warp = 8 column quads x 4 row quads: a direct store covers 4 rows x 128 B, a mirrored store 8 rows x 64 B);
//   * the 16 pair evaluations of a sub-tile are one straight-line block (the rare exact-division fallback and the
//     hits are handled after it), so independent division chains interleave;
//   * hits: matrix-producing launches append (row quad, column quad, 16 hit bits) to a shared-memory queue (one
//     warp-aggregated atomic per warp and sub-tile) that the whole CTA drains one tile later, one entry per thread --
//     the work no longer follows the lanes that happened to find the hits, which took 8 % off the kernel and halved
//     the time at the per-tile barrier; the hit-dense culled pass keeps the inline loop (the queue is slower there);
//   * matrix-producing launches use 128-thread CTAs (a thread owns two sub-tiles, 4 CTAs per SM): the per-tile barrier
//     joins 4 warps instead of 8, which was worth 10 % over the 256-thread shape; 64-thread CTAs, 5 or 6 CTAs per SM at 96 / 80
//     registers, several tiles per barrier, a warp-autonomous mbarrier/bulk-copy variant and a shared-memory ring drained by bulk stores
//     (cp.async.bulk shared -> global) were all slower and are not kept.
// One barrier per tile.  fp32-issue bound: ~67 issue slots per pair, 16 of them FMNMX at half rate.
struct TileArgs {
    int N, batch, nt, tiles_per_image, vec;
    const int32_t* n_per_image;
    const float* boxes;          // [batch, N, 8] 3D records or [batch, N, 4] 2D boxes (input order)
    char* ws;
    size_t ws_img_stride;
    float* out;                  // [batch, N, N] or nullptr
    float thr;
    const int32_t* tile_list;    // culled pass: [0] = count, [64..] = entries (blocks are then in spatial order)
    float inv_tpi, inv_w;        // 1 / tiles_per_image, 1 / (nt + 1): tile index decode without integer division
    int tiles_per_cta;           // matrix-only launches: > 0 = CTA c takes tiles [c k, (c+1) k) and retires (not persistent)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int kTT = 64;              // tile edge

__device__ __forceinline__ void tile_decode(int t, int nt, int& I, int& J) {
    float fn = 2.0f * nt + 1.0f;
    int i = (int)((fn - sqrtf(fn * fn - 8.0f * (float)t)) * 0.5f);
    i = max(0, min(i, nt - 1));
    while (i > 0 && i * nt - i * (i - 1) / 2 > t) --i;
    while ((i + 1) * nt - (i + 1) * i / 2 <= t) ++i;
    I = i;
    J = i + (t - (i * nt - i * (i - 1) / 2));
}

// Division-light enumeration of the nt(nt+1)/2 upper-triangle tiles: row i (nt - i tiles) is paired with row nt-1-i
// (i + 1 tiles) into one line of nt + 1 entries, so a flat index decodes with one reciprocal multiply (inv_w = 1/(nt+1)).
__device__ __forceinline__ int div_small(int t, int d, float inv_d) {      // t / d for t < 2^24-ish, d > 0
    int q = __float2int_rd(__int2float_rn(t) * inv_d);
    const int rem = t - q * d;
    q += (rem >= d) - (rem < 0);
    return q;
}
__device__ __forceinline__ void tile_decode_folded(int t, int nt, float inv_w, int& I, int& J) {
    const int i = div_small(t, nt + 1, inv_w), j = t - i * (nt + 1);
    if (j < nt - i) { I = i; J = i + j; }
    else { I = nt - 1 - i; J = I + (j - (nt - i)); }
}

template <int kSrc> struct RecOf;
template <> struct RecOf<kSrcBox3d> {
    typedef Rec3 type;
    static __device__ __forceinline__ Rec3 load(const float* p) { return load_rec3(p); }
    template <bool G, bool Aff> static __device__ __forceinline__ float fast(const Rec3& a, const Rec3& b, bool& u) {
        return iou3_fast<G, Aff>(a, b, inter_bev3(a, b), u);
    }
    template <bool G, bool Aff> static __device__ __forceinline__ float exact(const Rec3& a, const Rec3& b) {
        return iou3_exact_slow<G, Aff>(a, b);
    }
};
template <> struct RecOf<kSrcBox2d> {
    typedef Box2 type;
    static __device__ __forceinline__ Box2 load(const float* p) { return make_box2(*reinterpret_cast<const float4*>(p)); }
    template <bool G, bool Aff> static __device__ __forceinline__ float fast(const Box2& a, const Box2& b, bool& u) {
        return iou2_fast(a, b, u);
    }
    template <bool G, bool Aff> static __device__ __forceinline__ float exact(const Box2& a, const Box2& b) {
        return iou2_exact_slow(a, b);
    }
};

// Records of 4 consecutive boxes from a staged block ([field][64] arrays).
template <int kSrc> struct SoaOf;
template <> struct SoaOf<kSrcBox3d> {
    static constexpr int kFields = 8;
    // records of boxes first..first+3 from the SoA staging area (abev, field 7, is not needed by the 3D overlap)
    static __device__ __forceinline__ void load4(const float* soa, int first, Rec3 (&o)[4]) {
        float4 f[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) f[q] = *reinterpret_cast<const float4*>(soa + q * kTT + first);
        const float* p = reinterpret_cast<const float*>(f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[k].ymin = p[0 * 4 + k]; o[k].ymax = p[1 * 4 + k]; o[k].bx1 = p[2 * 4 + k]; o[k].bx2 = p[3 * 4 + k];
            o[k].bz1 = p[4 * 4 + k]; o[k].bz2 = p[5 * 4 + k]; o[k].vol = p[6 * 4 + k]; o[k].abev = 0.f;
        }
    }
};
template <> struct SoaOf<kSrcBox2d> {
    static constexpr int kFields = 4;
    static __device__ __forceinline__ void load4(const float* soa, int first, Box2 (&o)[4]) {
        float4 f[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) f[q] = *reinterpret_cast<const float4*>(soa + q * kTT + first);
        const float* p = reinterpret_cast<const float*>(f);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = make_box2(make_float4(p[0 * 4 + k], p[1 * 4 + k], p[2 * 4 + k], p[3 * 4 + k]));
    }
};

template <int kSrc, bool kGen, bool kAffine, bool kHasOut, bool kList, int kRB, bool kQueue>
__global__ void __launch_bounds__(256 / kRB, 2 * kRB) tile_kernel(TileArgs A) {
    constexpr int kThreads = 256 / kRB;                               // kRB row blocks of 64 / kRB rows per thread
    static_assert(!(kHasOut && kList), "the culled pass does not produce the matrix");
    typedef typename RecOf<kSrc>::type RecT;
    constexpr int kChunks = kBlkBytes / 16;                           // 16-byte chunks of one staged block (144)
    constexpr int kBufs = kQueue ? 3 : 2;                             // queued hits read the previous tile's rank arrays
    __shared__ __align__(16) float s_blk[kBufs][2][kBlkWords];      // [buffer][row/col][8 field arrays + rank array][64]
    __shared__ int s_ij[kBufs][4];
    __shared__ uint32_t s_q[kQueue ? 3 : 1][kQueue ? 256 : 1];      // hit queue of a tile: (row quad, column quad, 16 hit bits)
    __shared__ int s_nq[3];
    const int N = A.N, tid = threadIdx.x;
    const int NWt = (N + 31) / 32;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx = (lane & 7) + 8 * (warp & 1), ty = (lane >> 3) + 4 * (warp >> 1);      // 16 column quads x 16/kRB row quads
    const WsLayout L = ws_layout(N);
    const int total = kList ? A.tile_list[0] : A.tiles_per_image * A.batch;

    // Every thread decodes the tile (a handful of instructions) and copies 288 / kThreads 16-byte chunks of the two
    // 64-box blocks (blocked structure-of-arrays copies written by the rank / spatial kernel: 2304 contiguous bytes per
    // block, so staging is a few wide coalesced cp.async per thread and all warps carry the same work into the barrier).
    auto prefetch = [&](int t, int buf) {
        int b, I, J;
        if (kList) {
            const int e = A.tile_list[64 + t];
            b = e >> 16; I = (e >> 8) & 255; J = e & 255;
        } else {
            b = div_small(t, A.tiles_per_image, A.inv_tpi);
            tile_decode_folded(t - b * A.tiles_per_image, A.nt, A.inv_w, I, J);
        }
        const char* blk = A.ws + (size_t)b * A.ws_img_stride + L.blk;
#pragma unroll
        for (int c = tid; c < 2 * kChunks; c += kThreads) {
            const int side = c >= kChunks ? 1 : 0, cc = c - side * kChunks;
            cp_async16(reinterpret_cast<char*>(s_blk[buf][side]) + cc * 16, blk + (size_t)(side ? J : I) * kBlkBytes + cc * 16);
        }
        if (tid == 0) { s_ij[buf][0] = b; s_ij[buf][1] = I; s_ij[buf][2] = J; }
    };
    // preconditions of the straight-line division, checked on the chunks this thread staged itself (visible after
    // wait_all): coordinates |x| <= 2^19 (3D), volume / area within [2^-60, 2^60]
    auto own_bad = [&](int buf) -> bool {
        bool bad = false;
#pragma unroll
        for (int c = tid; c < 2 * kChunks; c += kThreads) {
            const int side = c >= kChunks ? 1 : 0, cc = c - side * kChunks, f = cc >> 4;
            const float4 q = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(s_blk[buf][side]) + cc * 16);
            if (kSrc == kSrcBox3d) {
                if (f < 6) bad = bad || !(fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fmaxf(fabsf(q.z), fabsf(q.w))) <= 524288.0f);
                else if (f == 6) bad = bad || outside_safe(q.x) || outside_safe(q.y) || outside_safe(q.z) || outside_safe(q.w);
            } else {
                if (f == 4) bad = bad || outside_safe(q.x) || outside_safe(q.y) || outside_safe(q.z) || outside_safe(q.w);
            }
        }
        return bad;
    };
    // hits -> suppression bits (sorted space).  Off-diagonal tiles see every unordered pair once; the diagonal tile sees
    // (i,j) and (j,i): only the orientation "row is the later box" emits.  Dead slots (box index >= live count, padding)
    // carry rank INT_MAX; ri == rj only for a box with itself.
    auto emit = [&](uint32_t hits, const int* rrank4, const int* crank4, bool mirror, uint32_t* mask) {
        while (hits) {
            const int bit = __ffs(hits) - 1;
            hits &= hits - 1u;
            const int ri = rrank4[bit >> 2], rj = crank4[bit & 3];
            const int later = max(ri, rj), earlier = min(ri, rj);
            if (later != INT_MAX && ri != rj && (mirror || ri > rj))
                atomicOr(mask + ((unsigned)earlier * (unsigned)NWt + (unsigned)(later >> 5)), 1u << (later & 31));
        }
    };
    // queued hits of the tile staged in buffer qb (its rank arrays are still there: three staging buffers): one entry per
    // thread, so the work is spread evenly over the CTA instead of following the lanes that happened to find the hits
    auto drain = [&](int qb) {
        const int nq = s_nq[qb];
        if (nq == 0) return;
        const int b = s_ij[qb][0];
        const bool mirror = s_ij[qb][1] != s_ij[qb][2];
        uint32_t* mask = reinterpret_cast<uint32_t*>(A.ws + (size_t)b * A.ws_img_stride + L.mask);
        const int* rrank = reinterpret_cast<const int*>(s_blk[qb][0]) + 8 * kTT;
        const int* crank = reinterpret_cast<const int*>(s_blk[qb][1]) + 8 * kTT;
        for (int e = tid; e < nq; e += kThreads) {
            const uint32_t ent = s_q[qb][e];
            emit(ent & 0xffffu, rrank + ((ent >> 20) & 15u) * 4, crank + ((ent >> 16) & 15u) * 4, mirror, mask);
        }
    };

    // persistent CTAs stride over the tile list
    const int t_step = (int)gridDim.x;
    const int t_end = total;
    int t = (int)blockIdx.x, buf = 0;
    if (kQueue && tid < 3) s_nq[tid] = 0;
    if (t < t_end) prefetch(t, 0);
    for (; t < t_end; t += t_step) {
        cp_async_wait_all();
        const bool bad = own_bad(buf);
        // publishes the staged blocks (and the previous tile's hit queue); every thread is done with the previous tile
        const bool tile_unsafe = __syncthreads_or(bad);
        const int nbuf = buf + 1 == kBufs ? 0 : buf + 1;              // next tile's buffer
        if (kQueue) {
            const int pbuf = buf == 0 ? 2 : buf - 1;                  // previous tile's buffer == next-next tile's buffer
            drain(pbuf);
            if (tid == 0) s_nq[nbuf] = 0;                             // last read one barrier ago; pushed to after the next one
        }
        const int b = s_ij[buf][0], I = s_ij[buf][1], J = s_ij[buf][2];
        if (t + t_step < t_end) prefetch(t + t_step, nbuf);

        const float* rsoa = s_blk[buf][0];
        const float* csoa = s_blk[buf][1];
        const int* rrank = reinterpret_cast<const int*>(rsoa) + 8 * kTT;
        const int* crank = reinterpret_cast<const int*>(csoa) + 8 * kTT;
        RecT cr[4];
        SoaOf<kSrc>::load4(csoa, 4 * tx, cr);
        const float thr = A.thr;
        float* out = kHasOut ? A.out + (size_t)b * N * N : nullptr;
        uint32_t* mask = reinterpret_cast<uint32_t*>(A.ws + (size_t)b * A.ws_img_stride + L.mask);
        const bool full_tile = kHasOut && A.vec && (I * kTT + kTT <= N) && (J * kTT + kTT <= N);
#pragma unroll 1
        for (int h = 0; h < kRB; ++h) {
            const int rl = (kTT / kRB) * h + 4 * ty;                 // first of this thread's 4 rows inside the tile
            RecT rr[4];
            SoaOf<kSrc>::load4(rsoa, rl, rr);
            float v[4][4];
            bool unsafe = tile_unsafe;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) v[r][k] = RecOf<kSrc>::template fast<kGen, kAffine>(rr[r], cr[k], unsafe);
            if (__builtin_expect(unsafe, 0)) {      // rare: outside div_rn_fast's proven range -> exact IEEE path
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[r][k] = RecOf<kSrc>::template exact<kGen, kAffine>(rr[r], cr[k]);
            }
            uint32_t hits = 0u;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) hits |= (uint32_t)(!(v[r][k] <= thr)) << (4 * r + k);

            // ---- consumer 2: the overlap matrix (optional), direct and mirrored, straight from registers
            if (kHasOut) {
                const int i0 = I * kTT + rl, j0 = J * kTT + 4 * tx;
                if (full_tile) {
                    float* drow = out + (int64_t)i0 * N + j0;
#pragma unroll
                    for (int r = 0; r < 4; ++r) st_cs_f4(drow + (int64_t)r * N, make_float4(v[r][0], v[r][1], v[r][2], v[r][3]));
                    if (I != J) {
                        float* dcol = out + (int64_t)j0 * N + i0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) st_cs_f4(dcol + (int64_t)k * N, make_float4(v[0][k], v[1][k], v[2][k], v[3][k]));
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (i0 + r < N && j0 + k < N) {
                                out[(int64_t)(i0 + r) * N + (j0 + k)] = v[r][k];
                                if (I != J) out[(int64_t)(j0 + k) * N + (i0 + r)] = v[r][k];
                            }
                        }
                }
            }
            // ---- consumer 1: threshold hits (~3 % of the pairs)
            if (kQueue) {
                // warp-aggregated append: one shared-memory atomic per warp and sub-tile
                const unsigned who = __ballot_sync(0xffffffffu, hits != 0u);
                if (who) {
                    const int leader = __ffs(who) - 1;
                    int base = 0;
                    if (lane == leader) base = atomicAdd(&s_nq[buf], __popc(who));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (hits) s_q[buf][base + __popc(who & ((1u << lane) - 1u))] = ((uint32_t)(rl >> 2) << 20) | ((uint32_t)tx << 16) | hits;
                }
            } else {
                emit(hits, rrank + rl, crank + 4 * tx, I != J, mask);
            }
        }
        buf = nbuf;
    }
    if (kQueue) {
        __syncthreads();
        drain(buf == 0 ? 2 : buf - 1);                               // the last tile's queue
    }
}

// ------------------------------------------------------------------------------------------ 2c-tall. matrix-only, tall tiles
// The matrix-only launch with tiles of 64 kQ rows x 64 columns (kQ = 4: 256 x 64 is the default): a thread evaluates
// 2 kQ sub-tiles of 4 x 4 pairs per tile against the same four column boxes, so the per-tile barrier, decode and column
// staging are spread over kQ times the pairs of the 64 x 64 shape (measured per 64 images of N = 4096: 1200 us with
// 64-row tiles, 1118 with 128, 1094 with 256).  Row block R meets the column blocks C >= kQ R.  With c = C - kQ R, a tile
// with c < kQ touches the diagonal: its 64-row quarters before quarter c are ordinary off-diagonal 64 x 64 tiles
// (mirrored), quarter c is a diagonal tile (not mirrored) and the quarters after it are skipped (they are the mirror
// images of quarters of the neighbouring tiles); every other tile is mirrored whole.
// kQ = 64-row quarters per row block (2: 128 x 64 tiles, 4: 256 x 64 tiles)
template <int kQ>
__device__ __forceinline__ void tall_decode(int t, int ntc, int& R, int& C) {
    // rows of the tile triangle have ntc - kQ R tiles: S(R) = R ntc - kQ R (R - 1) / 2 tiles come before row R
    const float bq = (float)ntc + 0.5f * kQ;
    int r = (int)((bq - sqrtf(fmaxf(bq * bq - 2.0f * kQ * (float)t, 0.f))) / (float)kQ);
    r = max(0, r);
    while (r > 0 && r * ntc - kQ * r * (r - 1) / 2 > t) --r;
    while ((r + 1) * ntc - kQ * (r + 1) * r / 2 <= t) ++r;
    R = r;
    C = kQ * r + (t - (r * ntc - kQ * r * (r - 1) / 2));
}
static inline int tall_tiles_per_image(int N, int kQ) {
    const int ntc = (N + 63) / 64, ntr = (N + 64 * kQ - 1) / (64 * kQ);
    int n = 0;
    for (int R = 0; R < ntr; ++R) n += ntc - kQ * R;
    return n;
}

template <int kSrc, bool kGen, bool kAffine, int kQ, bool kPacked = false>
__global__ void __launch_bounds__(128, 4) tile_tall_kernel(TileArgs A) {
    typedef typename RecOf<kSrc>::type RecT;
    constexpr int kNF = SoaOf<kSrc>::kFields;
    constexpr int kRows = 64 * kQ;
    __shared__ __align__(16) float s_row[2][kNF * kRows];            // [buffer][field][row box]
    __shared__ __align__(16) float s_col[2][kNF * kTT];             // [buffer][field][column box]
    __shared__ int s_ij[2][4];
    const int N = A.N, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx = (lane & 7) + 8 * (warp & 1), ty = (lane >> 3) + 4 * (warp >> 1);      // 16 column quads x 8 row quads
    const int total = A.tiles_per_image * A.batch;
    auto prefetch = [&](int t, int buf) {
        const int b = div_small(t, A.tiles_per_image, A.inv_tpi);
        int R, C;
        tall_decode<kQ>(t - b * A.tiles_per_image, A.nt, R, C);
        const float* bx = A.boxes + (size_t)b * N * kNF;
#pragma unroll
        for (int k = tid; k < kRows; k += 128) {                                 // kRows / 128 row records per thread
            const float* src = bx + (size_t)min(R * kRows + k, N - 1) * kNF;
            float* dst = &s_row[buf][k];
#pragma unroll
            for (int q = 0; q < kNF; ++q) cp_async4(dst + q * kRows, src + q);
        }
        if (tid < kTT) {                                                         // ... and one column record
            const float* src = bx + (size_t)min(C * kTT + tid, N - 1) * kNF;
            float* dst = &s_col[buf][tid];
#pragma unroll
            for (int q = 0; q < kNF; ++q) cp_async4(dst + q * kTT, src + q);
        }
        if (tid == 0) { s_ij[buf][0] = b; s_ij[buf][1] = R; s_ij[buf][2] = C; }
    };
    auto record_ok = [&](const float* f, int stride) -> bool {
        if constexpr (kSrc == kSrcBox3d)
            return rec3_sane(Rec3{f[0], f[stride], f[2 * stride], f[3 * stride], f[4 * stride], f[5 * stride], f[6 * stride], 0.f});
        else return box2_sane(make_box2(make_float4(f[0], f[stride], f[2 * stride], f[3 * stride])));
    };
    const bool chunked = A.tiles_per_cta > 0;
    const int t_step = chunked ? 1 : (int)gridDim.x;
    const int t_end = chunked ? min(total, ((int)blockIdx.x + 1) * A.tiles_per_cta) : total;
    int t = chunked ? (int)blockIdx.x * A.tiles_per_cta : (int)blockIdx.x, buf = 0;
    if (t < t_end) prefetch(t, 0);
    for (; t < t_end; t += t_step, buf ^= 1) {
        cp_async_wait_all();
        bool bad = false;
#pragma unroll
        for (int k = tid; k < kRows; k += 128) bad = bad || !record_ok(&s_row[buf][k], kRows);
        if (tid < kTT) bad = bad || !record_ok(&s_col[buf][tid], kTT);
        const bool tile_unsafe = __syncthreads_or(bad);
        const int b = s_ij[buf][0], R = s_ij[buf][1], C = s_ij[buf][2];
        if (t + t_step < t_end) prefetch(t + t_step, buf ^ 1);
        RecT cr[4];
        SoaOf<kSrc>::load4(s_col[buf], 4 * tx, cr);
        float* out = A.out + (size_t)b * N * N;
        const bool full_tile = A.vec && (R * kRows + kRows <= N) && (C * kTT + kTT <= N);
        const int c = C - kQ * R;                                     // < kQ: the tile touches the diagonal in quarter c
        const int h_end = c < kQ ? 2 * (c + 1) : 2 * kQ;              // quarters past the diagonal one are mirrors of other tiles
        const int j0 = C * kTT + 4 * tx;
        // store addresses are carried from one 32-row step to the next (byte pointers; N <= 8192, so row offsets fit 32 bits):
        // drow walks down 32 rows of the direct block, dcol walks 32 columns to the right in the mirrored block
        const uint32_t row_bytes = (uint32_t)N * 4u;
        char* drow = reinterpret_cast<char*>(out + (int64_t)(R * kRows + 4 * ty) * N + j0);
        char* dcol = reinterpret_cast<char*>(out + (int64_t)j0 * N + (R * kRows + 4 * ty));
        auto load_rows = [&](int rl, RecT* rr) {
            float4 f[7];
            constexpr int kUse = kSrc == kSrcBox3d ? 7 : 4;
#pragma unroll
            for (int q = 0; q < kUse; ++q) f[q] = *reinterpret_cast<const float4*>(&s_row[buf][q * kRows + rl]);
            const float* pf = reinterpret_cast<const float*>(f);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if constexpr (kSrc == kSrcBox3d) rr[k] = Rec3{pf[k], pf[4 + k], pf[8 + k], pf[12 + k], pf[16 + k], pf[20 + k], pf[24 + k], 0.f};
                else rr[k] = make_box2(make_float4(pf[k], pf[4 + k], pf[8 + k], pf[12 + k]));
            }
        };
        RecT rr[4];
#pragma unroll 1
        for (int h = 0; h < h_end; ++h) {
            const int rl = 32 * h + 4 * ty;
            if (R * kRows + 32 * h >= N) break;
            const bool mirror = (h >> 1) != c;                        // the diagonal 64 x 64 quarter is not mirrored
            load_rows(rl, rr);
            float v[4][4];
            bool unsafe = tile_unsafe;
            if constexpr (kPacked && kSrc == kSrcBox3d) {             // two column boxes per instruction (fp32x2)
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int k = 0; k < 4; k += 2) {
                        const F2 pv = iou3_fast2<kGen, kAffine>(rr[r], cr[k], cr[k + 1], unsafe);
                        v[r][k] = pv.x; v[r][k + 1] = pv.y;
                    }
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[r][k] = RecOf<kSrc>::template fast<kGen, kAffine>(rr[r], cr[k], unsafe);
            }
            if (__builtin_expect(unsafe, 0)) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[r][k] = RecOf<kSrc>::template exact<kGen, kAffine>(rr[r], cr[k]);
            }
            if (full_tile) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    st_cs_f4(reinterpret_cast<float*>(drow + r * row_bytes), make_float4(v[r][0], v[r][1], v[r][2], v[r][3]));
                if (mirror) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        st_cs_f4(reinterpret_cast<float*>(dcol + k * row_bytes), make_float4(v[0][k], v[1][k], v[2][k], v[3][k]));
                }
            } else {
                const int i0 = R * kRows + rl;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (i0 + r < N && j0 + k < N) {
                            out[(int64_t)(i0 + r) * N + (j0 + k)] = v[r][k];
                            if (mirror) out[(int64_t)(j0 + k) * N + (i0 + r)] = v[r][k];
                        }
                    }
            }
            drow += 32u * (size_t)row_bytes;
            dcol += 128;
        }
    }
}

// ------------------------------------------------------------------------------------------ 2d. matrix only, TMA tensor stores
// Same 256 x 64 symmetric tiles, same arithmetic, same bits as tile_tall_kernel; what changes is how the results leave the
// SM.  tile_tall_kernel stores every 4 x 4 sub-tile straight from registers (8 STG.128 + the address arithmetic they need)
// and then stalls on the registers the queued stores still read (ncu: all of its long_scoreboard time).  Here a warp owns
// a 32-row x 32-column block of the tile per step: its threads drop their 4 x 4 results into two shared-memory boxes of
// 32 rows x 128 bytes -- the block itself and its transpose (the mirror image across the matrix diagonal), both in the
// 128-byte-swizzled layout of the tensor map so that the row-wise and the transposed 16-byte writes spread over the banks
// -- and one lane hands both boxes to the copy engine (cp.async.bulk.tensor.3d ... shared::cta -> global, SASS UTMASTG).
// Registers are free right after the STS, the stores drain asynchronously while the warp evaluates its next block, there
// is no per-store address arithmetic (shared offsets are per-thread constants), partial tiles are clipped by the tensor
// map ([batch, N, N], so a tall tile never runs into the next image), and no barrier is added: a warp waits only for its
// own previous bulk store to have read its staging area (cp.async.bulk.wait_group.read), one 32 x 32 block later.
// Warp w of the 128-thread CTA takes column half w & 1 of the tile and the 32-row steps of parity w >> 1.
// Lane -> (tx, ty) = 4-column quad x 4-row quad inside the 32 x 32 block, in two 16-row sub-steps; lane bits (t0 t1 y0 t2 y1)
// make the direct writes conflict-free (a quarter warp covers 4 quads of 2 rows: 8 distinct swizzled chunks) and the
// transposed ones 2-way (no lane order makes both conflict-free: each needs its own third bit next to t0 and y0).
constexpr int kTmaStageBytes = 4 * 2 * 4096;                 // 4 warps x (direct box + mirror box) x 32 rows x 128 B

__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t smem_addr, int x, int y, int z, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;"
                 :: "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_addr), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int kSrc, bool kGen, bool kAffine, bool kPacked>
__global__ void __launch_bounds__(128, 4) tile_tma_kernel(TileArgs A, const __grid_constant__ CUtensorMap tmap) {
    typedef typename RecOf<kSrc>::type RecT;
    constexpr int kNF = SoaOf<kSrc>::kFields;
    constexpr int kQ = 4, kRows = 64 * kQ;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(smem_dyn) + 1023u) & ~1023u;        // swizzle atoms are 1 KB
    unsigned char* gbase = smem_dyn + (sbase - (uint32_t)__cvta_generic_to_shared(smem_dyn));
    float* s_row = reinterpret_cast<float*>(gbase + kTmaStageBytes);                                // [2][kNF * kRows]
    float* s_col = s_row + 2 * kNF * kRows;                                                         // [2][kNF * kTT]
    int* s_ij = reinterpret_cast<int*>(s_col + 2 * kNF * kTT);                                      // [2][4]
    const int N = A.N, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int half = warp & 1, hpar = warp >> 1;
    const int tx = (lane & 3) | ((lane >> 1) & 4), ty = ((lane >> 2) & 1) | ((lane >> 3) & 2);
    const int total = A.tiles_per_image * A.batch;
    // per-thread shared offsets (bytes) inside the warp's two boxes; sub-step 1 adds 16 rows (direct) / flips chunk bit 2 (mirror)
    const uint32_t dbox = sbase + (uint32_t)warp * 8192u, mbox = dbox + 4096u;
    uint32_t doff[4], moff[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = 4 * ty + r;                                        // + 16 * sub: (row & 7) is the same in both sub-steps
        doff[r] = dbox + (uint32_t)(row * 128 + ((tx ^ (row & 7)) << 4));
        const int mrow = 4 * tx + r;                                       // r plays k here: column k of the thread = mirror row
        moff[r] = mbox + (uint32_t)(mrow * 128 + ((ty ^ (mrow & 7)) << 4));
    }
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    auto prefetch = [&](int t, int buf) {
        const int b = div_small(t, A.tiles_per_image, A.inv_tpi);
        int R, C;
        tall_decode<kQ>(t - b * A.tiles_per_image, A.nt, R, C);
        const float* bx = A.boxes + (size_t)b * N * kNF;
#pragma unroll
        for (int k = tid; k < kRows; k += 128) {
            const float* src = bx + (size_t)min(R * kRows + k, N - 1) * kNF;
            float* dst = s_row + buf * kNF * kRows + k;
#pragma unroll
            for (int q = 0; q < kNF; ++q) cp_async4(dst + q * kRows, src + q);
        }
        if (tid < kTT) {
            const float* src = bx + (size_t)min(C * kTT + tid, N - 1) * kNF;
            float* dst = s_col + buf * kNF * kTT + tid;
#pragma unroll
            for (int q = 0; q < kNF; ++q) cp_async4(dst + q * kTT, src + q);
        }
        if (tid == 0) { s_ij[buf * 4 + 0] = b; s_ij[buf * 4 + 1] = R; s_ij[buf * 4 + 2] = C; }
    };
    auto record_ok = [&](const float* f, int stride) -> bool {
        if constexpr (kSrc == kSrcBox3d)
            return rec3_sane(Rec3{f[0], f[stride], f[2 * stride], f[3 * stride], f[4 * stride], f[5 * stride], f[6 * stride], 0.f});
        else return box2_sane(make_box2(make_float4(f[0], f[stride], f[2 * stride], f[3 * stride])));
    };
    const bool chunked = A.tiles_per_cta > 0;
    const int t_step = chunked ? 1 : (int)gridDim.x;
    const int t_end = chunked ? min(total, ((int)blockIdx.x + 1) * A.tiles_per_cta) : total;
    int t = chunked ? (int)blockIdx.x * A.tiles_per_cta : (int)blockIdx.x, buf = 0;
    if (t < t_end) prefetch(t, 0);
    for (; t < t_end; t += t_step, buf ^= 1) {
        cp_async_wait_all();
        const float* srow = s_row + buf * kNF * kRows;
        const float* scol = s_col + buf * kNF * kTT;
        bool bad = false;
#pragma unroll
        for (int k = tid; k < kRows; k += 128) bad = bad || !record_ok(srow + k, kRows);
        if (tid < kTT) bad = bad || !record_ok(scol + tid, kTT);
        const bool tile_unsafe = __syncthreads_or(bad);
        const int b = s_ij[buf * 4 + 0], R = s_ij[buf * 4 + 1], C = s_ij[buf * 4 + 2];
        if (t + t_step < t_end) prefetch(t + t_step, buf ^ 1);
        RecT cr[4];
        SoaOf<kSrc>::load4(scol, 32 * half + 4 * tx, cr);
        const int c = C - kQ * R;                                     // < kQ: the tile touches the diagonal in quarter c
        const int h_end = c < kQ ? 2 * (c + 1) : 2 * kQ;              // quarters past the diagonal one are mirrors of other tiles
        const int col0 = C * kTT + 32 * half;                         // first column of this warp's blocks
#pragma unroll 1
        for (int h = hpar; h < h_end; h += 2) {
            const int row0 = R * kRows + 32 * h;                      // first row of this block
            if (row0 >= N) break;
            const bool mirror = (h >> 1) != c;                        // the diagonal 64 x 64 quarter is not mirrored
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                const int rl = 32 * h + 16 * sub + 4 * ty;
                RecT rr[4];
                {
                    float4 f[7];
                    constexpr int kUse = kSrc == kSrcBox3d ? 7 : 4;
#pragma unroll
                    for (int q = 0; q < kUse; ++q) f[q] = *reinterpret_cast<const float4*>(srow + q * kRows + rl);
                    const float* pf = reinterpret_cast<const float*>(f);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if constexpr (kSrc == kSrcBox3d) rr[k] = Rec3{pf[k], pf[4 + k], pf[8 + k], pf[12 + k], pf[16 + k], pf[20 + k], pf[24 + k], 0.f};
                        else rr[k] = make_box2(make_float4(pf[k], pf[4 + k], pf[8 + k], pf[12 + k]));
                    }
                }
                float v[4][4];
                bool unsafe = tile_unsafe;
                if constexpr (kPacked && kSrc == kSrcBox3d) {         // two column boxes per instruction (fp32x2)
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int k = 0; k < 4; k += 2) {
                            const F2 pv = iou3_fast2<kGen, kAffine>(rr[r], cr[k], cr[k + 1], unsafe);
                            v[r][k] = pv.x; v[r][k + 1] = pv.y;
                        }
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int k = 0; k < 4; ++k) v[r][k] = RecOf<kSrc>::template fast<kGen, kAffine>(rr[r], cr[k], unsafe);
                }
                if (__builtin_expect(unsafe, 0)) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int k = 0; k < 4; ++k) v[r][k] = RecOf<kSrc>::template exact<kGen, kAffine>(rr[r], cr[k]);
                }
                if (sub == 0) {                                       // the previous block's bulk stores must have read the boxes
                    if (lane == 0) bulk_wait_read0();
                    __syncwarp();
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) sts_f4(doff[r] + (uint32_t)sub * 2048u, v[r][0], v[r][1], v[r][2], v[r][3]);
                if (mirror) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) sts_f4(moff[k] ^ ((uint32_t)sub << 6), v[0][k], v[1][k], v[2][k], v[3][k]);
                }
            }
            fence_async_smem();                                       // generic-proxy writes -> visible to the copy engine
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&tmap, dbox, col0, row0, b, policy);
                if (mirror) tma_store_3d(&tmap, mbox, row0, col0, b, policy);
                bulk_commit();
            }
        }
    }
    if (lane == 0) bulk_wait_read0();                                 // the staging area must outlive the copy engine's reads
}

// ------------------------------------------------------------------------------------------ 2d. spatial order + tile culling
// Matrix-free path only.  Suppression needs a pair's overlap only when it can exceed the threshold, and boxes that
// are far apart cannot: if two boxes are disjoint along an axis their intersection is EXACTLY 0 in fp32, so
//   2D IoU / 3D IoU:            v = 0           -> never a hit for thr >= 0
//   GIoU (v = io - h):          v = -h <= 0     -> never a hit for thr >= 0
//   0.5 * (1 + GIoU):           v = 0.5 (1 - h), h = (hull - V) / hull >= gap / (w_a + w_b + gap) along the disjoint
//                               axis, so v <= thr as soon as gap >= c (w_a + w_b), c = (1 - 2 thr + eps) / (2 thr - eps)
// (eps = 1e-4 swallows the fp32 rounding of hull, V and the division, which is < 1e-6 relative).  One CTA per image
// buckets the boxes by a 12-bit Morton code of their centre (counting sort in shared memory), writes the blocked
// structure-of-arrays copy of the records and ranks in that spatial order, reduces each run of 64 boxes to an AABB + max extents, and lists the
// tile pairs whose groups are NOT provably out of reach.  Groups holding a degenerate / non-finite box are never
// culled (0/0 = NaN counts as a hit, lib/groomed_nms.py:249-250).  On clustered boxes ~90 % of the tiles disappear.
// gt: [nt][10] group table (lo/hi per axis, max extents, bad flag); appends the tile pairs that are not provably out of reach
__device__ __forceinline__ void list_tiles(const float* gt, int nt, int naxes, float c, int b, int32_t* tile_list, int tid, int nthreads) {
    const int tpi = nt * (nt + 1) / 2;
    for (int t = tid; t < tpi; t += nthreads) {
        int I, J;
        tile_decode(t, nt, I, J);
        bool cull = false;
        if (I != J && gt[I * 10 + 9] == 0.f && gt[J * 10 + 9] == 0.f) {
            for (int ax = 0; ax < naxes; ++ax) {
                const float gap = fmaxf(gt[J * 10 + 2 * ax] - gt[I * 10 + 2 * ax + 1], gt[I * 10 + 2 * ax] - gt[J * 10 + 2 * ax + 1]);
                // empty groups have lo = +inf, hi = -inf: gap = +inf -> culled
                if (gap > 0.f && gap > c * (gt[I * 10 + 6 + ax] + gt[J * 10 + 6 + ax])) cull = true;
            }
        }
        if (!cull) tile_list[64 + atomicAdd(&tile_list[0], 1)] = (b << 16) | (I << 8) | J;
    }
}

struct SpatialArgs {
    int N, batch, nt, src;
    const int32_t* n_per_image;
    const float* boxes;          // [batch, N, 8] records or [batch, N, 4] boxes (input order)
    char* ws;
    size_t ws_img_stride;
    int32_t* tile_list;          // [0] = count (zeroed by the rank kernel), [64..] = entries
    float cull_c;                // gap factor c (0: any positive gap)
    int defer_list;              // direct-election flow: store the group table, list_failed_kernel lists tiles later
};

__global__ void __launch_bounds__(1024) spatial_kernel(SpatialArgs A) {
    __shared__ int s_hist[4096];
    __shared__ float s_red[4][32];
    __shared__ float s_gt[128][10];      // per group: xlo,xhi,ylo,yhi,zlo,zhi, wx,wy,wz (max extents), bad
    __shared__ int s_wsum[32];
    __shared__ uint16_t s_perm[GNMS_MAX_BOXES];                      // spatial position -> input index
    const int b = blockIdx.x, N = A.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = A.n_per_image ? min(A.n_per_image[b], N) : N;
    const WsLayout L = ws_layout(N);
    char* w = A.ws + (size_t)b * A.ws_img_stride;
    const int32_t* rank = reinterpret_cast<const int32_t*>(w + L.rank);
    float* blk = reinterpret_cast<float*>(w + L.blk);               // blocked SoA copy, here in SPATIAL order
    const bool is3d = A.src == kSrcBox3d;
    const int recf = is3d ? 8 : 4;
    const float* bx = A.boxes + (size_t)b * N * recf;
    constexpr int kPer = GNMS_MAX_BOXES / 1024;
    // centre of box i along the two bucketing axes (3D: BEV x,z ; 2D: x,y)
    auto centre = [&](int i, float& u, float& v) {
        if (is3d) {
            const float4 a = *reinterpret_cast<const float4*>(bx + (size_t)i * 8);
            const float2 c = *reinterpret_cast<const float2*>(bx + (size_t)i * 8 + 4);
            u = 0.5f * (a.z + a.w); v = 0.5f * (c.x + c.y);
        } else {
            const float4 a = *reinterpret_cast<const float4*>(bx + (size_t)i * 4);
            u = 0.5f * (a.x + a.z); v = 0.5f * (a.y + a.w);
        }
    };
    float cu[kPer], cv[kPer];
    float ulo = INFINITY, uhi = -INFINITY, vlo = INFINITY, vhi = -INFINITY;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int i = tid + q * 1024;
        cu[q] = 0.f; cv[q] = 0.f;
        if (i < n) {
            centre(i, cu[q], cv[q]);
            if (isfinite(cu[q]) && isfinite(cv[q])) {
                ulo = fminf(ulo, cu[q]); uhi = fmaxf(uhi, cu[q]); vlo = fminf(vlo, cv[q]); vhi = fmaxf(vhi, cv[q]);
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ulo = fminf(ulo, __shfl_xor_sync(0xffffffffu, ulo, d)); uhi = fmaxf(uhi, __shfl_xor_sync(0xffffffffu, uhi, d));
        vlo = fminf(vlo, __shfl_xor_sync(0xffffffffu, vlo, d)); vhi = fmaxf(vhi, __shfl_xor_sync(0xffffffffu, vhi, d));
    }
    if (lane == 0) { s_red[0][warp] = ulo; s_red[1][warp] = uhi; s_red[2][warp] = vlo; s_red[3][warp] = vhi; }
    for (int i = tid; i < 4096; i += 1024) s_hist[i] = 0;
    __syncthreads();
    ulo = s_red[0][lane]; uhi = s_red[1][lane]; vlo = s_red[2][lane]; vhi = s_red[3][lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ulo = fminf(ulo, __shfl_xor_sync(0xffffffffu, ulo, d)); uhi = fmaxf(uhi, __shfl_xor_sync(0xffffffffu, uhi, d));
        vlo = fminf(vlo, __shfl_xor_sync(0xffffffffu, vlo, d)); vhi = fmaxf(vhi, __shfl_xor_sync(0xffffffffu, vhi, d));
    }
    const float su = (uhi > ulo) ? 64.0f / (uhi - ulo) : 0.f, sv = (vhi > vlo) ? 64.0f / (vhi - vlo) : 0.f;
    int bucket[kPer], off[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int i = tid + q * 1024;
        bucket[q] = 4095; off[q] = 0;
        if (i < N) {
            if (i < n && isfinite(cu[q]) && isfinite(cv[q])) {
                const uint32_t qu = (uint32_t)min(63, max(0, (int)((cu[q] - ulo) * su)));
                const uint32_t qv = (uint32_t)min(63, max(0, (int)((cv[q] - vlo) * sv)));
                uint32_t m = 0;                                        // 12-bit Morton code
#pragma unroll
                for (int bit = 0; bit < 6; ++bit) m |= ((qu >> bit) & 1u) << (2 * bit) | ((qv >> bit) & 1u) << (2 * bit + 1);
                bucket[q] = (int)m;
            }
            off[q] = atomicAdd(&s_hist[bucket[q]], 1);
        }
    }
    __syncthreads();
    // exclusive scan of the 4096 bucket counts (4 per thread)
    {
        int loc[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { loc[q] = s_hist[tid * 4 + q]; sum += loc[q]; }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = s_wsum[lane], iv = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, iv, d);
                if (lane >= d) iv += t;
            }
            s_wsum[lane] = iv - v;
        }
        __syncthreads();
        int run = s_wsum[warp] + incl - sum;
#pragma unroll
        for (int q = 0; q < 4; ++q) { s_hist[tid * 4 + q] = run; run += loc[q]; }
    }
    __syncthreads();
    // spatial order: position -> input index in shared memory first, so that the blocked copy can be written with
    // 16-byte stores (a thread fills 4 consecutive slots of every field array)
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int i = tid + q * 1024;
        if (i < N) s_perm[s_hist[bucket[q]] + off[q]] = (uint16_t)i;
    }
    __syncthreads();
    for (int p4 = tid * 4; p4 < ((N + 63) & ~63); p4 += 4096) {
        float fld[8][4];
        int rk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int pos = p4 + e;
            if (pos < N) {
                const int i = s_perm[pos];
                if (is3d) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(bx + (size_t)i * 8));
                    const float4 c = __ldg(reinterpret_cast<const float4*>(bx + (size_t)i * 8) + 1);
                    fld[0][e] = a.x; fld[1][e] = a.y; fld[2][e] = a.z; fld[3][e] = a.w;
                    fld[4][e] = c.x; fld[5][e] = c.y; fld[6][e] = c.z; fld[7][e] = c.w;
                } else {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(bx) + i);
                    fld[0][e] = a.x; fld[1][e] = a.y; fld[2][e] = a.z; fld[3][e] = a.w;
                    fld[4][e] = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));      // area, for the range check
                    fld[5][e] = 0.f; fld[6][e] = 0.f; fld[7][e] = 0.f;
                }
                rk[e] = i < n ? rank[i] : INT_MAX;
            } else {                                                   // padding slots of the last block: a unit box
#pragma unroll
                for (int f = 0; f < 8; ++f) fld[f][e] = is3d ? (((f & 1) || f >= 6) ? 1.f : 0.f) : ((f >= 2 && f <= 4) ? 1.f : 0.f);
                rk[e] = INT_MAX;
            }
        }
        float* bg = blk + (size_t)(p4 >> 6) * kBlkWords + (p4 & 63);
        const int nf = is3d ? 8 : 5;
#pragma unroll
        for (int f = 0; f < 8; ++f)
            if (f < nf) *reinterpret_cast<float4*>(bg + f * 64) = make_float4(fld[f][0], fld[f][1], fld[f][2], fld[f][3]);
        *reinterpret_cast<int4*>(bg + 8 * 64) = make_int4(rk[0], rk[1], rk[2], rk[3]);
    }
    __threadfence_block();
    __syncthreads();
    // AABB + max extents of every run of 64 boxes (padded boxes, rank INT_MAX, do not count)
    const int nt = A.nt;
    for (int g = warp; g < nt; g += 32) {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, ext[3] = {0.f, 0.f, 0.f};
        bool bad = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pos = g * 64 + lane + 32 * h;
            const float* bg = blk + (size_t)g * kBlkWords + (lane + 32 * h);      // element of block g: field f at bg[f * 64]
            if (pos < N && reinterpret_cast<const int32_t*>(bg)[8 * 64] != INT_MAX) {
                float a[3][2];
                if (is3d) {
                    const Rec3 r = {bg[0], bg[64], bg[128], bg[192], bg[256], bg[320], bg[384], bg[448]};
                    a[0][0] = r.bx1; a[0][1] = r.bx2; a[1][0] = r.ymin; a[1][1] = r.ymax; a[2][0] = r.bz1; a[2][1] = r.bz2;
                    // the gap bound assumes vol = (BEV x extent)(y extent)(BEV z extent), true for y-rotated cuboids
                    // (lib/math_3d.py:364-435); records built from other corner sets are simply never culled
                    bad = bad || !rec3_sane(r) ||
                          r.vol != __fmul_rn(__fmul_rn(__fsub_rn(r.bx2, r.bx1), __fsub_rn(r.ymax, r.ymin)), __fsub_rn(r.bz2, r.bz1));
                } else {
                    const float4 q4 = make_float4(bg[0], bg[64], bg[128], bg[192]);
                    a[0][0] = q4.x; a[0][1] = q4.z; a[1][0] = q4.y; a[1][1] = q4.w; a[2][0] = 0.f; a[2][1] = 0.f;
                    bad = bad || !box2_sane(make_box2(q4)) || !(fabsf(q4.x) <= 1e18f && fabsf(q4.y) <= 1e18f && fabsf(q4.z) <= 1e18f && fabsf(q4.w) <= 1e18f);
                }
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    lo[ax] = fminf(lo[ax], a[ax][0]); hi[ax] = fmaxf(hi[ax], a[ax][1]);
                    ext[ax] = fmaxf(ext[ax], a[ax][1] - a[ax][0]);
                    bad = bad || !(a[ax][1] >= a[ax][0]);
                }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                lo[ax] = fminf(lo[ax], __shfl_xor_sync(0xffffffffu, lo[ax], d));
                hi[ax] = fmaxf(hi[ax], __shfl_xor_sync(0xffffffffu, hi[ax], d));
                ext[ax] = fmaxf(ext[ax], __shfl_xor_sync(0xffffffffu, ext[ax], d));
            }
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
            s_gt[g][0] = lo[0]; s_gt[g][1] = hi[0]; s_gt[g][2] = lo[1]; s_gt[g][3] = hi[1]; s_gt[g][4] = lo[2]; s_gt[g][5] = hi[2];
            s_gt[g][6] = ext[0]; s_gt[g][7] = ext[1]; s_gt[g][8] = ext[2]; s_gt[g][9] = bad ? 1.f : 0.f;
        }
    }
    __syncthreads();
    if (A.defer_list) {                                               // direct-election flow: keep the table, list later if needed
        float* gtab = reinterpret_cast<float*>(w + L.gtab);
        for (int i = tid; i < nt * 10; i += 1024) gtab[i] = s_gt[i / 10][i % 10];
        return;
    }
    list_tiles(&s_gt[0][0], nt, is3d ? 3 : 2, A.cull_c, b, A.tile_list, tid, 1024);
}

// tile pairs of the images elect_kernel gave up on (grid batch x 256)
__global__ void __launch_bounds__(256) list_failed_kernel(SpatialArgs A) {
    const int b = blockIdx.x;
    const WsLayout L = ws_layout(A.N);
    char* w = A.ws + (size_t)b * A.ws_img_stride;
    if (*reinterpret_cast<const int32_t*>(w + L.dflag) == 0) return;
    list_tiles(reinterpret_cast<const float*>(w + L.gtab), A.nt, A.src == kSrcBox3d ? 3 : 2, A.cull_c, b, A.tile_list, threadIdx.x, 256);
}

// has-earlier partials from the finished mask (column-major): CTA = 256 columns; work item = (row word, 32-column
// part): 32 independent coalesced loads per thread, parts combined in shared memory   (grid (ceil(N/256), batch) x 1024)
__device__ __forceinline__ void has_earlier_item(int N, int n, int chunk, const uint32_t* __restrict__ mask, uint32_t* __restrict__ has_earlier,
                                                 int he_slots, uint32_t* s_acc) {
    const int NW = (N + 31) / 32;
    const int l0 = chunk * 256;
    for (int jw = threadIdx.x; jw < NW; jw += 1024) s_acc[jw] = 0u;
    __syncthreads();
    for (int item = threadIdx.x; item < NW * 8; item += 1024) {
        const int jw = item % NW, part = item / NW;
        const int la = l0 + part * 32;
        uint32_t acc = 0u;
        if (jw * 32 < n && la < n && (la >> 5) <= jw) {        // a column only has bits in row words at or after its own
#pragma unroll
            for (int q = 0; q < 32; ++q) acc |= (la + q < n) ? mask[(size_t)(la + q) * NW + jw] : 0u;
        }
        if (acc) atomicOr(&s_acc[jw], acc);
    }
    __syncthreads();
    for (int jw = threadIdx.x; jw < NW; jw += 1024) has_earlier[(size_t)jw * he_slots + chunk] = s_acc[jw];
}

__global__ void __launch_bounds__(1024) has_earlier_kernel(int N, const int32_t* __restrict__ n_per_image, char* __restrict__ ws,
                                                           size_t ws_img_stride) {
    __shared__ uint32_t s_acc[GNMS_MAX_BOXES / 32];
    const int b = blockIdx.y;
    const int n = n_per_image ? min(n_per_image[b], N) : N;
    const WsLayout L = ws_layout(N);
    char* w = ws + (size_t)b * ws_img_stride;
    has_earlier_item(N, n, blockIdx.x, reinterpret_cast<const uint32_t*>(w + L.mask), reinterpret_cast<uint32_t*>(w + L.has_earlier),
                     L.he_slots, s_acc);
}

// Direct-election flow: only the images elect_kernel gave up on need this; the grid is small and loops over
// (image, 256-column chunk) items, so that a launch in which every image was elected costs next to nothing.
__global__ void __launch_bounds__(1024) has_earlier_failed_kernel(int N, int batch, const int32_t* __restrict__ n_per_image,
                                                                  char* __restrict__ ws, size_t ws_img_stride) {
    __shared__ uint32_t s_acc[GNMS_MAX_BOXES / 32];
    const int nchunks = (N + 255) / 256;
    const WsLayout L = ws_layout(N);
    for (int item0 = blockIdx.x; item0 < batch * nchunks; item0 += gridDim.x) {
        const int b = item0 / nchunks, chunk = item0 - b * nchunks;
        char* w = ws + (size_t)b * ws_img_stride;
        if (*reinterpret_cast<const int32_t*>(w + L.dflag) == 0) continue;
        __syncthreads();                                               // s_acc of the previous item has been written out
        has_earlier_item(N, n_per_image ? min(n_per_image[b], N) : N, chunk, reinterpret_cast<const uint32_t*>(w + L.mask),
                         reinterpret_cast<uint32_t*>(w + L.has_earlier), L.he_slots, s_acc);
    }
}

// ------------------------------------------------------------------------------------------ 2e. direct leader election
// Matrix-free path.  The greedy election only ever needs the overlap COLUMNS of the leaders (a few dozen per image on
// detector outputs), not all N^2 pairs: this kernel runs the reference's loop (lib/groomed_nms.py:247-262) as it is
// written -- pick the best box still in the pool, evaluate its overlap with the boxes still in the pool, take those with
// !(v <= thr) out (their first suppressor is this leader) -- with one CTA per image, all in shared memory:
//   * the image's blocked copy in SPATIAL order (blk, written by the spatial kernel just before: records + ranks) is
//     copied into shared memory once; the pool is a bitset over sorted positions that every warp scans redundantly, so a
//     step costs ONE barrier;
//   * per step every warp tests all 64-box groups against the leader with the conservative gap bound of the tile
//     culling (group table of the spatial kernel; a lane owns two groups) and gets the same 64-bit "near" mask; only
//     the boxes of the near groups are evaluated, one pair per thread, spread over the whole CTA -- so the critical
//     path of a step is one pair evaluation, not a warp's worth of them;
//   * a box that leaves the pool gets its first suppressor written straight to global memory (once per box).
// ~0.5 us per leader; images run side by side on different SMs.  If an image needs more than kElectMaxLeaders steps
// (boxes that barely overlap) the kernel gives up for that image and sets its flag: the list / tile / has-earlier
// kernels then do the work for it (they skip the images whose flag is clear) and chain_kernel takes the mask route.
constexpr int kElectMaxBoxes = 4096;
constexpr int kElectMaxLeaders = 384;
struct ElectArgs {
    int N, batch;
    const int32_t* n_per_image;
    char* ws;
    size_t ws_img_stride;
    float thr, cull_c;
};
static size_t elect_smem_bytes(int N) {
    const size_t nt = (size_t)((N + 63) / 64);
    return nt * kBlkBytes + nt * 64 * 2 + 3 * 128 * 4 + 64;           // blocks, rank -> slot map, alive / leader / bad bitsets
}

constexpr int kElectThreads = 256;

template <int kSrc, bool kGen, bool kAffine>
__global__ void __launch_bounds__(kElectThreads) elect_kernel(ElectArgs A) {
    typedef typename RecOf<kSrc>::type RecT;
    constexpr int kAxes = kSrc == kSrcBox3d ? 3 : 2;
    constexpr int kWarps = kElectThreads / 32;
    extern __shared__ __align__(16) unsigned char s_el[];
    const int b = blockIdx.x, N = A.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = A.n_per_image ? min(A.n_per_image[b], N) : N;
    const int nw = (n + 31) / 32, nt = (N + 63) / 64;
    const WsLayout L = ws_layout(N);
    char* w = A.ws + (size_t)b * A.ws_img_stride;
    const float* gtab = reinterpret_cast<const float*>(w + L.gtab);
    int32_t* dfsup = reinterpret_cast<int32_t*>(w + L.dfsup);
    float* s_blk = reinterpret_cast<float*>(s_el);                           // [group][8 fields + rank][64], spatial order
    uint16_t* s_slot = reinterpret_cast<uint16_t*>(s_blk + (size_t)nt * kBlkWords);   // sorted position -> spatial slot
    uint32_t* alive = reinterpret_cast<uint32_t*>(s_slot + (size_t)nt * 64); // boxes still in the pool, by sorted position
    uint32_t* leaderb = alive + 128;
    uint32_t* badb = leaderb + 128;                                          // by slot: record outside div_rn_fast's proven range
    {
        const uint4* src = reinterpret_cast<const uint4*>(w + L.blk);
        uint4* dst = reinterpret_cast<uint4*>(s_blk);
        for (int i = tid; i < nt * (kBlkBytes / 16); i += kElectThreads) dst[i] = src[i];
    }
    if (tid < 128) { alive[tid] = tid < nw ? valid_word(tid, n) : 0u; leaderb[tid] = 0u; badb[tid] = 0u; }
    for (int pos = tid; pos < n; pos += kElectThreads) dfsup[pos] = INT_MAX;
    // work items of a step: the 2 nt half-groups (32 spatial slots each); warp w owns items w, w + 8, w + 16, ... and
    // lane j keeps the reach of the group of item w + 8 j (lo / hi per axis, largest extents; a group holding an odd
    // box is never skipped).  Neighbouring items belong to different warps, so the few groups near a leader spread out.
    const int my_item = warp + kWarps * lane;
    const int my_group = my_item >> 1;
    float glo[kAxes], ghi[kAxes], gext[kAxes];
    bool gok = false;
    const bool item_valid = lane < 16 && my_group < nt;                      // 2 * 64 / 8 = 16 items per warp at most
#pragma unroll
    for (int ax = 0; ax < kAxes; ++ax) { glo[ax] = 0.f; ghi[ax] = 0.f; gext[ax] = 0.f; }
    if (item_valid) {
#pragma unroll
        for (int ax = 0; ax < kAxes; ++ax) { glo[ax] = gtab[my_group * 10 + 2 * ax]; ghi[ax] = gtab[my_group * 10 + 2 * ax + 1]; gext[ax] = gtab[my_group * 10 + 6 + ax]; }
        gok = A.cull_c >= 0.f && gtab[my_group * 10 + 9] == 0.f;
    }
    __shared__ uint32_t s_gbad[4];                                           // groups that may not be skipped
    if (tid < 4) s_gbad[tid] = 0u;
    __syncthreads();
    if (tid < nt && !(A.cull_c >= 0.f && gtab[tid * 10 + 9] == 0.f)) atomicOr(&s_gbad[tid >> 5], 1u << (tid & 31));
    for (int slot = tid; slot < nt * 64; slot += kElectThreads) {
        const float* bg = s_blk + (size_t)(slot >> 6) * kBlkWords + (slot & 63);
        const int r = reinterpret_cast<const int32_t*>(bg)[8 * 64];
        if (r == INT_MAX) continue;                                          // dead slot (padding / past the live count)
        s_slot[r] = (uint16_t)slot;
        bool bad;
        if constexpr (kSrc == kSrcBox3d) bad = !rec3_sane(Rec3{bg[0], bg[64], bg[128], bg[192], bg[256], bg[320], bg[384], 0.f});
        else bad = !box2_sane(make_box2(make_float4(bg[0], bg[64], bg[128], bg[192])));
        if (bad) atomicOr(&badb[slot >> 5], 1u << (slot & 31));
    }
    __syncthreads();
    auto record = [&](int slot) -> RecT {
        const float* bg = s_blk + (size_t)(slot >> 6) * kBlkWords + (slot & 63);
        if constexpr (kSrc == kSrcBox3d) return Rec3{bg[0], bg[64], bg[128], bg[192], bg[256], bg[320], bg[384], 0.f};
        else return make_box2(make_float4(bg[0], bg[64], bg[128], bg[192]));
    };
    int steps = 0;
    bool failed = false;
    while (true) {
        // first box still in the pool: every warp scans the bitset itself (no barrier, same answer everywhere)
        int l = -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t wv = alive[u * 32 + lane];
            const unsigned any = __ballot_sync(0xffffffffu, wv != 0u);
            if (any) {
                const int fl = __ffs(any) - 1;
                const uint32_t bits = __shfl_sync(0xffffffffu, wv, fl);
                l = (u * 32 + fl) * 32 + __ffs(bits) - 1;
                break;
            }
        }
        if (l < 0) break;
        if (++steps > kElectMaxLeaders) { failed = true; break; }
        const int lslot = s_slot[l];
        const RecT ldr = record(lslot);
        const bool lbad = (badb[lslot >> 5] >> (lslot & 31)) & 1u;
        const bool lok = !lbad && !((s_gbad[lslot >> 11] >> ((lslot >> 6) & 31)) & 1u);
        // which of my warp's items are within the leader's reach (gap along an axis > c * (leader extent + largest
        // extent in the group) => every overlap <= thr, see the tile culling; exact zeros when c == 0)
        bool is_near = item_valid;
        if (is_near && lok && gok) {
            float llo[3], lhi[3];
            if constexpr (kSrc == kSrcBox3d) { llo[0] = ldr.bx1; lhi[0] = ldr.bx2; llo[1] = ldr.ymin; lhi[1] = ldr.ymax; llo[2] = ldr.bz1; lhi[2] = ldr.bz2; }
            else { llo[0] = ldr.x1; lhi[0] = ldr.x2; llo[1] = ldr.y1; lhi[1] = ldr.y2; llo[2] = 0.f; lhi[2] = 0.f; }
#pragma unroll
            for (int ax = 0; ax < kAxes; ++ax) {
                const float gap = fmaxf(glo[ax] - lhi[ax], llo[ax] - ghi[ax]);
                if (gap > 0.f && gap > A.cull_c * ((lhi[ax] - llo[ax]) + gext[ax])) is_near = false;
            }
        }
        unsigned near = __ballot_sync(0xffffffffu, is_near);
        while (near) {                                                       // usually zero or one item per warp
            const int j = __ffs(near) - 1;
            near &= near - 1u;
            const int item = warp + kWarps * j;
            const int slot = (item >> 1) * 64 + 32 * (item & 1) + lane;      // one pair per lane
            const int r = reinterpret_cast<const int32_t*>(s_blk + (size_t)(item >> 1) * kBlkWords)[8 * 64 + (slot & 63)];
            if (r == INT_MAX || !((alive[r >> 5] >> (r & 31)) & 1u)) continue;
            bool out = false;
            if (r == l) {
                out = true;
                atomicOr(&leaderb[l >> 5], 1u << (l & 31));
            } else {
                const RecT mine = record(slot);
                bool unsafe = lbad || ((badb[slot >> 5] >> (slot & 31)) & 1u);
                float v = RecOf<kSrc>::template fast<kGen, kAffine>(mine, ldr, unsafe);
                if (__builtin_expect(unsafe, 0)) v = RecOf<kSrc>::template exact<kGen, kAffine>(mine, ldr);
                if (!(v <= A.thr)) { out = true; dfsup[r] = l; }             // NaN leaves the pool too (:249-250)
            }
            if (out) atomicAnd(&alive[r >> 5], ~(1u << (r & 31)));
        }
        __syncthreads();
    }
    if (tid == 0) *reinterpret_cast<int32_t*>(w + L.dflag) = failed ? 1 : 0;
    if (failed) return;
    uint32_t* dleader = reinterpret_cast<uint32_t*>(w + L.dleader);
    if (tid < (N + 31) / 32) dleader[tid] = tid < 128 ? leaderb[tid] : 0u;
}

// zero the suppression mask of the images elect_kernel gave up on (small grid that loops over (image, eighth) items)
__global__ void __launch_bounds__(256) zero_failed_kernel(int N, int batch, char* __restrict__ ws, size_t ws_img_stride) {
    const WsLayout L = ws_layout(N);
    const size_t tot4 = ((size_t)((N + 31) / 32) * N * 4 + 15) / 16, per = (tot4 + 7) / 8;
    for (int item = blockIdx.x; item < batch * 8; item += gridDim.x) {
        char* w = ws + (size_t)(item >> 3) * ws_img_stride;
        if (*reinterpret_cast<const int32_t*>(w + L.dflag) == 0) continue;
        uint4* m4 = reinterpret_cast<uint4*>(w + L.mask);
        const size_t lo = (size_t)(item & 7) * per, hi = lo + per < tot4 ? lo + per : tot4;
        for (size_t i = lo + threadIdx.x; i < hi; i += 256) m4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ------------------------------------------------------------------------------------------ 3. chain
// shared-memory carve-up of chain_kernel (host and device agree through these two functions)
__host__ __device__ inline size_t chain_wcol_end(int N) {
    size_t NW = (size_t)((N + 31) / 32);
    size_t NWa = (NW + 3) & ~(size_t)3, Na = ((size_t)N + 3) & ~(size_t)3;
    size_t P = 1;
    while ((int)P < N) P <<= 1;
    size_t keys = P * 8, win = (size_t)kWin * NW * 4;
    return 3 * NWa * 4 + Na * 4 * 2 + (keys > win ? keys : win);
}

struct ChainArgs {
    int N, batch;
    const int32_t* n_per_image;
    char* ws;
    size_t ws_img_stride;
    // overlap source for the (member, leader) gather
    int src;                       // BoxSrc
    const float* iou;              // kSrcMatrix
    int64_t ld, iou_img_stride;
    int generalized, affine;       // kSrcBox3d
    gnms_params p;
    // outputs
    int32_t* order;                // [batch,N] in (written by rank_kernel)
    float* sorted_scores;          // [batch,N] in
    float* prob;                   // [batch,N]
    int64_t* valid_idx;
    int64_t* invalid_idx;
    int32_t* counts;               // [batch,2]
    int32_t* lead;                 // saved
    float* pval;
    float* dpval;
    float* pre;
    int32_t* slot;                 // [batch,N] position of sorted pos in the sorted output (or NULL)
    // get_groups / hard-nms outputs (optional)
    int32_t* group_id;             // [N] by INPUT index
    int32_t* group_rank;           // [N] by INPUT index
    int32_t* n_groups;             // [1]
    int32_t* keep;                 // hard NMS: kept input indices in score order
    int32_t* n_keep;
    int leaders_only;              // hard NMS: stop after leader election
    int direct;                    // leaders / first suppressors come from elect_kernel (unless it gave up for this image)
    int have_v;                    // pval[] holds overlap(member, leader) on entry (batched election)
    int tril_input;                // GNMS_MODE_GROUP_MASK_INPUT_TRIL: a member that precedes its leader in INPUT order keeps its score
    int stage;                     // 0: grouping + closed-form rescore + lists (mode GROUP_MASK)
                                   // 1: grouping only, exports the group structure (mode GROUP_NOMASK, before the solve)
                                   // 2: lists only, from pre[] / lead[] already in global memory (after the solve)
};


// warp 0 only: list the positions of the first `limit` set bits of bits[0..nw) into out[], return the count
__device__ int warp_list_bits(const uint32_t* bits, int nw, int limit, int32_t* out) {
    const int lane = threadIdx.x & 31;
    int total = 0;
    for (int base = 0; base < nw && total < limit; base += 32) {
        int wi = base + lane;
        uint32_t w = wi < nw ? bits[wi] : 0u;
        int c = __popc(w);
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        int off = total + incl - c;
        while (w && off < limit) {
            int bpos = __ffs(w) - 1;
            w &= w - 1;
            out[off++] = wi * 32 + bpos;
        }
        total += __shfl_sync(0xffffffffu, incl, 31);
    }
    return min(total, limit);
}

// Optional phase clock (debug): thread 0 of image 0 stamps SM clock values at phase boundaries when enabled through
// gnms_debug_chain_clock(1); read back with gnms_debug_chain_clock_read.  Costs one predicated branch per phase.
#ifdef GNMS_DEBUG
__device__ long long g_chain_clk[64];
__device__ int g_chain_clk_on = 0;
#define GNMS_PHASE(k) do { if (g_chain_clk_on && threadIdx.x == 0 && blockIdx.x == 0) g_chain_clk[k] = clock64(); } while (0)
#else
#define GNMS_PHASE(k) do { } while (0)
#endif

// The sections of shared memory the chain works in.  chain_kernel carves them out of its own dynamic shared memory; the batched
// election (elect2_kernel) hands over its own leader bitset and first-suppressor array and carves the rest out of the record
// area it no longer needs, so that the election's result never leaves the SM.
struct ChainSmem {
    uint32_t* removed;             // NW
    uint32_t* leader;              // NW
    uint32_t* tmpbits;             // NW
    int32_t* fsup;                 // N   (later: lead[])
    int32_t* list;                 // N   (leader / candidate lists, later grank)
    uint32_t* wcol;                // max(kWin * NW words, P * 8 bytes)  (later: sort keys)
    float* ss_s;                   // N   sorted scores staged once
};
// state_in_smem: S.leader / S.fsup already hold the election's result (elect2_kernel calls this at its end)
__device__ __forceinline__ void chain_body(const ChainArgs& A, const int b, const ChainSmem S, const bool state_in_smem) {
    GNMS_PHASE(0);
    const int N = A.N;
    const int n = A.n_per_image ? min(A.n_per_image[b], N) : N;
    const int NW = (N + 31) / 32;
    const int nw = (n + 31) / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WsLayout L = ws_layout(N);
    char* w = A.ws + (size_t)b * A.ws_img_stride;
    const uint32_t* mask = reinterpret_cast<const uint32_t*>(w + L.mask);
    const uint32_t* has_earlier = reinterpret_cast<const uint32_t*>(w + L.has_earlier);
    const float* sbox = reinterpret_cast<const float*>(w + L.sbox);
    const int32_t* order = A.order + (size_t)b * N;
    const float* ss = A.sorted_scores + (size_t)b * N;

    const int NWa = (NW + 3) & ~3, Na = (N + 3) & ~3;                     // keep every section 16-byte aligned
    uint32_t* removed = S.removed;
    uint32_t* leader = S.leader;
    uint32_t* tmpbits = S.tmpbits;
    int32_t* fsup = S.fsup;
    int32_t* list = S.list;
    uint32_t* wcol = S.wcol;
    float* ss_s = S.ss_s;
    __shared__ int s_cnt;
    for (int pos = tid; pos < n; pos += kChainThreads) ss_s[pos] = ss[pos];   // consumed after several barriers

    int32_t* lead = fsup;
    float* pval = A.pval + (size_t)b * N;
    float* dpval = A.dpval + (size_t)b * N;
    const gnms_params P = A.p;
    if (A.stage == 2) {
        for (int pos = tid; pos < n; pos += kChainThreads) {
            fsup[pos] = A.lead[(size_t)b * N + pos];
            pval[pos] = 0.f;
            dpval[pos] = 0.f;
        }
        __syncthreads(); GNMS_PHASE(1);
    } else {
    const bool direct = state_in_smem || (A.direct && *reinterpret_cast<const int32_t*>(w + L.dflag) == 0);
    if (state_in_smem) {
        for (int i = tid; i < NW; i += kChainThreads) removed[i] = 0u;
    } else if (direct) {
        const uint32_t* dleader = reinterpret_cast<const uint32_t*>(w + L.dleader);
        const int32_t* dfsup = reinterpret_cast<const int32_t*>(w + L.dfsup);
        for (int i = tid; i < NW; i += kChainThreads) { removed[i] = 0u; leader[i] = i < nw ? dleader[i] : 0u; }
        for (int i = tid; i < N; i += kChainThreads) fsup[i] = i < n ? dfsup[i] : INT_MAX;
    } else {
    for (int i = tid; i < NW; i += kChainThreads) {
        removed[i] = 0u;
        uint32_t he = 0u;
        if (i < nw)
            for (int sl = 0; sl < L.he_slots; ++sl) he |= has_earlier[(size_t)i * L.he_slots + sl];
        leader[i] = (i < nw) ? (~he & valid_word(i, n)) : 0u;
    }
    for (int i = tid; i < N; i += kChainThreads) fsup[i] = INT_MAX;
    __syncthreads(); GNMS_PHASE(2);

    // ---- certain leaders L0 (no earlier overlapper at all): apply their columns in parallel
    for (int i = tid; i < nw; i += kChainThreads) tmpbits[i] = leader[i];
    __syncthreads(); GNMS_PHASE(3);
    if (warp == 0) {
        int c = warp_list_bits(tmpbits, nw, N, list);
        if (lane == 0) s_cnt = c;
    }
    __syncthreads(); GNMS_PHASE(4);
    {
        // one warp per certain leader: its mask column (all words in flight at once) -> first suppressor of every
        // box it overlaps (atomicMin, distinct addresses).  The `removed` bitset is derived from fsup afterwards:
        // OR-ing it here made 30 warps fight over the same 128 shared-memory words.
        const int cnt = s_cnt;
        for (int k = warp; k < cnt; k += kChainThreads / 32) {
            const int l = list[k];
            constexpr int kMaxW = GNMS_MAX_BOXES / 32 / 32;
            uint32_t mw[kMaxW];
#pragma unroll
            for (int u = 0; u < kMaxW; ++u) {
                const int jw = (l >> 5) + u * 32 + lane;
                mw[u] = jw < nw ? mask[(size_t)l * NW + jw] : 0u;
            }
#pragma unroll
            for (int u = 0; u < kMaxW; ++u) {
                const int jw = (l >> 5) + u * 32 + lane;
                uint32_t m = mw[u];
                while (m) {
                    int bp = __ffs(m) - 1;
                    m &= m - 1;
                    atomicMin(&fsup[jw * 32 + bp], l);
                }
            }
        }
        __syncthreads();
        for (int base = 0; base < nw * 32; base += kChainThreads) {
            const int pos = base + tid;
            const uint32_t bal = __ballot_sync(0xffffffffu, pos < n && fsup[pos] != INT_MAX);
            if (lane == 0 && (pos >> 5) < nw) removed[pos >> 5] = bal;
        }
    }
    // ---- unresolved boxes: windows of kWin candidates, resolved exactly in score order by warp 0
    while (true) {
        __syncthreads(); GNMS_PHASE(5);
        for (int i = tid; i < nw; i += kChainThreads) tmpbits[i] = ~(leader[i] | removed[i]) & valid_word(i, n);
        __syncthreads(); GNMS_PHASE(6);
        if (warp == 0) {
            int c = warp_list_bits(tmpbits, nw, kWin, list);
            if (lane == 0) s_cnt = c;
        }
        __syncthreads(); GNMS_PHASE(7);
        const int cnt = s_cnt;
        if (cnt == 0) break;
        for (int k = warp; k < cnt; k += kChainThreads / 32) {
            const int l = list[k];
            for (int jw = lane; jw < nw; jw += 32)
                wcol[k * NW + jw] = (jw >= (l >> 5)) ? mask[(size_t)l * NW + jw] : 0u;
        }
        __syncthreads(); GNMS_PHASE(8);
        if (warp == 0) {
            for (int k = 0; k < cnt; ++k) {
                const int l = list[k];
                const uint32_t rem = (removed[l >> 5] >> (l & 31)) & 1u;
                if (!rem) {
                    if (lane == 0) leader[l >> 5] |= 1u << (l & 31);
                    for (int jw = (l >> 5) + lane; jw < nw; jw += 32) {
                        uint32_t m = wcol[k * NW + jw];
                        if (m) {
                            removed[jw] |= m;
                            while (m) {
                                int bp = __ffs(m) - 1;
                                m &= m - 1;
                                int q = jw * 32 + bp;
                                fsup[q] = min(fsup[q], l);
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    }   // mask route
    __syncthreads(); GNMS_PHASE(9);

    // ---- hard NMS: the leaders are the keep set
    if (A.leaders_only) {
        if (warp == 0) {
            int c = warp_list_bits(leader, nw, N, list);
            if (lane == 0) { s_cnt = c; A.n_keep[b] = c; }
        }
        __syncthreads(); GNMS_PHASE(10);
        for (int k = tid; k < s_cnt; k += kChainThreads) A.keep[(size_t)b * N + k] = order[list[k]];
        return;
    }

    // ---- group of every box: leaders lead themselves, others follow their first suppressor; a NaN overlap with
    //      the leader means the box left the pool without joining the group (lib/groomed_nms.py:249-250)
    const float* __restrict__ sbox_r = sbox;
    const float* __restrict__ iou_r = A.iou ? A.iou + (size_t)b * A.iou_img_stride : nullptr;
    const int32_t* __restrict__ order_r = order;
    float* __restrict__ pval_r = pval;
    float* __restrict__ dpval_r = dpval;
#pragma unroll 4
    for (int pos = tid; pos < n; pos += kChainThreads) {
        const bool isl = (leader[pos >> 5] >> (pos & 31)) & 1u;
        int f = fsup[pos];
        int ld_ = isl ? pos : (f != INT_MAX ? f : -1);
        float pv = 0.f, dpv = 0.f;
        if (!isl && ld_ >= 0) {
            float v;
            if (A.have_v) {
                v = pval_r[pos];                                       // elect2_kernel left overlap(pos, first suppressor) here
            } else if (A.src == kSrcMatrix) {
                v = iou_r[(int64_t)order_r[pos] * A.ld + order_r[ld_]];
            } else {
                const float4* pa = reinterpret_cast<const float4*>(sbox_r + (size_t)pos * 8);
                const float4* pb = reinterpret_cast<const float4*>(sbox_r + (size_t)ld_ * 8);
                float4 a0 = pa[0], a1 = pa[1], c0 = pb[0], c1 = pb[1];
                if (A.src == kSrcBox3d) {
                    Rec3 ra = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    Rec3 rb = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                    float ib = inter_bev3(ra, rb);
                    bool unsafe = !rec3_sane(ra) || !rec3_sane(rb);              // straight-line division first
                    v = A.generalized ? (A.affine ? iou3_fast<true, true>(ra, rb, ib, unsafe) : iou3_fast<true, false>(ra, rb, ib, unsafe))
                                      : (A.affine ? iou3_fast<false, true>(ra, rb, ib, unsafe) : iou3_fast<false, false>(ra, rb, ib, unsafe));
                    if (__builtin_expect(unsafe, 0))
                        v = A.generalized ? (A.affine ? iou3_exact_slow<true, true>(ra, rb) : iou3_exact_slow<true, false>(ra, rb))
                                          : (A.affine ? iou3_exact_slow<false, true>(ra, rb) : iou3_exact_slow<false, false>(ra, rb));
                } else {
                    Box2 ra = {a0.x, a0.y, a0.z, a0.w, a1.x};
                    Box2 rb = {c0.x, c0.y, c0.z, c0.w, c1.x};
                    bool unsafe = !box2_sane(ra) || !box2_sane(rb);
                    v = iou2_fast(ra, rb, unsafe);
                    if (__builtin_expect(unsafe, 0)) v = iou2_exact_slow(ra, rb);
                }
            }
            if (v != v) {
                ld_ = -1;
            } else if (!(A.tril_input && order_r[pos] < order_r[ld_])) {      // (soft-sort variant: Phi is cut by INPUT position)
                pv = prune(v, P.pruning_method, P.nms_threshold, P.temperature);
                dpv = prune_grad(v, pv, P.pruning_method, P.temperature);
            }
        }
        lead[pos] = ld_;
        pval_r[pos] = pv;
        dpval_r[pos] = dpv;
    }
    __syncthreads(); GNMS_PHASE(11);

    // ---- in-group rank (0 = leader) and the group_size cap: only the first group_size+1 boxes stay (:254-255)
    // (leaders with members are listed first, then one warp per leader; the up-to-8 mask words a lane owns are all
    //  loaded before any is used)
    int32_t* grank = list;
    int32_t* llist = reinterpret_cast<int32_t*>(wcol);
    bool cap_done = false;
    if (!(A.group_id || A.stage == 1) && N >= 1024) {
        // Plain forward: only WHICH boxes exceed the cap matters, and only groups of more than group_size + 1 boxes have any.
        // Position-parallel, no mask columns: (1) group sizes -- the lanes of a 32-position word that share a leader add once;
        // (2) the big leaders get dense slots; (3) warp w walks ITS run of consecutive words in order and keeps, per slot, how
        // many boxes of the group precede each word inside the run; (4) an exclusive prefix over the 32 runs per slot;
        // (5) rank = boxes of the group before this one (the leader is entry 0): beyond group_size -> no group.
        constexpr int kCapSlots = 64, kCapW = GNMS_MAX_BOXES / 32 / 32;
        int32_t* gcount = grank;
        int32_t* bigpref = reinterpret_cast<int32_t*>(wcol);                       // [NW]
        uint16_t* tbl = reinterpret_cast<uint16_t*>(bigpref + NWa);                // [32 runs][kCapSlots]
        const int WPW = (nw + 31) / 32, gs1 = P.group_size;
        // Few leaders (a detector image): every leader gets a slot, so sizes and ranks come out of the same per-run pass and
        // no atomics are needed at all.
        __shared__ int s_gsize[kCapSlots];
        if (warp == 0) {
            int run = 0;
            for (int base = 0; base < nw; base += 32) {
                const int wi = base + lane;
                const int c = wi < nw ? __popc(leader[wi]) : 0;
                int incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (wi < nw) bigpref[wi] = run + incl - c;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) s_cnt = run;
        }
        for (int i = tid; i < 32 * kCapSlots; i += kChainThreads) tbl[i] = 0;
        __syncthreads();
        if (s_cnt <= kCapSlots) {
            cap_done = true;
            int slot[kCapW], pre[kCapW];
            uint32_t cls[kCapW];
#pragma unroll
            for (int u = 0; u < kCapW; ++u) {
                slot[u] = -1; pre[u] = 0; cls[u] = 0u;
                const int wd = warp * WPW + u;
                if (u < WPW && wd < nw) {                                          // (warp-uniform)
                    const int pos = wd * 32 + lane;
                    const int l = pos < n ? lead[pos] : -1;
                    if (l >= 0) slot[u] = bigpref[l >> 5] + __popc(leader[l >> 5] & ((1u << (l & 31)) - 1u));
                    cls[u] = __match_any_sync(0xffffffffu, slot[u]);
                    if (slot[u] >= 0) pre[u] = tbl[warp * kCapSlots + slot[u]];
                    __syncwarp();
                    if (slot[u] >= 0 && lane == __ffs(cls[u]) - 1) tbl[warp * kCapSlots + slot[u]] += (uint16_t)__popc(cls[u]);
                    __syncwarp();
                }
            }
            __syncthreads();
            for (int sl = warp; sl < s_cnt; sl += kChainThreads / 32) {            // lane = run
                const int v = tbl[lane * kCapSlots + sl];
                int incl = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                tbl[lane * kCapSlots + sl] = (uint16_t)(incl - v);
                if (lane == 31) s_gsize[sl] = incl;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < kCapW; ++u) {
                if (slot[u] >= 0 && s_gsize[slot[u]] > gs1 + 1) {
                    const int rk = (int)tbl[warp * kCapSlots + slot[u]] + pre[u] + __popc(cls[u] & ((1u << lane) - 1u));
                    if (rk > gs1) lead[(warp * WPW + u) * 32 + lane] = -1;
                }
            }
            __syncthreads(); GNMS_PHASE(17);
        }
        if (!cap_done) {
        for (int pos = tid; pos < n; pos += kChainThreads) gcount[pos] = 0;
        __syncthreads();
        int ldr[kCapW];
        uint32_t cls[kCapW];
#pragma unroll
        for (int u = 0; u < kCapW; ++u) {
            ldr[u] = -1; cls[u] = 0u;
            const int wd = warp * WPW + u;
            if (u < WPW && wd < nw) {                                              // (warp-uniform)
                const int pos = wd * 32 + lane;
                const int l = pos < n ? lead[pos] : -1;
                const uint32_t mm = __match_any_sync(0xffffffffu, l);
                ldr[u] = l; cls[u] = mm;
                if (l >= 0 && lane == __ffs(mm) - 1) atomicAdd(&gcount[l], __popc(mm));
            }
        }
        __syncthreads();
        for (int base = 0; base < nw * 32; base += kChainThreads) {
            const int pos = base + tid;
            const bool big = pos < n && lead[pos] == pos && gcount[pos] > gs1 + 1;
            const uint32_t bal = __ballot_sync(0xffffffffu, big);
            if (lane == 0 && (pos >> 5) < nw) tmpbits[pos >> 5] = bal;
        }
        __syncthreads();
        if (warp == 0) {
            int run = 0;
            for (int base = 0; base < nw; base += 32) {
                const int wi = base + lane;
                const int c = wi < nw ? __popc(tmpbits[wi]) : 0;
                int incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (wi < nw) bigpref[wi] = run + incl - c;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) s_cnt = run;
        }
        __syncthreads();
        const int nbig = s_cnt;
        if (nbig <= kCapSlots) {
            cap_done = true;
            if (nbig > 0) {
                int slot[kCapW], pre[kCapW];
#pragma unroll
                for (int u = 0; u < kCapW; ++u) {
                    slot[u] = -1; pre[u] = 0;
                    const int wd = warp * WPW + u;
                    if (u < WPW && wd < nw) {
                        const int l = ldr[u];
                        if (l >= 0 && ((tmpbits[l >> 5] >> (l & 31)) & 1u))
                            slot[u] = bigpref[l >> 5] + __popc(tmpbits[l >> 5] & ((1u << (l & 31)) - 1u));
                        if (slot[u] >= 0) pre[u] = tbl[warp * kCapSlots + slot[u]];
                        __syncwarp();
                        if (slot[u] >= 0 && lane == __ffs(cls[u]) - 1) tbl[warp * kCapSlots + slot[u]] += (uint16_t)__popc(cls[u]);
                        __syncwarp();
                    }
                }
                __syncthreads();
                for (int sl = warp; sl < nbig; sl += kChainThreads / 32) {         // lane = run
                    const int v = tbl[lane * kCapSlots + sl];
                    int incl = v;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += t;
                    }
                    tbl[lane * kCapSlots + sl] = (uint16_t)(incl - v);
                }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < kCapW; ++u) {
                    if (slot[u] >= 0) {
                        const int rk = (int)tbl[warp * kCapSlots + slot[u]] + pre[u] + __popc(cls[u] & ((1u << lane) - 1u));
                        if (rk > gs1) lead[(warp * WPW + u) * 32 + lane] = -1;
                    }
                }
            }
            __syncthreads(); GNMS_PHASE(17);
        }
        }   // more than kCapSlots leaders
    }
    if (!cap_done) {
    // member count per leader (shared-memory atomics); ranks are only needed where the cap can bite
    // (count > group_size) or when the caller wants the groups themselves
    for (int pos = tid; pos < n; pos += kChainThreads) grank[pos] = 0;
    __syncthreads(); GNMS_PHASE(12);
    for (int pos = tid; pos < n; pos += kChainThreads) {
        const int ld_ = lead[pos];
        if (ld_ >= 0 && ld_ != pos) atomicAdd(&grank[ld_], 1);
    }
    __syncthreads(); GNMS_PHASE(13);
    const int need_above = (A.group_id || A.stage == 1) ? 0 : P.group_size;
    for (int i = tid; i < nw; i += kChainThreads) {
        uint32_t lw = leader[i], keep = 0u;
        while (lw) {
            const int bp = __ffs(lw) - 1;
            lw &= lw - 1;
            if (grank[i * 32 + bp] > need_above) keep |= 1u << bp;
        }
        tmpbits[i] = keep;
    }
    __syncthreads(); GNMS_PHASE(14);
    for (int pos = tid; pos < n; pos += kChainThreads) grank[pos] = 0;
    __syncthreads(); GNMS_PHASE(15);
    if (warp == 0) {
        int c = warp_list_bits(tmpbits, nw, N, llist);
        if (lane == 0) s_cnt = c;
    }
    __syncthreads(); GNMS_PHASE(16);
    {
        // one warp per listed leader: its mask column names the candidates (few set bits); those that really follow
        // this leader are ranked by position with a warp scan of per-word popcounts
        const int nlead = s_cnt;
        constexpr int kMaxW = GNMS_MAX_BOXES / 32 / 32;          // mask words per lane per column (8)
        if (direct && nlead <= kWin && N >= 256) {
            // no mask on the direct route, few listed leaders (the usual case): position-parallel.  A warp takes 32 consecutive
            // positions, groups its lanes by leader (match.any) and records the size of each class in cnt[leader slot][word];
            // one warp per listed leader then turns its row of cnt into an exclusive prefix over the words; the rank of a member
            // (the leader counts as entry 0 of its group) is that prefix plus its place inside its class.
            uint16_t* cnt = reinterpret_cast<uint16_t*>(wcol + kWin);              // [kWin][nw]   (wcol holds kWin * NW words; llist sits in the first kWin)
            uint16_t* dmap = cnt + (size_t)kWin * NW;                              // [n] leader position -> slot, 0xffff otherwise
            for (int i = tid; i < nlead * nw; i += kChainThreads) cnt[(i / nw) * NW + (i % nw)] = 0;
            for (int pos = tid; pos < n; pos += kChainThreads) dmap[pos] = 0xffffu;
            __syncthreads();
            for (int k = tid; k < nlead; k += kChainThreads) dmap[llist[k]] = (uint16_t)k;
            __syncthreads();
            uint32_t cls[kMaxW];
            int slot_[kMaxW];
#pragma unroll
            for (int u = 0; u < kMaxW; ++u) {
                const int wd = warp + u * (kChainThreads / 32);
                cls[u] = 0u; slot_[u] = -1;
                if (wd < nw) {                                                      // (warp-uniform)
                    const int pos = wd * 32 + lane;
                    const int l = pos < n ? lead[pos] : -1;
                    const int d = l >= 0 ? (int)dmap[l] : 0xffff;
                    const uint32_t mm = __match_any_sync(0xffffffffu, d != 0xffff ? d : 0x10000 + lane);
                    if (d != 0xffff) {
                        cls[u] = mm; slot_[u] = d;
                        if (lane == __ffs(mm) - 1) cnt[(size_t)d * NW + wd] = (uint16_t)__popc(mm);
                    }
                }
            }
            __syncthreads();
            for (int k = warp; k < nlead; k += kChainThreads / 32) {
                int run = 0;
                for (int base = 0; base < nw; base += 32) {
                    const int wi = base + lane;
                    const int c = wi < nw ? (int)cnt[(size_t)k * NW + wi] : 0;
                    int incl = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += t;
                    }
                    if (wi < nw) cnt[(size_t)k * NW + wi] = (uint16_t)(run + incl - c);
                    run += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < kMaxW; ++u) {
                const int wd = warp + u * (kChainThreads / 32);
                if (slot_[u] >= 0) {
                    const int pos = wd * 32 + lane;
                    const int rk = (int)cnt[(size_t)slot_[u] * NW + wd] + __popc(cls[u] & ((1u << lane) - 1u));
                    if (rk > 0) grank[pos] = rk;                                   // (rank 0 is the leader itself)
                }
            }
        } else if (direct) {
            // many listed leaders: one warp per listed leader walks lead[] from the leader on, 32 positions a step
            for (int k = warp; k < nlead; k += kChainThreads / 32) {
                const int l = llist[k];
                int before = 0;
                for (int base = l & ~31; base < n; base += 32) {
                    const int pos = base + lane;
                    const bool memb = pos < n && pos != l && lead[pos] == l;
                    const unsigned bal = __ballot_sync(0xffffffffu, memb);
                    if (memb) grank[pos] = 1 + before + __popc(bal & ((1u << lane) - 1u));
                    before += __popc(bal);
                }
            }
        } else
        for (int k = warp; k < nlead; k += kChainThreads / 32) {
            const int l = llist[k];
            const int base0 = l >> 5;
            uint32_t mw[kMaxW];
#pragma unroll
            for (int u = 0; u < kMaxW; ++u) {
                const int jw = base0 + u * 32 + lane;
                mw[u] = jw < nw ? mask[(size_t)l * NW + jw] : 0u;
            }
            int carry = 0;
#pragma unroll
            for (int u = 0; u < kMaxW; ++u) {
                const int jw = base0 + u * 32 + lane;
                if (base0 + u * 32 >= nw) break;
                uint32_t m = mw[u], memb = 0u;
                while (m) {
                    int bp = __ffs(m) - 1;
                    m &= m - 1;
                    if (lead[jw * 32 + bp] == l) memb |= 1u << bp;
                }
                int c = __popc(memb), incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                const int before = carry + incl - c;
                uint32_t mm = memb;
                while (mm) {
                    int bp = __ffs(mm) - 1;
                    mm &= mm - 1;
                    grank[jw * 32 + bp] = 1 + before + __popc(memb & ((1u << bp) - 1u));
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
    }
    __syncthreads();
    const int gs = P.group_size;
    for (int pos = tid; pos < n; pos += kChainThreads) {
        if (lead[pos] >= 0 && lead[pos] != pos && grank[pos] > gs) lead[pos] = -1;
    }
    }   // !cap_done
    __syncthreads(); GNMS_PHASE(18);

    // ---- get_groups outputs
    if (A.group_id) {
        // exclusive prefix of leader popcounts per word (NW <= 256 words: warp 0, 8 words per lane)
        if (warp == 0) {
            int run = 0;
            for (int base = 0; base < nw; base += 32) {
                int wi = base + lane;
                int c = wi < nw ? __popc(leader[wi]) : 0, incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (wi < nw) tmpbits[wi] = run + incl - c;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) A.n_groups[b] = run;
        }
        __syncthreads(); GNMS_PHASE(19);
        for (int pos = tid; pos < n; pos += kChainThreads) {
            int ld_ = lead[pos];
            int gid = -1;
            if (ld_ >= 0) gid = tmpbits[ld_ >> 5] + __popc(leader[ld_ >> 5] & ((1u << (ld_ & 31)) - 1u));
            A.group_id[(size_t)b * N + order[pos]] = gid;
            A.group_rank[(size_t)b * N + order[pos]] = ld_ >= 0 ? grank[pos] : -1;
        }
        if (!A.prob) return;
        __syncthreads(); GNMS_PHASE(20);
    }

    // ---- mode GROUP_NOMASK: export the group structure for the per-group triangular solves and stop
    if (A.stage == 1) {
        __shared__ int s_wsum[32];
        int32_t* gsz = reinterpret_cast<int32_t*>(wcol);              // [Na] group size at the leader's position
        int32_t* gpre = gsz + Na;                                     // [Na] exclusive prefix
        int32_t* gbeg = reinterpret_cast<int32_t*>(w + L.gbeg);
        int32_t* members = reinterpret_cast<int32_t*>(w + L.members);
        int32_t* ngroups = reinterpret_cast<int32_t*>(w + L.ngroups);
        int32_t* lead_out = A.lead + (size_t)b * N;
        for (int pos = tid; pos < Na; pos += kChainThreads) gsz[pos] = 0;
        __syncthreads(); GNMS_PHASE(21);
        for (int pos = tid; pos < n; pos += kChainThreads) {
            const int ld_ = lead[pos];
            if (ld_ >= 0) atomicAdd(&gsz[ld_], 1);                   // leaders count themselves (lead[l] == l)
        }
        __syncthreads(); GNMS_PHASE(22);
        // block-wide exclusive scan: thread t owns the chunk [t*C, (t+1)*C)
        const int C = (n + kChainThreads - 1) / kChainThreads;
        int local = 0;
        for (int q = 0; q < C; ++q) { const int pos = tid * C + q; if (pos < n) local += gsz[pos]; }
        int incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads(); GNMS_PHASE(23);
        if (warp == 0) {
            int v = s_wsum[lane], iv = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, iv, d);
                if (lane >= d) iv += t;
            }
            s_wsum[lane] = iv - v;
            if (lane == 31) s_cnt = iv;                               // boxes that are in some group
        }
        __syncthreads(); GNMS_PHASE(24);
        int run = s_wsum[warp] + incl - local;
        for (int q = 0; q < C; ++q) { const int pos = tid * C + q; if (pos < n) { gpre[pos] = run; run += gsz[pos]; } }
        // dense group index of a leader = number of leaders before it
        if (warp == 0) {
            int runl = 0;
            for (int base = 0; base < nw; base += 32) {
                int wi = base + lane;
                int c = wi < nw ? __popc(leader[wi]) : 0, inc2 = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, inc2, d);
                    if (lane >= d) inc2 += t;
                }
                if (wi < nw) tmpbits[wi] = runl + inc2 - c;
                runl += __shfl_sync(0xffffffffu, inc2, 31);
            }
            if (lane == 0) { ngroups[0] = runl; gbeg[runl] = s_cnt; }
        }
        __syncthreads(); GNMS_PHASE(25);
        for (int pos = tid; pos < N; pos += kChainThreads) {
            const int ld_ = pos < n ? lead[pos] : -1;
            lead_out[pos] = ld_;
            if (ld_ >= 0) {
                members[gpre[ld_] + (ld_ == pos ? 0 : grank[pos])] = pos;
                if (ld_ == pos) gbeg[tmpbits[pos >> 5] + __popc(leader[pos >> 5] & ((1u << (pos & 31)) - 1u))] = gpre[pos];
            }
        }
        return;
    }
    }   // A.stage != 2

    // ---- rescore (mode GROUP_MASK closed form: row of I - Phi has two non-zeros), clamp, threshold
    float* pre_out = A.pre + (size_t)b * N;
    int32_t* lead_out = A.lead + (size_t)b * N;
    float* prob = A.prob + (size_t)b * N;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(wcol);
    const float vthr = P.valid_box_prob_threshold;
    // validity bitset in tmpbits via ballot (warp handles 32 consecutive positions)
    for (int base = 0; base < nw * 32; base += kChainThreads) {
        const int pos = base + tid;
        float r = 0.f, rt = 0.f;
        bool is_valid = false;
        if (pos < n) {
            const int ld_ = lead[pos];
            float prev = 0.f;
            if (A.stage == 2) prev = ld_ >= 0 ? pre_out[pos] : 0.f;          // from the triangular solve
            else if (ld_ == pos) prev = ss_s[pos];
            else if (ld_ >= 0) prev = __fsub_rn(ss_s[pos], __fmul_rn(pval[pos], ss_s[ld_]));
            r = fminf(fmaxf(prev, 0.f), 1.f);
            rt = (r < vthr) ? 0.f : r;
            is_valid = rt >= vthr;
            pre_out[pos] = prev;
            lead_out[pos] = ld_;
            if (!P.sorted_output) prob[pos] = P.thresholded_output ? rt : r;
            // stash the thresholded value for the sort below
            reinterpret_cast<float*>(list)[pos] = rt;
        }
        uint32_t bal = __ballot_sync(0xffffffffu, is_valid);
        if (lane == 0 && (pos >> 5) < nw) tmpbits[pos >> 5] = bal;
    }
    for (int pos = n + tid; pos < N; pos += kChainThreads) {
        pre_out[pos] = 0.f;
        lead_out[pos] = -1;
        prob[pos] = 0.f;
        pval[pos] = 0.f;
        dpval[pos] = 0.f;
        if (A.slot) A.slot[(size_t)b * N + pos] = pos;
    }
    __syncthreads(); GNMS_PHASE(26);
    // word prefix of valid counts -> tmp in `removed` (no longer needed)
    uint32_t* vpref = removed;
    if (warp == 0) {
        int run = 0;
        for (int base = 0; base < nw; base += 32) {
            int wi = base + lane;
            int c = wi < nw ? __popc(tmpbits[wi]) : 0, incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (wi < nw) vpref[wi] = run + incl - c;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_cnt = run;
    }
    __syncthreads(); GNMS_PHASE(27);
    const int V = s_cnt;
    if (tid == 0) { A.counts[b * 2 + 0] = V; A.counts[b * 2 + 1] = n - V; }
    int PV = 1;
    while (PV < V) PV <<= 1;
    for (int i = tid; i < PV; i += kChainThreads) keys[i] = ~0ull;
    __syncthreads(); GNMS_PHASE(28);
    const float* rthr = reinterpret_cast<const float*>(list);
    int64_t* invalid_idx = A.invalid_idx + (size_t)b * N;
    int64_t* valid_idx = A.valid_idx + (size_t)b * N;
    for (int pos = tid; pos < n; pos += kChainThreads) {
        const uint32_t wv = tmpbits[pos >> 5];
        const int before = vpref[pos >> 5] + __popc(wv & ((1u << (pos & 31)) - 1u));
        if ((wv >> (pos & 31)) & 1u) {
            keys[before] = ((unsigned long long)desc_key(rthr[pos]) << 32) | (uint32_t)pos;
        } else {
            const int irank = pos - before;
            invalid_idx[irank] = order[pos];
            if (A.slot) A.slot[(size_t)b * N + pos] = V + irank;
            if (P.sorted_output) prob[V + irank] = 0.f;
        }
    }
    __syncthreads(); GNMS_PHASE(29);
    if (V <= kChainThreads) {
        // short valid list (the usual case): rank by counting, one key per thread, no barriers
        if (tid < V) {
            const unsigned long long mine = keys[tid];
            int r = 0;
            for (int j = 0; j < V; ++j) r += (keys[j] < mine);
            const int pos = (int)(uint32_t)(mine & 0xffffffffull);
            valid_idx[r] = order[pos];
            if (A.slot) A.slot[(size_t)b * N + pos] = r;
            if (P.sorted_output) prob[r] = rthr[pos];
        }
        return;
    }
    for (int k = 2; k <= PV; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (PV >> 1); t += kChainThreads) {
                int i = 2 * t - (t & (j - 1));
                int l = i + j;
                unsigned long long a = keys[i], c = keys[l];
                bool up = ((i & k) == 0);
                if ((a > c) == up) { keys[i] = c; keys[l] = a; }
            }
            __syncthreads(); GNMS_PHASE(30);
        }
    }
    for (int i = tid; i < V; i += kChainThreads) {
        const int pos = (int)(uint32_t)(keys[i] & 0xffffffffull);
        valid_idx[i] = order[pos];
        if (A.slot) A.slot[(size_t)b * N + pos] = i;
        if (P.sorted_output) prob[i] = rthr[pos];
    }
}

__global__ void __launch_bounds__(kChainThreads) chain_kernel(ChainArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = A.N, NW = (N + 31) / 32;
    const int NWa = (NW + 3) & ~3, Na = (N + 3) & ~3;
    ChainSmem S;
    S.removed = reinterpret_cast<uint32_t*>(smem_raw);
    S.leader = S.removed + NWa;
    S.tmpbits = S.leader + NWa;
    S.fsup = reinterpret_cast<int32_t*>(S.tmpbits + NWa);
    S.list = S.fsup + Na;
    S.wcol = reinterpret_cast<uint32_t*>(S.list + Na);
    S.ss_s = reinterpret_cast<float*>(smem_raw + chain_wcol_end(N));
    chain_body(A, blockIdx.x, S, false);
}

// ------------------------------------------------------------------------------------------ 2f. batched leader election
// The per-leader loop above pays one barrier-bound step per leader.  The greedy election is exact under batching: take the
// first K <= 32 boxes still in the pool (score order).  Everything ahead of them has been decided, so among themselves they
// are resolved by the K x K overlap bits alone (a candidate is a leader iff every earlier candidate that overlaps it is itself
// suppressed); and every other box of the pool then leaves with the LOWEST new leader it overlaps -- exactly the leader the
// one-at-a-time loop (lib/groomed_nms.py:247-262) would have reached first.  A detector image needs 2-5 such steps instead of
// one per leader.  One CTA per image, 1024 threads, the image's sorted records in shared memory (N <= 4096):
//   A  warp 0 lists the first K pool positions
//   B  warp i evaluates candidate i against candidates j < i
//   C  warp 0 resolves the K x K bits and publishes the new leaders: their records, and for the two widest axes a table
//      "which of this step's leaders can reach a box whose lower edge falls in bin b" (64 bins per axis, conservative)
//   D  every pool box: two table look-ups name the few leaders that can reach it; for those the conservative gap bound of the
//      tile culling on every axis (gap > c * (sum of the two extents) => overlap <= thr; exact zeros when c == 0); the pairs
//      that survive go to a shared-memory queue
//   E  the queue is evaluated densely, one pair per thread, atomicMin of the leader position per box; the overlap of a box
//      with the leader that won is written out (it is the Phi entry the rescore needs: chain_kernel does not evaluate it again)
//   F  boxes with a suppressor leave the pool (one ballot per 32 positions)
// If the queue overflows the step is redone with the first quarter of its leaders (the rest of the candidates go back to the
// pool): with one leader it always fits, so the kernel never gives up -- no fallback kernels follow it.
__device__ int warp_list_bits(const uint32_t* bits, int nw, int limit, int32_t* out);     // (defined with chain_kernel below)
constexpr int kE2Threads = 1024;
constexpr int kE2Queue = 8192;
constexpr int kE2PerThread = kE2Queue / kE2Threads;
constexpr int kE2MaxK = kElectMaxBoxes / kE2Threads;        // pool positions per thread
constexpr int kE2Bins = 64;
struct Elect2Args {
    int N, batch;
    const int32_t* n_per_image;
    char* ws;
    size_t ws_img_stride;
    float thr, cull_c;
    float* vout;                   // [batch, N] by sorted position: overlap with the first suppressor (members only)
    int fuse_chain;                // != 0: run the chain (grouping cap, rescore, lists) right here, from shared memory
    ChainArgs chain;
};
static size_t elect2_smem_bytes(int N) {
    const size_t Np = ((size_t)N + kE2Threads - 1) / kE2Threads * kE2Threads;
    return Np * 32 + Np * 4 + (size_t)kE2Queue * 4 + 4 * 128 * 4 + 32 * 3 * 16 + 5 * 32 * 4 + 32 * 32 * 4 + 2 * kE2Bins * 4 + Np * 2 + 64;
}

template <int kSrc> struct AxesOf;
template <> struct AxesOf<kSrcBox3d> {          // depth first: detector outputs spread most along z, then x
    static constexpr int kAxes = 3;
    static __device__ __forceinline__ void get(const Rec3& r, float (&lo)[3], float (&hi)[3]) {
        lo[0] = r.bz1; hi[0] = r.bz2; lo[1] = r.bx1; hi[1] = r.bx2; lo[2] = r.ymin; hi[2] = r.ymax;
    }
};
template <> struct AxesOf<kSrcBox2d> {
    static constexpr int kAxes = 2;
    static __device__ __forceinline__ void get(const Box2& r, float (&lo)[3], float (&hi)[3]) {
        lo[0] = r.x1; hi[0] = r.x2; lo[1] = r.y1; hi[1] = r.y2; lo[2] = 0.f; hi[2] = 0.f;
    }
};
// bin of a lower edge: monotone in x for fixed (origin, scale >= 0), which is all the reach tables rely on
__device__ __forceinline__ int e2_bin(float x, float origin, float scale) {
    const float t = __fmul_rn(__fsub_rn(x, origin), scale);
    return t >= (float)(kE2Bins - 1) ? kE2Bins - 1 : (t > 0.f ? (int)t : 0);         // (NaN -> 0; such boxes never use the tables)
}

#ifdef GNMS_DEBUG
__device__ long long g_e2_clk[16];          // [0..7] cycles per phase (load, A, B, C, D, E, F, store), [8] steps, [9] leaders, [10] queued pairs, [11] redone steps
#define GNMS_E2_T(k) do { if (tid == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_e2_clk[k] += t_ - e2_t; e2_t = t_; } } while (0)
#define GNMS_E2_C(k, v) do { if (tid == 0 && blockIdx.x == 0) g_e2_clk[k] += (v); } while (0)
#else
#define GNMS_E2_T(k) do { } while (0)
#define GNMS_E2_C(k, v) do { } while (0)
#endif
template <int kSrc, bool kGen, bool kAffine>
__global__ void __launch_bounds__(kE2Threads) elect2_kernel(Elect2Args A) {
    typedef typename RecOf<kSrc>::type RecT;
    constexpr int kAxes = AxesOf<kSrc>::kAxes;
    extern __shared__ __align__(16) unsigned char s_e2[];
    const int b = blockIdx.x, N = A.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef GNMS_DEBUG
    long long e2_t = clock64();
#endif
    const int n = A.n_per_image ? min(A.n_per_image[b], N) : N;
    const int nw = (n + 31) / 32;
    const int Np = (N + kE2Threads - 1) / kE2Threads * kE2Threads, KK = Np / kE2Threads;
    const WsLayout L = ws_layout(N);
    char* w = A.ws + (size_t)b * A.ws_img_stride;
    float* vout = A.vout + (size_t)b * N;
    float4* rec4 = reinterpret_cast<float4*>(s_e2);                                  // [Np][2] sorted records
    int32_t* fsup = reinterpret_cast<int32_t*>(rec4 + (size_t)Np * 2);              // [Np] first suppressor (sorted position)
    uint32_t* queue = reinterpret_cast<uint32_t*>(fsup + Np);                        // [kE2Queue] (position << 5) | leader slot
    uint32_t* alive = queue + kE2Queue;                                              // [128] the pool, by sorted position
    uint32_t* leaderb = alive + 128;
    uint32_t* badb = leaderb + 128;                                                  // record outside div_rn_fast's proven range
    float4* ltab = reinterpret_cast<float4*>(badb + 128);                            // [32][3]: (lo, hi, extent, bad) per axis
    int32_t* cand = reinterpret_cast<int32_t*>(ltab + 32 * 3);                       // [32] candidate positions
    uint32_t* covm = reinterpret_cast<uint32_t*>(cand + 32);                         // [32] bit j of row i: !(overlap(i, j) <= thr)
    int32_t* lpos = reinterpret_cast<int32_t*>(covm + 32);                           // [32] position of leader slot q
    int32_t* lcidx = lpos + 32;                                                      // [32] candidate index of leader slot q
    float* covv = reinterpret_cast<float*>(lcidx + 32);                              // [32][32] overlap(candidate i, candidate j < i)
    uint32_t* reach = reinterpret_cast<uint32_t*>(covv + 32 * 32);                   // [2][kE2Bins] leader slots that reach a bin
    uint32_t* lbins = reach + 2 * kE2Bins;                                           // [32] leader slot q: first / last bin on axes 0, 1 (4 bytes)
    uint32_t* apref = lbins + 32;                                                    // [128] pool bits before each word
    uint16_t* bins = reinterpret_cast<uint16_t*>(apref + 128);                       // [Np] bin on axis 0 | bin on axis 1 << 8 (0xffff: bad box)
    __shared__ int s_K, s_m, s_qn, s_over;
    __shared__ uint32_t s_leaders;
    __shared__ float s_red[32][8];
    __shared__ float s_binp[8];               // per axis a in {0, 1}: origin, scale, largest extent, absolute margin
    // the image's sorted records (32 n bytes, contiguous): 1-D bulk copies by the copy engine, one thread issues them and an
    // mbarrier counts the bytes in -- no load instructions, no registers, the other threads set up the pool meanwhile
    __shared__ __align__(8) unsigned long long s_mbar;
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)n * 32u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
        const char* src = w + L.sbox;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(rec4);
        for (uint32_t off = 0; off < bytes; off += 16384u) {
            const uint32_t sz = bytes - off < 16384u ? bytes - off : 16384u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst + off), "l"(src + off), "r"(sz), "r"(mbar) : "memory");
        }
    }
    if (tid < 128) { alive[tid] = tid < nw ? valid_word(tid, n) : 0u; leaderb[tid] = 0u; }
    auto record = [&](int pos) -> RecT {
        const float4 u = rec4[2 * pos], v = rec4[2 * pos + 1];
        if constexpr (kSrc == kSrcBox3d) return Rec3{u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
        else return Box2{u.x, u.y, u.z, u.w, v.x};
    };
    auto is_bad = [&](int pos) -> bool { return (badb[pos >> 5] >> (pos & 31)) & 1u; };
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mbar) : "memory");
        }
    }
    __syncthreads();
    {
        // bad flags; per axis (0, 1): range of the lower edges, largest extent, largest magnitude -- over the sane boxes
        float mn0 = INFINITY, mx0 = -INFINITY, mn1 = INFINITY, mx1 = -INFINITY, ex0 = 0.f, ex1 = 0.f, mg = 0.f;
        for (int k = 0; k < KK; ++k) {
            const int pos = k * kE2Threads + tid;
            fsup[pos] = INT_MAX;
            bool bad = false;
            if (pos < n) {
                const RecT r = record(pos);
                if constexpr (kSrc == kSrcBox3d) bad = !rec3_sane(r);
                else bad = !box2_sane(r);
                if (!bad) {
                    float lo[3], hi[3];
                    AxesOf<kSrc>::get(r, lo, hi);
                    mn0 = fminf(mn0, lo[0]); mx0 = fmaxf(mx0, lo[0]); ex0 = fmaxf(ex0, __fsub_rn(hi[0], lo[0]));
                    mn1 = fminf(mn1, lo[1]); mx1 = fmaxf(mx1, lo[1]); ex1 = fmaxf(ex1, __fsub_rn(hi[1], lo[1]));
                    mg = fmaxf(mg, fmaxf(fmaxf(fabsf(lo[0]), fabsf(hi[0])), fmaxf(fabsf(lo[1]), fabsf(hi[1]))));
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, bad);
            if (lane == 0) badb[k * 32 + warp] = bal;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, d)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, d));
            mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, d)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, d));
            ex0 = fmaxf(ex0, __shfl_xor_sync(0xffffffffu, ex0, d)); ex1 = fmaxf(ex1, __shfl_xor_sync(0xffffffffu, ex1, d));
            mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, d));
        }
        if (lane == 0) { s_red[warp][0] = mn0; s_red[warp][1] = mx0; s_red[warp][2] = mn1; s_red[warp][3] = mx1; s_red[warp][4] = ex0; s_red[warp][5] = ex1; s_red[warp][6] = mg; }
        __syncthreads();
        if (warp == 0) {
            mn0 = s_red[lane][0]; mx0 = s_red[lane][1]; mn1 = s_red[lane][2]; mx1 = s_red[lane][3]; ex0 = s_red[lane][4]; ex1 = s_red[lane][5]; mg = s_red[lane][6];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, d)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, d));
                mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, d)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, d));
                ex0 = fmaxf(ex0, __shfl_xor_sync(0xffffffffu, ex0, d)); ex1 = fmaxf(ex1, __shfl_xor_sync(0xffffffffu, ex1, d));
                mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, d));
            }
            if (lane == 0) {
                // scale = bins / range (0 when the range is empty, not finite or absurd: then every box sits in bin 0 and the
                // tables cull nothing); margin: a few ulps of the largest coordinate, it absorbs the rounding of the reach edges
                const float r0 = __fsub_rn(mx0, mn0), r1 = __fsub_rn(mx1, mn1);
                s_binp[0] = mn0; s_binp[1] = (r0 > 0.f && r0 < 1e30f) ? __fdiv_rn((float)kE2Bins, r0) : 0.f; s_binp[2] = ex0;
                s_binp[4] = mn1; s_binp[5] = (r1 > 0.f && r1 < 1e30f) ? __fdiv_rn((float)kE2Bins, r1) : 0.f; s_binp[6] = ex1;
                s_binp[3] = s_binp[7] = __fmul_rn(__fadd_rn(mg, fmaxf(ex0, ex1)), 1.9073486e-6f);                    // 2^-19
            }
        }
    }
    const float thr = A.thr, cc = A.cull_c;
    __syncthreads();
    {
        const float bo0 = s_binp[0], bs0 = s_binp[1], bo1 = s_binp[4], bs1 = s_binp[5];
        for (int k = 0; k < KK; ++k) {
            const int pos = k * kE2Threads + tid;
            uint16_t bb = 0xffffu;
            if (pos < n && !is_bad(pos)) {
                float lo[3], hi[3];
                AxesOf<kSrc>::get(record(pos), lo, hi);
                bb = (uint16_t)(e2_bin(lo[0], bo0, bs0) | (e2_bin(lo[1], bo1, bs1) << 8));
            }
            bins[pos] = bb;
        }
    }
    GNMS_E2_T(0);
    while (true) {
        // ---- A: the first <= 32 positions of the pool.  Lane i counts words 4i .. 4i+3, a warp scan gives the number of pool
        //      bits before every word, and output slot o (lane o) finds its word by bisection and its bit with __fns.
        if (warp == 0) {
            const uint4 wv = reinterpret_cast<const uint4*>(alive)[lane];
            const int c0 = __popc(wv.x), c1 = __popc(wv.y), c2 = __popc(wv.z), c3 = __popc(wv.w);
            int incl = c0 + c1 + c2 + c3;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int before = incl - (c0 + c1 + c2 + c3);
            reinterpret_cast<uint4*>(apref)[lane] = make_uint4(before, before + c0, before + c0 + c1, before + c0 + c1 + c2);
            __syncwarp();
            const int K0 = min(total, 32);
            if (lane < K0) {
                int lo_ = 0, hi_ = 127;                                       // last word whose prefix is <= lane
                while (lo_ < hi_) {
                    const int mid = (lo_ + hi_ + 1) >> 1;
                    if ((int)apref[mid] <= lane) lo_ = mid; else hi_ = mid - 1;
                }
                cand[lane] = lo_ * 32 + __fns(alive[lo_], 0, lane - (int)apref[lo_] + 1);
            }
            covm[lane] = 0u;
            if (lane == 0) s_K = K0;
        }
        __syncthreads(); GNMS_E2_T(1);
        int K = s_K;
        if (K == 0) break;
        GNMS_E2_C(8, 1);
        // ---- B: candidate x earlier candidate
        if (warp < K) {
            bool hit = false;
            if (lane < warp) {
                const int pi = cand[warp], pj = cand[lane];
                const RecT mine = record(pi), other = record(pj);
                bool unsafe = is_bad(pi) || is_bad(pj);
                float v = RecOf<kSrc>::template fast<kGen, kAffine>(mine, other, unsafe);
                if (__builtin_expect(unsafe, 0)) v = RecOf<kSrc>::template exact<kGen, kAffine>(mine, other);
                hit = !(v <= thr);                                             // NaN leaves the pool too (:249-250)
                covv[warp * 32 + lane] = v;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) covm[warp] = bal;
        }
        __syncthreads(); GNMS_E2_T(2);
        // ---- C: resolve the candidates in score order (dependency rounds, usually 2-3)
        if (warp == 0) {
            const uint32_t row = lane < K ? covm[lane] : 0u;
            uint32_t undecided = K == 32 ? 0xffffffffu : ((1u << K) - 1u), leaders = 0u;
            while (undecided) {
                const bool mine = (undecided >> lane) & 1u;
                const bool isl = mine && (row & (undecided | leaders)) == 0u;
                const bool iss = mine && (row & leaders) != 0u;
                const uint32_t nl = __ballot_sync(0xffffffffu, isl), ns = __ballot_sync(0xffffffffu, iss);
                leaders |= nl;
                undecided &= ~(nl | ns);
            }
            if (lane < K) {
                const int c = cand[lane];
                if ((leaders >> lane) & 1u) {
                    const int q = __popc(leaders & ((1u << lane) - 1u));
                    lpos[q] = c; lcidx[q] = lane;
                    float lo[3], hi[3];
                    AxesOf<kSrc>::get(record(c), lo, hi);
                    const bool lbad = is_bad(c);
#pragma unroll
                    for (int ax = 0; ax < kAxes; ++ax) ltab[q * 3 + ax] = make_float4(lo[ax], hi[ax], __fsub_rn(hi[ax], lo[ax]), lbad ? 1.f : 0.f);
                    // reach tables: a box can pass the gap test on axis a only if its lower edge lies in
                    // [lo - R - emax - margin, hi + R + margin], R = c (extent + emax); bins of those two edges, inclusive
                    uint32_t lb = 0u;
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        int b0 = 0, b1 = kE2Bins - 1;
                        if (!lbad) {
                            const float emax = s_binp[a * 4 + 2], mgn = s_binp[a * 4 + 3];
                            const float R = __fmul_rn(cc, __fadd_rn(__fsub_rn(hi[a], lo[a]), emax));
                            b0 = e2_bin(__fsub_rn(__fsub_rn(__fsub_rn(lo[a], R), emax), mgn), s_binp[a * 4], s_binp[a * 4 + 1]);
                            b1 = e2_bin(__fadd_rn(__fadd_rn(hi[a], R), mgn), s_binp[a * 4], s_binp[a * 4 + 1]);
                        }
                        lb |= (uint32_t)(b0 | (b1 << 8)) << (16 * a);
                    }
                    lbins[q] = lb;
                    atomicOr(&leaderb[c >> 5], 1u << (c & 31));
                } else {
                    const int jf = __ffs(row & leaders) - 1;
                    fsup[c] = cand[jf];
                    vout[c] = covv[lane * 32 + jf];
                }
                atomicAnd(&alive[c >> 5], ~(1u << (c & 31)));
            }
            if (lane == 0) { s_m = __popc(leaders); s_leaders = leaders; s_qn = 0; s_over = 0; }
        }
        __syncthreads();
        {
            // reach tables: 8 threads per (axis, bin), 4 leader slots each, OR-reduced over the 8 neighbouring lanes
            const int a = tid >> 9, bin = (tid >> 3) & (kE2Bins - 1), part = tid & 7, m0 = s_m;
            uint32_t bits = 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = part * 4 + u;
                if (q < m0) {
                    const uint32_t lb = lbins[q] >> (16 * a);
                    if (bin >= (int)(lb & 0xffu) && bin <= (int)((lb >> 8) & 0xffu)) bits |= 1u << q;
                }
            }
            bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
            bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
            bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
            if (part == 0) reach[a * kE2Bins + bin] = bits;
        }
        __syncthreads(); GNMS_E2_T(3);
        GNMS_E2_C(9, s_m);
        // ---- D: pool x the leaders that can reach it
        while (true) {
            const int m = s_m;
            const uint32_t mall = m == 32 ? 0xffffffffu : ((1u << m) - 1u);
            // (the positions of a thread are swept first, all of them, then pushed with ONE warp scan and one atomic)
            uint32_t nearm[kE2MaxK];
#pragma unroll
            for (int k = 0; k < kE2MaxK; ++k) {
                nearm[k] = 0u;
                if (k < KK && ((alive[k * 32 + warp] >> lane) & 1u)) {
                    const int pos = k * kE2Threads + tid;
                    const uint32_t bb = bins[pos];
                    const bool mybad = bb == 0xffffu;
                    uint32_t cm = mall;
                    if (!mybad) cm &= reach[bb & 0xffu] & reach[kE2Bins + (bb >> 8)];
                    float lo[3], hi[3];
                    if (cm) AxesOf<kSrc>::get(record(pos), lo, hi);
                    while (cm) {
                        const int q = __ffs(cm) - 1;
                        cm &= cm - 1u;
                        bool nr = true;
                        if (ltab[q * 3].w == 0.f && !mybad) {
#pragma unroll
                            for (int ax = 0; ax < kAxes; ++ax) {
                                const float4 t = ltab[q * 3 + ax];
                                if (fmaxf(__fsub_rn(lo[ax], t.y), __fsub_rn(t.x, hi[ax])) > __fmul_rn(cc, __fadd_rn(t.z, __fsub_rn(hi[ax], lo[ax])))) nr = false;
                            }
                        }
                        nearm[k] |= (uint32_t)nr << q;
                    }
                }
            }
            {
                // one scan for the (up to four) positions of every lane: their counts travel as 16-bit fields of one word, so
                // that the queue keeps the pairs of 32 consecutive positions together (the evaluation reads their records)
                unsigned long long packed = 0ull;
#pragma unroll
                for (int k = 0; k < kE2MaxK; ++k) packed |= (unsigned long long)__popc(nearm[k]) << (16 * k);
                unsigned long long incl = packed;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                const unsigned long long totp = __shfl_sync(0xffffffffu, incl, 31);
                const int tot = (int)((totp & 0xffffu) + ((totp >> 16) & 0xffffu) + ((totp >> 32) & 0xffffu) + (totp >> 48));
                if (tot != 0) {                                                // (warp-uniform)
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_qn, tot);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base + tot > kE2Queue) {
                        if (lane == 0) s_over = 1;
                    } else {
                        const unsigned long long excl = incl - packed;
#pragma unroll
                        for (int k = 0; k < kE2MaxK; ++k) {
                            int at = base + (int)((excl >> (16 * k)) & 0xffffu);
                            uint32_t nm = nearm[k];
                            while (nm) {
                                const int q = __ffs(nm) - 1;
                                nm &= nm - 1u;
                                queue[at++] = ((uint32_t)(k * kE2Threads + tid) << 5) | (uint32_t)q;
                            }
                            base += (int)((totp >> (16 * k)) & 0xffffu);
                        }
                    }
                }
            }
            __syncthreads();
            if (!s_over) break;
            GNMS_E2_C(11, 1);
            __syncthreads();                                                   // (everybody has seen the flag)
            if (warp == 0) {
                // keep the first quarter of this step's leaders; later candidates go back to the pool unless one of the kept
                // leaders suppresses them (then their first suppressor is already right: it is the lowest leader of their row)
                const uint32_t leaders = s_leaders;
                const int keep = max(1, m >> 2);
                uint32_t rest = leaders;
                for (int t = 0; t < keep; ++t) rest &= rest - 1u;
                const int icut = __ffs(rest) - 1;                              // candidate index of the first dropped leader
                const uint32_t kept = leaders & ((1u << icut) - 1u);
                if (lane >= icut && lane < K) {
                    const int c = cand[lane];
                    const bool was_leader = (leaders >> lane) & 1u;
                    if (was_leader || (covm[lane] & kept) == 0u) {
                        atomicOr(&alive[c >> 5], 1u << (c & 31));
                        fsup[c] = INT_MAX;
                        if (was_leader) atomicAnd(&leaderb[c >> 5], ~(1u << (c & 31)));
                    }
                }
                __syncwarp();                                                  // (every lane has read s_leaders)
                if (lane == 0) { s_m = keep; s_leaders = kept; s_K = icut; s_qn = 0; s_over = 0; }
            }
            __syncthreads();
            K = s_K;
        }
        // ---- E: the queued pairs, one per thread and round
        GNMS_E2_T(4);
        const int qn = s_qn;
        GNMS_E2_C(10, qn);
        float vreg[kE2PerThread];
#pragma unroll
        for (int i = 0; i < kE2PerThread; ++i) {
            const int e = i * kE2Threads + tid;
            vreg[i] = 0.f;
            if (e < qn) {
                const uint32_t ent = queue[e];
                const int pos = (int)(ent >> 5), l = lpos[ent & 31u];
                const RecT mine = record(pos), ldr = record(l);
                bool unsafe = is_bad(pos) || is_bad(l);
                float v = RecOf<kSrc>::template fast<kGen, kAffine>(mine, ldr, unsafe);
                if (__builtin_expect(unsafe, 0)) v = RecOf<kSrc>::template exact<kGen, kAffine>(mine, ldr);
                if (!(v <= thr)) atomicMin(&fsup[pos], l);
                vreg[i] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kE2PerThread; ++i) {
            const int e = i * kE2Threads + tid;
            if (e < qn) {
                const uint32_t ent = queue[e];
                const int pos = (int)(ent >> 5);
                if (fsup[pos] == lpos[ent & 31u] && !(vreg[i] <= thr)) vout[pos] = vreg[i];
            }
        }
        GNMS_E2_T(5);
        // ---- F: boxes with a suppressor leave the pool
        for (int k = 0; k < KK; ++k) {
            const int pos = k * kE2Threads + tid;
            const bool a = ((alive[k * 32 + warp] >> lane) & 1u) && fsup[pos] == INT_MAX;
            const unsigned bal = __ballot_sync(0xffffffffu, a);
            __syncwarp();                                                      // (every lane has read the word lane 0 rewrites)
            if (lane == 0) alive[k * 32 + warp] = bal;
        }
        __syncthreads(); GNMS_E2_T(6);
    }
    if (A.fuse_chain) {
        // the chain continues in place: leader bitset and first suppressors stay where they are, everything else it needs is
        // carved out of the record area (dead now).  vout[] (global, written above) is read back by the same CTA: the barrier
        // that ended the loop ordered those writes.
        static_assert(kE2Threads == kChainThreads, "the chain runs with the election's thread block");
        const int NWc = (N + 31) / 32, NWa = (NWc + 3) & ~3, Na = (N + 3) & ~3;
        unsigned char* base = s_e2;
        ChainSmem S;
        S.removed = reinterpret_cast<uint32_t*>(base);
        S.tmpbits = S.removed + NWa;
        S.list = reinterpret_cast<int32_t*>(S.tmpbits + NWa);
        S.wcol = reinterpret_cast<uint32_t*>(S.list + Na);
        size_t P2 = 1;
        while ((int)P2 < N) P2 <<= 1;
        const size_t wbytes = P2 * 8 > (size_t)kWin * NWc * 4 ? P2 * 8 : (size_t)kWin * NWc * 4;
        S.ss_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(S.wcol) + wbytes);
        S.leader = leaderb;
        S.fsup = fsup;
        for (int pos = n + tid; pos < N; pos += kE2Threads) fsup[pos] = INT_MAX;      // (dead tail, as the chain's own loader leaves it)
        __syncthreads();
        chain_body(A.chain, b, S, true);
        return;
    }
    int32_t* dfsup = reinterpret_cast<int32_t*>(w + L.dfsup);
    uint32_t* dleader = reinterpret_cast<uint32_t*>(w + L.dleader);
    for (int pos = tid; pos < n; pos += kE2Threads) dfsup[pos] = fsup[pos];
    GNMS_E2_T(7);
    if (tid < (N + 31) / 32) dleader[tid] = tid < 128 ? leaderb[tid] : 0u;
    if (tid == 0) *reinterpret_cast<int32_t*>(w + L.dflag) = 0;
}


}  // namespace gnms
#include "solve.cuh"
namespace gnms {

// ------------------------------------------------------------------------------------------ 4. backward (mode GROUP_MASK)
struct BwdArgs {
    int N, batch;
    const int32_t* n_per_image;
    gnms_params p;
    const float* grad_prob;
    const int32_t* order;
    const float* sorted_scores;
    const int32_t* lead;
    const float* pval;
    const float* dpval;
    const float* pre;
    const int32_t* slot;
    float* grad_scores;
    float* grad_iou;
    int64_t ld_gi;
};

__global__ void __launch_bounds__(kChainThreads) backward_mask_kernel(BwdArgs A) {
    extern __shared__ __align__(16) float ds[];
    const int b = blockIdx.x, N = A.N, tid = threadIdx.x;
    const int n = A.n_per_image ? min(A.n_per_image[b], N) : N;
    const size_t o = (size_t)b * N;
    const float vthr = A.p.valid_box_prob_threshold;
    // g~ = g * 1[0 <= pre <= 1] * 1[in a group] (* 1[r >= valid_thr] when the returned vector is the thresholded one)
    for (int pos = tid; pos < n; pos += kChainThreads) {
        const float pre = A.pre[o + pos];
        const int ld_ = A.lead[o + pos];
        float g = A.p.sorted_output ? A.grad_prob[o + A.slot[o + pos]] : A.grad_prob[o + pos];
        bool pass = (pre >= 0.f) && (pre <= 1.f) && (ld_ >= 0);
        if (A.p.sorted_output || A.p.thresholded_output) {
            float r = fminf(fmaxf(pre, 0.f), 1.f);
            pass = pass && (r >= vthr);
        }
        ds[pos] = pass ? g : 0.f;
    }
    __syncthreads();
    // leaders: ds_l = g~_l - sum_m p_ml g~_m.  The sum must not depend on the order in which the members' atomics land
    // (deterministic run to run).  Shared-memory atomics on doubles (and on floats, and on 64-bit integers) are compare-and-swap
    // loops on this part -- a hundred members fighting over their leader's word took most of the kernel -- only 32-bit integer
    // adds are native.  So the products are accumulated in FIXED POINT: scaled by a power of two taken from the image's largest
    // product (|term| < 2^47), rounded to an integer and added as four 16-bit digits into four 32-bit counters per leader
    // (up to 2^13 members cannot overflow them; integer adds commute, so the sum is exact and order-free).  Terms keep 47 bits
    // below the largest one: the same guarantee as an fp64 sum up to a span of 2^23.  Infinite / NaN products take the fp64 path.
    float* red = reinterpret_cast<float*>(ds + ((N + 3) & ~3));                    // [32] block reduction
    uint32_t* dig = reinterpret_cast<uint32_t*>(red + 32);                         // [4][N]
    double* acc = reinterpret_cast<double*>(dig);                                  // (fp64 path: [N], aliases the digits)
    float mx = 0.f;
    for (int pos = tid; pos < n; pos += kChainThreads) {
        const int ld_ = A.lead[o + pos];
        if (ld_ >= 0 && ld_ != pos && ds[pos] != 0.f) {
            const float pr = fabsf(__fmul_rn(A.pval[o + pos], ds[pos]));
            mx = pr > mx || pr != pr ? pr : mx;                                    // (NaN sticks)
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const float t = __shfl_xor_sync(0xffffffffu, mx, d); mx = (t > mx || t != t) ? t : mx; }
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    for (int i = tid; i < 4 * N; i += kChainThreads) dig[i] = 0u;
    __syncthreads();
    mx = red[tid & 31];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const float t = __shfl_xor_sync(0xffffffffu, mx, d); mx = (t > mx || t != t) ? t : mx; }
    const bool fixed = mx < INFINITY;                                              // (false for inf and NaN; block-uniform)
    int ex = 0;
    if (fixed && mx > 0.f) (void)frexpf(mx, &ex);                                  // mx < 2^ex
    const int sh = 47 - ex;
    for (int pos = tid; pos < n; pos += kChainThreads) {
        const int ld_ = A.lead[o + pos];
        if (ld_ >= 0 && ld_ != pos) {
            const float gt = ds[pos];
            if (gt != 0.f) {
                const float pr = __fmul_rn(A.pval[o + pos], gt);
                if (fixed) {
                    const long long t = __double2ll_rn(ldexp((double)pr, sh));
                    atomicAdd(&dig[ld_], (uint32_t)(t & 0xffff));
                    atomicAdd(&dig[N + ld_], (uint32_t)((t >> 16) & 0xffff));
                    atomicAdd(&dig[2 * N + ld_], (uint32_t)((t >> 32) & 0xffff));
                    atomicAdd(&dig[3 * N + ld_], (uint32_t)(int32_t)(t >> 48));    // signed top digit
                } else {
                    atomicAdd(&acc[ld_], (double)pr);
                }
                if (A.grad_iou) {
                    const int64_t gi = (int64_t)b * N * A.ld_gi + (int64_t)A.order[o + pos] * A.ld_gi + A.order[o + ld_];
                    A.grad_iou[gi] = __fmul_rn(-__fmul_rn(A.sorted_scores[o + ld_], gt), A.dpval[o + pos]);
                }
            }
        }
    }
    __syncthreads();
    for (int pos = tid; pos < n; pos += kChainThreads) {
        double sum;
        if (fixed) {
            const long long v = (long long)dig[pos] + ((long long)dig[N + pos] << 16) + ((long long)dig[2 * N + pos] << 32) +
                                (long long)((unsigned long long)(long long)(int32_t)dig[3 * N + pos] << 48);
            sum = ldexp((double)v, -sh);
        } else {
            sum = acc[pos];
        }
        A.grad_scores[o + A.order[o + pos]] = __fsub_rn(ds[pos], (float)sum);
    }
    if (n < N) {
        // padded boxes (n_per_image < N): their input slots get zero gradient
        for (int i = tid; i < N; i += kChainThreads) {
            // slots not covered by order[0..n) are exactly the padded inputs n..N-1 of this image
            if (i >= n) A.grad_scores[o + i] = 0.f;
        }
    }
}

static size_t chain_smem_bytes(int N) { return chain_wcol_end(N) + (((size_t)N + 3) & ~(size_t)3) * 4 + 16; }

// One-time function attributes per device.  Idempotent, so a race between two first callers is harmless; the flag is
// atomic so that the store is never torn or reordered before the attribute calls.
static int configure_once() {
    static std::atomic<bool> done_dev[64];
    int dev = 0;
    GNMS_CUDA_TRY(cudaGetDevice(&dev));
    std::atomic<bool>& done = done_dev[dev & 63];
    if (done.load(std::memory_order_acquire)) return 0;
    GNMS_CUDA_TRY(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)chain_smem_bytes(GNMS_MAX_BOXES)));
    GNMS_CUDA_TRY(cudaFuncSetAttribute(backward_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * GNMS_MAX_BOXES + 256));
    GNMS_CUDA_TRY(cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem_bytes(GNMS_MAX_BOXES)));
    done.store(true, std::memory_order_release);
    return 0;
}

// Tile work list of the culled pass: one int per surviving tile (image << 16 | I << 8 | J), count in front.
static size_t tile_list_bytes(int N, int batch) {
    const size_t nt = (size_t)((N + kTT - 1) / kTT);
    return align_up(256 + (size_t)batch * (nt * (nt + 1) / 2) * 4);
}
static int32_t* tile_list_ptr(void* workspace, int N, int batch) {
    return reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) + ws_layout(N).total * (size_t)batch +
                                      align_up((size_t)batch * N * 4) + 4 * align_up((size_t)N * 4));
}

static size_t rank_smem_bytes(int N) { return (size_t)((N + 15) & ~15) * 4; }

template <int kSrc, int kCmp>
static void launch_mask_boxes(dim3 grid, cudaStream_t s, int generalized, int affine, int N, const int32_t* npi,
                              char* ws, size_t stride, float thr, float shift) {
    if (kSrc == kSrcBox3d) {
        if (generalized) {
            if (affine) mask_boxes_kernel<kSrcBox3d, true, true, kCmp><<<grid, kMaskThreads, 0, s>>>(N, npi, ws, stride, thr, shift);
            else mask_boxes_kernel<kSrcBox3d, true, false, kCmp><<<grid, kMaskThreads, 0, s>>>(N, npi, ws, stride, thr, shift);
        } else {
            if (affine) mask_boxes_kernel<kSrcBox3d, false, true, kCmp><<<grid, kMaskThreads, 0, s>>>(N, npi, ws, stride, thr, shift);
            else mask_boxes_kernel<kSrcBox3d, false, false, kCmp><<<grid, kMaskThreads, 0, s>>>(N, npi, ws, stride, thr, shift);
        }
    } else {
        mask_boxes_kernel<kSrc, false, false, kCmp><<<grid, kMaskThreads, 0, s>>>(N, npi, ws, stride, thr, shift);
    }
}

static int check_common(int N, int batch, const gnms_params* p) {
    if (N < 0 || batch < 0 || !p) return GNMS_E_BADARG;
    if (N > GNMS_MAX_BOXES) return GNMS_E_TOOLARGE;
    if (p->pruning_method < 0 || p->pruning_method > 2 || p->mode < 0 || p->mode > GNMS_MODE_GROUP_MASK_INPUT_TRIL || p->group_size < 0)
        return GNMS_E_BADARG;
    return 0;
}

}  // namespace gnms

using namespace gnms;

extern "C" int gnms_version(void) { return GNMS_VERSION; }

#ifdef GNMS_DEBUG
// phase clock of chain_kernel (see GNMS_PHASE); only in -DGNMS_DEBUG builds, not part of the public header
extern "C" int gnms_debug_chain_clock(int on) {
    return (int)cudaMemcpyToSymbol(g_chain_clk_on, &on, sizeof(int));
}
extern "C" int gnms_debug_elect2_clock_read(long long* out16, int reset) {
    GNMS_CUDA_TRY(cudaMemcpyFromSymbol(out16, gnms::g_e2_clk, sizeof(long long) * 16));
    if (reset) { long long z[16] = {0}; GNMS_CUDA_TRY(cudaMemcpyToSymbol(gnms::g_e2_clk, z, sizeof(z))); }
    return 0;
}
extern "C" int gnms_debug_chain_clock_read(long long* out64) {
    return (int)cudaMemcpyFromSymbol(out64, g_chain_clk, sizeof(long long) * 64);
}
#endif

extern "C" const char* gnms_error_string(int rc) {
    if (rc == 0) return "success";
    if (rc > 0) return cudaGetErrorString((cudaError_t)rc);
    switch (rc) {
        case GNMS_E_BADARG: return "gnms: bad argument";
        case GNMS_E_TOOLARGE: return "gnms: more boxes per image than GNMS_MAX_BOXES";
        case GNMS_E_ALIGN: return "gnms: pointer not 16-byte aligned";
        case GNMS_E_UNSUPPORTED: return "gnms: unsupported configuration";
    }
    return "gnms: unknown error";
}

// Layout: [batch image slices][slot int32[batch*N]][4 scratch arrays of N words (get_groups / hard NMS)]
extern "C" size_t gnms_workspace_bytes(int N, int batch) {
    if (N <= 0 || batch <= 0) return 256;
    return ws_layout(N).total * (size_t)batch + align_up((size_t)batch * N * 4) + 4 * align_up((size_t)N * 4) +
           tile_list_bytes(N, batch);
}

// Tensor map of the [batch, N, N] fp32 overlap matrix for tile_tma_kernel: boxes of 32 rows x 32 columns (128 bytes per
// row), 128-byte swizzle.  The encoder is the driver's cuTensorMapEncodeTiled, looked up through the runtime so that the
// library does not link against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode_matrix_tmap(CUtensorMap* m, float* out, int N, int batch) {
    static std::atomic<EncodeTiledFn> cached{nullptr};
    EncodeTiledFn fn = cached.load(std::memory_order_acquire);
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        GNMS_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) return GNMS_E_UNSUPPORTED;
        fn = reinterpret_cast<EncodeTiledFn>(p);
        cached.store(fn, std::memory_order_release);
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)batch};
    const cuuint64_t gstr[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * (cuuint64_t)N * 4};
    const cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : GNMS_E_UNSUPPORTED;
}

// Overlap matrices only (no bits, no workspace): the matrix-only instantiation of the tile kernel.  Used by the batched
// overlap entry points of overlap.cu and by the forward when the caller asks for the matrix.
int gnms_launch_overlap_tiles(const float* boxes, int src, int generalized, int affine, int N, int batch, float* out,
                              const GnmsLaunchOpts& O, cudaStream_t s) {
    if (N <= 0 || batch <= 0) return 0;
    if (!boxes || !out) return GNMS_E_BADARG;
    TileArgs T = {};
    T.N = N; T.batch = batch; T.nt = gnms_div_up(N, kTT); T.tiles_per_image = T.nt * (T.nt + 1) / 2;
    T.inv_tpi = 1.0f / (float)T.tiles_per_image; T.inv_w = 1.0f / (float)(T.nt + 1);
    T.vec = ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) && (N % 4 == 0);
    T.boxes = boxes; T.out = out; T.thr = INFINITY;
    {                                                                  // 256 x 64 tiles
        constexpr int kq = 4;
        T.tiles_per_image = tall_tiles_per_image(N, kq);
        T.inv_tpi = 1.0f / (float)T.tiles_per_image;
        const long long tot = (long long)T.tiles_per_image * batch;
        if (tot > 0x7fffffffLL) return GNMS_E_TOOLARGE;
        int grid = tot < 148 * 4 ? (int)tot : 148 * 4;
        T.tiles_per_cta = O.tiles_per_cta > 0 ? (O.tiles_per_cta + kq - 1) / kq : 0;
        if (T.tiles_per_cta > 0) grid = (int)((tot + T.tiles_per_cta - 1) / T.tiles_per_cta);
        const bool scalar = (O.flags & GNMS_OPT_SCALAR_MATH) != 0;
        // TMA tensor stores need a 16-byte aligned matrix with rows that are multiples of 16 bytes (T.vec) and the driver's
        // tensor-map encoder; anything else takes the register-direct kernel (same bits).
        bool use_tma = O.matrix_kernel != GNMS_MATRIX_KERNEL_DIRECT && T.vec;
        CUtensorMap tmap;
        if (use_tma) {
            const int erc = encode_matrix_tmap(&tmap, out, N, batch);
            if (erc != 0) {
                if (O.matrix_kernel == GNMS_MATRIX_KERNEL_TMA) return erc;       // asked for explicitly: report
                use_tma = false;
            }
        }
        const size_t tma_smem = 1024 + kTmaStageBytes + (size_t)2 * (src == kSrcBox3d ? 8 : 4) * (256 + kTT) * 4 + 64;
#define GNMS_TALL(SRC, G, AF)                                                           \
    do {                                                                                \
        if (use_tma) {                                                                  \
            if (scalar) { GNMS_TMA_ATTR((tile_tma_kernel<SRC, G, AF, false>)); tile_tma_kernel<SRC, G, AF, false><<<grid, 128, tma_smem, s>>>(T, tmap); } \
            else { GNMS_TMA_ATTR((tile_tma_kernel<SRC, G, AF, true>)); tile_tma_kernel<SRC, G, AF, true><<<grid, 128, tma_smem, s>>>(T, tmap); } \
        }                                                                               \
        else if (scalar) tile_tall_kernel<SRC, G, AF, 4><<<grid, 128, 0, s>>>(T);       \
        else tile_tall_kernel<SRC, G, AF, 4, true><<<grid, 128, 0, s>>>(T);             \
    } while (0)
#define GNMS_TMA_ATTR(KERNEL)                                                                                       \
    do {                                                                                                            \
        static std::atomic<bool> attr_done[64];                                                                     \
        int dev_ = 0;                                                                                               \
        GNMS_CUDA_TRY(cudaGetDevice(&dev_));                                                                        \
        if (!attr_done[dev_ & 63].load(std::memory_order_acquire)) {                                                \
            GNMS_CUDA_TRY(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem)); \
            attr_done[dev_ & 63].store(true, std::memory_order_release);                                            \
        }                                                                                                           \
    } while (0)
        if (src == kSrcBox3d) {
            if (generalized) { if (affine) GNMS_TALL(kSrcBox3d, true, true); else GNMS_TALL(kSrcBox3d, true, false); }
            else { if (affine) GNMS_TALL(kSrcBox3d, false, true); else GNMS_TALL(kSrcBox3d, false, false); }
        } else {
            GNMS_TALL(kSrcBox2d, false, false);
        }
#undef GNMS_TMA_ATTR
#undef GNMS_TALL
        GNMS_LAUNCH_CHECK();
    }
    return 0;
}

static int run_forward(const float* scores, int src, const float* iou, int64_t ld, const float* boxes,
                       float* overlap_out, int generalized, int affine, int N, int batch, const int32_t* npi,
                       const gnms_params* p,
                       float* prob, int64_t* valid_idx, int64_t* invalid_idx, int32_t* counts, gnms_saved sv,
                       int32_t* slot, void* workspace, const gnms_launch_opts* opts, cudaStream_t s) {
    GnmsLaunchOpts O;
    int rc = gnms_resolve_opts(opts, &O);
    if (rc) return rc;
    const bool g_direct = O.election != GNMS_ELECT_MASK, g_split_matrix = !(O.flags & GNMS_OPT_ONE_PASS),
               g_tile_queue = !(O.flags & GNMS_OPT_INLINE_HITS);
    const unsigned g_stage_mask = O.stage_mask;
    const int g_rank_by_sort = O.rank_method == GNMS_RANK_SORT ? 1 : O.rank_method == GNMS_RANK_COUNT ? 0 : -1;
    rc = check_common(N, batch, p);
    if (rc) return rc;
    if (N == 0 || batch == 0) return 0;
    if (!scores || !prob || !valid_idx || !invalid_idx || !counts || !workspace || !sv.order || !sv.sorted_scores ||
        !sv.lead || !sv.pval || !sv.dpval || !sv.pre)
        return GNMS_E_BADARG;
    if (p->sorted_output && !slot) return GNMS_E_BADARG;
    rc = configure_once();
    if (rc) return rc;
    const WsLayout L = ws_layout(N);
    char* ws = reinterpret_cast<char*>(workspace);
    const int NW = (N + 31) / 32;
    int box_stride = src == kSrcBox2d ? 4 : 8;
    const bool tril_input = p->mode == GNMS_MODE_GROUP_MASK_INPUT_TRIL;
    const int mode = tril_input ? GNMS_MODE_GROUP_MASK : p->mode;
    const bool need_groups = mode != GNMS_MODE_NOGROUP;
    const bool tiles = need_groups && src != kSrcMatrix;              // fused overlap + mask tile kernel
    // matrix-free pass: spatial order + culling of tile pairs that provably hold no pair above the threshold
    float cull_c = -1.f;                                               // < 0: culling not applicable
    // The caller wants the overlap matrix: it comes from the matrix-only tile kernel (no bits, no ranks -- it depends on nothing
    // else in this call; TMA tensor stores where the matrix layout allows) and the rest of the call is the matrix-free path:
    // direct leader election where applicable, else the (spatially culled) bits-only tile pass; mode NOGROUP needs neither.
    // The one-pass kernel that writes matrix and bits together is kept behind GNMS_OPT_ONE_PASS for comparison.
    if (overlap_out && src != kSrcMatrix && g_split_matrix && batch < 32768) {
        if (g_stage_mask & 4) {
            rc = gnms_launch_overlap_tiles(boxes, src, generalized, affine, N, batch, overlap_out, O, s);
            if (rc) return rc;
        }
        overlap_out = nullptr;
    }
    if (src != kSrcMatrix && !overlap_out && need_groups && N <= 128 * kTT && batch < 32768 && p->nms_threshold >= 0.f) {
        const float thr = p->nms_threshold;
        if (src == kSrcBox2d) cull_c = 0.f;                            // disjoint -> IoU = 0 <= thr
        else if (!affine) cull_c = 0.f;                                // IoU3D = 0 or GIoU <= 0 <= thr
        else if (thr >= 0.5f) cull_c = 0.f;                            // 0.5 (1 + v) <= 0.5 <= thr
        else if (generalized && thr > 0.05f) cull_c = (1.f - 2.f * thr + 1e-4f) / (2.f * thr - 1e-4f);
    }
    const bool culled = cull_c >= 0.f;
    // direct leader election first (elect_kernel): the mask kernels below then only work for the images it gave up on
    // (modes GROUP_MASK and GROUP_NOMASK share the grouping; the per-group solves of the latter gather their overlaps on the fly)
    // (batched election -- elect2_kernel -- by default; GNMS_ELECT_DIRECT keeps the one-leader-per-step kernel for comparison)
    const bool direct2 = (O.election == GNMS_ELECT_AUTO || O.election == GNMS_ELECT_BATCHED) && culled && need_groups && N <= kElectMaxBoxes;
    const bool direct = !direct2 && g_direct && culled && need_groups && N <= kElectMaxBoxes;
    // rank by counting costs batch * N^2 compares over the whole chip, the per-image radix sort a constant ~10 us
    const bool by_sort = g_rank_by_sort == 1 || (g_rank_by_sort < 0 && (double)batch * N * N >= 10.0 * 4096 * 4096);
    if (g_stage_mask & 1) {
        const int zm = ((tiles && !direct && !direct2) || (src != kSrcMatrix && overlap_out)) ? 1 : 0;
        if (by_sort) {
            sort_kernel<<<batch + (zm ? 4 * batch : 1), kSortThreads, sort_smem_bytes(N), s>>>(scores, N, N, batch, npi, sv.order,
                                                                                              sv.sorted_scores, ws, L.total, zm);
            GNMS_LAUNCH_CHECK();
        }
        rank_kernel<<<dim3(gnms_div_up(N, kRankElems), batch), 256, by_sort ? 0 : rank_smem_bytes(N), s>>>(
            scores, 1, N, N, npi, sv.order, sv.sorted_scores, ws, L.total, src == kSrcMatrix ? nullptr : boxes, src,
            (int64_t)N * box_stride, 0.f, by_sort ? 2 : 0, zm, tile_list_ptr(workspace, N, batch));
    }
    GNMS_LAUNCH_CHECK();
    if (src != kSrcMatrix && !boxes) return GNMS_E_BADARG;
    SpatialArgs SA = {};
    SA.N = N; SA.batch = batch; SA.nt = gnms_div_up(N, kTT); SA.src = src; SA.n_per_image = npi; SA.boxes = boxes; SA.ws = ws;
    SA.ws_img_stride = L.total; SA.tile_list = tile_list_ptr(workspace, N, batch); SA.cull_c = cull_c; SA.defer_list = direct ? 1 : 0;
    if (direct) {
        if (g_stage_mask & 2) spatial_kernel<<<batch, 1024, 0, s>>>(SA);     // spatial order + group table, no tile list yet
        GNMS_LAUNCH_CHECK();
    }
    ChainArgs A = {};
    A.N = N; A.batch = batch; A.n_per_image = npi; A.ws = ws; A.ws_img_stride = L.total;
    A.src = src; A.iou = iou; A.ld = ld; A.iou_img_stride = (int64_t)N * ld;
    A.generalized = generalized; A.affine = affine; A.p = *p;
    A.order = sv.order; A.sorted_scores = sv.sorted_scores; A.prob = prob; A.valid_idx = valid_idx;
    A.invalid_idx = invalid_idx; A.counts = counts; A.lead = sv.lead; A.pval = sv.pval; A.dpval = sv.dpval;
    A.pre = sv.pre; A.slot = slot; A.direct = (direct || direct2) ? 1 : 0; A.have_v = (direct2 && (g_stage_mask & 32)) ? 1 : 0;
    A.p.mode = mode; A.tril_input = tril_input ? 1 : 0;
    // mode GROUP_MASK after the batched election: the chain runs at the end of elect2_kernel, out of the same shared memory
    // (one launch; the election's result never goes through global memory).  A stage mask that asks for one of the two alone
    // keeps them apart.
    const bool chain_fused = direct2 && mode == GNMS_MODE_GROUP_MASK && (g_stage_mask & 32) && (g_stage_mask & 16) && !(O.flags & GNMS_OPT_SPLIT_CHAIN);
    if (direct2 && (g_stage_mask & 32)) {
        Elect2Args E = {};
        E.N = N; E.batch = batch; E.n_per_image = npi; E.ws = ws; E.ws_img_stride = L.total; E.thr = p->nms_threshold; E.cull_c = cull_c; E.vout = sv.pval;
        E.fuse_chain = chain_fused ? 1 : 0; E.chain = A; E.chain.stage = 0;
        const size_t esm = elect2_smem_bytes(N);
#define GNMS_ELECT2(SRC, G, AF)                                                                                       \
    do {                                                                                                              \
        static std::atomic<bool> attr_done[64];                                                                       \
        int dev_ = 0;                                                                                                 \
        GNMS_CUDA_TRY(cudaGetDevice(&dev_));                                                                          \
        if (!attr_done[dev_ & 63].load(std::memory_order_acquire)) {                                                  \
            GNMS_CUDA_TRY(cudaFuncSetAttribute(elect2_kernel<SRC, G, AF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)elect2_smem_bytes(kElectMaxBoxes)));                              \
            attr_done[dev_ & 63].store(true, std::memory_order_release);                                              \
        }                                                                                                             \
        elect2_kernel<SRC, G, AF><<<batch, kE2Threads, esm, s>>>(E);                                                  \
    } while (0)
        if (src == kSrcBox3d) {
            if (generalized) { if (affine) GNMS_ELECT2(kSrcBox3d, true, true); else GNMS_ELECT2(kSrcBox3d, true, false); }
            else { if (affine) GNMS_ELECT2(kSrcBox3d, false, true); else GNMS_ELECT2(kSrcBox3d, false, false); }
        } else {
            GNMS_ELECT2(kSrcBox2d, false, false);
        }
#undef GNMS_ELECT2
        GNMS_LAUNCH_CHECK();
    }
    if (direct && (g_stage_mask & 32)) {
        ElectArgs E = {};
        E.N = N; E.batch = batch; E.n_per_image = npi; E.ws = ws; E.ws_img_stride = L.total; E.thr = p->nms_threshold; E.cull_c = cull_c;
        const size_t esm = elect_smem_bytes(N);
#define GNMS_ELECT(SRC, G, AF)                                                                                       \
    do {                                                                                                             \
        static std::atomic<bool> attr_done[64];                                                                      \
        int dev_ = 0;                                                                                                \
        GNMS_CUDA_TRY(cudaGetDevice(&dev_));                                                                         \
        if (!attr_done[dev_ & 63].load(std::memory_order_acquire)) {                                                 \
            GNMS_CUDA_TRY(cudaFuncSetAttribute(elect_kernel<SRC, G, AF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)elect_smem_bytes(kElectMaxBoxes)));                              \
            attr_done[dev_ & 63].store(true, std::memory_order_release);                                             \
        }                                                                                                            \
        elect_kernel<SRC, G, AF><<<batch, kElectThreads, esm, s>>>(E);                                                       \
    } while (0)
        if (src == kSrcBox3d) {
            if (generalized) { if (affine) GNMS_ELECT(kSrcBox3d, true, true); else GNMS_ELECT(kSrcBox3d, true, false); }
            else { if (affine) GNMS_ELECT(kSrcBox3d, false, true); else GNMS_ELECT(kSrcBox3d, false, false); }
        } else {
            GNMS_ELECT(kSrcBox2d, false, false);
        }
#undef GNMS_ELECT
        GNMS_LAUNCH_CHECK();
        // images the election gave up on: zero their suppression masks (the sort / rank kernels skipped that), list their tiles
        zero_failed_kernel<<<(batch * 8 < 296 ? batch * 8 : 296), 256, 0, s>>>(N, batch, ws, L.total);
        GNMS_LAUNCH_CHECK();
        list_failed_kernel<<<batch, 256, 0, s>>>(SA);
        GNMS_LAUNCH_CHECK();
    }
    if (src == kSrcMatrix && (!iou || ld < N)) return GNMS_E_BADARG;
    if (need_groups && src == kSrcMatrix) {
        dim3 grid(gnms_div_up(N, kMaskThreads * 4), NW, batch);
        bool vec = ((reinterpret_cast<uintptr_t>(iou) & 15u) == 0) && (ld % 4 == 0) && (((int64_t)N * ld) % 4 == 0);
        if (!(g_stage_mask & 4)) {}
        else if (vec) mask_matrix_kernel<true><<<grid, kMaskThreads, 0, s>>>(iou, ld, (int64_t)N * ld, N, npi, sv.order, ws, L.total, p->nms_threshold);
        else mask_matrix_kernel<false><<<grid, kMaskThreads, 0, s>>>(iou, ld, (int64_t)N * ld, N, npi, sv.order, ws, L.total, p->nms_threshold);
        GNMS_LAUNCH_CHECK();
    } else if (src != kSrcMatrix && (need_groups || overlap_out) && !direct2) {
        // (mode NOGROUP needs no mask; the tile kernel still runs if the caller wants the overlap matrix; the batched election
        //  never gives up, so nothing of this block follows it)
        TileArgs T = {};
        T.N = N; T.batch = batch; T.nt = gnms_div_up(N, kTT); T.tiles_per_image = T.nt * (T.nt + 1) / 2;
        T.inv_tpi = 1.0f / (float)T.tiles_per_image; T.inv_w = 1.0f / (float)(T.nt + 1);
        T.vec = overlap_out && ((reinterpret_cast<uintptr_t>(overlap_out) & 15u) == 0) && (N % 4 == 0);
        T.n_per_image = npi; T.boxes = boxes; T.ws = ws; T.ws_img_stride = L.total; T.out = overlap_out;
        T.thr = need_groups ? p->nms_threshold : INFINITY;             // no bits wanted: nothing is > +inf ... NaN aside
        const int total = T.tiles_per_image * batch;
        const int grid_out = total < 148 * 4 ? total : 148 * 4;        // persistent: 128-thread CTAs, 4 per SM
        const int grid_bits = total < 148 * 2 ? total : 148 * 2;       // matrix-free: 256-thread CTAs, 2 per SM
        T.tile_list = tile_list_ptr(workspace, N, batch);
        const bool ho = overlap_out != nullptr;
        if (culled && !direct) {
            if (g_stage_mask & 2) spatial_kernel<<<batch, 1024, 0, s>>>(SA);
            GNMS_LAUNCH_CHECK();
        }
#define GNMS_TILE(SRC, G, AF)                                                              \
    do {                                                                                   \
        if (ho && g_tile_queue) tile_kernel<SRC, G, AF, true, false, 2, true><<<grid_out, 128, 0, s>>>(T);   \
        else if (ho) tile_kernel<SRC, G, AF, true, false, 2, false><<<grid_out, 128, 0, s>>>(T);             \
        else if (culled) tile_kernel<SRC, G, AF, false, true, 1, false><<<grid_bits, 256, 0, s>>>(T);        \
        else tile_kernel<SRC, G, AF, false, false, 1, false><<<grid_bits, 256, 0, s>>>(T);                   \
    } while (0)
        if (!(g_stage_mask & 4)) {
        } else if (src == kSrcBox3d) {
            if (generalized) {
                if (affine) GNMS_TILE(kSrcBox3d, true, true);
                else GNMS_TILE(kSrcBox3d, true, false);
            } else {
                if (affine) GNMS_TILE(kSrcBox3d, false, true);
                else GNMS_TILE(kSrcBox3d, false, false);
            }
        } else {
            GNMS_TILE(kSrcBox2d, false, false);
        }
#undef GNMS_TILE
        GNMS_LAUNCH_CHECK();
        if (need_groups && (g_stage_mask & 8)) {
            const int he_items = gnms_div_up(N, 256) * batch;
            if (direct) has_earlier_failed_kernel<<<(he_items < 148 ? he_items : 148), 1024, 0, s>>>(N, batch, npi, ws, L.total);
            else has_earlier_kernel<<<dim3(gnms_div_up(N, 256), batch), 1024, 0, s>>>(N, npi, ws, L.total);
            GNMS_LAUNCH_CHECK();
        }
    }
    if (!(g_stage_mask & 16) || chain_fused) return 0;
    if (mode == GNMS_MODE_GROUP_MASK) {
        A.stage = 0;
        chain_kernel<<<batch, kChainThreads, chain_smem_bytes(N), s>>>(A);
        GNMS_LAUNCH_CHECK();
        return 0;
    }
    // inverse modes: [grouping ->] triangular solve -> lists
    if (mode == GNMS_MODE_GROUP_NOMASK) {
        A.stage = 1;
        chain_kernel<<<batch, kChainThreads, chain_smem_bytes(N), s>>>(A);
        GNMS_LAUNCH_CHECK();
    }
    SolveArgs S = {};
    S.N = N; S.batch = batch; S.mode = mode; S.n_per_image = npi; S.ws = ws; S.ws_img_stride = L.total;
    S.ov.src = src; S.ov.iou = iou; S.ov.ld = ld; S.ov.generalized = generalized; S.ov.affine = affine;
    S.iou_img_stride = (int64_t)N * ld; S.p = *p; S.sorted_scores = sv.sorted_scores; S.order = sv.order;
    S.lead = sv.lead; S.pre = sv.pre;
    solve_fwd_kernel<<<dim3(mode == GNMS_MODE_NOGROUP ? 1 : N, batch), kSolveThreads, ((size_t)N + 40) * 4, s>>>(S);
    GNMS_LAUNCH_CHECK();
    A.stage = 2;
    chain_kernel<<<batch, kChainThreads, chain_smem_bytes(N), s>>>(A);
    GNMS_LAUNCH_CHECK();
    return 0;
}

// `slot` (position of every sorted box in the sorted output) lives at the tail of the pre-array contract: callers
// that use sorted_output pass a [batch,N] int32 buffer through gnms_saved.dpval's neighbour -- to keep the ABI flat
// it is carried in the workspace instead (first bytes after the per-image slices).
static int32_t* slot_ptr(void* workspace, int N, int batch) {
    return reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) + ws_layout(N).total * (size_t)batch);
}

extern "C" int gnms_forward_ex_f32(const float* scores, const float* iou, int64_t ld, int N, int batch,
                                   const int32_t* n_per_image, const gnms_params* p, float* prob, int64_t* valid_idx,
                                   int64_t* invalid_idx, int32_t* counts, gnms_saved saved, void* workspace,
                                   const gnms_launch_opts* opts, void* stream) {
    return run_forward(scores, kSrcMatrix, iou, ld, nullptr, nullptr, 0, 0, N, batch, n_per_image, p, prob, valid_idx,
                       invalid_idx, counts, saved, workspace ? slot_ptr(workspace, N, batch) : nullptr, workspace, opts,
                       (cudaStream_t)stream);
}
extern "C" int gnms_forward_f32(const float* scores, const float* iou, int64_t ld, int N, int batch,
                                const int32_t* n_per_image, const gnms_params* p, float* prob, int64_t* valid_idx,
                                int64_t* invalid_idx, int32_t* counts, gnms_saved saved, void* workspace,
                                void* stream) {
    return gnms_forward_ex_f32(scores, iou, ld, N, batch, n_per_image, p, prob, valid_idx, invalid_idx, counts, saved, workspace,
                               nullptr, stream);
}

extern "C" int gnms_forward_boxes_ex_f32(const float* scores, const float* boxes, int box_kind, int generalized,
                                         int affine, int N, int batch, const int32_t* n_per_image, const gnms_params* p,
                                         float* overlap_out, float* prob, int64_t* valid_idx, int64_t* invalid_idx,
                                         int32_t* counts, gnms_saved saved, void* workspace, const gnms_launch_opts* opts,
                                         void* stream) {
    if (box_kind != GNMS_BOX_2D && box_kind != GNMS_BOX_3D_REC) return GNMS_E_BADARG;
    if (boxes && (reinterpret_cast<uintptr_t>(boxes) & 15u)) return GNMS_E_ALIGN;
    return run_forward(scores, box_kind == GNMS_BOX_2D ? kSrcBox2d : kSrcBox3d, nullptr, 0, boxes, overlap_out, generalized, affine,
                       N, batch, n_per_image, p, prob, valid_idx, invalid_idx, counts, saved,
                       workspace ? slot_ptr(workspace, N, batch) : nullptr, workspace, opts, (cudaStream_t)stream);
}
extern "C" int gnms_forward_boxes_f32(const float* scores, const float* boxes, int box_kind, int generalized,
                                      int affine, int N, int batch, const int32_t* n_per_image, const gnms_params* p,
                                      float* overlap_out, float* prob, int64_t* valid_idx, int64_t* invalid_idx,
                                      int32_t* counts, gnms_saved saved, void* workspace, void* stream) {
    return gnms_forward_boxes_ex_f32(scores, boxes, box_kind, generalized, affine, N, batch, n_per_image, p, overlap_out, prob,
                                     valid_idx, invalid_idx, counts, saved, workspace, nullptr, stream);
}

extern "C" int gnms_backward_f32(const float* grad_prob, const float* prob, const float* iou, int64_t ld, int N,
                                 int batch, const int32_t* n_per_image, const gnms_params* p, gnms_saved sv,
                                 float* grad_scores, float* grad_iou, int64_t ld_gi, void* workspace, void* stream) {
    (void)prob;
    int rc = check_common(N, batch, p);
    if (rc) return rc;
    if (N == 0 || batch == 0) return 0;
    if (!grad_prob || !grad_scores || !sv.order || !sv.sorted_scores || !sv.lead || !sv.pval || !sv.dpval || !sv.pre)
        return GNMS_E_BADARG;
    if (grad_iou && ld_gi < N) return GNMS_E_BADARG;
    if (p->sorted_output && !workspace) return GNMS_E_BADARG;
    rc = configure_once();
    if (rc) return rc;
    if (p->mode != GNMS_MODE_GROUP_MASK && p->mode != GNMS_MODE_GROUP_MASK_INPUT_TRIL) {      // (the latter saved pval = 0 where Phi is cut)
        if (!workspace) return GNMS_E_BADARG;                          // holds the group structure / box records
        const WsLayout L = ws_layout(N);
        GNMS_CUDA_TRY(cudaMemsetAsync(grad_scores, 0, (size_t)batch * N * sizeof(float), (cudaStream_t)stream));
        SolveArgs S = {};
        S.N = N; S.batch = batch; S.mode = p->mode; S.n_per_image = n_per_image;
        S.ws = reinterpret_cast<char*>(workspace); S.ws_img_stride = L.total;
        S.ov.src = kSrcMatrix; S.ov.iou = iou; S.ov.ld = ld; S.iou_img_stride = (int64_t)N * ld; S.p = *p;
        S.sorted_scores = sv.sorted_scores; S.order = sv.order; S.lead = sv.lead; S.pre = sv.pre;
        S.grad_prob = grad_prob; S.slot = slot_ptr(workspace, N, batch); S.grad_scores = grad_scores;
        S.grad_iou = grad_iou; S.ld_gi = ld_gi;
        solve_bwd_kernel<<<dim3(p->mode == GNMS_MODE_NOGROUP ? 1 : N, batch), kSolveThreads, ((size_t)N + 40) * 4,
                           (cudaStream_t)stream>>>(S);
        GNMS_LAUNCH_CHECK();
        return 0;
    }
    BwdArgs A = {};
    A.N = N; A.batch = batch; A.n_per_image = n_per_image; A.p = *p; A.grad_prob = grad_prob; A.order = sv.order;
    A.sorted_scores = sv.sorted_scores; A.lead = sv.lead; A.pval = sv.pval; A.dpval = sv.dpval; A.pre = sv.pre;
    A.slot = workspace ? slot_ptr(workspace, N, batch) : nullptr;
    A.grad_scores = grad_scores; A.grad_iou = grad_iou; A.ld_gi = ld_gi;
    backward_mask_kernel<<<batch, kChainThreads, (size_t)N * 20 + 256, (cudaStream_t)stream>>>(A);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_get_groups_f32(const float* scores, const float* iou, int64_t ld, int N, float group_threshold,
                                   int group_size, int32_t* group_id, int32_t* group_rank, int32_t* n_groups,
                                   void* workspace, void* stream) {
    if (N < 0 || group_size < 0) return GNMS_E_BADARG;
    if (N > GNMS_MAX_BOXES) return GNMS_E_TOOLARGE;
    if (N == 0) return 0;
    if (!scores || !iou || ld < N || !group_id || !group_rank || !n_groups || !workspace) return GNMS_E_BADARG;
    int rc = configure_once();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const WsLayout L = ws_layout(N);
    char* ws = reinterpret_cast<char*>(workspace);
    // order / sorted scores / pval / dpval scratch live after the image slice (+ the slot array)
    char* extra = ws + L.total + align_up((size_t)N * 4);
    int32_t* order = reinterpret_cast<int32_t*>(extra);
    float* ss = reinterpret_cast<float*>(extra + align_up((size_t)N * 4));
    float* pv = reinterpret_cast<float*>(extra + 2 * align_up((size_t)N * 4));
    float* dpv = reinterpret_cast<float*>(extra + 3 * align_up((size_t)N * 4));
    const int NW = (N + 31) / 32;
    rank_kernel<<<dim3(gnms_div_up(N, kRankElems), 1), 256, rank_smem_bytes(N), s>>>(scores, 1, N, N, nullptr, order, ss, ws, L.total,
                                                                                    nullptr, kSrcMatrix, 0, 0.f, 0, 0, nullptr);
    GNMS_LAUNCH_CHECK();
    dim3 grid(gnms_div_up(N, kMaskThreads * 4), NW, 1);
    bool vec = ((reinterpret_cast<uintptr_t>(iou) & 15u) == 0) && (ld % 4 == 0);
    if (vec) mask_matrix_kernel<true><<<grid, kMaskThreads, 0, s>>>(iou, ld, 0, N, nullptr, order, ws, L.total, group_threshold);
    else mask_matrix_kernel<false><<<grid, kMaskThreads, 0, s>>>(iou, ld, 0, N, nullptr, order, ws, L.total, group_threshold);
    GNMS_LAUNCH_CHECK();
    ChainArgs A = {};
    A.N = N; A.batch = 1; A.ws = ws; A.ws_img_stride = L.total; A.src = kSrcMatrix; A.iou = iou; A.ld = ld;
    A.p.nms_threshold = group_threshold; A.p.group_size = group_size; A.p.pruning_method = GNMS_PRUNE_LINEAR;
    A.order = order; A.sorted_scores = ss; A.pval = pv; A.dpval = dpv;
    A.group_id = group_id; A.group_rank = group_rank; A.n_groups = n_groups;
    chain_kernel<<<1, kChainThreads, chain_smem_bytes(N), s>>>(A);
    GNMS_LAUNCH_CHECK();
    return 0;
}

static int hard_nms_impl(const float* dets, int N, float thresh, float shift, int cmp, int32_t* keep,
                         int32_t* n_keep, void* workspace, void* stream, int presorted) {
    if (N < 0 || cmp < 0 || cmp > 2) return GNMS_E_BADARG;
    if (N > GNMS_MAX_BOXES) return GNMS_E_TOOLARGE;
    if (!n_keep) return GNMS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 0) return (int)cudaMemsetAsync(n_keep, 0, sizeof(int32_t), s);
    if (!dets || !keep || !workspace) return GNMS_E_BADARG;
    int rc = configure_once();
    if (rc) return rc;
    const WsLayout L = ws_layout(N);
    char* ws = reinterpret_cast<char*>(workspace);
    char* extra = ws + L.total + align_up((size_t)N * 4);
    int32_t* order = reinterpret_cast<int32_t*>(extra);
    float* ss = reinterpret_cast<float*>(extra + align_up((size_t)N * 4));
    const int NW = (N + 31) / 32;
    rank_kernel<<<dim3(gnms_div_up(N, kRankElems), 1), 256, rank_smem_bytes(N), s>>>(dets + 4, 5, 0, N, nullptr, order, ss, ws, L.total,
                                                                                    dets, kSrcBoxShift, 0, shift, presorted, 0, nullptr);
    GNMS_LAUNCH_CHECK();
    dim3 grid(gnms_div_up(N, kMaskThreads), NW, 1);
    if (cmp == GNMS_CMP_GT) launch_mask_boxes<kSrcBoxShift, GNMS_CMP_GT>(grid, s, 0, 0, N, nullptr, ws, L.total, thresh, shift);
    else if (cmp == GNMS_CMP_GE) launch_mask_boxes<kSrcBoxShift, GNMS_CMP_GE>(grid, s, 0, 0, N, nullptr, ws, L.total, thresh, shift);
    else launch_mask_boxes<kSrcBoxShift, GNMS_CMP_NLE>(grid, s, 0, 0, N, nullptr, ws, L.total, thresh, shift);
    GNMS_LAUNCH_CHECK();
    ChainArgs A = {};
    A.N = N; A.batch = 1; A.ws = ws; A.ws_img_stride = L.total; A.src = kSrcBoxShift;
    A.order = order; A.sorted_scores = ss; A.keep = keep; A.n_keep = n_keep; A.leaders_only = 1;
    chain_kernel<<<1, kChainThreads, chain_smem_bytes(N), s>>>(A);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_hard_nms_f32(const float* dets, int N, float thresh, float shift, int cmp, int32_t* keep,
                                 int32_t* n_keep, void* workspace, void* stream) {
    return hard_nms_impl(dets, N, thresh, shift, cmp, keep, n_keep, workspace, stream, 0);
}

// Drop-in for `_nms` (lib/nms/gpu_nms.hpp:1-2, lib/nms/nms_kernel.cu:91-144): host pointers, boxes sorted by score.
extern "C" int gnms_nms_host(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                             float nms_overlap_thresh, int device_id) {
    if (!num_out || boxes_num < 0) return GNMS_E_BADARG;
    if (boxes_num == 0) { *num_out = 0; return 0; }
    if (!keep_out || !boxes_host || boxes_dim != 5) return GNMS_E_BADARG;
    if (boxes_num > GNMS_MAX_BOXES) return GNMS_E_TOOLARGE;
    int cur = 0;
    GNMS_CUDA_TRY(cudaGetDevice(&cur));
    if (cur != device_id) GNMS_CUDA_TRY(cudaSetDevice(device_id));
    const int N = boxes_num;
    size_t wsb = gnms_workspace_bytes(N, 1);
    size_t detb = align_up((size_t)N * 5 * 4), keepb = align_up((size_t)N * 4);
    char* dev = nullptr;
    int rc = (int)cudaMalloc(&dev, detb + keepb + 256 + wsb);
    if (rc) return rc;
    float* d_dets = reinterpret_cast<float*>(dev);
    int32_t* d_keep = reinterpret_cast<int32_t*>(dev + detb);
    int32_t* d_n = reinterpret_cast<int32_t*>(dev + detb + keepb);
    void* d_ws = dev + detb + keepb + 256;
    cudaStream_t s = 0;
    rc = (int)cudaMemcpyAsync(d_dets, boxes_host, (size_t)N * 5 * 4, cudaMemcpyHostToDevice, s);
    if (!rc) rc = hard_nms_impl(d_dets, N, nms_overlap_thresh, 1.0f, GNMS_CMP_GT, d_keep, d_n, d_ws, s, 1);
    int nk = 0;
    if (!rc) rc = (int)cudaMemcpyAsync(&nk, d_n, 4, cudaMemcpyDeviceToHost, s);
    if (!rc) rc = (int)cudaStreamSynchronize(s);
    if (!rc && nk > 0) rc = (int)cudaMemcpy(keep_out, d_keep, (size_t)nk * 4, cudaMemcpyDeviceToHost);
    if (!rc) *num_out = nk;
    cudaFree(dev);
    if (cur != device_id) cudaSetDevice(cur);
    return rc;
}
