// Soft-NMS (lib/nms_others.py:6-116) and AP loss (lib/loss/aploss.py:14-97) for sm_100a.
// Both are small, inherently sequential-over-rounds algorithms: one CTA per problem, block-wide reductions.
#include "common.cuh"
#include "polygon.cuh"

namespace gnms {

constexpr int kT = 1024;

__device__ __forceinline__ void block_argmax(double v, int idx, double* s_val, int* s_idx, double& out_v, int& out_i) {
    // (value desc, index asc): the reference's `maxscore < boxes[pos,4]` scan keeps the FIRST maximum
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        double ov = __shfl_down_sync(0xffffffffu, v, d);
        int oi = __shfl_down_sync(0xffffffffu, idx, d);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { s_val[warp] = v; s_idx[warp] = idx; }
    __syncthreads();
    if (warp == 0) {
        v = lane < (kT / 32) ? s_val[lane] : -INFINITY;
        idx = lane < (kT / 32) ? s_idx[lane] : INT_MAX;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            double ov = __shfl_down_sync(0xffffffffu, v, d);
            int oi = __shfl_down_sync(0xffffffffu, idx, d);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        if (lane == 0) { s_val[0] = v; s_idx[0] = idx; }
    }
    __syncthreads();
    out_v = s_val[0];
    out_i = s_idx[0];
    __syncthreads();
}

// One CTA.  Arrangement-free restatement of the reference's in-place selection sort with swaps: each round picks
// the live, not yet selected box with the highest (decayed) score, decays every other live box by the overlap
// weight and drops boxes whose score falls below `threshold`.  With distinct scores the selected sequence is
// independent of the physical row order the reference maintains (ties: lowest original index first).
__global__ void __launch_bounds__(kT)
soft_nms_kernel(const double* __restrict__ dets, int N, double sigma, double Nt, double threshold, int method,
                double shift, int32_t* __restrict__ keep, double* __restrict__ keep_scores, int32_t* __restrict__ n_keep,
                double* __restrict__ score, int32_t* __restrict__ state) {
    __shared__ double s_val[kT / 32];
    __shared__ int s_idx[kT / 32];
    const int tid = threadIdx.x;
    for (int i = tid; i < N; i += kT) { score[i] = dets[(size_t)i * 5 + 4]; state[i] = 0; }   // 0 live, 1 selected, 2 dropped
    __syncthreads();
    int nsel = 0;
    for (int round = 0; round < N; ++round) {
        double bv = -INFINITY;
        int bi = INT_MAX;
        for (int i = tid; i < N; i += kT) {
            if (state[i] == 0) {
                double v = score[i];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
        }
        double mv; int mi;
        block_argmax(bv, bi, s_val, s_idx, mv, mi);
        if (mi == INT_MAX) break;                       // nothing live
        if (tid == 0) { keep[nsel] = mi; keep_scores[nsel] = mv; state[mi] = 1; }
        ++nsel;
        const double tx1 = dets[(size_t)mi * 5 + 0], ty1 = dets[(size_t)mi * 5 + 1];
        const double tx2 = dets[(size_t)mi * 5 + 2], ty2 = dets[(size_t)mi * 5 + 3];
        __syncthreads();
        for (int i = tid; i < N; i += kT) {
            if (state[i] != 0) continue;
            const double x1 = dets[(size_t)i * 5 + 0], y1 = dets[(size_t)i * 5 + 1];
            const double x2 = dets[(size_t)i * 5 + 2], y2 = dets[(size_t)i * 5 + 3];
            const double area = __dmul_rn(__dadd_rn(__dsub_rn(x2, x1), shift), __dadd_rn(__dsub_rn(y2, y1), shift));
            const double iw = __dadd_rn(__dsub_rn(fmin(tx2, x2), fmax(tx1, x1)), shift);       // :73
            if (iw > 0) {
                const double ih = __dadd_rn(__dsub_rn(fmin(ty2, y2), fmax(ty1, y1)), shift);   // :75
                if (ih > 0) {
                    const double ua = __dsub_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dsub_rn(tx2, tx1), shift),
                                                                    __dadd_rn(__dsub_rn(ty2, ty1), shift)), area),
                                                __dmul_rn(iw, ih));                              // :77
                    const double ov = __ddiv_rn(__dmul_rn(iw, ih), ua);                          // :78
                    double wgt;
                    if (method == 1) wgt = ov > Nt ? __dsub_rn(1.0, ov) : 1.0;                   // :80-84
                    else if (method == 2) wgt = exp(-__ddiv_rn(__dmul_rn(ov, ov), sigma));       // :86
                    else wgt = ov > Nt ? 0.0 : 1.0;                                              // :88-91
                    const double sc = __dmul_rn(wgt, score[i]);                                  // :93
                    score[i] = sc;
                    if (sc < threshold) state[i] = 2;                                            // :97
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) *n_keep = nsel;
}

// ------------------------------------------------------------------------------------------ AP loss
// One CTA.  lib/loss/aploss.py:16-81 with delta = 1 (:18).  Workspace: fg index list (int32[n]), a+b / scale
// per positive (float[2n]), prec per positive (float[n]).
__global__ void __launch_bounds__(kT)
aploss_kernel(const float* __restrict__ logits, const float* __restrict__ targets, int n, float* __restrict__ loss,
              float* __restrict__ grad, int32_t* __restrict__ fg_list, float* __restrict__ fg_ab,
              float* __restrict__ fg_prec, unsigned long long* __restrict__ fg_keys) {
    __shared__ int s_nfg;
    __shared__ float s_red[kT / 32];
    __shared__ float s_thr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_nfg = 0;
    for (int i = tid; i < n; i += kT) grad[i] = 0.f;
    __syncthreads();
    // positives (targets == 1), listed in index order is not required: they are sorted by logit below
    for (int i = tid; i < n; i += kT)
        if (targets[i] == 1.f) fg_list[atomicAdd(&s_nfg, 1)] = i;
    __syncthreads();
    const int nfg = s_nfg;
    if (nfg == 0) {                                                          // :27-29  (max(targets) <= 0)
        if (tid == 0) *loss = 0.f;
        return;
    }
    // sort positives by (logit asc, index asc): torch.sort(fg_logits) (:47); small list -> rank by counting
    for (int a = tid; a < nfg; a += kT) {
        const int ia = fg_list[a];
        const float va = logits[ia];
        int r = 0;
        for (int c = 0; c < nfg; ++c) {
            const int ic = fg_list[c];
            const float vc = logits[ic];
            r += (vc < va) || (vc == va && ic < ia);
        }
        fg_keys[r] = (unsigned long long)(uint32_t)ia;
    }
    __syncthreads();
    float mn = INFINITY;
    for (int a = tid; a < nfg; a += kT) mn = fminf(mn, logits[(int)fg_keys[a]]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    if (lane == 0) s_red[warp] = mn;
    __syncthreads();
    if (tid == 0) {
        float m = s_red[0];
        for (int k = 1; k < kT / 32; ++k) m = fminf(m, s_red[k]);
        s_thr = __fsub_rn(m, 1.0f);                                          // :33 threshold_logit = min(fg) - delta
    }
    __syncthreads();
    const float thr = s_thr;
    // per positive (in ascending order): a = sum_fg clamp((fg - x)/2 + .5) + .5 ; b = sum_validbg clamp((bg - x)/2 + .5)
    for (int a = 0; a < nfg; ++a) {
        const float x = logits[(int)fg_keys[a]];
        float sa = 0.f, sb = 0.f;
        for (int i = tid; i < n; i += kT) {
            const float t = targets[i], v = logits[i];
            const float c = fminf(fmaxf(__fadd_rn(__fdiv_rn(__fsub_rn(v, x), 2.0f), 0.5f), 0.f), 1.f);
            if (t == 1.f) sa += c;
            else if (t == 0.f && v >= thr) sb += c;                          // :36 valid negatives
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, d);
            sb += __shfl_xor_sync(0xffffffffu, sb, d);
        }
        __syncthreads();
        if (lane == 0) { s_red[warp] = sa; }
        __syncthreads();
        float ta = 0.f;
        if (tid == 0) { for (int k = 0; k < kT / 32; ++k) ta += s_red[k]; }
        __syncthreads();
        if (lane == 0) { s_red[warp] = sb; }
        __syncthreads();
        if (tid == 0) {
            float tb = 0.f;
            for (int k = 0; k < kT / 32; ++k) tb += s_red[k];
            fg_ab[2 * a] = __fadd_rn(ta, 0.5f);                              // :57
            fg_ab[2 * a + 1] = tb;                                           // :59
        }
        __syncthreads();
    }
    // running max of the interpolated precision in ascending-logit order (:62-66), per-positive scale for bg grads
    if (tid == 0) {
        float max_prec = 0.f, sum_prec = 0.f;
        for (int a = 0; a < nfg; ++a) {
            const float A = fg_ab[2 * a], B = fg_ab[2 * a + 1];
            const float cur = __fdiv_rn(A, __fadd_rn(A, B));
            float scale = 1.f;
            if (max_prec <= cur) max_prec = cur;
            else scale = __fdiv_rn(__fsub_rn(1.f, max_prec), __fsub_rn(1.f, cur));
            fg_prec[a] = max_prec;
            fg_ab[2 * a] = __fadd_rn(A, B);      // reuse: a+b
            fg_ab[2 * a + 1] = scale;
            sum_prec += max_prec;
        }
        const float fn = (float)nfg;
        *loss = __fsub_rn(1.f, __fdiv_rn(sum_prec, fn));                     // :79-81
    }
    __syncthreads();
    const float fn = (float)nfg;
    for (int a = tid; a < nfg; a += kT)
        grad[(int)fg_keys[a]] = __fdiv_rn(-__fsub_rn(1.f, fg_prec[a]), fn);   // :71,75
    for (int i = tid; i < n; i += kT) {
        const float t = targets[i], v = logits[i];
        if (t == 0.f && v >= thr) {
            float g = 0.f;
            for (int a = 0; a < nfg; ++a) {                                   // same order as `valid_bg_grad += tmp2`
                const float x = logits[(int)fg_keys[a]];
                float c = fminf(fmaxf(__fadd_rn(__fdiv_rn(__fsub_rn(v, x), 2.0f), 0.5f), 0.f), 1.f);
                c = __fdiv_rn(c, fg_ab[2 * a]);                               // :60 tmp2 /= (a+b)
                const float sc = fg_ab[2 * a + 1];
                if (sc != 1.f) c = __fmul_rn(c, sc);                          // :66
                g = __fadd_rn(g, c);
            }
            grad[i] = __fdiv_rn(g, fn);                                       // :70,75
        }
    }
}

__global__ void prune_kernel(const float* __restrict__ x, int64_t n, int method, float thr, float temp,
                             float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = prune(x[i], method, thr, temp);
}

// A[ra,ca,:] = B[rb,cb,:]; product form when ca == nullptr (see header)
__global__ void indices_copy_kernel(float* __restrict__ A, int64_t colsA, const float* __restrict__ B, int64_t colsB,
                                    int64_t C, const int64_t* __restrict__ ra, const int64_t* __restrict__ ca,
                                    const int64_t* __restrict__ rb, const int64_t* __restrict__ cb, int64_t npairs) {
    const int64_t total = (ca ? npairs : npairs * npairs) * C;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        const int64_t k = t / C, c = t - k * C;
        int64_t r_a, c_a, r_b, c_b;
        if (ca) {
            r_a = ra[k]; c_a = ca[k]; r_b = rb[k]; c_b = cb[k];
        } else {
            const int64_t i = k / npairs, j = k - i * npairs;
            r_a = ra[i]; c_a = ra[j]; r_b = i; c_b = j;
        }
        A[(r_a * colsA + c_a) * C + c] = B[(r_b * colsB + c_b) * C + c];
    }
}


// ------------------------------------------------------------------------------------------ target-assignment overlaps
// lib/rpn_util.py:439-461 (compute_targets): ols = iou(rois, gts) [M,G] in float64 (numpy promotes the float32 rois
// against the float64 ground truths; only area_a stays float32 arithmetic, flag area_f32), its row max / argmax
// (first maximum, NaN counts as the maximum: np.amax / np.argmax) and, per ground truth, the best roi (first maximum
// over the rows).  kind = 1: iou_ign (lib/core.py:535-575), intersection over the roi's own area.  A thread owns one
// roi; the G <= 64 ground truths sit in shared memory; the per-column (value, row) maxima are reduced per CTA into
// partials, finalised by col_finalize_kernel.
constexpr int kTgtThreads = 256;
constexpr int kTgtMaxG = 64;
__device__ __forceinline__ bool better_max(double v, double best) {          // strict: the first maximum stays
    return (v != v && best == best) || (best == best && v > best);
}

__global__ void __launch_bounds__(kTgtThreads)
targets_overlaps_kernel(const double* __restrict__ rois, int64_t ld, int M, const double* __restrict__ gts, int G, int kind,
                        int area_f32, double* __restrict__ ols, double* __restrict__ row_max, int64_t* __restrict__ row_arg,
                        double* __restrict__ part_val, int32_t* __restrict__ part_idx) {
    __shared__ double s_g[kTgtMaxG][5];
    __shared__ double s_cv[kTgtThreads / 32][kTgtMaxG];
    __shared__ int s_ci[kTgtThreads / 32][kTgtMaxG];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int g = tid; g < G; g += kTgtThreads) {
        const double x1 = gts[g * 4 + 0], y1 = gts[g * 4 + 1], x2 = gts[g * 4 + 2], y2 = gts[g * 4 + 3];
        s_g[g][0] = x1; s_g[g][1] = y1; s_g[g][2] = x2; s_g[g][3] = y2;
        s_g[g][4] = __dmul_rn(__dsub_rn(x2, x1), __dsub_rn(y2, y1));
    }
    __syncthreads();
    const int i = blockIdx.x * kTgtThreads + tid;
    const bool live = i < M;
    double x1 = 0, y1 = 0, x2 = 0, y2 = 0, area = 0;
    if (live) {
        x1 = rois[(int64_t)i * ld + 0]; y1 = rois[(int64_t)i * ld + 1]; x2 = rois[(int64_t)i * ld + 2]; y2 = rois[(int64_t)i * ld + 3];
        if (area_f32) area = (double)__fmul_rn(__fsub_rn((float)x2, (float)x1), __fsub_rn((float)y2, (float)y1));
        else area = __dmul_rn(__dsub_rn(x2, x1), __dsub_rn(y2, y1));
    }
    double best = -INFINITY;
    int barg = 0;
    for (int g = 0; g < G; ++g) {
        double v = -INFINITY;
        if (live) {
            const double w = fmax(__dsub_rn(fmin(x2, s_g[g][2]), fmax(x1, s_g[g][0])), 0.0);     // lib/core.py:205-207
            const double h = fmax(__dsub_rn(fmin(y2, s_g[g][3]), fmax(y1, s_g[g][1])), 0.0);
            const double inter = __dmul_rn(w, h);                                                // :218
            double uni;
            if (kind == 0) uni = __dsub_rn(__dadd_rn(area, s_g[g][4]), inter);                   // :511
            else uni = __dsub_rn(__dadd_rn(area, __dmul_rn(s_g[g][4], 0.0)), __dmul_rn(inter, 0.0));   // :564
            v = __ddiv_rn(inter, uni);
            if (ols) ols[(int64_t)i * G + g] = v;
            if (g == 0 || better_max(v, best)) { best = v; barg = g; }
        }
        // column maximum over the rows of this CTA: (value, lowest row index)
        double cv = v;
        int ci = live ? i : INT_MAX;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, cv, d);
            const int oi = __shfl_down_sync(0xffffffffu, ci, d);
            if (oi != INT_MAX && (ci == INT_MAX || better_max(ov, cv) || (!(better_max(cv, ov)) && oi < ci))) { cv = ov; ci = oi; }
        }
        if (lane == 0) { s_cv[warp][g] = cv; s_ci[warp][g] = ci; }
    }
    if (live) { row_max[i] = best; row_arg[i] = barg; }
    __syncthreads();
    for (int g = tid; g < G; g += kTgtThreads) {
        double cv = s_cv[0][g];
        int ci = s_ci[0][g];
        for (int w = 1; w < kTgtThreads / 32; ++w) {
            const double ov = s_cv[w][g];
            const int oi = s_ci[w][g];
            if (oi != INT_MAX && (ci == INT_MAX || better_max(ov, cv) || (!(better_max(cv, ov)) && oi < ci))) { cv = ov; ci = oi; }
        }
        part_val[(size_t)blockIdx.x * G + g] = cv;
        part_idx[(size_t)blockIdx.x * G + g] = ci;
    }
}

__global__ void col_finalize_kernel(const double* __restrict__ part_val, const int32_t* __restrict__ part_idx, int nblocks, int G,
                                    double* __restrict__ col_max, int64_t* __restrict__ col_arg) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double cv = part_val[g];
    int ci = part_idx[g];
    for (int b = 1; b < nblocks; ++b) {                                     // blocks are in row order: strict keeps the first maximum
        const double ov = part_val[(size_t)b * G + g];
        const int oi = part_idx[(size_t)b * G + g];
        if (oi != INT_MAX && (ci == INT_MAX || better_max(ov, cv))) { cv = ov; ci = oi; }
    }
    col_max[g] = cv;
    col_arg[g] = ci == INT_MAX ? 0 : ci;
}

// keep lists for the host: the first min(count, K) entries of every image's int64 valid_idx as int32 (a keep list is a few
// dozen entries long, the [batch, N] int64 array it sits in 8 N bytes per image)
__global__ void pack_keep_kernel(const int64_t* __restrict__ valid_idx, const int32_t* __restrict__ counts, int N, int K,
                                 int32_t* __restrict__ out) {
    const int b = blockIdx.y;
    const int n = min(counts[2 * b], K);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x)
        out[(size_t)b * K + k] = k < n ? (int32_t)valid_idx[(size_t)b * N + k] : -1;
}

}  // namespace gnms

using namespace gnms;

extern "C" int gnms_pack_keep_i32(const int64_t* valid_idx, const int32_t* counts, int N, int batch, int K, int32_t* out, void* stream) {
    if (N < 0 || batch < 0 || K <= 0) return GNMS_E_BADARG;
    if (batch == 0) return 0;
    if (!valid_idx || !counts || !out) return GNMS_E_BADARG;
    pack_keep_kernel<<<dim3((K + 255) / 256, batch), 256, 0, (cudaStream_t)stream>>>(valid_idx, counts, N, K, out);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_prune_f32(const float* x, int64_t n, int pruning_method, float nms_threshold, float temperature,
                              float* out, void* stream) {
    if (n < 0 || pruning_method < 0 || pruning_method > 2) return GNMS_E_BADARG;
    if (n == 0) return 0;
    if (!x || !out) return GNMS_E_BADARG;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    prune_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, pruning_method, nms_threshold, temperature, out);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_indices_copy_f32(float* A, int64_t colsA, const float* B, int64_t colsB, int64_t C,
                                     const int64_t* ra, const int64_t* ca, const int64_t* rb, const int64_t* cb,
                                     int64_t npairs, void* stream) {
    if (npairs < 0 || C <= 0 || colsA <= 0 || colsB <= 0) return GNMS_E_BADARG;
    if (npairs == 0) return 0;
    if (!A || !B || !ra || (ca && (!rb || !cb))) return GNMS_E_BADARG;
    const int64_t total = (ca ? npairs : npairs * npairs) * C;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    indices_copy_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(A, colsA, B, colsB, C, ra, ca, rb, cb, npairs);
    GNMS_LAUNCH_CHECK();
    return 0;
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t gnms_soft_nms_workspace_bytes(int N) {
    if (N <= 0) return 256;
    return al256((size_t)N * 8) + al256((size_t)N * 4);
}

extern "C" int gnms_soft_nms_f64(const double* dets, int N, double sigma, double Nt, double threshold, int method,
                                 double shift, int32_t* keep, double* keep_scores, int32_t* n_keep, void* workspace,
                                 void* stream) {
    if (N < 0 || method < 0 || method > 2 || !n_keep) return GNMS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 0) return (int)cudaMemsetAsync(n_keep, 0, sizeof(int32_t), s);
    if (!dets || !keep || !keep_scores || !workspace) return GNMS_E_BADARG;
    double* score = reinterpret_cast<double*>(workspace);
    int32_t* state = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) + al256((size_t)N * 8));
    soft_nms_kernel<<<1, kT, 0, s>>>(dets, N, sigma, Nt, threshold, method, shift, keep, keep_scores, n_keep, score, state);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t gnms_aploss_workspace_bytes(int n) {
    if (n <= 0) return 256;
    return al256((size_t)n * 4) + al256((size_t)n * 8) + al256((size_t)n * 4) + al256((size_t)n * 8);
}

extern "C" int gnms_aploss_f32(const float* logits, const float* targets, int n, float* loss, float* grad,
                               void* workspace, size_t workspace_bytes, void* stream) {
    if (n < 0 || !loss) return GNMS_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return (int)cudaMemsetAsync(loss, 0, sizeof(float), s);
    if (!logits || !targets || !grad || !workspace || workspace_bytes < gnms_aploss_workspace_bytes(n)) return GNMS_E_BADARG;
    char* w = reinterpret_cast<char*>(workspace);
    int32_t* fg_list = reinterpret_cast<int32_t*>(w);
    float* fg_ab = reinterpret_cast<float*>(w + al256((size_t)n * 4));
    float* fg_prec = reinterpret_cast<float*>(w + al256((size_t)n * 4) + al256((size_t)n * 8));
    unsigned long long* fg_keys = reinterpret_cast<unsigned long long*>(w + al256((size_t)n * 4) + al256((size_t)n * 8) + al256((size_t)n * 4));
    aploss_kernel<<<1, kT, 0, s>>>(logits, targets, n, loss, grad, fg_list, fg_ab, fg_prec, fg_keys);
    GNMS_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------- 2D overlaps in fp64
// The reference's iou / intersect keep the dtype of their inputs: float64 numpy arrays (the detections of the inference
// path, lib/rpn_util.py:1295) are compared with the NMS threshold in float64.  One thread per output, the numpy branch's
// operation order (lib/core.py:196-206, 498-515), separately rounded.
__global__ void __launch_bounds__(256) overlap2d_f64_kernel(const double* __restrict__ a, int64_t ld_a, int M, const double* __restrict__ b,
                                                            int64_t ld_b, int N, int kind, int list_mode, int area_f32, double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = list_mode ? M : (int64_t)M * N;
    if (t >= total) return;
    const int i = list_mode ? (int)t : (int)(t / N), j = list_mode ? (int)t : (int)(t % N);
    const double* pa = a + (int64_t)i * ld_a;
    const double* pb = b + (int64_t)j * ld_b;
    const double iw = fmax(__dsub_rn(fmin(pa[2], pb[2]), fmax(pa[0], pb[0])), 0.0);
    const double ih = fmax(__dsub_rn(fmin(pa[3], pb[3]), fmax(pa[1], pb[1])), 0.0);
    const double inter = __dmul_rn(iw, ih);
    if (kind == GNMS_KIND_INTERSECT) { out[t] = inter; return; }
    // a side that was float32 in the caller's hands has its area computed in float32 (numpy / torch keep float32 for
    // float32 x float32 and promote only where the two sides meet; the coordinates are exact in both types)
    const double aa = (area_f32 & 1) ? (double)__fmul_rn(__fsub_rn((float)pa[2], (float)pa[0]), __fsub_rn((float)pa[3], (float)pa[1]))
                                     : __dmul_rn(__dsub_rn(pa[2], pa[0]), __dsub_rn(pa[3], pa[1]));
    const double ab = (area_f32 & 2) ? (double)__fmul_rn(__fsub_rn((float)pb[2], (float)pb[0]), __fsub_rn((float)pb[3], (float)pb[1]))
                                     : __dmul_rn(__dsub_rn(pb[2], pb[0]), __dsub_rn(pb[3], pb[1]));
    out[t] = __ddiv_rn(inter, __dsub_rn(__dadd_rn(aa, ab), inter));
}

extern "C" int gnms_overlap2d_f64(const double* a, int64_t ld_a, int M, const double* b, int64_t ld_b, int N, int kind, int list_mode,
                                  int area_f32, double* out, void* stream) {
    if (M < 0 || N < 0 || ld_a < 4 || ld_b < 4 || (kind != GNMS_KIND_IOU && kind != GNMS_KIND_INTERSECT) || (list_mode && M != N))
        return GNMS_E_BADARG;
    if (M == 0 || N == 0) return 0;
    if (!a || !b || !out) return GNMS_E_BADARG;
    const int64_t total = list_mode ? M : (int64_t)M * N;
    overlap2d_f64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, ld_a, M, b, ld_b, N, kind, list_mode, area_f32, out);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t gnms_targets_overlaps_workspace_bytes(int M, int G) {
    if (M <= 0 || G <= 0) return 256;
    const size_t nb = (size_t)((M + kTgtThreads - 1) / kTgtThreads);
    return nb * G * (sizeof(double) + sizeof(int32_t)) + 512;
}

extern "C" int gnms_targets_overlaps_f64(const double* rois, int64_t ld_rois, int M, const double* gts, int G, int kind, int area_f32,
                                         double* ols, double* row_max, int64_t* row_arg, double* col_max, int64_t* col_arg,
                                         void* workspace, void* stream) {
    if (M < 0 || G < 0 || ld_rois < 4 || kind < 0 || kind > 1) return GNMS_E_BADARG;
    if (G > kTgtMaxG) return GNMS_E_TOOLARGE;
    if (M == 0 || G == 0) return 0;
    if (!rois || !gts || !row_max || !row_arg || !col_max || !col_arg || !workspace) return GNMS_E_BADARG;
    const int nb = (M + kTgtThreads - 1) / kTgtThreads;
    double* pv = reinterpret_cast<double*>(workspace);
    int32_t* pi = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) + (((size_t)nb * G * sizeof(double) + 255) & ~(size_t)255));
    cudaStream_t s = (cudaStream_t)stream;
    targets_overlaps_kernel<<<nb, kTgtThreads, 0, s>>>(rois, ld_rois, M, gts, G, kind, area_f32, ols, row_max, row_arg, pv, pi);
    GNMS_LAUNCH_CHECK();
    col_finalize_kernel<<<(G + 63) / 64, 64, 0, s>>>(pv, pi, nb, G, col_max, col_arg);
    GNMS_LAUNCH_CHECK();
    return 0;
}

// ---- exact polygon iou3d (lib/core.py:246-302) --------------------------------------------------------------------
// One thread per pair.  The per-box preparation (face, y range, volume) is a few dozen flops and is redone per pair; the
// callers of the reference evaluate tens of detections against a handful of ground-truth boxes (lib/rpn_util.py:1776).
namespace gnms {

__global__ void __launch_bounds__(128)
iou3d_exact_kernel(const double* __restrict__ ca, int64_t ld_a, int M, const double* __restrict__ cb, int64_t ld_b, int N,
                   const double* __restrict__ vol, int list_mode, double* __restrict__ out_bev, double* __restrict__ out_3d) {
    const int64_t total = list_mode ? (int64_t)M : (int64_t)M * N;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = list_mode ? p : p / N, j = list_mode ? p : p % N;
        double a[24], b[24];
        for (int k = 0; k < 24; ++k) { a[k] = ca[i * ld_a + k]; b[k] = cb[j * ld_b + k]; }
        ExactBox b1, b2;
        exact_box_from_corners(a, b1);
        exact_box_from_corners(b, b2);
        double bev, v3;
        iou3d_exact_pair(b1, b2, vol != nullptr, vol ? vol[p] : 0.0, bev, v3);
        if (out_bev) out_bev[p] = bev;
        if (out_3d) out_3d[p] = v3;
    }
}

}  // namespace gnms

extern "C" int gnms_iou3d_exact_f64(const double* corners_a, int64_t ld_a, int M, const double* corners_b, int64_t ld_b, int N,
                                    const double* vol, int list_mode, double* out_bev, double* out_3d, void* stream) {
    if (M < 0 || N < 0 || ld_a < 24 || ld_b < 24) return GNMS_E_BADARG;
    if (list_mode && M != N) return GNMS_E_BADARG;
    if (M == 0 || N == 0) return 0;
    if (!corners_a || !corners_b || (!out_bev && !out_3d)) return GNMS_E_BADARG;
    const int64_t total = list_mode ? (int64_t)M : (int64_t)M * N;
    const int64_t want = (total + 127) / 128;
    const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
    iou3d_exact_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(corners_a, ld_a, M, corners_b, ld_b, N, vol, list_mode, out_bev, out_3d);
    GNMS_LAUNCH_CHECK();
    return 0;
}
