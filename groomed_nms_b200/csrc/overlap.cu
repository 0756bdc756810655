// Pairwise overlap kernels: 2D IoU / intersection, 7-DoF -> corners, corner -> per-box records, approximate
// (axis-aligned, optionally generalized) 3D IoU.  API-compatible products of lib/core.py:178-532 and
// lib/math_3d.py:364-435: the N x N result is materialised in HBM, so these kernels are HBM-write bound
// (4 bytes per pair) with an fp32 ALU co-limit of ~25 (2D) / ~55 (3D GIoU) issue slots per pair.
//
// Tiling: one CTA = kThreads threads x 4 consecutive columns (one 16 B streaming store per row per thread,
// 512 B contiguous per warp) x kRows rows.  Column boxes live in registers for the whole tile, row boxes are
// staged once per CTA in shared memory and read as warp-wide broadcasts.
#include "common.cuh"

namespace gnms {

constexpr int kThreads = 128;
constexpr int kColsPerThread = 4;
constexpr int kColsPerCta = kThreads * kColsPerThread;   // 512
constexpr int kRows = 16;

template <int kKind, bool kVec>
__global__ void __launch_bounds__(kThreads) overlap2d_kernel(const float* __restrict__ a, int M,
                                                             const float* __restrict__ b, int N,
                                                             float* __restrict__ out, int64_t ld,
                                                             int64_t box_img_stride, int64_t out_img_stride) {
    __shared__ Box2 rows[kRows];
    a += blockIdx.z * box_img_stride;      // batched self-overlap: image z (strides are 0 for single calls)
    b += blockIdx.z * box_img_stride;
    out += blockIdx.z * out_img_stride;
    const int i0 = blockIdx.y * kRows;
    const int j0 = blockIdx.x * kColsPerCta + threadIdx.x * kColsPerThread;
    if (threadIdx.x < kRows) {
        int i = i0 + threadIdx.x;
        if (i < M) rows[threadIdx.x] = make_box2(__ldg(reinterpret_cast<const float4*>(a) + i));
    }
    Box2 cb[kColsPerThread];
#pragma unroll
    for (int k = 0; k < kColsPerThread; ++k) {
        int j = j0 + k;
        cb[k] = make_box2(j < N ? __ldg(reinterpret_cast<const float4*>(b) + j) : make_float4(0, 0, 0, 0));
    }
    __syncthreads();
    if (j0 >= N) return;
    const int nrow = min(kRows, M - i0);
#pragma unroll 4
    for (int r = 0; r < nrow; ++r) {
        const Box2 ra = rows[r];
        float v[kColsPerThread];
#pragma unroll
        for (int k = 0; k < kColsPerThread; ++k)
            v[k] = (kKind == GNMS_KIND_IOU) ? iou2(ra, cb[k]) : intersect2(ra, cb[k]);
        float* dst = out + (int64_t)(i0 + r) * ld + j0;
        if (kVec && j0 + kColsPerThread <= N) {
            st_cs_f4(dst, make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
            for (int k = 0; k < kColsPerThread; ++k)
                if (j0 + k < N) dst[k] = v[k];
        }
    }
}

template <int kKind>
__global__ void overlap2d_list_kernel(const float* __restrict__ a, const float* __restrict__ b, int M,
                                      float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    Box2 x = make_box2(__ldg(reinterpret_cast<const float4*>(a) + i));
    Box2 y = make_box2(__ldg(reinterpret_cast<const float4*>(b) + i));
    out[i] = (kKind == GNMS_KIND_IOU) ? iou2(x, y) : intersect2(x, y);
}

// 7-DoF -> 8 corners.  Template corners (l on x for idx {1,3,5,6}; h on y for {2,3,6,7}; w on z for {4,5,6,7},
// centred), rotation about y, translation                                     lib/math_3d.py:369-435
// corners of one cuboid (lib/math_3d.py:364-435), x / y / z of the 8 corners
// kitti_order = false: the iou_3d_convention vertex order (:379-403); true: the other order of the reference (:405-426, x-high for
// corners {1,2,3,4}, y-high for {2,3,6,7}, z-high for {3,4,5,6})
__device__ __forceinline__ void corners_of(const float* __restrict__ p, float (&ox)[8], float (&oy)[8], float (&oz)[8],
                                           bool kitti_order = false) {
    float x = p[0], y = p[1], z = p[2], w = p[3], h = p[4], l = p[5], ry = p[6];
    float cs = cosf(ry), sn = sinf(ry);
    float hl = __fdiv_rn(l, 2.0f), hh = __fdiv_rn(h, 2.0f), hw = __fdiv_rn(w, 2.0f);
    float cx[2] = {__fsub_rn(0.0f, hl), __fsub_rn(l, hl)};
    float cy[2] = {__fsub_rn(0.0f, hh), __fsub_rn(h, hh)};
    float cz[2] = {__fsub_rn(0.0f, hw), __fsub_rn(w, hw)};
    // corner k uses: x-high for k in {1,3,5,6}; y-high for {2,3,6,7}; z-high for {4,5,6,7}
    const unsigned xmask = kitti_order ? 0x1Eu : 0x6Au, ymask = 0xCCu, zmask = kitti_order ? 0x78u : 0xF0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float px = cx[(xmask >> k) & 1], py = cy[(ymask >> k) & 1], pz = cz[(zmask >> k) & 1];
        // R = [[cos,0,sin],[0,1,0],[-sin,0,cos]] ; products rounded separately
        ox[k] = __fadd_rn(__fadd_rn(__fmul_rn(cs, px), __fmul_rn(sn, pz)), x);
        oy[k] = __fadd_rn(py, y);
        oz[k] = __fadd_rn(__fadd_rn(__fmul_rn(-sn, px), __fmul_rn(cs, pz)), z);
    }
}
// per-box record iou3d_approximate derives from the corners (lib/core.py:354-388)
__device__ __forceinline__ void record_of(float4 x0, float4 x1, float4 y0, float4 y1, float4 z0, float4 z1, float4& o0, float4& o1) {
    float xmin = fminf(fminf(fminf(x0.x, x0.y), fminf(x0.z, x0.w)), fminf(fminf(x1.x, x1.y), fminf(x1.z, x1.w)));
    float xmax = fmaxf(fmaxf(fmaxf(x0.x, x0.y), fmaxf(x0.z, x0.w)), fmaxf(fmaxf(x1.x, x1.y), fmaxf(x1.z, x1.w)));
    float ymin = fminf(fminf(fminf(y0.x, y0.y), fminf(y0.z, y0.w)), fminf(fminf(y1.x, y1.y), fminf(y1.z, y1.w)));
    float ymax = fmaxf(fmaxf(fmaxf(y0.x, y0.y), fmaxf(y0.z, y0.w)), fmaxf(fmaxf(y1.x, y1.y), fmaxf(y1.z, y1.w)));
    float zmin = fminf(fminf(fminf(z0.x, z0.y), fminf(z0.z, z0.w)), fminf(fminf(z1.x, z1.y), fminf(z1.z, z1.w)));
    float zmax = fmaxf(fmaxf(fmaxf(z0.x, z0.y), fmaxf(z0.z, z0.w)), fmaxf(fmaxf(z1.x, z1.y), fmaxf(z1.z, z1.w)));
    // BEV extents use corners [2,3,6,7] only (lib/core.py:383-388)
    float bx1 = fminf(fminf(x0.z, x0.w), fminf(x1.z, x1.w)), bx2 = fmaxf(fmaxf(x0.z, x0.w), fmaxf(x1.z, x1.w));
    float bz1 = fminf(fminf(z0.z, z0.w), fminf(z1.z, z1.w)), bz2 = fmaxf(fmaxf(z0.z, z0.w), fmaxf(z1.z, z1.w));
    float vol = __fmul_rn(__fmul_rn(__fsub_rn(xmax, xmin), __fsub_rn(ymax, ymin)), __fsub_rn(zmax, zmin));  // :448
    float abev = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(bz2, bz1));
    o0 = make_float4(ymin, ymax, bx1, bx2);
    o1 = make_float4(bz1, bz2, vol, abev);
}

__global__ void corners_kernel(const float* __restrict__ boxes7, int64_t ld, int N, float* __restrict__ corners, int kitti_order) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float ox[8], oy[8], oz[8];
    corners_of(boxes7 + (int64_t)n * ld, ox, oy, oz, kitti_order != 0);
    float4* o = reinterpret_cast<float4*>(corners + (int64_t)n * 24);
    o[0] = make_float4(ox[0], ox[1], ox[2], ox[3]);
    o[1] = make_float4(ox[4], ox[5], ox[6], ox[7]);
    o[2] = make_float4(oy[0], oy[1], oy[2], oy[3]);
    o[3] = make_float4(oy[4], oy[5], oy[6], oy[7]);
    o[4] = make_float4(oz[0], oz[1], oz[2], oz[3]);
    o[5] = make_float4(oz[4], oz[5], oz[6], oz[7]);
}
__global__ void records_kernel(float* __restrict__ corners, int N, float* __restrict__ rec, int mutate) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float4* c = reinterpret_cast<float4*>(corners + (int64_t)n * 24);
    float4 x0 = c[0], x1 = c[1], y0 = c[2], y1 = c[3], z0 = c[4], z1 = c[5];
    float4* o = reinterpret_cast<float4*>(rec + (int64_t)n * 8);
    record_of(x0, x1, y0, y1, z0, z1, o[0], o[1]);
    if (mutate) {  // the reference overwrites Y with Z in the caller's storage (lib/core.py:379-380)
        c[2] = z0;
        c[3] = z1;
    }
}
// both in one pass: 7-DoF boxes -> records (and, if asked, the corners as well); same arithmetic, same bits
__global__ void records7_kernel(const float* __restrict__ boxes7, int64_t ld, int N, float* __restrict__ rec, float* __restrict__ corners) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float ox[8], oy[8], oz[8];
    corners_of(boxes7 + (int64_t)n * ld, ox, oy, oz);
    const float4 x0 = make_float4(ox[0], ox[1], ox[2], ox[3]), x1 = make_float4(ox[4], ox[5], ox[6], ox[7]);
    const float4 y0 = make_float4(oy[0], oy[1], oy[2], oy[3]), y1 = make_float4(oy[4], oy[5], oy[6], oy[7]);
    const float4 z0 = make_float4(oz[0], oz[1], oz[2], oz[3]), z1 = make_float4(oz[4], oz[5], oz[6], oz[7]);
    float4* o = reinterpret_cast<float4*>(rec + (int64_t)n * 8);
    record_of(x0, x1, y0, y1, z0, z1, o[0], o[1]);
    if (corners) {
        float4* c = reinterpret_cast<float4*>(corners + (int64_t)n * 24);
        c[0] = x0; c[1] = x1; c[2] = y0; c[3] = y1; c[4] = z0; c[5] = z1;
    }
}

template <bool kGen, bool kAffine, bool kVec>
__global__ void __launch_bounds__(kThreads) overlap3d_kernel(const float* __restrict__ ra, int M,
                                                             const float* __restrict__ rb, int N,
                                                             float* __restrict__ out_bev, float* __restrict__ out_3d,
                                                             int64_t ld, const float* __restrict__ mul2d,
                                                             int64_t rec_img_stride, int64_t out_img_stride) {
    __shared__ Rec3 rows[kRows];
    ra += blockIdx.z * rec_img_stride;     // batched self-overlap: image z (strides are 0 for single calls)
    rb += blockIdx.z * rec_img_stride;
    if (out_bev) out_bev += blockIdx.z * out_img_stride;
    if (out_3d) out_3d += blockIdx.z * out_img_stride;
    if (mul2d) mul2d += blockIdx.z * out_img_stride;
    const int i0 = blockIdx.y * kRows;
    const int j0 = blockIdx.x * kColsPerCta + threadIdx.x * kColsPerThread;
    if (threadIdx.x < kRows) {
        int i = i0 + threadIdx.x;
        if (i < M) rows[threadIdx.x] = load_rec3(ra + (int64_t)i * 8);
    }
    Rec3 cb[kColsPerThread];
#pragma unroll
    for (int k = 0; k < kColsPerThread; ++k) {
        int j = min(j0 + k, N - 1);
        cb[k] = load_rec3(rb + (int64_t)j * 8);
    }
    __syncthreads();
    if (j0 >= N) return;
    const int nrow = min(kRows, M - i0);
    const bool full = kVec && (j0 + kColsPerThread <= N);
#pragma unroll 2
    for (int r = 0; r < nrow; ++r) {
        const Rec3 a = rows[r];
        float vb[kColsPerThread], v3[kColsPerThread];
#pragma unroll
        for (int k = 0; k < kColsPerThread; ++k) {
            float ibev = inter_bev3(a, cb[k]);
            if (out_bev) vb[k] = iou_bev3(a, cb[k], ibev);
            if (out_3d) v3[k] = iou3<kGen, kAffine>(a, cb[k], ibev);
        }
        const int64_t off = (int64_t)(i0 + r) * ld + j0;
        if (out_3d && mul2d) {
#pragma unroll
            for (int k = 0; k < kColsPerThread; ++k)
                if (j0 + k < N) v3[k] = __fmul_rn(__ldg(mul2d + off + k), v3[k]);   // lib/loss/rpn_3d.py:786
        }
        if (full) {
            if (out_bev) st_cs_f4(out_bev + off, make_float4(vb[0], vb[1], vb[2], vb[3]));
            if (out_3d) st_cs_f4(out_3d + off, make_float4(v3[0], v3[1], v3[2], v3[3]));
        } else {
#pragma unroll
            for (int k = 0; k < kColsPerThread; ++k)
                if (j0 + k < N) {
                    if (out_bev) out_bev[off + k] = vb[k];
                    if (out_3d) out_3d[off + k] = v3[k];
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Symmetric self-overlap tiles.  overlap(i,j) == overlap(j,i) bitwise (min/max and the fp32 additions commute), so
// only tile pairs I <= J of 64 x 64 are evaluated: the tile is stored directly (16 B streaming stores along rows)
// and, for I != J, a second time transposed through shared memory.  Halves the ALU work of the ALU-co-limited
// N x N kernels; HBM traffic stays the algorithmic 4 N^2 bytes.
// 256 threads: tx = tid & 15 owns columns 4tx..4tx+3, ty = tid >> 4 owns rows ty, ty+16, ty+32, ty+48.
constexpr int kST = 64;            // tile edge
constexpr int kSS = kST + 1;       // shared-memory row stride (floats)

__device__ __forceinline__ void sym_tile_coords(int t, int nt, int& I, int& J) {
    // upper-triangular enumeration: row I holds tiles (I,I)..(I,nt-1); t = I*nt - I*(I-1)/2 + (J - I)
    float fn = 2.0f * nt + 1.0f;
    int i = (int)((fn - sqrtf(fn * fn - 8.0f * (float)t)) * 0.5f);
    i = max(0, min(i, nt - 1));
    while (i > 0 && i * nt - i * (i - 1) / 2 > t) --i;
    while ((i + 1) * nt - (i + 1) * i / 2 <= t) ++i;
    I = i;
    J = i + (t - (i * nt - i * (i - 1) / 2));
}

template <bool kGen, bool kAffine>
__device__ __noinline__ float iou3_exact_noinline(Rec3 a, Rec3 b) { return iou3<kGen, kAffine>(a, b, inter_bev3(a, b)); }
__device__ __noinline__ float iou2_exact_noinline(Box2 a, Box2 b) { return iou2(a, b); }

template <typename Rec, typename F, typename FX>
__device__ __forceinline__ void sym_tile_body(int N, float* __restrict__ out, bool vec, const Rec (&rr)[4],
                                              const Rec (&cr)[4], int I, int J, float* tile, bool tile_unsafe,
                                              F pairfn, FX exactfn) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = I * kST, j0 = J * kST;
    float v[4][4];
    bool unsafe = tile_unsafe;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) v[r][k] = pairfn(rr[r], cr[k], unsafe);
    if (__builtin_expect(unsafe, 0)) {     // rare: operands outside div_rn_fast's proven range -> exact IEEE path
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) v[r][k] = exactfn(rr[r], cr[k]);
    }
    // direct tile
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty + 16 * r, j = j0 + 4 * tx;
        if (i < N && j < N) {
            float* dst = out + (int64_t)i * N + j;
            if (vec && j + 4 <= N) st_cs_f4(dst, make_float4(v[r][0], v[r][1], v[r][2], v[r][3]));
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k) if (j + k < N) dst[k] = v[r][k];
            }
        }
    }
    if (I == J) return;
    // transposed tile through shared memory: tile[c][r]
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) tile[(4 * tx + k) * kSS + ty + 16 * r] = v[r][k];
    __syncthreads();
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const int c = ty + 16 * pass;                 // row of the transposed tile = column index j0 + c
        const int i = j0 + c, j = i0 + 4 * tx;
        if (i < N && j < N) {
            const float* src = tile + c * kSS + 4 * tx;
            float* dst = out + (int64_t)i * N + j;
            if (vec && j + 4 <= N) st_cs_f4(dst, make_float4(src[0], src[1], src[2], src[3]));
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k) if (j + k < N) dst[k] = src[k];
            }
        }
    }
}

template <bool kGen, bool kAffine>
__global__ void __launch_bounds__(256) overlap3d_self_kernel(const float* __restrict__ rec, int N, float* __restrict__ out,
                                                             int nt, int vec) {
    __shared__ float tile[kST * kSS];
    rec += (int64_t)blockIdx.y * N * 8;
    out += (int64_t)blockIdx.y * N * N;
    int I, J;
    sym_tile_coords(blockIdx.x, nt, I, J);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    Rec3 rr[4], cr[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) rr[r] = load_rec3(rec + (int64_t)min(I * kST + ty + 16 * r, N - 1) * 8);
#pragma unroll
    for (int k = 0; k < 4; ++k) cr[k] = load_rec3(rec + (int64_t)min(J * kST + 4 * tx + k, N - 1) * 8);
    // box-level preconditions of the straight-line division, one record per thread, OR-ed over the CTA
    bool bad = false;
    if (threadIdx.x < 2 * kST) {
        const int idx = (threadIdx.x < kST ? I * kST + threadIdx.x : J * kST + threadIdx.x - kST);
        if (idx < N) bad = !rec3_sane(load_rec3(rec + (int64_t)idx * 8));
    }
    const bool tile_unsafe = __syncthreads_or(bad);
    sym_tile_body(N, out, vec != 0, rr, cr, I, J, tile, tile_unsafe,
                  [](const Rec3& a, const Rec3& b, bool& u) { return iou3_fast<kGen, kAffine>(a, b, inter_bev3(a, b), u); },
                  [](const Rec3& a, const Rec3& b) { return iou3_exact_noinline<kGen, kAffine>(a, b); });
}

__global__ void __launch_bounds__(256) overlap2d_self_kernel(const float* __restrict__ boxes, int N, float* __restrict__ out,
                                                             int nt, int vec) {
    __shared__ float tile[kST * kSS];
    boxes += (int64_t)blockIdx.y * N * 4;
    out += (int64_t)blockIdx.y * N * N;
    int I, J;
    sym_tile_coords(blockIdx.x, nt, I, J);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    Box2 rr[4], cr[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) rr[r] = make_box2(__ldg(reinterpret_cast<const float4*>(boxes) + min(I * kST + ty + 16 * r, N - 1)));
#pragma unroll
    for (int k = 0; k < 4; ++k) cr[k] = make_box2(__ldg(reinterpret_cast<const float4*>(boxes) + min(J * kST + 4 * tx + k, N - 1)));
    bool bad = false;
    if (threadIdx.x < 2 * kST) {
        const int idx = (threadIdx.x < kST ? I * kST + threadIdx.x : J * kST + threadIdx.x - kST);
        if (idx < N) bad = !box2_sane(make_box2(__ldg(reinterpret_cast<const float4*>(boxes) + idx)));
    }
    const bool tile_unsafe = __syncthreads_or(bad);
    sym_tile_body(N, out, vec != 0, rr, cr, I, J, tile, tile_unsafe,
                  [](const Box2& a, const Box2& b, bool& u) { return iou2_fast(a, b, u); },
                  [](const Box2& a, const Box2& b) { return iou2_exact_noinline(a, b); });
}

template <bool kGen, bool kAffine>
__global__ void overlap3d_list_kernel(const float* __restrict__ ra, const float* __restrict__ rb, int M,
                                      float* __restrict__ out_bev, float* __restrict__ out_3d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    Rec3 a = load_rec3(ra + (int64_t)i * 8), b = load_rec3(rb + (int64_t)i * 8);
    float ibev = inter_bev3(a, b);
    if (out_bev) out_bev[i] = iou_bev3(a, b, ibev);
    if (out_3d) out_3d[i] = iou3<kGen, kAffine>(a, b, ibev);
}

// d iou(a,b) / d coords, scaled by the upstream gradient g, accumulated into ga[4], gb[4].
__device__ __forceinline__ void iou2_grad(const float4 a, const float4 b, float g, float* ga, float* gb) {
    const float wa = a.z - a.x, ha = a.w - a.y, wb = b.z - b.x, hb = b.w - b.y;
    const float Aa = wa * ha, Ab = wb * hb;
    const float xlo = fmaxf(a.x, b.x), xhi = fminf(a.z, b.z), ylo = fmaxf(a.y, b.y), yhi = fminf(a.w, b.w);
    const float dw = xhi - xlo, dh = yhi - ylo;
    const float iw = fmaxf(dw, 0.f), ih = fmaxf(dh, 0.f);
    const float I = iw * ih, U = (Aa + Ab) - I;
    const float inv = 1.0f / (U * U);
    const float gI = g * (U + I) * inv;      // d(I/U)/dI with U = Aa + Ab - I
    const float gA = -g * I * inv;           // d(I/U)/dAa = d(I/U)/dAb
    const float giw = (dw >= 0.f) ? gI * ih : 0.f;
    const float gih = (dh >= 0.f) ? gI * iw : 0.f;
    // min/max: the selected operand gets the gradient, ties split it evenly (torch.minimum / maximum)
    const float sxhi = a.z < b.z ? 1.f : (a.z == b.z ? 0.5f : 0.f);
    const float sxlo = a.x > b.x ? 1.f : (a.x == b.x ? 0.5f : 0.f);
    const float syhi = a.w < b.w ? 1.f : (a.w == b.w ? 0.5f : 0.f);
    const float sylo = a.y > b.y ? 1.f : (a.y == b.y ? 0.5f : 0.f);
    ga[2] += giw * sxhi;          gb[2] += giw * (1.f - sxhi);
    ga[0] -= giw * sxlo;          gb[0] -= giw * (1.f - sxlo);
    ga[3] += gih * syhi;          gb[3] += gih * (1.f - syhi);
    ga[1] -= gih * sylo;          gb[1] -= gih * (1.f - sylo);
    ga[2] += gA * ha; ga[0] -= gA * ha; ga[3] += gA * wa; ga[1] -= gA * wa;
    gb[2] += gA * hb; gb[0] -= gA * hb; gb[3] += gA * wb; gb[1] -= gA * wb;
}

__global__ void iou2d_bwd_list_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g,
                                      int M, float* __restrict__ grad_a, float* __restrict__ grad_b) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float ga[4] = {0, 0, 0, 0}, gb[4] = {0, 0, 0, 0};
    iou2_grad(reinterpret_cast<const float4*>(a)[i], reinterpret_cast<const float4*>(b)[i], g[i], ga, gb);
    reinterpret_cast<float4*>(grad_a)[i] = make_float4(ga[0], ga[1], ga[2], ga[3]);
    reinterpret_cast<float4*>(grad_b)[i] = make_float4(gb[0], gb[1], gb[2], gb[3]);
}

// combinations: one warp per output box; kForA: reduce over j for grad_a[i], else over i for grad_b[j].
template <bool kForA>
__global__ void iou2d_bwd_comb_kernel(const float* __restrict__ a, int M, const float* __restrict__ b, int N,
                                      const float* __restrict__ g, float* __restrict__ grad) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nout = kForA ? M : N, nred = kForA ? N : M;
    if (warp >= nout) return;
    const float4 mine = reinterpret_cast<const float4*>(kForA ? a : b)[warp];
    float acc[4] = {0, 0, 0, 0}, dummy[4] = {0, 0, 0, 0};
    for (int k = lane; k < nred; k += 32) {
        const float4 other = reinterpret_cast<const float4*>(kForA ? b : a)[k];
        const float gv = kForA ? g[(int64_t)warp * N + k] : g[(int64_t)k * N + warp];
        if (kForA) iou2_grad(mine, other, gv, acc, dummy);
        else iou2_grad(other, mine, gv, dummy, acc);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], d);
    if (lane == 0) reinterpret_cast<float4*>(grad)[warp] = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// lib/math_3d.py:47-72.  Row products accumulated left to right in fp32 (a matmul's order is backend-defined).
__global__ void project_kernel(const float* __restrict__ p2, const float* __restrict__ pts, int64_t n, int pad_ones,
                               float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = pts[i], y = pts[n + i], z = pts[2 * n + i], w = pad_ones ? 1.0f : pts[3 * n + i];
    float o[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float acc = __fmul_rn(p2[r * 4 + 0], x);
        acc = __fadd_rn(acc, __fmul_rn(p2[r * 4 + 1], y));
        acc = __fadd_rn(acc, __fmul_rn(p2[r * 4 + 2], z));
        acc = __fadd_rn(acc, __fmul_rn(p2[r * 4 + 3], w));
        o[r] = acc;
    }
    if (fabsf(o[2]) > 1e-2f) {
        o[0] = __fdiv_rn(o[0], o[2]);
        o[1] = __fdiv_rn(o[1], o[2]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r * n + i] = o[r];
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace gnms

using namespace gnms;

extern "C" int gnms_overlap2d_f32(const float* a, int M, const float* b, int N, float* out, int64_t ld_out,
                                  int kind, void* stream) {
    if (M < 0 || N < 0 || ld_out < N || (kind != GNMS_KIND_IOU && kind != GNMS_KIND_INTERSECT)) return GNMS_E_BADARG;
    if (M == 0 || N == 0) return 0;
    if (!a || !b || !out) return GNMS_E_BADARG;
    if (!aligned16(a) || !aligned16(b)) return GNMS_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(gnms_div_up(N, kColsPerCta), gnms_div_up(M, kRows));
    bool vec = aligned16(out) && (ld_out % 4 == 0);
    if (kind == GNMS_KIND_IOU && a == b && M == N && ld_out == N && N >= 2 * kST) {   // self-overlap: symmetric tiles
        const int nt = gnms_div_up(N, kST);
        overlap2d_self_kernel<<<dim3(nt * (nt + 1) / 2, 1), 256, 0, s>>>(a, N, out, nt, vec && (N % 4 == 0));
        GNMS_LAUNCH_CHECK();
        return 0;
    }
    if (kind == GNMS_KIND_IOU) {
        if (vec) overlap2d_kernel<GNMS_KIND_IOU, true><<<grid, kThreads, 0, s>>>(a, M, b, N, out, ld_out, 0, 0);
        else overlap2d_kernel<GNMS_KIND_IOU, false><<<grid, kThreads, 0, s>>>(a, M, b, N, out, ld_out, 0, 0);
    } else {
        if (vec) overlap2d_kernel<GNMS_KIND_INTERSECT, true><<<grid, kThreads, 0, s>>>(a, M, b, N, out, ld_out, 0, 0);
        else overlap2d_kernel<GNMS_KIND_INTERSECT, false><<<grid, kThreads, 0, s>>>(a, M, b, N, out, ld_out, 0, 0);
    }
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_overlap2d_list_f32(const float* a, const float* b, int M, float* out, int kind, void* stream) {
    if (M < 0 || (kind != GNMS_KIND_IOU && kind != GNMS_KIND_INTERSECT)) return GNMS_E_BADARG;
    if (M == 0) return 0;
    if (!a || !b || !out) return GNMS_E_BADARG;
    if (!aligned16(a) || !aligned16(b)) return GNMS_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    if (kind == GNMS_KIND_IOU) overlap2d_list_kernel<GNMS_KIND_IOU><<<gnms_div_up(M, 256), 256, 0, s>>>(a, b, M, out);
    else overlap2d_list_kernel<GNMS_KIND_INTERSECT><<<gnms_div_up(M, 256), 256, 0, s>>>(a, b, M, out);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_corners_from_boxes7_f32(const float* boxes7, int64_t ld, int N, float* corners, void* stream) {
    return gnms_corners_from_boxes7_ex_f32(boxes7, ld, N, 1, corners, stream);
}
extern "C" int gnms_corners_from_boxes7_ex_f32(const float* boxes7, int64_t ld, int N, int iou_3d_convention, float* corners, void* stream) {
    if (N < 0 || ld < 7) return GNMS_E_BADARG;
    if (N == 0) return 0;
    if (!boxes7 || !corners) return GNMS_E_BADARG;
    if (!aligned16(corners)) return GNMS_E_ALIGN;
    corners_kernel<<<gnms_div_up(N, 128), 128, 0, (cudaStream_t)stream>>>(boxes7, ld, N, corners, iou_3d_convention ? 0 : 1);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_iou2d_backward_f32(const float* a, int M, const float* b, int N, const float* g, int list_mode,
                                       float* grad_a, float* grad_b, void* stream) {
    if (M < 0 || N < 0 || (list_mode && M != N)) return GNMS_E_BADARG;
    if (M == 0 || N == 0) return 0;
    if (!a || !b || !g || !grad_a || !grad_b) return GNMS_E_BADARG;
    if (!aligned16(a) || !aligned16(b) || !aligned16(grad_a) || !aligned16(grad_b)) return GNMS_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    if (list_mode) {
        iou2d_bwd_list_kernel<<<gnms_div_up(M, 256), 256, 0, s>>>(a, b, g, M, grad_a, grad_b);
    } else {
        iou2d_bwd_comb_kernel<true><<<gnms_div_up(M, 8), 256, 0, s>>>(a, M, b, N, g, grad_a);
        iou2d_bwd_comb_kernel<false><<<gnms_div_up(N, 8), 256, 0, s>>>(a, M, b, N, g, grad_b);
    }
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_project_points_f32(const float* p2, const float* pts, int64_t n, int pad_ones, float* out,
                                       void* stream) {
    if (n < 0) return GNMS_E_BADARG;
    if (n == 0) return 0;
    if (!p2 || !pts || !out) return GNMS_E_BADARG;
    project_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p2, pts, n, pad_ones, out);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_box3d_records_from_boxes7_f32(const float* boxes7, int64_t ld, int N, float* rec, float* corners, void* stream) {
    if (N < 0 || ld < 7) return GNMS_E_BADARG;
    if (N == 0) return 0;
    if (!boxes7 || !rec) return GNMS_E_BADARG;
    if (!aligned16(rec) || (corners && !aligned16(corners))) return GNMS_E_ALIGN;
    records7_kernel<<<gnms_div_up(N, 128), 128, 0, (cudaStream_t)stream>>>(boxes7, ld, N, rec, corners);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_box3d_records_f32(float* corners, int N, float* rec, int mutate_input, void* stream) {
    if (N < 0) return GNMS_E_BADARG;
    if (N == 0) return 0;
    if (!corners || !rec) return GNMS_E_BADARG;
    if (!aligned16(corners) || !aligned16(rec)) return GNMS_E_ALIGN;
    records_kernel<<<gnms_div_up(N, 128), 128, 0, (cudaStream_t)stream>>>(corners, N, rec, mutate_input);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_overlap3d_batched_f32(const float* rec, int N, int batch, float* out_3d, int generalized, int affine,
                                          void* stream);

template <bool G, bool A>
static void launch_overlap3d(dim3 grid, cudaStream_t s, bool vec, const float* ra, int M, const float* rb, int N,
                             float* ob, float* o3, int64_t ld, const float* mul2d, int64_t rstride = 0,
                             int64_t ostride = 0) {
    if (vec) overlap3d_kernel<G, A, true><<<grid, kThreads, 0, s>>>(ra, M, rb, N, ob, o3, ld, mul2d, rstride, ostride);
    else overlap3d_kernel<G, A, false><<<grid, kThreads, 0, s>>>(ra, M, rb, N, ob, o3, ld, mul2d, rstride, ostride);
}

extern "C" int gnms_overlap3d_f32(const float* rec_a, int M, const float* rec_b, int N, float* out_bev,
                                  float* out_3d, int64_t ld_out, int generalized, int affine, const float* mul2d,
                                  void* stream) {
    if (M < 0 || N < 0 || ld_out < N) return GNMS_E_BADARG;
    if (M == 0 || N == 0) return 0;
    if (!rec_a || !rec_b || (!out_bev && !out_3d)) return GNMS_E_BADARG;
    if (!aligned16(rec_a) || !aligned16(rec_b)) return GNMS_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    if (rec_a == rec_b && M == N && !out_bev && !mul2d && ld_out == N && N >= 2 * kST)   // self-overlap: symmetric tiles
        return gnms_overlap3d_batched_f32(rec_a, N, 1, out_3d, generalized, affine, stream);
    dim3 grid(gnms_div_up(N, kColsPerCta), gnms_div_up(M, kRows));
    bool vec = (ld_out % 4 == 0) && (!out_bev || aligned16(out_bev)) && (!out_3d || aligned16(out_3d));
    if (generalized) {
        if (affine) launch_overlap3d<true, true>(grid, s, vec, rec_a, M, rec_b, N, out_bev, out_3d, ld_out, mul2d);
        else launch_overlap3d<true, false>(grid, s, vec, rec_a, M, rec_b, N, out_bev, out_3d, ld_out, mul2d);
    } else {
        if (affine) launch_overlap3d<false, true>(grid, s, vec, rec_a, M, rec_b, N, out_bev, out_3d, ld_out, mul2d);
        else launch_overlap3d<false, false>(grid, s, vec, rec_a, M, rec_b, N, out_bev, out_3d, ld_out, mul2d);
    }
    GNMS_LAUNCH_CHECK();
    return 0;
}

// Batched self-overlap: image z of boxes[B,N,4] / rec[B,N,8] -> out[B,N,N] in one launch (grid.z = image).
// (the large cases go to the matrix-only instantiation of the tile kernel in gnms.cu: same values bit for bit, faster stores)
int gnms_launch_overlap_tiles(const float* boxes, int src, int generalized, int affine, int N, int batch, float* out,
                              const GnmsLaunchOpts& O, cudaStream_t s);
static const int kTileSrcBox2d = 1, kTileSrcBox3d = 2;                // gnms::BoxSrc

extern "C" int gnms_overlap2d_batched_f32(const float* boxes, int N, int batch, float* out, void* stream) {
    return gnms_overlap2d_batched_ex_f32(boxes, N, batch, out, nullptr, stream);
}
extern "C" int gnms_overlap2d_batched_ex_f32(const float* boxes, int N, int batch, float* out, const gnms_launch_opts* opts, void* stream) {
    GnmsLaunchOpts O;
    const int orc = gnms_resolve_opts(opts, &O);
    if (orc) return orc;
    if (N < 0 || batch < 0) return GNMS_E_BADARG;
    if (N == 0 || batch == 0) return 0;
    if (!boxes || !out) return GNMS_E_BADARG;
    if (!aligned16(boxes)) return GNMS_E_ALIGN;
    if ((long long)N * batch >= 2048) return gnms_launch_overlap_tiles(boxes, kTileSrcBox2d, 0, 0, N, batch, out, O, (cudaStream_t)stream);
    const int nt = gnms_div_up(N, kST);
    const int vec = aligned16(out) && (N % 4 == 0);
    overlap2d_self_kernel<<<dim3(nt * (nt + 1) / 2, batch), 256, 0, (cudaStream_t)stream>>>(boxes, N, out, nt, vec);
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_overlap3d_batched_f32(const float* rec, int N, int batch, float* out_3d, int generalized, int affine,
                                          void* stream) {
    return gnms_overlap3d_batched_ex_f32(rec, N, batch, out_3d, generalized, affine, nullptr, stream);
}
extern "C" int gnms_overlap3d_batched_ex_f32(const float* rec, int N, int batch, float* out_3d, int generalized, int affine,
                                             const gnms_launch_opts* opts, void* stream) {
    GnmsLaunchOpts O;
    const int orc = gnms_resolve_opts(opts, &O);
    if (orc) return orc;
    if (N < 0 || batch < 0) return GNMS_E_BADARG;
    if (N == 0 || batch == 0) return 0;
    if (!rec || !out_3d) return GNMS_E_BADARG;
    if (!aligned16(rec)) return GNMS_E_ALIGN;
    if ((long long)N * batch >= 2048) return gnms_launch_overlap_tiles(rec, kTileSrcBox3d, generalized, affine, N, batch, out_3d, O, (cudaStream_t)stream);
    const int nt = gnms_div_up(N, kST);
    const int vec = aligned16(out_3d) && (N % 4 == 0);
    dim3 grid(nt * (nt + 1) / 2, batch);
    cudaStream_t s = (cudaStream_t)stream;
    if (generalized) {
        if (affine) overlap3d_self_kernel<true, true><<<grid, 256, 0, s>>>(rec, N, out_3d, nt, vec);
        else overlap3d_self_kernel<true, false><<<grid, 256, 0, s>>>(rec, N, out_3d, nt, vec);
    } else {
        if (affine) overlap3d_self_kernel<false, true><<<grid, 256, 0, s>>>(rec, N, out_3d, nt, vec);
        else overlap3d_self_kernel<false, false><<<grid, 256, 0, s>>>(rec, N, out_3d, nt, vec);
    }
    GNMS_LAUNCH_CHECK();
    return 0;
}

extern "C" int gnms_overlap3d_list_f32(const float* rec_a, const float* rec_b, int M, float* out_bev, float* out_3d,
                                       int generalized, int affine, void* stream) {
    if (M < 0) return GNMS_E_BADARG;
    if (M == 0) return 0;
    if (!rec_a || !rec_b || (!out_bev && !out_3d)) return GNMS_E_BADARG;
    if (!aligned16(rec_a) || !aligned16(rec_b)) return GNMS_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    int g = gnms_div_up(M, 256);
    if (generalized) {
        if (affine) overlap3d_list_kernel<true, true><<<g, 256, 0, s>>>(rec_a, rec_b, M, out_bev, out_3d);
        else overlap3d_list_kernel<true, false><<<g, 256, 0, s>>>(rec_a, rec_b, M, out_bev, out_3d);
    } else {
        if (affine) overlap3d_list_kernel<false, true><<<g, 256, 0, s>>>(rec_a, rec_b, M, out_bev, out_3d);
        else overlap3d_list_kernel<false, false><<<g, 256, 0, s>>>(rec_a, rec_b, M, out_bev, out_3d);
    }
    GNMS_LAUNCH_CHECK();
    return 0;
}
