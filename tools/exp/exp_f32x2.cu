// Experiment: does Blackwell's packed fp32x2 arithmetic (add/mul/fma.rn.f32x2) speed up the pair evaluation of the tile
// kernel?  Two column boxes are evaluated against one row box per call, scalar vs packed, register resident; the packed
// version must be bitwise equal (every op is still an individually rounded IEEE fp32 op).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o exp_f32x2.bin exp_f32x2.cu && ./exp_f32x2.bin
#include "../../groomed_nms_b200/csrc/common.cuh"
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
using namespace gnms;

struct F2 { float x, y; };
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
    F2 r;
    asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ F2 sub2(F2 a, F2 b) { return add2(a, F2{-b.x, -b.y}); }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
    F2 r;
    asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmul.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    F2 r;
    asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\tfma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ F2 min2(F2 a, F2 b) { return F2{fminf(a.x, b.x), fminf(a.y, b.y)}; }
__device__ __forceinline__ F2 max2(F2 a, F2 b) { return F2{fmaxf(a.x, b.x), fmaxf(a.y, b.y)}; }
__device__ __forceinline__ F2 bc(float v) { return F2{v, v}; }
__device__ __forceinline__ F2 div2_fast(F2 a, F2 b) {
    F2 r = F2{rcp_approx(b.x), rcp_approx(b.y)};
    const F2 nb = F2{-b.x, -b.y};
    const F2 e = fma2(nb, r, bc(1.0f));
    r = fma2(r, e, r);
    const F2 q = fma2(a, r, bc(0.0f));
    const F2 m = fma2(nb, q, a);
    return fma2(r, m, q);
}
// two pairs at once: row box a against column boxes b0, b1 (generalized, affine)
__device__ __forceinline__ F2 iou3_fast2(const Rec3& a, const Rec3& b0, const Rec3& b1, bool& unsafe) {
    const F2 bx1 = F2{b0.bx1, b1.bx1}, bx2 = F2{b0.bx2, b1.bx2}, bz1 = F2{b0.bz1, b1.bz1}, bz2 = F2{b0.bz2, b1.bz2};
    const F2 ymin = F2{b0.ymin, b1.ymin}, ymax = F2{b0.ymax, b1.ymax}, vol = F2{b0.vol, b1.vol};
    const F2 iw = max2(sub2(min2(bc(a.bx2), bx2), max2(bc(a.bx1), bx1)), bc(0.f));
    const F2 ih = max2(sub2(min2(bc(a.bz2), bz2), max2(bc(a.bz1), bz1)), bc(0.f));
    const F2 ibev = mul2(iw, ih);
    const F2 yint = max2(bc(0.f), sub2(min2(bc(a.ymax), ymax), max2(bc(a.ymin), ymin)));
    const F2 i3d = mul2(ibev, yint);
    const F2 un = sub2(add2(bc(a.vol), vol), i3d);
    unsafe = unsafe || tiny_nonzero(i3d.x) || tiny_nonzero(i3d.y);
    F2 v = div2_fast(i3d, un);
    const F2 xh = sub2(max2(bc(a.bx2), bx2), min2(bc(a.bx1), bx1));
    const F2 yh = sub2(max2(bc(a.ymax), ymax), min2(bc(a.ymin), ymin));
    const F2 zh = sub2(max2(bc(a.bz2), bz2), min2(bc(a.bz1), bz1));
    const F2 vh = mul2(mul2(xh, yh), zh);
    v = sub2(v, div2_fast(sub2(vh, un), vh));
    return mul2(bc(0.5f), add2(bc(1.0f), v));
}

template <int MODE>
__global__ void __launch_bounds__(128, 4) k(const float* __restrict__ rec, int n, float* __restrict__ out, int iters, unsigned* diff) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Rec3 rr[4], cr[4];
    for (int q = 0; q < 4; ++q) { rr[q] = load_rec3(rec + (size_t)((tid * 4 + q) % n) * 8); cr[q] = load_rec3(rec + (size_t)((tid * 7 + 3 * q + 1) % n) * 8); }
    float acc = 0.f;
    unsigned nd = 0;
    for (int it = 0; it < iters; ++it) {
        float v[4][4];
        bool u = false;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int c = 0; c < 4; ++c) v[r][c] = iou3_fast<true, true>(rr[r], cr[c], inter_bev3(rr[r], cr[c]), u);
            }
            if (MODE == 1 || MODE == 2) {
#pragma unroll
                for (int c = 0; c < 4; c += 2) {
                    const F2 p = iou3_fast2(rr[r], cr[c], cr[c + 1], u);
                    if (MODE == 2) { nd += (__float_as_uint(p.x) != __float_as_uint(v[r][c])) + (__float_as_uint(p.y) != __float_as_uint(v[r][c + 1])); }
                    v[r][c] = p.x; v[r][c + 1] = p.y;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc += v[r][c];
        if (u) acc += 1.f;
        // new operands next round (cheap, keeps the compiler from hoisting)
#pragma unroll
        for (int q = 0; q < 4; ++q) { cr[q].bx1 += 0.01f; cr[q].bx2 += 0.01f; rr[q].bz1 -= 0.005f; rr[q].bz2 -= 0.005f; }
    }
    out[tid] = acc;
    if (MODE == 2 && nd) atomicAdd(diff, nd);
}

int main() {
    const int n = 4096, iters = 2000;
    float* h = (float*)malloc(n * 8 * 4);
    srand(1);
    for (int i = 0; i < n; ++i) {
        float x = -30 + 60.f * rand() / RAND_MAX, z = 5 + 65.f * rand() / RAND_MAX, y = 1.65f;
        float w = 2.f + 0.2f * rand() / RAND_MAX, hh = 1.5f, l = 4.f + 0.4f * rand() / RAND_MAX;
        float* p = h + i * 8;
        p[0] = y - hh; p[1] = y; p[2] = x - l / 2; p[3] = x + l / 2; p[4] = z - w / 2; p[5] = z + w / 2; p[6] = l * hh * w; p[7] = l * w;
    }
    float *rec, *out; unsigned* diff;
    CK(cudaMalloc(&rec, n * 8 * 4)); CK(cudaMalloc(&out, 592 * 128 * 4)); CK(cudaMalloc(&diff, 4));
    CK(cudaMemcpy(rec, h, n * 8 * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(diff, 0, 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto fn) {
        fn(); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); fn(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double pairs = 592.0 * 128 * 16 * iters;
        printf("%-28s %8.1f us   %6.2f G pairs/s  (tile kernel needs 8.52 M pairs per image)\n", name, ms * 1e3, pairs / ms / 1e6);
    };
    run("scalar", [&] { k<0><<<592, 128>>>(rec, n, out, iters, diff); });
    run("packed f32x2", [&] { k<1><<<592, 128>>>(rec, n, out, iters, diff); });
    k<2><<<592, 128>>>(rec, n, out, 50, diff);
    CK(cudaDeviceSynchronize());
    unsigned hd; CK(cudaMemcpy(&hd, diff, 4, cudaMemcpyDeviceToHost));
    printf("bitwise differences packed vs scalar: %u\n", hd);
    return 0;
}
