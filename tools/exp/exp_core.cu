// Experiment: where does the pair-evaluation time go?  Stripped tile loops (no stores, no bits), all pairs of B x N records.
#include "../../groomed_nms_b200/csrc/common.cuh"
#include <cstdio>
#include <vector>
#include <random>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
using namespace gnms;

template <int V> __device__ __forceinline__ float pairv(const Rec3& a, const Rec3& b, bool& u) {
    if (V == 0) return iou3_fast<true, true>(a, b, inter_bev3(a, b), u);           // production
    if (V == 1) return iou3<true, true>(a, b, inter_bev3(a, b));                    // library division
    // V == 2: no divisions at all (rcp-multiply, wrong results) -> cost of everything but the division chains
    float ibev = inter_bev3(a, b);
    float yint = fmaxf(0.0f, __fsub_rn(fminf(a.ymax, b.ymax), fmaxf(a.ymin, b.ymin)));
    float i3d = __fmul_rn(ibev, yint);
    float un = __fsub_rn(__fadd_rn(a.vol, b.vol), i3d);
    float v = __fmul_rn(i3d, un);
    float xh = __fsub_rn(fmaxf(a.bx2, b.bx2), fminf(a.bx1, b.bx1));
    float yh = __fsub_rn(fmaxf(a.ymax, b.ymax), fminf(a.ymin, b.ymin));
    float zh = __fsub_rn(fmaxf(a.bz2, b.bz2), fminf(a.bz1, b.bz1));
    float vh = __fmul_rn(__fmul_rn(xh, yh), zh);
    v = __fsub_rn(v, __fmul_rn(__fsub_rn(vh, un), vh));
    return __fmul_rn(0.5f, __fadd_rn(1.0f, v));
}

// ROWS rows of 4 pairs evaluated together (ILP = 4*ROWS); tile 64 x 64, 256 threads, records from shared memory
template <int V, int ROWS, int MINB>
__global__ void __launch_bounds__(256, MINB) core(const float* __restrict__ rec, int N, int nt, int total, float* __restrict__ sink) {
    __shared__ __align__(16) float s_rec[2][64 * 8];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float accum = 0.f;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int b = t / (nt * nt), tt = t % (nt * nt), I = tt / nt, J = tt % nt;
        __syncthreads();
        if (tid < 128) {
            const int side = tid >> 6, k = tid & 63;
            const float4* src = reinterpret_cast<const float4*>(rec + ((size_t)b * N + (side ? J : I) * 64 + k) * 8);
            float4* dst = reinterpret_cast<float4*>(&s_rec[side][k * 8]);
            dst[0] = src[0]; dst[1] = src[1];
        }
        __syncthreads();
        Rec3 cr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) cr[k] = load_rec3(&s_rec[1][(4 * tx + k) * 8]);
        bool u = false;
#pragma unroll
        for (int r0 = 0; r0 < 4; r0 += ROWS) {
            Rec3 rr[ROWS];
#pragma unroll
            for (int q = 0; q < ROWS; ++q) rr[q] = load_rec3(&s_rec[0][(ty + 16 * (r0 + q)) * 8]);
#pragma unroll
            for (int q = 0; q < ROWS; ++q)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float v = pairv<V>(rr[q], cr[k], u);
                    accum += (v > 0.4f) ? 1.f : 0.f;
                }
        }
        if (u) accum += 1000.f;
    }
    if (accum == 12345.678f) sink[0] = accum;
}

int main() {
    const int N = 4096, B = 8, nt = N / 64;
    float *rec, *sink;
    CK(cudaMalloc(&rec, (size_t)B * N * 8 * 4)); CK(cudaMalloc(&sink, 4));
    std::mt19937 g(1);
    std::uniform_real_distribution<float> u(0, 1);
    std::normal_distribution<float> nd(0, 1);
    std::vector<float> h((size_t)B * N * 8);
    for (int b = 0; b < B; ++b) {
        float cx[32], cz[32];
        for (int k = 0; k < 32; ++k) { cx[k] = -30 + 60 * u(g); cz[k] = 5 + 65 * u(g); }
        for (int i = 0; i < N; ++i) {
            int c = g() % 32;
            float x = cx[c] + 0.15f * nd(g), z = cz[c] + 0.15f * nd(g), y = 1.65f + 0.15f * nd(g);
            float w = 2.0f + 0.1f * nd(g), hh = 1.5f + 0.05f * nd(g), l = 4.0f + 0.2f * nd(g);
            float* p = &h[((size_t)b * N + i) * 8];
            p[0] = y - hh; p[1] = y; p[2] = x - l / 2; p[3] = x + l / 2; p[4] = z - w / 2; p[5] = z + w / 2; p[6] = l * hh * w; p[7] = l * w;
        }
    }
    CK(cudaMemcpy(rec, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    const int total = B * nt * nt / 2;            // same pair count as the symmetric kernel (half of all tiles)
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto fn) {
        for (int i = 0; i < 2; ++i) fn();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) fn();
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-52s %8.1f us\n", name, ms * 1e3);
    };
    run("fast div, 1 row/iter, minb 3, grid 444", [&] { core<0, 1, 3><<<444, 256>>>(rec, N, nt, total, sink); });
    run("fast div, 2 rows/iter, minb 3", [&] { core<0, 2, 3><<<444, 256>>>(rec, N, nt, total, sink); });
    run("fast div, 4 rows/iter, minb 2", [&] { core<0, 4, 2><<<296, 256>>>(rec, N, nt, total, sink); });
    run("fast div, 1 row/iter, minb 4", [&] { core<0, 1, 4><<<592, 256>>>(rec, N, nt, total, sink); });
    run("fast div, 2 rows/iter, minb 4", [&] { core<0, 2, 4><<<592, 256>>>(rec, N, nt, total, sink); });
    run("library div, 1 row/iter, minb 3", [&] { core<1, 1, 3><<<444, 256>>>(rec, N, nt, total, sink); });
    run("no div, 1 row/iter, minb 3", [&] { core<2, 1, 3><<<444, 256>>>(rec, N, nt, total, sink); });
    run("no div, 2 rows/iter, minb 4", [&] { core<2, 2, 4><<<592, 256>>>(rec, N, nt, total, sink); });
    CK(cudaDeviceSynchronize());
    return 0;
}
