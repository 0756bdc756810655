// Experiment: how fast can the tile kernel's STORE pattern alone run?  64 x 64 tiles of a [B][N][N] fp32 matrix, each
// written as a direct tile and a mirrored tile straight from registers (4 x 4 per thread, warp = 8 x 4 quads), no math.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_store.bin exp_store.cu && ./exp_store.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

template <int POLICY> __device__ __forceinline__ void st4(float* p, float4 v) {
    if (POLICY == 0) __stcs(reinterpret_cast<float4*>(p), v);
    else if (POLICY == 1) *reinterpret_cast<float4*>(p) = v;
    else if (POLICY == 2) __stcg(reinterpret_cast<float4*>(p), v);
    else __stwt(reinterpret_cast<float4*>(p), v);
}

// MODE 0: direct + mirrored (symmetric tiles, upper triangle); MODE 1: direct only, all tiles; MODE 2: mirrored only, all tiles
template <int POLICY, int MODE, int SPIN>
__global__ void __launch_bounds__(256, 2) k(float* out, int N, int nt, int total) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = (lane & 7) + 8 * (warp & 1), ty = (lane >> 3) + 4 * (warp >> 1);
    const int tpi = MODE == 0 ? nt * (nt + 1) / 2 : nt * nt;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int b = t / tpi, tt = t - b * tpi;
        int I, J;
        if (MODE == 0) {
            const int i = tt / (nt + 1), j = tt - i * (nt + 1);
            if (j < nt - i) { I = i; J = i + j; } else { I = nt - 1 - i; J = I + (j - (nt - i)); }
        } else { I = tt / nt; J = tt - I * nt; }
        float v = (float)(t + tid);
        // stand-in for the arithmetic between two tiles' stores: SPIN dependent FMAs per thread
#pragma unroll 1
        for (int s = 0; s < SPIN; ++s) v = v * 1.0001f + 0.5f;
        float* o = out + (size_t)b * N * N;
        if (MODE != 2) {
            float* drow = o + (size_t)(I * 64 + 4 * ty) * N + J * 64 + 4 * tx;
#pragma unroll
            for (int r = 0; r < 4; ++r) st4<POLICY>(drow + (size_t)r * N, make_float4(v, v + 1, v + 2, v + 3));
        }
        if (MODE == 2 || (MODE == 0 && I != J)) {
            float* dcol = o + (size_t)(J * 64 + 4 * tx) * N + I * 64 + 4 * ty;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) st4<POLICY>(dcol + (size_t)kk * N, make_float4(v, v + 1, v + 2, v + 3));
        }
    }
}

int main() {
    const int N = 4096, B = 16, nt = N / 64;
    float* out;
    CK(cudaMalloc(&out, (size_t)B * N * N * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto fn) {
        for (int i = 0; i < 2; ++i) fn();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) fn();
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-64s %8.1f us  %7.0f GB/s\n", name, ms * 1e3, (double)B * N * N * 4 / ms / 1e6);
    };
    const int sym = B * nt * (nt + 1) / 2, all = B * nt * nt;
    run("memset (cudaMemsetAsync)", [&] { cudaMemsetAsync(out, 0, (size_t)B * N * N * 4); });
    run("direct+mirror  st.cs   no math      grid 296", [&] { k<0, 0, 0><<<296, 256>>>(out, N, nt, sym); });
    run("direct+mirror  st      no math      grid 296", [&] { k<1, 0, 0><<<296, 256>>>(out, N, nt, sym); });
    run("direct+mirror  st.cg   no math      grid 296", [&] { k<2, 0, 0><<<296, 256>>>(out, N, nt, sym); });
    run("direct+mirror  st.wt   no math      grid 296", [&] { k<3, 0, 0><<<296, 256>>>(out, N, nt, sym); });
    run("direct only    st.cs   no math      grid 296", [&] { k<0, 1, 0><<<296, 256>>>(out, N, nt, all); });
    run("direct only    st      no math      grid 296", [&] { k<1, 1, 0><<<296, 256>>>(out, N, nt, all); });
    run("mirror only    st.cs   no math      grid 296", [&] { k<0, 2, 0><<<296, 256>>>(out, N, nt, all); });
    run("mirror only    st      no math      grid 296", [&] { k<1, 2, 0><<<296, 256>>>(out, N, nt, all); });
    run("direct+mirror  st.cs   200 FMA/tile grid 296", [&] { k<0, 0, 200><<<296, 256>>>(out, N, nt, sym); });
    run("direct+mirror  st      200 FMA/tile grid 296", [&] { k<1, 0, 200><<<296, 256>>>(out, N, nt, sym); });
    run("direct+mirror  st.cs   800 FMA/tile grid 296", [&] { k<0, 0, 800><<<296, 256>>>(out, N, nt, sym); });
    run("direct+mirror  st      800 FMA/tile grid 296", [&] { k<1, 0, 800><<<296, 256>>>(out, N, nt, sym); });
    run("direct+mirror  st.cs   no math      grid 592", [&] { k<0, 0, 0><<<592, 256>>>(out, N, nt, sym); });
    CK(cudaDeviceSynchronize());
    return 0;
}
