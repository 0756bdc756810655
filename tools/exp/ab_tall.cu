// A/B of the matrix-only tile kernel variants through the C-ABI, no Python (starts in about a second):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../include -o ab_tall.bin ab_tall.cu \
//        -L../../groomed_nms_b200 -lgroomed_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../groomed_nms_b200'
//   ./ab_tall.bin [images=32]
// For matrix_kernel = DIRECT (register-direct STG) and TMA (shared-memory staging + 2-D tensor stores), each with CTAs that take
// one 256 x 64 tile and retire (tiles_per_cta = 4) and persistent CTAs (tiles_per_cta = 0): time of one launch over `images` x
// N=4096 7-DoF boxes (CUDA events, 10 launches after 2 warm-ups) and the number of output words that differ from the first
// variant's (must be 0).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "groomed_nms_b200.h"


#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
#define RC(x) do { int rc = (x); if (rc) { printf("gnms rc %d at line %d\n", rc, __LINE__); return 1; } } while (0)

__global__ void count_diff(const uint4* a, const uint4* b, size_t n4, unsigned long long* out) {
    unsigned long long d = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = a[i], y = b[i];
        d += (x.x != y.x) + (x.y != y.y) + (x.z != y.z) + (x.w != y.w);
    }
    if (d) atomicAdd(out, d);
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 32, N = 4096;
    std::vector<float> b7((size_t)B * N * 7);
    uint64_t s = 88172645463325252ull;
    auto u = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((s >> 11) * (1.0 / 9007199254740992.0)); };
    for (size_t i = 0; i < (size_t)B * N; ++i) {
        float* p = &b7[i * 7];
        p[0] = -30.f + 60.f * u(); p[1] = 1.4f + 0.4f * u(); p[2] = 5.f + 65.f * u();
        p[3] = 1.4f + 0.4f * u(); p[4] = 1.4f + 0.2f * u(); p[5] = 3.5f + 1.f * u(); p[6] = -3.14159f + 6.28318f * u();
    }
    float *d_b7, *d_rec, *d_ref, *d_out;
    unsigned long long* d_cnt;
    const size_t mat = (size_t)B * N * N;
    CK(cudaMalloc(&d_b7, b7.size() * 4)); CK(cudaMalloc(&d_rec, (size_t)B * N * 8 * 4));
    CK(cudaMalloc(&d_ref, mat * 4)); CK(cudaMalloc(&d_out, mat * 4)); CK(cudaMalloc(&d_cnt, 8));
    CK(cudaMemcpy(d_b7, b7.data(), b7.size() * 4, cudaMemcpyHostToDevice));
    RC(gnms_box3d_records_from_boxes7_f32(d_b7, 7, B * N, d_rec, nullptr, nullptr));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct V { const char* name; int kernel, tpc; };
    const V variants[4] = {{"direct, 1 tile/CTA", GNMS_MATRIX_KERNEL_DIRECT, 4}, {"tma,    1 tile/CTA", GNMS_MATRIX_KERNEL_TMA, 4},
                           {"direct, persistent", GNMS_MATRIX_KERNEL_DIRECT, 0}, {"tma,    persistent", GNMS_MATRIX_KERNEL_TMA, 0}};
    for (int vi = 0; vi < 4; ++vi) {
        gnms_launch_opts o = {};
        o.struct_size = sizeof(o); o.matrix_kernel = variants[vi].kernel; o.tiles_per_cta = variants[vi].tpc;
        float* out = vi == 0 ? d_ref : d_out;
        CK(cudaMemset(out, 0xff, mat * 4));
        for (int i = 0; i < 2; ++i) RC(gnms_overlap3d_batched_ex_f32(d_rec, N, B, out, 1, 1, &o, nullptr));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 10; ++i) RC(gnms_overlap3d_batched_ex_f32(d_rec, N, B, out, 1, 1, &o, nullptr));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long diff = 0;
        if (vi > 0) {
            CK(cudaMemset(d_cnt, 0, 8));
            count_diff<<<148 * 8, 256>>>(reinterpret_cast<const uint4*>(d_ref), reinterpret_cast<const uint4*>(d_out), mat / 4, d_cnt);
            CK(cudaMemcpy(&diff, d_cnt, 8, cudaMemcpyDeviceToHost));
        }
        printf("%s  %8.1f us per launch (%d images)  %7.1f GB/s of matrix  words differing from the first variant: %llu\n",
               variants[vi].name, ms * 100.f, B, mat * 4.0 / (ms * 1e-4) / 1e9, diff);
        fflush(stdout);
    }
    return 0;
}
