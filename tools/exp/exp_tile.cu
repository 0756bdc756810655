// Experiment: production tile_kernel in isolation (B images of N 3D records), with / without matrix output.
#include "../../groomed_nms_b200/csrc/gnms.cu"
#include <cstdio>
#include <vector>
#include <numeric>
#include <random>
#include <algorithm>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
int main(int argc, char** argv) {
    const int N = 4096, B = 8;
    int ctas_per_sm = argc > 1 ? atoi(argv[1]) : 3;
    WsLayout L = ws_layout(N);
    float *rec, *out; char* ws;
    CK(cudaMalloc(&rec, (size_t)B * N * 8 * 4)); CK(cudaMalloc(&out, (size_t)B * N * N * 4)); CK(cudaMalloc(&ws, L.total * B));
    CK(cudaMemset(ws, 0, L.total * B));
    std::mt19937 g(1);
    std::uniform_real_distribution<float> u(0, 1);
    std::normal_distribution<float> nd(0, 1);
    std::vector<float> h((size_t)B * N * 8);
    std::vector<int> r(N);
    for (int b = 0; b < B; ++b) {
        float cx[32], cz[32];
        for (int k = 0; k < 32; ++k) { cx[k] = -30 + 60 * u(g); cz[k] = 5 + 65 * u(g); }
        for (int i = 0; i < N; ++i) {
            int c = g() % 32;
            float x = cx[c] + 0.15f * nd(g), z = cz[c] + 0.15f * nd(g), y = 1.65f + 0.15f * nd(g);
            float w = 2.0f + 0.1f * nd(g), hh = 1.5f + 0.05f * nd(g), l = 4.0f + 0.2f * nd(g);
            float* p = &h[((size_t)b * N + i) * 8];
            p[0] = y - hh; p[1] = y; p[2] = x - l / 2; p[3] = x + l / 2; p[4] = z - w / 2; p[5] = z + w / 2;
            p[6] = l * hh * w; p[7] = l * w;
        }
        std::iota(r.begin(), r.end(), 0); std::shuffle(r.begin(), r.end(), g);
        CK(cudaMemcpy(ws + b * L.total + L.rank, r.data(), N * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(rec, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    TileArgs T = {};
    T.N = N; T.batch = B; T.nt = N / kTT; T.tiles_per_image = T.nt * (T.nt + 1) / 2; T.vec = 1; T.n_per_image = nullptr;
    T.boxes = rec; T.ws = ws; T.ws_img_stride = L.total; T.out = out; T.thr = 0.4f;
    const int total = T.tiles_per_image * B;
    const int grid = std::min(total, 148 * ctas_per_sm);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto fn) {
        for (int i = 0; i < 3; ++i) fn();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) fn();
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        printf("%-40s grid %4d  %8.1f us  (%.2f ns/pair-lane)\n", name, grid, ms * 1e3, ms * 1e6 / ((double)total * 4096));
    };
    run("tile 3D gen+affine, matrix out", [&] { tile_kernel<kSrcBox3d, true, true, true><<<grid, 256>>>(T); });
    run("tile 3D gen+affine, no out", [&] { tile_kernel<kSrcBox3d, true, true, false><<<grid, 256>>>(T); });
    T.thr = 2.0f;
    run("tile 3D, no out, no hits (thr=2)", [&] { tile_kernel<kSrcBox3d, true, true, false><<<grid, 256>>>(T); });
    CK(cudaDeviceSynchronize());
    return 0;
}
