// Experiment 2: the production mask_matrix_kernel in isolation with (a) zero data (b) random data (c) real-ish clustered data
#include "../../groomed_nms_b200/csrc/gnms.cu"
#include <cstdio>
#include <vector>
#include <numeric>
#include <random>
#include <algorithm>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
int main() {
    const int N = 4096, B = 8;
    float* iou; int* order; char* ws;
    WsLayout L = ws_layout(N);
    CK(cudaMalloc(&iou, (size_t)B * N * N * 4)); CK(cudaMalloc(&order, B * N * 4)); CK(cudaMalloc(&ws, L.total * B));
    CK(cudaMemset(ws, 0, L.total * B));
    std::vector<int> o(B * N), r(N);
    std::mt19937 g(1);
    for (int b = 0; b < B; ++b) {
        std::iota(o.begin() + b * N, o.begin() + (b + 1) * N, 0); std::shuffle(o.begin() + b * N, o.begin() + (b + 1) * N, g);
        for (int i = 0; i < N; ++i) r[o[b * N + i]] = i;
        CK(cudaMemcpy(ws + b * L.total + L.rank, r.data(), N * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(order, o.data(), B * N * 4, cudaMemcpyHostToDevice));
    bool do_clear = false;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid(N / 1024, N / 32, B);
    auto clear = [&]() { for (int b = 0; b < B; ++b) { cudaMemsetAsync(ws + b * L.total + L.has_earlier, 0, (N / 32) * L.he_slots * 4); } };
    auto run = [&](const char* name) {
        for (int i = 0; i < 3; ++i) mask_matrix_kernel<true><<<grid, 256>>>(iou, N, (int64_t)N * N, N, nullptr, order, ws, L.total, 0.4f);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) { if (do_clear) clear(); mask_matrix_kernel<true><<<grid, 256>>>(iou, N, (int64_t)N * N, N, nullptr, order, ws, L.total, 0.4f); }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        printf("%-40s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, (double)B * N * N * 4 / ms / 1e6);
    };
    CK(cudaMemset(iou, 0, (size_t)B * N * N * 4));
    run("zeros");
    std::vector<float> h((size_t)N * N);
    std::uniform_real_distribution<float> u(0, 1);
    for (auto& x : h) x = u(g) * 0.39f;                       // no bits set
    for (int b = 0; b < B; ++b) CK(cudaMemcpy(iou + (size_t)b * N * N, h.data(), (size_t)N * N * 4, cudaMemcpyHostToDevice));
    run("random < thr (no bits)");
    for (auto& x : h) x = u(g) < 0.03f ? 0.9f : 0.1f;         // 3% bits (like 128-box clusters in 4096)
    for (int b = 0; b < B; ++b) CK(cudaMemcpy(iou + (size_t)b * N * N, h.data(), (size_t)N * N * 4, cudaMemcpyHostToDevice));
    run("3% random bits");
    for (auto& x : h) x = 0.9f;
    for (int b = 0; b < B; ++b) CK(cudaMemcpy(iou + (size_t)b * N * N, h.data(), (size_t)N * N * 4, cudaMemcpyHostToDevice));
    run("all bits");
    // clustered: 32 clusters, random membership, 0.9 inside a cluster
    std::vector<int> cid(N);
    for (auto& c : cid) c = g() % 32;
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) h[(size_t)i * N + j] = cid[i] == cid[j] ? 0.9f : 0.1f;
    for (int b = 0; b < B; ++b) CK(cudaMemcpy(iou + (size_t)b * N * N, h.data(), (size_t)N * N * 4, cudaMemcpyHostToDevice));
    run("clustered, flags persist");
    do_clear = true;
    run("clustered, flags cleared per call");
    return 0;
}
