// What the HOST side of the box allows for the e2e call pattern, with nothing of ours in the way (no torch, no kernels):
// G GPUs at once, each fed by its own host thread from its own pinned buffers with bare cudaMemcpyAsync -- per step
// `h2d` bytes in and `d2h` bytes out on one stream, 3 buffer sets in flight like HostPipeline.  Prints aggregate and per-GPU
// GB/s for G = 1, 2, 4, ... up to the GPUs present: if the aggregate stops growing, the host memory / PCIe root complex is
// the wall for e2e scaling, not the library.
//   nvcc -O3 -o h2d_wall.bin h2d_wall.cu -lpthread ;  ./h2d_wall.bin [h2d_bytes=8388608] [d2h_bytes=2228736] [steps=400]
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <vector>
#include <atomic>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Result { double seconds; };

static void worker(int dev, size_t h2d, size_t d2h, int steps, std::atomic<int>* ready, std::atomic<int>* go, Result* out) {
    CK(cudaSetDevice(dev));
    const int depth = 3;
    void *hin[depth], *hout[depth], *din[depth], *dout[depth];
    cudaStream_t st[depth];
    for (int k = 0; k < depth; ++k) {
        CK(cudaHostAlloc(&hin[k], h2d, cudaHostAllocDefault)); CK(cudaHostAlloc(&hout[k], d2h, cudaHostAllocDefault));
        CK(cudaMalloc(&din[k], h2d)); CK(cudaMalloc(&dout[k], d2h));
        CK(cudaStreamCreate(&st[k]));
        for (size_t i = 0; i < h2d; i += 4096) ((char*)hin[k])[i] = 1;        // first touch: pages local to this thread's node
    }
    for (int i = 0; i < 10; ++i) {
        CK(cudaMemcpyAsync(din[i % depth], hin[i % depth], h2d, cudaMemcpyHostToDevice, st[i % depth]));
        CK(cudaMemcpyAsync(hout[i % depth], dout[i % depth], d2h, cudaMemcpyDeviceToHost, st[i % depth]));
    }
    CK(cudaDeviceSynchronize());
    ready->fetch_add(1);
    while (go->load() == 0) std::this_thread::yield();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; ++i) {
        const int k = i % depth;
        CK(cudaMemcpyAsync(din[k], hin[k], h2d, cudaMemcpyHostToDevice, st[k]));
        CK(cudaMemcpyAsync(hout[k], dout[k], d2h, cudaMemcpyDeviceToHost, st[k]));
    }
    CK(cudaDeviceSynchronize());
    out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv) {
    const size_t h2d = argc > 1 ? strtoull(argv[1], 0, 10) : 8388608, d2h = argc > 2 ? strtoull(argv[2], 0, 10) : 2228736;
    const int steps = argc > 3 ? atoi(argv[3]) : 400;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    printf("h2d %zu B + d2h %zu B per step, %d steps, 3 buffer sets in flight per GPU, %d GPU(s) present\n", h2d, d2h, steps, ndev);
    for (int g = 1; g <= ndev; g *= 2) {
        std::vector<std::thread> th;
        std::vector<Result> res(g);
        std::atomic<int> ready(0), go(0);
        for (int d = 0; d < g; ++d) th.emplace_back(worker, d, h2d, d2h, steps, &ready, &go, &res[d]);
        while (ready.load() < g) std::this_thread::yield();
        go.store(1);
        for (auto& t : th) t.join();
        double worst = 0, sum_h2d = 0;
        for (int d = 0; d < g; ++d) { worst = res[d].seconds > worst ? res[d].seconds : worst; sum_h2d += h2d * (double)steps / res[d].seconds / 1e9; }
        printf("G=%d  aggregate H2D %.1f GB/s (sum of per-GPU rates %.1f), per GPU %.1f GB/s, D2H aggregate %.1f GB/s, steps/s per GPU %.0f\n", g,
               g * h2d * (double)steps / worst / 1e9, sum_h2d, h2d * (double)steps / worst / 1e9, g * d2h * (double)steps / worst / 1e9, steps / worst);
    }
    return 0;
}
