// Experiment: issue rates of the instruction classes the overlap tile kernel is made of, on B200 (sm_100a).
// Every test runs 8 independent dependency chains per thread, 8 warps per SM sub-partition, one wave of CTAs, and reports
// warp-instructions per cycle per SM sub-partition (clock64 deltas, averaged over CTAs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_pipes exp_pipes.cu && ./exp_pipes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

#define OP1(name, text) __device__ __forceinline__ void name(float& x, float y, float z) { asm volatile(text : "+f"(x) : "f"(y), "f"(z)); }
OP1(op_ffma, "fma.rn.f32 %0, %0, %1, %2;")
OP1(op_fmul, "mul.rn.f32 %0, %0, %1;")
OP1(op_fadd, "add.rn.f32 %0, %0, %2;")
OP1(op_fmin, "min.f32 %0, %0, %1;")
OP1(op_fmax, "max.f32 %0, %0, %2;")
OP1(op_rcp, "rcp.approx.ftz.f32 %0, %0;")
__device__ __forceinline__ void op_iadd(float& x, float y, float z) { asm volatile("{.reg .b32 t; mov.b32 t, %0; add.s32 t, t, 12345; mov.b32 %0, t;}" : "+f"(x)); }
__device__ __forceinline__ void op_lop(float& x, float y, float z) { asm volatile("{.reg .b32 t, u; mov.b32 t, %0; mov.b32 u, %1; xor.b32 t, t, u; mov.b32 %0, t;}" : "+f"(x) : "f"(y)); }
__device__ __forceinline__ void op_fmin3(float& x, float y, float z) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(y), "f"(z)); }
__device__ __forceinline__ void op_setp(float& x, float y, float z) { asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %0, %2, p;}" : "+f"(x) : "f"(y), "f"(z)); }

// MODE: which op sequence makes one "round" over the 8 chains
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float y, float z, int iters) {
    float a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 1.0f + 0.001f * (threadIdx.x + q);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll (MODE == 14 ? 1 : 4)
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (MODE == 0) op_ffma(a[q], y, z);
                if (MODE == 1) op_fmul(a[q], y, z);
                if (MODE == 2) op_fadd(a[q], y, z);
                if (MODE == 3) { if (rep & 1) op_fmax(a[q], y, z); else op_fmin(a[q], y, z); }
                if (MODE == 4) op_rcp(a[q], y, z);
                if (MODE == 5) op_iadd(a[q], y, z);
                if (MODE == 6) op_lop(a[q], y, z);
                if (MODE == 7) { op_ffma(a[q], y, z); op_fmin(a[q], y, z); }                       // 1 fma : 1 alu
                if (MODE == 8) { op_fadd(a[q], y, z); op_fmin(a[q], y, z); }
                if (MODE == 9) { op_ffma(a[q], y, z); op_ffma(a[q], z, y); op_fmin(a[q], y, z); }   // 2 fma : 1 alu
                if (MODE == 10) { op_fmin(a[q], y, z); op_fmax(a[q], y, z); op_fadd(a[q], y, z); }  // 2 alu : 1 fma
                if (MODE == 11) op_fmin3(a[q], y, z);
                if (MODE == 12) op_setp(a[q], y, z);
                if (MODE == 13) { op_fmul(a[q], y, z); op_fadd(a[q], y, z); }
                if (MODE == 14) {   // the tile kernel's mix per pair, scaled down: 15 minmax, 14 add, 7 mul, 10 fma, 2 rcp
                    op_fmin(a[q], y, z); op_fmax(a[q], y, z); op_fadd(a[q], y, z); op_fmax(a[q], y, z);
                    op_fmin(a[q], y, z); op_fmax(a[q], y, z); op_fadd(a[q], y, z); op_fmax(a[q], y, z); op_fmul(a[q], y, z);
                    op_fmin(a[q], y, z); op_fmax(a[q], y, z); op_fadd(a[q], y, z); op_fmax(a[q], y, z); op_fmul(a[q], y, z);
                    op_fadd(a[q], y, z); op_fadd(a[q], y, z);
                    op_rcp(a[q], y, z); op_ffma(a[q], y, z); op_ffma(a[q], y, z); op_ffma(a[q], y, z); op_ffma(a[q], y, z); op_ffma(a[q], y, z);
                    op_fmax(a[q], y, z); op_fmin(a[q], y, z); op_fadd(a[q], y, z);
                    op_fmax(a[q], y, z); op_fmin(a[q], y, z); op_fadd(a[q], y, z);
                    op_fmax(a[q], y, z); op_fmin(a[q], y, z); op_fadd(a[q], y, z);
                    op_fmul(a[q], y, z); op_fmul(a[q], y, z); op_fadd(a[q], y, z);
                    op_rcp(a[q], y, z); op_ffma(a[q], y, z); op_ffma(a[q], y, z); op_ffma(a[q], y, z); op_ffma(a[q], y, z); op_ffma(a[q], y, z);
                    op_fadd(a[q], y, z); op_fadd(a[q], y, z); op_fmul(a[q], y, z);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += a[q];
    if (s == 12345.678f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int ops_per_round, float* out, long long* cyc, int ctas_per_sm) {
    const int ctas = 148 * ctas_per_sm, iters = 200;
    k<MODE><<<ctas, 256>>>(out, cyc, 1.0001f, 0.9999f, iters);
    CK(cudaDeviceSynchronize());
    k<MODE><<<ctas, 256>>>(out, cyc, 1.0001f, 0.9999f, iters);
    CK(cudaDeviceSynchronize());
    static long long h[148 * 8];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < ctas; ++i) avg += (double)h[i];
    avg /= ctas;
    // per SMSP: ctas_per_sm * 8 warps / 4 = 2 * ctas_per_sm warps, each issuing iters*4*8*ops_per_round instructions
    const double winst = 2.0 * ctas_per_sm * (double)iters * 4 * 8 * ops_per_round;
    printf("%-44s %d warps/SMSP  %6.3f warp-inst/clk/SMSP\n", name, 2 * ctas_per_sm, winst / avg);
}

int main() {
    float* out; long long* cyc;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, sizeof(long long) * 148 * 8));
    for (int c = 2; c <= 4; c += 2) {
        run<0>("FFMA (3 reg)", 1, out, cyc, c);
        run<1>("FMUL", 1, out, cyc, c);
        run<2>("FADD", 1, out, cyc, c);
        run<3>("FMNMX", 1, out, cyc, c);
        run<11>("FMNMX3 (3-input min)", 1, out, cyc, c);
        run<4>("MUFU.RCP", 1, out, cyc, c);
        run<5>("IADD", 1, out, cyc, c);
        run<6>("LOP3", 1, out, cyc, c);
        run<12>("FSETP+FSEL", 2, out, cyc, c);
        run<7>("FFMA + FMNMX (1:1)", 2, out, cyc, c);
        run<8>("FADD + FMNMX (1:1)", 2, out, cyc, c);
        run<13>("FMUL + FADD (1:1)", 2, out, cyc, c);
        run<9>("2 FFMA + FMNMX", 3, out, cyc, c);
        run<10>("2 FMNMX + FADD", 3, out, cyc, c);
        run<14>("tile-kernel mix (50 ops/pair)", 50, out, cyc, c);
    }
    return 0;
}
