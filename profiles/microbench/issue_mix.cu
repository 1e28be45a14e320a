// Issue-slot probe (sm_100a): do the non-FP instructions of the pass kernels' inner loops cost time
// next to the packed FP32 stream, or do they hide in the second cycle of every FMUL2 / FFMA2?
// Each thread runs NACC independent packed chains (FMUL2 + FFMA2 by a run-time 1, as the kernels
// do); per 16 packed instructions it additionally executes K integer adds (ALU pipe), K register
// moves kept alive by asm volatile, or K shared-memory loads.  Prints cycles per 16 packed
// instructions per scheduler for K = 0 .. 16 and 1, 2, 3 warps per scheduler.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
struct P { u64 tap[8]; u64 one; int iters; unsigned long long* out; int* sink; };

template <int KIND, int K>
__global__ void __launch_bounds__(384, 1) k(const __grid_constant__ P p) {
    __shared__ int sm[1024];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    u64 acc[8], x[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) { acc[a] = 0; x[a] = (u64)(threadIdx.x + a) * 0x3f8000013f800001ull; }
    int e0 = threadIdx.x, e1 = 1, e2 = 2, e3 = 3;
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int a = 0; a < 8; ++a) acc[a] = fma2(acc[a], p.one, mul2(x[a], p.tap[a]));  // 16 packed instructions
#pragma unroll
        for (int q = 0; q < K; ++q) {
            if (KIND == 0) { asm volatile("add.s32 %0, %0, %1;" : "+r"(e0) : "r"(e1)); }
            else if (KIND == 1) { asm volatile("mov.b32 %0, %1;" : "=r"(e2) : "r"(e3)); asm volatile("mov.b32 %0, %1;" : "=r"(e3) : "r"(e2)); }
            else { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[(threadIdx.x + q) & 1023]))); e0 ^= v; }
        }
    }
    const long long t1 = clock64();
    u64 s = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a) s ^= acc[a];
    if (threadIdx.x == 0 && blockIdx.x == 0) p.out[0] = (unsigned long long)(t1 - t0);
    if (s == 0x1234567ull) p.sink[0] = e0 + e2 + e3;
}

template <int KIND, int K>
void run(P p, int threads, const char* kind) {
    k<KIND, K><<<148, threads>>>(p);
    cudaDeviceSynchronize();
    k<KIND, K><<<148, threads>>>(p);
    cudaDeviceSynchronize();
    unsigned long long cyc = 0;
    cudaMemcpy(&cyc, p.out, 8, cudaMemcpyDeviceToHost);
    const int warps_per_sched = threads / 128;
    printf("{\"extra\": \"%s\", \"k_per_16_packed\": %d, \"warps_per_scheduler\": %d, \"cycles_per_16_packed_per_scheduler\": %.2f}\n",
           kind, KIND == 1 ? 2 * K : K, warps_per_sched, (double)cyc / p.iters / 1.0 / 1.0 / 1.0 / 1.0 * 1.0 / warps_per_sched * 1.0);
}

template <int KIND>
void sweep(P p, const char* kind) {
    for (int threads : {128, 256, 384}) {
        run<KIND, 0>(p, threads, kind); run<KIND, 2>(p, threads, kind); run<KIND, 4>(p, threads, kind);
        run<KIND, 8>(p, threads, kind); run<KIND, 16>(p, threads, kind);
    }
}

int main() {
    P p;
    for (int t = 0; t < 8; ++t) { float v = (t & 1) ? -0.01f : 0.01f; unsigned u; memcpy(&u, &v, 4); p.tap[t] = ((u64)u << 32) | u; }
    { float v = 1.0f; unsigned u; memcpy(&u, &v, 4); p.one = ((u64)u << 32) | u; }
    p.iters = 20000;
    cudaMalloc(&p.out, 8); cudaMalloc(&p.sink, 4);
    sweep<0>(p, "iadd"); sweep<1>(p, "mov"); sweep<2>(p, "lds");
    printf("{\"note\": \"cycles per scheduler for one warp's 16 packed instructions; 32 = FP32 pipe bound (2 cycles each)\"}\n");
    return 0;
}
