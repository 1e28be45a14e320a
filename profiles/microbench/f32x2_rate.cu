// Issue-rate probe for the non-fused multiply-add the bit-exact kernels need (sm_100a):
//   mode 0: scalar FMUL + FADD per lane (what stream_kernel.cuh runs today)
//   mode 1: packed FMUL2 + FFMA2(acc, 1.0 from a parameter, prod)  -- one rounding per op, as mode 0
//   mode 2: packed FMUL2 + two scalar FADD
//   mode 3: two scalar FMUL + FFMA2(acc, 1.0, prod)
// Every thread runs NACC independent float2 chains; taps are kernel parameters.
// Prints ns, lane-ops (mul or add on one float) per clock per SM, and a checksum (all modes equal).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 pk(float2 v) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
__device__ __forceinline__ float2 up(u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
constexpr int NT = 8, NACC = 8;
struct P { float2 tap[NT]; float2 one; const float2* x; float2* out; int iters; };

template <int MODE>
__global__ void __launch_bounds__(384, 1) k(const __grid_constant__ P p) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float2 acc[NACC], x[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) { acc[a] = make_float2(0.f, 0.f); x[a] = p.x[tid * NACC + a]; }
    const u64 one = pk(p.one);
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
                if (MODE == 0) {
                    acc[a].x = __fadd_rn(acc[a].x, __fmul_rn(x[a].x, p.tap[t].x));
                    acc[a].y = __fadd_rn(acc[a].y, __fmul_rn(x[a].y, p.tap[t].x));
                } else if (MODE == 1) {
                    acc[a] = up(fma2(pk(acc[a]), one, mul2(pk(x[a]), pk(p.tap[t]))));
                } else if (MODE == 2) {
                    const float2 pr = up(mul2(pk(x[a]), pk(p.tap[t])));
                    acc[a].x = __fadd_rn(acc[a].x, pr.x);
                    acc[a].y = __fadd_rn(acc[a].y, pr.y);
                } else {
                    const float2 pr = make_float2(__fmul_rn(x[a].x, p.tap[t].x), __fmul_rn(x[a].y, p.tap[t].x));
                    acc[a] = up(fma2(pk(acc[a]), one, pk(pr)));
                }
            }
        }
#pragma unroll
        for (int a = 0; a < NACC; ++a) x[a] = make_float2(acc[(a + 1) % NACC].y, acc[(a + 2) % NACC].x); // both lanes loop-carried: no product is loop invariant
    }
    float2 s = acc[0];
#pragma unroll
    for (int a = 1; a < NACC; ++a) { s.x += acc[a].x; s.y += acc[a].y; }
    p.out[tid] = s;
}

template <int MODE>
void run(P p, int blocks, int threads, double mhz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(p); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const int n = blocks * threads;
    float2* h = (float2*)malloc(n * sizeof(float2));
    cudaMemcpy(h, p.out, n * sizeof(float2), cudaMemcpyDeviceToHost);
    double cs = 0; for (int i = 0; i < n; ++i) cs += (double)h[i].x + (double)h[i].y;
    const double laneops = (double)n * 2 /*lanes*/ * 2 /*mul+add*/ * NT * NACC * p.iters;
    printf("{\"mode\": %d, \"threads_per_sm\": %d, \"ms\": %.4f, \"laneops_per_clk_per_sm\": %.1f, \"checksum\": %.9g, \"err\": \"%s\"}\n",
           MODE, threads, ms, laneops / (ms * 1e-3 * mhz * 1e6) / blocks, cs, cudaGetErrorString(cudaGetLastError()));
    free(h);
}

int main() {
    int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double mhz = khz / 1000.0;
    const int blocks = pr.multiProcessorCount;
    for (int threads : {128, 256, 384}) {
        const int n = blocks * threads;
        P p; for (int t = 0; t < NT; ++t) { const float v = ((t & 1) ? -0.01f : 0.01f) * (t / 2 * 2 + 1); p.tap[t] = make_float2(v, v); } // taps sum to zero: values stay bounded
        p.one = make_float2(1.0f, 1.0f); p.iters = 4000;
        float2* hx = (float2*)malloc((size_t)n * NACC * sizeof(float2));
        for (int i = 0; i < n * NACC; ++i) hx[i] = make_float2((i % 97) * 0.013f, (i % 89) * 0.017f);
        float2 *dx, *dout; cudaMalloc(&dx, (size_t)n * NACC * sizeof(float2)); cudaMalloc(&dout, n * sizeof(float2));
        cudaMemcpy(dx, hx, (size_t)n * NACC * sizeof(float2), cudaMemcpyHostToDevice);
        p.x = dx; p.out = dout;
        { run<0>(p, blocks, threads, mhz); run<1>(p, blocks, threads, mhz); run<2>(p, blocks, threads, mhz); run<3>(p, blocks, threads, mhz); }
        cudaFree(dx); cudaFree(dout); free(hx);
    }
    printf("{\"sm_count\": %d, \"clock_mhz\": %.0f}\n", blocks, mhz);
    return 0;
}
