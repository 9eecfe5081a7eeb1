// Micro-benchmark: FP32 issue rates on sm_100a (scalar FFMA/FADD vs packed FFMA2/FADD2, operand patterns).
// Prints warp-instructions per clock per SM sub-partition and lane-ops per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fmas(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float subs(float a, float b) { float r; asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
constexpr int CH = 8, ITERS = 4096;
template <int MODE> __global__ void __launch_bounds__(512) k(const float* in, float* out, long long* clk) {
    float b = in[threadIdx.x], c = in[threadIdx.x + 512];
    float a[CH]; u64 A[CH];
    u64 B = ((u64)__float_as_uint(b) << 32) | __float_as_uint(c), Cc = ((u64)__float_as_uint(c) << 32) | __float_as_uint(b);
    for (int i = 0; i < CH; ++i) { a[i] = in[i]; A[i] = ((u64)__float_as_uint(in[i]) << 32) | __float_as_uint(in[i + 1]); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) a[i] = fmas(b, c, a[i]);                       // FFMA 3 distinct regs
            if (MODE == 1) A[i] = fma2(B, Cc, A[i]);                      // FFMA2 3 distinct pairs
            if (MODE == 2) a[i] = fmas(b, b, a[i]);                       // FFMA repeated operand
            if (MODE == 3) A[i] = fma2(B, B, A[i]);                       // FFMA2 repeated operand
            if (MODE == 4) a[i] = subs(a[i], b);                          // FADD
            if (MODE == 5) A[i] = sub2(A[i], B);                          // FADD2
            if (MODE == 6) { float d = subs(a[(i + 1) % CH], b); a[i] = fmas(d, d, a[i]); }      // scan pattern scalar (2 instr)
            if (MODE == 7) { u64 d = sub2(A[(i + 1) % CH], B); A[i] = fma2(d, d, A[i]); }        // scan pattern packed (2 instr)
            if (MODE == 8) a[i] = fmas(a[i], 1.0001f, 0.5f);              // FFMA imm forms
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < CH; ++i) s += a[i] + __uint_as_float((unsigned)A[i]) + __uint_as_float((unsigned)(A[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int instr_per, int lanes_per, const float* in, float* out, long long* clk, int threads) {
    k<MODE><<<148, threads>>>(in, out, clk); cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(in, out, clk); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    double winstr = (double)ITERS * CH * instr_per * (threads / 32);   // warp instructions per SM
    printf("%-34s threads %4d: %8.0f clk  %.3f warp-instr/clk/SMSP  %.1f lane-ops/clk/SM\n", name, threads, avg, winstr / avg / 4,
           winstr * 32 * lanes_per / instr_per / avg);
}
int main() {
    float *in, *out; long long* clk;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1.0f + i * 1e-3f; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    for (int threads : {128, 256, 512}) {
        run<0>("FFMA  a=fma(b,c,a)", 1, 1, in, out, clk, threads);
        run<1>("FFMA2 A=fma2(B,C,A)", 1, 2, in, out, clk, threads);
        run<2>("FFMA  a=fma(b,b,a)", 1, 1, in, out, clk, threads);
        run<3>("FFMA2 A=fma2(B,B,A)", 1, 2, in, out, clk, threads);
        run<4>("FADD  a=a-b", 1, 1, in, out, clk, threads);
        run<5>("FADD2 A=A-B", 1, 2, in, out, clk, threads);
        run<6>("scan scalar d=x-b; a=fma(d,d,a)", 2, 2, in, out, clk, threads);
        run<7>("scan packed D=X-B; A=fma2(D,D,A)", 2, 4, in, out, clk, threads);
        run<8>("FFMA imm a=fma(a,imm,imm)", 1, 1, in, out, clk, threads);
    }
    return 0;
}
