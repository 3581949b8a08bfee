// Microbenchmark: issue/throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a, alone and mixed with ALU-pipe work.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 r) { float2 d; asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r)); return d; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int ITERS = 4096, NCH = 8;

template <int MODE>   // 0: FFMA, 1: FFMA2, 2: FFMA + LOP3 (1:1), 3: FFMA2 + LOP3 (1:1), 4: FFMA2 + 2 LOP3
__global__ void k(float* out, float s, float t) {
    float a[NCH], b[NCH];
    u64 p[NCH];
    unsigned q[NCH];
    for (int i = 0; i < NCH; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = a[i] + 0.5f; p[i] = pk(a[i], b[i]); q[i] = threadIdx.x + i; }
    const u64 S = pk(s, s), T = pk(t, t);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0 || MODE == 2) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s), "f"(t)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(s), "f"(t)); }
            if (MODE == 1 || MODE == 3 || MODE == 4) p[i] = ffma2(p[i], S, T);
            if (MODE == 2) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(it), "r"(i * 77 + 1)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(it), "r"(i * 55 + 1)); }
            if (MODE == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(it), "r"(i * 77 + 1));
            if (MODE == 4) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(it), "r"(i * 77 + 1)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(it), "r"(i * 55 + 1)); }
        }
    }
    float r = 0;
    for (int i = 0; i < NCH; ++i) { float2 u = upk(p[i]); r += a[i] + b[i] + u.x + u.y + (float)q[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double fma_per_iter, double inst_per_iter) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 0.999f, 1e-3f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, 0.999f, 1e-3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double thr = 148.0 * 8 * 256, warps = thr / 32;
    const double fma = thr * ITERS * NCH * fma_per_iter, inst = warps * ITERS * NCH * inst_per_iter;
    printf("%-22s %8.3f ms  %7.2f TFLOP/s  %6.3f warp-inst/clk/SMSP (at 1.965 GHz)\n", name, ms, 2 * fma / ms * 1e-9, inst / (ms * 1e-3 * 1.965e9 * 148 * 4));
    cudaFree(out);
}

int main() {
    run<0>("FFMA", 2, 2);
    run<1>("FFMA2", 2, 1);
    run<2>("FFMA+LOP3 1:1", 2, 4);
    run<3>("FFMA2+LOP3 1:1", 2, 2);
    run<4>("FFMA2+2 LOP3", 2, 3);
    return 0;
}
