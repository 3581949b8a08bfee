// Microbenchmark: device -> pinned-host copy by a kernel (16-byte stores over PCIe) vs cudaMemcpyAsync, 88.5 MB.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const float4* __restrict__ s, float4* __restrict__ d, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __ldcs(s + i);
}
int main() {
    const size_t bytes = 88510464, n = bytes / 16;
    float4 *h, *d;
    cudaMallocHost(&h, bytes); cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost);
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost); cudaEventRecord(b); cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b); printf("memcpy D2H: %.3f ms  %.1f GB/s\n", ms / 5, bytes / (ms / 5) / 1e6);
    const int grids[] = {8, 16, 32, 64, 148, 592}, thr[] = {128, 256, 512};
    for (int t : thr) for (int g : grids) {
        k<<<g, t>>>(d, h, n);
        cudaEventRecord(a); for (int i = 0; i < 5; ++i) k<<<g, t>>>(d, h, n); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b); printf("kernel %4d x %3d: %.3f ms  %.1f GB/s\n", g, t, ms / 5, bytes / (ms / 5) / 1e6);
    }
    return 0;
}
