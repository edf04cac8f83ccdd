// micro-benchmark: FP64 / FP32 dependent-chain latency and throughput on this GPU (sizing the back-end)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dchain(double *out, int n, double a, double b) {
    double x = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { x = __dadd_rn(__dmul_rn(x, a), b); }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("DMUL+DADD dependent pair: %.1f cycles (threads %d x %d)\n", (double)(t1 - t0) / n, gridDim.x, blockDim.x);
}
__global__ void fchain(float *out, int n, float a, float b) {
    float x = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { x = __fadd_rn(__fmul_rn(x, a), b); }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("FMUL+FADD dependent pair: %.1f cycles\n", (double)(t1 - t0) / n);
}
__global__ void dthru(double *out, int n, double a, double b) {
    double x0 = out[0], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < n; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void ffma2thru(unsigned long long *out, int n, unsigned long long a) {
    unsigned long long x0 = out[0], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < n; i++) {
#define F(x) asm volatile("fma.rm.f32x2 %0, %0, %1, %0;" : "+l"(x) : "l"(a));
        F(x0) F(x1) F(x2) F(x3) F(x4) F(x5) F(x6) F(x7)
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    double *d; float *f; unsigned long long *u;
    cudaMalloc(&d, 1 << 26); cudaMalloc(&f, 1 << 26); cudaMalloc(&u, 1 << 26);
    cudaMemset(d, 0, 1 << 26); cudaMemset(f, 0, 1 << 26); cudaMemset(u, 0, 1 << 26);
    dchain<<<1, 1>>>(d, 20000, 0.999, 0.5); cudaDeviceSynchronize();
    dchain<<<1, 32>>>(d, 20000, 0.999, 0.5); cudaDeviceSynchronize();
    fchain<<<1, 1>>>(f, 20000, 0.999f, 0.5f); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); dthru<<<148 * 8, 256>>>(d, 4096, 0.999, 0.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA throughput: %.2f TFMA/s (%.3f ms)\n", 148.0 * 8 * 256 * 4096 * 8 / ms / 1e9, ms);
        cudaEventRecord(e0); ffma2thru<<<148 * 8, 256>>>(u, 4096, 0x3f7fbe773f7fbe77ull); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA2.RM throughput: %.2f T lane-FMA/s (%.3f ms)\n", 2 * 148.0 * 8 * 256 * 4096 * 8 / ms / 1e9, ms);
    }
    return 0;
}
