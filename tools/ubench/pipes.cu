// micro-benchmarks that size the front-end: which pipes can share the work of one raw IQ sample
//   I2F.S8 (byte -> float on the conversion pipe), FFMA2.RM, DFMA, PRMT/IADD and mixes of them
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define F2(x, a) asm volatile("fma.rm.f32x2 %0, %0, %1, %0;" : "+l"(x) : "l"(a));
#define I2F(f, w, b) asm volatile("{.reg .b8 x<4>; mov.b32 {x0,x1,x2,x3}, %1; cvt.rn.f32.s8 %0, x" #b ";}" : "=f"(f) : "r"(w));
__global__ void k_ffma2(u64 *out, int n, u64 a) {
    u64 x0 = out[0], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < n; i++) { F2(x0, a) F2(x1, a) F2(x2, a) F2(x3, a) F2(x4, a) F2(x5, a) F2(x6, a) F2(x7, a) }
    out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
// 8 I2F per iteration
__global__ void k_i2f(float *out, int n, uint32_t w) {
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    uint32_t v = w + threadIdx.x;
    for (int i = 0; i < n; i++) {
        float a, b, c, d, e, f, g, h;
        I2F(a, v, 0) I2F(b, v, 1) I2F(c, v, 2) I2F(d, v, 3)
        uint32_t v2 = v ^ 0x55aa55aau;
        I2F(e, v2, 0) I2F(f, v2, 1) I2F(g, v2, 2) I2F(h, v2, 3)
        s0 += a + e; s1 += b + f; s2 += c + g; s3 += d + h;   // 8 FADD per 8 I2F (keeps them live)
        v = v * 1664525u + 1013904223u;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = s0 + s1 + s2 + s3;
}
// the real mix: per 2 IQ samples (one 32-bit word): K of the 4 bytes via I2F, rest via PRMT+IADD; 18 FFMA2
template <int K>
__global__ void k_mix(u64 *out, int n, uint32_t w, u64 a) {
    u64 x0 = out[0], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5;
    uint32_t v = w + threadIdx.x;
    uint32_t acc = 0;
    for (int i = 0; i < n; i++) {
        float f0, f1, f2, f3;
        uint32_t vx = v ^ 0x80808080u;
        if (K >= 1) { I2F(f0, vx, 0) } else f0 = __uint_as_float(__byte_perm(v, 0, 0x4404) + 0x47038000u);
        if (K >= 2) { I2F(f2, vx, 2) } else f2 = __uint_as_float(__byte_perm(v, 0, 0x4424) + 0x47038000u);
        if (K >= 3) { I2F(f1, vx, 1) } else f1 = __uint_as_float(__byte_perm(v, 0, 0x4414) + 0x47038000u);
        if (K >= 4) { I2F(f3, vx, 3) } else f3 = __uint_as_float(__byte_perm(v, 0, 0x4434) + 0x47038000u);
        u64 p0, p1;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(f0), "f"(f1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(f2), "f"(f3));
        // 18 FFMA2 for two IQ samples, six independent chains
#define G(x, p) asm volatile("fma.rm.f32x2 %0, %1, %2, %0;" : "+l"(x) : "l"(p), "l"(a));
        G(x0, p0) G(x1, p1) G(x2, p0) G(x3, p1) G(x4, p0) G(x5, p1)
        G(x0, p1) G(x1, p0) G(x2, p1) G(x3, p0) G(x4, p1) G(x5, p0)
        G(x0, p0) G(x1, p1) G(x2, p0) G(x3, p1) G(x4, p0) G(x5, p1)
        acc += (uint32_t)(x0 >> 40);
        v = v * 1664525u + 1013904223u;   // IMAD on the fma pipe
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + x4 + x5 + acc;
}
__global__ void k_dfma(double *out, int n, double a, double b) {
    double x0 = out[0], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < n; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
// DFMA and FFMA2 together: do they overlap?
__global__ void k_dfma_ffma2(double *out, int n, double a, double b, u64 c) {
    double x0 = out[0], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    u64 y0 = (u64)out[1], y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3, y4 = y0 + 4, y5 = y0 + 5, y6 = y0 + 6, y7 = y0 + 7;
    for (int i = 0; i < n; i++) {
        x0 = fma(x0, a, b); F2(y0, c) F2(y1, c) x1 = fma(x1, a, b); F2(y2, c) F2(y3, c)
        x2 = fma(x2, a, b); F2(y4, c) F2(y5, c) x3 = fma(x3, a, b); F2(y6, c) F2(y7, c)
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + (double)(y0 + y1 + y2 + y3 + y4 + y5 + y6 + y7);
}
template <class F> static float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    void *d; cudaMalloc(&d, 1 << 26); cudaMemset(d, 0, 1 << 26);
    const int G = 148 * 8, T = 256, N = 4096;
    const double thr = (double)G * T * N;
    float ms;
    ms = timeit([&] { k_ffma2<<<G, T>>>((u64 *)d, N, 0x3f7fbe773f7fbe77ull); });
    printf("FFMA2.RM        : %.2f T lane-FMA/s  (%.1f warp-instr/clk/SM)\n", 2 * thr * 8 / ms / 1e9, thr * 8 / 32 / (ms * 1e-3) / 148 / 1.965e9);
    ms = timeit([&] { k_i2f<<<G, T>>>((float *)d, N, 12345u); });
    printf("I2F.S8 (+FADD)  : %.2f T cvt/s  (%.2f warp-instr/clk/SM)\n", thr * 8 / ms / 1e9, thr * 8 / 32 / (ms * 1e-3) / 148 / 1.965e9);
    ms = timeit([&] { k_dfma<<<G, T>>>((double *)d, N, 0.999, 0.5); });
    printf("DFMA            : %.2f T FMA/s  (%.2f warp-instr/clk/SM)\n", thr * 8 / ms / 1e9, thr * 8 / 32 / (ms * 1e-3) / 148 / 1.965e9);
    ms = timeit([&] { k_dfma_ffma2<<<G, T>>>((double *)d, N, 0.999, 0.5, 0x3f7fbe773f7fbe77ull); });
    printf("4 DFMA + 8 FFMA2: %.3f ms ; DFMA alone would be %.3f, FFMA2 alone %.3f (at the rates above)\n", ms, 0.f, 0.f);
    ms = timeit([&] { k_mix<0><<<G, T>>>((u64 *)d, N, 12345u, 0x3f7fbe773f7fbe77ull); });
    printf("mix K=0 (all PRMT+IADD): %.3f ms  -> %.2f T IQ samples/s\n", ms, thr * 2 / ms / 1e9);
    ms = timeit([&] { k_mix<1><<<G, T>>>((u64 *)d, N, 12345u, 0x3f7fbe773f7fbe77ull); });
    printf("mix K=1                : %.3f ms  -> %.2f T IQ samples/s\n", ms, thr * 2 / ms / 1e9);
    ms = timeit([&] { k_mix<2><<<G, T>>>((u64 *)d, N, 12345u, 0x3f7fbe773f7fbe77ull); });
    printf("mix K=2                : %.3f ms  -> %.2f T IQ samples/s\n", ms, thr * 2 / ms / 1e9);
    ms = timeit([&] { k_mix<4><<<G, T>>>((u64 *)d, N, 12345u, 0x3f7fbe773f7fbe77ull); });
    printf("mix K=4 (all I2F)      : %.3f ms  -> %.2f T IQ samples/s\n", ms, thr * 2 / ms / 1e9);
    return 0;
}
