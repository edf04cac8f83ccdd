// round 2, second probe: does a stream of SCALAR fma.rm.f32 leave issue slots for ALU work that the packed
// fma.rm.f32x2 stream does not?  Each variant: NF scalar FFMA (or NF/2 FFMA2) in 12 independent chains plus NA
// instructions of one ALU kind (8 independent registers), per loop iteration; cycles per iteration and sub-partition.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
enum { K_NONE, K_IADD, K_PRMT, K_CVT, K_IDP, K_LOP, K_OUT };
#define FS(f) asm volatile("fma.rm.f32 %0, %1, %2, %0;" : "+f"(f) : "f"(q), "f"(c));
#define FP(x) asm volatile("fma.rm.f32x2 %0, %1, %2, %0;" : "+l"(x) : "l"(q2), "l"(c2));
template <int PACKED, int NF, int KIND, int NA> __global__ void __launch_bounds__(256) k(u64 *out, int n, uint32_t w, float c, float q)
{
	float f[12];
	u64 x[6];
	uint32_t r[8];
#pragma unroll
	for (int i = 0; i < 12; i++) f[i] = (float)(out[0] + i);
#pragma unroll
	for (int i = 0; i < 6; i++) x[i] = out[1] + i;
#pragma unroll
	for (int i = 0; i < 8; i++) r[i] = (w + threadIdx.x) * (2 * i + 3);
	u64 q2 = ((u64)__float_as_uint(q) << 32) | __float_as_uint(q), c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
	for (int it = 0; it < n; it++) {
		// interleave: after every FFMA group a slice of the ALU work, as a scheduler would
#pragma unroll
		for (int g = 0; g < NF / 12; g++) {
			if (PACKED) {
#pragma unroll
				for (int i = 0; i < 6; i++) FP(x[i])
			} else {
#pragma unroll
				for (int i = 0; i < 12; i++) FS(f[i])
			}
#pragma unroll
			for (int a = g * NA / (NF / 12); a < (g + 1) * NA / (NF / 12); a++) {
				uint32_t &t = r[a & 7];
				if (KIND == K_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(t) : "r"(w));
				if (KIND == K_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x4414;" : "+r"(t) : "r"(w));
				if (KIND == K_LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(t) : "r"(w), "r"(r[(a + 1) & 7]));
				if (KIND == K_IDP) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(t) : "r"(0x00008000u), "r"(w));
				if (KIND == K_CVT) {   // the real conversion pair: PRMT then IADD (counts as two of NA)
					if (a & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(t) : "r"(w));
					else asm volatile("prmt.b32 %0, %0, %1, 0x4414;" : "+r"(t) : "r"(w));
				}
				if (KIND == K_OUT) {   // the per-output mix: IADD3, IABS, IADD, ISETP+@LOP, PRMT (cycled)
					switch (a % 6) {
					case 0: asm volatile("add.u32 %0, %0, %1;" : "+r"(t) : "r"(w)); break;
					case 1: asm volatile("abs.s32 %0, %0;" : "+r"(t)); break;
					case 2: asm volatile("add.u32 %0, %0, %1;" : "+r"(t) : "r"(r[(a + 3) & 7])); break;
					case 3: asm volatile("{.reg .pred p; setp.gt.s32 p, %1, %2; @p or.b32 %0, %0, 0x10;}" : "+r"(t) : "r"(r[(a + 1) & 7]), "r"(w)); break;
					case 4: asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(t) : "r"(r[(a + 2) & 7])); break;
					case 5: asm volatile("abs.s32 %0, %0;" : "+r"(t)); break;
					}
				}
			}
		}
	}
	u64 s = 0;
#pragma unroll
	for (int i = 0; i < 12; i++) s += __float_as_uint(f[i]);
#pragma unroll
	for (int i = 0; i < 6; i++) s += x[i];
#pragma unroll
	for (int i = 0; i < 8; i++) s += r[i];
	out[threadIdx.x + blockIdx.x * blockDim.x] = s;
}
static int g_ctas = 4;
template <int PACKED, int NF, int KIND, int NA> static void run(const char *name, void *d)
{
	const int G = 148 * g_ctas, T = 256, N = 2048;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<PACKED, NF, KIND, NA><<<G, T>>>((u64 *)d, N, 12345u, 0.999f, 1.5f); cudaDeviceSynchronize();
	cudaEventRecord(e0); k<PACKED, NF, KIND, NA><<<G, T>>>((u64 *)d, N, 12345u, 0.999f, 1.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	const double cyc = ms * 1e-3 * 1.965e9 / (2.0 * g_ctas * N);   // per iteration and warp slot of a sub-partition
	printf("%-34s %s NF=%2d NA=%2d : %7.3f ms  %6.1f cycles/iteration (FMA issue floor %d)\n", name, PACKED ? "FFMA2" : "FFMA ", NF, NA, ms, cyc, NF);
}
int main(int argc, char **argv)
{
	void *d; cudaMalloc(&d, 1 << 26); cudaMemset(d, 0, 1 << 26);
	for (g_ctas = 1; g_ctas <= 4; g_ctas *= 2) {
		printf("---- %d warps per sub-partition\n", 2 * g_ctas);
		run<0, 36, K_NONE, 0>("scalar alone", d);        run<1, 36, K_NONE, 0>("packed alone", d);
		run<0, 36, K_IADD, 8>("+ IADD", d);              run<1, 36, K_IADD, 8>("+ IADD", d);
		run<0, 36, K_IADD, 16>("+ IADD", d);             run<1, 36, K_IADD, 16>("+ IADD", d);
		run<0, 36, K_IADD, 24>("+ IADD", d);             run<1, 36, K_IADD, 24>("+ IADD", d);
		run<0, 36, K_PRMT, 8>("+ PRMT", d);              run<1, 36, K_PRMT, 8>("+ PRMT", d);
		run<0, 36, K_PRMT, 16>("+ PRMT", d);             run<1, 36, K_PRMT, 16>("+ PRMT", d);
		run<0, 36, K_CVT, 8>("+ PRMT/IADD pairs (2 samples)", d);  run<1, 36, K_CVT, 8>("+ PRMT/IADD pairs (2 samples)", d);
		run<0, 36, K_CVT, 16>("+ PRMT/IADD pairs x2", d);          run<1, 36, K_CVT, 16>("+ PRMT/IADD pairs x2", d);
		run<0, 36, K_IDP, 4>("+ IDP.4A (2 samples)", d); run<1, 36, K_IDP, 4>("+ IDP.4A (2 samples)", d);
		run<0, 36, K_IDP, 8>("+ IDP.4A", d);             run<1, 36, K_IDP, 8>("+ IDP.4A", d);
		run<0, 36, K_LOP, 12>("+ LOP3", d);              run<1, 36, K_LOP, 12>("+ LOP3", d);
		run<0, 36, K_OUT, 6>("+ output mix", d);         run<1, 36, K_OUT, 6>("+ output mix", d);
		run<0, 36, K_OUT, 12>("+ output mix x2", d);     run<1, 36, K_OUT, 12>("+ output mix x2", d);
	}
	return 0;
}
