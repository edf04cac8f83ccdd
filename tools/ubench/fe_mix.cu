// issue-slot cost of candidate front-end instructions next to a saturated FFMA2.RM stream (round 2).
// Every variant runs 18 FFMA2.RM (six independent chains) per loop iteration plus eight copies of ONE other
// instruction (four independent registers); cycles per iteration and sub-partition tell what the extra
// instruction costs: 36 + 8*c with c = 1 for a single issue slot, less if it hides under the FFMA2s.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define G(x, p) asm volatile("fma.rm.f32x2 %0, %1, %2, %0;" : "+l"(x) : "l"(p), "l"(a));
#define FMAS G(x0, p0) G(x1, p1) G(x2, p0) G(x3, p1) G(x4, p0) G(x5, p1) G(x0, p1) G(x1, p0) G(x2, p1) G(x3, p0) G(x4, p1) G(x5, p0) \
	G(x0, p0) G(x1, p1) G(x2, p0) G(x3, p1) G(x4, p0) G(x5, p1)
enum { NONE, IADD, PRMT, IDP4A, LOP3, IMAD, LDS, SHFL, FADD, FADD2, IABS, I2FP, FFMA36, VIMNMX, ISETP_SEL, DADD, DFMA1, STS };
template <int V> __global__ void __launch_bounds__(256) k_mix(u64 *out, int n, uint32_t w, u64 a)
{
	__shared__ uint32_t sm[1024];
	u64 x0 = out[0], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5;
	u64 p0 = out[1], p1 = p0 + 7;
	uint32_t r0 = w + threadIdx.x, r1 = r0 * 3, r2 = r0 * 5, r3 = r0 * 7;
	double d0 = (double)r0, d1 = (double)r1;
	sm[threadIdx.x] = r0; sm[threadIdx.x + 256] = r1; sm[threadIdx.x + 512] = r2; sm[threadIdx.x + 768] = r3;
	__syncthreads();
	const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm) + 16 * (threadIdx.x & 63);
	for (int i = 0; i < n; i++) {
		if (V == FFMA36) {
			float f0 = __uint_as_float((uint32_t)x0), f1 = __uint_as_float((uint32_t)x1), f2 = __uint_as_float((uint32_t)x2), f3 = __uint_as_float((uint32_t)x3);
			float c = __uint_as_float((uint32_t)a), q = __uint_as_float((uint32_t)p0);
#pragma unroll
			for (int j = 0; j < 9; j++) {
				asm volatile("fma.rm.f32 %0, %1, %2, %0;" : "+f"(f0) : "f"(q), "f"(c));
				asm volatile("fma.rm.f32 %0, %1, %2, %0;" : "+f"(f1) : "f"(q), "f"(c));
				asm volatile("fma.rm.f32 %0, %1, %2, %0;" : "+f"(f2) : "f"(q), "f"(c));
				asm volatile("fma.rm.f32 %0, %1, %2, %0;" : "+f"(f3) : "f"(q), "f"(c));
			}
			x0 = __float_as_uint(f0); x1 = __float_as_uint(f1); x2 = __float_as_uint(f2); x3 = __float_as_uint(f3);
		} else {
			FMAS
		}
#pragma unroll
		for (int j = 0; j < 2; j++) {
			if (V == IADD || V == FFMA36) {
				asm volatile("add.u32 %0, %0, %1;" : "+r"(r0) : "r"(w)); asm volatile("add.u32 %0, %0, %1;" : "+r"(r1) : "r"(w));
				asm volatile("add.u32 %0, %0, %1;" : "+r"(r2) : "r"(w)); asm volatile("add.u32 %0, %0, %1;" : "+r"(r3) : "r"(w));
			} else if (V == PRMT) {
				asm volatile("prmt.b32 %0, %0, %1, 0x4404;" : "+r"(r0) : "r"(w)); asm volatile("prmt.b32 %0, %0, %1, 0x4414;" : "+r"(r1) : "r"(w));
				asm volatile("prmt.b32 %0, %0, %1, 0x4424;" : "+r"(r2) : "r"(w)); asm volatile("prmt.b32 %0, %0, %1, 0x4434;" : "+r"(r3) : "r"(w));
			} else if (V == IDP4A) {
				asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r0) : "r"(0x00000080u), "r"(w)); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r1) : "r"(0x00008000u), "r"(w));
				asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r2) : "r"(0x00800000u), "r"(w)); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r3) : "r"(0x80000000u), "r"(w));
			} else if (V == LOP3) {
				asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r0) : "r"(w), "r"(r1)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r1) : "r"(w), "r"(r2));
				asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r2) : "r"(w), "r"(r3)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r3) : "r"(w), "r"(r0));
			} else if (V == IMAD) {
				asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r0) : "r"(w), "r"(r1)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r1) : "r"(w), "r"(r2));
				asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r2) : "r"(w), "r"(r3)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r3) : "r"(w), "r"(r0));
			} else if (V == LDS) {   // 2 x LDS.128 per j (4 per iteration): 16 B per lane, conflict free
				uint32_t a0, a1, a2, a3;
				asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(sa + (r0 & 0x800)));
				r0 ^= a0 ^ a1; r1 ^= a2 ^ a3;
				asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(sa + (r2 & 0x800)));
				r2 ^= a0 ^ a1; r3 ^= a2 ^ a3;
			} else if (V == STS) {
				asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(sa), "r"(r0), "r"(r1)); asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(sa + 8), "r"(r2), "r"(r3));
				asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(sa + 2048), "r"(r0), "r"(r1)); asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(sa + 2056), "r"(r2), "r"(r3));
			} else if (V == SHFL) {
				r0 = __shfl_up_sync(0xffffffffu, r0, 1); r1 = __shfl_up_sync(0xffffffffu, r1, 1);
				r2 = __shfl_up_sync(0xffffffffu, r2, 1); r3 = __shfl_up_sync(0xffffffffu, r3, 1);
			} else if (V == FADD) {
				float f; 
				asm volatile("add.rm.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(r0)), "f"(1.0f)); r0 = __float_as_uint(f);
				asm volatile("add.rm.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(r1)), "f"(1.0f)); r1 = __float_as_uint(f);
				asm volatile("add.rm.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(r2)), "f"(1.0f)); r2 = __float_as_uint(f);
				asm volatile("add.rm.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(r3)), "f"(1.0f)); r3 = __float_as_uint(f);
			} else if (V == FADD2) {   // 4 x add.f32x2 = 8 lane adds
				asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(a)); asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(a));
				u64 t0 = ((u64)r1 << 32) | r0, t1 = ((u64)r3 << 32) | r2;
				asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(t0) : "l"(a)); asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(t1) : "l"(a));
				r0 = (uint32_t)t0; r1 = (uint32_t)(t0 >> 32); r2 = (uint32_t)t1; r3 = (uint32_t)(t1 >> 32);
			} else if (V == IABS) {
				asm volatile("abs.s32 %0, %0;" : "+r"(r0)); asm volatile("abs.s32 %0, %0;" : "+r"(r1));
				asm volatile("abs.s32 %0, %0;" : "+r"(r2)); asm volatile("abs.s32 %0, %0;" : "+r"(r3));
			} else if (V == I2FP) {
				float f;
				asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(r0)); r0 = __float_as_uint(f);
				asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(r1)); r1 = __float_as_uint(f);
				asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(r2)); r2 = __float_as_uint(f);
				asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(r3)); r3 = __float_as_uint(f);
			} else if (V == VIMNMX) {
				asm volatile("max.s32 %0, %0, %1;" : "+r"(r0) : "r"(w)); asm volatile("max.s32 %0, %0, %1;" : "+r"(r1) : "r"(w));
				asm volatile("max.s32 %0, %0, %1;" : "+r"(r2) : "r"(w)); asm volatile("max.s32 %0, %0, %1;" : "+r"(r3) : "r"(w));
			} else if (V == ISETP_SEL) {   // 2 x (setp + predicated or)
				asm volatile("{.reg .pred p; setp.gt.s32 p, %1, %2; @p or.b32 %0, %0, 0x10;}" : "+r"(r0) : "r"(r1), "r"(w));
				asm volatile("{.reg .pred p; setp.gt.s32 p, %1, %2; @p or.b32 %0, %0, 0x20;}" : "+r"(r2) : "r"(r3), "r"(w));
			} else if (V == DADD) {   // 2 per j
				asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d0) : "d"(1.5)); asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d1) : "d"(1.5));
			} else if (V == DFMA1) {
				asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d0) : "d"(0.999), "d"(1.5)); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d1) : "d"(0.999), "d"(1.5));
			}
		}
	}
	out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + x4 + x5 + r0 + r1 + r2 + r3 + p0 + p1 + (u64)d0 + (u64)d1 + sm[threadIdx.x ^ 1];
}
template <class F> static float timeit(F f)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	f(); cudaDeviceSynchronize();
	cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int V> static void run(const char *name, int extra, void *d, double base)
{
	const int G = 148 * 4, T = 256, N = 4096;   // 4 CTAs x 8 warps per SM = 8 warps per sub-partition
	float ms = timeit([&] { k_mix<V><<<G, T>>>((u64 *)d, N, 12345u, 0x3f7fbe773f7fbe77ull); });
	int dev; cudaGetDevice(&dev); int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
	const double cyc = ms * 1e-3 * 1.965e9 / (8.0 * N);   // per iteration and warp slot of a sub-partition (8 warps share it)
	printf("%-10s %7.3f ms  %6.1f cycles/iteration  -> %+.2f cycles per extra instruction (%d of them)\n", name, ms, cyc,
	       extra ? (cyc - base) / extra : 0.0, extra);
}
int main()
{
	void *d; cudaMalloc(&d, 1 << 26); cudaMemset(d, 0, 1 << 26);
	const int G = 148 * 4, T = 256, N = 4096;
	float ms0 = timeit([&] { k_mix<NONE><<<G, T>>>((u64 *)d, N, 12345u, 0x3f7fbe773f7fbe77ull); });
	const double base = ms0 * 1e-3 * 1.965e9 / (8.0 * N);
	printf("18 FFMA2.RM alone: %.3f ms, %.1f cycles/iteration (36 = two issue cycles each)\n", ms0, base);
	run<IADD>("IADD", 8, d, base); run<PRMT>("PRMT", 8, d, base); run<IDP4A>("IDP.4A", 8, d, base); run<LOP3>("LOP3", 8, d, base);
	run<IMAD>("IMAD", 8, d, base); run<LDS>("LDS.128", 4, d, base); run<STS>("STS.64", 8, d, base); run<SHFL>("SHFL", 8, d, base);
	run<FADD>("FADD.RM", 8, d, base); run<FADD2>("FADD2.RM", 8, d, base); run<IABS>("IABS", 8, d, base); run<I2FP>("I2FP", 8, d, base);
	run<VIMNMX>("VIMNMX", 8, d, base); run<ISETP_SEL>("SETP+@OR", 8, d, base); run<DADD>("DADD", 4, d, base); run<DFMA1>("DFMA", 4, d, base);
	run<FFMA36>("36 FFMA+8 IADD", 8, d, base);
	return 0;
}
