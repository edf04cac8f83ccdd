// fir_taps.h - the reference's decimator tap tables as compile-time constants, shared by the fused front-end
// (frontend.cu) and the stand-alone downconvert stages (decim.cu).
#pragma once

namespace tfr {

// ------------------------------------------------------------------------------------------------
// tap tables (dsp_stuff.cpp:61-88 narrow, :91-117 wide, :119-130 first stage)
// ------------------------------------------------------------------------------------------------
#define TFR_T2 { 2443, 6339, 11036, 14254, 14254, 11036, 6339, 2443 }
#define TFR_T1N { -1087, -1082, -1065, -451, 912, 2997, 5556, 8157, 10285, 11484, 11484, 10285, 8157, 5556, 2997, 912, -451, -1065, -1082, -1087 }
#define TFR_T1W { 546, 451, -317, -1844, -3198, -2817, 494, 6469, 13074, 17421, 17421, 13074, 6469, 494, -2817, -3198, -1844, -317, 451, 546 }

__host__ __device__ constexpr int t2_tap(int n) { constexpr int t[8] = TFR_T2; return t[n]; }
__host__ __device__ constexpr int t1_tap(bool wide, int n)
{
	constexpr int tn[20] = TFR_T1N;
	constexpr int tw[20] = TFR_T1W;
	return wide ? tw[n] : tn[n];
}
__host__ __device__ constexpr int t2_sum() { int s = 0; for (int n = 0; n < 8; n++) s += t2_tap(n); return s; }
__host__ __device__ constexpr int t1_sum(bool wide, int lo, int hi) { int s = 0; for (int n = lo; n < hi; n++) s += t1_tap(wide, n); return s; }

}  // namespace tfr
