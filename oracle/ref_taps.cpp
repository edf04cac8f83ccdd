// Intermediate-value taps for the UNMODIFIED reference, attached at link time with
//   -Wl,--wrap=_Z6fm_deviiii -Wl,--wrap=_Z11fm_dev_nrzsiiii -Wl,--wrap=_ZN4iir24stepEd
// (fm_dev / fm_dev_nrzs: /root/reference/dsp_stuff.cpp:269-292, iir2::step: dsp_stuff.cpp:47-55).
// The reference objects are not edited; the linker redirects their cross-TU calls through the
// __wrap_ functions below, which forward to the real implementation and append (args, result)
// to the binary file named by $TFR_TAP_FILE.  TEST INFRASTRUCTURE ONLY.
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

extern "C" {
int __real__Z6fm_deviiii(int, int, int, int);
int __real__Z11fm_dev_nrzsiiii(int, int, int, int);
double __real__ZN4iir24stepEd(void *, double);
}

namespace {
FILE *tap_fd = NULL;
int tap_mask = -1;
struct tap_rec { int32_t kind, a, b, c, d, res; double din, dout; uint64_t obj; };

FILE *tap_file(void)
{
	if (!tap_fd) {
		const char *p = getenv("TFR_TAP_FILE");
		const char *m = getenv("TFR_TAP_MASK"); // bit0 fm_dev, bit1 fm_dev_nrzs, bit2 iir2::step
		tap_mask = m ? atoi(m) : 7;
		tap_fd = fopen(p ? p : "/dev/null", "wb");
		if (!tap_fd) { perror("TFR_TAP_FILE"); exit(-1); }
	}
	return tap_fd;
}
void put(int kind, int a, int b, int c, int d, int res, double din, double dout, void *obj)
{
	FILE *f = tap_file();
	if (!(tap_mask & (1 << kind)))
		return;
	tap_rec r;
	memset(&r, 0, sizeof(r));
	r.kind = kind; r.a = a; r.b = b; r.c = c; r.d = d; r.res = res; r.din = din; r.dout = dout;
	r.obj = (uint64_t)(uintptr_t)obj;
	fwrite(&r, sizeof(r), 1, f);
}
}

extern "C" int __wrap__Z6fm_deviiii(int ar, int aj, int br, int bj)
{
	int r = __real__Z6fm_deviiii(ar, aj, br, bj);
	put(0, ar, aj, br, bj, r, 0, 0, NULL);
	return r;
}
extern "C" int __wrap__Z11fm_dev_nrzsiiii(int ar, int aj, int br, int bj)
{
	int r = __real__Z11fm_dev_nrzsiiii(ar, aj, br, bj);
	put(1, ar, aj, br, bj, r, 0, 0, NULL);
	return r;
}
extern "C" double __wrap__ZN4iir24stepEd(void *self, double din)
{
	double r = __real__ZN4iir24stepEd(self, din);
	put(2, 0, 0, 0, 0, 0, din, r, self);
	return r;
}
