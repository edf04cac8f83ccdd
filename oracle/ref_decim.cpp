// Decimator-only driver for the UNMODIFIED reference `downconvert` (dsp_stuff.cpp:232-264),
// linked against the reference's own dsp_stuff.o.  Mirrors the -L replay loop of engine.cpp:67-85:
// read <block> bytes, (u8-128)<<6 into int16, process_iq in place, keep the first `ld` values.
// usage: ref_decim <in.u8> <out.s16> <filter 0|1> [block_bytes=65536] [passes=2]
// TEST INFRASTRUCTURE ONLY.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "dsp_stuff.h"

int main(int argc, char **argv)
{
	if (argc < 4) {
		fprintf(stderr, "usage: %s in.u8 out.s16 filter [block_bytes] [passes]\n", argv[0]);
		return 2;
	}
	int filter = atoi(argv[3]);
	int block = argc > 4 ? atoi(argv[4]) : 65536;
	int passes = argc > 5 ? atoi(argv[5]) : 2;
	FILE *in = fopen(argv[1], "rb"), *out = fopen(argv[2], "wb");
	if (!in || !out) { perror("open"); return 1; }
	downconvert dc(passes);
	std::vector<unsigned char> raw(block);
	std::vector<int16_t> s(block);
	while (fread(raw.data(), block, 1, in) == 1) {
		for (int n = 0; n < block; n++)
			s[n] = (raw[n] - 128) << 6;
		int ld = dc.process_iq(s.data(), block, filter);
		fwrite(s.data(), sizeof(int16_t), ld, out);
	}
	fclose(in);
	fclose(out);
	return 0;
}
