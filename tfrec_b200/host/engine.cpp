// engine.cpp - the sample pump of engine.cpp:46-94 (baycom/tfrec) for -L replay: read the dump in whole
// 65536-byte blocks, hand them to the device path, stop at the first short read (the reference drops the
// partial tail block too, engine.cpp:73-76).
#include "engine.h"

#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <vector>

#include "../../include/tfr.h"

engine::engine(int, uint32_t, int, int filter, fsk_demod *_fsk, int _dbg, int _dumpmode, char *_dumpfile)
	: fsk(_fsk), filter_type(filter), dbg(_dbg), dumpmode(_dumpmode), dumpfile(_dumpfile), batch_blocks(256)
{
	if (filter_type) puts("Wide filter");
	if (dumpmode)
		printf("Dumpmode %i (%s), dumpfile %s\n", dumpmode, dumpmode == 1 ? "SAVE" : (dumpmode == -1 ? "LOAD" : "NONE"), dumpfile);
	if (dumpmode >= 0) {
		fprintf(stderr, "live capture needs librtlsdr and is outside the accelerated path: use -L <file>\n");
		exit(-1);
	}
}

engine::~engine(void) {}

void engine::run(int timeout)
{
	FILE *fd = fopen(dumpfile, "rb");
	if (!fd) {
		perror(dumpfile);
		exit(-1);
	}
	const time_t start = time(0);
	std::vector<uint8_t> buf((size_t)batch_blocks * TFR_BLOCK_BYTES);
	for (;;) {
		const size_t got = fread(buf.data(), TFR_BLOCK_BYTES, batch_blocks, fd);   // whole blocks only
		if (got > 0 && fsk->process_raw(buf.data(), got * TFR_BLOCK_BYTES, filter_type) < 0) exit(-1);
		if (got < (size_t)batch_blocks) {
			printf("done reading dump\n");
			break;   // the reference exit(0)s here; returning lets -m 1 summaries run
		}
		if (timeout && (time(0) - start > timeout)) break;
	}
	fclose(fd);
}
