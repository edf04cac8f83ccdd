// engine.cpp - the sample pump of engine.cpp:46-94 (baycom/tfrec): read raw rtl-sdr u8 IQ in whole 65536-byte
// blocks, hand them to the device path, stop at the first short read (the reference drops the partial tail
// block too, engine.cpp:73-76).
//   -L <file>   replay a dump (dumpmode -1, engine.cpp:50-58, 67-81); "-" reads the dump from stdin
//   otherwise   LIVE: the reference pulls from librtlsdr (sdr.cpp:228-271), which is hardware I/O outside this
//               path; here the feed is raw u8 IQ at 1.536 MS/s on stdin, e.g.
//                   rtl_sdr -f 868250000 -s 1536000 -g 0 - | tfrec_b200_cli -e handler
//               (or `nc host 1234` behind an rtl_tcp header stripper).  -S <file> saves exactly the bytes that
//               are consumed, like the reference's dump writer (sdr.cpp:38-44, 233-234), so a later -L replay
//               of that file reproduces the live run.
#include "engine.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
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
		// live: latency matters more than launch efficiency - 8 blocks are 0.17 s of signal
		batch_blocks = 8;
		fprintf(stderr, "live mode: reading raw u8 IQ (1.536 MS/s) from stdin\n");
	}
}

engine::~engine(void) {}

void engine::run(int timeout)
{
	FILE *fd = stdin, *save_fd = NULL;
	if (dumpmode < 0 && strcmp(dumpfile, "-") != 0) {
		fd = fopen(dumpfile, "rb");
		if (!fd) {
			perror(dumpfile);
			exit(-1);
		}
	}
	if (dumpmode > 0 && dumpfile) {   // sdr::sdr, sdr.cpp:38-44
		save_fd = fopen(dumpfile, "wb");
		if (!save_fd) {
			perror(dumpfile);
			exit(-1);
		}
	}
	const time_t start = time(0);
	std::vector<uint8_t> buf((size_t)batch_blocks * TFR_BLOCK_BYTES);
	for (;;) {
		const size_t got = fread(buf.data(), TFR_BLOCK_BYTES, batch_blocks, fd);   // whole blocks only
		if (got > 0 && save_fd && fwrite(buf.data(), TFR_BLOCK_BYTES, got, save_fd) != got) {
			perror(dumpfile);
			exit(-1);
		}
		if (got > 0 && fsk->process_raw(buf.data(), got * TFR_BLOCK_BYTES, filter_type) < 0) exit(-1);
		if (got < (size_t)batch_blocks) {
			if (dumpmode < 0) printf("done reading dump\n");
			break;   // the reference exit(0)s here; returning lets -m 1 summaries run
		}
		if (timeout && (time(0) - start > timeout)) break;
	}
	if (save_fd) fclose(save_fd);
	if (fd != stdin) fclose(fd);
}
