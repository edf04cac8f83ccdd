// main.cpp - tfrec-compatible command line over the B200 decode path (flags of main.cpp:61-165, baycom/tfrec).
// Demodulators are registered in the reference's order with its samples-per-bit constants (main.cpp:171-218).
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <vector>

#include "engine.h"
#include "tfa1.h"
#include "tfa2.h"
#include "whb.h"

// -X: one message per line as hex bytes, '#' starts a comment (main.cpp:24-53)
static void replay_hex(vector<demodulator *> *demods, fsk_demod *fsk, int filter, const char *path)
{
	FILE *fd = fopen(path, "r");
	if (!fd) {
		perror("Can't open message file");
		exit(-1);
	}
	fsk->handle(filter);
	char line[1024];
	while (fgets(line, sizeof(line), fd)) {
		if (line[0] == '#') continue;
		unsigned char bytes[512];
		unsigned len = 0;
		for (char *tok = strtok(line, " \t\r\n"); tok && len < sizeof(bytes); tok = strtok(NULL, " \t\r\n"))
			bytes[len++] = (unsigned char)strtol(tok, NULL, 16);
		for (size_t n = 0; n < demods->size(); n++) {
			decoder *d = demods->at(n)->dec;
			d->store_bytes(bytes, (int)len);
			d->flush(0);
			puts("");
			d->flush_storage();
		}
	}
	fclose(fd);
}

static void usage(void)
{
	fprintf(stderr,
		"tfrec_b200 - B200 decode path for TFA IT+ (and compatible) sensors, tfrec compatible\n"
		"Options:\n"
		" -D          : Debug, print raw messages and some other stuff\n"
		" -e <exec>   : Executable to be called for every message (try echo)\n"
		" -t <thresh> : Set RF trigger threshold (default 0=auto)\n"
		" -m <mode>   : 0: exec handler for every message (default), 1: summary at program exit\n"
		" -w <timeout>: Run for <timeout> seconds (default: 0=forever)\n"
		" -W          : Wider filter, tolerate more frequency offset\n"
		" -T <types>  : HEX Bitmask of sensor types, default: 7 = TFA_1 | TFA_2 | TFA_3\n"
		"               %x: TFA_1, %x: TFA_2, %x: TFA_3, %x: TX22, %x: WeatherHub\n"
		" -q          : Quiet, do not print message to stdout\n"
		" -S <file>   : Live mode: save the raw IQ bytes that are decoded (replay them with -L)\n"
		" -L <file>   : Load IQ-file (rtl-sdr u8 dump, '-' = stdin) and decode it on the GPU\n"
		" -X <file>   : Load hexdump file and decode (test mode)\n"
		" without -L/-X: live mode, raw u8 IQ at 1.536 MS/s is read from stdin (rtl_sdr -s 1536000 - | ...)\n"
		" -d -f -g    : accepted for compatibility; tuning the stick is the feeder's job\n",
		1 << TFA_1, 1 << TFA_2, 1 << TFA_3, 1 << TX22, 1 << TFA_WHB);
}

int main(int argc, char **argv)
{
	int thresh = 0, debug = 0, timeout = 0, mode = 0, dumpmode = 0, types = 0x07, filter = 0;
	char *exec = NULL, *dumpfile = NULL, *hexfile = NULL;
	for (;;) {
		const int c = getopt(argc, argv, "d:Df:g:e:t:m:w:WqT:S:L:X:h");
		if (c == -1) break;
		switch (c) {
		case 'D': debug++; break;
		case 'e': exec = strdup(optarg); break;
		case 't': thresh = atoi(optarg); break;
		case 'm': mode = atoi(optarg); break;
		case 'w': timeout = atoi(optarg); break;
		case 'W': filter = 1; break;
		case 'q': debug = -1; break;
		case 'T': types = (int)strtol(optarg, NULL, 16); break;
		case 'S': dumpfile = strdup(optarg); dumpmode = 1; break;
		case 'L': dumpfile = strdup(optarg); dumpmode = -1; break;
		case 'X': hexfile = strdup(optarg); break;
		case 'd': case 'f': case 'g': break;
		default: usage(); return 0;
		}
	}

	vector<demodulator *> demods;
	if (types & (1 << TFA_1)) {
		printf("Registering demod for TFA_1 KlimaLoggPro\n");
		decoder *d = new tfa1_decoder(TFA_1);
		d->set_params(exec, mode, debug);
		demods.push_back(new tfa1_demod(d));
	}
	if (types & (1 << TFA_2)) {
		printf("Registering demod for TFA_2 sensors, 17240 bit/s\n");
		decoder *d = new tfa2_decoder(TFA_2);
		d->set_params(exec, mode, debug);
		demods.push_back(new tfa2_demod(d, (1536000 / 4.0) / 17240));
	}
	if (types & (1 << TFA_3)) {
		printf("Registering demod for TFA_3 sensors, 9600 bit/s\n");
		decoder *d = new tfa2_decoder(TFA_3);
		d->set_params(exec, mode, debug);
		demods.push_back(new tfa2_demod(d, (1536000 / 4.0) / 9600));
	}
	if (types & (1 << TX22)) {
		printf("Registering demod for TX22, 8842 bit/s\n");
		decoder *d = new tfa2_decoder(TX22);
		d->set_params(exec, mode, debug);
		demods.push_back(new tfa2_demod(d, (1536000 / 4.0) / 8842, 0.5));
	}
	if (types & (1 << TFA_WHB)) {
		printf("Registering demod for TFA_WHB sensors, 6000 bit/s\n");
		decoder *d = new whb_decoder(TFA_WHB);
		d->set_params(exec, mode, debug);
		demods.push_back(new whb_demod(d, (1536000 / 4.0) / 6000));
	}

	fsk_demod fsk(&demods, thresh, debug);
	if (hexfile) {
		replay_hex(&demods, &fsk, filter, hexfile);
		return 0;
	}
	engine e(0, 868250, -1, filter, &fsk, debug, dumpmode, dumpfile);
	e.run(timeout);
	if (mode)
		for (size_t n = 0; n < demods.size(); n++) demods.at(n)->dec->flush_storage();
	return 0;
}
