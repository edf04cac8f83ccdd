// engine.h - sample pump (reference engine.h:19-44).  Same constructor and run(): -L replay (dumpmode -1), and a
// live mode fed with raw u8 IQ on stdin in place of librtlsdr (hardware I/O, outside this path), -S saving the
// consumed bytes (dumpmode 1).
#ifndef TFRB200_HOST_ENGINE_H
#define TFRB200_HOST_ENGINE_H
#include <stdint.h>
#include <string>
#include "fm_demod.h"

using std::string;

class engine {
      public:
	engine(int device, uint32_t freq, int gain, int filter, fsk_demod *fsk, int dbg, int dmpmode, char *dumpfile);
	~engine(void);
	void run(int timeout);
	void get_properties(string &, string &, string &) {}
	// blocks handed to the device per tfr_process call (the reference processes one 65536-byte block at a time)
	void set_batch_blocks(int n) { batch_blocks = n > 0 ? n : 1; }

      private:
	fsk_demod *fsk;
	int filter_type;
	int dbg;
	int dumpmode;
	char *dumpfile;
	int batch_blocks;
};
#endif
