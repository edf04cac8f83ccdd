// fm_demod.h - fsk_demod (reference fm_demod.h:18-31): owns the registered demodulators and the trigger
// threshold.  On this path it is the place where the registration is turned into a device handle; the
// per-block work of fsk_demod::process (fm_demod.cpp:34-74) happens inside tfr_process().
#ifndef TFRB200_HOST_FM_DEMOD_H
#define TFRB200_HOST_FM_DEMOD_H
#include <vector>
#include "decoder.h"

using std::vector;

class fsk_demod {
      public:
	fsk_demod(vector<demodulator *> *_demods, int _thresh, int _dbg);
	~fsk_demod();
	// reference signature (fm_demod.h:21): int16 I,Q already decimated to 384 kS/s, len = number of int16, a multiple
	// of 16384 (one reference block) - for callers that keep their own decimator
	void process(int16_t *data_iq, int len);
	// the replacement for `dc.process_iq(data,len,filter); fsk->process(data,ld);` (engine.cpp:85-86):
	// nbytes of raw rtl-sdr u8 IQ, a multiple of 65536; results are delivered to the decoders' store_data()
	int process_raw(const uint8_t *iq, size_t nbytes, int filter_type);
	int types_mask(void) const { return types; }
	int threshold(void);
	tfr_handle *handle(int filter_type);
	int deliver(tfr_handle *hh);

      private:
	vector<demodulator *> *demods;
	int thresh, dbg, types;
	tfr_handle *h;
	int h_filter;
	size_t frames_seen;
};
#endif
