// dsp_stuff.h - host mirror of the reference's decimator interface (dsp_stuff.h:46-56) over the C ABI.
//
// `downconvert(p)` + `process_iq(buf, len, filter)` with the reference's meaning: int16 I,Q interleaved, decimated in
// place by 2^p (p-1 times the eight-tap filter, then the twenty-tap one, narrow or wide), every stage's history
// carried from call to call, returns the new len.  The work runs on the GPU (tfr_dc_process_i16, csrc/decim.cu); a
// caller that still holds the raw rtl-sdr bytes should use process_u8, which goes through the fused kernel
// (csrc/decim_fused.cu) and skips the int16 detour of engine.cpp:77-78.
// One restriction the reference does not have: the number of IQ pairs of a call must be a multiple of 2^p (the
// reference silently drops an odd trailing sample of a block at every stage; engine.cpp feeds 32768 pairs per block).
// There is no CPU fallback: the constructor throws std::runtime_error without an sm_100 device.
#ifndef _TFR_B200_DSP_STUFF_H
#define _TFR_B200_DSP_STUFF_H

#include <stdint.h>

#include <stdexcept>
#include <string>

#include "../../include/tfr.h"

class downconvert {
      public:
	downconvert(int p, int device = 0) : passes(p), h(nullptr)
	{
		if (tfr_dc_create(device, p, &h) != TFR_OK) throw std::runtime_error(std::string("downconvert: ") + tfr_last_error());
	}
	~downconvert(void) { tfr_dc_destroy(h); }
	downconvert(const downconvert &) = delete;
	downconvert &operator=(const downconvert &) = delete;

	// dsp_stuff.cpp:243-264; a negative return is a TFR_E_* code (tfr_last_error() has the text)
	int process_iq(int16_t *buf, int len, int filter = 0) { return (int)tfr_dc_process_i16(h, buf, len, filter); }
	// the same stream fed as raw offset-binary bytes: out receives (nbytes/2 >> p) I,Q pairs; returns the int16 count
	long process_u8(const uint8_t *iq, size_t nbytes, int16_t *out, int filter = 0) { return tfr_dc_process(h, iq, nbytes, filter, out, TFR_MEM_HOST); }

      private:
	int passes;
	tfr_dc *h;
};

#endif
