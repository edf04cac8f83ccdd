// tfa2.h - registration objects for TFA_2 / TFA_3 / TX22 (reference tfa2.h:11-44).  spb is 384000/baud as in
// main.cpp:186,194,202; the device uses the reference's as-built biquad coefficients for exactly those rates.
#ifndef TFRB200_HOST_TFA2_H
#define TFRB200_HOST_TFA2_H
#include <stdio.h>
#include "decoder.h"

class tfa2_decoder : public decoder {
      public:
	explicit tfa2_decoder(sensor_e _type = TFA_2) : decoder(_type) {}
};

class tfa2_demod : public demodulator {
      public:
	tfa2_demod(decoder *_dec, double _spb, double _iir_fac = 0.5) : demodulator(_dec), spb(_spb), iir_fac(_iir_fac)
	{
		printf("type 0x%x: Samples per bit: %.1f\n", _dec->get_type(), spb);   // tfa2.cpp:322
	}
	double samples_per_bit(void) const { return spb; }
	bool device_native(void) const { return iir_fac == 0.5; }

      private:
	double spb, iir_fac;
};
#endif
