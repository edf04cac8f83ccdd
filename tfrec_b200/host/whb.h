// whb.h - registration objects for TFA WeatherHub (reference whb.h:11-61).
#ifndef TFRB200_HOST_WHB_H
#define TFRB200_HOST_WHB_H
#include <stdio.h>
#include "decoder.h"

class whb_decoder : public decoder {
      public:
	explicit whb_decoder(sensor_e _type = TFA_WHB) : decoder(_type) {}
};

class whb_demod : public demodulator {
      public:
	whb_demod(decoder *_dec, double _spb) : demodulator(_dec), spb(_spb)
	{
		printf("WHB: Samples per bit: %.1f\n", spb);   // whb.cpp:612
	}
	double samples_per_bit(void) const { return spb; }
	bool device_native(void) const { return true; }

      private:
	double spb;
};
#endif
