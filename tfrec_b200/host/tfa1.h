// tfa1.h - registration objects for TFA_1 / KlimaLogg Pro (reference tfa1.h:11-33).  The demodulator
// (tfa1.cpp:143-190), framer (:120-134) and parser (:47-118) run on the device.
#ifndef TFRB200_HOST_TFA1_H
#define TFRB200_HOST_TFA1_H
#include "decoder.h"

class tfa1_decoder : public decoder {
      public:
	explicit tfa1_decoder(sensor_e _type) : decoder(_type) {}
};

class tfa1_demod : public demodulator {
      public:
	explicit tfa1_demod(decoder *_dec) : demodulator(_dec) {}
	double samples_per_bit(void) const { return 10.0; }   // BITPERIOD, tfa1.cpp:34
	bool device_native(void) const { return true; }
};
#endif
