// decoder.h - host-side mirror of the reference's demodulator/decoder plugin surface (decoder.h:11-73 of
// baycom/tfrec), kept source compatible: the same class names, virtuals, constructor arguments and the same
// sensordata_t, so that a main.cpp-style registration
//
//     decoder *d = new tfa1_decoder(TFA_1);  d->set_params(exec, mode, debug);
//     demods.push_back(new tfa1_demod(d));
//
// compiles unchanged.  What differs is where the work happens: the per-sample demod()/store_bit() state
// machines run on the GPU behind the C ABI (include/tfr.h); the objects here carry the registration (type,
// samples per bit) and receive the results - store_data() with its first-seen/duplicate rule and the -e exec
// contract (decoder.cpp:46-96) are implemented here, on the host, as in the reference.
#ifndef TFRB200_HOST_DECODER_H
#define TFRB200_HOST_DECODER_H

#include <stdint.h>
#include <sys/time.h>
#include <time.h>

#include <map>
#include <string>

#include "../../include/tfr.h"

// sensor types; the -T mask is the OR of 1<<value (main.cpp:146-148)
enum sensor_e {
	TFA_1 = 0,       // KlimaLogg Pro 30.3180/81/99, NRZS 38400
	TFA_2,           // 30.3143/44/46, NRZ 17240
	TFA_3,           // 30.3155, NRZ 9600
	TX22,            // LaCrosse TX22, NRZ 8842
	TFA_WHP,         // (never registered by the reference, main.cpp:204-212)
	TFA_WHB,         // WeatherHub, PSK/NRZS/G3RUH 6000
	FIREANGEL = 0x20
};

typedef struct {
	sensor_e type;
	uint64_t id;
	double temp;
	double humidity;
	int alarm;
	int flags;
	int sequence;
	time_t ts;
	int rssi;
} sensordata_t;

class decoder {
      public:
	explicit decoder(sensor_e _type);
	virtual ~decoder() {}
	void set_params(char *_handler, int _mode, int _dbg);
	// the two per-bit entry points exist for source compatibility; on this path bits never reach the host
	virtual void store_bit(int bit);
	// -X seam (main.cpp:45-50): parses the bytes given to store_bytes() with the DEVICE parser
	virtual void flush(int rssi, int offset = 0);
	virtual void store_data(sensordata_t &d);
	virtual void execute_handler(sensordata_t &d);
	virtual void flush_storage(void);
	// TFREC_EXEC=async (opt-in): pushes the commands queued so far to the handler shell; a no-op with the default
	// per-telegram system()
	static void flush_exec(void);
	virtual int has_sync(void) { return synced; }
	int count(void) { return (int)data.size(); }
	sensor_e get_type(void) { return type; }
	virtual void store_bytes(uint8_t *d, int len);

	// glue used by engine: where flush() finds a device parser, and result delivery
	void attach(tfr_handle *h) { handle = h; }
	void deliver_frame(const tfr_frame &f, sensordata_t *recs, int n_recs);
	int bad_count(void) const { return bad; }

      protected:
	int dbg;
	int bad;
	int synced;
	sensor_e type;
	uint8_t rdata[256];
	int byte_cnt;
	int snum;

      private:
	char *handler;
	int mode;
	std::map<uint64_t, sensordata_t> data;
	tfr_handle *handle;
};

class demodulator {
      public:
	explicit demodulator(decoder *_dec);
	virtual ~demodulator() {}
	virtual void start(int len);
	virtual void reset(void) {}
	// never called on this path: the GPU runs the built-in demodulators; a user-defined CPU demodulator cannot
	// be plugged into the device pipeline and engine refuses it (there is no CPU fallback)
	virtual int demod(int thresh, int pwr, int index, int16_t *iq);
	virtual double samples_per_bit(void) const { return 0; }
	virtual bool device_native(void) const { return false; }

	decoder *dec;

      protected:
	int last_bit_idx;
};

// format helpers shared by decoder.cpp and main.cpp
std::string tfr_format_line(const tfr_frame &f, const sensordata_t *recs, int n_recs, int dbg);

#endif
