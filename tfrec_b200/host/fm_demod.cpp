// fm_demod.cpp - fsk_demod over the C ABI: turns the registered demodulators into a device handle and pumps raw
// blocks through it; results come back ordered by (time, registration order) like the reference's output.
#include "fm_demod.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/tfr.h"

fsk_demod::fsk_demod(vector<demodulator *> *_demods, int _thresh, int _dbg)
	: demods(_demods), thresh(_thresh), dbg(_dbg), types(0), h(NULL), h_filter(-1), frames_seen(0)
{
	for (size_t n = 0; n < demods->size(); n++) {
		demodulator *d = demods->at(n);
		if (!d->device_native()) {
			fprintf(stderr, "fsk_demod: demodulator %zu is not one of the built-in device demodulators; "
					"this path has no CPU fallback\n", n);
			exit(-1);
		}
		types |= 1 << d->dec->get_type();
	}
}

fsk_demod::~fsk_demod()
{
	if (h) tfr_destroy(h);
}

tfr_handle *fsk_demod::handle(int filter_type)
{
	if (h && h_filter == filter_type) return h;
	if (h) tfr_destroy(h);
	tfr_config c;
	memset(&c, 0, sizeof(c));
	c.struct_size = sizeof(c);
	c.device = getenv("TFR_DEVICE") ? atoi(getenv("TFR_DEVICE")) : 0;
	c.types = types;
	c.filter = filter_type;
	c.thresh = thresh;
	c.n_streams = 1;
	if (tfr_create(&c, &h) != TFR_OK) {
		fprintf(stderr, "tfr_create: %s\n", tfr_last_error());
		exit(-1);
	}
	h_filter = filter_type;
	for (size_t n = 0; n < demods->size(); n++) demods->at(n)->dec->attach(h);
	return h;
}

// fm_demod.cpp:34-74 with the reference's own arguments: int16 I,Q at 384 kS/s, len = number of int16 (16384 per
// reference block).  For callers that keep their own decimator (dsp_stuff.h `downconvert`); the fused path from the
// raw bytes is process_raw().
void fsk_demod::process(int16_t *data_iq, int len)
{
	tfr_handle *hh = handle(h_filter < 0 ? 0 : h_filter);
	if (tfr_submit_decimated(hh, 0, data_iq, (size_t)len, TFR_MEM_HOST) || tfr_process(hh)) {
		fprintf(stderr, "tfr: %s\n", tfr_last_error());
		return;
	}
	deliver(hh);
}

int fsk_demod::threshold(void)
{
	int32_t t = thresh;
	if (h) tfr_get_thresh(h, 0, &t);
	return t;
}

int fsk_demod::process_raw(const uint8_t *iq, size_t nbytes, int filter_type)
{
	tfr_handle *hh = handle(filter_type);
	if (tfr_submit(hh, 0, iq, nbytes, TFR_MEM_HOST) || tfr_process(hh)) {
		fprintf(stderr, "tfr: %s\n", tfr_last_error());
		return -1;
	}
	return deliver(hh);
}

// the call's frames and records, in the reference's output order, to the registered decoders
int fsk_demod::deliver(tfr_handle *hh)
{
	long nf = tfr_poll_frames(hh, NULL, 0), nr = tfr_poll_records(hh, NULL, 0);
	if (nf < 0 || nr < 0) {
		fprintf(stderr, "tfr: %s\n", tfr_last_error());
		return -1;
	}
	vector<tfr_frame> frames((size_t)nf + 1);
	vector<tfr_record> recs((size_t)nr + 1);
	nf = tfr_poll_frames(hh, frames.data(), (size_t)nf);
	nr = tfr_poll_records(hh, recs.data(), (size_t)nr);
	if (dbg >= 3) {   // fm_demod.cpp:60-72
		const long nb = tfr_read_block_trace(hh, 0, NULL, 0);
		vector<tfr_block_trace> tr((size_t)(nb > 0 ? nb : 0) + 1);
		const long got = nb > 0 ? tfr_read_block_trace(hh, 0, tr.data(), (size_t)nb) : 0;
		for (long k = 0; k < got; k++) {
			printf("Trigger ratio %i/%i, avg %i \n", tr[k].triggered, 8192, tr[k].triggered_avg);
			const int next = (k + 1 < got) ? tr[k + 1].thresh : threshold();
			if (next > tr[k].thresh) printf("Increased trigger level to %i\n", next);
			if (next < tr[k].thresh) printf("Decreased trigger level to %i\n", next);
		}
	}
	for (long k = 0; k < nf; k++) {
		const tfr_frame &f = frames[k];
		decoder *dec = NULL;
		for (size_t n = 0; n < demods->size(); n++)
			if ((int)demods->at(n)->dec->get_type() == f.type) dec = demods->at(n)->dec;
		if (!dec) continue;
		sensordata_t sd[8];
		int n = 0;
		for (int j = 0; j < f.n_records && n < 8; j++) {
			const tfr_record &r = recs[f.first_record + j];
			sd[n].type = (sensor_e)r.type;
			sd[n].id = r.id;
			sd[n].temp = r.temp;
			sd[n].humidity = r.humidity;
			sd[n].alarm = r.alarm;
			sd[n].flags = r.flags;
			sd[n].sequence = r.sequence;
			sd[n].ts = (time_t)r.ts;
			sd[n].rssi = r.rssi;
			n++;
		}
		dec->deliver_frame(f, sd, n);
	}
	decoder::flush_exec();
	tfr_clear_results(hh);
	return (int)nf;
}
