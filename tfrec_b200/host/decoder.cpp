// decoder.cpp - host half of the plugin surface: result delivery, the first-seen / duplicate rule and the -e exec
// contract of decoder.cpp:46-109 (baycom/tfrec), and the stdout lines the reference prints from its flush()
// functions.  Parsing itself happened on the device; the line formatter below only re-reads the raw frame
// bytes for the values the reference prints but does not store in sensordata_t (PTEMP, PHUM, rain, ...).
#include "decoder.h"

#include <inttypes.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/tfr.h"

// ---------------------------------------------------------------------------------------------- helpers
namespace {

unsigned be(const uint8_t *p, int n)
{
	unsigned v = 0;
	while (n--) v = (v << 8) | *p++;
	return v;
}
// display-only CRCs for the BAD lines (tfa1.cpp:108, tfa2.cpp:208,268, whb.cpp:553); poly 0x31 / 0x04c11db7
unsigned crc8_31(const uint8_t *d, int len)
{
	unsigned c = 0;
	for (int n = 0; n < len; n++) {
		c ^= d[n];
		for (int m = 0; m < 8; m++) c = (c & 0x80) ? ((c << 1) ^ 0x31) & 0xff : (c << 1) & 0xff;
	}
	return c;
}
uint32_t crc32_msb(const uint8_t *d, int len, uint32_t c)
{
	for (int n = 0; n < len; n++) {
		c ^= (uint32_t)d[n] << 24;
		for (int m = 0; m < 8; m++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04c11db7u : c << 1;
	}
	return c;
}
bool whb_init(unsigned t, uint32_t &v)
{
	static const uint32_t tab[][2] = { { 0x02, 0x97d97a26 }, { 0x03, 0xf59c5a1e }, { 0x04, 0x98e1d11f }, { 0x06, 0xa7a41254 },
		{ 0x07, 0x3303fb1d }, { 0x08, 0x29f0f49b }, { 0x09, 0xa7a41254 }, { 0x0b, 0xe7720ae4 }, { 0x10, 0x62d0afc1 },
		{ 0x11, 0x8cba0708 }, { 0x12, 0x5a9e30ae } };
	for (auto &e : tab)
		if (e[0] == t) { v = e[1]; return true; }
	return false;
}
double t11(unsigned raw, bool ext = false)   // 11 (or 12) bit two's complement tenths, whb.cpp:109-123 as built (x*0.1)
{
	const unsigned sign = ext ? 0x800 : 0x400, mask = ext ? 0xfff : 0x7ff;
	raw &= mask;
	return (raw & sign) ? -(double)((int)((raw ^ mask) + 1)) * 0.1 : raw * 0.1;
}
double koffs(int offset) { return -1536.0 * offset / 131072; }

void append(std::string &s, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void append(std::string &s, const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	s += buf;
}

}  // namespace

// the dbg>=0 line of each flush(): tfa1.cpp:87, tfa2.cpp:140-153, 245-249, whb.cpp:136,152,182,215,244,274,307,334,362,388
std::string tfr_format_line(const tfr_frame &f, const sensordata_t *recs, int n_recs, int dbg)
{
	std::string s;
	const uint8_t *r = f.rdata;
	if (f.status != 0 || n_recs < 0) return s;
	switch (f.type) {
	case TFA_1:
		if (n_recs >= 1)
			append(s, "TFA1 ID %04x %+.1f %i%% seq %x lowbat %i RSSI %i", (int)recs[0].id, recs[0].temp, (int)recs[0].humidity,
			       recs[0].sequence, recs[0].alarm, f.rssi);
		break;
	case TFA_2:
	case TFA_3:
		if (n_recs >= 1)
			append(s, "TFA%i ID %06x %+.1lf %i%% RSSI %i Offset %.0lfkHz", f.type + 1, (int)recs[0].id, recs[0].temp,
			       (int)recs[0].humidity, f.rssi, koffs(f.offset));
		break;
	case TX22: {
		const int num = r[3] & 7;
		append(s, "TX22 ID %x, ", (TX22 << 28) | ((((r[2] & 0xf) << 2) | (r[3] >> 6)) << 4));
		for (int pass = 0; pass < 5; pass++)   // the reference prints temp, hum, rain, wind, gust in that order
			for (int n = num - 1; n >= 0; n--) {   // the last word of a kind wins (tfa2.cpp:107-147 overwrites)
				const uint8_t *w = r + 4 + 2 * n;
				if ((w[0] >> 4) != pass) continue;
				const int bcd = (w[0] & 0xf) * 100 + (w[1] >> 4) * 10 + (w[1] & 0xf), bin = ((w[0] & 0xf) << 8) + w[1];
				if (pass == 0) append(s, "temp %g, ", bcd * 0.1 - 40);
				if (pass == 1) append(s, "hum %g, ", (double)bcd);
				if (pass == 2) append(s, "rain %g, ", (double)bin);
				if (pass == 3) append(s, "speed %g, dir %g, ", w[1] * 0.1, (w[0] & 0xf) * 22.5);
				if (pass == 4) append(s, "gust %g, ", bin * 0.1);
				break;
			}
		append(s, "RSSI %i, offset %.0lfkHz", f.rssi, koffs(f.offset));
		break;
	}
	case TFA_WHB: {
		const unsigned t = r[5];
		const uint8_t *m = r + 11;
		uint64_t id = 0;
		for (int k = 0; k < 6; k++) id = (id << 8) | r[5 + k];
		const unsigned long long pid = id;
		switch (t) {
		case 0x02: append(s, "WHB02 ID %llx TEMP %g, PTEMP %g", pid, t11(be(m + 2, 2)), t11(be(m + 4, 2))); break;
		case 0x03:
			append(s, "WHB03 ID %llx TEMP %g HUM %i, PTEMP %g PHUM %i", pid, t11(be(m + 2, 2)), be(m + 4, 2) & 0xff,
			       t11(be(m + 6, 2)), be(m + 8, 2) & 0xff);
			break;
		case 0x04:
			append(s, "WHB04 ID %llx TEMP %g HUM %i WET %i, PTEMP %g PHUM %i PWET %i", pid, t11(be(m + 2, 2)), be(m + 4, 2) & 0xff,
			       (m[6] & 1) ^ 1, t11(be(m + 7, 2)), be(m + 9, 2) & 0xff, (m[11] & 1) ^ 1);
			break;
		case 0x06:
		case 0x09: {
			const bool x = (t == 0x09);
			append(s, "WHB0%i ID %llxTEMP %g HUM %i TEMP2 %g, PTEMP %g PHUM %i PTEMP2 %g", x ? 9 : 6, pid, t11(be(m + 2, 2)),
			       be(m + 6, 2) & 0xff, t11(be(m + 4, 2), x), t11(be(m + 8, 2)), be(m + 12, 2) & 0xff, t11(be(m + 10, 2), x));
			break;
		}
		case 0x07:
			append(s, "WHB07 ID %llx TEMP_IN %g HUM_IN %i TEMP_OUT %g HUM_OUT %i", pid, t11(be(m + 2, 2)), be(m + 4, 2) & 0xff,
			       t11(be(m + 6, 2)), be(m + 8, 2) & 0xff);
			if (dbg > 1)
				append(s, " PTEMP_IN %g PHUM_IN %i PTEMP_OUT %g PHUM_OUT %i", t11(be(m + 10, 2)), be(m + 12, 2) & 0xff,
				       t11(be(m + 14, 2)), be(m + 16, 2) & 0xff);
			break;
		case 0x08: {
			static const unsigned tu[4] = { 86400, 3600, 60, 1 };   // timeunit_tab, whb.cpp:65-70
			append(s, "WHB08 ID %llx cnt %i", pid, be(m + 4, 2));
			if (dbg > 1)   // whb.cpp:311-312, one line per stored time
				for (int i = 0; i < 10; i++) {
					const unsigned x = be(m + 6 + 2 * i, 2);
					append(s, "\nWHB08 ID %llx #%i time %i", pid, i, tu[(x >> 14) & 3] * (x & 0x3fff));
				}
			break;
		}
		case 0x0b:
			for (int i = 0; i < 6 && (i == 0 || dbg > 0); i++) {   // whb.cpp:344-348: the history entries with -D
				const unsigned v = be(m + 3 + 4 * i, 4);
				const float dir = 22.5 * (v >> 28), speed = (((v >> 16) & 0xff) + 256 * ((v >> 25) & 1)) * 0.1,
					    gust = (((v >> 8) & 0xff) + 256 * ((v >> 24) & 1)) * 0.1;
				append(s, "%sWHB0b ID %llx #%i DIR %f SPEED %f GUST %f time %i", i ? "\n" : "", pid, i, dir, speed, gust, (v & 0xff) * 2);
			}
			break;
		case 0x10: {
			static const unsigned tu[4] = { 86400, 3600, 60, 1 };
			for (int i = 0; i < 4 && (i == 0 || dbg > 0); i++) {   // whb.cpp:376-379
				const unsigned x = be(m + 2 + 2 * i, 2);
				append(s, "%sWHB10 ID %llx #%i %i %i", i ? "\n" : "", pid, i, x >> 15, tu[(x >> 13) & 3] * (x & 0x1fff));
			}
			break;
		}
		case 0x11:
			append(s, "WHB11 %llx TEMP1 %g HUM1 %i TEMP2 %g HUM2 %i TEMP3 %g HUM3 %i TEMP_IN %g HUM_IN %i", pid, t11(be(m + 2, 2)),
			       be(m + 4, 2) & 0xff, t11(be(m + 6, 2)), be(m + 8, 2) & 0xff, t11(be(m + 10, 2)), be(m + 12, 2) & 0xff,
			       t11(be(m + 14, 2)), be(m + 16, 2) & 0xff);
			if (dbg > 1)   // whb.cpp:410-412
				append(s, " PTEMP1 %g PHUM1 %i PTEMP2 %g PHUM2 %i PTEMP3 %g PHUM3 %i PTEMP_IN %g PHUM_IN %i", t11(be(m + 18, 2)),
				       be(m + 20, 2) & 0xff, t11(be(m + 22, 2)), be(m + 24, 2) & 0xff, t11(be(m + 26, 2)), be(m + 28, 2) & 0xff,
				       t11(be(m + 30, 2)), be(m + 32, 2) & 0xff);
			break;
		case 0x12:
			append(s, "WHB12 %llx TEMP %g HUM %i HUM3h %i HUM24h %i HUM7d %i HUM30d %i", pid, t11(be(m + 6, 2)), m[8] & 0x7f,
			       m[2] & 0x7f, m[3] & 0x7f, m[4] & 0x7f, m[5] & 0x7f);
			break;
		}
		break;
	}
	default: break;
	}
	return s;
}

// ---------------------------------------------------------------------------------------------- decoder
decoder::decoder(sensor_e _type)
	: dbg(0), bad(0), synced(0), type(_type), byte_cnt(0), snum(0), handler(NULL), mode(0), handle(NULL)
{
	memset(rdata, 0, sizeof(rdata));
}

void decoder::set_params(char *_handler, int _mode, int _dbg)
{
	handler = _handler;
	mode = _mode;
	dbg = _dbg;
}

void decoder::store_bit(int) {}

void decoder::store_bytes(uint8_t *d, int len)
{
	if (len > (int)sizeof(rdata)) len = (int)sizeof(rdata);
	memcpy(rdata, d, len);
	byte_cnt = len;
	synced = 1;
}

// -X path: hand the stored bytes to the device parser and present the result like the reference's flush()
void decoder::flush(int rssi, int offset)
{
	(void)offset;
	if (!handle) {
		fprintf(stderr, "decoder::flush: no device handle attached (the parsers run on the GPU; there is no CPU fallback)\n");
		return;
	}
	tfr_frame f;
	tfr_record rec[8];
	const int n = tfr_parse_bytes(handle, (int)type, rdata, byte_cnt, &f, rec, 8);
	if (n < 0) {
		fprintf(stderr, "tfr_parse_bytes: %s\n", tfr_last_error());
		return;
	}
	if (f.status == -1) {   // shorter than the type's minimum frame: the reference's flush() does nothing either
		byte_cnt = 0;
		synced = 0;
		return;
	}
	f.rssi = rssi;
	sensordata_t sd[8];
	for (int k = 0; k < n && k < 8; k++) {
		sd[k].type = (sensor_e)rec[k].type;
		sd[k].id = rec[k].id;
		sd[k].temp = rec[k].temp;
		sd[k].humidity = rec[k].humidity;
		sd[k].alarm = rec[k].alarm;
		sd[k].flags = rec[k].flags;
		sd[k].sequence = rec[k].sequence;
		sd[k].ts = (time_t)rec[k].ts;
		sd[k].rssi = rssi;
	}
	deliver_frame(f, sd, n);
	byte_cnt = 0;
	synced = 0;
}

// what the tail of each reference flush() does once the frame is parsed: hexdump (dbg != 0), the decoded or
// BAD line, and store_data() per record
void decoder::deliver_frame(const tfr_frame &f, sensordata_t *recs, int n_recs)
{
	const uint8_t *r = f.rdata;
	if (f.status == 3) {   // tfa2.cpp:294-300: printed unconditionally, once per inverted sync word
		for (int k = 0; k < f.byte_cnt; k++) printf("Inverted SYNC\n");
		return;
	}
	if (dbg) {
		int n = f.byte_cnt;
		if (type == TFA_1) n = 11;
		else if (type == TFA_2 || type == TFA_3) n = 7;
		if (n > TFR_MAX_RDATA) n = TFR_MAX_RDATA;
		if (type == TFA_WHB) printf("#%03i %u L=%i  ", snum++, (uint32_t)time(0), f.byte_cnt);
		else printf("#%03i %u  ", snum++, (uint32_t)time(0));
		for (int k = 0; k < n; k++) printf("%02x ", r[k]);
		if (type == TFA_1) printf("          ");
		else if (type == TX22) printf("      ");
		else if (type == TFA_WHB) printf(" RSSI %i ", f.rssi);
		else printf("                      ");
	}
	if (f.status == 0) {
		if (dbg >= 0) {
			const std::string line = tfr_format_line(f, recs, n_recs, dbg);
			if (!line.empty()) puts(line.c_str());
			fflush(stdout);
		}
		for (int k = 0; k < n_recs; k++) store_data(recs[k]);
		return;
	}
	bad++;
	// whb.cpp:505-508 prints this at dbg >= 0, i.e. also without -D
	uint32_t whb_unknown_init = 0;
	const bool whb_unknown = type == TFA_WHB && f.status != 1 && r[4] <= 60 && !whb_init(r[5], whb_unknown_init);
	if (whb_unknown && dbg >= 0) printf("WHB: Probably unsupported sensor type %02x! Please report\n", r[5]);
	if (!dbg) {
		fflush(stdout);
		return;
	}
	if (type == TFA_1) {
		if (f.status == 1) printf("TFA1 BAD %i RSSI %i (CRC %02x %02x)\n", bad, f.rssi, r[10], crc8_31(r + 2, 8));
		else printf("TFA1 BAD %i RSSI %i (SANITY)\n", bad, f.rssi);
	} else if (type == TFA_2 || type == TFA_3) {
		printf("TFA%i BAD %i RSSI %i  Offset %.0lfkHz (CRC %02x %02x)\n", type + 1, bad, f.rssi, koffs(f.offset), r[6], crc8_31(r + 2, 4));
	} else if (type == TX22) {
		const int num = r[3] & 7;
		if (f.status == 1)
			printf("TX22(%02x) BAD %i RSSI %i  Offset %.0lfkHz (CRC %02x %02x) len %i\n", 1 << type, bad, f.rssi, koffs(f.offset),
			       r[2 * num + 4], crc8_31(r + 2, 2 + 2 * num), f.byte_cnt);
		else
			printf("TX22(%02x) BAD %i RSSI %i  Offset %.0lfkHz len %i (SANITY)\n", 1 << type, bad, f.rssi, koffs(f.offset), f.byte_cnt);
	} else if (type == TFA_WHB) {
		const int plen = r[4];
		uint32_t init = 0;
		if (f.status == 1 && plen <= 60 && whb_init(r[5], init))
			printf("\nWHB BAD %i RSSI %i (CRC is %08x, should be %08x, len %i, plen %i)\n", bad, f.rssi, be(r + plen, 4),
			       crc32_msb(r + 4, plen > 4 ? plen - 4 : 0, init), f.byte_cnt, plen);
		else
			printf("\nWHB BAD %i RSSI %i (SANITY)\n", bad, f.rssi);
	}
	fflush(stdout);
}

// first appearance of an id is stored; a WeatherHub repeat with the same sequence number is not executed
// again (decoder.cpp:46-65)
void decoder::store_data(sensordata_t &d)
{
	bool repeat = false;
	std::map<uint64_t, sensordata_t>::iterator it = data.find(d.id);
	if (it == data.end()) {
		data.insert(std::make_pair(d.id, d));
	} else if (it->second.type == TFA_WHB) {
		if (it->second.sequence == d.sequence) repeat = true;
		else it->second.sequence = d.sequence;
	}
	if (mode == 0 && !repeat) execute_handler(d);
}

// Opt-in sink for high telegram rates (TFREC_EXEC=async in the environment): the reference forks the whole process and
// a shell for every telegram and waits for the handler (decoder.cpp:95 `system(cmd)`), which caps it at a few hundred
// telegrams per second.  Here ONE /bin/sh is started at the first telegram and every command line is written to its
// standard input: same commands, same order, executed one after the other by that shell, but the decoder does not
// wait for them.  The pipe is flushed after every delivered batch and closed (waited for) at exit.  Default: system().
namespace {
FILE *g_exec_sh = NULL;
int g_exec_mode = -1;   // -1 not looked up yet, 0 system() per telegram, 1 one shell fed through a pipe
void exec_close()
{
	if (g_exec_sh) {
		pclose(g_exec_sh);
		g_exec_sh = NULL;
	}
}
bool exec_async(const char *cmd)
{
	if (g_exec_mode < 0) {
		const char *e = getenv("TFREC_EXEC");
		g_exec_mode = (e && !strcmp(e, "async")) ? 1 : 0;
	}
	if (!g_exec_mode) return false;
	if (!g_exec_sh) {
		fflush(stdout);   // what has been printed so far stays in front of the handlers' output
		g_exec_sh = popen("/bin/sh", "w");
		if (!g_exec_sh) {
			perror("popen /bin/sh");
			g_exec_mode = 0;
			return false;
		}
		atexit(exec_close);
	}
	fputs(cmd, g_exec_sh);
	fputc('\n', g_exec_sh);
	return true;
}
}  // namespace

void decoder::flush_exec(void)
{
	if (g_exec_sh) fflush(g_exec_sh);
}

// "<handler> id temp hum seq alarm rssi flags ts" through system() (decoder.cpp:67-96)
void decoder::execute_handler(sensordata_t &d)
{
	if (!handler || !handler[0]) return;
	char cmd[512];
	if (type != TFA_WHB) {
		const uint64_t nid = d.id | (uint64_t)(int64_t)((int)d.type << 24);
		snprintf(cmd, sizeof(cmd), "%s %04" PRIx64 " %+.1f %g %i %i %i %i %li", handler, nid, d.temp, d.humidity, d.sequence, d.alarm,
			 d.rssi, d.flags, (long)d.ts);
	} else {
		snprintf(cmd, sizeof(cmd), "%s %013" PRIx64 " %+.1f %g %i %i %i %i %li", handler, d.id, d.temp, d.humidity, d.sequence,
			 d.alarm, d.rssi, d.flags, (long)d.ts);
	}
	if (dbg >= 1) printf("EXEC %s\n", cmd);
	if (exec_async(cmd)) return;
	if (system(cmd) == -1) perror("system");
}

void decoder::flush_storage(void)
{
	if (!mode) return;
	for (std::map<uint64_t, sensordata_t>::iterator it = data.begin(); it != data.end(); ++it) execute_handler(it->second);
	data.clear();
}

// ---------------------------------------------------------------------------------------------- demodulator
demodulator::demodulator(decoder *_dec) : dec(_dec), last_bit_idx(0) {}

void demodulator::start(int len)
{
	if (last_bit_idx) last_bit_idx -= len;
}

int demodulator::demod(int, int, int, int16_t *) { return 0; }
