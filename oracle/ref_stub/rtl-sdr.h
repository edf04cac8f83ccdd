/* Stand-in for librtlsdr's <rtl-sdr.h>, used ONLY to compile the unmodified reference
 * (/root/reference/sdr.cpp includes it, sdr.h:19) into oracle/_ref/.  librtlsdr is a hardware
 * I/O library that is absent in this image; in -L replay mode the reference never constructs
 * an sdr object (engine.cpp:30), so every entry point here just reports failure.
 * TEST INFRASTRUCTURE - never linked into the product library. */
#ifndef TFR_ORACLE_RTLSDR_STUB_H
#define TFR_ORACLE_RTLSDR_STUB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct rtlsdr_dev rtlsdr_dev_t;
typedef void (*rtlsdr_read_async_cb_t)(unsigned char *buf, uint32_t len, void *ctx);
static inline uint32_t rtlsdr_get_device_count(void) { return 0; }
static inline int rtlsdr_get_device_usb_strings(uint32_t i, char *m, char *p, char *s) { (void)i; (void)m; (void)p; (void)s; return -1; }
static inline int rtlsdr_open(rtlsdr_dev_t **d, uint32_t i) { (void)d; (void)i; return -1; }
static inline int rtlsdr_close(rtlsdr_dev_t *d) { (void)d; return -1; }
static inline int rtlsdr_reset_buffer(rtlsdr_dev_t *d) { (void)d; return -1; }
static inline int rtlsdr_cancel_async(rtlsdr_dev_t *d) { (void)d; return -1; }
static inline int rtlsdr_set_center_freq(rtlsdr_dev_t *d, uint32_t f) { (void)d; (void)f; return -1; }
static inline int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *d, int m) { (void)d; (void)m; return -1; }
static inline int rtlsdr_get_tuner_gains(rtlsdr_dev_t *d, int *g) { (void)d; (void)g; return -1; }
static inline int rtlsdr_set_tuner_gain(rtlsdr_dev_t *d, int g) { (void)d; (void)g; return -1; }
static inline int rtlsdr_set_freq_correction(rtlsdr_dev_t *d, int p) { (void)d; (void)p; return -1; }
static inline int rtlsdr_set_sample_rate(rtlsdr_dev_t *d, uint32_t r) { (void)d; (void)r; return -1; }
static inline int rtlsdr_read_async(rtlsdr_dev_t *d, rtlsdr_read_async_cb_t cb, void *ctx, uint32_t n, uint32_t l) { (void)d; (void)cb; (void)ctx; (void)n; (void)l; return -1; }
#ifdef __cplusplus
}
#endif
#endif
