/* TEST INFRASTRUCTURE ONLY -- minimal stand-in for <hackrf.h> so that the
 * reference vendor/BTLE/host/btle-tools/src/btle_rx.c compiles unmodified
 * (no radio is ever opened by the oracle harness).  Declares exactly the
 * libhackrf names btle_rx.c:489-660 touches. */
#ifndef SNRX_ORACLE_STUB_HACKRF_H
#define SNRX_ORACLE_STUB_HACKRF_H
#include <stdint.h>
typedef struct hackrf_device hackrf_device;
typedef struct {
    hackrf_device* device;
    uint8_t* buffer;
    int buffer_length;
    int valid_length;
    void* rx_ctx;
    void* tx_ctx;
} hackrf_transfer;
typedef int (*hackrf_sample_block_cb_fn)(hackrf_transfer*);
enum { HACKRF_SUCCESS = 0, HACKRF_TRUE = 1, HACKRF_ERROR_OTHER = -9999 };
static int hackrf_init(void) { return HACKRF_ERROR_OTHER; }
static int hackrf_exit(void) { return HACKRF_SUCCESS; }
static int hackrf_open(hackrf_device** d) { (void)d; return HACKRF_ERROR_OTHER; }
static int hackrf_close(hackrf_device* d) { (void)d; return HACKRF_SUCCESS; }
static int hackrf_set_freq(hackrf_device* d, uint64_t f) { (void)d; (void)f; return HACKRF_SUCCESS; }
static int hackrf_set_sample_rate(hackrf_device* d, double r) { (void)d; (void)r; return HACKRF_SUCCESS; }
static int hackrf_set_baseband_filter_bandwidth(hackrf_device* d, uint32_t b) { (void)d; (void)b; return HACKRF_SUCCESS; }
static int hackrf_set_vga_gain(hackrf_device* d, uint32_t g) { (void)d; (void)g; return HACKRF_SUCCESS; }
static int hackrf_set_lna_gain(hackrf_device* d, uint32_t g) { (void)d; (void)g; return HACKRF_SUCCESS; }
static int hackrf_stop_rx(hackrf_device* d) { (void)d; return HACKRF_SUCCESS; }
static int hackrf_start_rx(hackrf_device* d, hackrf_sample_block_cb_fn cb, void* ctx) { (void)d; (void)cb; (void)ctx; return HACKRF_ERROR_OTHER; }
static int hackrf_is_streaming(hackrf_device* d) { (void)d; return HACKRF_TRUE; }
static const char* hackrf_error_name(int e) { (void)e; return "stub"; }
#endif
