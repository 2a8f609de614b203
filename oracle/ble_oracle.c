/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into, imported by or
 * executed from the product path (snout_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Plain-C restatement ("port") of the BLE receive algorithm of the reference
 *   vendor/BTLE/host/btle-tools/src/btle_rx.c
 * for hosts where /root/reference is not available (the GPU box).  It is
 * pinned against the reference itself: tests/test_oracle_ble.py checks it,
 * frame for frame, against oracle/_ref/libbtle_ref.so (the unmodified
 * reference compiled by oracle/Makefile) on the reference's golden capture
 * vendor/BTLE/matlab/sample_iq_4msps.txt, on the usrp_replay_example vector
 * and on seeded synthetic captures, and against the committed fixtures in
 * tests/golden/.
 *
 * Form: the reference keeps four circular 32-entry bit histories and compares
 * them entry by entry (search_unique_bits, btle_rx.c:1369-1421).  Here the same
 * decision is taken on four 32-bit shift registers over a pre-sliced bit
 * array; the observable behaviour (which sample is reported, what is decoded
 * after it, where the search resumes, what a window may read) is the
 * reference's, including its zero-initialised history at every search origin
 * (btle_rx.c:1377), which accepts a 31-bit match one symbol before the origin.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../include/snoutrx.h"

#define SPS 4                       /* SAMPLE_PER_SYMBOL, btle_rx.c:176           */
#define WIN_IQ SNRX_BLE_WINDOW      /* LEN_BUF/2 int8 = 8192 IQ, btle_rx.c:180-182 */
#define SPAN_INT8 (31 * 8 + 16384)  /* receiver() buf_len, btle_rx.c:2382          */
#define DEMOD_LIMIT_INT8 19392      /* demod_buf_len, btle_rx.c:2025               */

/* ---- tables, built from their definitions (not copied) --------------------- */
static uint32_t g_crc_tab[256];      /* equals crc_table[], btle_rx.c:897-930        */
static uint8_t g_whiten[40][42];     /* equals scramble_table[][], scramble_table.h:1 */
static int g_tables_ready = 0;

static void build_tables(void) {
    if (g_tables_ready) return;
    /* CRC-24 of BLE, polynomial x^24+x^10+x^9+x^6+x^4+x^3+x+1 processed LSB first:
     * reflected polynomial 0xDA6000 (crc_table[128] == 0xda6000). */
    for (int i = 0; i < 256; i++) {
        uint32_t r = (uint32_t)i;
        for (int k = 0; k < 8; k++) r = (r & 1) ? (r >> 1) ^ 0xDA6000u : (r >> 1);
        g_crc_tab[i] = r & 0xFFFFFFu;
    }
    /* whitening: 7-bit LFSR x^7+x^4+1, position 0 = 1, positions 1..6 = channel
     * MSB first; output = position 6; bytes packed LSB first
     * (vendor/BTLE/matlab/scramble_gen.m:1-33). */
    for (int ch = 0; ch < 40; ch++) {
        int reg[7];
        reg[0] = 1;
        for (int k = 0; k < 6; k++) reg[1 + k] = (ch >> (5 - k)) & 1;
        for (int byte = 0; byte < 42; byte++) {
            uint8_t v = 0;
            for (int bit = 0; bit < 8; bit++) {
                int out = reg[6];
                v |= (uint8_t)(out << bit);
                /* shift: new[0] = old[6]; new[4] = old[3] ^ old[6] */
                int n0 = reg[6];
                for (int k = 6; k > 0; k--) reg[k] = reg[k - 1];
                reg[0] = n0;
                reg[4] ^= n0;
            }
            g_whiten[ch][byte] = v;
        }
    }
    g_tables_ready = 1;
}

/* crc_init_reorder, btle_rx.c:1801-1825: byte-swap the 24-bit value, then bit-reverse it */
static uint32_t crc_init_internal(uint32_t cmdline) {
    uint32_t swapped = ((cmdline & 0xFF) << 16) | (cmdline & 0xFF00) | ((cmdline >> 16) & 0xFF);
    uint32_t r = 0;
    for (int i = 0; i < 24; i++) r |= ((swapped >> i) & 1u) << (23 - i);
    return r;
}

static uint32_t crc24(const uint8_t* d, int n, uint32_t crc) {
    for (int i = 0; i < n; i++) crc = (g_crc_tab[(crc ^ d[i]) & 0xFF] ^ (crc >> 8)) & 0xFFFFFFu;
    return crc;                                     /* crc_update, btle_rx.c:1137-1148 */
}

/* ---- slicer ---------------------------------------------------------------- */
/* b[n] = (I[n]*Q[n+1] - I[n+1]*Q[n]) > 0   (btle_rx.c:1357-1361, 1385-1392).
 * `lead` zero samples are placed before sample 0, `n_total` bits are produced. */
static uint8_t* slice_bits(const int8_t* iq, int64_t n_iq, int64_t lead, int64_t n_total) {
    uint8_t* b = (uint8_t*)calloc((size_t)n_total, 1);
    if (!b) return NULL;
    for (int64_t n = 0; n + 1 < n_iq; n++) {
        int i0 = iq[2 * n], q0 = iq[2 * n + 1], i1 = iq[2 * n + 2], q1 = iq[2 * n + 3];
        b[lead + n] = (uint8_t)((i0 * q1 - i1 * q0) > 0);
    }
    return b;                  /* the last sample pairs with a zero sample -> bit 0 */
}

static void take_bytes(const uint8_t* b, int64_t first, int nbytes, uint8_t* out) {
    for (int i = 0; i < nbytes; i++) {             /* demod_byte, btle_rx.c:1348-1367 */
        uint8_t v = 0;
        for (int k = 0; k < 8; k++) v |= (uint8_t)(b[first + (int64_t)SPS * (8 * i + k)] << k);
        out[i] = v;
    }
}

/* One window.  `b` points at the bit of IQ sample 0 of the capture (negative
 * indices down to -4 are readable).  Appends to out[], returns new count. */
static int window_frames(const uint8_t* b, int64_t w, int channel, uint32_t aa, uint32_t aa_mask,
                         uint32_t crc_int, snrx_frame_t* out, int cap, int n) {
    const int64_t W = w * (int64_t)WIN_IQ;
    const int adv = (channel >= 37 && channel <= 39);         /* btle_rx.c:2034 */
    int eaten = 0;                                             /* int8 units, as the reference counts */
    for (;;) {
        int slots = (SPAN_INT8 - eaten) / (SPS * 2);           /* num_symbol_left, btle_rx.c:2032,2075,2123 */
        const int64_t P = W + eaten / 2;                       /* search origin in IQ samples */
        uint32_t reg[SPS] = {0, 0, 0, 0};                      /* memset, btle_rx.c:1377 */
        int64_t hit_n = -1;
        for (int t = 0; t < slots && hit_n < 0; t++) {
            for (int j = 0; j < SPS; j++) {
                int64_t pos = P + (int64_t)SPS * t + j;
                reg[j] = (reg[j] >> 1) | ((uint32_t)b[pos] << 31);
                if (((reg[j] ^ aa) & aa_mask) == 0) { hit_n = pos; break; }
            }
        }
        if (hit_n < 0) break;
        const int64_t s = hit_n - 31 * SPS;                    /* sample of AA bit 0, btle_rx.c:1409 */
        eaten = (int)(2 * (s - W)) + 32 * SPS * 2;              /* past the access address, btle_rx.c:2058 */
        const int64_t hdr_at = s + 32 * SPS;
        eaten += 16 * SPS * 2;                                 /* 2 header bytes, btle_rx.c:2066 */
        if (eaten > DEMOD_LIMIT_INT8) break;                   /* btle_rx.c:2067 */
        uint8_t bytes[48];
        take_bytes(b, hdr_at, 2, bytes);
        bytes[0] ^= g_whiten[channel][0];
        bytes[1] ^= g_whiten[channel][1];
        int len = adv ? (bytes[1] & 0x3F) : (bytes[1] & 0x1F); /* btle_rx.c:1794 / 1776 */
        if (adv && (len < 6 || len > 37)) continue;            /* btle_rx.c:2096-2104 */
        const int more = len + 3;
        eaten += 8 * more * SPS * 2;                           /* btle_rx.c:2113 */
        if (eaten > DEMOD_LIMIT_INT8) break;                   /* btle_rx.c:2115 */
        take_bytes(b, hdr_at + 16 * SPS, more, bytes + 2);
        for (int i = 0; i < more; i++) bytes[2 + i] ^= g_whiten[channel][2 + i];
        uint32_t calc = crc24(bytes, len + 2, crc_int);
        uint32_t recv = (uint32_t)bytes[len + 2] | ((uint32_t)bytes[len + 3] << 8) | ((uint32_t)bytes[len + 4] << 16);
        if (n < cap) {
            snrx_frame_t* f = &out[n];
            memset(f, 0, sizeof(*f));
            f->sample_index = s;
            f->window = (uint32_t)w;
            f->channel = (uint16_t)channel;
            f->proto = SNRX_PROTO_BLE;
            f->crc_ok = (uint8_t)(calc == recv);               /* crc_check returns "differs", btle_rx.c:1847 */
            f->phase = (uint8_t)(((s % 4) + 4) % 4);
            f->len = (uint16_t)(len + 5);
            f->access_addr = aa;
            memcpy(f->bytes, bytes, (size_t)(len + 5));
        }
        n++;
    }
    return n;
}

#define LEAD_BITS 8
#define TAIL_BITS (DEMOD_LIMIT_INT8 / 2 + 64)

int64_t ble_oracle_num_windows(int64_t n_iq) { return (n_iq + WIN_IQ - 1) / WIN_IQ; }

/* iq: n_iq interleaved int8 I,Q samples of one channel at 4 Msps; samples past
 * the end are zero.  Windows first_window .. first_window+n_windows-1 are
 * processed (n_windows <= 0: all).  Returns frames found (stores at most cap). */
int ble_oracle_windows_range(const int8_t* iq, int64_t n_iq, int channel, uint32_t aa, uint32_t aa_mask,
                             uint32_t crc_init_cmdline, int64_t first_window, int64_t n_windows,
                             snrx_frame_t* out, int cap) {
    build_tables();
    if (channel < 0 || channel > 39) return -1;
    int64_t nw_all = ble_oracle_num_windows(n_iq);
    if (n_windows <= 0) n_windows = nw_all - first_window;
    int64_t n_bits = LEAD_BITS + nw_all * (int64_t)WIN_IQ + TAIL_BITS;
    uint8_t* store = slice_bits(iq, n_iq, LEAD_BITS, n_bits);
    if (!store) return -1;
    const uint32_t crc_int = crc_init_internal(crc_init_cmdline);
    int n = 0;
    for (int64_t w = first_window; w < first_window + n_windows && w < nw_all; w++)
        n = window_frames(store + LEAD_BITS, w, channel, aa, aa_mask, crc_int, out, cap, n);
    free(store);
    return n;
}

int ble_oracle_windows(const int8_t* iq, int64_t n_iq, int channel, uint32_t aa,
                       uint32_t crc_init_cmdline, snrx_frame_t* out, int cap) {
    return ble_oracle_windows_range(iq, n_iq, channel, aa, 0xFFFFFFFFu, crc_init_cmdline, 0, 0, out, cap);
}

/* cf32 (interleaved float I,Q) -> int8 grid: q = clamp(rint(x*scale), -128, 127)
 * (SURVEY 8c "cf32 -> int8 for the BLE oracle"; round-half-even like lrintf). */
void ble_oracle_quantize(const float* x, int64_t n_floats, float scale, int8_t* q) {
    for (int64_t i = 0; i < n_floats; i++) {
        float v = x[i] * scale;
        float r = __builtin_rintf(v);
        if (r > 127.0f) r = 127.0f;
        if (r < -128.0f) r = -128.0f;
        q[i] = (int8_t)r;
    }
}

double ble_oracle_time(const int8_t* iq, int64_t n_iq, int channel, uint32_t aa,
                       uint32_t crc_init_cmdline, int reps, int* frames_out) {
    struct timespec t0, t1;
    int n = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; r++)
        n = ble_oracle_windows(iq, n_iq, channel, aa, crc_init_cmdline, NULL, 0);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (frames_out) *frames_out = n;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* table access for known-answer tests */
const uint8_t* ble_oracle_whiten_row(int ch) { build_tables(); return g_whiten[ch]; }
uint32_t ble_oracle_crc_table(int i) { build_tables(); return g_crc_tab[i & 255]; }
uint32_t ble_oracle_crc_init_internal(uint32_t x) { return crc_init_internal(x); }
uint32_t ble_oracle_crc24(const uint8_t* d, int n, uint32_t init_internal) { build_tables(); return crc24(d, n, init_internal); }
int ble_oracle_channel_mhz(int ch) {               /* get_freq_by_channel_number, btle_rx.c:932-948 */
    if (ch == 37) return 2402;
    if (ch == 38) return 2426;
    if (ch == 39) return 2480;
    if (ch >= 0 && ch <= 10) return 2404 + 2 * ch;
    if (ch >= 11 && ch <= 36) return 2428 + 2 * (ch - 11);
    return -1;
}
