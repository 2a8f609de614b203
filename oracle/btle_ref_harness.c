/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into the product.
 *
 * Compiles the UNMODIFIED reference BLE receiver from where it lies
 * (/root/reference/vendor/BTLE/host/btle-tools/src/btle_rx.c, passed as
 * -DREF_BTLE_RX_C=...) into oracle/_ref/libbtle_ref.so and drives it the way
 * its own main loop does (btle_rx.c:2341-2393): one receiver() call per
 * 16384-int8 half buffer, search span 31*8+16384, crc init reordered once.
 *
 * receiver() only prints.  To obtain records (bytes, CRC flag, sample index)
 * btle_ref_windows() walks the same steps as receiver() (btle_rx.c:2043-2154)
 * but every computation is done by the reference's own functions:
 * search_unique_bits, demod_byte, scramble_byte, parse_*_header_byte,
 * crc_check.  btle_ref_receiver_print() calls the real receiver() so a test
 * can compare its stdout with those records.
 */
#define main btle_rx_reference_main
#include REF_BTLE_RX_C
#undef main

#include "../include/snoutrx.h"

#define REF_HALF (LEN_BUF / 2)                                   /* 16384 int8 = 8192 IQ */
#define REF_SPAN ((LEN_DEMOD_BUF_ACCESS - 1) * 2 * SAMPLE_PER_SYMBOL + REF_HALF)
#define REF_DEMOD_LIMIT (LEN_BUF_MAX_NUM_PHY_SAMPLE + REF_HALF)  /* 19392 */
#define REF_LEAD 16                                              /* zero int8 before window 0 */
#define REF_TAIL (REF_DEMOD_LIMIT + 64)                          /* zero int8 after the capture */

static IQ_TYPE* padded_copy(const int8_t* iq, int64_t n_iq, int64_t n_windows) {
    int64_t total = REF_LEAD + n_windows * (int64_t)REF_HALF + REF_TAIL;
    IQ_TYPE* buf = (IQ_TYPE*)calloc((size_t)total, 1);
    if (!buf) return NULL;
    memcpy(buf + REF_LEAD, iq, (size_t)(2 * n_iq));
    return buf;
}

/* access-address mask of the next calls (-m option, btle_rx.c:2301); default all ones */
static uint32_t ref_mask = 0xFFFFFFFFu;
void btle_ref_set_mask(uint32_t m) { ref_mask = m; }

int64_t btle_ref_num_windows(int64_t n_iq) {
    return (n_iq + SNRX_BLE_WINDOW - 1) / SNRX_BLE_WINDOW;
}

/* one window, records instead of text */
static int one_window(IQ_TYPE* rxp_in, int64_t window, int channel, uint32_t aa,
                      uint32_t crc_init_internal, snrx_frame_t* out, int cap, int n) {
    IQ_TYPE* rxp = rxp_in;
    int buf_len = REF_SPAN;
    int num_symbol_left = buf_len / (SAMPLE_PER_SYMBOL * 2);
    int eaten = 0;
    int adv = (channel == 37 || channel == 38 || channel == 39);
    uint8_t bytes[2 + 37 + 3 + 8];

    uint32_to_bit_array(aa, access_bit);
    for (;;) {
        int hit = search_unique_bits(rxp, num_symbol_left, access_bit, access_bit_mask,
                                     LEN_DEMOD_BUF_ACCESS);
        if (hit == -1) break;
        eaten += hit;
        int64_t s_int8 = window * (int64_t)REF_HALF + eaten;        /* AA bit 0, int8 units */
        eaten += 8 * NUM_ACCESS_ADDR_BYTE * 2 * SAMPLE_PER_SYMBOL;
        rxp = rxp_in + eaten;
        eaten += 8 * 2 * 2 * SAMPLE_PER_SYMBOL;
        if (eaten > REF_DEMOD_LIMIT) break;
        demod_byte(rxp, 2, bytes);
        scramble_byte(bytes, 2, scramble_table[channel], bytes);
        rxp = rxp_in + eaten;
        num_symbol_left = (buf_len - eaten) / (SAMPLE_PER_SYMBOL * 2);

        int payload_len;
        if (adv) {
            ADV_PDU_TYPE t; int ta, ra;
            parse_adv_pdu_header_byte(bytes, &t, &ta, &ra, &payload_len);
            if (payload_len < 6 || payload_len > 37) continue;
        } else {
            LL_PDU_TYPE t; int a, b, c;
            parse_ll_pdu_header_byte(bytes, &t, &a, &b, &c, &payload_len);
        }
        int nb = payload_len + 3;
        eaten += 8 * nb * 2 * SAMPLE_PER_SYMBOL;
        if (eaten > REF_DEMOD_LIMIT) break;
        demod_byte(rxp, nb, bytes + 2);
        scramble_byte(bytes + 2, nb, scramble_table[channel] + 2, bytes + 2);
        rxp = rxp_in + eaten;
        num_symbol_left = (buf_len - eaten) / (SAMPLE_PER_SYMBOL * 2);
        int crc_flag = crc_check(bytes, payload_len + 2, crc_init_internal);

        if (n < cap) {
            snrx_frame_t* f = &out[n];
            memset(f, 0, sizeof(*f));
            f->sample_index = (s_int8 >= 0) ? s_int8 / 2 : -((-s_int8) / 2);
            f->window = (uint32_t)window;
            f->channel = (uint16_t)channel;
            f->proto = SNRX_PROTO_BLE;
            f->crc_ok = (uint8_t)(crc_flag == 0);
            f->phase = (uint8_t)(((s_int8 / 2) % 4 + 4) % 4);
            f->len = (uint16_t)(payload_len + 5);
            f->access_addr = aa;
            memcpy(f->bytes, bytes, (size_t)(payload_len + 5));
        }
        n++;
    }
    return n;
}

/* iq: n_iq interleaved int8 (I,Q) samples of ONE channel at 4 Msps.
 * Returns the number of frames found (may exceed cap; only cap are stored). */
int btle_ref_windows(const int8_t* iq, int64_t n_iq, int channel, uint32_t aa,
                     uint32_t crc_init_cmdline, snrx_frame_t* out, int cap) {
    int64_t nw = btle_ref_num_windows(n_iq);
    IQ_TYPE* buf = padded_copy(iq, n_iq, nw);
    if (!buf) return -1;
    uint32_to_bit_array(ref_mask, access_bit_mask);                  /* btle_rx.c:2301 */
    uint32_t crc_int = crc_init_reorder(crc_init_cmdline);           /* btle_rx.c:2335 */
    int n = 0;
    for (int64_t w = 0; w < nw; w++)
        n = one_window(buf + REF_LEAD + w * (int64_t)REF_HALF, w, channel, aa, crc_int, out, cap, n);
    free(buf);
    return n;
}

/* the real receiver(), printing to stdout exactly like btle_rx does */
int btle_ref_receiver_print(const int8_t* iq, int64_t n_iq, int channel, uint32_t aa,
                            uint32_t crc_init_cmdline) {
    int64_t nw = btle_ref_num_windows(n_iq);
    IQ_TYPE* buf = padded_copy(iq, n_iq, nw);
    if (!buf) return -1;
    uint32_to_bit_array(ref_mask, access_bit_mask);
    uint32_t crc_int = crc_init_reorder(crc_init_cmdline);
    for (int64_t w = 0; w < nw; w++) {
        receiver(buf + REF_LEAD + w * (int64_t)REF_HALF, REF_SPAN, channel, aa, crc_int, 0, 0);
        fflush(stdout);
    }
    free(buf);
    return 0;
}

/* time the reference path: `reps` passes over the capture, text to /dev/null is the
 * caller's business (it redirects stdout).  Returns seconds. */
double btle_ref_time(const int8_t* iq, int64_t n_iq, int channel, uint32_t aa,
                     uint32_t crc_init_cmdline, int reps, int* frames_out) {
    int64_t nw = btle_ref_num_windows(n_iq);
    IQ_TYPE* buf = padded_copy(iq, n_iq, nw);
    if (!buf) return -1.0;
    uint32_to_bit_array(ref_mask, access_bit_mask);
    uint32_t crc_int = crc_init_reorder(crc_init_cmdline);
    static snrx_frame_t scratch[4];
    struct timespec t0, t1;
    int n = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; r++) {
        n = 0;
        for (int64_t w = 0; w < nw; w++)
            n = one_window(buf + REF_LEAD + w * (int64_t)REF_HALF, w, channel, aa, crc_int, scratch, 0, n);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(buf);
    if (frames_out) *frames_out = n;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* tables and small functions of the reference, for known-answer tests */
const uint8_t* btle_ref_scramble_row(int channel) { return scramble_table[channel]; }
uint32_t btle_ref_crc_table(int i) { return (uint32_t)crc_table[i & 255]; }
uint32_t btle_ref_crc_init_reorder(uint32_t x) { return crc_init_reorder(x); }
uint32_t btle_ref_crc24(const uint8_t* b, int n, uint32_t init_internal) { return (uint32_t)crc24_byte((uint8_t*)b, n, init_internal); }
uint64_t btle_ref_freq(int channel) { return get_freq_by_channel_number(channel); }

/* the reference's own CONNECT_REQ field extraction (parse_adv_pdu_payload_byte, btle_rx.c:1476-1557) and
 * what receiver_controller would start tracking (receiver_status).  payload: the 34 payload bytes.  out[0..11] = AA (as the
 * receiver uses it), CRCInit, WinSize, WinOffset, Interval, Latency, Timeout, Hop, SCA, chm_is_full_map, receiver_status.hop,
 * receiver_status.interval; init_a / adv_a / chm: the byte arrays as the reference stores them (reversed). Returns its return code. */
int btle_ref_parse_connect_req(const uint8_t* payload, int n, uint32_t* out, uint8_t* init_a, uint8_t* adv_a, uint8_t* chm) {
    ADV_PDU_PAYLOAD_TYPE_5 p;
    memset(&p, 0, sizeof p);
    int rc = parse_adv_pdu_payload_byte((uint8_t*)payload, n, CONNECT_REQ, (void*)&p);
    if (rc != 0) return rc;
    out[0] = receiver_status.access_addr; out[1] = p.CRCInit; out[2] = p.WinSize; out[3] = p.WinOffset; out[4] = p.Interval;
    out[5] = p.Latency; out[6] = p.Timeout; out[7] = p.Hop; out[8] = p.SCA; out[9] = chm_is_full_map(receiver_status.chm) ? 1u : 0u;
    out[10] = (uint32_t)receiver_status.hop; out[11] = (uint32_t)receiver_status.interval;
    memcpy(init_a, p.InitA, 6); memcpy(adv_a, p.AdvA, 6); memcpy(chm, p.ChM, 5);
    return 0;
}
