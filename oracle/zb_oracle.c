/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into, imported by or
 * executed from the product path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * CPU restatement of Snout's IEEE 802.15.4 receive chain
 *   snout/modulations/Zigbee/hackrf/Zigbee_rx/top_block.py:52-89
 *     analog.quadrature_demod_cf(1)                              (:73)
 *     x - filter.single_pole_iir_filter_ff(0.00016)(x)           (:52,:70,:84-89)
 *     digital.clock_recovery_mm_ff(2, .000225, .5, .03, .0002)   (:69)
 *     ieee802_15_4.packet_sink(10)                               (:67)
 * and of the packet sink whose in-tree statement is
 *   scapy-radio/gnuradio/gr-zigbee/lib/packet_sink_scapy_impl.cc:55-374.
 *
 * PARITY STATUS.  The sink part is pinned: tests compare zb_sink_* below, chip
 * for chip and frame for frame, with the UNMODIFIED reference sink compiled
 * into oracle/_ref/libzbsink_ref.so.  The three GNU Radio 3.7.13.5 stream
 * blocks are NOT in the reference tree (un-vendored PyBOMBS dependency,
 * Makefile:40-41) and the reference holds no test or golden vector for them:
 * for that part this file restates the published block algorithms and says
 * "parity unpinned" (DESIGN.md section Oracle).  All float arithmetic here is
 * single operations in a fixed order (compile with -ffp-contract=off); the CUDA
 * kernels perform the same operations with __fmul_rn/__fadd_rn, so GPU and
 * oracle agree bit for bit.
 *
 * Deliberate restatement choices, part of the parity contract:
 *  - the single-pole DC tracker y[n] = a f[n] + (1-a) y[n-1] (double state) is
 *    evaluated in blocks of SNRX_IIR_BLOCK = 2048 samples, each started from a
 *    carried state folded from the 48 preceding blocks (memory cut after 98304
 *    samples = 15.7 time constants, residual weight e^-15.7 = 1.5e-7: the result
 *    equals the serial recurrence zb_oracle_dc_remove_serial to within a float
 *    ulp, bounded in tests/test_oracle_zigbee.py).  Blocks run in parallel on the
 *    GPU, and the output does not depend on where a time shard of a capture starts.
 *  - the clock recovery + packet sink run as independent chains, one per segment
 *    of the stream (zb_chains below); zb_oracle_receive_serial is the one
 *    unsegmented chain the reference flowgraph is, and the tests state how far
 *    the two can differ (not at all at high SNR; at marginal SNR by as much as
 *    the serial chain differs from itself when the capture starts one sample later).
 *  - the 8-tap interpolator dot product is summed as a balanced tree.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../include/snoutrx.h"
#include "zb_tables.h"

/* ---------------------------------------------------------------- front end */

/* table atan2 in the style of GNU Radio's fast_atan2f: octant reduction, 255-step
 * arctangent table with linear interpolation. */
static float tab_atan2f(float y, float x) {
    float ya = fabsf(y), xa = fabsf(x);
    if (!(ya > 0.0f || xa > 0.0f)) return 0.0f;
    float z = (ya < xa) ? ya / xa : xa / ya;
    float base;
    if (z < 0.003921569f) {
        base = z;
    } else {
        float alpha = z * 255.0f;
        int idx = ((int)alpha) & 0xFF;
        alpha = alpha - (float)idx;
        float lo = SNRX_ATAN_TAB[idx];
        float d = SNRX_ATAN_TAB[idx + 1] - lo;
        base = lo + d * alpha;
    }
    float ang;
    if (xa > ya) {
        if (x >= 0.0f) ang = (y >= 0.0f) ? base : -base;
        else ang = (y >= 0.0f) ? 3.14159265358979323846f - base : base - 3.14159265358979323846f;
    } else {
        if (y >= 0.0f) ang = (x >= 0.0f) ? 1.57079632679489661923f - base : 1.57079632679489661923f + base;
        else ang = (x >= 0.0f) ? -1.57079632679489661923f + base : -1.57079632679489661923f - base;
    }
    return ang;
}

/* f[n] = arg(x[n] * conj(x[n-1])), x[-1] = 0 (history of 2, gain 1). */
void zb_oracle_quad_demod(const float* iq, int64_t n, float* f) {
    float pr = 0.0f, pi = 0.0f;
    for (int64_t k = 0; k < n; k++) {
        float xr = iq[2 * k], xi = iq[2 * k + 1];
        float re = xr * pr + xi * pi;
        float im = xi * pr - xr * pi;
        f[k] = tab_atan2f(im, re);
        pr = xr; pi = xi;
    }
}

/* z = f - (float) y.  |z| <= pi + |DC| for any finite input; anything else (Inf / NaN samples in the capture) is replaced
 * by 0 so that the clock recovery state stays finite (same guard in csrc/zb.cuh zb_dc_out). */
static inline float dc_out(float f, double y) {
    float z = f - (float)y;
    return (fabsf(z) <= 16.0f) ? z : 0.0f;
}

/* The published block, stated serially (single_pole_iir<float,float,double>, then sub_ff):
 *   y[n] = a f[n] + (1-a) y[n-1]  (double state, y[-1] = 0),  z[n] = f[n] - (float) y[n].
 * This is what the reference flowgraph computes (top_block.py:52,70,84-89); the tests bound the blocked
 * evaluation below against it. */
void zb_oracle_dc_remove_serial(const float* f, int64_t n, float* z) {
    const double a = SNRX_IIR_ALPHA, b = SNRX_IIR_BETA;
    double y = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double t1 = a * (double)f[i];
        double t2 = b * y;
        y = t1 + t2;
        z[i] = dc_out(f[i], y);
    }
}

/* The same recurrence evaluated in blocks of SNRX_IIR_BLOCK samples on the absolute grid (what the GPU runs):
 *   e_b      = the recurrence over block b started from 0 (block-local end value),
 *   carry_b  = fold of the SNRX_IIR_MEMORY_BLOCKS preceding e_j, oldest first: c = e_j + (1-a)^BLOCK c,
 *   y        = the recurrence over block b started from carry_b,   z = f - (float) y.
 * With unlimited memory carry_b would be the serial state at the block start (up to double rounding); the memory
 * is cut after 48 blocks = 98304 samples = 15.7 time constants, where the forgotten part weighs e^-15.7 = 1.5e-7
 * of the DC -- so z equals the serial z to within a float ulp (tests/test_oracle_zigbee.py states the bound), and
 * z[n] depends on f[BLOCK*(b-48) .. n] only: any buffer that starts on the absolute block grid reproduces it
 * exactly once 48 blocks have gone by (time shards, include/snoutrx.h). */
void zb_oracle_dc_remove(const float* f, int64_t n, float* z) {
    const double a = SNRX_IIR_ALPHA, b = SNRX_IIR_BETA;
    double decay = 1.0;
    for (int i = 0; i < SNRX_IIR_BLOCK; i++) decay = decay * b;
    double ends[SNRX_IIR_MEMORY_BLOCKS];      /* block-local end values of the preceding blocks, oldest first */
    int n_ends = 0;
    for (int64_t n0 = 0; n0 < n; n0 += SNRX_IIR_BLOCK) {
        int64_t len = (n - n0 < SNRX_IIR_BLOCK) ? n - n0 : SNRX_IIR_BLOCK;
        double carry = 0.0;
        for (int j = 0; j < n_ends; j++) {
            double t = decay * carry;
            carry = ends[j] + t;
        }
        double l = 0.0, y = carry;
        for (int64_t i = 0; i < len; i++) {
            double t1 = a * (double)f[n0 + i];
            double t2 = b * l;
            l = t1 + t2;
            double t3 = b * y;
            y = t1 + t3;
            z[n0 + i] = dc_out(f[n0 + i], y);
        }
        if (n_ends == SNRX_IIR_MEMORY_BLOCKS) {
            for (int j = 1; j < n_ends; j++) ends[j - 1] = ends[j];
            n_ends--;
        }
        ends[n_ends++] = l;
    }
}

/* ----------------------------------------------------------------- the sink */

static const uint32_t* chip_map(void) {
    /* discriminator-domain chip words, derived from the 802.15.4 PN sequences:
     * bit (31-k) = c[k]^c[k-1]^(k&1); equals CHIP_MAPPING[] & 0x7FFFFFFE of
     * packet_sink_scapy_impl.h:28-45 (tests/test_oracle_zigbee.py checks). */
    static uint32_t map[16];
    static int ready = 0;
    if (!ready) {
        const char* pn0 = "11011001110000110101001000101110";
        for (int s = 0; s < 16; s++) {
            int c[32];
            for (int k = 0; k < 32; k++) {
                int src = (k - 4 * (s & 7) + 64) % 32;
                c[k] = pn0[src] - '0';
                if ((s & 8) && (k & 1)) c[k] ^= 1;
            }
            uint32_t v = 0;
            for (int k = 1; k < 32; k++) v |= (uint32_t)(c[k] ^ c[k - 1] ^ (k & 1)) << (31 - k);
            map[s] = v & 0x7FFFFFFEu;
        }
        ready = 1;
    }
    return map;
}

static inline int dist(uint32_t reg, uint32_t word) {
    return __builtin_popcount((reg & 0x7FFFFFFEu) ^ word);
}

typedef struct {
    int state;                 /* 0 search, 1 have sync (PHR), 2 have header (PSDU) */
    uint32_t reg;
    int preamble_cnt, chip_cnt;
    int byte, nibble_idx;      /* d_packet_byte, d_packet_byte_index */
    int frame_len, got;        /* d_packetlen, d_payload_cnt */
    unsigned lqi_sum, lqi_n;
    int threshold;
    uint8_t psdu[128];
    int64_t sync_chip;         /* index of the chip that completed the SFD */
    int64_t sync_pos;          /* its input position */
} zb_sink_t;

static void sink_search(zb_sink_t* s) {        /* enter_search, packet_sink_scapy_impl.cc:55-65 */
    s->state = 0; s->reg = 0; s->preamble_cnt = 0; s->chip_cnt = 0; s->byte = 0;
}

void zb_sink_init(zb_sink_t* s, int threshold) {
    memset(s, 0, sizeof(*s));
    s->threshold = threshold;
    sink_search(s);
}

static int sink_decode(zb_sink_t* s) {         /* decode_chips, packet_sink_scapy_impl.cc:95-125 */
    const uint32_t* map = chip_map();
    int best = 0xFF, best_d = 33;
    for (int i = 0; i < 16; i++) {
        int d = dist(s->reg, map[i]);
        if (d < best_d) { best = i; best_d = d; }
    }
    if (best_d < s->threshold) {
        if (s->lqi_n < 8) { s->lqi_sum += 32 - best_d; s->lqi_n++; }
        return best & 0xF;
    }
    return 0xFF;
}

/* Push one hard chip.  Returns 1 when a frame was completed (s->psdu, s->got). */
int zb_sink_push(zb_sink_t* s, int chip, int64_t chip_index, int64_t pos) {
    const uint32_t* map = chip_map();
    s->reg = (s->reg << 1) | (uint32_t)(chip & 1);
    if (s->state == 0) {                                    /* STATE_SYNC_SEARCH, :176-245 */
        if (s->preamble_cnt > 0) s->chip_cnt++;
        if (s->preamble_cnt == 0) {
            if (dist(s->reg, map[0]) < s->threshold) s->preamble_cnt = 1;
        } else if (s->chip_cnt == 32) {
            s->chip_cnt = 0;
            if (s->byte == 0) {
                if (dist(s->reg, map[0]) <= s->threshold) s->preamble_cnt++;
                else if (dist(s->reg, map[7]) <= s->threshold) s->byte = 7 << 4;
                else sink_search(s);
            } else {
                if (dist(s->reg, map[10]) <= s->threshold) {
                    s->state = 1;                           /* enter_have_sync, :67-79 */
                    s->got = 0; s->byte = 0; s->nibble_idx = 0;
                    s->lqi_sum = 0; s->lqi_n = 0;
                    s->sync_chip = chip_index; s->sync_pos = pos;
                } else sink_search(s);
            }
        }
        return 0;
    }
    if (s->state == 1) {                                    /* STATE_HAVE_SYNC, :247-291 */
        s->chip_cnt++;
        if (s->chip_cnt != 32) return 0;
        s->chip_cnt = 0;
        int c = sink_decode(s);
        if (c == 0xFF) { sink_search(s); return 0; }
        if (s->nibble_idx == 0) s->byte = c; else s->byte |= c << 4;
        s->nibble_idx++;
        if (s->nibble_idx % 2 == 0) {
            if (s->byte <= 127) {                           /* enter_have_header, :81-92 */
                s->state = 2; s->frame_len = s->byte; s->got = 0; s->byte = 0; s->nibble_idx = 0;
            } else sink_search(s);
        }
        return 0;
    }
    /* STATE_HAVE_HEADER, :293-359 */
    s->chip_cnt = (s->chip_cnt + 1) % 32;
    if (s->chip_cnt != 0) return 0;
    int c = sink_decode(s);
    if (c == 0xFF) { sink_search(s); return 0; }
    if (s->nibble_idx == 0) s->byte = c; else s->byte |= c << 4;
    s->nibble_idx++;
    if (s->nibble_idx % 2 != 0) return 0;
    s->psdu[s->got++] = (uint8_t)s->byte;
    s->nibble_idx = 0;
    if (s->got >= s->frame_len) { sink_search(s); return 1; }
    return 0;
}

static unsigned sink_lqi(const zb_sink_t* s) {              /* :334-335 */
    unsigned scaled = (s->lqi_sum / 8) << 3;
    return scaled >= 256 ? 255 : scaled;
}

static uint16_t fcs16(const uint8_t* d, int n) {            /* Dot15d4FCS.compute_fcs, dot15d4.py:151-164 */
    uint16_t crc = 0;
    for (int i = 0; i < n; i++) {
        unsigned c = d[i];
        unsigned q = (crc ^ c) & 15;
        crc = (uint16_t)((crc >> 4) ^ (q * 4225));
        q = (crc ^ (c >> 4)) & 15;
        crc = (uint16_t)((crc >> 4) ^ (q * 4225));
    }
    return crc;
}
uint16_t zb_oracle_fcs16(const uint8_t* d, int n) { return fcs16(d, n); }
uint32_t zb_oracle_chip_word(int s) { return chip_map()[s & 15]; }
int zb_oracle_sink_size(void) { return (int)sizeof(zb_sink_t); }
const uint8_t* zb_sink_psdu(const zb_sink_t* s) { return s->psdu; }
int zb_sink_len(const zb_sink_t* s) { return s->got; }

/* ------------------------------------------------- clock recovery + one chain */

typedef struct { float mu, omega, last; int64_t ii; } zb_mm_t;

static inline float mm_interp(const float* in, float mu) {
    int imu = (int)rintf(mu * (float)SNRX_MMSE_NSTEPS);
    const float* t = SNRX_MMSE_TAPS[imu];
    float p0 = t[0] * in[0], p1 = t[1] * in[1], p2 = t[2] * in[2], p3 = t[3] * in[3];
    float p4 = t[4] * in[4], p5 = t[5] * in[5], p6 = t[6] * in[6], p7 = t[7] * in[7];
    float s01 = p0 + p1, s23 = p2 + p3, s45 = p4 + p5, s67 = p6 + p7;
    float a = s01 + s23, b = s45 + s67;
    return a + b;
}

/* One Mueller-Mueller step at position st->ii; returns the soft chip. */
static inline float mm_step(zb_mm_t* st, const float* z) {
    const float omega_mid = 2.0f, gain_omega = 0.000225f, gain_mu = 0.03f;
    const float omega_lim = 2.0f * 0.0002f;
    float out = mm_interp(z + st->ii, st->mu);
    float sl = (st->last < 0.0f) ? -1.0f : 1.0f;
    float so = (out < 0.0f) ? -1.0f : 1.0f;
    float t1 = sl * out, t2 = so * st->last;
    float mm = t1 - t2;
    st->last = out;
    float om = st->omega + gain_omega * mm;
    float dv = om - omega_mid;
    float hi = fabsf(dv + omega_lim), lo = fabsf(dv - omega_lim);
    float cl = 0.5f * (hi - lo);                            /* branchless clip */
    st->omega = omega_mid + cl;
    float g = gain_mu * mm;
    float m1 = st->mu + st->omega;
    float m2 = m1 + g;
    float fl = floorf(m2);
    /* the advance is 1, 2 or 3 samples for every finite input (m2 lies in (1.6, 3.4)); the clamp only guards the loop */
    int adv = (int)fl;
    st->ii += adv < 1 ? 1 : adv > 3 ? 3 : adv;
    st->mu = m2 - fl;
    return out;
}

/* Run one chain over z[begin, end): fresh clock recovery at `begin`; the sink stays in its initial (search,
 * empty register) state until the chain reaches position `hold` (0 = from the start) -- the state the
 * sequential sink is in right after it finished a frame there.  Frames whose SFD-completing chip lies at a
 * position in [body_lo, body_hi) are kept.  The chain stops at `end`, or once it has passed its body with
 * the sink searching (a later sync belongs to the next chain).  *stop_out = max(position where it stopped,
 * hold): up to there the receiver was busy with a frame of this chain (or an earlier one).
 * chips_out (optional) receives up to chips_cap soft chips, chip_pos_out their positions.
 * Returns the number of chips (n_frames in/out counts, stores at most cap). */
int64_t zb_oracle_chain_hold(const float* z, int64_t begin, int64_t end, int64_t body_lo, int64_t body_hi,
                             int64_t hold, int threshold, int channel, uint32_t segment,
                             snrx_frame_t* out, int cap, int* n_frames,
                             float* chips_out, int64_t* chip_pos_out, int64_t chips_cap, int64_t* stop_out) {
    zb_mm_t mm = {0.5f, 2.0f, 0.0f, begin};
    zb_sink_t sink;
    zb_sink_init(&sink, threshold);
    int64_t nchips = 0;
    while (mm.ii + 8 <= end && !(mm.ii >= body_hi && sink.state == 0)) {
        int64_t pos = mm.ii;
        float soft = mm_step(&mm, z);
        if (chips_out && nchips < chips_cap) { chips_out[nchips] = soft; if (chip_pos_out) chip_pos_out[nchips] = pos; }
        if (pos >= hold && zb_sink_push(&sink, soft > 0.0f, nchips, pos)) {
            if (sink.sync_pos >= body_lo && sink.sync_pos < body_hi) {
                if (*n_frames < cap) {
                    snrx_frame_t* f = &out[*n_frames];
                    memset(f, 0, sizeof(*f));
                    f->sample_index = sink.sync_pos;
                    f->window = segment;
                    f->channel = (uint16_t)channel;
                    f->proto = SNRX_PROTO_ZIGBEE;
                    f->len = (uint16_t)sink.got;
                    f->lqi = (uint8_t)sink_lqi(&sink);
                    memcpy(f->bytes, sink.psdu, (size_t)sink.got);
                    if (sink.got >= 2) {
                        uint16_t c = fcs16(sink.psdu, sink.got - 2);
                        f->crc_ok = (uint8_t)(c == (uint16_t)(sink.psdu[sink.got - 2] | (sink.psdu[sink.got - 1] << 8)));
                    }
                }
                (*n_frames)++;
            }
        }
        nchips++;
    }
    if (stop_out) *stop_out = mm.ii > hold ? mm.ii : hold;
    return nchips;
}

int64_t zb_oracle_chain(const float* z, int64_t begin, int64_t end, int64_t body_lo, int64_t body_hi,
                        int threshold, int channel, uint32_t segment,
                        snrx_frame_t* out, int cap, int* n_frames,
                        float* chips_out, int64_t* chip_pos_out, int64_t chips_cap) {
    return zb_oracle_chain_hold(z, begin, end, body_lo, body_hi, 0, threshold, channel, segment, out, cap, n_frames,
                                chips_out, chip_pos_out, chips_cap, NULL);
}

#define ZB_POST_HALO 16448      /* PHR + 127 bytes = 256 symbols * 64 samples + 64 */
/* While the reference's sequential sink decodes a frame it cannot lock onto anything else.  A chain that starts inside a
 * foreign frame can (payload chips that look like 0,7,A), so a CRC-failed record whose sync lies inside the span of an
 * earlier CRC-ok record of the same stream -- PHR (2 symbols) + len bytes (2 symbols each) of 64 samples after the
 * SFD-completing chip -- is an artefact of restarting the sink per segment and is not reported.  Records of one stream
 * in position order, compacted in place; returns the number kept. */
int zb_span_filter(snrx_frame_t* fr, int n) {
    int64_t good_end = 0;
    int w = 0;
    for (int k = 0; k < n; k++) {
        int keep = 1;
        if (fr[k].crc_ok) {
            int64_t e = fr[k].sample_index + (int64_t)(2 + 2 * fr[k].len) * 64;
            if (e > good_end) good_end = e;
        } else if (fr[k].sample_index < good_end) keep = 0;
        if (keep) { if (w != k) fr[w] = fr[k]; w++; }
    }
    return w;
}

#define ZB_SINK_LEAD 1024       /* the sink starts this many samples before the body: SHR (640) + alignment slack */

/* The chains of one stream.  Chain k covers the body [k*segment, (k+1)*segment): its clock recovery starts
 * `prehalo` samples early (warm-up), its sink ZB_SINK_LEAD samples early (long enough to catch a preamble
 * that began before the body, short enough that a chain starting inside a foreign frame rarely locks onto
 * payload chips before its body), and it may run ZB_POST_HALO samples past its body to finish a frame.
 * Chains are independent, so a time shard reproduces the whole-capture result bit for bit.  Frames are
 * appended in (segment, position) order. */
static int zb_chains(const float* z, int64_t n, int channel, int threshold, int64_t segment, int64_t prehalo,
                     snrx_frame_t* out, int cap) {
    int nf = 0;
    uint32_t seg = 0;
    for (int64_t lo = 0; lo < n; lo += segment, seg++) {
        const int64_t hi = lo + segment < n ? lo + segment : n;
        const int64_t begin = lo - prehalo > 0 ? lo - prehalo : 0;
        const int64_t end = hi + ZB_POST_HALO < n ? hi + ZB_POST_HALO : n;
        zb_oracle_chain_hold(z, begin, end, lo, hi, lo - ZB_SINK_LEAD, threshold, channel, seg, out, cap, &nf, NULL, NULL, 0, NULL);
    }
    if (out && nf <= cap) nf = zb_span_filter(out, nf);
    return nf;
}

/* Whole receive chain over one buffer of channel-rate cf32: DC tracker fresh at sample 0, then the chains. */
int zb_oracle_receive(const float* iq, int64_t n, int channel, int threshold,
                      int64_t segment, int64_t prehalo, snrx_frame_t* out, int cap) {
    float* f = (float*)malloc(sizeof(float) * (size_t)(n + 8));
    float* z = (float*)malloc(sizeof(float) * (size_t)(n + 8));
    if (!f || !z) { free(f); free(z); return -1; }
    zb_oracle_quad_demod(iq, n, f);
    zb_oracle_dc_remove(f, n, z);
    int nf = zb_chains(z, n, channel, threshold, segment, prehalo, out, cap);
    free(f); free(z);
    return nf;
}

/* The reference flowgraph as it is: serial DC tracker, ONE clock-recovery + sink chain over the whole stream
 * (fresh at sample 0), no span rule needed.  The yardstick the segmented receiver is measured against. */
int zb_oracle_receive_serial(const float* iq, int64_t n, int channel, int threshold, snrx_frame_t* out, int cap) {
    float* f = (float*)malloc(sizeof(float) * (size_t)(n + 8));
    float* z = (float*)malloc(sizeof(float) * (size_t)(n + 8));
    if (!f || !z) { free(f); free(z); return -1; }
    zb_oracle_quad_demod(iq, n, f);
    zb_oracle_dc_remove_serial(f, n, z);
    int nf = 0;
    zb_oracle_chain_hold(z, 0, n, 0, n, 0, threshold, channel, 0, out, cap, &nf, NULL, NULL, 0, NULL);
    free(f); free(z);
    return nf;
}

/* same, on an already demodulated + DC-removed stream (used on GPU-produced streams) */
int zb_oracle_receive_z(const float* z, int64_t n, int channel, int threshold,
                        int64_t segment, int64_t prehalo, snrx_frame_t* out, int cap) {
    return zb_chains(z, n, channel, threshold, segment, prehalo, out, cap);
}

double zb_oracle_time(const float* iq, int64_t n, int channel, int64_t segment, int64_t prehalo,
                      int reps, int* frames_out) {
    struct timespec t0, t1;
    int nf = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; r++) nf = zb_oracle_receive(iq, n, channel, 10, segment, prehalo, NULL, 0);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (frames_out) *frames_out = nf;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
