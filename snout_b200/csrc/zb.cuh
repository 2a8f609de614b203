// zb.cuh -- IEEE 802.15.4 (Zigbee) receive chain on the GPU.
//
// Reference chain (snout/modulations/Zigbee/hackrf/Zigbee_rx/top_block.py:52-89):
//   analog.quadrature_demod_cf(1) -> x - single_pole_iir_filter_ff(0.00016)(x)
//   -> digital.clock_recovery_mm_ff(2, 0.000225, 0.5, 0.03, 0.0002) -> ieee802_15_4.packet_sink(10)
// with the sink's in-tree statement scapy-radio/gnuradio/gr-zigbee/lib/packet_sink_scapy_impl.cc
// (enter_search 55-65, enter_have_sync 67-79, enter_have_header 81-92, decode_chips 95-125,
// general_work 158-374).  The three GNU Radio stream blocks are not vendored in the reference
// tree; their published algorithms are restated here (and, independently, in oracle/zb_oracle.c,
// which the kernels must match bit for bit: every float operation below is an explicitly rounded
// single operation in a fixed order).
//
// Parallel decomposition (DESIGN.md "Zigbee"):
//   k_zb_quad      one thread per sample: table atan2 of x[n] conj(x[n-1])
//   k_zb_iir_*     the single-pole DC tracker evaluated in blocks of 4096 samples (block-local
//                  recurrence + a carried state folded from the 8 preceding blocks), so blocks run
//                  in parallel and the result does not depend on where a time shard starts
//   k_zb_chain     one thread per (capture, channel, segment): Mueller-Mueller clock recovery feeding
//                  the packet-sink state machine; segments overlap by a warm-up pre-halo and a
//                  longest-frame post-halo, a frame belongs to the segment holding its SFD position
//   k_zb_gather    compaction of the per-chain frame slots into the batch frame list
#pragma once
#include <string>

#include "common.cuh"
#include "scan.cuh"
#define SNRX_TABLE_QUAL static const
#include "zb_tables.h"

namespace snrx {

constexpr int kZbPostHalo = 16448;          // PHR + 127 bytes = 256 symbols * 64 samples, + 64
constexpr int kZbMinFrameSamples = 896;     // 10 SHR + 2 PHR + 2 PSDU symbols of 64 samples
constexpr int kZbSinkLead = 1024;           // the sink starts this many samples before the body: SHR (640) + alignment
                                            // slack; the clock recovery starts `prehalo` samples early (warm-up only)

SNRX_HD float tab_atan2(float y, float x, const float* tab /*[257]*/) {
    const float ya = fabsf(y), xa = fabsf(x);
    if (!(ya > 0.0f || xa > 0.0f)) return 0.0f;
    const float z = (ya < xa) ? f_div(ya, xa) : f_div(xa, ya);
    float base;
    if (z < 0.003921569f) {
        base = z;
    } else {
        float alpha = f_mul(z, 255.0f);
        const int idx = ((int)alpha) & 0xFF;
        alpha = f_sub(alpha, (float)idx);
        const float lo = tab[idx];
        const float d = f_sub(tab[idx + 1], lo);
        base = f_add(lo, f_mul(d, alpha));
    }
    const float pi = 3.14159265358979323846f, hpi = 1.57079632679489661923f;
    float ang;
    if (xa > ya) {
        if (x >= 0.0f) ang = (y >= 0.0f) ? base : -base;
        else ang = (y >= 0.0f) ? f_sub(pi, base) : f_sub(base, pi);
    } else {
        if (y >= 0.0f) ang = (x >= 0.0f) ? f_sub(hpi, base) : f_add(hpi, base);
        else ang = (x >= 0.0f) ? f_add(-hpi, base) : f_sub(-hpi, base);
    }
    return ang;
}

// f = arg(x * conj(p))
SNRX_HD float quad_demod(float xr, float xi, float pr, float pi, const float* tab) {
    const float re = f_add(f_mul(xr, pr), f_mul(xi, pi));
    const float im = f_sub(f_mul(xi, pr), f_mul(xr, pi));
    return tab_atan2(im, re, tab);
}

// discriminator-domain chip words of the 16 data symbols (derived from the 802.15.4 PN
// sequences; equal CHIP_MAPPING[] & 0x7FFFFFFE of packet_sink_scapy_impl.h:28-45)
struct ChipMap { uint32_t w[16]; };
inline ChipMap make_chip_map() {
    ChipMap m;
    const char* pn0 = "11011001110000110101001000101110";
    for (int s = 0; s < 16; s++) {
        int c[32];
        for (int k = 0; k < 32; k++) {
            c[k] = pn0[(k - 4 * (s & 7) + 64) % 32] - '0';
            if ((s & 8) && (k & 1)) c[k] ^= 1;
        }
        uint32_t v = 0;
        for (int k = 1; k < 32; k++) v |= (uint32_t)(c[k] ^ c[k - 1] ^ (k & 1)) << (31 - k);
        m.w[s] = v & 0x7FFFFFFEu;
    }
    return m;
}

SNRX_HD int popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

struct ZbSink {
    int state;                 // 0 search, 1 have sync (PHR), 2 have header (PSDU)
    uint32_t reg;
    int preamble_cnt, chip_cnt;
    int byte, nibble_idx;
    int frame_len, got;
    unsigned lqi_sum, lqi_n;
    int32_t sync_pos;
};

SNRX_HD void zb_sink_search(ZbSink& s) { s.state = 0; s.reg = 0; s.preamble_cnt = 0; s.chip_cnt = 0; s.byte = 0; }
SNRX_HD void zb_sink_init(ZbSink& s) { s.got = 0; s.frame_len = 0; s.nibble_idx = 0; s.lqi_sum = 0; s.lqi_n = 0; s.sync_pos = 0; zb_sink_search(s); }

SNRX_HD int zb_dist(uint32_t reg, uint32_t word) { return popc32((reg & 0x7FFFFFFEu) ^ word); }

SNRX_HD int zb_decode_symbol(ZbSink& s, const uint32_t* map, int threshold) {     // decode_chips :95-125
    int best = 0xFF, best_d = 33;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int d = zb_dist(s.reg, map[i]);
        if (d < best_d) { best = i; best_d = d; }
    }
    if (best_d < threshold) {
        if (s.lqi_n < 8) { s.lqi_sum += 32 - best_d; s.lqi_n++; }
        return best & 0xF;
    }
    return 0xFF;
}

// push one hard chip; returns 1 when a frame is complete (psdu[0..got))
SNRX_HD int zb_sink_push(ZbSink& s, int chip, int32_t pos, const uint32_t* map, int threshold, uint8_t* psdu) {
    s.reg = (s.reg << 1) | (uint32_t)(chip & 1);
    if (s.state == 0) {                                           // STATE_SYNC_SEARCH :176-245
        if (s.preamble_cnt > 0) s.chip_cnt++;
        if (s.preamble_cnt == 0) {
            if (zb_dist(s.reg, map[0]) < threshold) s.preamble_cnt = 1;
        } else if (s.chip_cnt == 32) {
            s.chip_cnt = 0;
            if (s.byte == 0) {
                if (zb_dist(s.reg, map[0]) <= threshold) s.preamble_cnt++;
                else if (zb_dist(s.reg, map[7]) <= threshold) s.byte = 7 << 4;
                else zb_sink_search(s);
            } else {
                if (zb_dist(s.reg, map[10]) <= threshold) {       // enter_have_sync :67-79
                    s.state = 1; s.got = 0; s.byte = 0; s.nibble_idx = 0; s.lqi_sum = 0; s.lqi_n = 0;
                    s.sync_pos = pos;
                } else zb_sink_search(s);
            }
        }
        return 0;
    }
    if (s.state == 1) {                                           // STATE_HAVE_SYNC :247-291
        s.chip_cnt++;
        if (s.chip_cnt != 32) return 0;
        s.chip_cnt = 0;
        const int c = zb_decode_symbol(s, map, threshold);
        if (c == 0xFF) { zb_sink_search(s); return 0; }
        if (s.nibble_idx == 0) s.byte = c; else s.byte |= c << 4;
        s.nibble_idx++;
        if (s.nibble_idx % 2 == 0) {
            if (s.byte <= 127) { s.state = 2; s.frame_len = s.byte; s.got = 0; s.byte = 0; s.nibble_idx = 0; }   // :81-92
            else zb_sink_search(s);
        }
        return 0;
    }
    s.chip_cnt = (s.chip_cnt + 1) % 32;                           // STATE_HAVE_HEADER :293-359
    if (s.chip_cnt != 0) return 0;
    const int c = zb_decode_symbol(s, map, threshold);
    if (c == 0xFF) { zb_sink_search(s); return 0; }
    if (s.nibble_idx == 0) s.byte = c; else s.byte |= c << 4;
    s.nibble_idx++;
    if (s.nibble_idx % 2 != 0) return 0;
    psdu[s.got++] = (uint8_t)s.byte;
    s.nibble_idx = 0;
    if (s.got >= s.frame_len) { zb_sink_search(s); return 1; }
    return 0;
}

SNRX_HD uint16_t zb_fcs16(const uint8_t* d, int n) {              // Dot15d4FCS.compute_fcs, dot15d4.py:151-164
    uint32_t crc = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t c = d[i];
        uint32_t q = (crc ^ c) & 15u;
        crc = ((crc >> 4) ^ (q * 4225u)) & 0xFFFFu;
        q = (crc ^ (c >> 4)) & 15u;
        crc = ((crc >> 4) ^ (q * 4225u)) & 0xFFFFu;
    }
    return (uint16_t)crc;
}

// state carried into block b of the DC tracker: fold of the (at most) SNRX_IIR_MEMORY_BLOCKS preceding
// block-local end values, oldest first; decay = (1-alpha)^4096
SNRX_HD double zb_iir_fold(const double* ends, int b, double decay) {
    double carry = 0.0;
    for (int j = (b > SNRX_IIR_MEMORY_BLOCKS ? b - SNRX_IIR_MEMORY_BLOCKS : 0); j < b; j++) carry = d_add(ends[j], d_mul(decay, carry));
    return carry;
}

struct ZbMm { float mu, omega, last; int32_t ii; };     // ii: position in the stream (n_out < 2^31)

// row of the MMSE interpolator table for the fractional delay mu: rint(mu * 128).  mu * 128 lies in [0, 128],
// so adding 1.5 * 2^23 rounds it to the nearest integer (ties to even, exactly rintf) in the low mantissa bits:
// one FADD instead of a conversion on the chain's critical path.
SNRX_HD int zb_mm_row(float mu) {
    const float t = f_add(f_mul(mu, (float)SNRX_MMSE_NSTEPS), 12582912.0f);
#ifdef __CUDA_ARCH__
    return __float_as_int(t) & 0x1FF;
#else
    uint32_t u; memcpy(&u, &t, 4); return (int)(u & 0x1FFu);
#endif
}

// one Mueller-Mueller step on the 8 samples in[0..8) = z[ii .. ii+8) with the interpolator row t = taps[zb_mm_row(mu)];
// returns the soft chip
SNRX_HD float zb_mm_step(ZbMm& st, const float (&in)[8], const float (&t)[8]) {
    const float omega_mid = 2.0f, gain_omega = 0.000225f, gain_mu = 0.03f;
    const float omega_lim = 2.0f * 0.0002f;
    const float p0 = f_mul(t[0], in[0]), p1 = f_mul(t[1], in[1]), p2 = f_mul(t[2], in[2]), p3 = f_mul(t[3], in[3]);
    const float p4 = f_mul(t[4], in[4]), p5 = f_mul(t[5], in[5]), p6 = f_mul(t[6], in[6]), p7 = f_mul(t[7], in[7]);
    const float s01 = f_add(p0, p1), s23 = f_add(p2, p3), s45 = f_add(p4, p5), s67 = f_add(p6, p7);
    const float out = f_add(f_add(s01, s23), f_add(s45, s67));
    const float sl = (st.last < 0.0f) ? -1.0f : 1.0f;
    const float so = (out < 0.0f) ? -1.0f : 1.0f;
    const float mm = f_sub(f_mul(sl, out), f_mul(so, st.last));
    st.last = out;
    const float om = f_add(st.omega, f_mul(gain_omega, mm));
    const float dv = f_sub(om, omega_mid);
    const float hi = fabsf(f_add(dv, omega_lim)), lo = fabsf(f_sub(dv, omega_lim));
    st.omega = f_add(omega_mid, f_mul(0.5f, f_sub(hi, lo)));
    const float g = f_mul(gain_mu, mm);
    const float m2 = f_add(f_add(st.mu, st.omega), g);
    const float fl = floorf(m2);
    st.ii += (int32_t)fl;
    st.mu = f_sub(m2, fl);
    return out;
}

// where a chain reads its samples from: straight from the stream (host stepping harness) ...
struct ZbDirectSrc {
    const float* z;
    const float* taps;     // [129][8]
    SNRX_HD void get8(int32_t ii, float (&in)[8]) const {
#pragma unroll
        for (int k = 0; k < 8; k++) in[k] = z[ii + k];
    }
    SNRX_HD void row(int r, float (&t)[8]) const {
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = taps[r * SNRX_MMSE_NTAPS + k];
    }
};

// Samples a frame occupies after its SFD-completing chip: PHR (2 symbols) + len bytes (2 symbols each), 64 samples per symbol.
// While the reference's sequential sink decodes such a frame it cannot lock onto anything else, so a CRC-failed record whose
// sync lies inside the span of an earlier CRC-ok record of the same stream is an artefact of restarting the sink per segment
// and is not reported (k_zb_span_filter here, zb_span_filter in oracle/zb_oracle.c, stream.zb_span_filter across shards).
SNRX_HD int64_t zb_frame_end(int64_t sample_index, int len) { return sample_index + (int64_t)(2 + 2 * len) * 64; }

struct ZbChainParams {
    int32_t n_out;            // channel-rate samples per capture in the buffer
    int32_t origin;           // local index where segment `first_segment` starts (pre halo length)
    int32_t body;             // body length (local samples from origin)
    int32_t segment, prehalo;
    int32_t n_segments;
    uint32_t first_segment, first_capture;
    uint32_t n_captures, n_channels;
    int32_t threshold;
    uint32_t slots_per_chain;
    size_t z_stride;          // floats between (capture, channel) streams
};

// One chain.  src: samples of this (capture, channel) stream.  Frames are written to slots[0..), returns count.
// The chain ends at the post halo, or as soon as it has passed its body with the sink back in the search
// state: a sync found from there on completes at a position >= hi and belongs to the next segment, so
// nothing this chain could still report is lost (oracle/zb_oracle.c zb_oracle_chain stops at the same step).
template <class Src>
SNRX_HD uint32_t zb_run_chain(Src& src, const ZbChainParams& p, int seg, const uint32_t* map,
                              int channel_number, uint32_t capture_id, snrx_frame_t* slots, float* chips_dbg,
                              int64_t chips_cap, int64_t* nchips_out, int64_t* good_end_out = nullptr) {
    int64_t good_end = 0;            // end (whole-capture index) of the last CRC-ok frame this chain reports
    // all positions are < n_out + segment + post halo < 2^31 (zb_create bounds max_out)
    const int32_t lo = p.origin + seg * p.segment;
    int32_t hi = lo + p.segment;
    const int32_t body_end = p.origin + p.body;
    if (hi > body_end) hi = body_end;
    int32_t begin = lo - p.prehalo; if (begin < 0) begin = 0;
    int32_t end = hi + kZbPostHalo; if (end > p.n_out) end = p.n_out;
    const int32_t sink_from = lo - kZbSinkLead;
    ZbMm mm; mm.mu = 0.5f; mm.omega = 2.0f; mm.last = 0.0f; mm.ii = begin;
    ZbSink sink; zb_sink_init(sink);
    uint8_t psdu[128];
    uint32_t nf = 0;
    int32_t nchips = 0;
    while (mm.ii + 8 <= end && !(mm.ii >= hi && sink.state == 0)) {
        const int32_t pos = mm.ii;
        float in[8], t[8];
        src.row(zb_mm_row(mm.mu), t);
        src.get8(pos, in);
        const float soft = zb_mm_step(mm, in, t);
        if (chips_dbg && nchips < chips_cap) chips_dbg[nchips] = soft;
        nchips++;
        if (pos >= sink_from && zb_sink_push(sink, soft > 0.0f, pos, map, p.threshold, psdu)) {
            if (sink.sync_pos >= lo && sink.sync_pos < hi) {
                if (nf < p.slots_per_chain) {
                    snrx_frame_t& f = slots[nf];
                    f.sample_index = (int64_t)(sink.sync_pos - p.origin) + (int64_t)p.first_segment * p.segment;
                    f.capture_id = capture_id;
                    f.window = p.first_segment + (uint32_t)seg;
                    f.channel = (uint16_t)channel_number;
                    f.proto = SNRX_PROTO_ZIGBEE;
                    unsigned scaled = (sink.lqi_sum / 8) << 3;                  // :334-335
                    f.lqi = (uint8_t)(scaled >= 256 ? 255 : scaled);
                    f.phase = 0;
                    f.len = (uint16_t)sink.got;
                    f.access_addr = 0;
                    f.crc_ok = 0;
                    if (sink.got >= 2) {
                        const uint16_t c = zb_fcs16(psdu, sink.got - 2);
                        f.crc_ok = (uint8_t)(c == (uint16_t)(psdu[sink.got - 2] | (psdu[sink.got - 1] << 8)));
                    }
                    for (int i = 0; i < 132; i++) f.bytes[i] = (i < sink.got) ? psdu[i] : 0;
                    if (f.crc_ok) { const int64_t e = zb_frame_end(f.sample_index, sink.got); if (e > good_end) good_end = e; }
                }
                nf++;
            }
        }
    }
    if (nchips_out) *nchips_out = nchips;
    if (good_end_out) *good_end_out = good_end;
    return nf;
}

// Filter of one chain's records given the ends of the CRC-ok frames of the `lookback` preceding chains of its stream.
// Records are compacted in place; returns the number kept.
SNRX_HD uint32_t zb_filter_chain(snrx_frame_t* slots, uint32_t n, int64_t good_end) {
    uint32_t w = 0;
    for (uint32_t k = 0; k < n; k++) {
        const snrx_frame_t& f = slots[k];
        bool keep = true;
        if (f.crc_ok) { const int64_t e = zb_frame_end(f.sample_index, f.len); if (e > good_end) good_end = e; }
        else if (f.sample_index < good_end) keep = false;
        if (keep) { if (w != k) slots[w] = slots[k]; w++; }
    }
    return w;
}
SNRX_HD int zb_filter_lookback(int segment) { return (16384 + segment - 1) / segment + 1; }

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------ kernels

struct ZbQuadArgs {
    const float2* x; uint64_t stride; int64_t n; uint32_t n_captures;
    float* f; size_t f_stride;            // [cap][1][f_stride]
    const float* atan_tab;
};

__global__ void __launch_bounds__(256) k_zb_quad(ZbQuadArgs a) {
    __shared__ float tab[257];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) tab[i] = a.atan_tab[i];
    __syncthreads();
    const uint64_t total = (uint64_t)a.n_captures * (uint64_t)a.n;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t cap = (uint32_t)(idx / (uint64_t)a.n);
        const int64_t n = (int64_t)(idx % (uint64_t)a.n);
        const float2* xc = a.x + (size_t)cap * a.stride;
        const float2 cur = __ldg(xc + n);
        float2 prev = make_float2(0.f, 0.f);
        if (n > 0) prev = __ldg(xc + n - 1);
        a.f[(size_t)cap * a.f_stride + n] = quad_demod(cur.x, cur.y, prev.x, prev.y, tab);
    }
}

struct ZbIirArgs {
    const float* f; float* z; size_t stride;      // per (capture, channel) stream stride
    int32_t n; int32_t n_blocks; uint32_t n_streams;
    double* block_end;    // [stream][n_blocks]  block-local recurrence value at the block end
    double* carry_in;     // [stream][n_blocks]  carried state entering each block
    const double* pw;     // [4096] (1-alpha)^(i+1)
};

// A (stream, block) unit is one thread's 4096-sample serial recurrence.  Its samples are contiguous and
// 128-byte aligned, so the thread streams them as float4 with the next 32 samples (8 loads) already in
// flight while the current 32 go through the dependent double-precision chain.
constexpr int kIirPf = 8;             // float4 loads in flight per thread

__global__ void __launch_bounds__(64) k_zb_iir_sum(ZbIirArgs a) {
    const uint32_t total = a.n_streams * (uint32_t)a.n_blocks;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    // consecutive threads take different streams (DRAM page spread, as in k_zb_chain)
    const uint32_t b = idx / a.n_streams, s = idx % a.n_streams;
    const float* f = a.f + (size_t)s * a.stride + (size_t)b * SNRX_IIR_BLOCK;
    const int len = min(SNRX_IIR_BLOCK, a.n - (int)b * SNRX_IIR_BLOCK);
    const int n4 = len >> 2;
    const float4* f4 = reinterpret_cast<const float4*>(f);
    double l = 0.0;
    float4 cur[kIirPf], nxt[kIirPf];
#pragma unroll
    for (int k = 0; k < kIirPf; k++) nxt[k] = (k < n4) ? __ldcs(f4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i4 = 0; i4 < n4; i4 += kIirPf) {
#pragma unroll
        for (int k = 0; k < kIirPf; k++) cur[k] = nxt[k];
#pragma unroll
        for (int k = 0; k < kIirPf; k++) nxt[k] = (i4 + kIirPf + k < n4) ? __ldcs(f4 + i4 + kIirPf + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < kIirPf; k++) {
            if (i4 + k < n4) {
                l = d_add(d_mul(SNRX_IIR_ALPHA, (double)cur[k].x), d_mul(SNRX_IIR_BETA, l));
                l = d_add(d_mul(SNRX_IIR_ALPHA, (double)cur[k].y), d_mul(SNRX_IIR_BETA, l));
                l = d_add(d_mul(SNRX_IIR_ALPHA, (double)cur[k].z), d_mul(SNRX_IIR_BETA, l));
                l = d_add(d_mul(SNRX_IIR_ALPHA, (double)cur[k].w), d_mul(SNRX_IIR_BETA, l));
            }
        }
    }
    for (int i = n4 << 2; i < len; i++) l = d_add(d_mul(SNRX_IIR_ALPHA, (double)f[i]), d_mul(SNRX_IIR_BETA, l));
    a.block_end[(size_t)s * a.n_blocks + b] = l;
}

// carried state entering block b: folded from the SNRX_IIR_MEMORY_BLOCKS preceding blocks (all full
// length), oldest first -- one thread per (stream, block), no serial pass over the stream
__global__ void __launch_bounds__(128) k_zb_iir_carry(ZbIirArgs a) {
    const uint32_t total = a.n_streams * (uint32_t)a.n_blocks;
    const double decay = a.pw[SNRX_IIR_BLOCK - 1];
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const uint32_t s = idx / (uint32_t)a.n_blocks;
        const int b = (int)(idx % (uint32_t)a.n_blocks);
        a.carry_in[idx] = zb_iir_fold(a.block_end + (size_t)s * a.n_blocks, b, decay);
    }
}

__global__ void __launch_bounds__(64) k_zb_dc(ZbIirArgs a) {
    __shared__ double pw_s[SNRX_IIR_BLOCK];                   // (1-alpha)^(i+1): same index for every thread -> broadcast
    for (int i = threadIdx.x; i < SNRX_IIR_BLOCK; i += blockDim.x) pw_s[i] = a.pw[i];
    __syncthreads();
    const uint32_t total = a.n_streams * (uint32_t)a.n_blocks;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const uint32_t b = idx / a.n_streams, s = idx % a.n_streams;
    const float* f = a.f + (size_t)s * a.stride + (size_t)b * SNRX_IIR_BLOCK;
    float* z = a.z + (size_t)s * a.stride + (size_t)b * SNRX_IIR_BLOCK;
    const int len = min(SNRX_IIR_BLOCK, a.n - (int)b * SNRX_IIR_BLOCK);
    const int n4 = len >> 2;
    const float4* f4 = reinterpret_cast<const float4*>(f);
    float4* z4 = reinterpret_cast<float4*>(z);
    const double carry = a.carry_in[(size_t)s * a.n_blocks + b];
    double l = 0.0;
    float4 cur[kIirPf], nxt[kIirPf];
#pragma unroll
    for (int k = 0; k < kIirPf; k++) nxt[k] = (k < n4) ? __ldcs(f4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i4 = 0; i4 < n4; i4 += kIirPf) {
#pragma unroll
        for (int k = 0; k < kIirPf; k++) cur[k] = nxt[k];
#pragma unroll
        for (int k = 0; k < kIirPf; k++) nxt[k] = (i4 + kIirPf + k < n4) ? __ldcs(f4 + i4 + kIirPf + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < kIirPf; k++) {
            if (i4 + k < n4) {
                const int i = (i4 + k) << 2;
                float4 o;
                l = d_add(d_mul(SNRX_IIR_ALPHA, (double)cur[k].x), d_mul(SNRX_IIR_BETA, l));
                o.x = f_sub(cur[k].x, (float)d_add(l, d_mul(pw_s[i], carry)));
                l = d_add(d_mul(SNRX_IIR_ALPHA, (double)cur[k].y), d_mul(SNRX_IIR_BETA, l));
                o.y = f_sub(cur[k].y, (float)d_add(l, d_mul(pw_s[i + 1], carry)));
                l = d_add(d_mul(SNRX_IIR_ALPHA, (double)cur[k].z), d_mul(SNRX_IIR_BETA, l));
                o.z = f_sub(cur[k].z, (float)d_add(l, d_mul(pw_s[i + 2], carry)));
                l = d_add(d_mul(SNRX_IIR_ALPHA, (double)cur[k].w), d_mul(SNRX_IIR_BETA, l));
                o.w = f_sub(cur[k].w, (float)d_add(l, d_mul(pw_s[i + 3], carry)));
                z4[i4 + k] = o;
            }
        }
    }
    for (int i = n4 << 2; i < len; i++) {
        const float fv = f[i];
        l = d_add(d_mul(SNRX_IIR_ALPHA, (double)fv), d_mul(SNRX_IIR_BETA, l));
        z[i] = f_sub(fv, (float)d_add(l, d_mul(pw_s[i], carry)));
    }
}

// ... or, on the device, through a per-thread ring in shared memory that cp.async keeps kZbAhead samples
// ahead of the chain.  The M&M loop is one long dependent chain per thread; with the samples already on
// chip its step costs shared-memory latency instead of an L2 / HBM round trip every few steps.
// Layout: ring row r of lane l at ring[r * 32 + l] -> every access of a warp is bank-conflict free whatever
// the lanes' positions; rows 0..7 are mirrored at kZbRing.. so the 8 taps never wrap.
constexpr int kZbRing = 128;          // samples held per chain (power of two)
constexpr int kZbChunk = 16;          // samples per cp.async group
constexpr int kZbAhead = 64;          // a group is waited for 3 groups (48 samples) after its issue
constexpr int kZbChainThreads = 64;

__device__ __forceinline__ void cp_async4(uint32_t smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

struct ZbRingSrc {
    const float* z;        // stream
    uint32_t ring;         // shared-memory byte address of this lane's column
    uint32_t taps;         // shared-memory byte address of the interpolator table (kept in a register: the
                           // compiler would otherwise rebuild it from SR_CgaCtaId inside the loop)
    int32_t begin;         // stream index of ring position 0
    int32_t fetched;       // samples requested so far (relative to begin, multiple of kZbChunk)
    int32_t last;          // last readable stream index (requests beyond it are clamped; never consumed)

    __device__ __forceinline__ void issue() {
        const int r = fetched & (kZbRing - 1);
        const uint32_t dst = ring + (uint32_t)r * 128u;
        const int32_t g0 = begin + fetched;
        if (g0 + kZbChunk - 1 <= last) {
            const float* g = z + g0;
#pragma unroll
            for (int k = 0; k < kZbChunk; k++) cp_async4(dst + 128u * k, g + k);
            if (r == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) cp_async4(dst + 128u * (kZbRing + k), g + k);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kZbChunk; k++) cp_async4(dst + 128u * k, z + min(g0 + k, last));
            if (r == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) cp_async4(dst + 128u * (kZbRing + k), z + min(g0 + k, last));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        fetched += kZbChunk;
    }
    __device__ __forceinline__ void prime() {
        while (fetched < 8 + kZbAhead) issue();
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    // the chain advances by at most 3 samples per step and one group brings 16, so one issue per step
    // keeps `fetched >= rel + 8 + kZbAhead`; everything but the 3 newest groups has landed after the wait
    __device__ __forceinline__ void get8(int32_t ii, float (&in)[8]) {
        const int rel = ii - begin;
        if (fetched < rel + 8 + kZbAhead) issue();
        asm volatile("cp.async.wait_group 3;\n" ::: "memory");
        const uint32_t p = ring + (uint32_t)(rel & (kZbRing - 1)) * 128u;
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(in[k]) : "r"(p + 128u * k));
    }
    __device__ __forceinline__ void row(int r, float (&t)[8]) const {
        const uint32_t a = taps + (uint32_t)r * (SNRX_MMSE_NTAPS * 4);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t[0]), "=f"(t[1]), "=f"(t[2]), "=f"(t[3]) : "r"(a));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t[4]), "=f"(t[5]), "=f"(t[6]), "=f"(t[7]) : "r"(a + 16u));
    }
};

__global__ void __launch_bounds__(kZbChainThreads) k_zb_chain(const float* __restrict__ z, ZbChainParams p,
                                                 const float* __restrict__ taps_g, ChipMap map_arg,
                                                 const int32_t* __restrict__ channel_numbers,
                                                 snrx_frame_t* __restrict__ slots, uint32_t* __restrict__ counts,
                                                 int64_t* __restrict__ good_end,
                                                 float* chips_dbg, int64_t chips_cap_per_chain, int64_t* nchips_dbg) {
    __shared__ __align__(16) float taps[(SNRX_MMSE_NSTEPS + 1) * SNRX_MMSE_NTAPS];
    __shared__ float ring[(kZbChainThreads / 32) * (kZbRing + 8) * 32];
    for (int i = threadIdx.x; i < (SNRX_MMSE_NSTEPS + 1) * SNRX_MMSE_NTAPS; i += blockDim.x) taps[i] = taps_g[i];
    __syncthreads();
    const uint32_t total = p.n_captures * p.n_channels * (uint32_t)p.n_segments;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const uint32_t* map = map_arg.w;            // kernel parameter: the chip words are constant-bank operands
    // consecutive threads take different streams so that a warp touches many DRAM pages at once
    const uint32_t seg = idx / (p.n_captures * p.n_channels);
    const uint32_t sc = idx % (p.n_captures * p.n_channels);
    const uint32_t cap = sc / p.n_channels, ch = sc % p.n_channels;
    const uint32_t chain = (cap * p.n_channels + ch) * (uint32_t)p.n_segments + seg;    // output order
    ZbRingSrc src;
    src.z = z + (size_t)sc * p.z_stride;
    src.ring = (uint32_t)__cvta_generic_to_shared(ring + (threadIdx.x >> 5) * ((kZbRing + 8) * 32) + (threadIdx.x & 31));
    src.taps = (uint32_t)__cvta_generic_to_shared(taps);
    asm volatile("" : "+r"(src.ring), "+r"(src.taps));          // opaque: stay in registers
    {
        const int32_t lo = p.origin + (int32_t)seg * p.segment;
        src.begin = lo - p.prehalo > 0 ? lo - p.prehalo : 0;                  // = the chain's first sample
    }
    src.fetched = 0;
    src.last = p.n_out > 0 ? p.n_out - 1 : 0;
    src.prime();
    int64_t nchips = 0;
    const uint32_t nf = zb_run_chain(src, p, (int)seg, map, channel_numbers[ch], p.first_capture + cap,
                                     slots + (size_t)chain * p.slots_per_chain,
                                     chips_dbg ? chips_dbg + (size_t)chain * chips_cap_per_chain : nullptr,
                                     chips_cap_per_chain, &nchips, good_end + chain);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    counts[chain] = nf < p.slots_per_chain ? nf : p.slots_per_chain;
    if (nchips_dbg) nchips_dbg[chain] = nchips;
}

// one thread per chain: drop the CRC-failed records that lie inside a CRC-ok frame of this or a preceding chain
__global__ void __launch_bounds__(128) k_zb_span_filter(snrx_frame_t* __restrict__ slots, uint32_t slots_per_chain,
                                                        uint32_t* __restrict__ counts, const int64_t* __restrict__ good_end,
                                                        uint32_t n_chains, uint32_t n_segments, int lookback) {
    const uint32_t chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    const uint32_t n = counts[chain];
    if (n == 0) return;
    const uint32_t seg = chain % n_segments;
    int64_t ge = 0;
    for (int j = 1; j <= lookback && (uint32_t)j <= seg; j++) { const int64_t e = good_end[chain - j]; if (e > ge) ge = e; }
    const uint32_t w = zb_filter_chain(slots + (size_t)chain * slots_per_chain, n, ge);
    if (w != n) counts[chain] = w;
}

__global__ void __launch_bounds__(128) k_zb_gather(const snrx_frame_t* __restrict__ slots, uint32_t slots_per_chain,
                                                   const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                                   uint32_t n_chains, snrx_frame_t* __restrict__ frames, uint32_t frame_cap,
                                                   uint32_t* __restrict__ totals, int after_ble) {
    // one warp per chain, each frame copied as 10 x 16 bytes
    const int lane = threadIdx.x & 31;
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t base = after_ble ? totals[0] : 0u;
    for (uint32_t c = blockIdx.x * warps_per_block + (threadIdx.x >> 5); c < n_chains; c += gridDim.x * warps_per_block) {
        const uint32_t n = counts[c], off = base + offsets[c];
        for (uint32_t k = 0; k < n; k++) {
            if (off + k >= frame_cap) break;
            const uint4* src = reinterpret_cast<const uint4*>(slots + (size_t)c * slots_per_chain + k);
            uint4* dst = reinterpret_cast<uint4*>(frames + off + k);
            if (lane < 10) dst[lane] = src[lane];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) totals[2] = offsets[n_chains];
}

// ------------------------------------------------------------------------------------ host side
struct ZbState {
    bool ready = false;
    uint32_t n_ch = 0, max_caps = 0, max_out = 0;
    size_t stride = 0;                 // floats per (capture, channel) stream
    float *d_f = nullptr, *d_z = nullptr;
    double *d_block_end = nullptr, *d_carry = nullptr, *d_pw = nullptr;
    float *d_atan = nullptr, *d_mmse = nullptr;
    int32_t* d_channels = nullptr;
    snrx_frame_t* d_slots = nullptr; size_t slots_bytes = 0;
    uint32_t *d_counts = nullptr, *d_offsets = nullptr, *d_scratch = nullptr;
    int64_t* d_good_end = nullptr;
    uint32_t max_chains = 0, slots_per_chain = 0;
    float* d_chips = nullptr; int64_t* d_nchips = nullptr; int64_t chips_cap = 0;
    float *d_wb_taps_rho = nullptr, *d_wb_taps_flat = nullptr, *d_wb_taps_pass = nullptr; float2* d_wb_cf = nullptr; int wb_nt = 16;
    bool wb_cta_kernel = false;   // wideband front end
    ChipMap map;
    uint32_t last_chains = 0;
};

inline void zb_free(ZbState& s) {
    void* bufs[] = {s.d_f, s.d_z, s.d_block_end, s.d_carry, s.d_pw, s.d_atan, s.d_mmse, s.d_channels, s.d_slots,
                    s.d_counts, s.d_offsets, s.d_scratch, s.d_good_end, s.d_chips, s.d_nchips, s.d_wb_taps_rho, s.d_wb_taps_flat, s.d_wb_taps_pass, s.d_wb_cf};
    for (void* b : bufs) if (b) cudaFree(b);
    s = ZbState();
}

#define ZCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            err = b__;                                                                             \
            return e__ == cudaErrorMemoryAllocation ? SNRX_ENOMEM : SNRX_ECUDA;                    \
        }                                                                                          \
    } while (0)

inline int zb_create(ZbState& s, const snrx_config_t& cfg, bool wideband, uint32_t n_ch, uint32_t max_caps,
                     uint32_t max_out, int sm_count, std::string& err) {
    (void)sm_count;
    s.n_ch = n_ch; s.max_caps = max_caps; s.max_out = max_out;
    if ((uint64_t)max_out + cfg.zb_segment + kZbPostHalo >= (1ull << 31)) { err = "zigbee: capture too long for 32-bit chain positions"; return SNRX_ERANGE; }
    s.stride = ((size_t)max_out + 8 + 31) & ~(size_t)31;
    const size_t streams = (size_t)max_caps * n_ch;
    ZCK(cudaMalloc((void**)&s.d_f, streams * s.stride * sizeof(float)));
    ZCK(cudaMalloc((void**)&s.d_z, streams * s.stride * sizeof(float)));
    const size_t nblk = (max_out + SNRX_IIR_BLOCK - 1) / SNRX_IIR_BLOCK;
    ZCK(cudaMalloc((void**)&s.d_block_end, streams * nblk * sizeof(double)));
    ZCK(cudaMalloc((void**)&s.d_carry, streams * nblk * sizeof(double)));
    std::vector<double> pw(SNRX_IIR_BLOCK);
    { volatile double p = 1.0; const double b = SNRX_IIR_BETA; for (int i = 0; i < SNRX_IIR_BLOCK; i++) { p = p * b; pw[i] = p; } }
    ZCK(cudaMalloc((void**)&s.d_pw, sizeof(double) * SNRX_IIR_BLOCK));
    ZCK(cudaMemcpy(s.d_pw, pw.data(), sizeof(double) * SNRX_IIR_BLOCK, cudaMemcpyHostToDevice));
    ZCK(cudaMalloc((void**)&s.d_atan, sizeof(SNRX_ATAN_TAB)));
    ZCK(cudaMemcpy(s.d_atan, SNRX_ATAN_TAB, sizeof(SNRX_ATAN_TAB), cudaMemcpyHostToDevice));
    ZCK(cudaMalloc((void**)&s.d_mmse, sizeof(SNRX_MMSE_TAPS)));
    ZCK(cudaMemcpy(s.d_mmse, SNRX_MMSE_TAPS, sizeof(SNRX_MMSE_TAPS), cudaMemcpyHostToDevice));
    int32_t chans[16];
    for (int i = 0; i < 16; i++) chans[i] = wideband ? 11 + i : cfg.channel;
    ZCK(cudaMalloc((void**)&s.d_channels, sizeof chans));
    ZCK(cudaMemcpy(s.d_channels, chans, sizeof chans, cudaMemcpyHostToDevice));
    const uint32_t max_segs = (max_out + cfg.zb_segment - 1) / cfg.zb_segment;
    s.max_chains = (uint32_t)(streams * max_segs);
    s.slots_per_chain = cfg.zb_segment / kZbMinFrameSamples + 2;
    s.slots_bytes = (size_t)s.max_chains * s.slots_per_chain * sizeof(snrx_frame_t);
    ZCK(cudaMalloc((void**)&s.d_slots, s.slots_bytes));
    ZCK(cudaMalloc((void**)&s.d_counts, sizeof(uint32_t) * ((size_t)s.max_chains + 1)));
    ZCK(cudaMalloc((void**)&s.d_offsets, sizeof(uint32_t) * ((size_t)s.max_chains + 1)));
    ZCK(cudaMalloc((void**)&s.d_scratch, sizeof(uint32_t) * scan_scratch_items(s.max_chains)));
    ZCK(cudaMalloc((void**)&s.d_good_end, sizeof(int64_t) * ((size_t)s.max_chains + 1)));
    if (cfg.flags & SNRX_F_KEEP_STREAMS) {
        s.chips_cap = ((int64_t)cfg.zb_segment + cfg.zb_prehalo + kZbPostHalo) / 2 + 64;
        ZCK(cudaMalloc((void**)&s.d_chips, sizeof(float) * (size_t)s.max_chains * (size_t)s.chips_cap));
        ZCK(cudaMalloc((void**)&s.d_nchips, sizeof(int64_t) * (size_t)s.max_chains));
    }
    s.map = make_chip_map();
    s.ready = true;
    return SNRX_OK;
}

// wideband front end (channelizer + discriminator): defined in pfb_zb.cuh
inline int zb_wideband_front(ZbState& s, const snrx_config_t& cfg, const float2* x, uint32_t n_captures, uint64_t n_samples,
                      uint64_t stride, uint32_t n_out, cudaStream_t st, int& launches, std::string& err);

inline int zb_process(ZbState& s, const snrx_config_t& cfg, const float2* x, uint32_t n_captures, uint64_t n_samples,
                      uint64_t stride, uint32_t n_out, uint32_t pre_out, uint32_t body_out, uint32_t first_segment,
                      uint32_t first_capture, snrx_frame_t* frames, uint32_t frame_cap, uint32_t* totals, bool after_ble,
                      cudaStream_t st, int sm_count, int& launches, std::string& err, cudaEvent_t ev_front_done = nullptr) {
    const bool wideband = (cfg.mode != SNRX_MODE_ZB_NB);
    const uint32_t streams = n_captures * s.n_ch;
    if (wideband) {
        int r = zb_wideband_front(s, cfg, x, n_captures, n_samples, stride, n_out, st, launches, err);
        if (r != SNRX_OK) return r;
    } else {
        ZbQuadArgs q;
        q.x = x; q.stride = stride; q.n = (int64_t)n_samples; q.n_captures = n_captures;
        q.f = s.d_f; q.f_stride = s.stride; q.atan_tab = s.d_atan;
        const uint64_t total = (uint64_t)n_captures * n_samples;
        const int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)sm_count * 8);
        k_zb_quad<<<grid, 256, 0, st>>>(q);
        launches++;
    }
    if (ev_front_done) ZCK(cudaEventRecord(ev_front_done, st));      // front end = channelizer + discriminator (stats.gpu_ms_frontend)
    ZbIirArgs ia;
    ia.f = s.d_f; ia.z = s.d_z; ia.stride = s.stride; ia.n = (int32_t)n_out;
    ia.n_blocks = (int32_t)((n_out + SNRX_IIR_BLOCK - 1) / SNRX_IIR_BLOCK); ia.n_streams = streams;
    ia.block_end = s.d_block_end; ia.carry_in = s.d_carry; ia.pw = s.d_pw;
    const uint32_t nb_total = streams * (uint32_t)ia.n_blocks;
    k_zb_iir_sum<<<(nb_total + 63) / 64, 64, 0, st>>>(ia);
    k_zb_iir_carry<<<(nb_total + 127) / 128, 128, 0, st>>>(ia);
    k_zb_dc<<<(nb_total + 63) / 64, 64, 0, st>>>(ia);
    launches += 3;

    ZbChainParams p{};
    p.n_out = (int32_t)n_out; p.origin = (int32_t)pre_out; p.body = (int32_t)body_out;
    p.segment = (int32_t)cfg.zb_segment; p.prehalo = (int32_t)cfg.zb_prehalo;
    p.n_segments = (int32_t)((body_out + cfg.zb_segment - 1) / cfg.zb_segment);
    p.first_segment = first_segment; p.first_capture = first_capture;
    p.n_captures = n_captures; p.n_channels = s.n_ch; p.threshold = cfg.zb_threshold;
    p.slots_per_chain = s.slots_per_chain; p.z_stride = s.stride;
    const uint32_t n_chains = streams * (uint32_t)p.n_segments;
    if (n_chains > s.max_chains) { err = "zigbee: more chains than capacity"; return SNRX_ERANGE; }
    k_zb_chain<<<(n_chains + kZbChainThreads - 1) / kZbChainThreads, kZbChainThreads, 0, st>>>(s.d_z, p, s.d_mmse, s.map, s.d_channels, s.d_slots, s.d_counts,
                                                  s.d_good_end, s.d_chips, s.chips_cap, s.d_nchips);
    k_zb_span_filter<<<(n_chains + 127) / 128, 128, 0, st>>>(s.d_slots, s.slots_per_chain, s.d_counts, s.d_good_end, n_chains,
                                                             (uint32_t)p.n_segments, zb_filter_lookback(p.segment));
    launches += 2 + exclusive_scan(s.d_counts, n_chains, s.d_offsets, s.d_scratch, st);
    k_zb_gather<<<std::max(1u, std::min<uint32_t>((n_chains + 3) / 4, (uint32_t)sm_count * 8)), 128, 0, st>>>(
        s.d_slots, s.slots_per_chain, s.d_counts, s.d_offsets, n_chains, frames, frame_cap, totals, after_ble ? 1 : 0);
    launches++;
    ZCK(cudaGetLastError());
    s.last_chains = n_chains;
    return SNRX_OK;
}

inline int zb_debug_stage(ZbState& s, int stage, uint32_t caps, uint32_t n_out, const void** src, uint64_t* bytes) {
    (void)n_out;
    if (!s.ready) return SNRX_ESTATE;
    if (stage == SNRX_STAGE_ZB_DISC) { *src = s.d_z; *bytes = (uint64_t)caps * s.n_ch * s.stride * sizeof(float); return SNRX_OK; }
    if (stage == SNRX_STAGE_ZB_CHIPS) {
        if (!s.d_chips) return SNRX_ESTATE;
        *src = s.d_chips; *bytes = (uint64_t)s.last_chains * (uint64_t)s.chips_cap * sizeof(float); return SNRX_OK;
    }
    if (stage == SNRX_STAGE_ZB_NCHIPS) {
        if (!s.d_nchips) return SNRX_ESTATE;
        *src = s.d_nchips; *bytes = (uint64_t)s.last_chains * sizeof(int64_t); return SNRX_OK;
    }
    if (stage == SNRX_STAGE_ZB_F) { *src = s.d_f; *bytes = (uint64_t)caps * s.n_ch * s.stride * sizeof(float); return SNRX_OK; }
    if (stage == SNRX_STAGE_CHAN_CF32) {
        if (!s.d_wb_cf) return SNRX_ESTATE;
        *src = s.d_wb_cf; *bytes = (uint64_t)caps * s.n_ch * n_out * sizeof(float2); return SNRX_OK;
    }
    return SNRX_EINVAL;
}
#endif  // __CUDACC__

}  // namespace snrx
