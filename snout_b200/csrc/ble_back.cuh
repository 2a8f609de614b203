// ble_back.cuh -- BLE back end on sliced bit streams: access-address sliding correlation,
// per-candidate de-whitening + CRC-24, and the per-window resolver that replays the
// order-dependent acceptance rules of the reference receiver.
//
// Reference behaviour reproduced (vendor/BTLE/host/btle-tools/src/btle_rx.c):
//   search_unique_bits()  1369-1421   sliding 32-bit compare over 4 sample phases, history
//                                     zeroed at every call (1377)
//   demod_byte()          1348-1367   LSB-first packing of symbol-spaced bit decisions
//   scramble_byte()       1158-1163   XOR with scramble_table[channel]
//   crc_update/crc_check  1137-1156, 1826-1848
//   parse_*_header_byte   1771-1795   length fields, ADV gate 6..37 in receiver() 2096-2104
//   receiver()            2020-2155   search -> header -> payload -> resume after the frame
//   main loop             2341-2393   one receiver() call per 8192-IQ half buffer
//
// The kernels work on bits only: the slicer decision b[n] = (I[n]Q[n+1] - I[n+1]Q[n]) > 0 is all
// that search_unique_bits() and demod_byte() ever look at, so the front ends (narrow-band slicer,
// wideband channelizer) hand over one bit per channel-rate sample in natural order (BitsLayout,
// common.cuh): bit (n & 31) of word kBitsLeadWords + (n >> 5) is the decision of sample n.
#pragma once
#include "common.cuh"

namespace snrx {

SNRX_HD int hi_bit_plus1(uint32_t d) {   // 0 for d == 0, else index of highest set bit + 1
#ifdef __CUDA_ARCH__
    return 32 - __clz(d);
#else
    int n = 0; while (d) { n++; d >>= 1; } return n;
#endif
}

// number of low access-address positions that a zeroed history satisfies "for free"
// (btle_rx.c:1377,1395-1401): positions p with aa[p] == 0 or mask[p] == 0, counted from p = 0
SNRX_HD int aa_virtual_bits(uint32_t aa, uint32_t mask) {
    uint32_t v = aa & mask;
    int z = 0;
    while (z < 31 && !((v >> z) & 1u)) z++;
    return z;
}

// Sliding correlation over the 32 start positions held by one word of a bit stream, bit-parallel
// ("shift-and"): bit i of the result is set when the 32 symbol-spaced decisions (every 4th sample)
// starting at the word's sample i equal the access address in every masked position >= z (full matches
// and the "virtual" matches that are usable only at a search origin).  w[0] is the word itself, w[1..4]
// the four that follow: access-address bit p is compared with the stream shifted by 4p samples.  After
// bit p has been applied a random position survives with probability 2^-(p+1); `keep_going` lets a warp
// stop as soon as none of its lanes has a survivor.
// The per-position constants of the correlation, computed once on the host and passed as a kernel parameter: in the kernel they
// are constant-bank operands of ONE LOP3 per access-address position (computing them from aa / mask in the loop cost three
// uniform-datapath instructions per position: 143 of the kernel's 560 instructions).
struct AaTables {
    uint32_t want0[32];      // all ones when AA bit p is 0
    uint32_t dont_care[32];  // all ones when position p is not compared
};
inline AaTables make_aa_tables(uint32_t aa, uint32_t mask_hi /* mask & ~((1<<z)-1) */) {
    AaTables t;
    for (int p = 0; p < 32; p++) { t.want0[p] = ((aa >> p) & 1u) - 1u; t.dont_care[p] = ((mask_hi >> p) & 1u) - 1u; }
    return t;
}

template <class KEEP>
SNRX_HD uint32_t aa_word_hits(const uint32_t (&w)[5], const AaTables& t, KEEP keep_going) {
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int p = 0; p < 32; p++) {
        const uint32_t s = (p & 7) ? funnel_r(w[p >> 3], w[(p >> 3) + 1], 4 * (p & 7)) : w[p >> 3];   // bit i = sample i + 4p
        m &= (s ^ t.want0[p]) | t.dont_care[p];
        if ((p & 3) == 3 && p >= 11 && p < 31) { if (!keep_going(m)) return 0u; }
    }
    return m;
}

// Finish the decode of a candidate from its de-whitened 4-byte chunks (chunk c = bytes 4c..4c+3
// counted from the PDU header).  Mirrors receiver() btle_rx.c:2066-2125.  Written with compile-time
// indices only (the byte loop is unrolled and predicated) so that one GPU thread per candidate keeps
// everything in registers.
SNRX_HD void ble_finish(const uint32_t (&chunk)[11], bool adv_channel, uint32_t crc_init_internal,
                        const uint32_t* crc_tab, Dec& d) {
    uint32_t* bw = reinterpret_cast<uint32_t*>(d.bytes);            // Dec::bytes is 4-byte aligned
#pragma unroll
    for (int c = 0; c < 11; c++) bw[c] = chunk[c];
    const int b1 = (int)((chunk[0] >> 8) & 0xFFu);
    const int len = adv_channel ? (b1 & 0x3F) : (b1 & 0x1F);        // btle_rx.c:1794 / 1776
    d.len = (uint8_t)len;
    d.emit = (uint8_t)(adv_channel ? (len >= 6 && len <= 37) : 1);   // btle_rx.c:2096
    uint32_t crc = crc_init_internal, recv = 0;
#pragma unroll
    for (int i = 0; i < 42; i++) {
        const uint32_t byte = (chunk[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        const int k = i - (len + 2);
        if (k < 0) crc = (crc_tab[(crc ^ byte) & 0xFFu] ^ (crc >> 8)) & 0xFFFFFFu;       // crc_update, btle_rx.c:1137-1148
        else if (k < 3) recv |= byte << (8 * k);                                        // crc_check, btle_rx.c:1826-1848
    }
    d.crc_ok = (uint8_t)(d.emit && crc == recv);
}

// Replay of receiver() for one 8192-IQ window starting at local sample W.
//   cands[c0, c1) : hits of this (capture, channel), ascending in s
//   EMIT(idx)     : called for every accepted, printed frame with the candidate index
// Returns the number of frames.
template <class EMIT>
SNRX_HD int ble_resolve_window(const Cand* cands, const Dec* decs, int c0, int c1, int W, int z, EMIT emit) {
    int n = 0;
    int eaten = 0;                                   // int8 units past W, as the reference counts
    for (;;) {
        int slots = (kSpanInt8 - eaten) / (kSps * 2);    // num_symbol_left, btle_rx.c:2032,2075,2123
        if (slots <= 0) break;
        const int P = W + eaten / 2;                 // search origin (local sample index)
        // first candidate with s >= P - 4z
        int lo = c0, hi = c1;
        const int s_min = P - 4 * z;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (cands[mid].s < s_min) lo = mid + 1; else hi = mid; }
        int found = -1;
        for (int k = lo; k < c1; k++) {
            const int s = cands[k].s;
            const int vneed = cands[k].vneed;        // number of low AA bits that do NOT match (0 = full)
            if (s >= P) {
                int t = 31 + (s - P) / kSps;
                if (t >= slots) break;               // beyond this call's search span
                if (vneed == 0) { found = k; break; }
            } else {
                int v = (P - s + kSps - 1) / kSps;   // history slots still zero when this position is tested
                int t = 31 - v;
                if (v <= z && vneed <= v && t < slots) { found = k; break; }
            }
        }
        if (found < 0) break;
        const int s = cands[found].s;
        eaten = 2 * (s - W) + 32 * kSps * 2 + 16 * kSps * 2;     // AA + 2 header bytes, btle_rx.c:2058,2066
        if (eaten > kDemodLimitInt8) break;          // btle_rx.c:2067
        const Dec& d = decs[found];
        if (!d.emit) continue;                       // ADV length gate: resume after the header
        eaten += 8 * ((int)d.len + 3) * kSps * 2;    // btle_rx.c:2113
        if (eaten > kDemodLimitInt8) break;          // btle_rx.c:2115
        emit(found);
        n++;
    }
    return n;
}

// Build the 160-byte frame record in registers and store it as ten 16-byte words.
SNRX_HD void ble_fill_frame(snrx_frame_t& f, const Cand& c, const Dec& d, const BleParams& p, int window_local,
                            int channel_number) {
    uint32_t w[40];
    const int64_t si = (int64_t)(c.s - p.m_origin) + (int64_t)kWindow * p.first_window;
    w[0] = (uint32_t)(uint64_t)si;
    w[1] = (uint32_t)((uint64_t)si >> 32);
    w[2] = p.first_capture + c.cap;
    w[3] = p.first_window + (uint32_t)window_local;
    w[4] = (uint32_t)(channel_number & 0xFFFF) | ((uint32_t)SNRX_PROTO_BLE << 16) | ((uint32_t)d.crc_ok << 24);
    const uint32_t nb = (uint32_t)d.len + 5u;
    w[5] = ((uint32_t)(((c.s % 4) + 4) % 4) << 8) | (nb << 16);          // lqi = 0 | phase | len
    w[6] = p.aa;
    const uint32_t* db = reinterpret_cast<const uint32_t*>(d.bytes);     // Dec::bytes is 4-byte aligned
#pragma unroll
    for (int j = 0; j < 11; j++) {
        const int left = (int)nb - 4 * j;
        const uint32_t keep = left >= 4 ? 0xFFFFFFFFu : left <= 0 ? 0u : ((1u << (8 * left)) - 1u);
        w[7 + j] = db[j] & keep;
    }
#pragma unroll
    for (int j = 18; j < 40; j++) w[j] = 0u;
    uint4* dst = reinterpret_cast<uint4*>(&f);
#pragma unroll
    for (int j = 0; j < 10; j++) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------ kernels

// Sliding access-address correlation.  One warp per (capture, channel, chunk of 32 words); each lane
// owns one word (32 start positions) and reads the four words that follow it.  Writes the hit masks
// (bit i of hits[w] = the 32 symbols starting at sample 32 (w - kBitsLeadWords) + i match) and the number
// of hits per chunk; after the prefix sum k_aa_fill turns the masks into the candidate list, ascending in
// s, without atomics or a sort.
__global__ void __launch_bounds__(256) k_aa_search(const uint32_t* __restrict__ bits, BitsLayout lay, BleParams p,
                                                   uint32_t n_chunks, uint32_t* __restrict__ counts,
                                                   uint32_t* __restrict__ hits_out, const __grid_constant__ AaTables tabs) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t n_items = p.n_captures * p.n_channels * n_chunks;
    const uint32_t nw = lay.words_per_stream;
    for (uint32_t item = blockIdx.x * warps_per_block + (threadIdx.x >> 5); item < n_items;
         item += gridDim.x * warps_per_block) {
        const uint32_t chunk = item % n_chunks;
        const uint32_t ch = (item / n_chunks) % p.n_channels;
        const uint32_t cap = item / (n_chunks * p.n_channels);
        const uint32_t w = chunk * 32 + lane;                       // word owned by this lane
        const uint32_t* pw = bits + lay.index(cap, ch, 0);
        uint32_t ww[5];
#pragma unroll
        for (int d = 0; d < 5; d++) ww[d] = (w + d < nw) ? __ldg(pw + w + d) : 0u;
        uint32_t h = aa_word_hits(ww, tabs, [](uint32_t m) { return __any_sync(0xffffffffu, m != 0u) != 0; });
        // start positions beyond the capture carry no data
        const int nvalid = p.n_out - 32 * ((int)w - kBitsLeadWords);
        if (nvalid <= 0 || w >= nw) h = 0u; else if (nvalid < 32) h &= (1u << nvalid) - 1u;
        if (w < nw) hits_out[lay.index(cap, ch, w)] = h;
        int cnt = __popc(h);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) counts[item] = (uint32_t)cnt;
    }
}

// Candidate list from the hit masks.  One warp per chunk that has hits; lanes own words, an exclusive
// prefix over the lanes gives each lane its write position; bit order is sample order.
__global__ void __launch_bounds__(256) k_aa_fill(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ hits_in,
                                                 BitsLayout lay, BleParams p, uint32_t n_chunks,
                                                 const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                                 Cand* __restrict__ cands, uint32_t cand_cap) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t n_items = p.n_captures * p.n_channels * n_chunks;
    // a warp takes 32 consecutive items at a time and reads their counts in ONE coalesced load: most chunks hold no hit, and
    // a dependent global load per item made this kernel a chain of L2 round trips
    for (uint32_t base = (blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * 32u; base < n_items; base += gridDim.x * warps_per_block * 32u) {
      const uint32_t mine = (base + lane < n_items) ? __ldg(counts + base + lane) : 0u;
      uint32_t todo = __ballot_sync(0xffffffffu, mine != 0u);
      while (todo) {
        const uint32_t item = base + (uint32_t)(__ffs(todo) - 1);
        todo &= todo - 1u;
        const uint32_t chunk = item % n_chunks;
        const uint32_t ch = (item / n_chunks) % p.n_channels;
        const uint32_t cap = item / (n_chunks * p.n_channels);
        const uint32_t w = chunk * 32 + lane;
        uint32_t hits = (w < lay.words_per_stream) ? __ldg(hits_in + lay.index(cap, ch, w)) : 0u;
        const int cnt = __popc(hits);
        int pre = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += v; }
        pre -= cnt;
        uint32_t dst = offsets[item] + (uint32_t)pre;
        const uint32_t* pw = bits + lay.index(cap, ch, 0);
        while (hits) {
            const int i = __ffs(hits) - 1;
            hits &= hits - 1u;
            const int s = 32 * ((int)w - kBitsLeadWords) + i;
            const uint32_t d = (symbols32(pw, s) ^ p.aa) & p.aa_mask;
            if (dst < cand_cap) {
                Cand c;
                c.s = s; c.ch_idx = (uint16_t)ch; c.vneed = (uint8_t)hi_bit_plus1(d); c.pad = 0; c.cap = cap;
                cands[dst] = c;
            }
            dst++;
        }
      }
    }
}

// One THREAD per candidate: pulls the 11 x 32 symbol decisions of the longest possible frame out of the
// bit stream (45 consecutive words), de-whitens them, applies the header rules and runs the CRC-24 from
// a shared-memory table.
__global__ void __launch_bounds__(128) k_ble_decode(const uint32_t* __restrict__ bits, BitsLayout lay, BleParams p,
                                                    const uint32_t* __restrict__ n_cands_dev, uint32_t cand_cap,
                                                    const Cand* __restrict__ cands, Dec* __restrict__ decs,
                                                    const uint32_t* __restrict__ crc_tab,
                                                    const uint32_t* __restrict__ whiten /*[40][11]*/,
                                                    const int32_t* __restrict__ channel_numbers) {
    __shared__ uint32_t crc_s[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) crc_s[i] = crc_tab[i];
    __syncthreads();
    uint32_t n = *n_cands_dev;
    if (n > cand_cap) n = cand_cap;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const Cand c = cands[k];
        const int chn = channel_numbers[c.ch_idx];
        const uint32_t* pw = bits + lay.index(c.cap, c.ch_idx, 0);
        // header starts 32 symbols = 128 samples after AA bit 0
        const int n0 = c.s + 128 + 32 * kBitsLeadWords, w0 = n0 >> 5, sh = n0 & 31;
        uint32_t chunk[11];
        uint32_t lo = __ldg(pw + w0);
#pragma unroll
        for (int q = 0; q < 11; q++) {
            uint32_t r = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t hi = __ldg(pw + w0 + 4 * q + b + 1);
                r |= compress4(funnel_r(lo, hi, sh)) << (8 * b);
                lo = hi;
            }
            chunk[q] = r ^ __ldg(whiten + chn * 11 + q);
        }
        Dec d;
        d.s = c.s; d.resume = 0; d.vneed = c.vneed;
        ble_finish(chunk, chn >= 37 && chn <= 39, p.crc_init_internal, crc_s, d);
        decs[k] = d;
    }
}

// One thread per (capture, channel, window): replay of receiver().  COUNT pass stores the
// number of frames of the window, FILL pass writes the frame records at the scanned offsets.
template <bool FILL>
__global__ void __launch_bounds__(256) k_ble_resolve(const Cand* __restrict__ cands, const Dec* __restrict__ decs,
                                                     const uint32_t* __restrict__ cand_offsets, uint32_t n_chunks,
                                                     BleParams p, uint32_t* counts,
                                                     const uint32_t* __restrict__ frame_offsets,
                                                     snrx_frame_t* __restrict__ frames, uint32_t frame_cap,
                                                     const int32_t* __restrict__ channel_numbers, uint32_t cand_cap) {
    const uint32_t n_items = p.n_captures * p.n_channels * (uint32_t)p.n_windows;
    const int z = aa_virtual_bits(p.aa, p.aa_mask);
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
        if (FILL && counts[item] == 0u) continue;
        const uint32_t w = item % (uint32_t)p.n_windows;
        const uint32_t ch = (item / (uint32_t)p.n_windows) % p.n_channels;
        const uint32_t cap = item / ((uint32_t)p.n_windows * p.n_channels);
        const uint32_t base = (cap * p.n_channels + ch) * n_chunks;
        uint32_t c0 = cand_offsets[base], c1 = cand_offsets[base + n_chunks];
        if (c0 > cand_cap) c0 = cand_cap;
        if (c1 > cand_cap) c1 = cand_cap;
        int nf = 0;
        if (c1 > c0) {
            const int W = p.m_origin + kWindow * (int)w;
            if (!FILL) {
                nf = ble_resolve_window(cands, decs, (int)c0, (int)c1, W, z, [](int) {});
            } else {
                uint32_t dst = frame_offsets[item];
                const int chn = channel_numbers[ch];
                nf = ble_resolve_window(cands, decs, (int)c0, (int)c1, W, z, [&](int k) {
                    if (dst < frame_cap) ble_fill_frame(frames[dst], cands[k], decs[k], p, (int)w, chn);
                    dst++;
                });
            }
        }
        if (!FILL) counts[item] = (uint32_t)nf;
    }
}
#endif  // __CUDACC__

}  // namespace snrx
