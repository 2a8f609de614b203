/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into the product.
 *
 * Compiles the UNMODIFIED reference 802.15.4 packet sink
 *   scapy-radio/gnuradio/gr-zigbee/lib/packet_sink_scapy_impl.cc
 * (path passed as -DREF_ZB_SINK_CC=...) against the stub headers in oracle/stub and
 * exposes a C interface that feeds it soft chips one at a time, so that the chip at
 * which every blob is published is known.
 */
#define HAVE_CONFIG_H 1
#include REF_ZB_SINK_CC

#include <cstring>

using gr::zigbee::packet_sink_scapy_impl;

extern "C" {

void* zb_ref_new(int threshold) { return new packet_sink_scapy_impl(threshold); }
void zb_ref_delete(void* p) { delete (packet_sink_scapy_impl*)p; }

/* Feed n soft chips.  For every published blob stores its end chip index (index of the chip
 * that completed the frame, relative to the first chip ever fed to this instance + base),
 * its length and its bytes (the 8-byte header is stripped) into the out arrays.
 * Returns the number of blobs published during this call. */
int zb_ref_feed(void* p, const float* chips, int64_t n, int64_t base,
                int64_t* end_chip, int32_t* len, uint8_t* bytes /* cap*128 */, int cap) {
    packet_sink_scapy_impl* s = (packet_sink_scapy_impl*)p;
    int k = 0;
    gr_vector_int ni(1, 1);
    gr_vector_const_void_star in(1);
    gr_vector_void_star out;
    for (int64_t i = 0; i < n; i++) {
        size_t before = s->published.size();
        in[0] = chips + i;
        s->general_work(0, ni, in, out);
        for (size_t j = before; j < s->published.size(); j++) {
            const std::vector<uint8_t>& b = s->published[j];
            if (k < cap) {
                end_chip[k] = base + i;
                len[k] = (int32_t)b.size() - 8;
                std::memset(bytes + 128 * k, 0, 128);
                if (b.size() > 8) std::memcpy(bytes + 128 * k, b.data() + 8, b.size() - 8);
            }
            k++;
        }
    }
    s->published.clear();
    return k;
}

/* same but in one general_work() call (checks that chunking does not matter) */
int zb_ref_feed_block(void* p, const float* chips, int64_t n) {
    packet_sink_scapy_impl* s = (packet_sink_scapy_impl*)p;
    gr_vector_int ni(1, (int)n);
    gr_vector_const_void_star in(1, chips);
    gr_vector_void_star out;
    s->general_work(0, ni, in, out);
    int k = (int)s->published.size();
    return k;
}
int zb_ref_take(void* p, int idx, uint8_t* bytes128) {
    packet_sink_scapy_impl* s = (packet_sink_scapy_impl*)p;
    if (idx < 0 || idx >= (int)s->published.size()) return -1;
    const std::vector<uint8_t>& b = s->published[idx];
    std::memset(bytes128, 0, 128);
    if (b.size() > 8) std::memcpy(bytes128, b.data() + 8, b.size() - 8);
    return (int)b.size() - 8;
}

unsigned int zb_ref_chip_mapping(int i) { return CHIP_MAPPING[i & 15]; }

}  // extern "C"
