// zb_mac.cuh -- SURVEY 8(f) row N2: the Zigbee consumer path on the GPU-decoded 802.15.4 records.
//
// What Snout does with every datagram of the receive flowgraph (snout/util/zigbee.py:194-202 handle_packet ->
// RFtap(pkt) -> Dot15d4FCS dissection, snout/core/message.py:258-304 ZigbeeMessage.fromraw -> src_addr / dest_addr /
// seqnum / src_panid / dest_panid, and the touchlink scan's `haslayer(ZLLScanResponse)`, zigbee.py:176-192): one scapy
// object tree per packet.  Here one thread per record walks the same header: the MAC header exactly as the reference's
// dissector reads it (scapy-radio/scapy/scapy/layers/dot15d4.py: Dot15d4FCS :140-172 and the address rules of
// Dot15d4Data :216-231, Dot15d4Beacon :255-287, Dot15d4Cmd :294-331, util_srcpanid_present :359-364) and, for data frames,
// the inter-PAN path down to the ZLL commissioning command (dot15d4.py:233-245 with conf.dot15d4_protocol = 'zigbee',
// snout/cli.py:24; scapy/layers/zigbee.py ZigbeeNWKStub :717-731, ZigbeeAppDataPayloadStub :734-763,
// ZigbeeZLLCommissioningCluster :1027-1059).  A field the record is too short to hold sets SNRX_ZBMAC_MALFORMED
// (the reference dissector stops with what it has).  oracle/zbmac_oracle.py restates the rules in Python and is pinned
// against the imported reference dissector (tests/golden/zbmac_ref.json).
#pragma once
#include "common.cuh"

namespace snrx {

SNRX_HD uint64_t zb_le(const uint8_t* p, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

// One 802.15.4 record -> summary.  psdu = MHR | payload | FCS as in snrx_frame_t.bytes, len = valid bytes.
SNRX_HD void zb_mac_parse(const uint8_t* psdu, int len, snrx_zbmac_t& o) {
    o.dest_addr = 0; o.src_addr = 0; o.dest_panid = 0; o.src_panid = 0; o.fcf = 0; o.present = 0; o.seqnum = 0;
    o.frame_type = 0xFF; o.dest_mode = 0; o.src_mode = 0; o.cmd_id = 0xFF; o.payload_off = 0; o.zll_command = 0xFF;
    o.reserved = 0; o.cluster = 0; o.profile = 0;
    if (len < 5) { o.present |= SNRX_ZBMAC_MALFORMED; return; }      // FCF (2) + sequence number + FCS (2)
    const int n = len - 2;                                           // bytes in front of the FCS
    const unsigned b0 = psdu[0], b1 = psdu[1];
    o.fcf = (uint16_t)(b0 | (b1 << 8));
    o.frame_type = b0 & 7;
    const bool security = (b0 >> 3) & 1, compress = (b0 >> 6) & 1;
    if (security) o.present |= SNRX_ZBMAC_SECURITY;
    if ((b0 >> 4) & 1) o.present |= SNRX_ZBMAC_PENDING;
    if ((b0 >> 5) & 1) o.present |= SNRX_ZBMAC_ACKREQ;
    if (compress) o.present |= SNRX_ZBMAC_PANID_COMPRESS;
    o.dest_mode = (b1 >> 2) & 3;
    o.src_mode = (b1 >> 6) & 3;
    o.seqnum = psdu[2];
    int pos = 3;
    auto take = [&](int k) -> bool { if (pos + k > n) { o.present |= SNRX_ZBMAC_MALFORMED; pos = n; return false; } return true; };
    auto addr_len = [](int mode) { return mode == 2 ? 2 : mode == 3 ? 8 : 0; };   // dot15d4AddressField.lengthFromAddrMode
    // An address field whose mode has no length (0 = none where the field is unconditional, 1 = reserved) makes the
    // reference's dot15d4AddressField.getfield raise (dot15d4.py:60-63): the dissector then keeps the whole MAC payload as
    // raw bytes and Snout sees none of the addressing fields.
    if (((o.frame_type == 1 || o.frame_type == 3) && (o.dest_mode < 2 || o.src_mode == 1)) || (o.frame_type == 0 && o.src_mode < 2)) {
        o.present |= SNRX_ZBMAC_NO_ADDRESSING;
        o.payload_off = (uint8_t)pos;
        return;
    }
    if (o.frame_type == 1 || o.frame_type == 3) {                    // Data / Command: dest_panid is ALWAYS read (:217, :296)
        if (!take(2)) return;
        o.dest_panid = (uint16_t)zb_le(psdu + pos, 2); pos += 2; o.present |= SNRX_ZBMAC_DEST_PANID;
        const int dl = addr_len(o.dest_mode);
        if (dl) { if (!take(dl)) return; o.dest_addr = zb_le(psdu + pos, dl); pos += dl; o.present |= SNRX_ZBMAC_DEST_ADDR; }
        if (o.src_mode != 0 && !compress) {                          // util_srcpanid_present
            if (!take(2)) return;
            o.src_panid = (uint16_t)zb_le(psdu + pos, 2); pos += 2; o.present |= SNRX_ZBMAC_SRC_PANID;
        }
        if (o.src_mode != 0) {
            const int sl = addr_len(o.src_mode);                     // mode 1 (reserved): a field of length 0
            if (sl) { if (!take(sl)) return; o.src_addr = zb_le(psdu + pos, sl); pos += sl; o.present |= SNRX_ZBMAC_SRC_ADDR; }
        }
    } else if (o.frame_type == 0) {                                  // Beacon: src_panid, src_addr (:256-257)
        if (!take(2)) return;
        o.src_panid = (uint16_t)zb_le(psdu + pos, 2); pos += 2; o.present |= SNRX_ZBMAC_SRC_PANID;
        const int sl = addr_len(o.src_mode);
        if (sl) { if (!take(sl)) return; o.src_addr = zb_le(psdu + pos, sl); pos += sl; o.present |= SNRX_ZBMAC_SRC_ADDR; }
    } else {                                                         // Ack: no fields; types 4..7: raw payload
        o.payload_off = (uint8_t)pos;
        return;
    }
    // The auxiliary security header is NOT skipped: the reference's ConditionalField tests
    // `getfieldval("fcf_security") is True` (dot15d4.py:227-228, :265-266, :305-306) on a value that is the int 1, so its
    // dissector never parses that header and the MAC payload (and a command id) starts right behind the addresses.
    (void)security;
    if (o.frame_type == 3) {
        if (!take(1)) return;
        o.cmd_id = psdu[pos]; pos += 1;
    }
    o.payload_off = (uint8_t)pos;
    if (o.frame_type != 1) return;
    // ---- data frame payload: inter-PAN stub headers down to the ZLL commissioning command
    const uint8_t* p = psdu + pos;
    const int m = n - pos;
    if (m < 1 || (p[0] & 3) != 3) return;                            // Dot15d4Data.guess_payload_class: NWK frame type 0b11
    o.present |= SNRX_ZBMAC_INTERPAN;
    int q = 2;                                                       // ZigbeeNWKStub: 2 bytes
    if (m < q + 1) { o.present |= SNRX_ZBMAC_MALFORMED; return; }
    const unsigned aps = p[q]; q += 1;                               // ZigbeeAppDataPayloadStub frame control
    if (((aps >> 2) & 3) == 3) q += 2;                               // group address
    if (m < q + 4) { o.present |= SNRX_ZBMAC_MALFORMED; return; }
    o.cluster = (uint16_t)zb_le(p + q, 2);
    o.profile = (uint16_t)zb_le(p + q + 2, 2);
    q += 4;
    if ((aps & 3) != 3 || o.profile != 0xC05E || o.cluster != 0x1000) return;
    if (m < q + 1) { o.present |= SNRX_ZBMAC_MALFORMED; return; }
    const unsigned zcl = p[q]; q += 1;                               // ZigbeeZLLCommissioningCluster frame control
    if ((zcl >> 2) & 1) q += 2;                                      // manufacturer code
    if (m < q + 2) { o.present |= SNRX_ZBMAC_MALFORMED; return; }
    o.zll_command = p[q + 1];                                        // transaction sequence, then the command
    o.present |= SNRX_ZBMAC_ZLL;
    if (o.zll_command == 0x01) o.present |= SNRX_ZBMAC_ZLL_SCAN_RESPONSE;
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256) k_zb_mac_summary(const snrx_frame_t* __restrict__ frames, uint32_t n, snrx_zbmac_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const snrx_frame_t& f = frames[i];
    snrx_zbmac_t o;
    if (f.proto == SNRX_PROTO_ZIGBEE) {
        uint8_t b[132];
        const uint4* s = reinterpret_cast<const uint4*>(&f);       // the record is 16-byte aligned; bytes[] starts at offset 28
        uint32_t w[40];
#pragma unroll
        for (int k = 0; k < 10; k++) { const uint4 v = s[k]; w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
        for (int k = 0; k < 132; k++) b[k] = (uint8_t)(w[(28 + k) >> 2] >> (8 * ((28 + k) & 3)));
        zb_mac_parse(b, f.len, o);
    } else {
        zb_mac_parse(nullptr, 0, o);
        o.present = SNRX_ZBMAC_NOT_ZIGBEE;
    }
    o.frame = i;
    out[i] = o;
}
#endif

}  // namespace snrx
