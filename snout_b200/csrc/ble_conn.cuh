// ble_conn.cuh -- SURVEY 8(f) row N3: CONNECT_REQ records of a batch -> the parameters of the connections they open
// (parse_adv_pdu_payload_byte, vendor/BTLE/host/btle-tools/src/btle_rx.c:1476-1557; what receiver_controller :2167-2282
// starts tracking).  The second half of the row -- receiving the data channels with the learned access address -- reuses the
// back-end kernels of ble_back.cuh on the batch's bit streams (snrx_ble_follow in snrx.cu).
#pragma once
#include "common.cuh"

namespace snrx {

static_assert(sizeof(snrx_conn_t) == 56, "snrx_conn_t layout (include/snoutrx.h, _abi.CONN_DTYPE)");

// pdu = header(2) | payload | crc(3) as in snrx_frame_t.bytes.  Returns false unless this is a CONNECT_REQ with a 34-byte payload.
SNRX_HD bool ble_conn_parse(const uint8_t* pdu, int len, snrx_conn_t& o) {
    if (len != 2 + 34 + 3 || (pdu[0] & 0x0F) != 5 || (pdu[1] & 0x3F) != 34) return false;     // btle_rx.c:1477-1480
    const uint8_t* p = pdu + 2;
    for (int i = 0; i < 6; i++) { o.init_a[i] = p[i]; o.adv_a[i] = p[6 + i]; }
    o.access_addr = (uint32_t)p[12] | ((uint32_t)p[13] << 8) | ((uint32_t)p[14] << 16) | ((uint32_t)p[15] << 24);   // :1543-1546
    o.crc_init = ((uint32_t)p[16] << 16) | ((uint32_t)p[17] << 8) | (uint32_t)p[18];                                // :1505-1507
    o.win_size = p[19];
    o.win_offset = (uint16_t)(p[20] | (p[21] << 8));
    o.interval = (uint16_t)(p[22] | (p[23] << 8));
    o.latency = (uint16_t)(p[24] | (p[25] << 8));
    o.timeout = (uint16_t)(p[26] | (p[27] << 8));
    for (int i = 0; i < 5; i++) o.chm[i] = p[28 + i];
    o.hop = p[33] & 0x1F;
    o.sca = (p[33] >> 5) & 7;
    // chm_is_full_map :2158-2163 tests the REVERSED bytes {1F, FF, FF, FF, FF}: as transmitted that is FF FF FF FF 1F
    o.chm_full = (uint8_t)(p[28] == 0xFF && p[29] == 0xFF && p[30] == 0xFF && p[31] == 0xFF && p[32] == 0x1F);
    o.reserved[0] = o.reserved[1] = 0;
    return true;
}

#if defined(__CUDACC__)
// flags[i] = 1 where record i is a CRC-ok CONNECT_REQ (the scan of the flags gives the output order = record order)
__global__ void __launch_bounds__(256) k_ble_conn_flag(const snrx_frame_t* __restrict__ frames, uint32_t n, uint32_t* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const snrx_frame_t& f = frames[i];
    flags[i] = (f.proto == SNRX_PROTO_BLE && f.crc_ok && f.channel >= 37 && f.len == 39 && (f.bytes[0] & 0x0F) == 5 && (f.bytes[1] & 0x3F) == 34) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_ble_conn_fill(const snrx_frame_t* __restrict__ frames, uint32_t n, const uint32_t* __restrict__ flags,
                                                       const uint32_t* __restrict__ offsets, snrx_conn_t* __restrict__ out, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i] || offsets[i] >= cap) return;
    const snrx_frame_t& f = frames[i];
    uint8_t pdu[39];
    for (int k = 0; k < 39; k++) pdu[k] = f.bytes[k];
    snrx_conn_t c;
    if (!ble_conn_parse(pdu, f.len, c)) return;
    c.sample_index = f.sample_index; c.capture_id = f.capture_id; c.frame = i; c.channel = (uint8_t)f.channel;
    out[offsets[i]] = c;
}
#endif

}  // namespace snrx
