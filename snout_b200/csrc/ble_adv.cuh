// ble_adv.cuh -- SURVEY 8(f) row N1: advertising analytics on the GPU-decoded BLE records.
//
// What Snout does with every btle_rx line (snout/core/message.py:205-237 BtleMessage.fromraw ->
// snout/core/protocols/btle/advertising.py:113-307 BtlePDUPayload -> snout/core/device.py Device.get_unique):
// split the text, walk the AD structures of the payload, pick out flags / service data / manufacturer data
// (Apple Continuity TLVs, Microsoft), and keep one Device per sender address.  At 10^7..10^8 frames per second a
// Python object per line is the bottleneck, so the same two steps run on the decoded records in HBM:
//   k_ble_adv_summary   one thread per record: a fixed 32-byte snrx_adv_t (sender, PDU type, AD-structure fields)
//   k_ble_adv_devices   the summaries folded into an open-addressing hash table keyed by (AdvA, TxAdd):
//                       packet / CRC-ok counts, channel / PDU-type / AD / Apple-type masks, first / last position, and the
//                       company id of the LATEST packet (64-bit atomicMax on position | value): every field is folded
//                       with an order-independent atomic, so the table does not depend on thread or batch order
// Parsing rules follow the reference parser statement by statement (oracle/adv_oracle.py restates them in Python and is
// pinned against the imported reference, tests/golden/adv_ref.json); where the reference raises IndexError or never
// terminates on truncated input (advertising.py:104-110) the record is flagged SNRX_ADV_MALFORMED instead.
#pragma once
#include "common.cuh"

namespace snrx {

// PDU types whose payload is AdvA + AD structures (btle_rx.c print format "AdvA:.. Data:..", message.py:226-233)
SNRX_HD bool ble_pdu_has_adv_data(int pdu_type) { return pdu_type == 0 || pdu_type == 2 || pdu_type == 4 || pdu_type == 6; }

// AppleTypeParser.get_type_data (advertising.py:93-110) + parse_man_data_apple (:224-290): (type, len, data) TLVs
SNRX_HD void ble_adv_apple(const uint8_t* d, int n, snrx_adv_t& o) {
    o.apple_types = 0;
    o.apple_action = 0xFF;
    o.hints &= (uint8_t)~SNRX_HINT_NEARBY_MASK;                                   // the dict entry is replaced: last Apple AD wins
    int pos = 0;
    while (pos < n) {
        if (pos + 1 >= n) { o.present |= SNRX_ADV_MALFORMED; break; }           // the reference loops forever here
        const int t = d[pos], l = d[pos + 1];
        const uint8_t* v = d + pos + 2;
        int vl = n - (pos + 2); if (vl > l) vl = l;                               // slices truncate silently
        pos += 2 + l;
        if (t < 32) o.apple_types |= 1u << t;
        if (t == 0x10) {                                                          // Nearby: action code (:262-266)
            if (vl >= 1) o.apple_action = v[0] & 0x0F; else o.present |= SNRX_ADV_MALFORMED;
            if (vl >= 1 && !(o.hints & SNRX_HINT_NEARBY_MASK)) {                  // Device.os stops at the FIRST Nearby record
                int hint = 1;                                                     // nearby_data = apple_data[1:], :272-281
                if (vl - 1 == 1 && v[1] == 0x00) hint = 2;
                if (vl - 1 == 4) { if (v[1] == 0x10) hint = 3; if (v[1] == 0x18 || v[1] == 0x1C) hint = 4; }
                o.hints |= (uint8_t)hint;
            }
        }
        if (t == 0x0C && vl < 3) o.present |= SNRX_ADV_MALFORMED;                 // Handoff reads 3 bytes (:247-249): IndexError
    }
}

// `'ba5689a6fabfa2bd01467d6e00fbabad' in data.hex()` (device.py:187-190, 210-213): a substring test on the hex text, so the
// 32 hex digits may start at either nibble of a byte
SNRX_HD bool ble_adv_has_fitbit_uuid(const uint8_t* v, int n) {
    const uint8_t pat[16] = {0xba, 0x56, 0x89, 0xa6, 0xfa, 0xbf, 0xa2, 0xbd, 0x01, 0x46, 0x7d, 0x6e, 0x00, 0xfb, 0xab, 0xad};
    for (int s = 0; s + 32 <= 2 * n; s++) {                                       // s: first hex digit of the candidate match
        bool ok = true;
        for (int k = 0; k < 32 && ok; k++) {
            const int i = s + k;
            const int have = (i & 1) ? (v[i >> 1] & 0x0F) : (v[i >> 1] >> 4);
            const int want = (k & 1) ? (pat[k >> 1] & 0x0F) : (pat[k >> 1] >> 4);
            ok = have == want;
        }
        if (ok) return true;
    }
    return false;
}

// What one packet contributes to Device.vendor / .model / .os (device.py:171-246): 0 = the packet does not decide.
// Packets the reference parser raises on never become messages (MALFORMED), and only the AdvA + AD-structure PDU types do.
SNRX_HD uint32_t ble_adv_vendor_vote(const snrx_adv_t& a) {                       // 2^16 | company id, or 1 = FitBit
    if (a.present & SNRX_ADV_MALFORMED) return 0;
    if (a.present & SNRX_ADV_MANUFACTURER) return (1u << 16) | a.company_id;      // company_name is never empty ('??')
    if (a.hints & SNRX_HINT_FITBIT) return 1u;
    return 0;
}
// key of "first deciding packet": larger for earlier packets; vote < 2^17 (position at 2-sample resolution: 47 bits)
SNRX_HD unsigned long long ble_dev_first_key(unsigned long long pos, uint32_t vote) {
    return ((((1ull << 47) - 1ull) - (pos >> 1)) << 17) | vote;
}
SNRX_HD uint32_t ble_adv_model_vote(const snrx_adv_t& a) {
    if (a.present & SNRX_ADV_MALFORMED) return 0;
    if (a.hints & SNRX_HINT_FITBIT) return SNRX_MODEL_FITBIT_CHARGE;
    if ((a.present & SNRX_ADV_MANUFACTURER) && a.company_id == 0x004C && (a.apple_types & (1u << 7))) return SNRX_MODEL_AIRPODS;
    return 0;
}
SNRX_HD uint32_t ble_adv_os_vote(const snrx_adv_t& a) {
    if ((a.present & SNRX_ADV_MALFORMED) || !(a.present & SNRX_ADV_MANUFACTURER)) return 0;
    if (a.company_id == 0x004C && (a.hints & SNRX_HINT_NEARBY_MASK)) return (uint32_t)(a.hints & SNRX_HINT_NEARBY_MASK);   // 1..4 = SNRX_OS_UNDECIDED..IOS12
    if (a.company_id == 0x0006) return SNRX_OS_WINDOWS10;
    return 0;
}

// One BLE record -> summary.  pdu = header(2) | payload | crc(3) as in snrx_frame_t.bytes, len = valid bytes.
SNRX_HD void ble_adv_parse(const uint8_t* pdu, int len, snrx_adv_t& o) {
    for (int i = 0; i < 6; i++) o.adv_a[i] = 0;
    o.pdu_type = 0xFF; o.tx_add = 0; o.rx_add = 0; o.adv_len = 0; o.n_ad = 0; o.ad_flags = 0; o.present = 0;
    o.company_id = 0xFFFF; o.service_uuid = 0xFFFF; o.unknown_type = 0; o.apple_action = 0xFF; o.apple_types = 0;
    o.oob_flags = 0; o.hints = 0;
    if (len < 5) { o.present |= SNRX_ADV_MALFORMED; return; }
    o.pdu_type = pdu[0] & 0x0F;
    o.tx_add = (pdu[0] >> 6) & 1;
    o.rx_add = (pdu[0] >> 7) & 1;
    const int n = len - 5;                                                        // payload bytes present
    const uint8_t* p = pdu + 2;
    if (!ble_pdu_has_adv_data(o.pdu_type)) {
        // sender of the other advertising PDUs: ADV_DIRECT_IND / SCAN_REQ carry it first, CONNECT_REQ's AdvA second
        if (o.pdu_type == 5 && n >= 12) { for (int i = 0; i < 6; i++) o.adv_a[i] = p[6 + i]; o.present |= SNRX_ADV_SENDER; }
        else if ((o.pdu_type == 1 || o.pdu_type == 3) && n >= 6) { for (int i = 0; i < 6; i++) o.adv_a[i] = p[i]; o.present |= SNRX_ADV_SENDER; }
        return;
    }
    if (n < 6) { o.present |= SNRX_ADV_MALFORMED; return; }                       // btle_rx.c:1428-1432
    for (int i = 0; i < 6; i++) o.adv_a[i] = p[i];
    o.present |= SNRX_ADV_SENDER;
    const uint8_t* a = p + 6;
    const int an = n - 6;
    o.adv_len = (uint8_t)an;
    int pos = 0;
    while (pos < an) {                                                            // AdvDataParser.get_ad_structure :74-91
        const int l = a[pos];
        const uint8_t* s = a + pos + 1;
        int sl = an - (pos + 1); if (sl > l) sl = l;
        pos += 1 + l;
        o.n_ad++;
        if (sl <= 0) continue;                                                    // parse_ad_structure: `if data:`
        const int t = s[0];
        const uint8_t* v = s + 1;
        const int vl = sl - 1;
        if (t == 0x01) {                                                          // flags :161-176
            if (vl >= 1) { o.ad_flags = v[0]; o.present |= SNRX_ADV_FLAGS; } else o.present |= SNRX_ADV_MALFORMED;
        } else if (t == 0x06) {
            o.present |= SNRX_ADV_UUID128;
            o.hints &= (uint8_t)~SNRX_HINT_FITBIT;                                // last AD 0x06 wins
            if (ble_adv_has_fitbit_uuid(v, vl)) o.hints |= SNRX_HINT_FITBIT;
        } else if (t == 0x11) {                                                   // security manager OOB flags :183-201
            if (vl >= 1) { o.oob_flags = v[0]; o.present |= SNRX_ADV_OOB; } else o.present |= SNRX_ADV_MALFORMED;
        } else if (t == 0x16) {                                                   // service data :203-210
            if (vl >= 2) { o.service_uuid = (uint16_t)(v[0] | (v[1] << 8)); o.present |= SNRX_ADV_SERVICE_DATA; }
            else o.present |= SNRX_ADV_MALFORMED;
        } else if (t == 0xFF) {                                                   // manufacturer specific :212-232
            if (vl >= 2) {
                o.company_id = (uint16_t)(v[0] | (v[1] << 8));
                o.present |= SNRX_ADV_MANUFACTURER;
                if (o.company_id == 0x004C) ble_adv_apple(v + 2, vl - 2, o);
            } else o.present |= SNRX_ADV_MALFORMED;
        } else {
            o.unknown_type = (uint8_t)t;
            o.present |= SNRX_ADV_UNKNOWN;
        }
    }
}

SNRX_HD uint64_t ble_dev_key(const snrx_adv_t& a) {
    uint64_t k = 0;
    for (int i = 0; i < 6; i++) k |= (uint64_t)a.adv_a[i] << (8 * i);
    return k | ((uint64_t)a.tx_add << 48) | (1ull << 63);                        // bit 63: slot in use
}
SNRX_HD uint32_t ble_dev_hash(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}

// device-side accumulator of one sender: every field is folded with an order-independent atomic, so the table does not
// depend on thread or batch order ("latest company id" = atomicMax on position | value)
struct DevSlot {
    unsigned long long key;           // 0 = empty
    unsigned long long chan_mask;     // bit c: seen on BLE channel c
    unsigned long long first_inv;     // max over packets of (2^63 - pos): zero-initialised "min"
    unsigned long long last_pos;      // max over packets of pos
    unsigned long long last_company;  // max over packets with manufacturer data of (pos << 16 | company id)
    unsigned int packets, crc_ok;
    unsigned int pdu_mask, apple_types;
    unsigned int present, ad_flags;   // OR over packets
    // FIRST deciding packet of Device.vendor / .model / .os: max over deciding packets of ble_dev_first_key(pos, vote)
    unsigned long long first_vendor, first_model, first_os;
    unsigned long long pad[5];
};
static_assert(sizeof(DevSlot) == 128, "one slot per 128-byte line");

// order of appearance inside a job: capture, then channel-rate position (sample_index >= -4)
SNRX_HD unsigned long long ble_dev_pos(uint32_t capture_id, int64_t sample_index) {
    return ((unsigned long long)(capture_id & 0xFFFF) << 32) | (unsigned long long)((sample_index + 16) & 0xFFFFFFFFll);
}

SNRX_HD void ble_dev_export(const DevSlot& d, snrx_device_t& o) {
    for (int i = 0; i < 6; i++) o.adv_a[i] = (uint8_t)(d.key >> (8 * i));
    o.tx_add = (uint8_t)((d.key >> 48) & 1);
    o.ad_flags = (uint8_t)d.ad_flags;
    o.packets = d.packets; o.crc_ok = d.crc_ok;
    o.chan_mask = d.chan_mask;
    const unsigned long long first = (1ull << 63) - d.first_inv;
    o.first_capture = (uint32_t)(first >> 32); o.first_index = (int64_t)(first & 0xFFFFFFFFull) - 16;
    o.last_capture = (uint32_t)(d.last_pos >> 32); o.last_index = (int64_t)(d.last_pos & 0xFFFFFFFFull) - 16;
    o.pdu_mask = (uint16_t)d.pdu_mask;
    o.present = (uint16_t)d.present;
    o.company_id = d.last_company ? (uint16_t)(d.last_company & 0xFFFF) : 0xFFFF;
    o.vendor_company = (uint16_t)(d.first_vendor & 0xFFFF);
    o.vendor_kind = (uint8_t)(((d.first_vendor >> 16) & 1u) ? SNRX_VENDOR_COMPANY : d.first_vendor ? SNRX_VENDOR_FITBIT : SNRX_VENDOR_NONE);
    if (o.vendor_kind != SNRX_VENDOR_COMPANY) o.vendor_company = 0xFFFF;
    o.model = (uint8_t)(d.first_model & 0xFF);
    o.os = (uint8_t)(d.first_os & 0xFF);
    o.pad = 0;
    o.apple_types = d.apple_types;
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256) k_ble_adv_summary(const snrx_frame_t* __restrict__ frames, uint32_t n,
                                                         snrx_adv_t* __restrict__ out, uint32_t* __restrict__ n_out) {
    // BLE records come first in a batch (mixed mode appends the 802.15.4 records after them), so the summary index is
    // the record index; non-BLE records get pdu_type 0xFF and are not counted
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const snrx_frame_t& f = frames[i];
    snrx_adv_t o;
    if (f.proto == SNRX_PROTO_BLE) {
        uint8_t pdu[48];
        const int len = f.len < 47 ? f.len : 47;
        for (int k = 0; k < len; k++) pdu[k] = f.bytes[k];
        ble_adv_parse(pdu, len, o);
        atomicAdd(n_out, 1u);
    } else {
        ble_adv_parse(nullptr, 0, o);
        o.present = 0;                      // not a BLE record: pdu_type 0xff, nothing else
    }
    o.frame = i;
    out[i] = o;
}

__global__ void __launch_bounds__(256) k_ble_adv_devices(const snrx_frame_t* __restrict__ frames, const snrx_adv_t* __restrict__ adv,
                                                         uint32_t n, DevSlot* __restrict__ table, uint32_t table_mask,
                                                         uint32_t* __restrict__ counters /* [0] new devices, [1] dropped: table full */) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const snrx_adv_t a = adv[i];
    if (!(a.present & SNRX_ADV_SENDER)) return;
    const snrx_frame_t& f = frames[i];
    const unsigned long long key = ble_dev_key(a);
    uint32_t s = ble_dev_hash(key) & table_mask;
    for (uint32_t probe = 0; probe <= table_mask; probe++, s = (s + 1) & table_mask) {
        const unsigned long long prev = atomicCAS(&table[s].key, 0ull, key);
        if (prev == 0ull || prev == key) {
            DevSlot& d = table[s];
            if (prev == 0ull) atomicAdd(&counters[0], 1u);
            const unsigned long long pos = ble_dev_pos(f.capture_id, f.sample_index);
            atomicAdd(&d.packets, 1u);
            if (f.crc_ok) atomicAdd(&d.crc_ok, 1u);
            atomicOr(&d.chan_mask, 1ull << (f.channel & 63));
            atomicOr(&d.pdu_mask, 1u << a.pdu_type);
            atomicOr(&d.present, (unsigned int)a.present);
            atomicOr(&d.ad_flags, (unsigned int)a.ad_flags);
            atomicOr(&d.apple_types, a.apple_types);
            atomicMax(&d.first_inv, (1ull << 63) - pos);
            atomicMax(&d.last_pos, pos);
            if (a.present & SNRX_ADV_MANUFACTURER) atomicMax(&d.last_company, (pos << 16) | a.company_id);
            if (f.crc_ok) {                                       // only CRC0 lines become messages (message.py:225-226)
                if (const uint32_t v = ble_adv_vendor_vote(a)) atomicMax(&d.first_vendor, ble_dev_first_key(pos, v));
                if (const uint32_t v = ble_adv_model_vote(a)) atomicMax(&d.first_model, ble_dev_first_key(pos, v));
                if (const uint32_t v = ble_adv_os_vote(a)) atomicMax(&d.first_os, ble_dev_first_key(pos, v));
            }
            return;
        }
    }
    atomicAdd(&counters[1], 1u);
}

// compaction of the used slots (order = slot order; the host sorts by address)
__global__ void __launch_bounds__(256) k_ble_adv_export(const DevSlot* __restrict__ table, uint32_t slots, snrx_device_t* __restrict__ out,
                                                        uint32_t cap, uint32_t* __restrict__ n_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slots || table[i].key == 0ull) return;
    const uint32_t k = atomicAdd(n_out, 1u);
    if (k < cap) ble_dev_export(table[i], out[k]);
}
#endif  // __CUDACC__

}  // namespace snrx
