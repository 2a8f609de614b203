/*
 * snoutrx.h -- C ABI of libsnoutrx.so, the B200-native IQ -> packet receive engine
 * that replaces Snout's two receive engines behind their own boundaries.
 *
 * Plain C: pointers and sizes only, no torch / C++ types.  Every entry point
 * returns 0 on success or a negative SNRX_E* code (snrx_strerror() explains it).
 * There is no CPU fallback anywhere behind this interface: if no CUDA device
 * is usable snrx_create() fails with SNRX_ENODEV.
 *
 * Which reference interface each entry point replaces (paths relative to the
 * nislab/snout tree):
 *
 *   snrx_create / snrx_destroy
 *       BLE   : process start of `btle_rx -c <ch> -g 6 -a 8e89bed6 -k 555555`
 *               (snout/util/btle.py:53-69; option table
 *               vendor/BTLE/host/btle-tools/src/btle_rx.c:1209-1294; board
 *               set-up config_run_board() btle_rx.c:623; crc_init_reorder()
 *               btle_rx.c:1801-1825 applied once at start, btle_rx.c:2335).
 *       Zigbee: construction of the flowgraph `top_block(channel=...)`
 *               (snout/modulations/Zigbee/hackrf/Zigbee_rx/top_block.py:29-89).
 *   snrx_process
 *       BLE   : rx_callback() filling the ring (btle_rx.c:489-498) + the
 *               half-buffer loop calling receiver() (btle_rx.c:2341-2393,
 *               receiver() 2020-2155, search_unique_bits() 1369-1421,
 *               demod_byte() 1348-1367, scramble_byte() 1158-1163,
 *               crc_check() 1826-1848).
 *       Zigbee: the stream edges source -> quadrature_demod_cf -> (x - iir(x))
 *               -> clock_recovery_mm_ff -> packet_sink (top_block.py:52-73,
 *               80-89) and the sink's general_work()
 *               (scapy-radio/gnuradio/gr-zigbee/lib/packet_sink_scapy_impl.cc:158-374).
 *       Wideband modes additionally run the 96 Msps polyphase channelizer,
 *       which has no counterpart in the reference (it retunes one channel at
 *       a time: snout/util/btle.py:62, snout/core/radio.py:415).
 *   snrx_process_sc8
 *       the same with the reference's native sample format: int8 I,Q as delivered by
 *       hackrf_transfer.buffer to rx_callback() (btle_rx.c:489-498, IQ_TYPE btle_rx.c:204).
 *   snrx_poll
 *       BLE   : the printf()/fflush() of one line per frame (btle_rx.c:2137-2145,
 *               2383) that snout/core/pcontroller.py:115-131 reads back.
 *       Zigbee: message_port_pub() of the PSDU blob
 *               (packet_sink_scapy_impl.cc:333-349) -> UDP 127.0.0.1:52002
 *               (top_block.py:71) -> GnuradioSocket.recv
 *               (scapy-radio/scapy/scapy/modules/gnuradio.py:67-73).
 *   snrx_set_channel
 *       Zigbee: top_block.set_channel() via XMLRPC (top_block.py:94-96).
 *   snrx_debug_stage
 *       no reference counterpart: exposes intermediate streams (quantised
 *       channel streams, discriminator output, soft chips) for parity tests.
 */
#ifndef SNOUTRX_H
#define SNOUTRX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNRX_ABI_VERSION 2

/* protocol ids = scapy-radio GnuradioPacket.proto values
 * (scapy-radio/scapy/scapy/layers/gnuradio.py:19-25) */
#define SNRX_PROTO_ZIGBEE 2
#define SNRX_PROTO_BLE    3

/* error codes */
#define SNRX_OK        0
#define SNRX_EINVAL   -1   /* bad argument / configuration            */
#define SNRX_ENODEV   -2   /* no usable CUDA device                   */
#define SNRX_ENOMEM   -3   /* host or device allocation failed        */
#define SNRX_ECUDA    -4   /* CUDA runtime error (see snrx_last_error) */
#define SNRX_ERANGE   -5   /* input larger than the configured capacity */
#define SNRX_EOVERFLOW -6  /* more candidates / frames than capacity  */
#define SNRX_ESTATE   -7   /* call out of sequence                    */

/* receive modes */
#define SNRX_MODE_BLE_NB     0  /* one BLE channel, 4 Msps cf32 in          (BASELINE config 1) */
#define SNRX_MODE_ZB_NB      1  /* one 802.15.4 channel, 4 Msps cf32 in     (config 2)          */
#define SNRX_MODE_ZB_WB16    2  /* 96 Msps in -> 16 Zigbee receivers        (config 3)          */
#define SNRX_MODE_BLE_WB40   3  /* 96 Msps in -> 40 BLE receivers           (config 4)          */
#define SNRX_MODE_MIXED_WB56 4  /* 96 Msps in -> 40 BLE + 16 Zigbee         (config 5)          */

/* geometry constants of the wideband channelizer (DESIGN.md "Channelizer") */
#define SNRX_WB_RATE      96000000
#define SNRX_WB_CENTER    2440000000ull
#define SNRX_WB_DECIM     24        /* 96 Msps -> 4 Msps per channel          */
#define SNRX_NB_RATE      4000000
#define SNRX_BLE_WINDOW   8192      /* btle_rx half buffer in IQ samples (btle_rx.c:180-182) */

/* One decoded frame.  Fixed 160 bytes so it can be all-gathered as-is. */
typedef struct snrx_frame {
    int64_t  sample_index;  /* channel-rate (4 Msps) IQ index, capture relative:
                               BLE: sample carrying access-address bit 0 (may be
                               -4..-1 at a window origin, btle_rx.c:1409);
                               Zigbee: input position of the chip completing the SFD */
    uint32_t capture_id;    /* index of the capture inside the processed batch */
    uint32_t window;        /* BLE: 8192-IQ window that reported it; Zigbee: segment */
    uint16_t channel;       /* BLE 0..39 / 802.15.4 11..26                    */
    uint8_t  proto;         /* SNRX_PROTO_*                                   */
    uint8_t  crc_ok;        /* BLE: CRC-24 matches (btle_rx prints CRC0); Zigbee: FCS-16 matches */
    uint8_t  lqi;           /* Zigbee: min(255,(sum/8)<<3) (packet_sink_scapy_impl.cc:334-335) */
    uint8_t  phase;         /* BLE: sample phase 0..3 of the hit                */
    uint16_t len;           /* valid bytes in bytes[]: BLE 2+payload+3, Zigbee PSDU length */
    uint32_t access_addr;   /* BLE                                            */
    uint8_t  bytes[132];    /* BLE: header|payload|crc (de-whitened); Zigbee: PSDU incl. FCS */
} snrx_frame_t;

/* Default Zigbee chain geometry: one clock-recovery + packet-sink chain per 4096 channel-rate samples
 * (1 ms), started 2048 samples early.  Part of the parity contract: the oracle runs the same segments.
 * (Measured on the oracle, tests/test_oracle_zigbee.py: how long the warm-up is does not change how close the
 * segmented receiver is to one unsegmented chain -- 2048 .. 131072 samples give the same frame agreement.) */
#define SNRX_ZB_SEGMENT_DEFAULT 4096
#define SNRX_ZB_PREHALO_DEFAULT 2048
/* Blocked DC tracker (a11): block length and memory in blocks; a time shard needs
 * (SNRX_ZB_IIR_MEMORY_BLOCKS + 1) * SNRX_ZB_IIR_BLOCK + zb_prehalo channel samples of pre halo. */
#define SNRX_ZB_IIR_BLOCK 2048
#define SNRX_ZB_IIR_MEMORY_BLOCKS 48
/* Span rule (also part of the contract): a CRC-failed 802.15.4 record whose sample_index lies inside the span of an
 * earlier CRC-ok record of the same (capture, channel) -- (2 + 2 * len) * 64 samples after its sample_index -- is not
 * reported: the reference's sequential packet sink is busy with that frame (packet_sink_scapy_impl.cc:247-359) and only
 * a chain restarted inside it can lock onto its payload chips.  Applied inside every batch; a caller that concatenates
 * time shards applies it once more across the shard boundary (snout_b200/stream.py zb_span_filter). */

typedef struct snrx_config {
    uint32_t abi_version;    /* must be SNRX_ABI_VERSION                                   */
    int32_t  device;         /* CUDA device ordinal                                        */
    int32_t  mode;           /* SNRX_MODE_*                                                */
    int32_t  channel;        /* NB modes: BLE 0..39 / Zigbee 11..26 ; ignored in WB modes  */
    uint32_t access_addr;    /* BLE `-a`  (default 0x8E89BED6, btle_rx.c:188)              */
    uint32_t crc_init;       /* BLE `-k`  as typed (0x555555, btle_rx.c:189)               */
    int32_t  zb_threshold;   /* packet sink threshold (10, top_block.py:67)                */
    float    quant_scale;    /* BLE: q = clamp(rint(x*quant_scale), -128, 127)             */
    uint64_t max_samples;    /* capacity: input samples per capture                        */
    uint32_t max_captures;   /* capacity: captures per snrx_process batch (wideband modes: <= 65535) */
    uint32_t max_frames;     /* capacity: frames per snrx_process batch                    */
    uint32_t zb_segment;     /* Zigbee: chain segment body, channel-rate samples (0 = SNRX_ZB_SEGMENT_DEFAULT) */
    uint32_t zb_prehalo;     /* Zigbee: chain warm-up, channel-rate samples (0 = SNRX_ZB_PREHALO_DEFAULT)   */
    uint32_t pfb_taps;       /* WB: prototype length, 384 or 768 (0 = default 384)         */
    uint32_t flags;          /* SNRX_F_*                                                   */
    uint32_t access_mask;    /* BLE `-m`: access-address bits that take part in the match
                                (0 = default 0xFFFFFFFF, btle_rx.c:1203,2301)                */
    uint32_t reserved;       /* must be 0                                                  */
} snrx_config_t;

#define SNRX_F_KEEP_STREAMS 1u  /* also store channel streams so snrx_debug_stage can return them */

typedef struct snrx_stats {
    uint64_t samples_in;      /* input-rate samples consumed by the last snrx_process */
    uint64_t channel_samples; /* channel-rate samples produced (all channels)         */
    uint32_t candidates;      /* BLE access-address hits / Zigbee chains run          */
    uint32_t frames;          /* frames emitted                                       */
    uint32_t frames_crc_ok;
    uint32_t kernel_launches; /* kernels launched by the last snrx_process            */
    float    gpu_ms;          /* device time of the last snrx_process (CUDA events)   */
    float    gpu_ms_frontend; /* of which: channelizer / slicer kernel                */
} snrx_stats_t;

/* debug stages for snrx_debug_stage() */
#define SNRX_STAGE_BLE_Q8      1  /* int8 I,Q quantised channel streams [cap][ch][n][2]      */
#define SNRX_STAGE_BLE_BITS    2  /* uint32 slicer bit words [cap][ch][words], sample order  */
#define SNRX_STAGE_CHAN_CF32   3  /* cf32 channel streams [cap][ch][n] (WB modes)             */
#define SNRX_STAGE_ZB_DISC     4  /* f32 discriminator minus DC [cap][ch][stride] (stride = bytes/(4*cap*ch)) */
#define SNRX_STAGE_ZB_CHIPS    5  /* f32 soft chips [chain][chips_cap], chain = (cap, ch, segment)  */
#define SNRX_STAGE_ZB_F        6  /* f32 discriminator output before DC removal [cap][ch][stride] */
#define SNRX_STAGE_ZB_NCHIPS   7  /* int64 number of chips each chain produced [chain]            */

typedef struct snrx_handle snrx_t;

int  snrx_abi_version(void);
const char* snrx_strerror(int code);
const char* snrx_last_error(snrx_t* h);          /* detail text of the last failure (may be "") */

int  snrx_device_count(int* n);
int  snrx_create(snrx_t** h, const snrx_config_t* cfg);
void snrx_destroy(snrx_t* h);

/* Time shard of a longer capture (multi-GPU / streaming): the buffer handed to snrx_process
 * starts `pre_samples` before the shard body (read-only halo: channelizer history, Zigbee
 * warm-up) and may extend past it (post halo, so frames that start in the body can be
 * decoded completely).  Only frames anchored inside the body are reported, with the
 * sample_index / window they have in the whole capture, so the union over shards equals
 * the result of one call on the whole capture, bit for bit.  Requirements (channel rate):
 *   BLE    pre halo multiple of 128 (>= 128 unless the shard starts the capture), post halo
 *          >= 2048, body on the 8192-sample window grid;
 *   Zigbee pre halo multiple of 2048 and >= 100352 + zb_prehalo unless the shard starts the
 *          capture (the DC tracker remembers 48 blocks of 2048 and the first block of a buffer
 *          starts without discriminator / channelizer history), post halo >= 16448, body on
 *          the zb_segment grid; zb_segment and zb_prehalo multiples of 2048, zb_segment a divisor
 *          or a multiple of 8192. */
typedef struct snrx_shard {
    uint64_t pre_samples;      /* input-rate samples of pre halo                                   */
    uint64_t body_samples;     /* input-rate samples of the body (0 = rest of the buffer)          */
    uint32_t first_window;     /* index inside the capture of the body's first 8192-sample window
                                  (channel rate); Zigbee segment = first_window * 8192 / zb_segment */
    uint32_t first_capture_id; /* capture_id reported for capture 0 of the batch                   */
} snrx_shard_t;

/* Run the receive path over a batch of `n_captures` captures, each `n_samples`
 * interleaved cf32 IQ samples long and `stride_samples` apart (input rate; 0 = n_samples).
 * `iq` is a host pointer (pageable or pinned; copied in chunks overlapped with
 * compute) or, if is_device_ptr != 0, a device pointer on cfg.device; 16-byte aligned.
 * `shard` is NULL for whole captures.  Asynchronous with respect to the host when
 * the input is a device pointer; results are collected with snrx_poll. */
int  snrx_process(snrx_t* h, const float* iq, uint32_t n_captures,
                  uint64_t n_samples, uint64_t stride_samples,
                  const snrx_shard_t* shard, int is_device_ptr);

/* snrx_process for captures stored as interleaved signed 8-bit I,Q -- the transfer format of the HackRF the
 * reference receives from (IQ_TYPE int8_t, btle_rx.c:204; rx_callback() copies exactly these bytes into the ring,
 * btle_rx.c:489-498).  Sample value x = q / 128 (exact), so with the narrow-band default quant_scale = 128 the
 * BLE path sees the reference's own int8 samples.  A quarter of the host->device bytes of cf32.  `iq` 16-byte
 * aligned, even stride for batches; sizes and strides count complex samples as in snrx_process. */
int  snrx_process_sc8(snrx_t* h, const int8_t* iq, uint32_t n_captures,
                      uint64_t n_samples, uint64_t stride_samples,
                      const snrx_shard_t* shard, int is_device_ptr);

/* Up to two batches may be queued (snrx_process, snrx_process, snrx_poll, ...): results are
 * collected oldest first.  snrx_poll waits for the oldest queued batch; *n_out = its number of
 * frames, in reference order (capture, channel, window, sample_index).  With out == NULL it only
 * reports the count; with out != NULL it copies up to `cap` frames and retires the batch.
 * snrx_poll_view retires the batch and returns a pointer to the engine's pinned host copy of the
 * frames (written by the kernels directly, no extra copy), valid until the second-next
 * snrx_process. */
int  snrx_poll(snrx_t* h, snrx_frame_t* out, uint32_t cap, uint32_t* n_out);
int  snrx_poll_view(snrx_t* h, const snrx_frame_t** frames, uint32_t* n_out);

/* Device-memory frame list / totals of the most recent snrx_process (valid once that batch has been
 * polled; the kernels write the list in HBM and k_export_frames copies it to the pinned host buffer). */
int  snrx_frames_device(snrx_t* h, void** frames_dev, void** count_dev);

/* Device-memory list of the frames of the batch most recently retired by snrx_poll / snrx_poll_view and
 * their count: the send buffer of the multi-GPU frame all-gather (no host round trip, no copy).
 * The buffer holds max_frames records (those past the count are unspecified); the 160 bytes in front of it hold
 * {uint64 count, uint64 batch number (0, 1, ... per handle)}, so [header | records] can be sent as one block.  It stays
 * valid for two further snrx_process calls: it is overwritten by the third (each of the two lanes alternates between two lists). */
int  snrx_polled_frames_device(snrx_t* h, const snrx_frame_t** frames_dev, uint32_t* n_out);

/* ---- advertising analytics on the decoded BLE records (what Snout does per btle_rx line:
 * BtleMessage.fromraw message.py:205-237 -> BtlePDUPayload advertising.py:113-307 -> Device device.py) ---- */
#define SNRX_ADV_FLAGS         0x0001  /* AD 0x01 present (ad_flags)                                   */
#define SNRX_ADV_UUID128       0x0002  /* AD 0x06                                                      */
#define SNRX_ADV_OOB           0x0004  /* AD 0x11 (oob_flags)                                          */
#define SNRX_ADV_SERVICE_DATA  0x0008  /* AD 0x16 (service_uuid)                                       */
#define SNRX_ADV_MANUFACTURER  0x0010  /* AD 0xff (company_id; Apple: apple_types / apple_action)      */
#define SNRX_ADV_UNKNOWN       0x0020  /* some other AD type (unknown_type = the last one)             */
#define SNRX_ADV_MALFORMED     0x0040  /* truncated where the reference parser raises / never returns  */
#define SNRX_ADV_SENDER        0x0080  /* adv_a holds the sender address                               */

/* snrx_adv_t.hints.  Low nibble: the first Nearby TLV (Apple type 0x10) of the record's Apple data -- 0 none, 1 present
 * without a version hint, 2 / 3 / 4 = the reference's 'iOS Version Hint' '10' / '11' / '12' (advertising.py:267-281). */
#define SNRX_HINT_NEARBY_MASK  0x0f
#define SNRX_HINT_FITBIT       0x10    /* the AD 0x06 field contains ba5689a6fabfa2bd01467d6e00fbabad (device.py:187-190)  */
/* snrx_device_t.model / .os codes */
#define SNRX_MODEL_NONE 0
#define SNRX_MODEL_FITBIT_CHARGE 1     /* 'Charge / Charge HR' (device.py:210-213)                             */
#define SNRX_MODEL_AIRPODS 2           /* Apple record of type 0x07 (device.py:214-219)                        */
#define SNRX_OS_NONE 0
#define SNRX_OS_UNDECIDED 1            /* a Nearby record without a version hint: the reference answers '-'    */
#define SNRX_OS_IOS10 2
#define SNRX_OS_IOS11 3
#define SNRX_OS_IOS12 4
#define SNRX_OS_WINDOWS10 5            /* company id 0x0006 (device.py:243-245)                                */
#define SNRX_VENDOR_NONE    0
#define SNRX_VENDOR_COMPANY 1          /* vendor_company holds the company id whose name Device.vendor returns */
#define SNRX_VENDOR_FITBIT  2          /* no company id in the first deciding packet, but the FitBit UUID: "FitBit" */

typedef struct snrx_adv {           /* one per record of the batch, 32 bytes */
    uint8_t  adv_a[6];      /* sender address as transmitted (btle_rx prints it reversed, btle_rx.c:1434-1441) */
    uint8_t  pdu_type;      /* header & 0x0f; 0xff: not a BLE record                                  */
    uint8_t  tx_add, rx_add;
    uint8_t  adv_len;       /* bytes of AD structures (PDU types 0, 2, 4, 6)                           */
    uint8_t  n_ad;          /* AD structures walked (AdvDataParser.get_ad_structure :74-91)            */
    uint8_t  ad_flags;      /* AD 0x01 value (last one wins, as in the reference's dict)               */
    uint16_t present;       /* SNRX_ADV_*                                                              */
    uint16_t company_id;    /* AD 0xff company (word16be() of the reference is little endian), 0xffff none */
    uint16_t service_uuid;  /* AD 0x16, 0xffff none                                                    */
    uint8_t  unknown_type;
    uint8_t  apple_action;  /* Apple Nearby (type 0x10) action code, 0xff none                         */
    uint8_t  oob_flags;     /* AD 0x11 value                                                           */
    uint8_t  hints;         /* SNRX_HINT_*: what Device.vendor / .model / .os look at (device.py:171-265)      */
    uint32_t apple_types;   /* bit t: Apple Continuity TLV type t (< 32) present (last Apple AD wins)   */
    uint32_t frame;         /* index of the record in the batch                                        */
} snrx_adv_t;

typedef struct snrx_device {        /* one per sender (AdvA, TxAdd), 64 bytes */
    uint8_t  adv_a[6];
    uint8_t  tx_add;
    uint8_t  ad_flags;      /* OR over its packets                                                     */
    uint32_t packets, crc_ok;
    uint64_t chan_mask;     /* bit c: seen on BLE channel c                                            */
    int64_t  first_index, last_index;     /* channel-rate position of its first / last packet (kept modulo 2^32 samples and
                                             2^16 captures inside the table: a stream longer than 17.9 minutes at 4 Msps
                                             should advance capture_id -- snrx_shard_t.first_capture_id -- per 2^32 samples) */
    uint32_t first_capture, last_capture; /* ... and their capture ids                                 */
    uint16_t pdu_mask;      /* bit t: PDU type t seen                                                  */
    uint16_t present;       /* OR of SNRX_ADV_* over its packets                                       */
    uint16_t company_id;    /* of its latest packet with manufacturer data, 0xffff none                */
    uint16_t vendor_company;/* Device.vendor (device.py:171-191): company id of its FIRST packet that carries one
                               (valid when vendor_kind == SNRX_VENDOR_COMPANY)                              */
    uint32_t apple_types;   /* OR over its packets                                                     */
    uint8_t  model;         /* Device.model (device.py:193-220): SNRX_MODEL_* decided by its first deciding packet */
    uint8_t  os;            /* Device.os (device.py:222-246): SNRX_OS_* decided by its first deciding packet   */
    uint8_t  vendor_kind;   /* SNRX_VENDOR_*                                                           */
    uint8_t  pad;
} snrx_device_t;

/* Summaries of the batch most recently retired by snrx_poll / snrx_poll_view (computed on the GPU from the device frame
 * list; call before the second-next snrx_process).  out may be NULL (only folds the batch into the device table).
 * *n_out = records of the batch (BLE and others; others have pdu_type 0xff). */
int  snrx_ble_adv_summary(snrx_t* h, snrx_adv_t* out, uint32_t cap, uint32_t* n_out);
/* The sender table accumulated by every snrx_ble_adv_summary call so far (unordered); reset != 0 clears it afterwards.
 * SNRX_EOVERFLOW if more than 2^18 distinct senders were seen (the extra ones were not recorded). */
int  snrx_ble_devices(snrx_t* h, snrx_device_t* out, uint32_t cap, uint32_t* n_out, int reset);

/* ---- BLE connections.  btle_rx -o follows ONE connection by retuning its single channel on a wall-clock
 * schedule (receiver_controller, btle_rx.c:2167-2282: on a CRC-ok CONNECT_REQ it takes AA, CRCInit, ChM, Hop, Interval from
 * the fields parsed at btle_rx.c:1476-1557, hops hop_chan = (hop_chan + hop) % 37 and receives with the new access address
 * and crc_init_reorder(CRCInit)).  With all 40 channels channelized at once nothing has to be retuned or timed: the
 * CONNECT_REQs are picked out of the decoded records, and the data channels of the batch -- whose slicer bit streams are
 * still in HBM -- are searched and decoded AGAIN with the learned access address and CRC init. ---- */
typedef struct snrx_conn {          /* one per CRC-ok CONNECT_REQ (ADV PDU type 5, payload 34 bytes), 56 bytes */
    int64_t  sample_index;  /* of the CONNECT_REQ record                                                       */
    uint32_t capture_id;
    uint32_t frame;         /* index of the record in the batch                                                */
    uint32_t access_addr;   /* AA of the connection (payload bytes 12..15, little endian; btle_rx.c:1543-1546)  */
    uint32_t crc_init;      /* CRCInit as btle_rx.c:1505-1507 assembles it (byte 16 is the most significant):
                               the value to pass as crc_init (`-k`) for this connection                         */
    uint8_t  init_a[6], adv_a[6];   /* as transmitted (the reference stores them reversed for printing)          */
    uint16_t win_offset, interval, latency, timeout;
    uint8_t  chm[5];        /* channel map as transmitted (payload bytes 28..32)                               */
    uint8_t  win_size, hop, sca;
    uint8_t  channel;       /* advertising channel the request was heard on                                    */
    uint8_t  chm_full;      /* all 37 data channels used (the only maps btle_rx -o follows, btle_rx.c:2158-2163) */
    uint8_t  reserved[2];
} snrx_conn_t;

/* CONNECT_REQs (CRC ok) among the records of the batch most recently retired by snrx_poll / snrx_poll_view, in record order. */
int  snrx_ble_connections(snrx_t* h, snrx_conn_t* out, uint32_t cap, uint32_t* n_out);
/* Searches and decodes the BLE channels of that same batch again with another access address / CRC init (`-a` / `-k` of a
 * connection learned from snrx_ble_connections): *n_out frames in reference order, copied to out (up to cap).  The batch's
 * slicer bit streams are reused: nothing is channelized twice.  Call before the second-next snrx_process. */
int  snrx_ble_follow(snrx_t* h, uint32_t access_addr, uint32_t crc_init, snrx_frame_t* out, uint32_t cap, uint32_t* n_out);

/* ---- the Zigbee consumer path on the decoded 802.15.4 records (what Snout does per datagram:
 * RFtap(pkt) -> Dot15d4FCS dissection, snout/util/zigbee.py:194-202; ZigbeeMessage.fromraw, snout/core/message.py:258-304;
 * the touchlink scan's haslayer(ZLLScanResponse), zigbee.py:176-192) ---- */
#define SNRX_ZBMAC_SECURITY           0x0001  /* FCF security enabled                                        */
#define SNRX_ZBMAC_PENDING            0x0002  /* FCF frame pending                                           */
#define SNRX_ZBMAC_ACKREQ             0x0004  /* FCF acknowledgement request                                 */
#define SNRX_ZBMAC_PANID_COMPRESS     0x0008  /* FCF PAN id compression                                      */
#define SNRX_ZBMAC_DEST_PANID         0x0010  /* dest_panid holds a value                                    */
#define SNRX_ZBMAC_DEST_ADDR          0x0020  /* dest_addr holds a value (dest_mode 2: 16 bit, 3: 64 bit)    */
#define SNRX_ZBMAC_SRC_PANID          0x0040
#define SNRX_ZBMAC_SRC_ADDR           0x0080
#define SNRX_ZBMAC_INTERPAN           0x0100  /* data frame whose payload is a ZigBee inter-PAN stub NWK frame */
#define SNRX_ZBMAC_ZLL                0x0200  /* ... carrying a ZLL commissioning cluster command (zll_command) */
#define SNRX_ZBMAC_ZLL_SCAN_RESPONSE  0x0400  /* ... which is a scan response (what the touchlink scan collects) */
#define SNRX_ZBMAC_NO_ADDRESSING      0x0800  /* an address mode without a length (none / reserved): the reference dissector
                                                 raises (dot15d4.py:60-63) and yields no addressing fields               */
#define SNRX_ZBMAC_MALFORMED          0x4000  /* the record ends inside a field the header announces          */
#define SNRX_ZBMAC_NOT_ZIGBEE         0x8000  /* the record is not an 802.15.4 record                         */

typedef struct snrx_zbmac {         /* one per record of the batch, 40 bytes */
    uint64_t dest_addr, src_addr;   /* short addresses in the low 16 bits                                      */
    uint16_t dest_panid, src_panid;
    uint16_t fcf;                   /* frame control field as transmitted (little endian)                      */
    uint16_t present;               /* SNRX_ZBMAC_*                                                            */
    uint8_t  seqnum;
    uint8_t  frame_type;            /* 0 beacon, 1 data, 2 ack, 3 command (0xff: record too short)             */
    uint8_t  dest_mode, src_mode;   /* addressing modes 0 none, 2 short, 3 long                                */
    uint8_t  cmd_id;                /* MAC command id (command frames), else 0xff                              */
    uint8_t  payload_off;           /* offset of the MAC payload inside the PSDU                               */
    uint8_t  zll_command;           /* ZLL commissioning command id, 0xff none                                 */
    uint8_t  reserved;
    uint16_t cluster, profile;      /* inter-PAN APS stub cluster / profile                                    */
    uint32_t frame;                 /* index of the record in the batch                                        */
} snrx_zbmac_t;

/* Summaries of the batch most recently retired by snrx_poll / snrx_poll_view, computed on the GPU from the device frame
 * list (call before the second-next snrx_process).  *n_out = records of the batch (BLE records get SNRX_ZBMAC_NOT_ZIGBEE). */
int  snrx_zb_mac_summary(snrx_t* h, snrx_zbmac_t* out, uint32_t cap, uint32_t* n_out);

/* ---- the one exchange of the path -- every engine of a node gets every engine's frame records.
 * No reference counterpart (the reference runs on one host CPU); north_star: "NCCL over NVLink is used only to allgather
 * the decoded-frame records".  Here the gather needs no collective kernel at all: once the engines are connected, the
 * kernel that exports a batch's frame list (k_export_frames) also stores the records -- only the 16-byte pieces a record
 * uses -- and then a {count, batch} header straight into a receive slot in every peer's HBM over NVLink / NVSwitch
 * (peer pointers obtained through CUDA IPC).  Nothing is launched or waited for per batch on the host, and no SM-resident
 * collective waits for the slowest rank beside the channelizer.
 *
 *   snrx_exchange_create   allocates this engine's receive area (8 slots x world x (1 + cap_records) records) and returns
 *                          its 64-byte CUDA IPC handle; the caller passes the handles around (MPI, torch.distributed,
 *                          a file: 64 bytes per rank) ...
 *   snrx_exchange_connect  ... and hands all of them in, in rank order; every later batch is pushed to every rank.
 *   snrx_allgather         waits until batch `batch_no` (0, 1, ... = the order of snrx_process calls on every engine) of
 *                          all ranks has arrived here; counts[r] = records of rank r; with out != NULL the records are
 *                          copied out in rank order (*n_out = their number).  Collective discipline: every rank calls it
 *                          for every batch, in order, and before it queues batch_no + 4 (a slot is reused after 8 batches).
 *                          SNRX_EOVERFLOW if some rank had more than cap_records records in that batch (the caller falls back
 *                          to its own two-phase gather for that batch), SNRX_ESTATE after timeout_ms without the batch. */
#define SNRX_XCHG_HANDLE_BYTES 64
#define SNRX_XCHG_SLOTS 8
#define SNRX_XCHG_MAX_WORLD 16
int  snrx_exchange_create(snrx_t* h, uint32_t rank, uint32_t world, uint32_t cap_records, void* handle_out);
int  snrx_exchange_connect(snrx_t* h, const void* handles);
int  snrx_allgather(snrx_t* h, uint64_t batch_no, snrx_frame_t* out, uint32_t cap, uint32_t* counts, uint32_t* n_out,
                    uint32_t timeout_ms);

/* ---- the transmit side on the GPU -- synthetic / replay captures generated where they are consumed.
 * Waveforms as the reference's transmitters state them: GFSK per vendor/BTLE/host/btle-tools/src/btle_tx.c:1111-1149
 * (gen_sample_from_phy_bit, float version; h = 0.5, 4 samples per symbol, 16 Gaussian taps :103-128), half-sine O-QPSK per
 * snout/grc-blocks/transmitter_OQPSK.py:96-111 with the SHR / PHR of gr-zigbee preamble_prefixer_scapy_impl.cc:48-52,67-88.
 * Every burst is one frame: the caller provides its bits / bytes and where it goes (the schedule is a few bytes per frame
 * and is drawn on the host, snout_b200/synth.py); the GPU modulates, places the bursts in their 4 Msps channel streams,
 * lifts the streams to 96 Msps through a 96-bin synthesis filterbank (x24 polyphase interpolation, 384 taps, and the
 * rotation to the bin centre) and adds white Gaussian noise from a counter-based generator (a function of seed and
 * sample index only: the same call gives the same capture). */
typedef struct snrx_tx_burst {
    int64_t  start;        /* channel-rate (4 Msps) sample index of the burst's first sample                              */
    uint32_t data_offset;  /* offset of the burst's data in `data`                                                       */
    uint32_t n_units;      /* BLE: PHY bits (preamble | access address | whitened PDU + CRC, LSB first, 8 per data byte);
                              802.15.4: PPDU bytes (SHR | PHR | PSDU)                                                     */
    uint16_t bin_slot;     /* index into `bins`: the filterbank bin (snrx_*_channel_bin) the burst is sent on             */
    uint8_t  proto;        /* SNRX_PROTO_BLE (GFSK) or SNRX_PROTO_ZIGBEE (O-QPSK)                                         */
    uint8_t  reserved;
    float    cfo_hz, phase0, amp;   /* carrier offset, initial phase (rad), amplitude                                     */
} snrx_tx_burst_t;                  /* 32 bytes */

/* Writes n_steps * 24 cf32 samples of a 96 Msps capture (centre 2440 MHz) to iq_out (device pointer if out_is_device,
 * else host).  bins[n_bins]: filterbank bins in use (n_bins <= 48); taps[384]: the synthesis prototype (gain 24);
 * gauss[16]: Gaussian taps of the GFSK modulator; sigma: noise standard deviation per I / Q component (0 = none). */
int  snrx_synth_wideband(int device, const snrx_tx_burst_t* bursts, uint32_t n_bursts, const uint8_t* data, uint64_t n_data,
                         const int32_t* bins, uint32_t n_bins, const float* taps, const double* gauss, uint64_t n_steps,
                         float sigma, uint64_t seed, float* iq_out, int out_is_device);

int  snrx_set_channel(snrx_t* h, int channel);           /* NB modes */
int  snrx_set_stream(snrx_t* h, void* cuda_stream);       /* run on a caller stream */
int  snrx_sync(snrx_t* h);
int  snrx_stats(snrx_t* h, snrx_stats_t* s);
int  snrx_debug_stage(snrx_t* h, int stage, void* out, uint64_t cap_bytes, uint64_t* n_bytes);

/* prototype low-pass of the wideband channelizer (taps = 384 or 768 doubles) */
int  snrx_pfb_prototype(int mode, uint32_t taps, double* out);

/* pinned host staging for callers that want full PCIe rate */
int  snrx_host_alloc(void** p, uint64_t bytes);
int  snrx_host_free(void* p);

/* channel plan helpers (btle_rx.c:932-948; top_block.py:56,94-96) */
int  snrx_ble_channel_mhz(int channel);          /* 37->2402 ... ; <0 on error */
int  snrx_zigbee_channel_mhz(int channel);       /* 11->2405 ... 26->2480      */
int  snrx_ble_channel_bin(int channel);          /* PFB bin (0..95) of a BLE channel  */
int  snrx_zigbee_channel_bin(int channel);

#ifdef __cplusplus
}
#endif
#endif /* SNOUTRX_H */
