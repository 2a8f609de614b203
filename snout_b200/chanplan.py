"""Channel plan of the 2.4 GHz ISM band receivers.

BLE channel index -> RF centre follows get_freq_by_channel_number()
(vendor/BTLE/host/btle-tools/src/btle_rx.c:932-948); the 802.15.4 channel ->
centre follows top_block.set_channel()
(snout/modulations/Zigbee/hackrf/Zigbee_rx/top_block.py:56,94-96).

The wideband channelizer (no reference counterpart) is centred on 2440 MHz at
96 Msps with 96 bins of 1 MHz and decimation 24 (4 Msps per channel).
"""

WB_RATE = 96_000_000
WB_CENTER_MHZ = 2440
WB_BINS = 96
WB_DECIM = 24
NB_RATE = 4_000_000
BLE_WINDOW = 8192

BLE_CHANNELS = tuple(range(40))
ZIGBEE_CHANNELS = tuple(range(11, 27))
BLE_ADV_CHANNELS = (37, 38, 39)
BLE_ADV_AA = 0x8E89BED6
BLE_ADV_CRC_INIT = 0x555555


def ble_channel_mhz(ch: int) -> int:
    if ch == 37:
        return 2402
    if ch == 38:
        return 2426
    if ch == 39:
        return 2480
    if 0 <= ch <= 10:
        return 2404 + 2 * ch
    if 11 <= ch <= 36:
        return 2428 + 2 * (ch - 11)
    raise ValueError(f"BLE channel {ch} out of range 0..39")


def zigbee_channel_mhz(ch: int) -> int:
    if not 11 <= ch <= 26:
        raise ValueError(f"802.15.4 channel {ch} out of range 11..26")
    return 2400 + 5 * (ch - 10)


def ble_channel_bin(ch: int) -> int:
    return (ble_channel_mhz(ch) - WB_CENTER_MHZ) % WB_BINS


def zigbee_channel_bin(ch: int) -> int:
    return (zigbee_channel_mhz(ch) - WB_CENTER_MHZ) % WB_BINS


def ble_hop_channels(hop: int, n_events: int, last: int = 0) -> list[int]:
    """Data channel of each of the next connection events as btle_rx -o steps through them with a full channel map:
    hop_chan = (hop_chan + hop) % 37 (btle_rx.c:2194, 2227, 2254)."""
    out = []
    for _ in range(n_events):
        last = (last + hop) % 37
        out.append(last)
    return out
