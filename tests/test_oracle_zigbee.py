"""Zigbee oracle: the sink restatement is pinned against the UNMODIFIED reference packet sink
(scapy-radio/gnuradio/gr-zigbee/lib/packet_sink_scapy_impl.cc) through fixtures generated from it;
the GNU Radio stream blocks are un-vendored, parity with them is unpinned."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from snout_b200 import _abi, synth


def test_chip_mapping_equals_reference(oracle_mod, golden):
    ref = golden("zb_sink_ref.npz")["chip_mapping"]
    assert np.array_equal(oracle_mod.zb_chip_words("port"), ref & 0x7FFFFFFE)
    assert np.array_equal(synth.zb_chip_mapping(), ref & 0x7FFFFFFE)
    assert (ref < 2 ** 31).all()          # why masking CHIP_MAPPING with 0xFFFFFFFE == 0x7FFFFFFE (:208-227)


def test_fcs16_kats(oracle_mod):
    k = json.load(open(os.path.join(GOLDEN, "kats.json")))
    for item in k["fcs16"]:
        frame = bytes.fromhex(item["frame_hex"])
        assert oracle_mod.zb_fcs16(frame[:-2]) == item["fcs"] == int.from_bytes(frame[-2:], "little")
        assert synth.fcs16(frame[:-2]) == item["fcs"]


@pytest.mark.parametrize("seed", [2001, 2002, 2003])
def test_sink_port_matches_reference_sink_fixture(oracle_mod, golden, seed):
    g = golden("zb_sink_ref.npz")
    s, esn0, ch, n = g[f"params_{seed}"]
    cap = synth.zigbee_capture(n=int(n), channel=int(ch), seed=int(s), esn0_db=float(esn0))
    z = oracle_mod.zb_dc_remove(oracle_mod.zb_quad_demod(cap.iq))
    frames, chips, pos = oracle_mod.zb_chain(z, 0, len(z), 0, len(z), want_chips=True)
    assert len(frames) == len(g[f"len_{seed}"]) > 0
    for f, ln, by in zip(frames, g[f"len_{seed}"], g[f"bytes_{seed}"]):
        assert f["len"] == ln and np.array_equal(f["bytes"][:ln], by[:ln])
    # the frame ends (2 + 2*max(len,1)) symbols of 32 chips after the chip that completed the SFD
    for f, end in zip(frames, g[f"end_chip_{seed}"]):
        sync_chip = end - 64 * (1 + max(int(f["len"]), 1))
        assert pos[sync_chip] == f["sample_index"]


def test_recall_and_segmentation(oracle_mod):
    cap = synth.zigbee_capture(n=1_500_000, channel=15, seed=31, esn0_db=15.0)
    whole = oracle_mod.zb_receive(cap.iq, 15, segment=1 << 40, prehalo=0)
    seg = oracle_mod.zb_receive(cap.iq, 15, segment=65536, prehalo=4096)
    truth = [bytes(t.data) for t in cap.truth]
    for fr in (whole, seg):
        got = [bytes(f["bytes"][:f["len"]]) for f in fr]
        assert got == truth and fr["crc_ok"].all()
    # a restarted clock recovery may place the same chip one or two input samples away
    assert (np.abs(whole["sample_index"] - seg["sample_index"]) <= 2).all()
    assert (seg["window"] == seg["sample_index"] // 65536).all()
    assert (np.abs(whole["sample_index"] - np.array([t.anchor for t in cap.truth])) <= 16).all()


def test_dc_tracker_is_shard_invariant(oracle_mod):
    """The blocked DC tracker remembers exactly the 48 preceding 2048-sample blocks: a buffer that
    starts anywhere on the 2048 grid reproduces the whole-capture stream from its 49th block on."""
    B, M = _abi.ZB_IIR_BLOCK, _abi.ZB_IIR_MEMORY_BLOCKS
    rng = np.random.default_rng(5)
    f = (0.08 + 0.7 * rng.standard_normal((M + 12) * B + 1234)).astype(np.float32)
    z = oracle_mod.zb_dc_remove(f)
    for start_block in (1, 3, 7):
        zs = oracle_mod.zb_dc_remove(f[start_block * B:])
        assert np.array_equal(zs[M * B:], z[(start_block + M) * B:])
        assert not np.array_equal(zs[:B], z[start_block * B:(start_block + 1) * B])


@pytest.mark.parametrize("seed,esn0", [(2001, 30.0), (2002, 12.0), (2003, 6.0)])
def test_dc_tracker_equals_the_serial_recurrence(oracle_mod, seed, esn0):
    """a11 against the published block: z of the blocked tracker vs y[n] = a f[n] + (1-a) y[n-1] run serially
    (single_pole_iir_filter_ff + sub_ff, top_block.py:52,70) on the BASELINE config-2 captures (CFO up to +-40 kHz, i.e. a
    DC of up to 0.063 rad/sample under the discriminator).  north_star tolerance: 1e-4 of rms.  Measured: the two differ by
    at most one float ulp of z (the 48-block memory forgets 1.5e-7 of the DC), five orders below the tolerance."""
    cap = synth.zigbee_capture(n=3_000_000, channel=11, seed=seed, esn0_db=esn0)
    f = oracle_mod.zb_quad_demod(cap.iq)
    zs, zb = oracle_mod.zb_dc_remove_serial(f), oracle_mod.zb_dc_remove(f)
    rms = float(np.sqrt(np.mean(zs.astype(np.float64) ** 2)))
    err = float(np.abs(zs.astype(np.float64) - zb).max())
    assert err / rms <= 1e-4                       # the stated tolerance
    assert err <= 2.0 ** -21                       # what it actually is: an ulp of |z| <= 4
    # a pure tone (constant discriminator output): the worst case for a truncated memory
    c = np.full(400_000, 0.0628, np.float32)
    assert np.abs(oracle_mod.zb_dc_remove_serial(c) - oracle_mod.zb_dc_remove(c)).max() <= 2.0 ** -24


def _frameset(fr):
    return {(bytes(f["bytes"][: f["len"]]), int(f["crc_ok"])) for f in fr}


@pytest.mark.parametrize("esn0,max_diff", [(30.0, 0.0), (15.0, 0.005), (12.0, 0.01)])
def test_segmented_receiver_vs_the_serial_flowgraph(oracle_mod, esn0, max_diff):
    """a12 against the flowgraph as it is (zb_oracle_receive_serial: serial DC tracker, ONE clock-recovery + sink chain from
    sample 0): how many frames differ when the chain restarts every 4096 samples with a 2048-sample warm-up (the engine's
    definition).  The clock recovery never forgets its past completely, so at marginal SNR the reported sets differ by a
    few frames -- by about as much as the serial flowgraph differs from ITSELF when the capture starts one sample later
    (`self_diff`, the yardstick).  Longer warm-ups (up to 131072 samples were tried) do not reduce the difference."""
    n_frames = n_diff = n_self = 0
    for seed in (5000, 5001, 5002):
        cap = synth.zigbee_capture(n=3_000_000, channel=11, seed=seed, esn0_db=esn0)
        serial = _frameset(oracle_mod.zb_receive_serial(cap.iq, 11))
        seg = _frameset(oracle_mod.zb_receive(cap.iq, 11))
        shifted = _frameset(oracle_mod.zb_receive_serial(cap.iq[1:], 11))
        n_frames += len(serial)
        n_diff += len(serial ^ seg)
        n_self += len(serial ^ shifted)
    assert n_frames > 250
    assert n_diff <= max_diff * n_frames + (0 if esn0 >= 30 else max(2, 2 * n_self)), (n_frames, n_diff, n_self)


def test_windowed_sink_equals_reference_sink_fixture(oracle_mod, emu, golden):
    """The engine's sink (csrc/zb.cuh zb_sink_window: 32 chips at a time) on the chips the committed fixture was made
    from: same frames at the same chips as the UNMODIFIED packet_sink_scapy_impl.cc (tests/golden/zb_sink_ref.npz)."""
    import ctypes
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    g = golden("zb_sink_ref.npz")
    for seed in (2001, 2002, 2003):
        s, esn0, ch, n = g[f"params_{seed}"]
        cap = synth.zigbee_capture(n=int(n), channel=int(ch), seed=int(s), esn0_db=float(esn0))
        z = oracle_mod.zb_dc_remove(oracle_mod.zb_quad_demod(cap.iq))
        _, chips, _ = oracle_mod.zb_chain(z, 0, len(z), 0, len(z), want_chips=True)
        hard = (chips > 0).astype(np.uint8)
        lens, by, endc = np.zeros(4096, np.int32), np.zeros((4096, 128), np.uint8), np.zeros(4096, np.int64)
        k = emu.emu_zb_sink_chips(P(hard), ctypes.c_int64(len(hard)), 10, P(lens), P(by), P(endc), 4096)
        assert k == len(g[f"len_{seed}"]) > 0
        assert np.array_equal(lens[:k], g[f"len_{seed}"]) and np.array_equal(endc[:k], g[f"end_chip_{seed}"])
        for i in range(k):
            assert np.array_equal(by[i, : lens[i]], g[f"bytes_{seed}"][i, : lens[i]])


def test_edge_cases(oracle_mod):
    assert len(oracle_mod.zb_receive(np.zeros(5, np.complex64), 11)) == 0
    assert len(oracle_mod.zb_receive(np.zeros(100_003, np.complex64), 11)) == 0
    rng = np.random.default_rng(3)
    n = (rng.standard_normal(200_000) + 1j * rng.standard_normal(200_000)).astype(np.complex64)
    assert oracle_mod.zb_receive(n, 11)["crc_ok"].sum() == 0


ref = pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libzbsink_ref.so")),
                         reason="oracle/_ref not built (needs /root/reference)")


@ref
def test_sink_port_equals_reference_sink_live(oracle_mod):
    for seed, esn0 in ((41, 10.0), (42, 8.0)):          # low SNR: aborted frames, false locks
        cap = synth.zigbee_capture(n=800_000, channel=11, seed=seed, esn0_db=esn0, gap=(500, 8000))
        z = oracle_mod.zb_dc_remove(oracle_mod.zb_quad_demod(cap.iq))
        frames, chips, _ = oracle_mod.zb_chain(z, 0, len(z), 0, len(z), want_chips=True)
        refout = oracle_mod.zb_sink_reference(chips)
        assert [bytes(f["bytes"][:f["len"]]) for f in frames] == [b for _, b in refout]


@ref
def test_reference_sink_random_chips(oracle_mod):
    """Random hard chips with embedded valid symbols: both sinks must publish the same blobs."""
    rng = np.random.default_rng(5)
    words = oracle_mod.zb_chip_words("port")
    chips = []
    for _ in range(60):
        chips += list(rng.integers(0, 2, int(rng.integers(10, 400))))
        syms = [0] * 8 + [7, 10] + [3, 0] + [int(x) for x in rng.integers(0, 16, 6)]   # SHR, PHR=3, 3 bytes
        for s in syms:
            bits = [(int(words[s]) >> (31 - k)) & 1 for k in range(32)]
            flip = rng.integers(0, 32, int(rng.integers(0, 4)))
            for f in flip:
                bits[f] ^= 1
            chips += bits
    soft = (np.array(chips, dtype=np.float32) * 2 - 1)
    refout = oracle_mod.zb_sink_reference(soft)
    # drive the port sink through zb_chain is not possible on raw chips; use the C API directly
    import ctypes
    lib = oracle_mod._lib("port")
    st = ctypes.create_string_buffer(lib.zb_oracle_sink_size())
    lib.zb_sink_init(st, 10)
    lib.zb_sink_psdu.restype = ctypes.POINTER(ctypes.c_uint8)
    got = []
    for i, c in enumerate(soft):
        if lib.zb_sink_push(st, int(c > 0), ctypes.c_int64(i), ctypes.c_int64(i)):
            n = lib.zb_sink_len(st)
            got.append((i, bytes(lib.zb_sink_psdu(st)[:n])))
    assert len(refout) > 20 and got == refout


@ref
def test_windowed_sink_equals_reference_sink_live(oracle_mod, emu):
    """zb_sink_window against the UNMODIFIED reference sink on random chips with embedded (damaged) symbols: short and
    empty frames, truncated preambles, false locks, aborts."""
    import ctypes
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    rng = np.random.default_rng(5)
    words = oracle_mod.zb_chip_words("port")
    total = 0
    for trial in range(12):
        chips = []
        for _ in range(60):
            chips += list(rng.integers(0, 2, int(rng.integers(1, 400))))
            ln = int(rng.integers(0, 6))
            syms = [0] * int(rng.integers(1, 9)) + [7, 10] + [ln, 0] + [int(x) for x in rng.integers(0, 16, 2 * max(ln, 1))]
            for s in syms:
                bits = [(int(words[s]) >> (31 - k)) & 1 for k in range(32)]
                for f in rng.integers(0, 32, int(rng.integers(0, 5))):
                    bits[f] ^= 1
                chips += bits
        hard = np.array(chips, dtype=np.uint8)
        refout = oracle_mod.zb_sink_reference(hard.astype(np.float32) * 2 - 1, cap=8192)
        lens, by, endc = np.zeros(8192, np.int32), np.zeros((8192, 128), np.uint8), np.zeros(8192, np.int64)
        k = emu.emu_zb_sink_chips(P(hard), ctypes.c_int64(len(hard)), 10, P(lens), P(by), P(endc), 8192)
        assert [(int(endc[i]), bytes(by[i, : lens[i]])) for i in range(k)] == refout
        total += k
    assert total > 300


def test_span_filter_host_equals_oracle_and_is_shard_invariant(oracle_mod):
    """stream.zb_span_filter (numpy, across shards) == the oracle's zb_span_filter (C, one stream) on random record
    lists; cutting the list anywhere and carrying the state gives the same result; idempotent; BLE records untouched."""
    import ctypes
    from snout_b200 import _abi, stream
    rng = np.random.default_rng(5)
    lib = oracle_mod._lib("port")
    for trial in range(20):
        n = int(rng.integers(1, 60))
        f = np.zeros(n, _abi.FRAME_DTYPE)
        f["proto"], f["channel"], f["capture_id"] = 2, 11, 0
        f["sample_index"] = np.sort(rng.integers(0, 200_000, n))
        f["len"] = rng.integers(5, 128, n)
        f["crc_ok"] = rng.integers(0, 2, n)
        want = f.copy()
        k = lib.zb_span_filter(want.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n))
        want = want[:k]
        got = stream.zb_span_filter(f)
        assert got.tobytes() == want.tobytes()
        assert stream.zb_span_filter(got).tobytes() == got.tobytes()
        cut = int(rng.integers(0, n + 1))
        st = {}
        parts = [stream.zb_span_filter(f[:cut], st), stream.zb_span_filter(f[cut:], st)]
        assert np.concatenate(parts).tobytes() == want.tobytes()
        # two interleaved streams + BLE records
        g = np.concatenate([f, f])
        g["channel"][n:] = 12
        g["proto"][-1] = 3
        out = stream.zb_span_filter(g)
        assert out[out["channel"] == 11].tobytes() == want.tobytes()


def test_segmented_receiver_equals_unsegmented_at_high_snr(oracle_mod):
    """With the span rule the 4096-sample chains report exactly what one unsegmented chain reports (20 dB, dense traffic)."""
    cap = synth.zigbee_capture(n=2_000_000, channel=11, seed=2005, esn0_db=20.0, gap=(500, 6000))
    a = oracle_mod.zb_receive(cap.iq, 11)
    b = oracle_mod.zb_receive(cap.iq, 11, segment=1 << 40, prehalo=0)
    assert len(a) == len(b) > 100 and np.array_equal(a["bytes"], b["bytes"]) and a["crc_ok"].all()


def _reference_psdus(oracle_mod):
    """The 802.15.4 frames of the reference's own test captures (tests/golden/zb_ref_frames.json), FCS appended where the
    capture has none (DLT 230)."""
    g = json.load(open(os.path.join(GOLDEN, "zb_ref_frames.json")))
    out = [bytes.fromhex(h) for h in g["with_fcs"]]
    for h in g["without_fcs"]:
        b = bytes.fromhex(h)
        c = oracle_mod.zb_fcs16(b)
        out.append(b + bytes([c & 0xFF, c >> 8]))
    return g, out


def test_reference_pcap_frames_fcs_and_air_round_trip(oracle_mod):
    g, psdus = _reference_psdus(oracle_mod)
    for h in g["with_fcs"]:                                               # captured with its FCS: a KAT of the FCS-16
        b = bytes.fromhex(h)
        assert oracle_mod.zb_fcs16(b[:-2]) == b[-2] | (b[-1] << 8)
    assert len(psdus) == 55
    # transmit every frame once (O-QPSK per transmitter_OQPSK.py, 20 dB), receive with the oracle
    rng = np.random.default_rng(77)
    sig, truth = synth.zb_baseband(1_000_000, 15, rng, gap=(1500, 6000), psdus=psdus)
    x = (sig + synth._awgn(len(sig), rng, 2.0 / 10 ** 2.0)).astype(np.complex64)
    sent = [bytes(t.data) for t in truth]
    assert len(sent) >= 55 and sent[:55] == psdus
    got = oracle_mod.zb_receive(x, 15)
    assert [bytes(f["bytes"][: f["len"]]) for f in got] == sent and got["crc_ok"].all()
