"""Host logic of the streaming front door (no GPU): the shard plan obeys the halo rules of
include/snoutrx.h, every stream sample lands in exactly one shard body, halos carry the right
samples, two shards are in flight, and the btle_rx / Zigbee_rx hosts drive it correctly.
A recording stand-in replaces the engine here; the DSP itself is covered by `-m gpu` tests."""
import io
import socket
import threading
import time
import xmlrpc.client

import numpy as np
import pytest

from snout_b200 import _abi, btle_cli, chanplan, formats, stream


from fake_engine import MARK, RecordingEngine


def _check_cover(eng, x, decim, unit, pre, post):
    """Bodies tile the stream; halos hold the neighbouring samples; ABI constraints hold."""
    pos = 0
    for i, (buf, sh) in enumerate(eng.calls):
        last = i == len(eng.calls) - 1
        b0 = sh["first_window"] * 8192 * decim
        assert b0 == pos and (b0 // decim) % unit == 0
        lo = b0 - sh["pre_samples"]
        assert sh["pre_samples"] == min(b0, pre * decim)          # a full pre halo, or everything back to the start of the stream
        body = sh["body_samples"] if sh["body_samples"] else len(buf) - sh["pre_samples"]
        assert (sh["body_samples"] == 0) == last
        n_expect = len(buf)
        assert np.array_equal(buf, x[lo: lo + n_expect]), f"shard {i} content"
        if not last:
            assert len(buf) == sh["pre_samples"] + body + post * decim
        pos = b0 + body
    tail = len(x) - len(x) % decim
    assert pos == tail


@pytest.mark.parametrize("mode,n,units", [("ble_nb", 8192 * 37 + 555, 8), ("ble_nb", 8192 * 16, 8), ("ble_nb", 1000, 4),
                                          ("zb_nb", 65536 * 7 + 4321, 2), ("ble_wb40", 24 * (8192 * 9 + 100) + 7, 3),
                                          ("zb_wb16", 24 * (16384 * 11 + 5), 4), ("mixed_wb56", 24 * 65536 * 3 + 48, 1)])
def test_streamer_covers_stream_exactly(mode, n, units):
    seg = 16384 if mode == "zb_wb16" else 65536
    decim = 24 if "wb" in mode else 1
    n_ble = 0 if mode.startswith("zb") else 1
    n_zb = 0 if mode.startswith("ble") else 1
    unit, pre, post = stream.shard_geometry(n_ble, n_zb, seg, 4096)
    eng = RecordingEngine(mode, max_samples=(units * unit + pre + post) * decim, zb_segment=seg, zb_prehalo=4096)
    st = stream.ShardStreamer(eng, units_per_shard=units)
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    got, off = [], 0
    while off < n:                                   # ragged block sizes
        step = int(rng.integers(1, 3 * unit * decim))
        for fr in st.feed(x[off: off + step]):
            got.append(fr)
        off += step
    for fr in st.flush():
        got.append(fr)
    st.close()
    assert eng.max_queue == (2 if len(eng.calls) > 1 else 1) and not eng.queue
    _check_cover(eng, x, decim, unit, pre, post)
    assert [int(f["window"][0]) for f in got] == [c[1]["first_window"] for c in eng.calls]      # stream order
    # the offline planner cuts the same shards
    plan = stream.plan_shards(len(x), decim, unit, pre, post, units)
    assert [(p["pre_samples"], p["body_samples"], p["first_window"]) for p in plan] == \
           [(c[1]["pre_samples"], c[1]["body_samples"], c[1]["first_window"]) for c in eng.calls]
    assert [p["hi"] - p["lo"] for p in plan[:-1]] == [len(c[0]) for c in eng.calls[:-1]]


def test_shard_geometry_rules():
    assert stream.shard_geometry(1, 0) == (8192, 128, 2048)
    unit, pre, post = stream.shard_geometry(0, 1, 65536, 4096)
    need = (_abi.ZB_IIR_MEMORY_BLOCKS + 1) * _abi.ZB_IIR_BLOCK + 4096       # tracker memory + the buffer's first block + chain warm-up
    assert unit == 65536 and pre == need == 104448 and pre % _abi.ZB_IIR_BLOCK == 0 and post >= 16448
    assert stream.shard_geometry(40, 16, 65536, 4096) == (65536, pre, post)
    # library defaults (4096-sample segments, 2048 warm-up): bodies stay on the 8192-sample window grid
    assert stream.shard_geometry(0, 1) == (8192, need - 2048, post) and stream.shard_geometry(40, 16)[0] == 8192
    with pytest.raises(ValueError):
        stream.shard_geometry(0, 1, 10000, 4096)
    with pytest.raises(ValueError):
        stream.ShardStreamer(RecordingEngine("zb_nb", max_samples=65536, zb_segment=65536, zb_prehalo=4096))


def test_iq_blocks_formats(tmp_path):
    x = (np.arange(20, dtype=np.float32) / 128.0).view(np.complex64)
    p = tmp_path / "a.cf32"
    x.tofile(p)
    assert np.array_equal(np.concatenate(list(stream.iq_blocks(str(p), "cf32", block_samples=3))), x)
    q = np.array([[1, -2], [127, -128], [0, 5]], np.int8)
    p8 = tmp_path / "a.sc8"
    q.tofile(p8)
    y = np.concatenate(list(stream.iq_blocks(str(p8), "sc8", block_samples=2)))
    assert np.array_equal(np.rint(y.view(np.float32) * 128).astype(np.int8).reshape(-1, 2), q)


# ------------------------------------------------------------------------------------ btle_rx host
def test_btle_cli_option_table_matches_reference():
    out = io.StringIO()
    o = btle_cli.parse_commandline(["-c", "38", "-g", "6", "-a", "8e89bed6", "-k", "555555"], out)   # snout/util/btle.py:53
    assert (o.chan, o.gain, o.access_addr, o.crc_init, o.access_mask) == (38, 6, 0x8E89BED6, 0x555555, 0xFFFFFFFF)
    assert out.getvalue() == formats.BTLE_RX_BANNER
    o = btle_cli.parse_commandline(["--chan=5", "--access", "0xAF9A9356", "--crcinit", "abcdef", "-m", "ffffff00", "-v",
                                    "-s", "x.pcap", "-f", "2402000000"], io.StringIO())
    assert (o.chan, o.access_addr, o.crc_init, o.access_mask, o.verbose, o.filename_pcap, o.freq_hz) == \
           (5, 0xAF9A9356, 0xABCDEF, 0xFFFFFF00, 1, "x.pcap", 2402000000)
    for bad, msg in ((["-c", "40"], "channel number must be within 0~39!"), (["-g", "63"], "rx gain must be within 0~62!"),
                     (["-c", "37", "extra"], "Error: unknown/extra arguments specified on command line!")):
        out = io.StringIO()
        assert btle_cli.parse_commandline(bad, out) is None
        assert msg in out.getvalue() and "Usage:" in out.getvalue()
    assert btle_cli.parse_commandline(["-c", "3x"], io.StringIO()).chan == 3          # strtol semantics


def test_btle_cli_run_prints_reference_lines(tmp_path):
    made = []

    def factory(mode, **kw):
        made.append(RecordingEngine(mode, **kw))
        return made[-1]
    out = io.StringIO()
    o = btle_cli.parse_commandline(["-c", "37", "-s", str(tmp_path / "p.pcap"), "--shard-windows", "4"], out)
    x = np.zeros(8192 * 9 + 17, np.complex64)
    rc = btle_cli.run(o, out, engine_factory=factory, blocks=[x[:30000], x[30000:]])
    assert rc == 0 and made[0].closed and made[0].mode == "ble_nb"
    assert made[0].kw["access_addr"] == 0x8E89BED6 and made[0].kw["crc_init"] == 0x555555
    lines = out.getvalue().splitlines()
    assert lines[0] == "BLE sniffer. Xianjun Jiao. putaoshu@msn.com" and lines[1] == ""
    assert lines[2] == "Cmd line input: chan 37, freq 2402MHz, access addr 8e89bed6, crc init 555555 raw 0 verbose 0 rx 6dB (B200) file=%s" % (tmp_path / "p.pcap")
    assert lines[3].startswith("will store packets to: ")
    pk = [ln for ln in lines if " Pkt" in ln]
    assert len(pk) == 3 == len(made[0].calls)
    for i, ln in enumerate(pk):                      # what BtleMessage.fromraw needs (message.py:226-235)
        tok = (ln + "\n").split(" ")
        assert len(tok) == 11 and tok[-1] == "CRC0\n" and tok[1] == f"Pkt{i + 1}" and tok[2] == "Ch37"
        assert tok[8] == "AdvA:060504030201"
    assert lines[-1] == "Exit main loop ..."
    lt, recs = formats.read_pcap(open(tmp_path / "p.pcap", "rb").read())
    assert lt == 256 and len(recs) == 3
    # refused options and missing source
    for argv in (["-o"], ["-r"], []):
        out = io.StringIO()
        assert btle_cli.run(btle_cli.parse_commandline(argv, out), out) == 1


def test_btle_cli_wideband_channel_filter():
    class Wb(RecordingEngine):
        def process(self, iq, shard=None):
            super().process(iq, shard)
            f = np.repeat(self.queue[-1], 3)
            f["channel"] = [37, 38, 5]
            self.queue[-1] = f
            return self
    made = []
    out = io.StringIO()
    o = btle_cli.parse_commandline(["-c", "38", "--wideband", "--shard-windows", "2"], out)
    btle_cli.run(o, out, engine_factory=lambda m, **k: made.append(Wb(m, **k)) or made[-1],
                 blocks=[np.zeros(24 * 8192 * 2, np.complex64)])
    assert made[0].mode == "ble_wb40"
    assert [ln.split(" ")[2] for ln in out.getvalue().splitlines() if " Pkt" in ln] == ["Ch38"] * len(made[0].calls)
    out = io.StringIO()
    o = btle_cli.parse_commandline(["--wideband", "--all-channels", "--shard-windows", "2"], out)
    btle_cli.run(o, out, engine_factory=lambda m, **k: Wb(m, **k), blocks=[np.zeros(24 * 8192, np.complex64)])
    pk = [ln.split(" ") for ln in out.getvalue().splitlines() if " Pkt" in ln]
    assert [(t[1], t[2]) for t in pk] == [("Pkt1", "Ch37"), ("Pkt1", "Ch38"), ("Pkt1", "Ch5")]       # numbering per channel


# ------------------------------------------------------------------------------------ Zigbee_rx host
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_zigbee_top_block_interface_xmlrpc_and_udp():
    from snout_b200.zigbee_rx import top_block
    rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rx.bind(("127.0.0.1", 0))
    rx.settimeout(5)
    port = _free_port()
    made = []
    x = np.zeros(65536 * 5 + 99, np.complex64)
    tb = top_block(channel=11, xmlrpc_addr=("localhost", port), udp_dest=rx.getsockname(), segments_per_shard=2, zb_segment=65536,
                   engine_factory=lambda m, **k: made.append(RecordingEngine(m, **k)) or made[-1], blocks=[x])
    srv = xmlrpc.client.ServerProxy(f"http://localhost:{port}")
    assert srv.get_channel() == 11 and srv.get_samp_rate() == 4000000
    srv.set_channel(20)                                         # gnuradio_set_vars(channel=20), gnuradio.py:323-338
    assert srv.get_channel() == 20 and tb.get_center_freq() == 2450000000
    with pytest.raises(xmlrpc.client.Fault):
        srv.get_startup_var()                                   # wait_for_radio() accepts a Fault (gnuradio.py:342-381)
    tb.start()
    got = [rx.recvfrom(4096)[0] for _ in range(3)]
    tb.stop()
    tb.wait()
    assert made[0].mode == "zb_nb" and made[0].closed and tb.frames_sent == 3 == len(made[0].calls)
    for d in got:
        p = formats.parse_rftap(d)
        assert d[:16] == bytes.fromhex("5246746104000101c30000000000803f") and p["dlt"] == 195 and p["qual"] == 1.0
        assert p["payload"] == bytes(MARK)
    # GnuradioPacket encapsulation of the older flowgraphs: first byte 2 -> GnuradioSocket.recv builds a GnuradioPacket
    tb = top_block(channel=26, serve_xmlrpc=False, udp_dest=rx.getsockname(), encap="gnuradio", segments_per_shard=2, zb_segment=65536,
                   engine_factory=lambda m, **k: RecordingEngine(m, **k), blocks=[x[:70000]])
    tb.start()
    d = rx.recvfrom(4096)[0]
    tb.wait()
    assert d[:8] == bytes([2, 0, 0, 0, 0, 0, 0, 0]) and d[8:] == bytes(MARK)
