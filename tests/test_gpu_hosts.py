"""The reference-facing hosts on the real CUDA engine (`pytest -m gpu`): streaming in shards equals
one call over the whole capture bit for bit; `btle_rx` prints the lines the reference prints;
`Zigbee_rx` serves XMLRPC and sends the RFtap datagrams of the frames the engine decodes."""
import io
import os
import socket
import subprocess
import sys
import xmlrpc.client

import numpy as np
import pytest

from conftest import ROOT, assert_frames_equal, ble_expected
from snout_b200 import _abi, btle_cli, formats, stream, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Engine():
    from snout_b200.engine import RxEngine
    assert _abi.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return RxEngine


def _stream_all(eng, x, units, block=50_000):
    st = stream.ShardStreamer(eng, units_per_shard=units)
    out = []
    for off in range(0, len(x), block):
        out += list(st.feed(x[off: off + block]))
    out += list(st.flush())
    st.close()
    return np.concatenate(out) if out else np.zeros(0, _abi.FRAME_DTYPE)


def test_stream_equals_whole_ble_nb(Engine, oracle_mod):
    cap = synth.ble_capture(n=8192 * 41 + 3001, channel=38, seed=81, esn0_db=28, gap=(100, 1500))
    with Engine("ble_nb", channel=38, max_samples=len(cap.iq)) as e:
        whole = e.run(cap.iq)
    with Engine("ble_nb", channel=38, max_samples=8192 * 6 + 128 + 2048) as e:
        got = _stream_all(e, cap.iq, units=6, block=33_333)
    assert len(whole) > 150
    assert_frames_equal(got, whole, what="BLE narrow band: streamed shards vs whole")
    assert_frames_equal(whole, ble_expected(oracle_mod, oracle_mod.ble_quantize(cap.iq, 128.0), 38), what="vs oracle")


def test_stream_equals_whole_ble_wb40(Engine):
    cap = synth.wideband_capture(seconds=0.03, kind="ble", seed=4400, gap=(200, 2500))
    with Engine("ble_wb40", max_samples=len(cap.iq)) as e:
        whole = e.run(cap.iq)
    with Engine("ble_wb40", max_samples=(8192 * 4 + 128 + 2048) * 24) as e:
        got = _stream_all(e, cap.iq, units=4, block=1_000_003)
    order = np.lexsort((got["sample_index"], got["window"], got["channel"]))
    assert len(whole) > 500
    assert_frames_equal(got[order], whole, what="BLE wideband: streamed shards vs whole")


def test_stream_equals_whole_zigbee(Engine):
    cap = synth.zigbee_capture(n=1_000_000, channel=15, seed=91, esn0_db=15.0, gap=(500, 9000))
    with Engine("zb_nb", channel=15, max_samples=len(cap.iq)) as e:
        whole = e.run(cap.iq)
    from snout_b200 import stream
    unit, pre, post = stream.shard_geometry(0, 1)
    with Engine("zb_nb", channel=15, max_samples=24 * unit + pre + post) as e:
        got = _stream_all(e, cap.iq, units=24, block=77_777)
    assert len(whole) > 30
    assert_frames_equal(got, whole, what="Zigbee: streamed shards vs whole")


def test_stream_equals_whole_mixed(Engine):
    cap = synth.wideband_capture(seconds=0.045, kind="mixed", seed=5100, esn0_db=25.0, gap=(400, 5000))
    with Engine("mixed_wb56", max_samples=len(cap.iq), zb_segment=16384) as e:
        whole = e.run(cap.iq)
    from snout_b200 import stream
    unit, pre, post = stream.shard_geometry(40, 16, 16384)
    with Engine("mixed_wb56", max_samples=(3 * unit + pre + post) * 24, zb_segment=16384) as e:
        got = _stream_all(e, cap.iq, units=3, block=2_000_000)
    key = lambda f: np.lexsort((f["sample_index"], f["window"], f["channel"], 255 - f["proto"].astype(int)))   # noqa: E731
    assert len(whole) > 100 and (whole["proto"] == 2).sum() > 4
    assert_frames_equal(got[key(got)], whole[key(whole)], what="mixed wideband: streamed shards vs whole")


def test_ble_access_mask(Engine, oracle_mod, golden):
    """-m: only masked access-address bits take part in the match (btle_rx.c:1395-1401, 2301)."""
    g = golden("btle_mask_ref.npz")
    cap = synth.ble_capture(n=400_000, channel=37, seed=12, esn0_db=18, gap=(100, 1500))
    q = oracle_mod.ble_quantize(cap.iq, 128.0)
    for mask in (0xFFFFFF00, 0x00FFFFFF, 0xFFFF0000, 0xFFFFFFFE):
        with Engine("ble_nb", channel=37, max_samples=len(cap.iq), access_mask=mask, max_frames=1 << 16) as e:
            got = e.run(cap.iq)
        assert len(got) > 10
        assert_frames_equal(got, g[f"frames_{mask:08x}"], what=f"mask {mask:08x} vs reference fixture")
        assert_frames_equal(got, ble_expected(oracle_mod, q, 37, aa_mask=mask), what=f"mask {mask:08x} vs oracle")


# ------------------------------------------------------------------------------------ btle_rx
def test_btle_rx_cli_prints_reference_lines(Engine, golden, tmp_path):
    """The executable on the golden capture: same lines as the reference receiver() printed
    (fixture made from the unmodified btle_rx.c), from a cf32 file and from the HackRF int8 format."""
    g = golden("btle_synth_ref.npz")
    s, esn0, ch, n = g["params_1002"]
    cap = synth.ble_capture(n=int(n), channel=int(ch), seed=int(s), esn0_db=float(esn0))
    want = [formats.strip_timestamp(ln) for ln in str(g["stdout_1002"]).splitlines(keepends=True)]
    f32 = tmp_path / "cap.cf32"
    cap.iq.tofile(f32)
    q = np.clip(np.rint(cap.iq.view(np.float32) * 128.0), -128, 127).astype(np.int8)
    f8 = tmp_path / "cap.sc8"
    q.tofile(f8)
    for path, fmt in ((f32, "cf32"), (f8, "sc8")):
        out = io.StringIO()
        o = btle_cli.parse_commandline(["-c", str(int(ch)), "-g", "6", "-a", "8e89bed6", "-k", "555555", "--iq", str(path),
                                        "--format", fmt, "--shard-windows", "16", "-s", str(tmp_path / "o.pcap")], out)
        assert btle_cli.run(o, out) == 0
        lines = out.getvalue().splitlines(keepends=True)
        got = [formats.strip_timestamp(ln) for ln in lines if " Pkt" in ln]
        assert got == want and len(want) > 50
        assert lines[-1] == "Exit main loop ...\n"
        lt, recs = formats.read_pcap(open(tmp_path / "o.pcap", "rb").read())
        assert lt == 256 and len(recs) == len(want)
    # as a child process, the way PController runs it (snout/core/pcontroller.py:54,115)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "btle_rx"), "-c", str(int(ch)), "-g", "6", "-a", "8e89bed6",
                        "-k", "555555", "--iq", str(f32)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert [formats.strip_timestamp(ln) for ln in r.stdout.splitlines(keepends=True) if " Pkt" in ln] == want


def test_btle_rx_cli_wideband(Engine, tmp_path):
    cap = synth.wideband_capture(seconds=0.02, kind="ble", seed=4500, gap=(200, 2500))
    path = tmp_path / "wb.cf32"
    cap.iq.tofile(path)
    with Engine("ble_wb40", max_samples=len(cap.iq)) as e:
        whole = e.run(cap.iq)
    out = io.StringIO()
    o = btle_cli.parse_commandline(["-c", "39", "--iq", str(path), "--wideband", "--shard-windows", "8"], out)
    assert btle_cli.run(o, out) == 0
    sel = whole[whole["channel"] == 39]
    want = [formats.strip_timestamp(ln) for ln in formats.btle_rx_lines(sel)]
    got = [formats.strip_timestamp(ln) for ln in out.getvalue().splitlines(keepends=True) if " Pkt" in ln]
    assert got == want and len(want) > 10


# ------------------------------------------------------------------------------------ Zigbee_rx
def test_zigbee_rx_flowgraph_end_to_end(Engine, tmp_path):
    from snout_b200.zigbee_rx import top_block
    cap = synth.zigbee_capture(n=700_000, channel=11, seed=2001, esn0_db=25.0)
    path = tmp_path / "zb.cf32"
    cap.iq.tofile(path)
    with Engine("zb_nb", channel=11, max_samples=len(cap.iq)) as e:
        whole = e.run(cap.iq)
    assert len(whole) > 10
    rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rx.bind(("127.0.0.1", 0))
    rx.settimeout(20)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    tb = top_block(channel=11, iq=str(path), xmlrpc_addr=("localhost", port), udp_dest=rx.getsockname(), segments_per_shard=4)
    srv = xmlrpc.client.ServerProxy(f"http://localhost:{port}")
    assert srv.get_channel() == 11
    tb.start()
    got = [rx.recvfrom(4096)[0] for _ in range(len(whole))]
    tb.wait()
    assert tb.error is None and tb.frames_sent == len(whole)
    assert got == [formats.rftap_datagram(f) for f in whole]
    for d, t in zip(got, cap.truth):
        assert formats.parse_rftap(d)["payload"] == bytes(t.data)


def test_zigbee_rx_wideband_channel_select(Engine, tmp_path):
    from snout_b200.zigbee_rx import top_block
    cap = synth.wideband_capture(seconds=0.03, kind="zigbee", seed=3200, esn0_db=22.0, gap=(400, 5000))
    with Engine("zb_wb16", max_samples=len(cap.iq), zb_segment=16384) as e:
        whole = e.run(cap.iq)
    ch = int(np.bincount(whole["channel"]).argmax())
    sel = whole[whole["channel"] == ch]
    rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rx.bind(("127.0.0.1", 0))
    rx.settimeout(20)
    tb = top_block(channel=ch, wideband=True, serve_xmlrpc=False, udp_dest=rx.getsockname(), blocks=[cap.iq],
                   segments_per_shard=3, zb_segment=16384)
    tb.start()
    got = [rx.recvfrom(4096)[0] for _ in range(len(sel))]
    tb.wait()
    assert tb.error is None and tb.frames_sent == len(sel) > 2
    assert got == [formats.rftap_datagram(f) for f in sel]


def test_device_frame_list_and_header(Engine):
    """snrx_polled_frames_device: the HBM list equals the host records, the record in front of it carries {count, batch number},
    and a polled list survives two further process() calls (each lane alternates two lists)."""
    import torch
    from snout_b200.dist import _DevView
    cap = synth.ble_capture(n=400_000, channel=37, seed=1001, esn0_db=25)
    other = synth.ble_capture(n=400_000, channel=37, seed=1002, esn0_db=25)
    with Engine("ble_nb", channel=37, max_samples=400_000) as e:
        for batch in range(3):
            fr = e.process(cap.iq).poll()
            ptr, records, n = e.polled_frames_device()
            assert n == len(fr) > 10 and records >= n
            view = torch.as_tensor(_DevView(ptr - 160, (n + 1) * 160), device="cuda")
            hdr = view[:16].cpu().numpy().view(np.uint64)
            assert (int(hdr[0]), int(hdr[1])) == (n, 3 * batch)          # three process() calls per round of this loop
            e.process(other.iq)                                   # two further batches (both lanes) ...
            e.poll()
            e.process(other.iq)
            dev = view[160:].cpu().numpy().view(_abi.FRAME_DTYPE)  # ... and the polled list is still intact
            assert dev.tobytes() == fr.tobytes()
            e.poll()
