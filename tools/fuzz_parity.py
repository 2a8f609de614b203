#!/usr/bin/env python3
"""One-off soak of the stage-wise parity statements over many seeds / SNRs / geometries (not a pytest: minutes of GPU time).
BLE wideband: frames == the unmodified btle_rx.c (where compiled) and its restatement on the engine's own quantised streams.
Zigbee wideband: discriminator, DC-removed stream and frames == the oracle on the engine's own channel streams, default and odd
chain geometries.  Mixed: == the two single-protocol engines.  Prints a summary; exits non-zero on the first mismatch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from conftest import assert_frames_equal, ble_expected
from snout_b200 import _abi, synth, chanplan
from snout_b200.engine import RxEngine
oracle.build()
n_ble = n_zb = 0
t0 = time.time()
for seed in range(7000, 7000 + int(os.environ.get("FUZZ_BLE", 12))):
    rng = np.random.default_rng(seed)
    esn0 = float(rng.choice([9.0, 12.0, 15.0, 20.0, 30.0]))
    spread = float(rng.choice([0.0, 10.0, 25.0]))
    sec = float(rng.choice([0.0045, 0.007, 0.0101]))
    cap = synth.wideband_capture(seconds=sec, kind="ble", seed=seed, esn0_db=esn0, gap=(200, int(rng.integers(1500, 6000))), amp_db_spread=spread)
    x = cap.iq[: len(cap.iq) - 24 * int(rng.integers(0, 200))]
    taps = int(rng.choice([384, 768]))
    with RxEngine("ble_wb40", max_samples=len(x), pfb_taps=taps, keep_streams=True) as e:
        got = e.run(x); q8 = e.debug_stage(_abi.STAGE_BLE_Q8)[0]
    want = np.concatenate([ble_expected(oracle, q8[c], c) for c in range(40)])
    assert_frames_equal(got, want, what=f"ble seed {seed} esn0 {esn0} taps {taps}")
    with RxEngine("ble_wb40", max_samples=len(x), pfb_taps=taps) as e:
        assert_frames_equal(e.run(x), want, what=f"ble production kernel seed {seed}")
    n_ble += len(want)
for seed in range(7100, 7100 + int(os.environ.get("FUZZ_ZB", 8))):
    rng = np.random.default_rng(seed)
    esn0 = float(rng.choice([9.0, 12.0, 20.0, 30.0]))
    seg, pre = [(0, 0), (4096, 2048), (8192, 2048), (16384, 4096), (6144, 2048)][int(rng.integers(0, 5))]
    cap = synth.wideband_capture(seconds=0.03, kind="zigbee", seed=seed, esn0_db=esn0, gap=(400, int(rng.integers(3000, 20000))))
    x = cap.iq[: len(cap.iq) - 24 * int(rng.integers(0, 100))]
    with RxEngine("zb_wb16", max_samples=len(x), keep_streams=True, zb_segment=seg, zb_prehalo=pre) as e:
        got = e.run(x); y = e.debug_stage(_abi.STAGE_CHAN_CF32)[0]; f = e.debug_stage(_abi.STAGE_ZB_F)[0]; z = e.debug_stage(_abi.STAGE_ZB_DISC)[0]
    want = []
    kw = dict(segment=seg, prehalo=pre) if seg else {}
    for c in range(16):
        fo = oracle.zb_quad_demod(y[c]); assert np.array_equal(f[c], fo), (seed, c, "disc")
        zo = oracle.zb_dc_remove(fo); assert np.array_equal(z[c], zo), (seed, c, "dc")
        want.append(oracle.zb_receive_z(zo, 11 + c, **kw))
    want = np.concatenate(want)
    assert_frames_equal(got, want, what=f"zb seed {seed} esn0 {esn0} seg {seg}/{pre}")
    with RxEngine("zb_wb16", max_samples=len(x), zb_segment=seg, zb_prehalo=pre) as e:
        assert_frames_equal(e.run(x), want, what=f"zb production kernel seed {seed}")
    n_zb += len(want)
n_nb = 0
for seed in range(7300, 7300 + int(os.environ.get("FUZZ_NB", 6))):
    rng = np.random.default_rng(seed)
    ch = int(rng.choice([37, 38, 39, 5, 22]))
    n = int(rng.integers(60_000, 400_000))
    cap = synth.ble_capture(n=n, channel=ch, seed=seed, esn0_db=float(rng.choice([8.0, 12.0, 20.0, 30.0])))
    with RxEngine("ble_nb", channel=ch, max_samples=n) as e:
        got = e.run(cap.iq)
    want = ble_expected(oracle, oracle.ble_quantize(cap.iq, 128.0), ch)
    assert_frames_equal(got, want, what=f"ble_nb seed {seed} ch {ch} n {n}")
    zch = int(rng.integers(11, 27))
    zc = synth.zigbee_capture(n=n, channel=zch, seed=seed, esn0_db=float(rng.choice([9.0, 12.0, 20.0, 30.0])))
    with RxEngine("zb_nb", channel=zch, max_samples=n) as e:
        gz = e.run(zc.iq)
    assert_frames_equal(gz, oracle.zb_receive(zc.iq, zch), what=f"zb_nb seed {seed} ch {zch} n {n}")
    n_nb += len(want) + len(gz)
for seed in range(7200, 7204):
    cap = synth.wideband_capture(seconds=0.0125, kind="mixed", seed=seed, esn0_db=20.0, gap=(400, 5000))
    with RxEngine("ble_wb40", max_samples=len(cap.iq)) as e: a = e.run(cap.iq)
    with RxEngine("zb_wb16", max_samples=len(cap.iq)) as e: b = e.run(cap.iq)
    with RxEngine("mixed_wb56", max_samples=len(cap.iq)) as e: m = e.run(cap.iq)
    assert_frames_equal(m, np.concatenate([a, b]), what=f"mixed seed {seed}")
print(f"fuzz ok: {n_ble} BLE frames over {os.environ.get('FUZZ_BLE', 12)} captures, {n_zb} 802.15.4 frames over {os.environ.get('FUZZ_ZB', 8)} captures, 4 mixed captures, {n_nb} narrow-band frames over 2 x {os.environ.get('FUZZ_NB', 6)} captures, {time.time() - t0:.0f} s")
