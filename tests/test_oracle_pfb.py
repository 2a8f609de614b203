"""The channelizer has no reference counterpart; its CPU statement (oracle/pfb_oracle.c) is checked
against the textbook definition evaluated in numpy and against its own fast factorisation."""
import numpy as np

from snout_b200 import chanplan, synth


def _taps():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from gen_tables import PFB_DESIGNS, kaiser_lowpass
    return {k: kaiser_lowpass(*v) for k, v in PFB_DESIGNS.items()}


def test_direct_matches_numpy_definition(oracle_mod):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(24 * 600) + 1j * rng.standard_normal(24 * 600)).astype(np.complex64)
    h = _taps()["BLE_384"]
    n = np.arange(len(h))
    for k in (0, 2, 58, 95, 33):
        y = oracle_mod.pfb(x, h, [k], m0=40, m1=48)[0]
        for i, m in enumerate(range(40, 48)):
            want = np.sum(h * np.exp(2j * np.pi * k * n / 96) * x[24 * m - n].astype(np.complex128)) * (-1j) ** (k * m)
            assert abs(y[i] - want) < 1e-5


def test_fast_matches_direct(oracle_mod):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(24 * 3000) + 1j * rng.standard_normal(24 * 3000)).astype(np.complex64)
    for name in ("BLE_384", "ZB_768"):
        h = _taps()[name]
        bins = [chanplan.ble_channel_bin(c) for c in (37, 0, 12, 39)] + [chanplan.zigbee_channel_bin(c) for c in (11, 26)]
        a = oracle_mod.pfb(x, h, bins)
        b = oracle_mod.pfb(x, h, bins, fast=True)
        assert np.abs(a - b).max() / np.sqrt(np.mean(np.abs(a) ** 2)) < 1e-5


def test_tone_lands_in_its_bin(oracle_mod):
    h = _taps()["BLE_384"]
    n = np.arange(24 * 2000)
    for ch in (37, 17, 39):
        f = (chanplan.ble_channel_mhz(ch) - chanplan.WB_CENTER_MHZ) * 1e6 + 100e3
        x = np.exp(2j * np.pi * f * n / chanplan.WB_RATE).astype(np.complex64)
        bins = [chanplan.ble_channel_bin(c) for c in range(40)]
        y = oracle_mod.pfb(x, h, bins, m0=100, m1=1900)
        p = np.mean(np.abs(y) ** 2, axis=1)
        assert np.argmax(p) == ch and abs(p[ch] - 1.0) < 5e-3      # pass-band ripple
        assert np.sort(p)[-2] < 1e-4                       # >= 40 dB to every other channel
        # residual 100 kHz offset: phase advances 2 pi * 0.1/4 per channel sample
        d = np.angle(y[ch, 1:] * np.conj(y[ch, :-1]))
        assert np.allclose(d, 2 * np.pi * 0.1 / 4, atol=1e-3)


def test_wideband_chain_recall(oracle_mod):
    cap = synth.wideband_capture(seconds=0.0045, kind="ble", seed=4000, gap=(200, 2000))
    h = _taps()["BLE_384"]
    y = oracle_mod.pfb(cap.iq, h, [chanplan.ble_channel_bin(c) for c in range(40)], fast=True)
    ok = tot = 0
    for c in range(40):
        fr = oracle_mod.ble_decode(oracle_mod.ble_quantize(y[c], 100.0), c)
        truth = {bytes(t.data) for t in cap.truth if t.channel == c}
        ok += len(truth & {bytes(f["bytes"][:f["len"]]) for f in fr if f["crc_ok"]})
        tot += len(truth)
    assert tot > 200 and ok >= 0.97 * tot
