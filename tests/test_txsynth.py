"""the transmit side on the GPU (csrc/synth.cuh, snrx_synth_wideband).  The frame schedule is drawn on the
host with the very draws of the numpy generator, so the noise-free GPU capture must equal snout_b200/synth.py's
wideband_capture sample for sample within float32 rounding; with noise it must decode to the frames that were sent."""
import math

import numpy as np
import pytest

from snout_b200 import _abi, chanplan, synth


def test_schedule_places_the_frames_of_the_numpy_generator():
    for kind, gap in (("ble", (200, 2500)), ("zigbee", None), ("mixed", (400, 5000))):
        cap = synth.wideband_capture(seconds=0.0045, kind=kind, seed=4000, gap=gap)
        bursts, blob, bins, truth, n_ch = synth.wideband_schedule(0.0045, kind, 4000, gap=gap)
        assert [(t.channel, t.start, t.anchor, bytes(t.data), t.proto) for t in truth] == \
               [(t.channel, t.start, t.anchor, bytes(t.data), t.proto) for t in cap.truth]
        assert n_ch * 24 == len(cap.iq) and len(bursts) == len(truth) and bursts.dtype.itemsize == 32
        assert (np.diff(bursts["data_offset"].astype(np.int64)) > 0).all() and len(bins) == len(set(bins.tolist())) <= 48


@pytest.mark.gpu
@pytest.mark.parametrize("kind,gap", [("ble", (200, 2500)), ("mixed", (400, 5000))])
def test_gpu_capture_equals_numpy_statement(kind, gap):
    import torch
    want = synth.wideband_capture(seconds=0.0085, kind=kind, seed=4100, gap=gap, esn0_db=300.0).iq      # noise ~ 1e-15
    got, truth = synth.wideband_capture_gpu(seconds=0.0085, kind=kind, seed=4100, gap=gap, esn0_db=None)
    got = got.cpu().numpy()
    assert got.shape == want.shape and len(truth) > 50
    rms = math.sqrt(float(np.mean(np.abs(want) ** 2)))
    assert np.abs(got - want).max() / rms < 1e-4
    host, _ = synth.wideband_capture_gpu(seconds=0.0085, kind=kind, seed=4100, gap=gap, esn0_db=None, to_host=True)
    assert np.array_equal(host, got)                     # host and device output paths, chunking: same samples
    del torch


@pytest.mark.gpu
def test_gpu_capture_noise_and_round_trip():
    from snout_b200.engine import RxEngine
    clean, truth = synth.wideband_capture_gpu(seconds=0.02, kind="ble", seed=5300, esn0_db=None, gap=(400, 5000))
    noisy, _ = synth.wideband_capture_gpu(seconds=0.02, kind="ble", seed=5300, esn0_db=25.0, gap=(400, 5000))
    again, _ = synth.wideband_capture_gpu(seconds=0.02, kind="ble", seed=5300, esn0_db=25.0, gap=(400, 5000))
    assert bool((noisy == again).all())                  # a function of (schedule, seed) only
    w = (noisy - clean).cpu().numpy()
    sigma2 = 4.0 * chanplan.WB_DECIM / (10.0 ** 2.5)     # synth.wideband_capture: sps * 24 / EsN0, sps = 4 for BLE
    assert abs(w.real.var() / (sigma2 / 2) - 1) < 0.01 and abs(w.imag.var() / (sigma2 / 2) - 1) < 0.01
    assert abs(w.mean()) < 4 * math.sqrt(sigma2 / len(w))
    assert abs(np.mean(w.real * w.imag)) < 0.01 * sigma2 and abs(np.mean(w[1:] * np.conj(w[:-1]))) < 0.01 * sigma2
    assert abs(np.mean(np.abs(w.real) > 2 * math.sqrt(sigma2 / 2)) - 0.0455) < 0.003        # Gaussian tails
    with RxEngine("ble_wb40", max_samples=len(noisy)) as e:
        fr = e.run(noisy)
    sent = {(t.channel, bytes(t.data)) for t in truth if t.start + 2000 < len(noisy) // 24}
    dec = {(int(f["channel"]), bytes(f["bytes"][: f["len"]])) for f in fr if f["crc_ok"]}
    assert len(sent) > 600 and len(sent & dec) >= 0.97 * len(sent)
    # mixed BLE + 802.15.4 capture (the two protocols overlap in frequency, so not every frame survives): the engine decodes
    # the GPU-made capture and the numpy-made one to the same frames
    g, _ = synth.wideband_capture_gpu(seconds=0.02, kind="mixed", seed=5300, esn0_db=None, gap=(400, 5000))
    c = synth.wideband_capture(seconds=0.02, kind="mixed", seed=5300, esn0_db=300.0, gap=(400, 5000)).iq
    with RxEngine("mixed_wb56", max_samples=len(c)) as e:
        a, b = e.run(g), e.run(c)
    key = lambda f: {(int(r["channel"]), int(r["proto"]), int(r["sample_index"]), bytes(r["bytes"][: r["len"]])) for r in f if r["crc_ok"]}   # noqa: E731
    assert len(key(b)) > 250 and (key(b) & {k for k in key(b) if k[1] == 2}) and len(key(a) ^ key(b)) <= 0.01 * len(key(b))
