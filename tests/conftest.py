import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu() -> bool:
    try:
        from snout_b200 import _abi
        return _abi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a device must fail loudly, not skip: the product has no CPU path
    pass


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def emu():
    import ctypes
    path = os.path.join(ROOT, "tests", "emu", "libemu.so")
    r = subprocess.run(["make", "-C", os.path.dirname(path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return ctypes.CDLL(path)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


FRAME_KEYS = ("sample_index", "window", "channel", "proto", "crc_ok", "lqi", "phase", "len", "access_addr", "bytes")


def assert_frames_equal(got, want, keys=FRAME_KEYS, what=""):
    assert len(got) == len(want), f"{what}: {len(got)} frames, expected {len(want)}"
    for k in keys:
        if not np.array_equal(got[k], want[k]):
            bad = np.nonzero(np.any(np.atleast_2d(got[k] != want[k]).reshape(len(got), -1), axis=1))[0][:5]
            raise AssertionError(f"{what}: field {k} differs at frames {bad.tolist()}: {got[k][bad]} vs {want[k][bad]}")


def ble_expected(oracle_mod, q8, channel, **kw):
    """Expected BLE frames of one int8 channel stream: the restatement (oracle/ble_oracle.c), and -- wherever the UNMODIFIED
    btle_rx.c was compiled (oracle/_ref/libbtle_ref.so travels to the GPU box as a prebuilt file) -- the reference itself,
    which must say the same."""
    want = oracle_mod.ble_decode(q8, channel, **kw)
    if oracle_mod.have_ref("btle_ref") and not (set(kw) - {"aa", "crc_init", "aa_mask"}):
        assert_frames_equal(oracle_mod.ble_decode(q8, channel, impl="reference", **kw), want, what=f"unmodified btle_rx.c vs its restatement, channel {channel}")
    return want
