"""The C-ABI library loads and exports every symbol include/snoutrx.h declares; the product tree
never touches the oracle; no compute is called here (no GPU needed)."""
import ctypes
import os
import re

import numpy as np

from conftest import ROOT
from snout_b200 import _abi, chanplan


def _header_functions():
    src = open(os.path.join(ROOT, "include", "snoutrx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snrx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/snoutrx.h but not exported"
    assert sorted(n for n, _, _ in _abi.SYMBOLS) == declared
    assert lib.snrx_abi_version() == 2


def test_struct_sizes_match_header():
    assert _abi.FRAME_DTYPE.itemsize == 160
    assert ctypes.sizeof(_abi.Config) == 72 and ctypes.sizeof(_abi.Shard) == 24 and ctypes.sizeof(_abi.Stats) == 40


def test_channel_plan_helpers():
    lib = _abi.load()
    for ch in range(40):
        assert lib.snrx_ble_channel_mhz(ch) == chanplan.ble_channel_mhz(ch)
        assert lib.snrx_ble_channel_bin(ch) == chanplan.ble_channel_bin(ch)
    for ch in range(11, 27):
        assert lib.snrx_zigbee_channel_mhz(ch) == chanplan.zigbee_channel_mhz(ch)
        assert lib.snrx_zigbee_channel_bin(ch) == chanplan.zigbee_channel_bin(ch)
    assert lib.snrx_ble_channel_mhz(40) < 0 and lib.snrx_zigbee_channel_mhz(27) < 0
    assert lib.snrx_strerror(-2).decode().startswith("no usable CUDA device")
    h = _abi.pfb_prototype(_abi.MODE_BLE_WB40, 384)
    assert abs(h.sum() - 1.0) < 1e-12 and np.allclose(h, h[::-1])


def test_product_never_imports_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "snout_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M) or re.search(r'#include\s+"[^"]*oracle/', text):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_tables_identical_for_kernels_and_oracle():
    a = open(os.path.join(ROOT, "snout_b200", "csrc", "zb_tables.h")).read()
    b = open(os.path.join(ROOT, "oracle", "zb_tables.h")).read()
    assert a == b
