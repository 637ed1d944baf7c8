"""The C-ABI library without a GPU: it loads, exports every symbol include/fmrx.h
declares, answers host-only queries, and refuses to compute when no device exists
(there is no CPU fallback in the product)."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "fmrx.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmrx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(pkg):
    fm = pkg.binding
    lib = fm.load()
    declared = _declared_symbols()
    assert len(declared) >= 28
    out = subprocess.run(["nm", "-D", "--defined-only", str(fm.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (fmrx_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/fmrx.h but not exported: {missing}"
    assert sorted(fm.ABI_SYMBOLS) == declared, "binding.ABI_SYMBOLS out of date with include/fmrx.h"
    assert lib.fmrx_abi_version() == 1


def test_library_is_sm100a_cuda_code(pkg):
    """The shipped library carries sm_100a SASS for the hot-path kernels."""
    out = subprocess.run(["cuobjdump", "-lelf", str(pkg.binding.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN4fmrx5k_pllENS_7PllArgsE", str(pkg.binding.LIB_PATH)],
                          capture_output=True, text=True).stdout
    assert "DFMA" in sass and "Function : _ZN4fmrx5k_pllENS_7PllArgsE" in sass


def test_mode_table_matches_oracle(pkg, port):
    for mode in range(4):
        for taps in (51, 101, 301):
            a, b = pkg.binding.mode_table(mode, taps), port.mode(mode, taps)
            assert a.__dict__ == b.__dict__
    with pytest.raises(pkg.binding.FmrxError):
        pkg.binding.mode_table(4, 51)
    with pytest.raises(pkg.binding.FmrxError):
        pkg.binding.mode_table(0, 5)
    assert pkg.binding.mode_table(0, 0).taps == 51          # 0 -> the binary's default


def test_mode_table_values_from_reference_source(pkg):
    """src/project.cpp:327-364, literal expectations."""
    m = pkg.binding.mode_table
    assert (m(0).block_size, m(0).if_per_block, m(0).audio_per_block) == (12800, 640, 128)
    assert (m(1).block_size, m(1).if_per_block, m(1).audio_per_block) == (6144, 768, 128)
    assert (m(2).block_size, m(2).if_per_block, m(2).audio_per_block, m(2).audio_taps, m(2).if_fs) == \
        (2048000, 102400, 18816, 7497, 35280000)
    assert (m(3).block_size, m(3).if_per_block, m(3).audio_per_block, m(3).audio_taps, m(3).if_fs) == \
        (5898240, 327680, 56448, 22491, 112896000)


def test_no_cpu_fallback(pkg):
    fm = pkg.binding
    if fm.device_count() > 0:
        pytest.skip("a GPU is visible; the refusal path is for boxes without one")
    with pytest.raises(fm.FmrxError) as e:
        fm.Pipeline(0, 51, 1)
    assert e.value.status == fm.ERR_NO_DEVICE
    with pytest.raises(fm.FmrxError):
        fm.mixer(np.ones(4, np.float32), np.ones(4, np.float32))
    with pytest.raises(fm.FmrxError):
        fm.PLL(np.ones(4, np.float32), 19000, 240e3)


def test_argument_errors_do_not_need_a_device(pkg):
    fm = pkg.binding
    lib = fm.load()
    assert lib.fmrx_create(None, None) == fm.ERR_ARG
    cfg = fm.ConfigStruct(0, 51, 0, -1, 0, 0)                 # n_captures = 0
    h = ctypes.c_void_p()
    assert lib.fmrx_create(ctypes.byref(h), ctypes.byref(cfg)) == fm.ERR_ARG
    cfg = fm.ConfigStruct(7, 51, 1, -1, 0, 0)                 # bad mode
    assert lib.fmrx_create(ctypes.byref(h), ctypes.byref(cfg)) == fm.ERR_ARG
    assert lib.fmrx_strerror(fm.ERR_NO_DEVICE).decode().startswith("no usable CUDA")
    assert lib.fmrx_state_size(None) == 0
    assert lib.fmrx_destroy(None) == fm.OK
