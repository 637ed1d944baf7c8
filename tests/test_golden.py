"""Golden fixtures (tests/golden/*.npz, generated FROM THE REFERENCE by
tests/golden/make_golden.py) against (a) the oracle, on the CPU, and (b) the CUDA
path through the C ABI, on the GPU.  The reference checkout is not needed here."""
import hashlib
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import assert_bits_equal

GOLD = Path(__file__).resolve().parent / "golden"
STAGES = ("i_ds", "q_ds", "demod", "chan", "pilot", "nco", "mixer", "mono", "mono_shift",
          "stereo", "left", "right")
CHAIN_FILES = sorted(p.name for p in GOLD.glob("chain_*.npz"))


def _load_chain(name, synth, mode_info):
    g = np.load(GOLD / name)
    mode, taps, nb, seed = int(g["mode"]), int(g["taps"]), int(g["blocks"]), int(g["seed"])
    info = mode_info(mode, taps)
    if "iq" in g:
        iq = g["iq"]
    else:
        iq = synth.synth_iq(nb * info.block_size // 2, info.rf_fs, seed=seed)
    digest = np.frombuffer(hashlib.sha256(iq.tobytes()).digest(), np.uint8)
    if not np.array_equal(digest, g["iq_sha256"]):
        pytest.skip("synthetic input is not reproducible on this numpy build (SHA-256 differs)")
    return g, mode, taps, nb, iq


def _check_stages(g, got, what):
    for s in STAGES:
        if "stage_" + s in g:
            assert_bits_equal(got[s], g["stage_" + s], f"{what} stage {s}")
        else:
            assert_bits_equal(got[s][:4096], g[f"stage_{s}_head"], f"{what} stage {s} (head)")
            d = np.frombuffer(hashlib.sha256(np.ascontiguousarray(got[s]).tobytes()).digest(), np.uint8)
            assert np.array_equal(d, g[f"stage_{s}_sha256"]), f"{what} stage {s}: SHA-256 differs"


def test_fixtures_present():
    assert len(CHAIN_FILES) >= 6 and (GOLD / "taps.npz").exists() and (GOLD / "pll_kat.npz").exists()


@pytest.mark.parametrize("name", CHAIN_FILES)
def test_oracle_matches_reference_golden(name, port, synth):
    g, mode, taps, nb, iq = _load_chain(name, synth, port.mode)
    pcm, d = port.chain(mode, taps).run(iq, STAGES)
    _check_stages(g, d, f"oracle {name}")
    assert np.array_equal(pcm, g["pcm"])


def test_oracle_taps_match_reference_golden(port):
    g = np.load(GOLD / "taps.npz")
    for key in g.files:
        f = key.split("_")
        if f[0] == "lpf":
            got = port.lpf_taps(float(f[1]), float(f[2]), int(f[3]), int(f[4]))
        else:
            got = port.bpf_taps(float(f[1]), float(f[2]), float(f[3]), int(f[4]))
        assert_bits_equal(got, g[key], key)


def test_product_tap_design_matches_reference_golden(pkg):
    """fmrx_impulse_response_lpf/bpf are host code: checked without a GPU."""
    fm = pkg.binding
    g = np.load(GOLD / "taps.npz")
    for key in g.files:
        f = key.split("_")
        if f[0] == "lpf":
            got = fm.impulseResponseLPF(float(f[1]), float(f[2]), int(f[3]), int(f[4]))
        else:
            got = fm.impulseResponseBPF(float(f[1]), float(f[2]), float(f[3]), int(f[4]))
        assert_bits_equal(got, g[key], key)


def test_oracle_pll_matches_reference_golden(port):
    g = np.load(GOLD / "pll_kat.npz")
    nco, _, st = port.pll(g["pilot"], 19000, 240e3, 2, 0, 0.01)
    assert_bits_equal(nco, g["nco"], "pll nco")
    assert_bits_equal(st, g["state"], "pll state")
    nco, _, st = port.pll(g["pilot"][:2000], 19000, 240e3, 2, 0, 0.01, g["sat_state_in"])
    assert_bits_equal(nco, g["sat_nco"], "pll nco (saturating counter)")
    assert_bits_equal(st, g["sat_state"], "pll state (saturating counter)")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CHAIN_FILES)
def test_cuda_pipeline_matches_reference_golden(name, fm, synth):
    g, mode, taps, nb, iq = _load_chain(name, synth, fm.mode_table)
    with fm.Pipeline(mode, taps, 1, keep_stages=True) as p:
        pcm, d = p.process_stages(iq, STAGES)
        pll = p.pll_state()
    _check_stages(g, {k: v[0] for k, v in d.items()}, f"cuda {name}")
    assert np.array_equal(pcm[0], g["pcm"])
    assert_bits_equal(pll, g["pll_state"], "pll state")


@pytest.mark.gpu
def test_cuda_pll_matches_reference_golden(fm):
    g = np.load(GOLD / "pll_kat.npz")
    nco, st = fm.PLL(g["pilot"], 19000, 240e3, 2, 0, 0.01)
    assert_bits_equal(nco, g["nco"], "pll nco")
    assert_bits_equal(st, g["state"], "pll state")
    nco, st = fm.PLL(g["pilot"][:2000], 19000, 240e3, 2, 0, 0.01, g["sat_state_in"])
    assert_bits_equal(nco, g["sat_nco"], "pll nco (saturating counter)")
    assert_bits_equal(st, g["sat_state"], "pll state (saturating counter)")
