"""The reference's RDS sketch (src/project.cpp:200-271: rds_thread, compiled but never started): 54-60 kHz
channel band-pass, squarer, 113.5-114.5 kHz band-pass, PLL(114000, bp_fs, 0.5, 0, 0.01), channel delay, mixer.
CPU: the oracle's restatement against the reference's own operators called in rds_thread's order.
GPU: the product (fmrx_rds_*: operator FIR kernel, k_pll unchanged, two element-wise kernels) against the oracle."""
import numpy as np
import pytest

import pyoracle
from conftest import assert_bits_equal


def _demod_with_rds(port, synth, n_blocks, seed=0):
    """Demodulated FM (the chain's own `demod` stage) with a 57 kHz BPSK-ish subcarrier added, so that the sketch's
    squarer finds a 114 kHz line to lock its PLL to."""
    info = port.mode(0, 51)
    iq = synth.synth_iq_exact(n_blocks * info.block_size // 2, float(info.rf_fs), station=seed)
    _, d = port.chain(0, 51).run(iq, ("demod",))
    x = d["demod"]
    t = np.arange(len(x), dtype=np.float64) / info.if_fs
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, len(x) // 202 + 2) * 2 - 1                   # 1187.5 baud at 240 kHz: 202 samples per bit
    data = np.repeat(bits, 202)[:len(x)].astype(np.float64)
    return (x + 0.05 * data * np.cos(2 * np.pi * 57000.0 * t)).astype(np.float32), info


@pytest.mark.parametrize("taps,delay", [(51, 5), (101, 12)])
def test_oracle_rds_sketch_matches_reference_operators(port, reference, synth, taps, delay):
    x, info = _demod_with_rds(port, synth, 40)
    a = pyoracle.RdsSketch(port, float(info.if_fs), taps, delay)
    b = pyoracle.RdsSketch(reference, float(info.if_fs), taps, delay)
    for blk in range(40):
        seg = x[blk * info.if_per_block:(blk + 1) * info.if_per_block]
        oa, ca, na = a.block(seg)
        ob, cb, nb = b.block(seg)
        assert_bits_equal(ca, cb, f"channel, block {blk}")
        assert_bits_equal(na, nb, f"carrier after the PLL, block {blk}")
        assert_bits_equal(oa, ob, f"mixer, block {blk}")
    assert np.abs(oa).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("taps,delay,block_mult", [(51, 5, 1), (101, 12, 3), (51, 0, 2)])
def test_cuda_rds_sketch_matches_oracle(fm, port, synth, taps, delay, block_mult):
    x, info = _demod_with_rds(port, synth, 60, seed=3)
    ref = pyoracle.RdsSketch(port, float(info.if_fs), taps, delay)
    n = info.if_per_block * block_mult
    with fm.RdsFront(float(info.if_fs), taps, delay) as rds:
        for blk in range(len(x) // n):
            seg = x[blk * n:(blk + 1) * n]
            go, gc, _ = rds.process(seg)
            oo, oc, _ = ref.block(seg)
            assert_bits_equal(gc, oc, f"channel, block {blk}")
            assert_bits_equal(go, oo, f"mixer, block {blk}")
        with pytest.raises(fm.FmrxError):
            rds.process(x[:taps - 2])                                     # shorter than the filters: the reference reads out of bounds
