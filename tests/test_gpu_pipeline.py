"""GPU parity of the fused pipeline (fmrx_create / fmrx_process) against the
oracle's block loop: every intermediate bit-exact, PCM identical, across modes,
tap counts, chunkings, call boundaries, captures and state hand-off."""
import numpy as np
import pytest

import pyoracle
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu

ALL_STAGES = ("i_ds", "q_ds", "demod", "chan", "pilot", "trig", "nco", "mixer",
              "mono", "mono_shift", "stereo", "left", "right")


def _capture(synth, info, nblocks, seed, extra=0):
    return synth.synth_iq(nblocks * info.block_size // 2 + extra, info.rf_fs, seed=seed)


@pytest.mark.parametrize("mode,taps,nblocks,chunk", [
    (0, 51, 24, 0), (0, 51, 24, 5), (0, 101, 12, 4), (0, 301, 8, 3),
    (1, 51, 18, 7), (1, 101, 6, 0), (2, 51, 2, 1), (3, 51, 2, 1), (2, 101, 2, 0),
])
def test_all_stages_bitwise(fm, port, synth, mode, taps, nblocks, chunk):
    info = port.mode(mode, taps)
    iq = _capture(synth, info, nblocks, seed=mode * 7 + taps)
    opcm, od = port.chain(mode, taps).run(iq, ALL_STAGES)
    with fm.Pipeline(mode, taps, 1, chunk_blocks=chunk, keep_stages=True) as p:
        assert p.info.block_size == info.block_size and p.info.audio_per_block == info.audio_per_block
        gpcm, gd = p.process_stages(iq, ALL_STAGES)
    for s in ALL_STAGES:
        assert_bits_equal(gd[s][0], od[s], f"mode {mode} taps {taps} stage {s}")
    assert np.array_equal(gpcm[0], opcm)


@pytest.mark.parametrize("mode,taps,nblocks", [(0, 7, 5), (0, 29, 5), (0, 30, 5), (0, 31, 5), (0, 60, 5), (0, 61, 5), (0, 151, 6), (0, 256, 5),
                                               (0, 512, 5), (1, 13, 5), (1, 200, 4), (3, 37, 1), (2, 64, 2)])
def test_unusual_tap_counts(fm, port, synth, mode, taps, nblocks):
    """`project --taps N` takes any N in 7..512.  The FIR kernels pad their tap loops (K1 to a multiple of 3*decim,
    K2 to a multiple of 12) and size their windows and halos from N: counts at, just below and just above the
    padding periods, tiny and maximal, for every decimation the modes have."""
    info = port.mode(mode, taps)
    iq = _capture(synth, info, nblocks, seed=taps)
    stages = ("demod", "chan", "pilot", "trig", "mixer", "mono", "stereo")
    opcm, od = port.chain(mode, taps).run(iq, stages)
    with fm.Pipeline(mode, taps, 1, chunk_blocks=2, keep_stages=True) as p:
        gpcm, gd = p.process_stages(iq, stages)
    for s_ in stages:
        assert_bits_equal(gd[s_][0], od[s_], f"mode {mode} taps {taps} stage {s_}")
    assert np.array_equal(gpcm[0], opcm)


def test_partial_trailing_block_is_dropped(fm, port, synth):
    info = port.mode(0, 51)
    iq = _capture(synth, info, 6, seed=2, extra=1234)
    opcm, _ = port.chain(0, 51).run(iq)
    with fm.Pipeline(0, 51, 1) as p:
        g = p.process(iq)
    assert g.shape[1] == 6 * 2 * info.audio_per_block
    assert np.array_equal(g[0], opcm)


def test_streaming_calls_equal_one_call(fm, port, synth):
    """State carried across fmrx_process calls == the reference's loop-carried state,
    including the shared audio-state quirk at every block boundary."""
    info = port.mode(0, 51)
    iq = _capture(synth, info, 30, seed=5)
    opcm, _ = port.chain(0, 51).run(iq)
    with fm.Pipeline(0, 51, 1, chunk_blocks=4) as p:
        parts = []
        for lo, hi in ((0, 1), (1, 2), (2, 9), (9, 10), (10, 30)):
            parts.append(p.process(iq[lo * info.block_size:hi * info.block_size])[0])
        assert np.array_equal(np.concatenate(parts), opcm)
        p.reset()
        assert np.array_equal(p.process(iq)[0], opcm)


def test_batched_captures_are_independent(fm, port, synth):
    info = port.mode(0, 51)
    C = 5
    iq = np.stack([_capture(synth, info, 10, seed=100 + c) for c in range(C)])
    with fm.Pipeline(0, 51, C, chunk_blocks=3) as p:
        g = p.process(iq)
        launches = p.kernel_launches
    assert launches == 4 * 4          # 4 kernels per chunk, ceil(10/3) chunks
    for c in range(C):
        o, _ = port.chain(0, 51).run(iq[c])
        assert np.array_equal(g[c], o), f"capture {c}"


def test_state_handoff_between_pipelines(fm, port, synth):
    """Time-sharding: shard 1 starts from shard 0's state blob and continues bit-exactly."""
    info = port.mode(0, 51)
    iq = _capture(synth, info, 16, seed=8)
    opcm, _ = port.chain(0, 51).run(iq)
    cut = 7 * info.block_size
    with fm.Pipeline(0, 51, 1) as a, fm.Pipeline(0, 51, 1) as b:
        first = a.process(iq[:cut])[0]
        blob = a.get_state()
        assert len(blob) == a.state_size()
        b.set_state(blob)
        second = b.process(iq[cut:])[0]
        assert_bits_equal(b.pll_state(), port_pll_state(port, iq), "pll state after hand-off")
    assert np.array_equal(np.concatenate([first, second]), opcm)


def port_pll_state(port, iq):
    ch = port.chain(0, 51)
    ch.run(iq)
    st = ch.get_state()
    t1 = 50
    off = 4 * t1 + 2
    return st[off:off + 6]


def test_state_blob_rejected_by_other_mode(fm):
    with fm.Pipeline(0, 51, 1) as a, fm.Pipeline(1, 51, 1) as b:
        with pytest.raises(fm.FmrxError) as e:
            b.set_state(a.get_state())
        assert e.value.status == fm.ERR_STATE


def test_pll_saturation_inside_pipeline(fm, port, synth):
    """Start the PLL counter just below 2^24 via the state blob on both sides."""
    import struct
    info = port.mode(0, 51)
    iq = _capture(synth, info, 6, seed=12)
    with fm.Pipeline(0, 51, 1) as p:
        p.process(iq[:2 * info.block_size])
        blob = bytearray(p.get_state())
        # the 8 PLL floats are the last 32 bytes; trigOffset is float #5
        struct.pack_into("<f", blob, len(blob) - 32 + 5 * 4, 16777216.0 - 900.0)
        p.set_state(bytes(blob))
        g = p.process(iq[2 * info.block_size:])[0]
        gst = p.pll_state()
    ch = port.chain(0, 51)
    ch.run(iq[:2 * info.block_size])
    st = ch.get_state()
    st[4 * 50 + 2 + 5] = 16777216.0 - 900.0
    ch.set_state(st)
    o, _ = ch.run(iq[2 * info.block_size:])
    assert gst[5] == 16777216.0
    assert np.array_equal(g, o)


def test_device_pointer_entry(fm, port, synth):
    torch = pytest.importorskip("torch")
    info = port.mode(0, 51)
    C, nb = 3, 9
    iq = np.stack([_capture(synth, info, nb, seed=40 + c) for c in range(C)])
    d_iq = torch.from_numpy(iq).cuda()
    d_pcm = torch.zeros((C, nb * 2 * info.audio_per_block), dtype=torch.int16, device="cuda")
    with fm.Pipeline(0, 51, C, chunk_blocks=4) as p:
        s = torch.cuda.current_stream()
        p.process_device(d_iq.data_ptr(), d_iq.stride(0), nb, d_pcm.data_ptr(), d_pcm.stride(0), s.cuda_stream)
        s.synchronize()
    g = d_pcm.cpu().numpy()
    for c in range(C):
        o, _ = port.chain(0, 51).run(iq[c])
        assert np.array_equal(g[c], o)


def test_long_capture_roundtrip_properties(fm, port, synth):
    """A longer capture (10 s): PCM equals the oracle, the decoded tones are the ones
    that were modulated (R = 3 kHz on channel 0, L = 1 kHz on channel 1)."""
    info = port.mode(0, 51)
    nb = 1875       # 10 s
    iq = synth.synth_iq(nb * info.block_size // 2, info.rf_fs, seed=0, f_l=1000.0, f_r=3000.0)
    opcm, _ = port.chain(0, 51).run(iq)
    with fm.Pipeline(0, 51, 1) as p:
        g = p.process(iq)[0]
    assert np.array_equal(g, opcm)
    x = g.astype(np.float64).reshape(-1, 2)[48000:48000 * 5]
    f = np.fft.rfftfreq(len(x), 1 / 48000.0)
    for ch, tone in ((0, 3000.0), (1, 1000.0)):
        spec = np.abs(np.fft.rfft(x[:, ch] * np.hanning(len(x))))
        assert abs(f[np.argmax(spec)] - tone) < 2.0
