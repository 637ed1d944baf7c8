"""One capture time-sharded over devices (fmrx_long_*, the C++ host orchestration in libfmrx_b200.so):
feed-forward stages of all shards from FIR halos at once, PLL state handed from shard to shard with peer
copies, PCM gathered on the first device.  The result must be bit-identical to the oracle's single pass --
and so to one Pipeline over the whole capture.  On a one-GPU box the shards share the device (the halo and
hand-off logic is the same); with more GPUs visible each shard gets its own."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _devices(fm, n):
    have = fm.device_count()
    return [r % have for r in range(n)]


@pytest.mark.parametrize("mode,taps,n_blocks,n_shards", [(0, 51, 23, 3), (0, 301, 13, 4), (1, 101, 9, 2), (2, 51, 5, 2), (0, 51, 4, 4)])
def test_time_sharded_capture_is_bit_identical(fm, port, synth, mode, taps, n_blocks, n_shards):
    info = port.mode(mode, taps)
    iq = synth.synth_iq_exact(n_blocks * info.block_size // 2, float(info.rf_fs), station=mode + taps)
    want, _ = port.chain(mode, taps).run(iq)
    ch = port.chain(mode, taps)
    ch.run(iq)
    with fm.LongCapture(mode, taps, _devices(fm, n_shards), n_blocks) as lc:
        shards = [lc.shard(r) for r in range(n_shards)]
        assert shards[0][0] == 0 and shards[0][2] == 0 and sum(s[1] for s in shards) == n_blocks
        assert all(shards[r][0] == shards[r - 1][0] + shards[r - 1][1] for r in range(1, n_shards))
        got = lc.process(iq)
        pll = lc.pll_state()
        assert np.array_equal(got, want), f"first difference at PCM index {int(np.nonzero(got != want)[0][0])}"
        assert np.array_equal(lc.process(iq), want)        # and again: the handle is reusable
    st = ch.get_state()
    off = 4 * (taps - 1) + 2
    assert np.array_equal(pll.view(np.uint32), st[off:off + 6].view(np.uint32))


def test_time_sharded_device_buffers_and_timing(fm, port, synth):
    torch = pytest.importorskip("torch")
    mode, taps, n_blocks, n_shards = 0, 51, 40, 3
    info = port.mode(mode, taps)
    iq = synth.synth_iq_exact(n_blocks * info.block_size // 2, float(info.rf_fs), station=9)
    want, _ = port.chain(mode, taps).run(iq)
    devs = _devices(fm, n_shards)
    with fm.LongCapture(mode, taps, devs, n_blocks) as lc:
        bufs = []
        for r in range(n_shards):
            first, cnt, halo = lc.shard(r)
            piece = iq[(first - halo) * info.block_size:(first + cnt) * info.block_size]
            bufs.append(torch.from_numpy(piece.copy()).to(f"cuda:{devs[r]}"))
        pcm = torch.zeros(n_blocks * 2 * info.audio_per_block, dtype=torch.int16, device=f"cuda:{devs[0]}")
        torch.cuda.synchronize()
        ms = lc.process_device([b.data_ptr() for b in bufs], pcm.data_ptr())
        assert ms > 0
        assert np.array_equal(pcm.cpu().numpy(), want)
