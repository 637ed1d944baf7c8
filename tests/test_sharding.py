"""N>1 host logic on CPU ranks (gloo, world_size 2 and 3): capture sharding, PCM
gather, and the time-shard chain with state hand-off.  The engine on these ranks is a
stand-in backed by the oracle (tests may use it); on a GPU box the same functions are
driven with binding.Pipeline (tests/test_gpu_multi.py)."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG_NAME = "software-defined-radio-course-project_b200"


class OracleEngine:
    """Same three methods as binding.Pipeline, computed by the oracle chain(s)."""

    def __init__(self, mode, taps, n_captures=1):
        sys.path.insert(0, str(ROOT / "oracle"))
        import pyoracle
        self.port = pyoracle.Port()
        self.chains = [self.port.chain(mode, taps) for _ in range(n_captures)]

    def process(self, iq):
        iq = np.atleast_2d(iq)
        return np.stack([c.run(row)[0] for c, row in zip(self.chains, iq)])

    def get_state(self):
        return self.chains[0].get_state().tobytes()

    def set_state(self, blob):
        self.chains[0].set_state(np.frombuffer(blob, np.float32))


def test_shard_ranges_cover_and_balance(pkg):
    par = pkg.parallel
    for n in (0, 1, 7, 64, 1350000):
        for world in (1, 2, 3, 8):
            r = [par.shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    assert par.capture_shards(64, 8) == [(8 * k, 8 * k + 8) for k in range(8)]
    with pytest.raises(ValueError):
        par.shard_range(4, 2, 2)


def _worker(rank, world, port_no, mode, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    pkg = importlib.import_module(PKG_NAME)
    par, synth = pkg.parallel, pkg.synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        probe = OracleEngine(mode, 51)
        info = probe.chains[0].info
        # --- independent captures: 5 captures over `world` ranks, ragged shards ---
        n_cap, nb = 5, 6
        lo, hi = par.shard_range(n_cap, world, rank)
        iq = np.stack([synth.synth_iq(nb * info.block_size // 2, info.rf_fs, seed=200 + c) for c in range(lo, hi)]) \
            if hi > lo else np.zeros((0, nb * info.block_size), np.uint8)
        eng = OracleEngine(mode, 51, hi - lo)
        parts = par.gather_pcm(eng.process(iq) if hi > lo else np.zeros((0, nb * 2 * info.audio_per_block), np.int16))
        if rank == 0:
            got = np.concatenate(parts)
            want = np.stack([OracleEngine(mode, 51).process(synth.synth_iq(nb * info.block_size // 2, info.rf_fs, seed=200 + c))[0]
                             for c in range(n_cap)])
            assert np.array_equal(got, want), "batched gather differs from single-rank result"
        # --- one capture, time-sharded chain with state hand-off ---
        total_blocks = 11
        whole = synth.synth_iq(total_blocks * info.block_size // 2, info.rf_fs, seed=300)
        b0, b1 = par.time_shards(total_blocks, world)[rank]
        pcm = par.run_time_sharded(OracleEngine(mode, 51), whole[b0 * info.block_size:b1 * info.block_size])
        if rank == 0:
            want = OracleEngine(mode, 51).process(whole)[0]
            assert np.array_equal(pcm, want), "time-sharded chain differs from a single pass"
            Path(tmp).write_text("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, 0), (3, 1)])
def test_gather_and_time_shard_chain_gloo(tmp_path, world, mode):
    import torch.multiprocessing as mp
    marker = tmp_path / "done"
    port_no = 29600 + world * 10 + mode
    mp.spawn(_worker, args=(world, port_no, mode, str(marker)), nprocs=world, join=True)
    assert marker.read_text() == "ok"
