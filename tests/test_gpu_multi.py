"""Multi-GPU host logic with the real engine (binding.Pipeline) over NCCL: runs only
when the box has >= 2 GPUs (gpurun --gpus 2); otherwise the same logic is covered by the
gloo tests in tests/test_sharding.py."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
PKG_NAME = "software-defined-radio-course-project_b200"


def _worker(rank, world, port_no, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    import torch
    import torch.distributed as dist
    import pyoracle
    pkg = importlib.import_module(PKG_NAME)
    fm, par, synth = pkg.binding, pkg.parallel, pkg.synth
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        port = pyoracle.Port()
        info = port.mode(0, 51)
        total = 21
        whole = synth.synth_iq(total * info.block_size // 2, info.rf_fs, seed=400)
        b0, b1 = par.time_shards(total, world)[rank]
        with fm.Pipeline(0, 51, 1, device=rank) as eng:
            class One:                                     # Pipeline.process returns [1, n]
                def process(self, iq): return eng.process(iq)[0]
                def get_state(self): return eng.get_state()
                def set_state(self, b): eng.set_state(b)
            pcm = par.run_time_sharded(One(), whole[b0 * info.block_size:b1 * info.block_size], device=dev)
        n_cap = 5
        lo, hi = par.shard_range(n_cap, world, rank)
        iq = np.stack([synth.synth_iq(6 * info.block_size // 2, info.rf_fs, seed=500 + c) for c in range(lo, hi)])
        with fm.Pipeline(0, 51, hi - lo, device=rank) as eng:
            parts = par.run_capture_batch(eng, iq, device=dev)
        if rank == 0:
            assert np.array_equal(pcm, port.chain(0, 51).run(whole)[0])
            got = np.concatenate(parts)
            for c in range(n_cap):
                want = port.chain(0, 51).run(synth.synth_iq(6 * info.block_size // 2, info.rf_fs, seed=500 + c))[0]
                assert np.array_equal(got[c], want)
            Path(tmp).write_text("ok")
    finally:
        dist.destroy_process_group()


def test_time_shard_chain_and_gather_over_nccl(fm, tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs on the box (covered on CPU ranks by tests/test_sharding.py)")
    marker = tmp_path / "done"
    mp.spawn(_worker, args=(world, 29733, str(marker)), nprocs=world, join=True)
    assert marker.read_text() == "ok"
