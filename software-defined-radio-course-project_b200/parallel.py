"""Host-side sharding of the receive chain over the GPUs of one box (one process per
GPU, ``torch.distributed``).  SURVEY.md section 8(e):

* independent captures/stations shard trivially -- each rank runs the whole chain on
  its own captures; the only collective is the gather of the PCM to one rank;
* one long capture is cut into consecutive runs of whole blocks; the feed-forward
  stages could run ahead, but the PLL is one sequential recurrence, so the shards run
  as a chain: rank r receives rank r-1's carried state (``get_state`` blob: IQ / demod /
  channel / trigArg tails, PLL scalars, block counter), continues bit-exactly, and
  passes its own state on.  No speed-up for the PLL is possible; this
  buys capacity and keeps results identical to a single pass.

The functions take an *engine* -- anything with ``process(iq) -> pcm``,
``get_state() -> bytes`` and ``set_state(bytes)`` (``binding.Pipeline`` on a GPU) -- so
the sharding and hand-off logic is testable on CPU ranks (gloo) with a stand-in engine.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of ``range(n_items)``: the first ``n % world`` ranks get one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def capture_shards(n_captures: int, world: int) -> list[tuple[int, int]]:
    return [shard_range(n_captures, world, r) for r in range(world)]


def time_shards(n_blocks: int, world: int) -> list[tuple[int, int]]:
    """Whole-block time shards of one capture, in stream order."""
    return [shard_range(n_blocks, world, r) for r in range(world)]


def _as_tensor(a, device):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def gather_pcm(pcm_local: np.ndarray, dst: int = 0, device="cpu", group=None):
    """Gather every rank's PCM (int16, any shape, sizes may differ) to ``dst``.
    Returns the list of arrays on ``dst`` (rank order), ``None`` elsewhere."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shape = torch.tensor(list(pcm_local.shape) + [0] * (4 - pcm_local.ndim), dtype=torch.int64, device=device)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    ndim = pcm_local.ndim
    sizes = [int(np.prod([int(v) for v in s[:ndim]])) for s in shapes]
    # moved as bytes: NCCL has no 16-bit integer type
    mine = _as_tensor(np.ascontiguousarray(pcm_local, np.int16).reshape(-1).view(np.uint8), device)
    out = None
    # point-to-point: sizes differ per rank, and this is an output funnel, not a reduction
    if rank == dst:
        out = []
        for r in range(world):
            if r == rank:
                out.append(pcm_local.copy())
                continue
            buf = torch.empty(2 * sizes[r], dtype=torch.uint8, device=device)
            if sizes[r]:
                dist.recv(buf, src=r, group=group)
            out.append(buf.cpu().numpy().view(np.int16).reshape([int(v) for v in shapes[r][:ndim]]))
    elif mine.numel():
        dist.send(mine, dst=dst, group=group)
    return out


def run_capture_batch(engine, iq_local: np.ndarray, dst: int = 0, device="cpu", group=None):
    """Independent captures: this rank's shard through its engine, PCM gathered to ``dst``."""
    pcm = engine.process(iq_local)
    return gather_pcm(np.asarray(pcm), dst=dst, device=device, group=group)


def run_time_sharded(engine, iq_shard: np.ndarray, dst: int = 0, device="cpu", group=None):
    """One capture cut into consecutive block runs, one per rank, processed as a chain with
    state hand-off.  Returns the whole capture's PCM on ``dst`` (``None`` elsewhere)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank > 0:
        n = torch.zeros(1, dtype=torch.int64, device=device)
        dist.recv(n, src=rank - 1, group=group)
        blob = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
        dist.recv(blob, src=rank - 1, group=group)
        engine.set_state(blob.cpu().numpy().tobytes())
    pcm = np.asarray(engine.process(iq_shard))
    if rank + 1 < world:
        blob = np.frombuffer(engine.get_state(), np.uint8)
        dist.send(torch.tensor([len(blob)], dtype=torch.int64, device=device), dst=rank + 1, group=group)
        dist.send(_as_tensor(blob, device), dst=rank + 1, group=group)
    parts = gather_pcm(pcm.reshape(-1), dst=dst, device=device, group=group)
    return np.concatenate(parts) if parts is not None else None
