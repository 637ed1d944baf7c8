"""How fast is the NCCL path between the ranks of this box?  (torchrun --nproc-per-node N tools/nccl_probe.py)
Times the PCM gather bench.py does (737 MB of int32 per rank to rank 0) and a ring send/recv."""
import os, time
import torch, torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
n = 737_280_000 // 4
x = torch.full((n,), rank, dtype=torch.int32, device="cuda")
outs = [torch.empty_like(x) for _ in range(world)] if rank == 0 else None
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    dist.gather(x, outs, dst=0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if rank == 0:
        print(f"gather of {world - 1} x {x.numel() * 4 / 1e6:.0f} MB to rank 0: {dt * 1e3:.1f} ms = {(world - 1) * x.numel() * 4 / dt / 1e9:.1f} GB/s", flush=True)
y = torch.empty_like(x)
for it in range(2):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    ops = [dist.P2POp(dist.isend, x, (rank + 1) % world), dist.P2POp(dist.irecv, y, (rank - 1) % world)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if rank == 0:
        print(f"ring send/recv of {x.numel() * 4 / 1e6:.0f} MB per rank: {dt * 1e3:.1f} ms = {x.numel() * 4 / dt / 1e9:.1f} GB/s per link", flush=True)
dist.destroy_process_group()
