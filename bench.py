#!/usr/bin/env python
"""Benchmark of the FM receive chain hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path

For N > 1 the driver launches it under torchrun (one rank per GPU, NCCL); RANK /
LOCAL_RANK / WORLD_SIZE come from the environment.  One JSON line is printed by rank 0.

Workload (config.workload): mode 0 stereo, the reference binary's 51 taps, a batch of
64 independent synthetic 60 s FM-stereo captures PER GPU (BASELINE.json configs[3];
weak scaling: captures are independent, so ranks share nothing on the data path; the
only collective is the NCCL gather of the PCM to rank 0, inside the timed region).
A "step" is one pass of the whole chain over the batch: 64 x 144 M IQ samples.

  value : IQ Msamples/s, inputs resident in HBM, CUDA-event timed, max over ranks.
  e2e   : same metric through the C ABI entry fmrx_process() with HOST (pinned)
          buffers: H2D of the u8 IQ and D2H of the int16 PCM inside the timed region.
  roofline     : the dominant kernel (k_pll, the per-capture PLL recurrence; >95 % of
                 the step) against measured HBM bandwidth.  The kernel is latency
                 bound by construction (one dependent chain per capture), so the
                 fraction is tiny; pll_ns_per_sample is the number that matters.
  cpu_baseline : the reference's own `project` binary (oracle/_ref, built from the
                 reference sources) on the box's host cores, one process per core, on
                 a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "input IQ Msamples/s (mode 0 stereo)"
UNIT = "Msamples/s"
MODE, TAPS = 0, 51
RF_FS = 2.4e6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--captures", type=int, default=64, help="captures per GPU")
    ap.add_argument("--seconds", type=float, default=60.0, help="seconds of signal per capture")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="signal per process for the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(captures, seconds):
    return (f"mode 0 stereo, {TAPS} taps, {captures} independent synthetic FM-stereo captures x "
            f"{seconds:g} s (2.4 Msps u8 IQ) per GPU")


# ----------------------------------------------------------------------------------------
# CPU reference arm / baseline
# ----------------------------------------------------------------------------------------

def _reference_binary():
    sys.path.insert(0, str(ROOT / "oracle"))
    import pyoracle
    return pyoracle, pyoracle.Reference.binary(TAPS)


def run_cpu_sample(iq_bytes: bytes, n_proc: int):
    """Time the reference's CPU path on `n_proc` host cores, one capture per process.
    Returns (msps, kind, seconds_wall)."""
    pyoracle, exe = _reference_binary()
    n_pairs = len(iq_bytes) // 2
    if exe is not None:
        t0 = time.perf_counter()
        procs = [subprocess.Popen([str(exe), str(MODE), "2"], stdin=subprocess.PIPE,
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                 for _ in range(n_proc)]

        def feed(p):
            try:
                p.stdin.write(iq_bytes)
                p.stdin.close()
            except BrokenPipeError:
                pass
            p.wait()
        th = [threading.Thread(target=feed, args=(p,)) for p in procs]
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        return n_proc * n_pairs / dt / 1e6, "reference", dt
    # the reference could not be compiled on this box: time the C restatement instead
    import numpy as np
    port = pyoracle.Port()
    iq = np.frombuffer(iq_bytes, np.uint8)
    chains = [port.chain(MODE, TAPS) for _ in range(n_proc)]
    t0 = time.perf_counter()
    th = [threading.Thread(target=c.run, args=(iq,)) for c in chains]   # ctypes drops the GIL
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return n_proc * n_pairs / dt / 1e6, "port", dt


def cpu_sample_input(seconds: float) -> bytes:
    pkg = importlib.import_module("software-defined-radio-course-project_b200")
    info_block = 12800
    n_blocks = max(8, int(seconds * RF_FS * 2 / info_block))
    return pkg.synth.synth_iq(n_blocks * info_block // 2, RF_FS, seed=0).tobytes()


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_proc = max(1, min(cores, args.captures))
    iq = cpu_sample_input(args.cpu_seconds)
    vals = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        msps, kind, dt = run_cpu_sample(iq, n_proc)
        if i >= args.warmup:
            vals.append((msps, dt))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(d for _, d in vals) / len(vals)
    sample = (f"{n_proc} processes x {len(iq) // 2 / RF_FS:g} s of the same synthetic mode-0 capture, "
              f"{'reference project binary (oracle/_ref)' if kind == 'reference' else 'oracle C port'}, stdin->/dev/null")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.captures, args.seconds), "sample": sample},
        "real_time_factor": value * 1e6 / RF_FS,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_proc, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------

def main_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    pkg = importlib.import_module("software-defined-radio-course-project_b200")
    fm = pkg.binding
    fm.load()   # raises if the CUDA library is not built: there is no CPU fallback

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or fm.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    info = fm.mode_table(MODE, TAPS)
    C = args.captures
    nb = max(1, int(args.seconds * info.rf_fs * 2 / info.block_size))
    n_bytes = nb * info.block_size
    n_pairs = n_bytes // 2
    n_pcm = nb * 2 * info.audio_per_block
    samples_per_step = C * n_pairs                       # per GPU

    # ---- synthetic input, resident in HBM: station k = rank*C + c ----
    iq = pkg.synth.synth_iq_torch(n_pairs, C, dev, info.rf_fs, first_station=rank * C, seed=1234 + rank)
    # PCM is double-buffered so that the gather of step k (NCCL, its own stream) runs under step k+1
    pcm_bufs = [torch.zeros((C, n_pcm), dtype=torch.int16, device=dev) for _ in range(2 if world > 1 else 1)]
    pcm = pcm_bufs[0]
    gathered = None
    if world > 1 and rank == 0:
        gathered = [torch.empty((C, n_pcm // 2), dtype=torch.int32, device=dev) for _ in range(world)]

    pipe = fm.Pipeline(MODE, TAPS, C, device=local)
    stream = torch.cuda.current_stream()
    pending = []                                          # outstanding gather of the previous step
    step_no = [0]

    def step_device():
        buf = pcm_bufs[step_no[0] % len(pcm_bufs)]
        step_no[0] += 1
        pipe.reset()
        pipe.process_device(iq.data_ptr(), iq.stride(0), nb, buf.data_ptr(), buf.stride(0), stream.cuda_stream)
        if world > 1:
            while pending:                                # the gather before last must be done: its
                pending.pop().wait()                      # receive buffers are about to be reused
            # the only collective: PCM to rank 0 (one R,L frame per int32 word; NCCL has no int16)
            pending.append(dist.gather(buf.view(torch.int32), gathered, dst=0, async_op=True))

    def drain():
        while pending:
            pending.pop().wait()

    # ---- parity spot check against the oracle (first capture, first blocks) ----
    parity = "skipped"
    if rank == 0:
        try:
            sys.path.insert(0, str(ROOT / "oracle"))
            import pyoracle
            chk_blocks = min(nb, 64)
            with fm.Pipeline(MODE, TAPS, 1, device=local) as p1:
                host_iq = iq[0, :chk_blocks * info.block_size].cpu().numpy()
                got = p1.process(host_iq)[0]
            ref, _ = pyoracle.Port().chain(MODE, TAPS).run(host_iq)
            parity = "bit-identical" if np.array_equal(got, ref) else f"MISMATCH ({int((got != ref).sum())} samples)"
        except Exception as e:   # the checker is optional at bench time
            parity = f"unavailable ({type(e).__name__})"

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        step_device()
    drain()
    torch.cuda.synchronize()
    barrier()
    pipe.set_timing(True)
    launches0 = pipe.kernel_launches
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    kern = {"rf_demod_ms": 0.0, "bandpass_ms": 0.0, "pll_ms": 0.0, "audio_ms": 0.0}
    torch.cuda.synchronize()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        t_host0 = time.perf_counter()
        step_device()
        # per-kernel device times of this step (event reads only; the pipeline call above
        # already made `stream` wait for the step, so this does not add work to the region)
        torch.cuda.synchronize()
        t = pipe.last_timing()
        for k_ in kern:
            kern[k_] += t[k_]
        if world > 1:                                    # per-rank step time, for diagnosing a slow rank
            print(f"[bench rank {rank}] device step {time.perf_counter() - t_host0:.3f} s (k_pll {t['pll_ms']:.0f} ms)",
                  file=sys.stderr, flush=True)
    t_host0 = time.perf_counter()
    drain()                                              # the last gather is inside the timed region
    if world > 1:
        print(f"[bench rank {rank}] last gather {time.perf_counter() - t_host0:.3f} s", file=sys.stderr, flush=True)
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = pipe.kernel_launches - launches0
    clk = clocks.stop() if rank == 0 else None
    pipe.set_timing(False)
    t_max = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_step = float(t_max.item()) / args.steps
    value = world * samples_per_step / (ms_step * 1e-3) / 1e6

    # ---- end to end through fmrx_process(): pinned host in, pinned host out ----
    e2e = None
    if not args.no_e2e:
        import psutil
        need = C * n_bytes + C * n_pcm * 2
        e2e_nb = nb
        avail = psutil.virtual_memory().available / max(1, world)
        if need * 1.3 > avail:                       # not enough host RAM for the full batch
            e2e_nb = max(1, int(nb * avail / (need * 1.3)))
        h_iq = torch.empty((C, e2e_nb * info.block_size), dtype=torch.uint8, pin_memory=True)
        h_pcm = torch.empty((C, e2e_nb * 2 * info.audio_per_block), dtype=torch.int16, pin_memory=True)
        h_iq.copy_(iq[:, :e2e_nb * info.block_size])
        torch.cuda.synchronize()

        def step_host():
            pipe.reset()
            pipe.process_raw(h_iq.data_ptr(), h_iq.stride(0), e2e_nb, h_pcm.data_ptr(), h_pcm.stride(0))

        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e_samples = C * e2e_nb * info.block_size // 2
        e2e = {"value": world * e2e_samples * args.steps / float(t_e.item()) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": world * C * e2e_nb * info.block_size,
               "d2h_bytes_per_step": world * C * e2e_nb * 2 * info.audio_per_block * 2,
               "seconds_per_capture": e2e_nb * info.block_size / 2 / info.rf_fs,
               "api": "fmrx_process (C ABI, pinned host buffers)"}
        if rank == 0 and parity == "bit-identical":
            # the host path must produce the same PCM as the device path
            same = bool(torch.equal(h_pcm[0], pcm[0, :h_pcm.shape[1]].cpu()))
            e2e["matches_device_path"] = same
        del h_iq, h_pcm

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_proc = max(1, min(cores, C))
        cpu_blocks = max(8, int(args.cpu_seconds * info.rf_fs * 2 / info.block_size))
        cpu_blocks = min(cpu_blocks, nb)
        sample_iq = iq[0, :cpu_blocks * info.block_size].cpu().numpy().tobytes()
        msps, kind, dt = run_cpu_sample(sample_iq, n_proc)
        cpu = {"value": msps, "unit": UNIT, "cores": n_proc, "kind": kind,
               "sample": f"{n_proc} processes x {cpu_blocks * info.block_size / 2 / info.rf_fs:g} s of capture 0 "
                         f"({'reference project binary' if kind == 'reference' else 'oracle C port'}), {dt:.1f} s wall"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_pll) ----
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak = float(json.loads(peaks_path.read_text())["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    n_if_step = C * nb * info.if_per_block                          # IF samples per step per GPU
    pll_launches = launches // 4 // max(1, args.steps)              # per step
    pll_ms_step = kern["pll_ms"] / args.steps
    alg_bytes_launch = 8.0 * n_if_step / max(1, pll_launches)       # 4 B pilot in + 4 B trigArg out per IF sample
    achieved = 8.0 * n_if_step / (pll_ms_step * 1e-3) / 1e9
    traffic = None
    prof = ROOT / "profiles" / "pll_traffic.json"
    if prof.exists():
        traffic = json.loads(prof.read_text())["dram_bytes_per_if_sample"] * n_if_step / max(1, pll_launches)
    roofline = {"kernel": "k_pll", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_launch, "launch_ms": pll_ms_step / max(1, pll_launches),
                "share_of_step": pll_ms_step / ms_step,
                "note": "bound by the instruction issue of one warp per capture (one dependent recurrence per capture); see pll_ns_per_sample"}
    total_k = sum(kern.values()) / args.steps
    # FP32 issue-rate view of the FIR kernels: one MAC = FMUL + FADD (bit-exact, unfused)
    macs_rf = 2.0 * TAPS / info.rf_decim            # per IQ sample (I and Q)
    macs_bp = 2.0 * TAPS / info.rf_decim
    macs_au = 2.0 * TAPS * info.audio_interp / info.audio_decim / info.rf_decim
    fp32_peak_tflops = 148 * 128 * 2 * 1.965e9 / 1e12
    kernels = {
        "k_rf_demod": {"ms_per_step": kern["rf_demod_ms"] / args.steps,
                       "fp32_frac_of_ffma_peak": 2 * macs_rf * samples_per_step / (kern["rf_demod_ms"] / args.steps * 1e-3) / 1e12 / fp32_peak_tflops},
        "k_bandpass_pair": {"ms_per_step": kern["bandpass_ms"] / args.steps,
                            "fp32_frac_of_ffma_peak": 2 * macs_bp * samples_per_step / (kern["bandpass_ms"] / args.steps * 1e-3) / 1e12 / fp32_peak_tflops},
        "k_pll": {"ms_per_step": pll_ms_step, "ns_per_if_sample_per_chain": pll_ms_step * 1e6 / (nb * info.if_per_block)},
        "k_audio": {"ms_per_step": kern["audio_ms"] / args.steps,
                    "fp32_frac_of_ffma_peak": 2 * macs_au * samples_per_step / (kern["audio_ms"] / args.steps * 1e-3) / 1e12 / fp32_peak_tflops},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(C, nb * info.block_size / 2 / info.rf_fs), "mode": MODE, "taps": TAPS,
                   "captures_per_gpu": C, "blocks_per_capture": nb, "iq_bytes_per_gpu": C * n_bytes,
                   "l2": "inputs larger than L2 (no flush needed)", "parallelism": f"{world} x independent capture shards",
                   "collective": "NCCL gather of PCM to rank 0, step k's gather under step k+1 (all inside the timed region)" if world > 1 else "none"},
        "real_time_factor": value * 1e6 / info.rf_fs,
        "real_time_factor_per_capture": (samples_per_step / C) / (ms_step * 1e-3) / info.rf_fs,
        "pll_ns_per_sample": kernels["k_pll"]["ns_per_if_sample_per_chain"],
        "roofline": roofline, "kernels": kernels, "kernel_ms_per_step_sum": total_k,
        "cpu_baseline": cpu, "e2e": e2e, "clocks": clk, "gpu_launches": launches,
        "parity_check": parity,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
