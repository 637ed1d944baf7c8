#!/usr/bin/env python
"""How many PLL chains one B200 carries: device-resident throughput of the chain against the number of
captures per GPU (one k_pll CTA -- one chain warp -- per capture; up to 148 - 32 CTAs get an SM of their own).

    python tools/capacity_sweep.py [--taps 51] [--seconds 10] [--captures 64,96,116,128,148,192,256] > gpurun_out/capacity.jsonl
"""
from __future__ import annotations

import argparse
import importlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taps", type=int, default=51)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--captures", default="64,96,116,128,148,192,256")
    a = ap.parse_args()
    import torch
    pkg = importlib.import_module("software-defined-radio-course-project_b200")
    fm = pkg.binding
    fm.load()
    dev = torch.device("cuda", 0)
    info = fm.mode_table(a.mode, a.taps)
    nb = max(1, int(a.seconds * info.rf_fs * 2 / info.block_size))
    n_pairs = nb * info.block_size // 2
    counts = [int(c) for c in a.captures.split(",")]
    cmax = max(counts)
    # 64 distinct stations, repeated (the tone plan repeats with period 64 anyway)
    base = pkg.synth.synth_iq_exact_torch(n_pairs, min(cmax, 64), dev, float(info.rf_fs), first_station=0)
    for C in counts:
        iq = base.repeat((C + base.shape[0] - 1) // base.shape[0], 1)[:C].contiguous()
        pcm = torch.zeros((C, nb * 2 * info.audio_per_block), dtype=torch.int16, device=dev)
        s = torch.cuda.current_stream()
        with fm.Pipeline(a.mode, a.taps, C, device=0) as p:
            best = None
            for rep in range(3):
                p.reset()
                p.set_timing(True)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                p.process_device(iq.data_ptr(), iq.stride(0), nb, pcm.data_ptr(), pcm.stride(0), s.cuda_stream)
                e1.record(s)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                k = p.last_timing()
                if rep and (best is None or ms < best[0]):
                    best = (ms, k)
        ms, k = best
        same = bool(torch.equal(pcm[0], pcm[min(C - 1, 64 if C > 64 else 0)])) if C > 64 else None
        print(json.dumps({"captures": C, "mode": a.mode, "taps": a.taps, "seconds_per_capture": nb * info.block_size / 2 / info.rf_fs,
                          "iq_msps": C * n_pairs / (ms * 1e-3) / 1e6, "ms": ms, "kernels_ms": k,
                          "pll_ns_per_if_sample_per_chain": k["pll_ms"] * 1e6 / (nb * info.if_per_block),
                          "whole_sm_per_chain": C + 32 <= 148, "capture_64_equals_capture_0": same}), flush=True)
        del iq, pcm
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
