#!/usr/bin/env python
"""Device-resident throughput of the chain on every BASELINE.json configuration that is
not the bench workload (bench.py times configs[3]; these are the parity-test cases):
one JSON line per configuration, CUDA-event timed, synthetic IQ generated on the GPU.

    python tools/config_sweep.py [--hour-seconds 3600] > gpurun_out/config_sweep.jsonl

Every line also carries a parity spot check of the first blocks against the oracle
(checker only) and the per-kernel device times of the pass.
"""
from __future__ import annotations

import argparse
import importlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hour-seconds", type=float, default=3600.0, help="length of the configs[4] capture")
    ap.add_argument("--only", default="", help="comma-separated config names")
    args = ap.parse_args()

    import numpy as np
    import torch
    pkg = importlib.import_module("software-defined-radio-course-project_b200")
    fm = pkg.binding
    fm.load()
    if not torch.cuda.is_available():
        raise SystemExit("config_sweep.py: no CUDA device -- the product path has no CPU fallback")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    try:
        import pyoracle
        port = pyoracle.Port()
    except Exception:
        port = None

    # name, BASELINE.json config, mode, taps, captures, seconds
    cases = [
        ("m0_t101_1x60s", "configs[0] mode 0 mono, 101 taps (the reference ignores `channels`: same R,L stream)", 0, 101, 1, 60.0),
        ("m0_t51_1x60s", "configs[1] mode 0 stereo, 51 taps, one capture", 0, 51, 1, 60.0),
        ("m1_t51_1x120s", "mode 1 (1.44 Msps, decim 4), 51 taps, one capture", 1, 51, 1, 120.0),
        ("m2_t51_1x120s", "configs[2] mode 2 (147/800 polyphase to 44.1 kHz), 51 taps, one capture", 2, 51, 1, 120.0),
        ("m3_t51_1x120s", "configs[2] mode 3 (2.304 Msps, 441/2560 polyphase), 51 taps, one capture", 3, 51, 1, 120.0),
        ("m0_t301_64x8s", "mode 0, 301 taps, 64 captures x 8 s (FIR-heavy case)", 0, 301, 64, 8.0),
        ("m0_t301_1hour", f"configs[4] one {args.hour_seconds:g} s capture, 301 taps, on one GPU "
                          "(the PLL is one chain: time shards on more GPUs add capacity, not speed)", 0, 301, 1, args.hour_seconds),
    ]
    only = {s for s in args.only.split(",") if s}
    for name, what, mode, taps, C, seconds in cases:
        if only and name not in only:
            continue
        info = fm.mode_table(mode, taps)
        nb = max(1, int(seconds * info.rf_fs * 2 / info.block_size))
        n_pairs = nb * info.block_size // 2
        iq = pkg.synth.synth_iq_torch(n_pairs, C, dev, info.rf_fs, first_station=0, seed=77)
        pcm = torch.zeros((C, nb * 2 * info.audio_per_block), dtype=torch.int16, device=dev)
        stream = torch.cuda.current_stream()
        parity = "skipped"
        with fm.Pipeline(mode, taps, C, device=0) as pipe:
            warm = max(1, min(nb, int(1.0 * info.rf_fs * 2 / info.block_size)))   # about a second of signal
            for _ in range(2):                             # warm-up on a prefix
                pipe.reset()
                pipe.process_device(iq.data_ptr(), iq.stride(0), warm, pcm.data_ptr(), pcm.stride(0), stream.cuda_stream)
            torch.cuda.synchronize()
            pipe.reset()
            pipe.set_timing(True)
            l0 = pipe.kernel_launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            pipe.process_device(iq.data_ptr(), iq.stride(0), nb, pcm.data_ptr(), pcm.stride(0), stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            kern = pipe.last_timing()
            launches = pipe.kernel_launches - l0
            st = pipe.pll_state(0)
        if port is not None:
            chk = min(nb, 48 if mode < 2 else 2)
            host_iq = iq[0, :chk * info.block_size].cpu().numpy()
            ref, _ = port.chain(mode, taps).run(host_iq)
            got = pcm[0, :chk * 2 * info.audio_per_block].cpu().numpy()
            parity = "bit-identical" if np.array_equal(got, ref) else f"MISMATCH ({int((got != ref).sum())} samples)"
        n_if = nb * info.if_per_block
        msps = C * n_pairs / (ms * 1e-3) / 1e6
        print(json.dumps({
            "config": name, "what": what, "mode": mode, "taps": taps, "captures": C, "seconds_per_capture": nb * info.block_size / 2 / info.rf_fs,
            "iq_msps": msps, "real_time_factor": msps * 1e6 / info.rf_fs, "real_time_factor_per_capture": msps * 1e6 / info.rf_fs / C,
            "ms": ms, "kernels_ms": kern, "pll_ns_per_if_sample_per_chain": kern["pll_ms"] * 1e6 / n_if,
            "gpu_launches": launches, "trigOffset_end": float(st[5]), "parity_first_blocks": parity,
        }), flush=True)
        del iq, pcm
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
