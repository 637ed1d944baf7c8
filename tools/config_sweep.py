#!/usr/bin/env python
"""Device-resident throughput AND whole-run parity of the chain on every BASELINE.json
configuration that is not the bench workload (bench.py times configs[3]): one JSON line per
configuration, CUDA-event timed, the capture made on the GPU by the integer synthesiser and the
complete PCM compared by SHA-256 with the reference's (tests/golden/long_runs.json: oracle port
and the reference's own compiled filter.cpp, precomputed for exactly these bytes).

    python tools/config_sweep.py [--only name,name] > gpurun_out/config_sweep.jsonl
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
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
    golden = json.loads((ROOT / "tests" / "golden" / "long_runs.json").read_text())

    # name, BASELINE.json config, golden fixture(s) (one per capture)
    cases = [
        ("m0_t101_1x60s", "configs[0] mode 0 mono, 101 taps (the reference ignores `channels`: same R,L stream)", ["long_m0_t101_60s"]),
        ("m0_t51_1x80s", "configs[1] mode 0 stereo, 51 taps, one capture, across the counter's saturation (69.9 s)", ["long_m0_t51_80s"]),
        ("m1_t51_1x65s", "mode 1 (1.152 Msps, decim 4), 51 taps, across saturation (58.3 s)", ["long_m1_t51_65s"]),
        ("m2_t51_1x75s", "configs[2] mode 2 (147/800 polyphase to 44.1 kHz), 51 taps, across saturation (69.9 s)", ["long_m2_t51_75s"]),
        ("m3_t51_1x70s", "configs[2] mode 3 (2.304 Msps, 441/2560 polyphase), 51 taps, across saturation (65.5 s)", ["long_m3_t51_70s"]),
        ("m0_t301_1x25s", "mode 0, 301 taps, one capture", ["long_m0_t301_25s"]),
        ("m0_t51_nopilot_60s", "mode 0, a capture without a pilot (the loop never locks)", ["hostile_m0_t51_60s_nopilot"]),
        ("m0_t51_noise_60s", "mode 0, noise only", ["hostile_m0_t51_60s_noise"]),
        ("m2_t51_noise_20s", "mode 2, noise only", ["hostile_m2_t51_20s_noise"]),
        ("m0_t51_64x60s", "configs[3] the bench batch: 64 stations x 60 s", [f"bench_m0_t51_60s_station{k}" for k in range(64)]),
        ("m0_t301_1hour", "configs[4] one 3600 s capture, 301 taps, on one GPU", ["hour_m0_t301_3600s"]),
    ]
    only = {s for s in args.only.split(",") if s}
    for name, what, fixtures in cases:
        if only and name not in only:
            continue
        if any(f not in golden for f in fixtures):
            print(json.dumps({"config": name, "skipped": "fixture not generated"}), flush=True)
            continue
        gs = [golden[f] for f in fixtures]
        g0 = gs[0]
        mode, taps, C, nb = g0["mode"], g0["taps"], len(gs), g0["n_blocks"]
        info = fm.mode_table(mode, taps)
        n_pairs = nb * info.block_size // 2
        iq = torch.empty((C, 2 * n_pairs), dtype=torch.uint8, device=dev)
        for c, g in enumerate(gs):
            pkg.synth.synth_iq_exact_torch(n_pairs, 1, dev, float(info.rf_fs), first_station=g["station"], kinds=[g["kind"]], out=iq[c:c + 1])
        pcm = torch.zeros((C, nb * 2 * info.audio_per_block), dtype=torch.int16, device=dev)
        stream = torch.cuda.current_stream()
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
        same = 0
        for c, g in enumerate(gs):
            same += hashlib.sha256(pcm[c].cpu().numpy().tobytes()).hexdigest() == g["pcm_sha256"]
        n_if = nb * info.if_per_block
        msps = C * n_pairs / (ms * 1e-3) / 1e6
        print(json.dumps({
            "config": name, "what": what, "mode": mode, "taps": taps, "captures": C, "seconds_per_capture": nb * info.block_size / 2 / info.rf_fs,
            "iq_msps": msps, "real_time_factor": msps * 1e6 / info.rf_fs, "real_time_factor_per_capture": msps * 1e6 / info.rf_fs / C,
            "ms": ms, "kernels_ms": kern, "pll_ns_per_if_sample_per_chain": kern["pll_ms"] * 1e6 / n_if,
            "gpu_launches": launches, "trigOffset_end": float(st[5]),
            "parity_whole_run": f"{same}/{C} captures bit-identical to the reference PCM (sha256 over {nb * 2 * info.audio_per_block} int16 each)",
        }), flush=True)
        if name == "m0_t301_1hour":
            # the same capture time-sharded (fmrx_long_*: feed-forward stages of all shards at once from FIR halos, the
            # PLL state handed from shard to shard, PCM gathered on the first device) over every visible GPU -- two
            # shards on the one GPU if there is only one
            n_dev = torch.cuda.device_count()
            devs = list(range(n_dev)) if n_dev > 1 else [0, 0]
            with fm.LongCapture(mode, taps, devs, nb) as lc:
                ptrs, keep = [], []
                for r, d in enumerate(devs):
                    first, cnt, halo = lc.shard(r)
                    piece = iq[0, (first - halo) * info.block_size:(first + cnt) * info.block_size]
                    t = piece if d == 0 else piece.to(f"cuda:{d}")
                    keep.append(t)
                    ptrs.append(t.data_ptr())
                pcm.zero_()
                torch.cuda.synchronize()
                ms_l = lc.process_device(ptrs, pcm.data_ptr())
                ok = hashlib.sha256(pcm[0].cpu().numpy().tobytes()).hexdigest() == g0["pcm_sha256"]
            print(json.dumps({"config": name + "_time_sharded", "what": f"the same capture as {len(devs)} time shards on devices {devs}",
                              "iq_msps": n_pairs / (ms_l * 1e-3) / 1e6, "ms": ms_l, "single_pipeline_ms": ms,
                              "parity_whole_run": "bit-identical to the reference PCM (sha256)" if ok else "MISMATCH"}), flush=True)
            del keep
        del iq, pcm
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
