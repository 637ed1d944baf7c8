#!/usr/bin/env python
"""Latency and sustained real-time factor of the streaming CLI (host/project_main.cpp: reader / process / writer
threads over a pinned ring, ~0.2 s chunks by default) for modes 0-3, through a pipe as `rtl_sdr | project | aplay`
would run it.  One JSON line per mode.

    python tools/stream_stats.py [--seconds 8] > gpurun_out/stream_stats.jsonl
"""
import argparse
import importlib
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("software-defined-radio-course-project_b200")
CLI = ROOT / "software-defined-radio-course-project_b200" / "bin" / "project"

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=8.0)
ap.add_argument("--taps", type=int, default=51)
a = ap.parse_args()
for mode in range(4):
    info = pkg.binding.mode_table(mode, a.taps)
    nb = max(3, int(a.seconds * info.rf_fs * 2 / info.block_size))
    iq = pkg.synth.synth_iq_exact(nb * info.block_size // 2, float(info.rf_fs), station=mode).tobytes()
    for chunk in ([], ["--chunk-blocks", "1"] if mode < 2 else []):
        t0 = time.perf_counter()
        r = subprocess.run([str(CLI), str(mode), "s", "--stats", "--taps", str(a.taps), *chunk], input=iq, capture_output=True, timeout=600)
        wall = time.perf_counter() - t0
        line = [ln for ln in r.stderr.decode().splitlines() if ln.startswith("fmrx stats: ")]
        st = json.loads(line[0][len("fmrx stats: "):]) if line else {"error": r.stderr.decode()[-300:]}
        st["wall_including_process_start_s"] = wall
        st["pcm_bytes"] = len(r.stdout)
        print(json.dumps(st), flush=True)
