#!/usr/bin/env python
"""Generate tests/golden/long_runs.json: SHA-256 of the reference's PCM for captures of a
minute to an hour, made with the integer synthesiser (synth.synth_iq_exact: the GPU box
regenerates the SAME BYTES on the device in seconds, so the full-length parity tests and
bench.py need neither /root/reference nor minutes of CPU oracle time there).

Every case is run through BOTH checkers where the reference library is available:
  port : oracle/fmrx_oracle.c (the C restatement)
  ref  : oracle/_ref/libref_fm.so = the reference's own src/filter.cpp, compiled unmodified,
         driven by the block loop of src/project.cpp:146-193 (oracle/ref_shim.cpp)
and the entry records whether the two PCM streams are identical (they must be), so each
fixture is pinned to the reference itself.

    python tests/golden/make_long_runs.py [--only NAME_PREFIX] [--jobs N] [--no-hour]

Takes ~5 minutes on 8 cores without the one-hour case, ~75 minutes with it (one core per
checker).  Merges into the existing JSON, so cases can be (re)generated separately.
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import json
import sys
import time
from concurrent.futures import ProcessPoolExecutor, as_completed
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
OUT = Path(__file__).resolve().parent / "long_runs.json"
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def cases(include_hour=True):
    c = []

    def add(name, mode, taps, seconds, station=0, kind="stereo", seg_seconds=1.0, ref=True):
        c.append(dict(name=name, mode=mode, taps=taps, seconds=seconds, station=station, kind=kind,
                      seg_seconds=seg_seconds, ref=ref))
    # BASELINE.json configs[0..2] at lengths that cross the float counter's saturation (2^24 IF samples:
    # 69.9 s at 240 kHz, 58.3 s at 288 kHz, 65.5 s at 256 kHz) and hundreds of block boundaries
    add("long_m0_t51_80s", 0, 51, 80.0)
    add("long_m0_t101_60s", 0, 101, 60.0)
    add("long_m1_t51_65s", 1, 51, 65.0)
    add("long_m2_t51_75s", 2, 51, 75.0)
    add("long_m3_t51_70s", 3, 51, 70.0)
    add("long_m0_t301_25s", 0, 301, 25.0)
    # loops that never lock / lock on the wrong tone (bench.py worst_case and mixed legs)
    for kind in ("noise", "offtune", "nopilot"):
        add(f"hostile_m0_t51_60s_{kind}", 0, 51, 60.0, kind=kind)
        add(f"hostile_m2_t51_20s_{kind}", 2, 51, 20.0, kind=kind)
    # BASELINE.json configs[3]: the benchmark's 64 stations (port only: 64 x 9 s of CPU)
    for k in range(64):
        add(f"bench_m0_t51_60s_station{k}", 0, 51, 60.0, station=k, seg_seconds=10.0, ref=False)
    # BASELINE.json configs[4]
    if include_hour:
        add("hour_m0_t301_3600s", 0, 301, 3600.0, seg_seconds=10.0)
    return c


def run_job(case, impl):
    import pyoracle
    pkg = importlib.import_module("software-defined-radio-course-project_b200")
    port = pyoracle.Port()
    info = port.mode(case["mode"], case["taps"])
    nb = int(case["seconds"] * info.rf_fs * 2 / info.block_size)
    seg_blocks = max(1, int(round(case["seg_seconds"] * info.rf_fs * 2 / info.block_size)))
    gen = pkg.synth.ExactSynth(float(info.rf_fs), case["station"], case["kind"])
    h_iq, h_pcm = hashlib.sha256(), hashlib.sha256()
    segs = []
    t0 = time.time()
    if impl == "port":
        chain = port.chain(case["mode"], case["taps"])
    else:
        import ctypes as C
        ref = pyoracle.Reference()
        h = ref.lib.ref_chain_create(case["mode"], case["taps"])
        assert ref.lib.ref_chain_block_size(h) == info.block_size
    for b0 in range(0, nb, seg_blocks):
        n = min(seg_blocks, nb - b0)
        iq = gen.read(n * info.block_size // 2)
        h_iq.update(iq.tobytes())
        if impl == "port":
            pcm, _ = chain.run(iq)
        else:
            pcm = np.zeros(n * 2 * info.audio_per_block, np.int16)
            for b in range(n):
                blk = iq[b * info.block_size:(b + 1) * info.block_size]
                ref.lib.ref_chain_block(h, blk.ctypes.data_as(C.POINTER(C.c_uint8)),
                                        pcm[b * 2 * info.audio_per_block:].ctypes.data_as(C.POINTER(C.c_int16)), None)
        raw = pcm.tobytes()
        h_pcm.update(raw)
        segs.append(hashlib.sha256(raw).hexdigest()[:16])
    out = dict(impl=impl, n_blocks=nb, seg_blocks=seg_blocks, iq_sha256=h_iq.hexdigest(), pcm_sha256=h_pcm.hexdigest(),
               pcm_seg_sha=segs, cpu_seconds=round(time.time() - t0, 1))
    if impl == "port":
        st = chain.get_state()
        t1 = case["taps"] - 1
        off = 4 * t1 + 2          # oracle state layout: rf I/Q, pilot, channel states (taps-1 each), prev I/Q, then the PLL six
        out["pll_state_hex"] = [f"{int(v):08x}" for v in st[off:off + 6].view(np.uint32)]
    else:
        st = np.zeros(6, np.float32)
        ref.lib.ref_chain_get_pll(h, st.ctypes.data_as(C.POINTER(C.c_float)))
        out["pll_state_hex"] = [f"{int(v):08x}" for v in st.view(np.uint32)]
        ref.lib.ref_chain_destroy(h)
    return case["name"], out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--jobs", type=int, default=7)
    ap.add_argument("--no-hour", action="store_true")
    a = ap.parse_args()
    import pyoracle
    have_ref = pyoracle.Reference.available()
    todo = [c for c in cases(not a.no_hour) if c["name"].startswith(a.only)]
    # longest first
    todo.sort(key=lambda c: -c["seconds"] * c["taps"])
    results = {}
    with ProcessPoolExecutor(a.jobs) as ex:
        futs = []
        for c in todo:
            futs.append(ex.submit(run_job, c, "port"))
            if c["ref"] and have_ref:
                futs.append(ex.submit(run_job, c, "ref"))
        for f in as_completed(futs):
            name, r = f.result()
            results.setdefault(name, {})[r["impl"]] = r
            print(f"{name} [{r['impl']}] {r['cpu_seconds']} s pcm {r['pcm_sha256'][:16]}", flush=True)
    db = json.loads(OUT.read_text()) if OUT.exists() else {}
    for c in todo:
        r = results[c["name"]]
        p = r["port"]
        entry = dict(mode=c["mode"], taps=c["taps"], seconds=c["seconds"], station=c["station"], kind=c["kind"],
                     n_blocks=p["n_blocks"], seg_blocks=p["seg_blocks"], iq_sha256=p["iq_sha256"],
                     pcm_sha256=p["pcm_sha256"], pcm_seg_sha=p["pcm_seg_sha"], pll_state_hex=p["pll_state_hex"],
                     oracle_cpu_seconds=p["cpu_seconds"])
        if "ref" in r:
            q = r["ref"]
            assert q["iq_sha256"] == p["iq_sha256"]
            entry["reference_lib_identical"] = bool(q["pcm_sha256"] == p["pcm_sha256"] and q["pcm_seg_sha"] == p["pcm_seg_sha"])
            # (ncoOut_state, index 4, is dead in the reference and not kept by either side the same way)
            entry["reference_lib_pll_state_identical"] = bool(
                [v for i, v in enumerate(q["pll_state_hex"]) if i != 4] == [v for i, v in enumerate(p["pll_state_hex"]) if i != 4])
            entry["reference_cpu_seconds"] = q["cpu_seconds"]
            if not entry["reference_lib_identical"]:
                first = next((i for i, (x, y) in enumerate(zip(p["pcm_seg_sha"], q["pcm_seg_sha"])) if x != y), None)
                print(f"!! {c['name']}: oracle port and reference library DIFFER (first segment {first})", flush=True)
        db[c["name"]] = entry
    OUT.write_text(json.dumps(db, indent=0, sort_keys=True) + "\n")
    bad = [k for k, v in db.items() if v.get("reference_lib_identical") is False]
    print(f"wrote {OUT} ({len(db)} cases); oracle != reference in: {bad or 'none'}")


if __name__ == "__main__":
    main()
