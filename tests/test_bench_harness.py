"""bench.py's own machinery on the CPU: the whole-run parity check (golden hashes + live oracle) must call a
correct PCM correct and a wrong one wrong, and the reference arm must print the contract's JSON line."""
import importlib.util
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_parity_full_counts_differences(port, synth):
    b = _bench()
    info = port.mode(0, 51)
    nb = 40
    iq = np.stack([synth.synth_iq_exact(nb * info.block_size // 2, 2.4e6, station=k) for k in (0, 1, 2)])
    pcm = np.stack([port.chain(0, 51).run(iq[c])[0] for c in range(3)])
    seconds = nb * info.block_size / 2 / info.rf_fs
    good = b.parity_full(np, iq, pcm, [0, 1, 2], ["stereo"] * 3, seconds, 60.0, 2)
    assert good["captures_compared"] == 3 and good["samples_differ"] == 0 and good["max_abs_lsb"] == 0
    assert good["golden_captures"] == 0                       # (fixtures exist for 60 s captures only)
    bad_pcm = pcm.copy()
    bad_pcm[1, 1000] += 3
    bad_pcm[2, 5] -= 1
    bad = b.parity_full(np, iq, bad_pcm, [0, 1, 2], ["stereo"] * 3, seconds, 60.0, 2)
    assert bad["samples_differ"] == 2 and bad["max_abs_lsb"] == 3 and bad["captures_gt_1lsb"] == 1
    assert bad["first_difference"] == {"capture": 1, "pcm_index": 1000}
    assert good["input_sha256"] == bad["input_sha256"]


def test_golden_fixtures_cover_the_bench_batch():
    b = _bench()
    g = b.load_golden()
    for k in range(64):
        e = g[f"bench_m0_t51_60s_station{k}"]
        assert e["station"] == k and e["n_blocks"] == 22500 and len(e["pcm_sha256"]) == 64


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-seconds", "1"], capture_output=True, text=True, timeout=300)
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "Msamples/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["higher_is_better"] is True and line["gpu_launches"] == 0
