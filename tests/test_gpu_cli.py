"""The host programs on the GPU: the `project` CLI (host/project_main.cpp) and the
drop-in link of the UNMODIFIED reference main against the filter.h shim."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
CLI = ROOT / "software-defined-radio-course-project_b200" / "bin" / "project"
DROPIN = ROOT / "oracle" / "_ref" / "project_dropin"


def _run(exe, args, data, timeout=300):
    return subprocess.run([str(exe), *args], input=data, capture_output=True, timeout=timeout)


@pytest.mark.parametrize("mode,chan,taps,nblocks,extra", [(0, "2", 51, 40, 0), (0, "s", 51, 33, 4321), (1, "m", 51, 21, 0),
                                                          (0, "1", 101, 12, 100), (2, "s", 51, 2, 0)])
def test_cli_stdin_to_stdout_matches_oracle(fm, port, synth, mode, chan, taps, nblocks, extra):
    assert CLI.exists(), "bin/project not built (run __graft_entry__.build())"
    info = port.mode(mode, taps)
    iq = synth.synth_iq(nblocks * info.block_size // 2 + extra, info.rf_fs, seed=50 + mode)
    args = [str(mode), chan] + (["--taps", str(taps)] if taps != 51 else []) + ["--chunk-blocks", "7"]
    r = _run(CLI, args, iq.tobytes())
    assert r.returncode == 1                                   # the reference's exit status at EOF
    err = r.stderr.decode()
    label = "mono" if chan in ("1", "m") else "stereo"
    assert f"Operating in mode {mode}, {label}" in err and "End of input stream reached!" in err
    ref, _ = port.chain(mode, taps).run(iq)                    # channel argument changes nothing (reference quirk i)
    out = np.frombuffer(r.stdout, np.int16)
    assert np.array_equal(out, ref)


def test_cli_argument_handling_like_reference(fm):
    assert b"Invalid mode: 7!" in _run(CLI, ["7", "1"], b"").stderr
    assert b"Invaild channel: 3!" in _run(CLI, ["0", "3"], b"").stderr
    r = _run(CLI, ["2"], b"")                                  # a single argument is ignored -> mode 0
    assert b"Operating in default mode 0, mono" in r.stderr and r.returncode == 1
    assert _run(CLI, ["0", "1", "2"], b"").returncode == 1     # usage


def test_unmodified_reference_main_links_against_the_shim(fm, port, synth):
    """src/project.cpp (threads, queue and all) + filter_shim + iofunc_shim + libfmrx_b200:
    its stdout is an exact prefix of the oracle PCM, short by the <=4 blocks the reference
    itself loses at EOF."""
    if not DROPIN.exists():
        pytest.skip("oracle/_ref/project_dropin not built (needs the reference checkout at build time)")
    info = port.mode(0, 51)
    iq = synth.synth_iq(40 * info.block_size // 2, info.rf_fs, seed=77)
    import json, time
    t0 = time.perf_counter()
    r = _run(DROPIN, ["0", "2"], iq.tobytes(), timeout=600)
    w_short = time.perf_counter() - t0
    assert r.returncode == 1 and b"End of input stream reached!" in r.stderr
    # how fast the level-1 drop-in is (eleven operator calls per block, each with its own copies): the marginal
    # time per block between a 40- and a 2000-block run (process start and context creation cancel).  Recorded, not a gate.
    long_iq = np.tile(iq, 50)
    t0 = time.perf_counter()
    r2 = _run(DROPIN, ["0", "2"], long_iq.tobytes(), timeout=600)
    w_long = time.perf_counter() - t0
    assert r2.returncode == 1
    block_s = info.block_size / 2 / info.rf_fs
    rec = {"what": "unmodified reference main + filter.h shim + libfmrx_b200 (operator entry points), mode 0, 51 taps",
           "wall_40_blocks_s": w_short, "wall_2000_blocks_s": w_long, "ms_per_block": 1e3 * (w_long - w_short) / 1960,
           "real_time_factor": 1960 * block_s / max(1e-9, w_long - w_short)}
    print(rec)
    out_dir = ROOT / "gpurun_out"
    if out_dir.is_dir():
        (out_dir / "dropin_rtf.json").write_text(json.dumps(rec) + "\n")
    out = np.frombuffer(r.stdout, np.int16)
    ref, _ = port.chain(0, 51).run(iq)
    per_block = 2 * info.audio_per_block
    assert len(out) % per_block == 0 and 0 <= len(ref) - len(out) <= 4 * per_block and len(out) > 20 * per_block
    assert np.array_equal(out, ref[:len(out)])


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_cli_streaming_stats_and_latency(fm, port, synth, mode):
    """The live path (src/project.cpp:392-393: rtl_sdr | project | aplay): default ~0.2 s chunks through the
    reader / process / writer ring.  PCM identical to the oracle; --stats reports a sustained real-time factor
    above 1 (it must keep up with a live dongle) and a per-chunk latency."""
    import json
    info = port.mode(mode, 51)
    nb = max(3, int(3.0 * info.rf_fs * 2 / info.block_size))
    iq = synth.synth_iq_exact(nb * info.block_size // 2, float(info.rf_fs), station=mode)
    r = _run(CLI, [str(mode), "s", "--stats"], iq.tobytes())
    assert r.returncode == 1
    ref, _ = port.chain(mode, 51).run(iq)
    assert np.array_equal(np.frombuffer(r.stdout, np.int16), ref)
    line = [ln for ln in r.stderr.decode().splitlines() if ln.startswith("fmrx stats: ")]
    assert line, r.stderr.decode()
    st = json.loads(line[0][len("fmrx stats: "):])
    print(st)
    assert st["real_time_factor"] > 1.0 and st["latency_ms"]["max"] > 0
