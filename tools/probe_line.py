"""Condense tools/gpu_probe.py output (stdin) to one line per run; pass kernel debug lines through."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for l in sys.stdin:
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l)
        print(tag, "mode", d["mode"], "taps", d["taps"], "Msps", round(d["iq_msps"]), "pll_ns", round(d["pll_ns_per_sample"], 1),
              "first_call_ns", round(d.get("first_call_pll_ns_per_sample", 0), 1), "diag", d["pll_groups_last_chunk"],
              d["pll_groups_redone_last_chunk"], "rf", round(d["rf_demod_ms"], 2), "bp", round(d["bandpass_ms"], 2), "au",
              round(d["audio_ms"], 2))
    elif l.startswith("pll dbg"):
        print(l[:330])
