"""Synthetic FM-stereo IQ captures (the reference's .raw captures are absent).

Signal model, at fs = rf_fs:

    L = 0.5 sin(2 pi f_L t),  R = 0.5 sin(2 pi f_R t)
    mpx = 0.45 (L+R) + 0.1 sin(2 pi 19000 t) + 0.45 (L-R) sin(2 pi 38000 t)
    phi = 2 pi 75000 cumsum(mpx) / fs          (float64, carried across chunks)
    I = cos phi + sigma n_I,  Q = sin phi + sigma n_Q,   sigma = 0.01
    u8 = clip(rint(x*100 + 128), 0, 255),  interleaved I,Q

Station k of a batch uses seed=k, f_L = 1000+37(k mod 64) Hz, f_R = 3000+53(k mod 64) Hz.

`synth_iq` is the numpy (host) generator used by the parity tests and the
golden fixtures; `synth_iq_torch` produces the same signal model on a CUDA
device for the benchmark (closed-form phase integral, torch RNG -- same
statistics, not the same bytes).
"""
from __future__ import annotations

import numpy as np

PILOT_HZ = 19000.0
DEVIATION_HZ = 75000.0


def station_tones(k: int) -> tuple[float, float]:
    """Audio tones of station k.  Stations 0..63 are SURVEY 8(d)'s batch; beyond that (multi-GPU
    batches) the tone pair repeats with period 64 -- only the noise differs -- because 1000 + 37 k and
    3000 + 53 k would leave the 15 kHz audio band of FM broadcast from k ~ 230 on and then sit on
    the 19 kHz pilot itself: not an FM-stereo signal any more (the reference's PLL locks onto the
    tone instead of the pilot)."""
    k %= 64
    return 1000.0 + 37.0 * k, 3000.0 + 53.0 * k


def synth_iq(n_pairs: int, rf_fs: float = 2.4e6, seed: int = 0, f_l: float | None = None,
             f_r: float | None = None, sigma: float = 0.01, chunk: int = 1 << 20,
             pilot_hz: float = PILOT_HZ) -> np.ndarray:
    """Return ``2*n_pairs`` uint8 (I,Q interleaved)."""
    if f_l is None or f_r is None:
        f_l, f_r = station_tones(seed)
    rng = np.random.default_rng(seed)
    out = np.empty(2 * n_pairs, np.uint8)
    phi0 = 0.0
    for s in range(0, n_pairs, chunk):
        e = min(n_pairs, s + chunk)
        t = np.arange(s, e, dtype=np.float64) / rf_fs
        left = 0.5 * np.sin(2 * np.pi * f_l * t)
        right = 0.5 * np.sin(2 * np.pi * f_r * t)
        mpx = (0.45 * (left + right) + 0.1 * np.sin(2 * np.pi * pilot_hz * t)
               + 0.45 * (left - right) * np.sin(2 * np.pi * 2 * pilot_hz * t))
        phi = phi0 + 2 * np.pi * DEVIATION_HZ * np.cumsum(mpx) / rf_fs
        phi0 = float(phi[-1])
        i = np.cos(phi) + sigma * rng.standard_normal(e - s)
        q = np.sin(phi) + sigma * rng.standard_normal(e - s)
        out[2 * s:2 * e:2] = np.clip(np.rint(i * 100 + 128), 0, 255).astype(np.uint8)
        out[2 * s + 1:2 * e:2] = np.clip(np.rint(q * 100 + 128), 0, 255).astype(np.uint8)
    return out


def synth_iq_torch(n_pairs: int, n_captures: int, device, rf_fs: float = 2.4e6,
                   first_station: int = 0, sigma: float = 0.01, seed: int = 0,
                   chunk: int = 1 << 22):
    """``[n_captures, 2*n_pairs]`` uint8 on ``device`` (torch is device plumbing
    here: it only fabricates benchmark input).  The FM phase is the closed-form
    integral of the multiplex, evaluated in float64."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_captures, 2 * n_pairs), dtype=torch.uint8, device=device)
    w = 2 * np.pi
    kdev = w * DEVIATION_HZ
    for c in range(n_captures):
        f_l, f_r = station_tones(first_station + c)
        for s in range(0, n_pairs, chunk):
            e = min(n_pairs, s + chunk)
            t = torch.arange(s, e, dtype=torch.float64, device=device) / rf_fs

            def icos(f, amp):  # integral of amp*sin(2 pi f t) dt from 0
                return amp * (1.0 - torch.cos(w * f * t)) / (w * f)

            def isin_prod(fa, fb, amp):  # integral of amp*sin(a t) sin(b t), a != b
                a, b = w * fa, w * fb
                return amp * 0.5 * (torch.sin((a - b) * t) / (a - b) - torch.sin((a + b) * t) / (a + b))

            integ = (icos(f_l, 0.225) + icos(f_r, 0.225) + icos(PILOT_HZ, 0.1)
                     + isin_prod(f_l, 2 * PILOT_HZ, 0.225) - isin_prod(f_r, 2 * PILOT_HZ, 0.225))
            phi = kdev * integ
            n = e - s
            i = torch.cos(phi) + sigma * torch.randn(n, dtype=torch.float64, device=device, generator=g)
            q = torch.sin(phi) + sigma * torch.randn(n, dtype=torch.float64, device=device, generator=g)
            out[c, 2 * s:2 * e:2] = torch.clamp(torch.round(i * 100 + 128), 0, 255).to(torch.uint8)
            out[c, 2 * s + 1:2 * e:2] = torch.clamp(torch.round(q * 100 + 128), 0, 255).to(torch.uint8)
    return out
