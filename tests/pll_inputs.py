"""PLL inputs that defeat k_pll's speculation one way or another: the predictor misses, a
candidate's guard fails (pilot sample 0 or subnormal, angle at the +-pi seam), trigArg changes
binade all the time.  Used on the CPU (oracle against the compiled reference) and on the GPU
(CUDA path against the oracle): everything must come out bit-identical."""
import numpy as np

KINDS = ["noise", "zeros_and_tiny", "constant_sign", "dropouts", "am_pilot", "huge"]


def hostile_pilot(kind: str, n: int = 40000) -> np.ndarray:
    t = np.arange(n, dtype=np.float64)
    rng = np.random.default_rng(11)
    tone = np.sin(2 * np.pi * 19000.3 / 240e3 * t + 0.7)
    if kind == "noise":                       # no pilot at all: the loop never locks
        x = rng.uniform(-1, 1, n)
    elif kind == "zeros_and_tiny":            # exact zeros (atan2(+-0, +-0)), subnormal products
        x = 0.1 * tone
        x[rng.integers(0, n, 400)] = 0.0
        x[rng.integers(0, n, 400)] = -0.0
        x[rng.integers(0, n, 400)] = 1e-41
        x[5000:5200] = 0.0
    elif kind == "constant_sign":             # pilot riding on a DC offset: never changes sign
        x = 0.5 + 0.1 * tone
    elif kind == "dropouts":                  # lock, lose it, re-acquire
        x = 0.1 * tone
        x[10000:14000] = 0.02 * rng.standard_normal(4000)
        x[25000:25100] *= -1.0
    elif kind == "am_pilot":                  # deep amplitude modulation with sign flips through zero
        x = 0.1 * tone * np.sin(2 * np.pi * 37.0 / 240e3 * t)
    elif kind == "huge":                      # float-range amplitudes
        x = 1e30 * tone
        x[::97] = 3e38
    else:
        raise ValueError(kind)
    return x.astype(np.float32)
