"""TEST INFRASTRUCTURE ONLY: a voiced-looking test signal shared by the feature and CLI tests."""
import numpy as np


def utterance(seed: int, n: int = 40000, lead: int = 5000, tail: int = 7000) -> np.ndarray:
    """Voiced-looking test signal (harmonics with a slow envelope + noise) between two near-silent stretches."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 22050.0
    f0 = 110.0 + 40.0 * rng.random()
    body = sum(np.sin(2 * np.pi * f0 * k * t + rng.random()) / k for k in range(1, 9))
    body *= 0.2 * (0.6 + 0.4 * np.sin(2 * np.pi * 3.0 * t)) * np.hanning(n)
    body += 0.003 * rng.standard_normal(n)
    quiet = lambda m: 2e-5 * rng.standard_normal(m)
    return np.concatenate([quiet(lead), body, quiet(tail)]).astype(np.float32)


