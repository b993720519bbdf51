"""Deterministic synthetic PCM for parity tests and bench.py.

Per clip: 2-4 slowly FM-modulated tones per channel + low-passed Gaussian noise (about -20 dBFS)
+ ~2 decaying noise bursts per second (forces short blocks) + inter-channel correlation rho drawn
from {0, 0.5, 0.9, 1.0} per 1.5 s segment + 0.5 s of digital silence, peak <= 0.7 FS.  Output is int16, shape
(nsamples, nch), identical for a given (seed, seconds, samprate, nch).
"""
import numpy as np


def synth_pcm(seed: int, seconds: float, samprate: int, nch: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = int(round(seconds * samprate))
    t = np.arange(n, dtype=np.float64) / samprate
    rho = float(rng.choice([0.0, 0.5, 0.9, 1.0]))

    def one_channel():
        x = np.zeros(n)
        for _ in range(int(rng.integers(2, 5))):
            f0 = float(rng.uniform(80.0, min(6000.0, 0.2 * samprate)))
            fm = float(rng.uniform(0.05, 0.8))
            dev = float(rng.uniform(0.0, 0.05)) * f0
            amp = float(rng.uniform(0.03, 0.18))
            ph = 2 * np.pi * (f0 * t + dev / (2 * np.pi * fm) * np.sin(2 * np.pi * fm * t)) + rng.uniform(0, 6.28)
            x += amp * np.sin(ph)
        noise = rng.standard_normal(n)
        k = np.array([0.2, 0.3, 0.3, 0.2])          # mild low-pass
        noise = np.convolve(noise, k, mode="same") * 0.1
        x += noise
        nb = max(1, int(2 * seconds))
        for _ in range(nb):
            p = int(rng.integers(0, max(1, n - 1)))
            ln = int(min(n - p, rng.integers(int(0.01 * samprate), int(0.08 * samprate) + 1)))
            env = np.exp(-np.arange(ln) / (0.012 * samprate)) * float(rng.uniform(0.3, 1.2))
            x[p:p + ln] += env * rng.standard_normal(ln)
        return x

    common = one_channel()
    # correlation changes every ~1.5 s so that both L/R and M/S frames occur inside one clip
    nseg = int(np.ceil(seconds / 1.5))
    seg_rho = rng.choice([0.0, 0.5, 0.9, 1.0], size=nseg)
    seg_rho[0] = rho
    rho_t = np.repeat(seg_rho, int(np.ceil(n / nseg)))[:n]
    chans = []
    for _ in range(nch):
        own = one_channel()
        chans.append(rho_t * common + (1.0 - rho_t) * own)
    x = np.stack(chans, axis=1)
    if n > samprate:                                 # 0.5 s of digital silence
        s0 = int(rng.integers(0, n - samprate // 2))
        x[s0:s0 + samprate // 2, :] = 0.0
    peak = np.max(np.abs(x)) + 1e-12
    x *= min(1.0, 0.7 / peak)
    return np.round(x * 32767.0).astype(np.int16)
