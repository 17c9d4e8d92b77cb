"""Synthetic conditioning features with the structure of the real ones.

x[B, 256+1+n_spk, T] as assembled by Net._forward
(/root/reference/egs/vc/libritts/local/tuning/hifigan.py:83-97): channels 0..255 are the
ASR bottleneck (rows of a 48-entry VQ codebook, held for geometric run lengths), channel 256
the CMVN-normalised F0 (0 on unvoiced frames), channels 257.. the one-hot target speaker,
constant in time.  numpy's PCG64 generator makes it reproducible on any host.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

N_BN = 256
N_CODEWORDS = 48
N_SPEAKERS = 247


def codebook(seed: int = 1234) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((N_CODEWORDS, N_BN)).astype(np.float32)


def utterance_parts(rng: np.random.Generator, frames: int, n_spk: int = N_SPEAKERS, voiced_fraction: float = 0.6):
    """One utterance in the compact form: (idx uint8 [frames] VQ codeword per frame, f0 float32 [frames], spk int)."""
    idx = np.empty(frames, dtype=np.int64)
    f0 = np.zeros(frames, dtype=np.float32)
    t = 0
    while t < frames:                       # VQ index held for a geometric run (mean 4 frames)
        run = int(rng.geometric(0.25))
        idx[t:t + run] = rng.integers(N_CODEWORDS)
        t += run
    t = 0
    while t < frames:                       # alternating voiced / unvoiced segments
        seg = int(rng.integers(5, 40))
        if rng.random() < voiced_fraction:
            f0[t:t + seg] = rng.standard_normal(min(seg, frames - t)).astype(np.float32)
        t += seg
    return idx.astype(np.uint8), f0, int(rng.integers(n_spk))


def assemble(idx: np.ndarray, f0: np.ndarray, spk: int, n_spk: int = N_SPEAKERS, cb: Optional[np.ndarray] = None) -> np.ndarray:
    """x = [codebook[idx] | f0 | one_hot(spk)] as float32 [256+1+n_spk, frames] (hifigan.py:83-97)."""
    cb = codebook() if cb is None else cb
    frames = idx.shape[0]
    x = np.zeros((N_BN + 1 + n_spk, frames), dtype=np.float32)
    x[:N_BN] = cb[idx.astype(np.int64)].T
    x[N_BN] = f0
    x[N_BN + 1 + spk] = 1.0
    return x


def utterance(rng: np.random.Generator, frames: int, n_spk: int = N_SPEAKERS,
              cb: Optional[np.ndarray] = None, voiced_fraction: float = 0.6) -> np.ndarray:
    """One utterance: float32 [256+1+n_spk, frames]."""
    idx, f0, spk = utterance_parts(rng, frames, n_spk, voiced_fraction)
    return assemble(idx, f0, spk, n_spk, cb)


def quant_awgn_f0(f0: np.ndarray, rng: np.random.Generator, bins: int = 16, noise_db: float = 2.0) -> np.ndarray:
    """Synthetic stand-in for f0_transformation="quant_16_awgn_2" (BASELINE configs[2]) on a normalised F0 track:
    quantize_f0 = round(x * bins) / bins with unvoiced (== 0) frames kept at 0, then awgn_f0 = + N(0, sqrt(10^(dB/10)))
    with unvoiced frames kept at 0 (/root/reference/satools/satools/hifigan/nn.py:28-62).  The reference draws the
    noise from torch's global CPU RNG; here it comes from `rng` so the workload is reproducible on any host."""
    uv = f0 == 0
    q = np.round(f0 * bins) / bins
    q[uv] = 0
    q = q + rng.normal(0.0, np.sqrt(10.0 ** (noise_db / 10.0)), size=q.shape).astype(np.float32)
    q[uv] = 0
    return q.astype(np.float32)


def batch(seed: int, frames: Sequence[int], n_spk: int = N_SPEAKERS, pad_to: Optional[int] = None,
          f0_transformation: str = "") -> np.ndarray:
    """Batch padded to the longest item the way the pipeline pads (zeros in BN/F0, but the
    speaker one-hot stays on: it is interpolated over the padded length, hifigan.py:94-97)."""
    rng = np.random.default_rng(seed)
    cb = codebook()
    T = max(frames) if pad_to is None else pad_to
    out = np.zeros((len(frames), N_BN + 1 + n_spk, T), dtype=np.float32)
    for b, n in enumerate(frames):
        u = utterance(rng, n, n_spk, cb)
        out[b, :, :n] = u
        out[b, N_BN + 1:, n:] = u[N_BN + 1:, :1]
    if f0_transformation:
        if f0_transformation != "quant_16_awgn_2":
            raise ValueError("only quant_16_awgn_2 is generated")
        out[:, N_BN] = quant_awgn_f0(out[:, N_BN], np.random.default_rng(seed + 7919))
    return out


def waveform(seed: int, seconds: float, sr: int = 16000) -> np.ndarray:
    """Speech-like synthetic waveform, float32 [n] in [-1, 1]: voiced segments (a harmonic series on a slowly moving F0 of
    80-300 Hz under a formant-like envelope), unvoiced noise bursts and silences -- what the YAAPT front end
    (`satools/hifigan/yaapt.py`) has to tell apart.  numpy PCG64: the same samples on any host."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    y = np.zeros(n, dtype=np.float64)
    t = 0
    while t < n:
        seg = int(rng.integers(int(0.08 * sr), int(0.45 * sr)))
        seg = min(seg, n - t)
        kind = rng.random()
        if kind < 0.55:                                     # voiced
            f0 = rng.uniform(80.0, 300.0) * (1.0 + 0.15 * np.sin(2 * np.pi * rng.uniform(1.0, 4.0) * np.arange(seg) / sr + rng.uniform(0, 6.28)))
            phase = 2 * np.pi * np.cumsum(f0) / sr
            s = np.zeros(seg)
            for h in range(1, 13):
                s += np.sin(h * phase + rng.uniform(0, 6.28)) / (h ** rng.uniform(0.8, 1.6))
            env = np.hanning(seg + 2)[1:-1] ** 0.5
            y[t:t + seg] = rng.uniform(0.1, 0.5) * env * s / 3.0
        elif kind < 0.8:                                    # unvoiced burst
            y[t:t + seg] = rng.uniform(0.01, 0.08) * rng.standard_normal(seg) * np.hanning(seg + 2)[1:-1]
        t += seg                                            # else: silence
    y += 1e-4 * rng.standard_normal(n)
    return np.clip(y, -1.0, 1.0).astype(np.float32)
