"""Shared test helpers: golden loading, seeded generator weights, error metrics."""
import hashlib
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def state_sha256(state) -> str:
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(state[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


_GEN_CACHE = {}


def seeded_generator(seed: int, imput_dim: int = 504, **kw):
    """satools_b200.CoreHifiGan with the reference's random init for `seed` (CPU parameters)."""
    from satools_b200 import CoreHifiGan
    key = (seed, imput_dim, tuple(sorted(kw.items())))
    if key not in _GEN_CACHE:
        torch.manual_seed(seed)
        _GEN_CACHE[key] = CoreHifiGan(imput_dim=imput_dim, **kw)
    return _GEN_CACHE[key]


def case_input(name: str, meta: dict) -> np.ndarray:
    """Conditioning tensor of a golden generator case: regenerated from the seed (Cin 504), or shipped with the
    fixture (other input widths, tiny); the fixture's SHA-256 of x proves it is the tensor the reference saw."""
    from satools_b200 import conditioning
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    x = g["x"] if "x" in g.files else conditioning.batch(meta["cond_seed"], meta["frames"])
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest() == g["x_sha256"].tobytes(), "conditioning drifted"
    return x


def numpy_state(gen):
    return {k: v.detach().cpu().numpy() for k, v in gen.state_dict().items()}


def snr_db(ref, test) -> float:
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    den = float((err ** 2).sum())
    return float("inf") if den == 0 else 10.0 * np.log10(float((ref ** 2).sum()) / den)


def max_abs(ref, test) -> float:
    return float(np.abs(np.asarray(test, dtype=np.float64) - np.asarray(ref, dtype=np.float64)).max())


def stage_slices(t: np.ndarray) -> np.ndarray:
    """Same sampling as oracle/make_golden.py:slices."""
    C = t.shape[1]
    ch = [0, 1, C // 2, C - 1]
    return np.concatenate([t[:, ch, :48], t[:, ch, -48:]], axis=2)
