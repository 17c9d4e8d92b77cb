#!/usr/bin/env python3
"""Per-section device time of one forward (CUDA events around every launch, averaged over repeats)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200")):
    sys.path.insert(0, p)
import torch
from satools_b200 import CoreHifiGan, conditioning

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(0)
gen = CoreHifiGan(imput_dim=504, precision="fp16").to("cuda:0")
x = torch.from_numpy(conditioning.batch(7, [750] * B)).to("cuda:0")
for _ in range(3):
    gen(x)
ragged = os.environ.get("RAGGED") == "1"            # RAGGED=1: 10-15 s items padded to 15 s, true lengths passed along
if ragged:
    import numpy as np
    frames = [int(v) for v in np.random.default_rng(0).integers(500, 751, size=B)]
    x = torch.from_numpy(conditioning.batch(7, frames)[:, :, :750]).to("cuda:0")
    if x.shape[2] < 750:
        x = torch.nn.functional.pad(x, (0, 750 - x.shape[2]))
    fwd = gen.forward
    gen.forward = lambda t: fwd(t, frames_per_item=frames)
prof = gen.profile(x, repeats=reps)
sec = {}
for tag, ms in prof:
    sec[tag // 16] = sec.get(tag // 16, 0.0) + ms
print("sections(ms):", " ".join(f"{k}:{v:.3f}" for k, v in sorted(sec.items())), "total:", f"{sum(sec.values()):.3f}")
if len(sys.argv) > 3:
    print("launches:", " ".join(f"{t}:{ms:.3f}" for t, ms in prof))
