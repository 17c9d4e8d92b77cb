#!/usr/bin/env python3
"""Tiny padded + ragged forwards for compute-sanitizer (memcheck): tools/sanitize_target.py [precision]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200")):
    sys.path.insert(0, p)
import torch
from satools_b200 import CoreHifiGan, conditioning

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
torch.manual_seed(0)
gen = CoreHifiGan(imput_dim=504, precision=precision).to("cuda:0")
frames = [70, 9, 41]
x = torch.from_numpy(conditioning.batch(3, frames)).to("cuda:0")
y0 = gen(x)[0]
y1 = gen(x, frames_per_item=frames)[0]
gen.check()
ok = all(torch.equal(y0[b, 0, :320 * f + 1], y1[b, 0, :320 * f + 1]) for b, f in enumerate(frames))
print("kept samples identical:", ok, "launches:", gen.last_launch_count)
