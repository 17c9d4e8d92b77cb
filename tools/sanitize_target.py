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

# YAAPT front end + SHC on a small ragged batch
import numpy as np
from satools_b200 import yaapt_frontend as yf
lens = [9000, 4000, 700]
w = np.zeros((3, 9000), dtype=np.float32)
for b, n in enumerate(lens):
    w[b, :n] = conditioning.waveform(40 + b, n / 16000.0)[:n]
fr = yf.nlfer(torch.from_numpy(w).to("cuda:0"), lengths=lens, frame_length=35.0, frame_space=20.0)
shc, cp, cm = yf.spec_shc(fr, lengths=lens, candidates=True, frame_length=35.0, frame_space=20.0)
sp, sd = yf.spec_track(fr, lengths=lens, frame_length=35.0, frame_space=20.0) if min(fr.nframes) >= 4 else (None, None)
print("yaapt front end:", fr.nframes, int(fr.vuv.sum()), "voiced; SHC", tuple(shc.shape), bool(torch.isfinite(shc).all()))
f0 = yf.yaapt(torch.from_numpy(w[:2]).to("cuda:0"), lengths=lens[:2], frame_length=35.0, frame_space=20.0, nccf_thresh1=0.25, tda_frame_length=25.0)
print("yaapt final pitch:", tuple(f0.shape), float(f0.max()))
