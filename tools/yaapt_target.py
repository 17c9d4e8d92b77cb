#!/usr/bin/env python3
"""Fixed workload for ncu captures of the YAAPT front end: B utterances of 10-15 s, N calls."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from satools_b200 import conditioning, yaapt_frontend as yf

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lens = [int(v) * 320 for v in np.random.default_rng(0).integers(500, 751, size=B)]
x = torch.zeros(B, max(lens))
for b, m in enumerate(lens):
    x[b, :m] = torch.from_numpy(conditioning.waveform(500 + b, m / 16000.0)[:m])
x = x.to("cuda:0")
for _ in range(n):
    r = yf.nlfer(x, lengths=lens, frame_length=35.0, frame_space=20.0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    r = yf.nlfer(x, lengths=lens, frame_length=35.0, frame_space=20.0)
e1.record()
torch.cuda.synchronize()
opts = dict(frame_length=35.0, frame_space=20.0)
for _ in range(2):
    shc, cp, cm = yf.spec_shc(r, lengths=lens, candidates=True, **opts)
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(5):
    shc, cp, cm = yf.spec_shc(r, lengths=lens, candidates=True, **opts)
s1.record()
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    sp, sd = yf.spec_track(r, lengths=lens, **opts)
t1.record()
torch.cuda.synchronize()
fo = dict(opts, nccf_thresh1=0.25, tda_frame_length=25.0)
for _ in range(2):
    f0 = yf.yaapt(x, lengths=lens, **fo)
y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
y0.record()
for _ in range(5):
    f0 = yf.yaapt(x, lengths=lens, **fo)
y1.record()
torch.cuda.synchronize()
print(f"yaapt (whole extractor): {y0.elapsed_time(y1) / 5:.3f} ms per batch = {sum(lens) / 16000.0 / (y0.elapsed_time(y1) / 5e3):.0f} audio-s/s, "
      f"{float((f0 > 0).float().mean()):.2f} of the frames voiced")
print(f"spec_track (SHC + peaks + DP): {t0.elapsed_time(t1) / 5:.3f} ms per batch")
print(f"SHC: {s0.elapsed_time(s1) / 5:.3f} ms per batch ({int(r.vuv.sum())} voiced frames of {sum(r.nframes)})")
print(f"B={B}: {e0.elapsed_time(e1) / 5:.3f} ms per batch, {sum(lens) / 16000.0 / (e0.elapsed_time(e1) / 5e3):.0f} audio-s/s, voiced {float(r.vuv.float().mean()):.2f}")
