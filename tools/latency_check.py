#!/usr/bin/env python3
"""Single-utterance latency (configs[0]: one 5 s utterance, T = 250) of the device entry, direct and as a CUDA graph,
for the current environment switches.  usage: [ENV=..] python tools/latency_check.py [T]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from satools_b200 import CoreHifiGan, conditioning

T = int(sys.argv[1]) if len(sys.argv) > 1 else 250
torch.manual_seed(0)
gen = CoreHifiGan(imput_dim=504, precision="fp16").to("cuda:0")
x = torch.from_numpy(conditioning.batch(7, [T])).to("cuda:0")
for _ in range(5):
    y = gen(x)[0]
torch.cuda.synchronize()
ts = []
for _ in range(50):
    t0 = time.perf_counter(); gen(x); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
g = gen.graphed(1, T)
for _ in range(5):
    g(x)
torch.cuda.synchronize()
tg = []
for _ in range(50):
    t0 = time.perf_counter(); g(x); torch.cuda.synchronize(); tg.append(time.perf_counter() - t0)
sw = {k: v for k, v in os.environ.items() if k.startswith("SATOOLS_B200")}
print(f"T={T} {sw}: direct {1e3*np.median(ts):.3f} ms, graph {1e3*np.median(tg):.3f} ms, launches {gen.last_launch_count}")
