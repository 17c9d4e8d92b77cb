#!/usr/bin/env python3
"""Small fixed workload for ncu captures: N forwards of a [B,504,750] batch in one precision."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200")):
    sys.path.insert(0, p)
import torch
from satools_b200 import CoreHifiGan, conditioning

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
gen = CoreHifiGan(imput_dim=504, precision=precision).to("cuda:0")
x = torch.from_numpy(conditioning.batch(7, [750] * B)).to("cuda:0")
for _ in range(n):
    gen(x)
gen.check()
print("launches per forward:", gen.last_launch_count)
