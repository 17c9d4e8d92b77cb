#!/usr/bin/env python3
"""Print the in-kernel cycle accounting of the fused ResBlock kernels (SATOOLS_B200_CHAIN_TIMING=1)."""
import ctypes as C
import os
import sys

os.environ["SATOOLS_B200_CHAIN_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200")):
    sys.path.insert(0, p)
import torch
from satools_b200 import CoreHifiGan, conditioning, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
gen = CoreHifiGan(imput_dim=504, precision="fp16").to("cuda:0")
x = torch.from_numpy(conditioning.batch(7, [750] * B)).to("cuda:0")
for _ in range(2):
    gen(x)
gen.check()
lib = _lib.load()
buf = (C.c_int64 * (64 * 16))()
n = lib.sa_hifigan_chain_timing(gen._handle, buf, 64)
names = ["mma_total", "mma_wait_ready", "mma_wait_w", "mma_issue", "epi_total", "epi_load_x", "epi_wait_acc", "epi_work", "epi_tmem_ld", "epi_fence"]
labels = []
print("cycle counters per launch (timed launches in order: conv_tc launches = MMA warp [total, wait A, wait W, wait acc-empty], epilogue warp 0 [total, wait acc-full]; fused kernels as in sa_hifigan.h), summed over CTAs / 148")
for i in range(n):
    v = [buf[i * 16 + j] / 148.0 for j in range(14)]
    tot = v[0] or 1.0
    print(f"launch {i:2d} " + " ".join(f"{val/1e3:7.0f}k({100*val/ (tot if j < 4 else (v[4] or 1)):3.0f}%)" for j, val in enumerate(v)))

# Per-launch device time of the same forward (CUDA events) next to the MMA-warp cycle count: cycles / time is
# the SM clock the kernel actually ran at (clock64 counts SM cycles; under a power cap it is well below the
# nominal 1965 MHz), and the difference to the in-kernel total is launch + ramp + tail overhead.
prof = gen.profile(x, repeats=3)
kern = [(t, ms) for t, ms in prof if (t % 16) != 15 and t < 16 * 6]      # drop pack_input and the tail
print("launch  tag   event_ms   mma_total_kcycles   implied_MHz")
for i, (t, ms) in enumerate(kern[:n]):
    cyc = buf[i * 16 + 0] / 148.0
    print(f"{i:4d} {t:5d} {ms:9.4f} {cyc/1e3:12.0f} {cyc / (ms * 1e3) if ms > 0 else 0:12.0f}")
