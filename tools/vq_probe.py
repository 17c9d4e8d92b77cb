#!/usr/bin/env python3
"""Time sa_hifigan_vq_assign on the configs[1]-sized input (64 x 750 rows of 256, 48 codes); run under ncu for the profile."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sa-toolkit_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from satools_b200 import conditioning  # noqa: E402

gen = helpers.seeded_generator(2).to("cuda:0")
gen.precision = "fp16"
cb = conditioning.codebook()
gen.set_codebook(torch.from_numpy(cb))
rng = np.random.default_rng(9)
for rows in (64 * 750, 8 * 64 * 750):
    x = torch.from_numpy((cb[rng.integers(0, 48, size=rows)] + 0.5 * rng.standard_normal((rows, 256))).astype(np.float32)).cuda()
    for q in (False, True):
        for _ in range(3):
            gen.vq_assign(x, return_quantized=q)
        n = 5 if len(sys.argv) > 1 else 50
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(n):
            gen.vq_assign(x, return_quantized=q)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / n
        by = rows * 256 * 4 * (2 if q else 1) + rows
        print(f"rows {rows} quantized={q}: {ms * 1e3:.1f} us, {by / ms / 1e6:.0f} GB/s algorithmic, "
              f"{2 * rows * 256 * 48 / ms / 1e9:.1f} TFLOP/s fp32")
