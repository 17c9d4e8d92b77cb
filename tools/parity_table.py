#!/usr/bin/env python3
"""Measured parity of the CUDA path per precision mode and seed (GPU box): SNR and max-abs error against the fp64
oracle (oracle/hifigan_torch_cpu.py, pinned to the reference by tests/test_oracle.py) on 5 s utterances, plus one
full-size item of the bench batch.  Output goes to stdout; commit it as profiles/rN_parity_table.txt.

    python tools/parity_table.py > gpurun_out/parity_table.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np
import torch

import helpers
from oracle import hifigan_torch_cpu as otc
from satools_b200 import CoreHifiGan, conditioning


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    print("parity of the sm_100a path vs the fp64 oracle (random-init reference weights; SNR dB / max-abs)")
    print(f"{'case':44s} {'fp32':>22s} {'fp16':>22s} {'bf16':>22s}")
    cases = []
    for seed in (0, 1, 2):
        cases.append((f"seed {seed}: 1 x 5 s (configs[0] shape)", seed, conditioning.batch(100 + seed, [250])))
    cases.append(("seed 0: 1 x 5 s, quant_16_awgn_2 F0", 0, conditioning.batch(103, [250], f0_transformation="quant_16_awgn_2")))
    cases.append(("seed 0: 4 x 15 s dense randn input", 0, np.random.default_rng(9).standard_normal((4, 504, 750)).astype(np.float32)))
    rng = np.random.default_rng(42)
    frames = rng.integers(500, 751, size=64).tolist()
    xb = conditioning.batch(4242, frames, pad_to=750)
    cases.append(("seed 0: item 63 of the 64 x 15 s bench batch", 0, xb))
    for name, seed, x in cases:
        torch.manual_seed(seed)
        gen = CoreHifiGan(imput_dim=504)
        p64 = otc.fold(gen.state_dict(), torch.float64)
        gen = gen.to("cuda:0")
        pick = slice(63, 64) if x.shape[0] == 64 else slice(None)
        ref = otc.generator_forward(p64, torch.from_numpy(x[pick]).double()).numpy()
        row = []
        for precision in ("fp32", "fp16", "bf16"):
            if precision == "fp32" and x.shape[0] == 64:
                row.append("(fp32 mode: see 4 x 15 s)")
                continue
            gen.precision = precision
            y = gen(torch.from_numpy(x).to("cuda:0"))[0][pick].cpu().numpy()
            gen.check()
            row.append(f"{helpers.snr_db(ref, y):6.1f} dB / {helpers.max_abs(ref, y):.1e}")
        print(f"{name:44s} {row[0]:>22s} {row[1]:>22s} {row[2]:>22s}", flush=True)
        gen.release()


if __name__ == "__main__":
    main()
