#!/usr/bin/env python3
"""Bring-up check of the grouped (block-Toeplitz) narrow-stage kernel on a B200: stage taps and waveform against the fp64
oracle and against the per-tap kernels (SATOOLS_B200_GROUP=0), padded and ragged, then per-section times of both."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

os.environ.setdefault("SATOOLS_B200_GROUP_MIN_TILES", "0")     # small inputs here: dispatch the grouped kernels anyway

import helpers
from oracle import hifigan_numpy as onp
from satools_b200 import CoreHifiGan, conditioning


def make(group, precision="fp16"):
    os.environ["SATOOLS_B200_GROUP"] = "1" if group else "0"
    torch.manual_seed(0)
    g = CoreHifiGan(imput_dim=504, precision=precision).to("cuda:0")
    return g


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"):
        frames = [40, 33]
        x = conditioning.batch(12, frames)
        torch.manual_seed(0)
        state = {k: v.numpy() for k, v in CoreHifiGan(imput_dim=504).state_dict().items()}
        y_ref, stages = onp.generator_forward(state, x, return_stages=True)
        xd = torch.from_numpy(x).cuda()
        for precision in ("fp16", "bf16"):
            outs = {}
            for group in (0, 1):
                gen = make(group, precision)
                for tap in (4, 5):
                    _, act = gen.forward_with_tap(xd, tap)
                    a = act.cpu().numpy()
                    print(f"{precision} group={group} stage tap {tap}: SNR {helpers.snr_db(stages[tap], a):.1f} dB max-abs {helpers.max_abs(stages[tap], a):.2e} finite={np.isfinite(a).all()}", flush=True)
                y = gen(xd)[0]
                gen.check()
                outs[group] = y.cpu().numpy()
                print(f"{precision} group={group} waveform: SNR {helpers.snr_db(y_ref, outs[group]):.1f} dB max-abs {helpers.max_abs(y_ref, outs[group]):.2e} launches {gen.last_launch_count}", flush=True)
                yr = gen(xd, frames_per_item=frames)[0].cpu().numpy()
                gen.check()
                for b, f in enumerate(frames):
                    n = 320 * f + 1
                    assert np.array_equal(yr[b, 0, :n], outs[group][b, 0, :n]), f"ragged != padded, item {b}"
                gen.release()
            print(f"{precision} grouped vs per-tap kernels: SNR {helpers.snr_db(outs[0], outs[1]):.1f} dB", flush=True)
    if what in ("all", "time"):
        B = 64
        xb = torch.from_numpy(conditioning.batch(7, [750] * B)).cuda()
        for group in (0, 1):
            gen = make(group)
            for _ in range(3):
                gen(xb)
            gen.check()
            prof = gen.profile(xb, repeats=3)
            sec = {}
            for tag, ms in prof:
                sec[tag // 16] = sec.get(tag // 16, 0.0) + ms
            print(f"group={group} sections(ms):", " ".join(f"{k}:{v:.3f}" for k, v in sorted(sec.items())), "total:", f"{sum(sec.values()):.3f}", flush=True)
            print("   launches:", " ".join(f"{t}:{ms:.3f}" for t, ms in prof if t >= 64), flush=True)
            gen.release()


if __name__ == "__main__":
    main()
