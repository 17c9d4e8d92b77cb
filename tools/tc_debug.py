#!/usr/bin/env python3
"""GPU debugging aid: per-stage SNR of a precision mode against the golden stage slices and the
fp32 CUDA path (full tensors), so a mismatch can be localised to one layer family."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import helpers
from satools_b200 import CoreHifiGan, conditioning

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
frames = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [41]
torch.manual_seed(0)
gen = CoreHifiGan(imput_dim=504).to("cuda:0")
x = torch.from_numpy(conditioning.batch(1001, frames)).to("cuda:0")
for tap in range(6):
    try:
        gen.precision = "fp32"
        y32, a32 = gen.forward_with_tap(x, tap)
        gen.precision = precision
        y, a = gen.forward_with_tap(x, tap)
        gen.check()
    except Exception as e:  # noqa
        print(f"tap {tap}: ERROR {e}")
        break
    a32n, an = a32.cpu().numpy(), a.cpu().numpy()
    err = np.abs(an - a32n)
    print(f"tap {tap}: shape {tuple(an.shape)} SNR {helpers.snr_db(a32n, an):.1f} dB  max-abs {err.max():.3e} "
          f"finite {np.isfinite(an).all()}  ref-rms {np.sqrt((a32n**2).mean()):.3e}")
    if helpers.snr_db(a32n, an) < 30:
        # where is it wrong?  per-channel-chunk and per-time-block error map
        B, C, L = an.shape
        e_c = err.reshape(B, C // 8, 8, L).max(axis=(0, 2, 3))
        print("   worst channel chunks:", np.argsort(-e_c)[:8].tolist(), "err", np.sort(e_c)[::-1][:4])
        e_t = err.max(axis=(0, 1))
        blk = 32
        nb = (L + blk - 1) // blk
        e_tb = [e_t[i * blk:(i + 1) * blk].max() for i in range(nb)]
        print("   err by 32-row block:", " ".join(f"{v:.1e}" for v in e_tb[:24]))
        print("   sample got:", an[0, :4, :6].round(4).tolist())
        print("   sample ref:", a32n[0, :4, :6].round(4).tolist())
        break
yn, y32n = y.cpu().numpy(), y32.cpu().numpy()
print(f"output: SNR {helpers.snr_db(y32n, yn):.1f} dB max-abs {np.abs(yn - y32n).max():.3e}")
