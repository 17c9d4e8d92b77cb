#!/usr/bin/env python3
"""Small vq_assign calls for compute-sanitizer (both kernels, ragged tiles, code groups, quantised output)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sa-toolkit_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import helpers  # noqa: E402
from oracle import vq_numpy as ovq  # noqa: E402

gen = helpers.seeded_generator(2).to("cuda:0")
rng = np.random.default_rng(1)
for n_codes, dim, rows in [(48, 256, 300), (100, 64, 129), (20, 250, 77), (255, 32, 1), (3, 4, 513)]:
    cb = rng.standard_normal((n_codes, dim)).astype(np.float32)
    x = (cb[rng.integers(0, n_codes, size=rows)] + 0.5 * rng.standard_normal((rows, dim))).astype(np.float32)
    gen.set_codebook(torch.from_numpy(cb))
    idx, q = gen.vq_assign(torch.from_numpy(x).cuda(), return_quantized=True)
    torch.cuda.synchronize()
    want, wq = ovq.assign(x, cb)
    ok = ovq.margin(x, cb) > 1e-5
    assert np.array_equal(idx.cpu().numpy().astype(np.int64)[ok], want[ok]) and np.array_equal(q.cpu().numpy()[ok], wq[ok])
    print(f"codes {n_codes} dim {dim} rows {rows}: ok")
