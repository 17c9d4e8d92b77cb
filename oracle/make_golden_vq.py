#!/usr/bin/env python3
"""Mint tests/golden/vq_assign.npz by RUNNING THE REFERENCE's VectorQuantizerEMA (build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_vq.py

Imports /root/reference/satools (read-only), builds `VectorQuantizerEMA(48, 256, 0.25, 0.99)` as the VQ layer of the
ASR-BN extractor does (egs/asr/librispeech/local/chain/tuning/tdnnf_wav2vec2_vq.py:99-101), puts it in eval mode (what
`extract_bn` runs under, satools/satools/infer_helper.py:57-58) and records, per case, inputs [N, T, 256], the embedding,
`encoding_indices` and the returned `quantized`.  Inputs are noisy codewords (what a trained linearB emits) mixed with
unstructured rows.
"""
import os
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/satools")
sys.path.insert(0, ROOT)

import numpy as np
import torch

import satools.chain.nn as sann  # noqa: E402  (the reference)
from oracle import vq_numpy as ovq  # noqa: E402

CASES = [(0, 2, 37, 0.5), (1, 1, 1, 0.1), (2, 3, 129, 1.5), (3, 1, 64, 0.0)]     # (seed, N, T, noise)


def main():
    out = {}
    for seed, N, T, noise in CASES:
        torch.manual_seed(seed)
        vq = sann.VectorQuantizerEMA(48, 256, 0.25, 0.99).eval()
        cb = vq._embedding.weight.detach().clone()
        pick = torch.randint(0, 48, (N, T))
        x = cb[pick] + noise * torch.randn(N, T, 256)
        x[:, ::5] = torch.randn(N, (T + 4) // 5, 256)                              # unstructured rows
        with torch.no_grad():
            _, quantized, _, _, dist, enc = vq(x)
        idx = enc.reshape(N, T).numpy()
        oi, oq = ovq.assign(x.numpy(), cb.numpy())
        m = ovq.margin(x.numpy(), cb.numpy())
        print(f"case {seed}: [{N},{T},256] oracle == reference indices: {np.array_equal(oi, idx)}, quantized equal: "
              f"{np.array_equal(oq, quantized.numpy())}, min margin {m.min():.2e}")
        out[f"c{seed}_inputs"] = x.numpy()
        out[f"c{seed}_codebook"] = cb.numpy()
        out[f"c{seed}_indices"] = idx.astype(np.int64)
        out[f"c{seed}_quantized"] = quantized.numpy()
    path = os.path.join(ROOT, "tests", "golden", "vq_assign.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
