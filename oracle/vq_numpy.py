"""TEST INFRASTRUCTURE -- CPU restatement of the last step of the ASR-BN extractor: the eval-mode
VectorQuantizerEMA.forward that `extract_bn` ends with (egs/asr/librispeech/local/chain/tuning/tdnnf_wav2vec2_vq.py:96-112,
:312; satools/satools/chain/nn.py:402-477).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

Parity pinned: tests/golden/vq_assign.npz holds inputs / codebook / encoding_indices / quantized produced by running the
reference's own module (oracle/make_golden_vq.py); tests/test_oracle.py checks this file against them.
"""
import numpy as np


def distances(flat_input, codebook, dtype=np.float64):
    """chain/nn.py:423-428: |x|^2 + |e|^2 - 2 x e^T, rows of flat_input against rows of the codebook."""
    x = np.asarray(flat_input, dtype=dtype)
    e = np.asarray(codebook, dtype=dtype)
    return (x * x).sum(axis=1, keepdims=True) + (e * e).sum(axis=1) - 2.0 * (x @ e.T)


def assign(inputs, codebook):
    """inputs [..., dim] -> (encoding_indices [...], quantized [..., dim] fp32).

    encoding_indices: chain/nn.py:436 (argmin over the codes, first minimum).  quantized: what the module returns in eval
    mode, chain/nn.py:448-456 -- the selected codeword passed through `inputs + (quantized - inputs)` in fp32 (the
    straight-through form; it differs from the codeword itself by at most one rounding)."""
    x = np.asarray(inputs, dtype=np.float32)
    e = np.asarray(codebook, dtype=np.float32)
    flat = x.reshape(-1, e.shape[1])
    idx = np.argmin(distances(flat, e), axis=1)
    q = (flat + (e[idx] - flat)).astype(np.float32)
    return idx.reshape(x.shape[:-1]), q.reshape(x.shape)


def margin(inputs, codebook):
    """Gap between the two smallest distances of every row (fp64), relative to |x|^2 + max |e|^2 -- the magnitude the fp32
    formula of chain/nn.py:423-428 rounds at.  Rows with a gap near fp32 resolution (1e-7) are the ones where two correct
    fp32 implementations may pick different codes."""
    e = np.asarray(codebook, dtype=np.float64)
    x = np.asarray(inputs, dtype=np.float64).reshape(-1, e.shape[1])
    d = np.sort(distances(x, e), axis=1)
    scale = (x * x).sum(axis=1) + (e * e).sum(axis=1).max()
    return (d[:, 1] - d[:, 0]) / np.maximum(scale, 1e-30)
