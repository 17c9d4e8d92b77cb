"""CPU ORACLE (test infrastructure, not product code) -- numpy restatement of the
SA-toolkit HiFi-GAN generator forward.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may
import this file.  The product path (``sa-toolkit_b200``) never does.

What it restates (reference file:line, under /root/reference):
  * CoreHifiGan.forward_resnet      satools/satools/hifigan/archi.py:77-91
  * ResBlock1.forward               satools/satools/hifigan/nn.py:168-175
  * get_padding                     satools/satools/hifigan/nn.py:17-18
  * weight_norm (old style, dim=0)  torch.nn.utils.weight_norm as used at archi.py:40,50,70-72
                                    and nn.py:98-165  (w = g * v / ||v||, norm over all dims but 0)
The arithmetic of the reference lives in PyTorch/ATen (Conv1d, ConvTranspose1d,
leaky_relu, ReflectionPad1d, tanh; reference pins torch==2.1.2, install.sh:27-28).
Their published semantics are restated here with plain numpy in float64.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4).  This oracle is pinned against outputs of the reference module
itself, executed in the build container by ``oracle/make_golden.py`` and committed
under ``tests/golden/`` (see tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np

UPSAMPLE_RATES = (5, 4, 4, 2, 2)            # hifigan.py:47
UPSAMPLE_KERNELS = (11, 8, 8, 4, 4)         # hifigan.py:48
RESBLOCK_KERNELS = (3, 7, 11)               # archi.py:27
RESBLOCK_DILATIONS = (1, 3, 5)              # archi.py:28
INITIAL_CHANNELS = 512                      # archi.py:26
LRELU_SLOPE = 0.1                           # archi.py:80, nn.py:170,172
FINAL_SLOPE = 0.01                          # archi.py:87 (F.leaky_relu default)


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """nn.py:17-18."""
    return int((kernel_size * dilation - dilation) / 2)


def leaky_relu(x: np.ndarray, slope: float) -> np.ndarray:
    return np.where(x >= 0, x, x * slope)


def fold_weight_norm(g: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Old-style weight_norm with dim=0: w = g * v / ||v||_2, the norm taken over every
    dim except 0 (dim 0 is Cout for Conv1d and Cin for ConvTranspose1d)."""
    v = np.asarray(v)
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
    return np.asarray(g, dtype=np.float64) * v.astype(np.float64) / norm   # float64; caller casts


def conv1d(x, w, b, dilation=1, padding=0):
    """torch.nn.functional.conv1d, stride 1, zero padding.
    x [B,Cin,L], w [Cout,Cin,k], b [Cout] -> [B,Cout,L + 2p - d(k-1)]."""
    B, Cin, L = x.shape
    Cout, Cin2, k = w.shape
    assert Cin == Cin2
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding)))
    Lout = L + 2 * padding - dilation * (k - 1)
    y = np.empty((B, Cout, Lout), dtype=x.dtype)
    y[:] = b[None, :, None]
    for j in range(k):
        seg = xp[:, :, j * dilation: j * dilation + Lout]
        y += np.matmul(w[None, :, :, j], seg)
    return y


def conv_transpose1d(x, w, b, stride, padding):
    """torch.nn.functional.conv_transpose1d restated as a polyphase filter bank.
    x [B,Cin,L], w [Cin,Cout,k] -> y [B,Cout,(L-1)*stride - 2*padding + k] with
      y[co, n] = b[co] + sum_ci sum_{j : (n + p - j) % stride == 0} x[ci, (n+p-j)/stride] * w[ci,co,j].
    Phase phi = n % stride uses taps j = (phi+p) % stride + stride*m, reading input
    index q + (phi+p)//stride - m for n = stride*q + phi."""
    B, Cin, L = x.shape
    Cin2, Cout, k = w.shape
    assert Cin == Cin2
    Lout = (L - 1) * stride - 2 * padding + k
    y = np.empty((B, Cout, Lout), dtype=x.dtype)
    y[:] = b[None, :, None]
    for phi in range(stride):
        n_idx = np.arange(phi, Lout, stride)
        q = n_idx // stride
        j0 = (phi + padding) % stride
        off = (phi + padding) // stride
        for m, j in enumerate(range(j0, k, stride)):
            i = q + off - m
            ok = (i >= 0) & (i < L)
            if not ok.any():
                continue
            seg = np.zeros((B, Cin, len(n_idx)), dtype=x.dtype)
            seg[:, :, ok] = x[:, :, i[ok]]
            y[:, :, n_idx] += np.matmul(w[:, :, j].T[None], seg)
    return y


def resblock1(x, convs1, convs2, k):
    """nn.py:168-175.  convs1/convs2: lists of (w, b) for dilations 1,3,5 / 1,1,1."""
    for m, d in enumerate(RESBLOCK_DILATIONS):
        w1, b1 = convs1[m]
        w2, b2 = convs2[m]
        xt = leaky_relu(x, LRELU_SLOPE)
        xt = conv1d(xt, w1, b1, dilation=d, padding=get_padding(k, d))
        xt = leaky_relu(xt, LRELU_SLOPE)
        xt = conv1d(xt, w2, b2, dilation=1, padding=get_padding(k, 1))
        x = xt + x
    return x


def folded_params(state: dict, dtype=np.float64) -> dict:
    """state: name -> ndarray using the reference's state-dict keys
    ({conv_pre,ups.N,resblocks.M.convs{1,2}.K,conv_post}.{weight_g,weight_v,bias}, or
    '.weight' after remove_weight_norm, archi.py:109-115).  Returns name -> (w, b)."""
    out = {}
    names = sorted({k.rsplit(".", 1)[0] for k in state})
    for n in names:
        if n + ".weight" in state:
            w = np.asarray(state[n + ".weight"])
        else:
            w = fold_weight_norm(np.asarray(state[n + ".weight_g"]), np.asarray(state[n + ".weight_v"]))
        out[n] = (w.astype(dtype), np.asarray(state[n + ".bias"]).astype(dtype))
    return out


def generator_forward(state: dict, x: np.ndarray, dtype=np.float64, return_stages=False):
    """archi.py:77-91 (+ :93-107 wrapper).  x [B,Cin,T] -> y [B,1,320*T+1].
    With return_stages also returns [conv_pre out, stage0..4 outputs]."""
    p = folded_params(state, dtype)
    h = conv1d(x.astype(dtype), *p["conv_pre"], dilation=1, padding=3)           # archi.py:78
    stages = [h]
    nk = len(RESBLOCK_KERNELS)
    for i, (u, ku) in enumerate(zip(UPSAMPLE_RATES, UPSAMPLE_KERNELS)):
        h = leaky_relu(h, LRELU_SLOPE)                                            # archi.py:80
        h = conv_transpose1d(h, *p[f"ups.{i}"], stride=u, padding=(ku - u) // 2)  # archi.py:81
        xs = np.zeros_like(h)                                                     # archi.py:82
        for j, k in enumerate(RESBLOCK_KERNELS):
            r = i * nk + j
            c1 = [p[f"resblocks.{r}.convs1.{m}"] for m in range(3)]
            c2 = [p[f"resblocks.{r}.convs2.{m}"] for m in range(3)]
            xs = xs + resblock1(h, c1, c2, k)                                     # archi.py:85
        h = xs / nk                                                               # archi.py:86
        stages.append(h)
    h = leaky_relu(h, FINAL_SLOPE)                                                # archi.py:87
    h = np.concatenate([h[:, :, 1:2], h], axis=2)                                 # ReflectionPad1d((1,0)) archi.py:75,88
    y = np.tanh(conv1d(h, *p["conv_post"], dilation=1, padding=3))                # archi.py:89-90
    if return_stages:
        return y, stages
    return y


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    den = float((err ** 2).sum())
    if den == 0.0:
        return float("inf")
    return 10.0 * np.log10(float((ref ** 2).sum()) / den)
