"""CPU ORACLE / CPU BASELINE (test + bench infrastructure, not product code).

Functional torch-CPU port of the reference generator forward: the very ATen ops the
reference executes on CPU (conv1d / conv_transpose1d / leaky_relu / reflection pad /
tanh in fp32, oneDNN/MKL underneath), driven from a plain state dict so that it can
travel to the GPU box where /root/reference does not exist.

Follows /root/reference/satools/satools/hifigan/archi.py:77-91 and nn.py:168-175.
It is what ``bench.py`` times as ``cpu_baseline`` (kind "port") and as the
``--impl reference`` arm.  Pinned against the reference module's own output by
tests/test_oracle.py (golden fixtures from oracle/make_golden.py).

Unlike the reference module it folds weight-norm once (the reference re-folds on every
forward through the weight_norm pre-forward hooks) -- this favours the CPU baseline.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

UPSAMPLE_RATES = (5, 4, 4, 2, 2)
UPSAMPLE_KERNELS = (11, 8, 8, 4, 4)
RESBLOCK_KERNELS = (3, 7, 11)
RESBLOCK_DILATIONS = (1, 3, 5)


def fold(state: dict, dtype=torch.float32) -> dict:
    out = {}
    names = sorted({k.rsplit(".", 1)[0] for k in state})
    for n in names:
        if n + ".weight" in state:
            w = state[n + ".weight"].detach().to("cpu", torch.float64)
        else:
            g = state[n + ".weight_g"].detach().to("cpu", torch.float64)
            v = state[n + ".weight_v"].detach().to("cpu", torch.float64)
            norm = v.pow(2).sum(dim=tuple(range(1, v.dim())), keepdim=True).sqrt()
            w = g * v / norm
        out[n] = (w.to(dtype).contiguous(), state[n + ".bias"].detach().to("cpu", dtype).contiguous())
    return out


@torch.no_grad()
def generator_forward(p: dict, x: torch.Tensor) -> torch.Tensor:
    """p = fold(state).  x [B,Cin,T] on CPU -> y [B,1,320T+1]."""
    h = F.conv1d(x, *p["conv_pre"], padding=3)
    for i, (u, ku) in enumerate(zip(UPSAMPLE_RATES, UPSAMPLE_KERNELS)):
        h = F.leaky_relu(h, 0.1)
        h = F.conv_transpose1d(h, *p[f"ups.{i}"], stride=u, padding=(ku - u) // 2)
        xs = torch.zeros_like(h)
        for j, k in enumerate(RESBLOCK_KERNELS):
            r = 3 * i + j
            xr = h
            for m, d in enumerate(RESBLOCK_DILATIONS):
                xt = F.leaky_relu(xr, 0.1)
                xt = F.conv1d(xt, *p[f"resblocks.{r}.convs1.{m}"], dilation=d, padding=(k * d - d) // 2)
                xt = F.leaky_relu(xt, 0.1)
                xt = F.conv1d(xt, *p[f"resblocks.{r}.convs2.{m}"], padding=(k - 1) // 2)
                xr = xt + xr
            xs += xr
        h = xs / 3
    h = F.leaky_relu(h)
    h = F.pad(h, (1, 0), mode="reflect")
    return torch.tanh(F.conv1d(h, *p["conv_post"], padding=3))
