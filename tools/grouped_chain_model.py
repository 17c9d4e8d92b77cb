#!/usr/bin/env python3
"""Numpy model of the data flow of group_chain_kernel (sa-toolkit_b200/csrc/chain_group_tc.cuh): the index arithmetic
of the block-Toeplitz ("grouped") fused ResBlock -- staged-tile layout, slice schedule, Toeplitz weight blocks, the
d-major position permutation that turns the dilated convs into dilation-1 convs, halo / zero-padding / keep rules --
executed element by element exactly as the kernel addresses shared memory and TMEM (flat 16-bit element offsets), and
compared with the oracle's ResBlock1 (oracle/hifigan_numpy.py).  Test infrastructure; run on CPU:

    python tools/grouped_chain_model.py

Layout.  C = 16 or 32 channels, G = 64 / C positions per 128-byte row.  A staged tile holds R = MS * 128 * G positions
(+ PAD rows each side).  One MMA (M = 128 rows, N = 64 = (g', co), K = 16) consumes "slice" q: the 16 channels
h = q % (C/16) of the position at offset c = q / (C/16) from the row's first position (minus the conv's left reach):
  A[m, kk]  = buf[ row0 + m, (c - pad) * C + h * 16 + kk ]          (flat element offset: + m * 64)
  B_q[g' * C + co, kk] = W[co, h * 16 + kk, j = c - g']  if 0 <= j < k else 0
so that D[m, g' * C + co] = sum_j sum_ci W[co, ci, j] * X[G * m + g' + j - pad, ci]: the k-tap conv of positions
G * m + g'.  n_slices = (G + k - 1) * C / 16 instead of G * k * C / 16 narrow MMAs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hifigan_numpy as onp  # noqa: E402

PAD_ROWS = 8          # 128-byte rows of slack on both sides of a staged tile (one 1024-byte swizzle atom)


def toeplitz_blocks(w, C, G):
    """w [Cout, Cin, k] -> B [n_slices][64][16] (fp32 here; 16-bit, SWIZZLE_32B K-major on the device)."""
    k = w.shape[2]
    cpp = C // 16
    n_slices = (G + k - 1) * cpp
    B = np.zeros((n_slices, 64, 16), dtype=w.dtype)
    for q in range(n_slices):
        c, h = divmod(q, cpp)
        for g in range(G):
            j = c - g
            if 0 <= j < k:
                B[q, g * C:(g + 1) * C, :] = w[:, h * 16:(h + 1) * 16, j]
    return B


def perm(tau, d, Q):
    """natural tile-local time -> d-major position"""
    return (tau % d) * Q + tau // d


def perm_inv(p, d, Q):
    return d * (p % Q) + p // Q


class Tile:
    def __init__(self, C, MS):
        self.C, self.G = C, 64 // C
        self.R = MS * 128 * self.G
        self.rows = MS * 128
        self.pad_el = PAD_ROWS * 64
        self.n_el = (self.rows + 2 * PAD_ROWS) * 64

    def new_buf(self):
        return np.zeros(self.n_el, dtype=np.float64)

    def store_pos(self, buf, pos, vec):            # one position (C channels) at position index pos (may exceed R)
        o = self.pad_el + pos * self.C
        assert 0 <= o and o + self.C <= self.n_el
        buf[o:o + self.C] = vec

    def mma_conv(self, buf, Bq, k, acc=None):
        """All sub-tiles of one conv: returns D [rows, 64]."""
        C, G = self.C, self.G
        cpp = C // 16
        pad = (k - 1) // 2
        D = np.zeros((self.rows, 64)) if acc is None else acc.copy()
        m = np.arange(self.rows)
        for q in range(Bq.shape[0]):
            a_off = self.pad_el + (-pad) * C + q * 16          # linear in q: (c - pad) * C + h * 16 with C = 16 * cpp
            idx = a_off + m[:, None] * 64 + np.arange(16)[None, :]
            A = buf[idx]
            D += A @ Bq[q].T
        return D


def half(v, bf16=False):
    if bf16:
        u = np.asarray(v, dtype=np.float32).view(np.uint32)
        r = ((u >> 16) & 1) + 0x7FFF
        return ((u + r) & 0xFFFF0000).view(np.float32).astype(np.float64)
    return np.asarray(v, dtype=np.float32).astype(np.float16).astype(np.float64)


def lrelu16(v, keep):
    h = half(v)
    h = np.maximum(h, half(h * half(0.1)))
    return np.where(keep, h, 0.0)


def run_block(x, convs1, convs2, k, dils, MS=2, halo=None, exact=False):
    """x [L, C] fp64 (one item).  Returns the ResBlock output [L, C] computed tile by tile the way the kernel does.
    exact=True skips the 16-bit rounding (checks the index math to 1e-12)."""
    L, C = x.shape
    T = Tile(C, MS)
    G, R = T.G, T.R
    n_pairs = len(dils)
    pads = []
    for d in dils:
        pads += [(k - 1) // 2 * d, (k - 1) // 2]
    H = sum(pads) if halo is None else halo
    assert H % G == 0
    valid = R - 2 * H
    out = np.zeros_like(x)
    rnd = (lambda v, keep: np.where(keep, np.where(v >= 0, v, 0.1 * v), 0.0)) if exact else lrelu16
    wq = (lambda w: w) if exact else (lambda w: half(w))
    B1 = [toeplitz_blocks(wq(convs1[m][0]), C, G) for m in range(n_pairs)]
    B2 = [toeplitz_blocks(wq(convs2[m][0]), C, G) for m in range(n_pairs)]
    n_tiles = (L + valid - 1) // valid
    tau = np.arange(R)
    for mt in range(n_tiles):
        t0 = mt * valid - H
        t = t0 + tau
        inside = (t >= 0) & (t < L)
        keep = inside & (tau >= H) & (tau < R - H)
        bufA, bufT = T.new_buf(), T.new_buf()
        # x load: residual -> TMEM region D2 [rows, 64]; lrelu(x) -> bufA natural
        xin = np.where(inside[:, None], x[np.clip(t, 0, L - 1)], 0.0)
        D2 = xin.reshape(T.rows, 64).copy()
        a0 = rnd(xin, inside[:, None])
        for i in range(R):
            T.store_pos(bufA, i, a0[i])
        cb = np.zeros(C)
        for m, d in enumerate(dils):
            Q = -(-R // d)
            # ---- conv1 (dilation d as a dilation-1 conv over d-major positions) ----
            D1 = T.mma_conv(bufA, B1[m], k)
            v = D1.reshape(R, C) + convs1[m][1][None, :]
            if d == 1:
                src_tau = tau
            else:
                src_tau = perm_inv(tau, d, Q)                   # the row owner of position p holds time tau = inv(p)
            ok = src_tau < R
            tt = t0 + src_tau
            ins = ok & (tt >= 0) & (tt < L)
            a = rnd(v, ins[:, None])
            for p in range(R):
                if ok[p]:
                    T.store_pos(bufT, int(src_tau[p]), a[p])     # natural order for conv2
            # ---- conv2 accumulates onto the residual in TMEM ----
            D2 = T.mma_conv(bufT, B2[m], k, acc=D2)
            cb = cb + convs2[m][1]
            xnew = D2.reshape(R, C) + cb[None, :]
            if m + 1 < n_pairs:
                dn = dils[m + 1]
                Qn = -(-R // dn)
                a = rnd(xnew, inside[:, None])
                # bufA keeps the previous (differently ordered) contents: stale positions only reach halo rows
                for i in range(R):
                    T.store_pos(bufA, int(perm(i, dn, Qn)) if dn > 1 else i, a[i])
            else:
                out[t[keep]] = xnew[keep]
    return out


def main():
    rng = np.random.default_rng(0)
    worst = 0.0
    for C in (16, 32):
        for k in (3, 7, 11):
            L = 3000 if C == 16 else 1700
            x = rng.standard_normal((L, C))
            c1 = [(rng.standard_normal((C, C, k)) * 0.1, rng.standard_normal(C) * 0.1) for _ in range(3)]
            c2 = [(rng.standard_normal((C, C, k)) * 0.1, rng.standard_normal(C) * 0.1) for _ in range(3)]
            ref = onp.resblock1(x.T[None], c1, c2, k)[0].T
            got = run_block(x, c1, c2, k, (1, 3, 5), exact=True)
            err = np.abs(got - ref).max()
            worst = max(worst, err)
            got16 = run_block(x, c1, c2, k, (1, 3, 5), exact=False)
            snr = onp.snr_db(ref, got16)
            # a wider common halo (whole-stage variant: all chains use the k = 11 halo)
            got_h = run_block(x, c1, c2, k, (1, 3, 5), halo=60, exact=True)
            print(f"C={C} k={k}: exact-mode max-abs {err:.2e} (halo 60: {np.abs(got_h - ref).max():.2e}); fp16 operands SNR {snr:.1f} dB")
    assert worst < 1e-10
    print("ok")


if __name__ == "__main__":
    main()
