"""Host-side corpus driver for the generator: the synthesis part of the anonymize pipeline.

The reference batches utterances in file order, pads every item to the longest of its batch, generates the padded
length, copies the whole fp32 batch back, trims on the host and writes 16-bit PCM
(/root/reference/satools/satools/bin/pipeline.py:43-66,148-163), one process per GPU slot over count-balanced slices
(bin/anonymize:80-93).  This driver takes the conditioning of a set of utterances -- the assembled tensor
x = [BN | F0 | speaker] (hifigan.py:83-97) or the compact form (VQFeatures: code index + F0 per frame, speaker id) -- and

  * shards them over the ranks by length (scheduler.shard, no communication),
  * batches utterances of similar length (scheduler.batches) and pads with the pipeline's semantics (BN / F0 zero,
    speaker kept on); utterances longer than `chunk_frames` are cut into windows with the generator's receptive-field
    halo (20 frames) that are batched among themselves: the stitched waveform equals the unchunked one,
  * stages every batch into reusable pinned slabs on worker threads while the GPU works on the previous ones, submits it
    to the two-slot HostPipeline, and gets back ONLY the kept samples of every item (320 * frames + 1), as 16-bit PCM by
    default -- what pipeline.py:156-160 trims and writes -- so the D2H traffic is ~40 % of the padded fp32 batch,
  * hands every finished waveform to `sink(utt_id, samples)` (or collects them in the returned dict), trimmed to
    `original_len[utt]` samples when given (pipeline.py:156).
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import scheduler
from .pipeline import HostPipeline

N_BN_F0 = 257     # channels 0..256 are zero in padding; the speaker one-hot stays on (hifigan.py:94-97)
PAD_CODE = 255    # VQ index of a padding frame: >= n_codes selects the zero vector (include/sa_hifigan.h)


@dataclass
class VQFeatures:
    """Compact conditioning of one utterance: idx uint8 [T] (VQ codeword per frame), f0 float32 [T]
    (normalised / transformed, at T frames), spk (target speaker index)."""
    idx: np.ndarray
    f0: np.ndarray
    spk: int

    @property
    def frames(self) -> int:
        return int(self.idx.shape[0])


Features = Union[np.ndarray, VQFeatures]
_NP_OF = {torch.float32: np.float32, torch.float16: np.float16, torch.int16: np.int16}


def _frames(f: Features) -> int:
    return f.frames if isinstance(f, VQFeatures) else int(f.shape[1])


def _pad_batch(items: Sequence[np.ndarray], frames: int) -> np.ndarray:
    """Reference padding of a batch of assembled tensors (kept for tests and small callers)."""
    cin = items[0].shape[0]
    out = np.zeros((len(items), cin, frames), dtype=np.float32)
    for b, x in enumerate(items):
        n = x.shape[1]
        out[b, :, :n] = x
        out[b, N_BN_F0:, n:] = x[N_BN_F0:, -1:]
    return out


class _Slabs:
    """Reusable pinned staging memory: `n` input slabs (dense x, or idx + f0 + spk) and `n` output slabs."""

    def __init__(self, n: int, cin: int, max_items: int, max_padded_frames: int, max_out_samples: int, out_dtype, vq: bool):
        self.n = n
        if vq:
            self.idx = [torch.empty(max_padded_frames, dtype=torch.uint8).pin_memory() for _ in range(n)]
            self.f0 = [torch.empty(max_padded_frames, dtype=torch.float32).pin_memory() for _ in range(n)]
            self.spk = [torch.empty(max_items, dtype=torch.int32).pin_memory() for _ in range(n)]
        else:
            self.x = [torch.empty(cin * max_padded_frames, dtype=torch.float32).pin_memory() for _ in range(n)]
        self.out = [torch.empty(max_out_samples, dtype=out_dtype).pin_memory() for _ in range(n)]


_SLAB_CACHE: dict = {}


def _get_slabs(n, cin, max_items, max_padded_frames, max_out_samples, out_dtype, vq) -> _Slabs:
    """Pinned allocations cost milliseconds each: keep the largest set per (input form, output type) for the next call."""
    key = (vq, cin, out_dtype)
    c = _SLAB_CACHE.get(key)
    if c is None or c[0] < max_items or c[1] < max_padded_frames or c[2] < max_out_samples:
        caps = (max(max_items, c[0] if c else 0), max(max_padded_frames, c[1] if c else 0), max(max_out_samples, c[2] if c else 0))
        _SLAB_CACHE[key] = (*caps, _Slabs(n, cin, caps[0], caps[1], caps[2], out_dtype, vq))
    return _SLAB_CACHE[key][3]


@torch.no_grad()
def synthesize_corpus(gen, feats: Dict[str, Features], *, rank: int = 0, world_size: int = 1,
                      max_items: int = 64, max_padded_frames: int = 64 * 750, chunk_frames: int = 3000,
                      out_dtype: torch.dtype = torch.int16, device: Optional[str] = None,
                      original_len: Optional[Dict[str, int]] = None,
                      sink: Optional[Callable[[str, np.ndarray], None]] = None,
                      stats: Optional[dict] = None, staging_threads: Optional[int] = None) -> Dict[str, np.ndarray]:
    """gen: satools_b200.CoreHifiGan on a CUDA device (set_codebook() done when feats holds VQFeatures).
    feats: utterance id -> float32 [Cin, frames] or VQFeatures.  Returns utterance id -> waveform [samples] for the
    utterances of this rank (empty when `sink` consumes them).  out_dtype: torch.int16 (PCM16, default), float16, float32."""
    t_start = time.perf_counter()
    if staging_threads is None:
        # dense fp32 features are ~100 MB per batch: one copying thread moves ~5 GB/s, and with 8 ranks on one host the
        # staging of a batch (21 ms) was longer than its GPU step (10.5 ms); the items of a batch are copied by up to 4 threads
        staging_threads = max(2, min(4, (os.cpu_count() or 8) // max(1, world_size) - 1))
    ids = sorted(feats)
    lengths = [_frames(feats[u]) for u in ids]
    mine = scheduler.shard(lengths, world_size)[rank]
    dev = torch.device(device) if device is not None else next(gen.parameters()).device
    vq = bool(ids) and isinstance(feats[ids[0]], VQFeatures)
    cin = gen.imput_dim
    halo = scheduler.RECEPTIVE_HALO_FRAMES
    out: Dict[str, np.ndarray] = {}

    # ---- plan: work items (utt, read_lo, read_hi, keep_lo, keep_hi) grouped into batches -------------------------------
    # short utterances: one window each, batched by similar length and padded; long ones: halo windows, batched only with
    # windows of the same read length (padding a window would replace the true end-of-utterance context by padding frames)
    plan: List[Tuple[int, List[Tuple[int, int, int, int, int]]]] = []          # (padded frames T, items)
    short = [i for i in mine if lengths[i] <= chunk_frames]
    for batch in scheduler.batches(short, lengths, max_items=max_items, max_padded_frames=max_padded_frames):
        plan.append((max(lengths[i] for i in batch), [(i, 0, lengths[i], 0, lengths[i]) for i in batch]))
    long_wav: Dict[int, np.ndarray] = {}
    long_left: Dict[int, int] = {}
    by_len: Dict[int, List[Tuple[int, int, int, int, int]]] = {}
    for i in mine:
        if lengths[i] <= chunk_frames:
            continue
        windows = scheduler.chunks(lengths[i], chunk_frames, halo)
        long_left[i] = len(windows)
        for (rlo, rhi, klo, khi) in windows:
            by_len.setdefault(rhi - rlo, []).append((i, rlo, rhi, klo, khi))
    per_batch = max(1, min(max_items, max_padded_frames // max(1, chunk_frames + 2 * halo)))
    for T, items in sorted(by_len.items()):
        for k in range(0, len(items), per_batch):
            plan.append((T, items[k:k + per_batch]))
    if not plan:
        return out

    max_b = max(len(items) for _, items in plan)
    max_pf = max(T * len(items) for T, items in plan)
    max_os = max(sum(gen.output_length(rhi - rlo) for (_, rlo, rhi, _, _) in items) for _, items in plan)
    n_slabs = 3
    slabs = _get_slabs(n_slabs, cin, max_b, max_pf, max_os, out_dtype, vq)
    st = {"batches": len(plan), "utterances": len(mine), "frames": int(sum(lengths[i] for i in mine)),
          "padded_frames": int(sum(T * len(items) for T, items in plan)), "stage_s": 0.0, "stage_wait_s": 0.0,
          "collect_s": 0.0, "gpu_wait_s": 0.0}

    def stage(k: int):
        """Fill input slab k % n_slabs with batch k (runs on a worker thread; torch copies release the GIL)."""
        t0 = time.perf_counter()
        T, items = plan[k]
        B = len(items)
        s = k % n_slabs
        if vq:
            idx = slabs.idx[s][:B * T].view(B, T)
            f0 = slabs.f0[s][:B * T].view(B, T)
            spk = slabs.spk[s][:B]
            for b, (i, rlo, rhi, _, _) in enumerate(items):
                f = feats[ids[i]]
                n = rhi - rlo
                idx[b, :n] = torch.from_numpy(f.idx[rlo:rhi])
                f0[b, :n] = torch.from_numpy(f.f0[rlo:rhi])
                if n < T:
                    idx[b, n:] = PAD_CODE
                    f0[b, n:] = 0.0
                spk[b] = int(f.spk)
            view = (idx, f0, spk)
        else:
            x = slabs.x[s][:B * cin * T].view(B, cin, T)

            def copy_items(b0: int, b1: int):
                for b in range(b0, b1):
                    i, rlo, rhi, _, _ = items[b]
                    src = torch.from_numpy(feats[ids[i]])
                    n = rhi - rlo
                    x[b, :, :n] = src[:, rlo:rhi]
                    if n < T:
                        x[b, :N_BN_F0, n:] = 0.0
                        x[b, N_BN_F0:, n:] = src[N_BN_F0:, rhi - 1:rhi]

            n_par = min(staging_threads, B) if B * T * cin * 4 >= (32 << 20) else 1     # small batches: the hand-over costs more
            if n_par > 1:                                  # the items of one batch, copied by several threads
                cuts = [B * j // n_par for j in range(n_par + 1)]
                for fu in [copy_pool.submit(copy_items, cuts[j], cuts[j + 1]) for j in range(n_par)]:
                    fu.result()
            else:
                copy_items(0, B)
            view = (x,)
        return view, time.perf_counter() - t0

    def emit(i: int, wav: np.ndarray) -> None:
        u = ids[i]
        if original_len and u in original_len:
            wav = wav[:original_len[u]]
        if sink is not None:
            sink(u, wav)
        else:
            out[u] = wav.copy() if wav.base is not None else wav

    def collect(k: int, ticket: int) -> None:
        t0 = time.perf_counter()
        y = pipe.result(ticket)
        t1 = time.perf_counter()
        st["gpu_wait_s"] += t1 - t0
        y = y.numpy()
        T, items = plan[k]
        off = 0
        for (i, rlo, rhi, klo, khi) in items:
            n = gen.output_length(rhi - rlo)
            seg = y[off:off + n]
            off += n
            if i not in long_left:
                emit(i, seg)
                continue
            if i not in long_wav:
                long_wav[i] = np.empty(gen.output_length(lengths[i]), dtype=_NP_OF[out_dtype])
            lo, hi = 320 * (klo - rlo), 320 * (khi - rlo)
            if klo == 0:
                long_wav[i][:1 + 320 * khi] = seg[:1 + hi]
            else:
                long_wav[i][1 + 320 * klo:1 + 320 * khi] = seg[1 + lo:1 + hi]
            long_left[i] -= 1
            if long_left[i] == 0:
                emit(i, long_wav.pop(i))
        st["collect_s"] += time.perf_counter() - t1

    pipe = HostPipeline(gen, depth=2, device=dev)
    pool = ThreadPoolExecutor(max_workers=2)                       # two batches staged ahead
    copy_pool = ThreadPoolExecutor(max_workers=max(1, staging_threads))
    try:
        futures = {k: pool.submit(stage, k) for k in range(min(2, len(plan)))}
        pending: Optional[Tuple[int, int]] = None
        for k, (T, items) in enumerate(plan):
            t0 = time.perf_counter()
            view, dt = futures.pop(k).result()
            st["stage_wait_s"] += time.perf_counter() - t0
            st["stage_s"] += dt
            fpi = [rhi - rlo for (_, rlo, rhi, _, _) in items]
            o = slabs.out[k % n_slabs]
            if vq:
                ticket = pipe.submit_vq(view[0], view[1], view[2], fpi, out=o, out_dtype=out_dtype)
            else:
                ticket = pipe.submit(view[0], out=o, out_dtype=out_dtype, frames_per_item=fpi, trimmed=True)
            if pending is not None:
                collect(*pending)                 # batch k - 1: its input slab (k - 1) % 3 is free after this
            pending = (k, ticket)
            if k + 2 < len(plan):                 # slab (k + 2) % 3 == (k - 1) % 3: released by the collect above
                futures[k + 2] = pool.submit(stage, k + 2)
        if pending is not None:
            collect(*pending)
        pipe.drain()
    finally:
        pool.shutdown(wait=True)
        copy_pool.shutdown(wait=True)
    st["h2d_bytes"], st["d2h_bytes"], st["gpu_launches"] = pipe.h2d_bytes, pipe.d2h_bytes, pipe.launches
    st["seconds"] = time.perf_counter() - t_start
    if stats is not None:
        stats.update(st)
    return out
