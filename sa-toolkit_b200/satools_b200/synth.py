"""Host-side corpus driver for the generator: the synthesis part of the anonymize pipeline.

The reference batches utterances in file order, pads every item to the longest of its batch,
generates the padded length, copies the whole batch back and trims at write time
(/root/reference/satools/satools/bin/pipeline.py:43-66,148-156), one process per GPU slot over
count-balanced slices (bin/anonymize:80-93).  This driver takes the conditioning tensors
(x = [BN | F0 | speaker], hifigan.py:83-97) of a set of utterances and

  * shards them over the ranks by length (scheduler.shard, no communication),
  * batches utterances of similar length (scheduler.batches) and pads with the pipeline's
    semantics (BN/F0 zero, speaker one-hot kept on),
  * splits utterances longer than `chunk_frames` into windows with the generator's receptive-field
    halo (20 frames): the stitched waveform equals the unchunked one,
  * trims every waveform to 320 * frames (+1: the reflect-pad sample, archi.py:88), or to
    `original_len` samples when given (pipeline.py:156),
  * returns float32 / float16 / PCM16 waveforms on the host.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import scheduler

N_BN_F0 = 257     # channels 0..256 are zero in padding; the speaker one-hot stays on (hifigan.py:94-97)


def _pad_batch(items: Sequence[np.ndarray], frames: int) -> np.ndarray:
    cin = items[0].shape[0]
    out = np.zeros((len(items), cin, frames), dtype=np.float32)
    for b, x in enumerate(items):
        n = x.shape[1]
        out[b, :, :n] = x
        out[b, N_BN_F0:, n:] = x[N_BN_F0:, -1:]
    return out


@torch.no_grad()
def synthesize_corpus(gen, feats: Dict[str, np.ndarray], *, rank: int = 0, world_size: int = 1,
                      max_items: int = 64, max_padded_frames: int = 64 * 750, chunk_frames: int = 3000,
                      out_dtype: torch.dtype = torch.float32, device: Optional[str] = None,
                      original_len: Optional[Dict[str, int]] = None) -> Dict[str, np.ndarray]:
    """gen: satools_b200.CoreHifiGan on a CUDA device.  feats: utterance id -> float32 [Cin, frames].
    Returns utterance id -> waveform [samples] for the utterances of this rank."""
    ids = sorted(feats)
    lengths = [int(feats[u].shape[1]) for u in ids]
    mine = scheduler.shard(lengths, world_size)[rank]
    dev = torch.device(device) if device is not None else next(gen.parameters()).device
    out: Dict[str, np.ndarray] = {}

    def run(batch_np: np.ndarray) -> np.ndarray:
        xh = torch.from_numpy(batch_np).pin_memory()
        y = gen.synthesize_host(xh, out_dtype=out_dtype, device=dev)
        return y.numpy()

    # Batches go through the two-deep pipeline: batch k+1 is padded, pinned and enqueued (its H2D copy runs)
    # while the kernels of batch k are still busy; the waveforms of batch k are trimmed after that.
    from .pipeline import HostPipeline
    pipe = HostPipeline(gen, depth=2, device=dev)

    def collect(ticket: int, batch: Sequence[int]) -> None:
        y = pipe.result(ticket).numpy()
        for b, i in enumerate(batch):
            out[ids[i]] = y[b, 0, :320 * lengths[i] + 1].copy()

    short = [i for i in mine if lengths[i] <= chunk_frames]
    pending = None
    for batch in scheduler.batches(short, lengths, max_items=max_items, max_padded_frames=max_padded_frames):
        T = max(max(lengths[i] for i in batch), 2)
        xh = torch.from_numpy(_pad_batch([feats[ids[i]] for i in batch], T)).pin_memory()
        ticket = pipe.submit(xh, out_dtype=out_dtype, frames_per_item=[max(lengths[i], 1) for i in batch])
        if pending is not None:
            collect(*pending)
        pending = (ticket, batch)
    if pending is not None:
        collect(*pending)
    pipe.drain()
    for i in mine:
        if lengths[i] <= chunk_frames:
            continue
        x, n = feats[ids[i]], lengths[i]
        wav = np.zeros(320 * n + 1, dtype=out[ids[short[0]]].dtype if short else
                       {torch.float32: np.float32, torch.float16: np.float16, torch.int16: np.int16}[out_dtype])
        windows = scheduler.chunks(n, chunk_frames)
        # Windows are batched only with windows of the same length: padding a window with extra frames is NOT
        # what the unchunked run sees at the end of the utterance (there the convs zero-pad; padded frames would
        # carry the speaker one-hot), and the difference would reach the last 20 frames.
        by_len: Dict[int, List[int]] = {}
        for k, (rlo, rhi, _, _) in enumerate(windows):
            by_len.setdefault(rhi - rlo, []).append(k)
        for T, ks in sorted(by_len.items()):
            for w0 in range(0, len(ks), max_items):
                group = ks[w0:w0 + max_items]
                y = run(np.stack([np.ascontiguousarray(x[:, windows[k][0]:windows[k][1]]) for k in group]))
                for j, k in enumerate(group):
                    rlo, rhi, klo, khi = windows[k]
                    lo, hi = 320 * (klo - rlo), 320 * (khi - rlo)
                    if klo == 0:
                        wav[:1 + 320 * khi] = y[j, 0, :1 + hi]
                    else:
                        wav[1 + 320 * klo:1 + 320 * khi] = y[j, 0, 1 + lo:1 + hi]
        out[ids[i]] = wav
    if original_len:
        for u in out:
            if u in original_len:
                out[u] = out[u][:original_len[u]]
    return out
