"""Length-balanced utterance scheduling for multi-GPU anonymization.

The reference shards ``wav.scp`` into contiguous slices of equal *count*, one process per
GPU slot, with no communication (/root/reference/satools/satools/bin/anonymize:80-93,
script_utils.py split_dict) and batches in file order, padding every item to the longest of
its batch (bin/pipeline.py:43-66).  Here utterances are spread by *length*:

  * shard(): longest-processing-time-first greedy assignment onto the ranks (cost = frames +
    a fixed per-utterance term), deterministic, no collective needed: every rank computes the
    same assignment from the same length list and takes its own part;
  * batches(): within a rank, sort by length and cut batches bounded by item count and by
    padded frames so padding waste stays small;
  * chunks(): split a long utterance into windows with the generator's receptive-field halo
    (+-20 frames) so chunked synthesis reproduces the unchunked result.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

RECEPTIVE_HALO_FRAMES = 20


def shard(lengths: Sequence[int], world_size: int, per_item_cost: int = 8) -> List[List[int]]:
    """Return world_size lists of utterance indices (LPT greedy, ties by lowest rank)."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i]) + per_item_cost
    return out


def shard_loads(lengths: Sequence[int], assignment: List[List[int]]) -> List[int]:
    return [sum(int(lengths[i]) for i in part) for part in assignment]


def batches(indices: Sequence[int], lengths: Sequence[int], max_items: int = 64,
            max_padded_frames: int = 64 * 750) -> List[List[int]]:
    """Group `indices` into batches of similar length: at most max_items items and at most
    max_padded_frames = items * longest frames per batch."""
    order = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    out: List[List[int]] = []
    cur: List[int] = []
    cur_max = 0
    for i in order:
        n = int(lengths[i])
        new_max = max(cur_max, n)
        if cur and (len(cur) + 1 > max_items or (len(cur) + 1) * new_max > max_padded_frames):
            out.append(cur)
            cur, new_max = [], n
        cur.append(i)
        cur_max = new_max
    if cur:
        out.append(cur)
    return out


def chunks(n_frames: int, chunk_frames: int, halo: int = RECEPTIVE_HALO_FRAMES) -> List[Tuple[int, int, int, int]]:
    """Windows (read_lo, read_hi, keep_lo, keep_hi) in frames: synthesize frames
    [read_lo, read_hi), keep output of frames [keep_lo, keep_hi)."""
    if chunk_frames < 1:
        raise ValueError("chunk_frames must be >= 1")
    out = []
    lo = 0
    while lo < n_frames:
        hi = min(lo + chunk_frames, n_frames)
        out.append((max(0, lo - halo), min(n_frames, hi + halo), lo, hi))
        lo = hi
    return out


def padding_waste(batch_list: Iterable[Sequence[int]], lengths: Sequence[int]) -> float:
    """Fraction of computed frames that are padding."""
    real = padded = 0
    for b in batch_list:
        m = max(int(lengths[i]) for i in b)
        padded += m * len(b)
        real += sum(int(lengths[i]) for i in b)
    return 0.0 if padded == 0 else 1.0 - real / padded
