"""CPU tests of the multi-GPU host logic: length-balanced sharding (no collective on the data
path), batching and chunking.  The N>1 path is exercised with world_size-2 gloo processes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from satools_b200 import scheduler


def _lengths(n=300, seed=0):
    rng = np.random.default_rng(seed)
    return np.clip(rng.lognormal(np.log(12.3 * 50) - 0.18, 0.6, n), 50, 1750).astype(int).tolist()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_is_a_partition_and_balanced(world):
    L = _lengths()
    parts = scheduler.shard(L, world)
    assert sorted(i for p in parts for i in p) == list(range(len(L)))
    loads = scheduler.shard_loads(L, parts)
    assert max(loads) - min(loads) <= max(L)          # LPT bound
    assert max(loads) <= 1.02 * sum(L) / world + max(L) / world


def test_shard_beats_reference_contiguous_split():
    L = sorted(_lengths(), reverse=True)              # adversarial order for a count-balanced split
    world = 8
    n = len(L)
    contiguous = [sum(L[r * n // world:(r + 1) * n // world]) for r in range(world)]
    lpt = scheduler.shard_loads(L, scheduler.shard(L, world))
    assert max(lpt) < 0.6 * max(contiguous)


def test_shard_edge_cases():
    assert scheduler.shard([], 4) == [[], [], [], []]
    assert scheduler.shard([10], 2) == [[0], []]
    with pytest.raises(ValueError):
        scheduler.shard([1], 0)


def test_batches_bound_items_and_padding():
    L = _lengths(500, 3)
    bs = scheduler.batches(range(len(L)), L, max_items=64, max_padded_frames=64 * 750)
    assert sorted(i for b in bs for i in b) == list(range(len(L)))
    for b in bs:
        assert len(b) <= 64 and len(b) * max(L[i] for i in b) <= 64 * 750 or len(b) == 1
    assert scheduler.padding_waste(bs, L) < 0.15
    naive = [list(range(i, min(i + 64, len(L)))) for i in range(0, len(L), 64)]
    assert scheduler.padding_waste(bs, L) < scheduler.padding_waste(naive, L)


def test_chunks_cover_and_carry_halo():
    cs = scheduler.chunks(3000, 512)
    assert cs[0] == (0, 532, 0, 512) and cs[-1][3] == 3000
    assert [c[2] for c in cs[1:]] == [c[3] for c in cs[:-1]]
    for rlo, rhi, klo, khi in cs:
        assert rlo == max(0, klo - 20) and rhi == min(3000, khi + 20)
    assert scheduler.chunks(10, 512) == [(0, 10, 0, 10)]


def _worker(rank, world, port, lengths, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = scheduler.shard(lengths, world)[rank]          # every rank derives the same plan locally
    done = torch.tensor([float(sum(lengths[i] for i in mine)), float(len(mine))])
    dist.barrier()
    dist.all_reduce(done)                                  # reporting only, as bench.py does
    torch.save({"mine": mine, "total": done.tolist()}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_covers_corpus(tmp_path):
    L = _lengths(101, 9)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, L, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt") for r in (0, 1))
    assert sorted(r0["mine"] + r1["mine"]) == list(range(len(L)))
    assert r0["total"] == r1["total"] == [float(sum(L)), float(len(L))]


def test_receptive_field_is_exactly_20_frames_each_side():
    """RECEPTIVE_HALO_FRAMES by exact perturbation of the fp64 oracle (random weights, so every tap is non-zero):
    changing frame f changes output samples of frames f-20 .. f+20 and nothing else, i.e. sample 1 + 320 k + s depends
    on frames [k - 20, k + 20] -- a window [klo - 20, khi + 20) reproduces frames [klo, khi)."""
    import numpy as np
    import torch
    import helpers
    from oracle import hifigan_torch_cpu as otc
    p = otc.fold(helpers.seeded_generator(0).state_dict(), torch.float64)
    T, f = 100, 50
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((1, 504, T)))
    y0 = otc.generator_forward(p, x).numpy()[0, 0]
    x[:, :, f] += 1.0
    y1 = otc.generator_forward(p, x).numpy()[0, 0]
    changed = np.nonzero(y0 != y1)[0]
    first_frame, last_frame = (changed.min() - 1) // 320, (changed.max() - 1) // 320
    assert (first_frame, last_frame) == (f - scheduler.RECEPTIVE_HALO_FRAMES, f + scheduler.RECEPTIVE_HALO_FRAMES)
