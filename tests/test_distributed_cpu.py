"""Host-side logic of the multi-GPU path, run on CPU with world_size 2 over gloo (no CUDA involved)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D_BLOB = 1024      # MPL_PEER_BLOB_BYTES (include/modppl_b200.h)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from modppl_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, n_loc = D.shard_range(1 << 10, rank, world)
        blob = bytes([rank]) * D.PEER_BLOB_BYTES
        blobs = D.all_gather_bytes(blob)
        mx = D.max_over_ranks(float(rank) + 0.5)
        q.put((rank, off, n_loc, [b[0] for b in blobs], [len(b) for b in blobs], mx))
    finally:
        dist.destroy_process_group()


def test_shard_range():
    sys.path.insert(0, ROOT)
    from modppl_b200.distributed import shard_range
    assert shard_range(1 << 24, 0, 8) == (0, 1 << 21)
    assert shard_range(1 << 24, 7, 8) == (7 << 21, 1 << 21)
    assert shard_range(1000, 1, 2) == (500, 500)
    with pytest.raises(ValueError):
        shard_range(1001, 0, 2)
    with pytest.raises(ValueError):
        shard_range(1000, 2, 2)


def test_blob_exchange_and_reduction_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][:3] == (0, 0, 512) and res[1][:3] == (1, 512, 512)
    for r in res:
        assert r[3] == [0, 1] and r[4] == [D_BLOB, D_BLOB]      # blobs arrive in rank order, intact
        assert r[5] == 1.5                                  # max over ranks


def test_island_resampling_plan_is_a_pure_function_with_survivors_in_place():
    sys.path.insert(0, ROOT)
    import numpy as np
    from modppl_b200.distributed import island_resampling_plan, island_seed
    # equal weights: nothing happens, the log mean weight is the common value
    r, anc, lm, ess = island_resampling_plan([-3.0] * 8, 0, 7)
    assert not r and list(anc) == list(range(8)) and abs(lm + 3.0) < 1e-12 and abs(ess - 8) < 1e-9
    # one island carries everything: every other island continues from it, and it stays where it is
    r, anc, lm, ess = island_resampling_plan([0.0] + [-50.0] * 7, 1, 7)
    assert r and list(anc) == [0] * 8 and abs(lm - np.log(1 / 8)) < 1e-9 and ess < 1.01
    # general: identical on every call (all ranks take the same decision), offspring counts within one of G w_g / sum w, survivors keep their place
    d = np.array([0.0, -1.0, -6.0, 0.5, -7.0, -0.2, -9.0, -8.0])
    a1 = island_resampling_plan(d, 5, 123)
    a2 = island_resampling_plan(d.copy(), 5, 123)
    assert a1[0] and np.array_equal(a1[1], a2[1]) and a1[2] == a2[2]
    w = np.exp(d - d.max()); w /= w.sum()
    counts = np.bincount(a1[1], minlength=8)
    assert counts.sum() == 8 and np.all(np.abs(counts - 8 * w) < 1.0 + 1e-9)
    assert all(a1[1][g] == g for g in range(8) if counts[g] > 0)
    assert abs(a1[2] - (d.max() + np.log(np.mean(np.exp(d - d.max()))))) < 1e-12
    assert len({island_seed(9, g) for g in range(8)}) == 8


def _island_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import torch.distributed as dist
    from modppl_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the gather every comparison is built on: one (delta, live buffer) pair per rank, in rank order
        t = torch.tensor([-1.5 * rank, float(rank & 1)], dtype=torch.float64)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        table = np.stack([o.numpy() for o in out])
        plan = D.island_resampling_plan(table[:, 0], 0, 3)
        q.put((rank, table.tolist(), bool(plan[0]), plan[1].tolist(), float(plan[2])))
    finally:
        dist.destroy_process_group()


def test_island_comparison_agrees_across_ranks_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29100 + (os.getpid() % 500)
    procs = [ctx.Process(target=_island_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:] == res[1][1:]                       # same table, same decision, same ancestors on both ranks
    assert res[0][1] == [[0.0, 0.0], [-1.5, 1.0]]
