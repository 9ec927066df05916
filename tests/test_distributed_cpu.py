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
