"""CPU: the N>1 path -- region chunks dealt to ranks, per-chunk result blobs gathered to the writer rank in chunk
order -- with world_size 2 (and 3) over the gloo backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from longcalld_b200 import shard


def test_deal_chunks_blocks_and_balance():
    for n, world in ((0, 2), (1, 2), (17, 2), (100, 8), (6000, 8)):
        parts = shard.deal_chunks(n, world)
        allc = np.sort(np.concatenate(parts)) if n else np.zeros(0, int)
        assert np.array_equal(allc, np.arange(n))
        own = shard.owner_of(n, world)
        for r, p in enumerate(parts):
            assert (own[p] == r).all()
        if n >= world * shard.CHUNK_BLOCK * 4:
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= shard.CHUNK_BLOCK
        # neighbouring chunks stay together inside a block
        for p in parts:
            if len(p) > 1:
                assert ((np.diff(p) == 1) | (np.diff(p) > shard.CHUNK_BLOCK - 1)).all()


def _blob(c):
    rng = np.random.default_rng(1000 + c)
    return rng.integers(0, 256, int(rng.integers(0, 5000)), dtype=np.uint8).tobytes()


def _worker(rank, world, port, n_chunks, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.deal_chunks(n_chunks, world)[rank]
    got = shard.gather_chunk_results(mine, [_blob(c) for c in mine], n_chunks, dst=0)
    if rank == 0:
        q.put([g == _blob(c) for c, g in enumerate(got)])
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_chunks", [(2, 37), (3, 50), (2, 3)])
def test_gather_chunk_results_gloo(world, n_chunks):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_chunks, q)) for r in range(world)]
    for p in procs: p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120); assert p.exitcode == 0
    assert len(ok) == n_chunks and all(ok)
