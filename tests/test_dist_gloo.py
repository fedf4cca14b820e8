"""CPU, world_size 2, gloo: the data-parallel host logic (batch sharding + the one init-time
broadcast) without GPUs."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tf2_b200.dist import broadcast_bytes, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 1023):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


class _FakeNetWork:
    """Stands in for tf2_b200.network.NetWork (which needs a GPU): the packed weight blob is 3000 bytes
    derived from the model handed to rank 0."""

    def __init__(self):
        self.calls, self.blob = [], None

    def InitFromCodes(self, model, q, max_images, variant):
        self.calls.append(("codes", max_images, variant))
        self.q = q
        self.blob = (np.arange(3000, dtype=np.int64) * model % 251).astype(np.uint8)

    def weight_blob_bytes(self):
        return self.blob.size

    def export_weight_blob(self, ptr):
        import ctypes
        ctypes.memmove(ptr, self.blob.ctypes.data, self.blob.size)

    def InitFromBlob(self, ptr, nbytes, max_images, variant, q=None):
        import ctypes
        self.calls.append(("blob", max_images, variant))
        self.q = q
        self.blob = np.frombuffer(ctypes.string_at(ptr, nbytes), dtype=np.uint8).copy()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 owns the "weight blob"; the others learn its size and content from one broadcast
        blob = torch.arange(0, 5000, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.empty(0, dtype=torch.uint8)
        got = broadcast_bytes(blob, dist, 0)
        ok = bool(torch.equal(got, torch.arange(0, 5000, dtype=torch.int64).to(torch.uint8)))
        # every rank processes its shard of a 9-image batch; gathering reproduces the batch order
        data = torch.arange(9 * 4, dtype=torch.float32).reshape(9, 4)
        a, b = shard_range(9, rank, world)
        mine = data[a:b] * 2
        parts = [torch.empty((shard_range(9, r, world)[1] - shard_range(9, r, world)[0], 4)) for r in range(world)]
        dist.all_gather(parts, mine) if len({p.shape for p in parts}) == 1 else None
        if len({p.shape for p in parts}) != 1:
            # uneven shards: exchange through padded tensors
            pad = torch.zeros((5, 4)); pad[:b - a] = mine
            gathered = [torch.zeros((5, 4)) for _ in range(world)]
            dist.all_gather(gathered, pad)
            parts = [gathered[r][:shard_range(9, r, world)[1] - shard_range(9, r, world)[0]] for r in range(world)]
        ok = ok and bool(torch.equal(torch.cat(parts), data * 2))
        # init_network_distributed: rank 0 loads the model, the others import the broadcast blob
        from tf2_b200.dist import init_network_distributed
        qtab = (np.arange(6 * 11, dtype=np.int64) % 9 - 5).astype(np.int8).reshape(6, 11)
        nw = init_network_distributed(_FakeNetWork(), dist, "cpu", model=7 if rank == 0 else None,
                                      q=qtab if rank == 0 else None, max_images=4, variant=1)
        ok = ok and nw.calls == [("codes" if rank == 0 else "blob", 4, 1)]
        ok = ok and np.array_equal(nw.q, qtab)          # every rank ends up with rank 0's Q table
        ok = ok and np.array_equal(nw.blob, (np.arange(3000, dtype=np.int64) * 7 % 251).astype(np.uint8))
        # device-timed numbers are combined as the max over ranks
        t = torch.tensor([1.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t) == float(world)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
