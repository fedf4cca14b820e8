"""Data-parallel plumbing: images are independent (the reference iterates frames sequentially,
`Runtime_Engine/cnn/device/src/sequencer.cl:58-62`), so a batch shards over ranks with no data-path
collective.  The only collective is the init-time broadcast of the packed weight blob from rank 0
(SURVEY.md 8e).  `torch.distributed` is the transport (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) slice of `n_items` owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_bytes(buf, dist, src: int = 0):
    """Broadcasts a 1-D uint8 tensor whose length only `src` knows.  Returns the tensor every rank
    ends up with (same device / dtype as `buf` on src; other ranks pass an empty tensor on the
    target device)."""
    import torch
    n = torch.tensor([buf.numel() if dist.get_rank() == src else 0], dtype=torch.int64, device=buf.device)
    dist.broadcast(n, src)
    if dist.get_rank() != src:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=buf.device)
    dist.broadcast(buf, src)
    return buf


def broadcast_q(q, dist, device, src: int = 0):
    """The parsed Q table (int8 [rows][max_out_channel], a few KB) from `src` to every rank."""
    import numpy as np
    import torch
    if dist.get_rank() == src:
        has = torch.tensor([0 if q is None else 1, 0 if q is None else q.shape[0], 0 if q is None else q.shape[1]],
                           dtype=torch.int64, device=device)
    else:
        has = torch.zeros(3, dtype=torch.int64, device=device)
    dist.broadcast(has, src)
    if int(has[0]) == 0:
        return None
    if dist.get_rank() == src:
        buf = torch.from_numpy(np.ascontiguousarray(q, dtype=np.int8).view(np.uint8).reshape(-1)).to(device)
    else:
        buf = torch.empty(int(has[1]) * int(has[2]), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    return buf.cpu().numpy().view(np.int8).reshape(int(has[1]), int(has[2])).copy()


def init_network_distributed(nw, dist, device, model=None, q=None, max_images: int = 1, variant: int = 0):
    """Rank 0 loads the model (`model` = LoadModel output), every other rank receives the packed
    weights through one broadcast and imports them (tf2b_import_weight_blob).  Rank 0's Q table travels
    too (Runner.Run needs its first entry to quantise float images on every rank)."""
    import torch
    rank = dist.get_rank()
    if rank == 0:
        nw.InitFromCodes(model, q, max_images=max_images, variant=variant)
        blob = torch.empty(nw.weight_blob_bytes(), dtype=torch.uint8, device=device)
        nw.export_weight_blob(blob.data_ptr())
    else:
        blob = torch.empty(0, dtype=torch.uint8, device=device)
    blob = broadcast_bytes(blob, dist, 0)
    q = broadcast_q(q, dist, device, 0)
    if rank != 0:
        nw.InitFromBlob(blob.data_ptr(), int(blob.numel()), max_images=max_images, variant=variant, q=q)
    return nw
