"""Multi-GPU plumbing: the sealed route table is built on rank 0 and broadcast as bytes (NCCL on GPUs, gloo in the CPU
tests); requests are sharded contiguously with no data-path collective."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def broadcast_table_image(image: bytes | None, rank: int, device=None) -> bytes:
    """Rank 0 passes the serialized table; every rank returns the same bytes."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert image is not None
        return image
    dev = device if device is not None else torch.device("cpu")
    ln = torch.tensor([len(image) if rank == 0 else 0], dtype=torch.int64, device=dev)
    dist.broadcast(ln, 0)
    if rank == 0:
        buf = torch.frombuffer(bytearray(image), dtype=torch.uint8).to(dev)
    else:
        buf = torch.empty(int(ln.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, 0)
    return buf.cpu().numpy().tobytes()


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous request-index range of `rank`: outputs concatenate in order."""
    return n * rank // world, n * (rank + 1) // world


def max_over_ranks(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device if device is not None else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
