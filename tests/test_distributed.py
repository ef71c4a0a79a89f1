"""The N>1 path on CPU: two gloo ranks — table sealed on rank 0 and broadcast, requests sharded contiguously, each rank
serves its shard (kernel device code via tests/emu), shards concatenate to the unsharded oracle result."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, outdir: str):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gofr_b200 import dist as gd
    from gofr_b200 import spec as S
    from gofr_b200 import synth
    from gofr_b200.table import Table
    from tests.emu import emu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    image = Table(synth.config4_spec()).serialize() if rank == 0 else None
    image = gd.broadcast_table_image(image, rank)
    n = 6001
    lo, hi = gd.shard_range(n, rank, world)
    shard = synth.config4_batch(hi - lo, start=lo)  # request i is a pure function of (seed, i)
    out, off, meta = emu.serve(image, shard, S.http_date(1789974595))
    np.save(os.path.join(outdir, f"out{rank}.npy"), out[:int(off[-1])])
    np.save(os.path.join(outdir, f"meta{rank}.npy"), meta)
    t = gd.max_over_ranks(float(rank + 1))
    assert t == float(world)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_broadcast_and_shard():
    import torch.multiprocessing as mp
    from gofr_b200 import spec as S
    from gofr_b200 import synth
    from tests import oracle as O
    with tempfile.TemporaryDirectory() as d:
        port = _free_port()
        mp.spawn(_worker, args=(2, port, d), nprocs=2, join=True)
        outs = [np.load(os.path.join(d, f"out{r}.npy")) for r in range(2)]
        metas = [np.load(os.path.join(d, f"meta{r}.npy")) for r in range(2)]
    whole = synth.config4_batch(6001)
    o, f, m = O.OracleTable(synth.config4_spec()).serve(whole, S.http_date(1789974595))
    assert np.array_equal(np.concatenate(outs), o[:int(f[-1])])
    assert np.array_equal(np.concatenate(metas), m)


def test_stream_is_a_pure_function_of_the_index():
    from gofr_b200 import synth
    a = synth.config2_batch(1000, start=0)
    b = synth.config2_batch(400, start=600)
    assert np.array_equal(a.trace_ids[600:], b.trace_ids)
    assert np.array_equal(a.arena[600 * synth.C2_REQ_STRIDE:1000 * synth.C2_REQ_STRIDE], b.arena[:400 * synth.C2_REQ_STRIDE])
