"""Multi-GPU parity on real GPUs (BASELINE configs 4 and 5): the sealed table is broadcast from rank 0 over NCCL, every
rank serves its contiguous shard of the stream, rank 0 gathers the shards and compares their concatenation with the
oracle serving the unsharded stream.   torchrun --nproc-per-node N scratch/multi_gpu_parity.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from gofr_b200 import _abi, dist as gd, spec as S, synth
from gofr_b200.engine import Engine
from gofr_b200.table import Table

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
date = S.http_date(1_700_000_000)

# ---- config 4: mixed 64-route table, 262 144 requests sharded by contiguous index range ----
N = 262144
spec = synth.config4_spec()
image = Table(spec).serialize() if rank == 0 else None
image = gd.broadcast_table_image(image, rank, dev)
eng = Engine(Table(image=image), local)
lo, hi = gd.shard_range(N, rank, world)
batch = synth.config4_batch(hi - lo, start=lo)
resp = eng.alloc_responses(batch.n, batch.n * 700 + 4096)
eng.serve_device(eng.upload(batch), date, resp)
out, off, meta = resp.to_host()
from tests import oracle as O
_o, _f, _m = O.OracleTable(spec).serve(batch, date)
_ok = np.array_equal(off, _f) and np.array_equal(meta, _m) and out.tobytes() == _o[:_f[-1]].tobytes()
print(f"rank {rank}: shard [{lo},{hi}) local parity {_ok}; bytes {int(off[-1])} vs {int(_f[-1])}", flush=True)
total = torch.tensor([int(off[-1])], dtype=torch.int64, device=dev)
sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
dist.all_gather(sizes, total)
sizes = [int(s.item()) for s in sizes]
cap = max(sizes)
buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
buf[:int(off[-1])] = torch.from_numpy(out).to(dev)
gathered = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
dist.gather(buf, gathered, dst=0)
if rank == 0:
    from tests import oracle as O
    full = synth.config4_batch(N)
    o1, f1, m1 = O.OracleTable(spec).serve(full, date)  # single thread: the threaded oracle (bench baseline) leaves gaps between its shards
    cat = b"".join(g[:s].cpu().numpy().tobytes() for g, s in zip(gathered, sizes))
    print("sizes", sizes, "sum", sum(sizes), "oracle", int(f1[-1]), flush=True)
    if cat != o1[:f1[-1]].tobytes():
        a = np.frombuffer(cat, np.uint8); b = o1[:len(cat)]
        k = int(np.nonzero(a != b[:len(a)])[0][0]) if len(a) <= len(b) else -1
        print("first difference at byte", k, bytes(a[max(0, k - 40):k + 40]), bytes(b[max(0, k - 40):k + 40]), flush=True)
    assert len(cat) == int(f1[-1]) and cat == o1[:f1[-1]].tobytes(), "config 4: sharded GPUs != unsharded oracle"
    print(f"config 4 on {world} GPUs: {N} requests, {len(cat)} response bytes, shards concatenate to the unsharded oracle output")

# ---- config 5: gRPC Hello frames, 1 Mi sharded ----
N5 = 1 << 20
lo, hi = gd.shard_range(N5, rank, world)
frames, foff = synth.config5_frames(hi - lo, start=lo) if "start" in synth.config5_frames.__code__.co_varnames else synth.config5_frames(hi - lo)
n5 = len(foff) - 1
d_in = torch.from_numpy(np.concatenate([frames, np.zeros(64, np.uint8)])).to(dev)
d_off = torch.from_numpy(foff.view(np.int32)).to(dev)
cap5 = int(frames.size) + 40 * n5
d_out = torch.zeros(cap5 + 64, dtype=torch.uint8, device=dev)
d_ooff = torch.zeros(n5 + 1, dtype=torch.int32, device=dev)
d_meta = torch.zeros(n5, dtype=torch.int32, device=dev)
_abi.check(_abi.lib().gofr_grpc_hello_device(eng._e, d_in.data_ptr(), d_off.data_ptr(), n5, d_out.data_ptr(), cap5, d_ooff.data_ptr(),
                                             d_meta.data_ptr(), torch.cuda.current_stream().cuda_stream), "grpc")
torch.cuda.synchronize()
from tests import oracle as O
o5, f5, m5 = O.grpc_hello(frames, foff)
ok = bool(np.array_equal(d_ooff.cpu().numpy().view(np.uint32), f5)) and d_out[:int(f5[-1])].cpu().numpy().tobytes() == o5[:f5[-1]].tobytes()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    assert int(flag.item()) == 1, "config 5: a rank disagrees with the oracle on its shard"
    print(f"config 5 on {world} GPUs: {N5} frames, every rank's shard identical to the oracle")
dist.barrier()
dist.destroy_process_group()
