#!/usr/bin/env python
"""pcie_probe.py --gpus N [--bind 0|1] — the platform floor of the end-to-end path: N processes, one per GPU, each moving one
bench step's worth of bytes (256 MB host->device and 562 MB device->host, concurrently, pinned buffers, no kernel) per
step.  With --bind 1 every process first binds itself to the CPUs / memory of its GPU's NUMA node (gofr_bind_host_thread),
as bench.py does.  Prints one JSON line: ms per step (max over processes) and the aggregate link rates."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, n, bind, steps, h2d_bytes, d2h_bytes, barrier, out):
    import ctypes as C
    import torch
    torch.cuda.set_device(rank)
    node = -1
    if bind:
        from gofr_b200 import _abi
        nn = C.c_int(-1)
        rc = _abi.lib().gofr_bind_host_thread(rank, C.byref(nn))
        node = nn.value if rc == 0 else -2
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    h_in.fill_(1); h_out.fill_(0)
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    d_out = torch.ones(d2h_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def step():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        s1.synchronize(); s2.synchronize()
    for _ in range(3):
        step()
    barrier.wait(timeout=120)   # a worker that died must not leave the others (and the GPU box) waiting
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    barrier.wait(timeout=120)
    out.put((rank, dt / steps * 1e3, node, sorted(os.sched_getaffinity(0))[:2]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--bind", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--h2d", type=int, default=256 << 20)
    ap.add_argument("--d2h", type=int, default=(1 << 20) * 536)
    a = ap.parse_args()
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    barrier, out = ctx.Barrier(a.gpus), ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, a.gpus, a.bind, a.steps, a.h2d, a.d2h, barrier, out)) for r in range(a.gpus)]
    [p.start() for p in ps]
    res = sorted(out.get(timeout=300) for _ in ps)
    [p.join(timeout=30) for p in ps]
    ms = max(r[1] for r in res)
    print(json.dumps({"probe": "concurrent H2D + D2H, pinned, no kernel", "gpus": a.gpus, "numa_bound": bool(a.bind), "ms_per_step_max": round(ms, 3),
                      "ms_per_step_per_gpu": [round(r[1], 3) for r in res], "numa_node_per_gpu": [r[2] for r in res],
                      "h2d_GBps_aggregate": round(a.gpus * a.h2d / ms / 1e6, 1), "d2h_GBps_aggregate": round(a.gpus * a.d2h / ms / 1e6, 1),
                      "h2d_bytes": a.h2d, "d2h_bytes": a.d2h, "steps": a.steps}))


if __name__ == "__main__":
    main()
