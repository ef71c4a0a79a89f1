#!/bin/bash
# multi_gpu_e2e.sh N — the end-to-end scaling evidence at N GPUs: link probes with and without NUMA binding, then the bench
# line with and without it.  Results under gpurun_out/mg_N_*.json
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# never rebuild on the GPU box: a stale library must fail the run at once, not N processes racing nvcc
python -c "from gofr_b200 import _build, _abi; import sys; sys.exit(1 if _build._stale() else 0)" || { echo "library is stale: build it before gpurun"; exit 1; }
nvidia-smi topo -m > gpurun_out/mg_${N}_topo.txt 2>&1
cat /sys/fs/cgroup/cpu.max > gpurun_out/mg_${N}_cpumax.txt 2>&1
timeout 120 python scratch/pcie_probe.py --gpus $N --bind 1 --steps 10 | tee gpurun_out/mg_${N}_probe_bound.json
timeout 120 python scratch/pcie_probe.py --gpus $N --bind 0 --steps 10 | tee gpurun_out/mg_${N}_probe_unbound.json
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@"; }
run 2>gpurun_out/mg_${N}_bench_bound.err | tail -1 > gpurun_out/mg_${N}_bench_bound.json
run --no-numa-bind --no-extras 2>gpurun_out/mg_${N}_bench_unbound.err | tail -1 > gpurun_out/mg_${N}_bench_unbound.json
python - <<PY
import json
for k in ("bound", "unbound"):
    try:
        d = json.load(open("gpurun_out/mg_${N}_bench_%s.json" % k))
        print(k, "value %.2f G req/s" % (d["value"] / 1e9), "e2e %.1f M req/s  %.2f ms/step  link floor %.2f ms" % (d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["link_floor_ms_per_step"]), d.get("numa"), [kk for kk in d if "sharded" in kk])
    except Exception as e:
        print(k, "failed", e)
PY
