#!/bin/bash
# closed-loop front-end load on one GPU; run from the repo root (on the GPU box)
set -e
D=gpurun_out/frontend; mkdir -p $D
python - <<PY
from gofr_b200 import synth
from gofr_b200.table import Table
spec, b = synth.config2_spec(), synth.config2_batch(65536)
open("$D/table.img", "wb").write(Table(spec).serialize())
b.desc.tofile("$D/desc.bin"); b.trace_ids.tofile("$D/ids.bin"); b.arena.tofile("$D/arena.bin")
PY
LIB=$(python -c "from gofr_b200 import _build; print(_build.LIB)")
g++ -O2 -std=c++17 -pthread scratch/frontend_bench/frontend_bench.cpp -o $D/frontend_bench "$LIB" -Wl,-rpath,$(dirname "$LIB")
CFGS=${CFGS:-"1,1,0 64,64,50 256,256,100 512,512,100 1024,1024,200 4096,4096,200 4096,1024,200"}
echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max /sys/fs/cgroup/cpu/cpu.cfs_quota_us /sys/fs/cgroup/cpu/cpu.cfs_period_us 2>/dev/null | tr "\n" " ")  nproc: $(nproc)"
for cfg in $CFGS; do
  thr0=$(grep -E "nr_throttled|throttled_usec|throttled_time" /sys/fs/cgroup/cpu.stat /sys/fs/cgroup/cpu/cpu.stat 2>/dev/null | tr "\n" " ")
  IFS=, read -r T B W <<< "$cfg"
  GOFR_FRONTEND_DEBUG=${DEBUG:-0} FRONTEND_BENCH_SLOW=${SLOW:-} timeout 120 $D/frontend_bench $D/table.img $D/desc.bin $D/ids.bin $D/arena.bin $T $B $W 3 | tee -a $D/results.jsonl
  echo "  cgroup throttling before: $thr0 after: $(grep -E 'nr_throttled|throttled_usec|throttled_time' /sys/fs/cgroup/cpu.stat /sys/fs/cgroup/cpu/cpu.stat 2>/dev/null | tr '\n' ' ')"
done
