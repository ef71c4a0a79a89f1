#!/bin/bash
# round-2 validation call: GPU tests, kernel timings (tree + residency overrides + 3-CTA variant), ncu captures, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab.jsonl
timeout 420 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r2_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu.log
kb() { timeout 60 python scratch/kbench.py --check "$@" | tee -a gpurun_out/ab.jsonl; }
kb --workload config2 --tag tree
kb --workload config4 --n 262144 --tag tree
kb --workload config3 --n 65536 --tag tree
GOFR_SLOT_CTAS=5 kb --workload config2 --tag force5
GOFR_SLOT_CTAS=4 kb --workload config4 --n 262144 --tag force4
GOFR_LIB_PATH=scratch/variants/libgofr_w3.so kb --workload config2 --tag w3
kb --workload config2 --layout packed --tag tree_packed
bash scratch/ncu_capture.sh r2_c2
bash scratch/ncu_capture.sh r2_c4 --workload config4 --n 262144
bash scratch/ncu_capture.sh r2_c3 --workload config3 --n 65536
( timeout 240 python __graft_entry__.py --smoke 2>&1 | tail -12 ) > gpurun_out/r2_smoke.log
tail -3 gpurun_out/r2_smoke.log
