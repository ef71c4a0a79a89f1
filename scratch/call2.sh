#!/bin/bash
# round-2 second validation call: GPU tests, kernel timings after the values/no-values kernel split, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab.jsonl
timeout 300 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r2_pytest_gpu2.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu2.log
kb() { timeout 60 python scratch/kbench.py --check "$@" | tee -a gpurun_out/ab.jsonl; }
kb --workload config2 --tag tree
GOFR_SLOT_CTAS=5 kb --workload config2 --tag force5
kb --workload config2 --layout packed --tag tree_packed
GOFR_LIB_PATH=scratch/variants/libgofr_no256.so kb --workload config2 --layout packed --tag no256_packed
kb --workload config4 --n 262144 --tag tree
kb --workload config3 --n 65536 --tag tree
( timeout 240 python __graft_entry__.py --smoke 2>&1 | tail -12 ) > gpurun_out/r2_smoke2.log
tail -3 gpurun_out/r2_smoke2.log
