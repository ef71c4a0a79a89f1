#!/bin/bash
# ab.sh [reps] -- VARIANT...   runs kbench with the in-tree library and each scratch/variants/libgofr_VARIANT.so, alternating
reps=${1:-2}; shift; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for r in $(seq 1 $reps); do
  python scratch/kbench.py --check $KBENCH_ARGS | tee -a gpurun_out/ab.jsonl
  for v in "$@"; do
    GOFR_LIB_PATH=scratch/variants/libgofr_$v.so python scratch/kbench.py $KBENCH_ARGS | tee -a gpurun_out/ab.jsonl
  done
done
