#!/bin/bash
# ab_frozen.sh VARIANT... — kbench with prebuilt libraries only (nothing is built on the GPU box); KBENCH_ARGS as in ab.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  GOFR_LIB_PATH=scratch/variants/libgofr_$v.so timeout 60 python scratch/kbench.py --check $KBENCH_ARGS --tag $v | tee -a gpurun_out/ab.jsonl
done
