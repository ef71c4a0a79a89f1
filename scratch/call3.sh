#!/bin/bash
# round-2 final call: GPU tests, one full ncu capture of the headline kernel (traffic), the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/final_pytest_gpu.log 2>&1
tail -3 gpurun_out/final_pytest_gpu.log
bash scratch/ncu_capture.sh r2_final
( time timeout 240 python bench.py --gpus 1 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2>&1 | tail -4
tail -c 1500 gpurun_out/final_bench.json
