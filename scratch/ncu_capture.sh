#!/bin/bash
# ncu_capture.sh NAME [kbench args]  — one full ncu capture of the slot-layout serve kernel (1 launch) on the GPU box
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:serve_slots_kernel -s 4 -c 1 -o gpurun_out/$name -f \
    python scratch/kbench.py --steps 2 "$@" > gpurun_out/ncu_$name.log 2>&1
tail -2 gpurun_out/ncu_$name.log
