#!/bin/sh
# end-of-round evidence: full captures of the kernels, the launch list of the default bench command, secondary workloads
set -x
timeout 300 ncu --set full --import-source on --clock-control none -k regex:serve_slots_kernel -s 3 -c 1 -o gpurun_out/r01_final_slots python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/r01_final_slots.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k "regex:serve_kernel" -s 3 -c 1 -o gpurun_out/r01_final_packed python bench.py --steps 3 --warmup 3 --no-e2e --layout packed > gpurun_out/r01_final_packed.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r01_final_launches.log 2>&1
for w in config3 config4 config5 reqlog; do timeout 200 python bench.py --workload $w --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r01_final_$w.json; done
timeout 200 python bench.py --workload config4 --layout packed --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r01_final_config4_packed.json
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r01_final_main.json 2> gpurun_out/r01_final_main.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r01_final_ref.json 2> gpurun_out/r01_final_ref.err
cat gpurun_out/r01_final_main.json
