#!/bin/sh
# end-of-round evidence: full captures of the three kernels, the launch list of the default bench command, secondary workloads
set -x
timeout 300 ncu --set full --import-source on --clock-control none -k regex:serve_kernel -s 3 -c 1 -o gpurun_out/r01_serve_v11 python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/r01_serve_v11.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:reqlog_kernel -s 3 -c 1 -o gpurun_out/r01_reqlog python bench.py --workload reqlog --steps 3 --warmup 3 > gpurun_out/r01_reqlog.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r01_launches_bench.log 2>&1
for w in config3 config4 config5 reqlog; do timeout 200 python bench.py --workload $w --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r01_bench_$w.json; done
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r01_bench_main.json 2> gpurun_out/r01_bench_main.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r01_bench_ref.json 2> gpurun_out/r01_bench_ref.err
cat gpurun_out/r01_bench_main.json
