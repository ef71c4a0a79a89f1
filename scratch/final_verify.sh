#!/bin/sh
# round-end verification on one GPU: full gpu test suite, smoke, default bench, reference arm, secondary workloads,
# one ncu capture of the proto encoder
set -x
O=gpurun_out/final
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
for w in config3 config4 config5 proto reqlog http; do timeout 200 python bench.py --workload $w --steps 20 --warmup 3 2>/dev/null | tail -1 > $O/$w.json; done
timeout 200 ncu --set full --import-source on --clock-control none -k regex:proto_encode_kernel -s 3 -c 1 -o $O/proto_encode python bench.py --workload proto --steps 3 --warmup 3 > $O/ncu_proto.log 2>&1
ncu -i $O/proto_encode.ncu-rep --page raw --csv > $O/proto_encode_raw.csv 2>/dev/null
cut -c1-900 $O/bench_default.json
