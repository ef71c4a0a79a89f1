#!/bin/sh
# one measurement round on the GPU box: parity tests, resident bench, optional ncu capture ($1 = capture name)
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e > gpurun_out/iter.json 2> gpurun_out/iter.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/iter.json"))
print("kernel_ms", d["roofline"]["kernel_ms_per_launch"], "frac", d["roofline"]["frac"], "value", d["value"])
PY
if [ -n "$1" ]; then
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:serve_kernel -s 3 -c 1 -o gpurun_out/$1 python bench.py --steps 3 --warmup 3 --no-e2e --requests 262144 > gpurun_out/ncu_$1.log 2>&1
fi
