#!/bin/bash
# round-2 call after the geometry fix (fifth CTA restored): GPU tests, kernel timings, then ncu + bench with the faster
# slot-layout residency for config 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab.jsonl
timeout 200 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/final_pytest_gpu.log 2>&1
tail -3 gpurun_out/final_pytest_gpu.log
kb() { timeout 60 python scratch/kbench.py --check "$@" | tee -a gpurun_out/ab.jsonl; }
GOFR_SLOT_CTAS=4 kb --workload config2 --tag wide4
GOFR_SLOT_CTAS=5 kb --workload config2 --tag five
kb --workload config2 --layout packed --tag packed
kb --workload config4 --n 262144 --tag c4
kb --workload config3 --n 65536 --tag c3
best=$(python - <<'PY'
import json
r = {}
for l in open("gpurun_out/ab.jsonl"):
    d = json.loads(l)
    if d["workload"] == "config2" and d["layout"] == "slots": r[d["tag"]] = d["kernel_ms"]
print(5 if r.get("five", 9) < r.get("wide4", 9) else 4)
PY
)
echo "faster residency for config 2: $best" | tee gpurun_out/final_residency.txt
export GOFR_SLOT_CTAS=$best
bash scratch/ncu_capture.sh r2_final
( time timeout 240 python bench.py --gpus 1 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2>&1 | tail -4
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(d['value'], d['roofline']['frac'], d['roofline']['kernel'], d['other_layout']['kernel_ms_per_launch'], d['e2e']['value'], d['geometry'])"
