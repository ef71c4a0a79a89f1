#!/bin/bash
# summarize.sh REPORT.ncu-rep OUTDIR PREFIX [KERNEL]  — key metrics, per-function and per-line summaries of an ncu capture
# of a kernel in serve_kernel.cu (the SASS with line info comes from the CURRENT in-tree library: capture and library
# must be the same build)
rep=$1; out=$2; pre=$3; export KERNEL=${4:-serve_slots_kernel}
mkdir -p $out /tmp/ncu_x
sh profiles/extract.sh $rep /tmp/ncu_x
python profiles/key_metrics.py /tmp/ncu_x/raw.csv > $out/${pre}_key_metrics.txt
python - >> $out/${pre}_key_metrics.txt <<'PY'
import csv
rows = list(csv.reader(open("/tmp/ncu_x/raw.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
print("\n# warp stall cycles per issued instruction")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        print(f"{h:75s} {vals[i]}")
PY
python profiles/by_function.py /tmp/ncu_x/src.csv /tmp/ncu_x/cub/serve_kernel.sass > $out/${pre}_by_function.txt
: > $out/${pre}_stalls_by_line.txt
for c in stall_long_sb stall_wait stall_no_inst stall_short_sb stall_branch_resolving stall_barrier; do
  python profiles/stall_by_line.py /tmp/ncu_x/src.csv /tmp/ncu_x/cub/serve_kernel.sass $c 10 >> $out/${pre}_stalls_by_line.txt 2>/dev/null; echo >> $out/${pre}_stalls_by_line.txt
done
cat $out/${pre}_key_metrics.txt; cat $out/${pre}_by_function.txt
