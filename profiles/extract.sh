#!/bin/sh
# usage: extract.sh <report.ncu-rep> <outdir>  — raw metrics, SASS page and nvdisasm of the current library
set -e
mkdir -p "$2/cub"; (cd "$2/cub" && rm -f *.cubin *.sass && cuobjdump -xelf all /root/repo/gofr_b200/libgofr_b200.so >/dev/null 2>&1 && for f in *serve_kernel.sm_100a.cubin; do nvdisasm --print-line-info $f > serve_kernel.sass; done)
ncu -i "$1" --page raw --csv > "$2/raw.csv" 2>/dev/null
ncu -i "$1" --page source --csv > "$2/src.csv" 2>/dev/null
