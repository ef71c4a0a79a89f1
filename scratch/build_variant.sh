#!/bin/bash
# build_variant.sh NAME [nvcc flags...] — builds the library with extra flags into scratch/variants/libgofr_NAME.so
# (A/B measurements through GOFR_LIB_PATH in one gpurun call; the in-tree library is rebuilt afterwards).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
GOFR_EXTRA_NVCC="$*" python -m gofr_b200._build >/dev/null
cp gofr_b200/libgofr_b200.so scratch/variants/libgofr_$name.so
python -m gofr_b200._build >/dev/null
echo scratch/variants/libgofr_$name.so
