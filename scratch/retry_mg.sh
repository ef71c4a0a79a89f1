#!/bin/bash
# keeps asking for an N-GPU box until one call goes through (the pod is often busy)
N=$1; tries=${2:-12}
cd /root/repo
for i in $(seq 1 $tries); do
  /usr/local/graft/bin/gpurun --gpus $N --timeout 900 -- "scratch/multi_gpu_e2e.sh $N" > gpurun_out/mg${N}_call.log 2>&1
  if grep -q "status=ok" gpurun_out/mg${N}_call.log; then echo "done after $i tries"; tail -6 gpurun_out/mg${N}_call.log; exit 0; fi
  sleep 240
done
echo "gave up"; tail -3 gpurun_out/mg${N}_call.log
