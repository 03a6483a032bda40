#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --cpis 128 --no-cpu --no-e2e --no-stream --fmt complex64"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k3_cfar -s 2 -c 1 -o gpurun_out/prof_k3d $CMD > gpurun_out/ncu_k3d.log 2>&1
echo "ncu exit $?"
