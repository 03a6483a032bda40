#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --cpis 128 --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k1_range|k2_doppler" -s 4 -c 2 -o gpurun_out/prof_k12b $CMD > gpurun_out/ncu_k12b.log 2>&1
echo "ncu exit $?"
