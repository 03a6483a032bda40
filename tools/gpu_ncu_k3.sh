#!/bin/bash
# one `ncu --set full` capture of a K3 launch (FMT = int16 | complex64), after the same command ran clean without ncu
mkdir -p gpurun_out
FMT=${FMT:-int16}
CMD="python bench.py --steps 1 --warmup 1 --cpis 128 --no-cpu --no-e2e --no-stream --fmt $FMT"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k3_cfar -s 2 -c 1 -o gpurun_out/prof_k3_$FMT $CMD > gpurun_out/ncu_k3_$FMT.log 2>&1
echo "ncu exit $?"
