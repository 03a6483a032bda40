#!/bin/bash
# evidence for profiles/: one full ncu capture of the dominant kernel and the launch list of a short bench run
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --cpis 128 --no-cpu --no-e2e --no-stream"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k3_cfar -s 2 -c 1 -o gpurun_out/prof_k3_final $CMD > gpurun_out/ncu_k3_final.log 2>&1
echo "ncu exit $?"
CMD2="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-stream"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/launches_final.csv
