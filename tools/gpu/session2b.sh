#!/bin/bash
# round 2: ncu evidence of the current build -- launch list of one bench step, one full capture each of K3, K1, K2, K4.
# The reports are exported to CSV on the box (raw metrics + per-source-line counters) and only K3's report travels back.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --cpis 128 --no-cpu --no-e2e --no-stream ${EXTRA:-}"
$CMD > gpurun_out/plain_r2.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_r2.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches.csv $CMD > gpurun_out/ncu_launches_r2.log 2>&1; echo "launch list exit $?"
for K in ${KERNELS:-k3_cfar k1_range k2_doppler k4_emit k4_row_offsets}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o /tmp/r2_prof_$K $CMD > gpurun_out/ncu_r2_$K.log 2>&1; echo "$K exit $?"
  ncu -i /tmp/r2_prof_$K.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/r2_prof_$K.ncu-rep --page source --csv > gpurun_out/r2_ncu_${K}_source.csv 2>/dev/null
  ls -la /tmp/r2_prof_$K.ncu-rep
done
du -sh gpurun_out
