#!/bin/bash
# one `ncu --set full` capture of one launch of a kernel, after the same command ran clean without ncu
# usage: gpurun -- 'KERNEL=k3_cfar FMT=int16 [SKIP=2] bash tools/gpu/ncu_kernel.sh'
mkdir -p gpurun_out
FMT=${FMT:-int16}; KERNEL=${KERNEL:-k3_cfar}; SKIP=${SKIP:-2}
CMD="python bench.py --steps 1 --warmup 1 --cpis 128 --no-cpu --no-e2e --no-stream --fmt $FMT ${EXTRA:-}"
$CMD > gpurun_out/plain_$FMT.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s $SKIP -c 1 -o gpurun_out/prof_${KERNEL}_$FMT $CMD > gpurun_out/ncu_${KERNEL}_$FMT.log 2>&1
echo "ncu exit $?"
