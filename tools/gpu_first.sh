#!/bin/bash
# First GPU pass: parity tests, compare-idiom micro-benchmark, a short bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 120 ./tools/microbench/mb_cmpacc > gpurun_out/mb_cmpacc.log 2>&1; cat gpurun_out/mb_cmpacc.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
