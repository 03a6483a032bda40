#!/bin/bash
mkdir -p gpurun_out
for ch in 16 24 32 48 64; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-stream --chunk $ch > gpurun_out/bench_chunk$ch.json 2> gpurun_out/bench_chunk$ch.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_chunk$ch.json")); print("chunk $ch CPI/s %.0f ms/step %.2f" % (d["value"], d["ms_per_step"]))
except Exception as e: print("chunk $ch fail", e)
PY
done
