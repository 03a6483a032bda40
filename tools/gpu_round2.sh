#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for ch in 0 2 16 64; do
  timeout 600 python bench.py --steps 10 --warmup 3 --chunk $ch --no-cpu --no-e2e > gpurun_out/bench_chunk$ch.json 2> gpurun_out/bench_chunk$ch.err; echo "chunk $ch exit $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_chunk$ch.json"))
    st=d["roofline"]["stages"]
    print("chunk", d["config"]["chunk_cpis"], "CPI/s %.0f" % d["value"], {k:(round(v["ms_total_per_step"],2), round(v.get("frac",0),3)) for k,v in st.items()}, "dets/cpi", d["detections_per_cpi_mean"], "status", d["status_or"])
except Exception as e:
    print("parse fail", e); print(open("gpurun_out/bench_chunk$ch.err").read()[-2000:])
PY
done
