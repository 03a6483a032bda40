#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu $BENCH_ARGS > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json"))
st=d["roofline"]["stages"]
print("chunk", d["config"]["chunk_cpis"], "CPI/s %.0f" % d["value"], "e2e", d["e2e"] and round(d["e2e"]["value"]), {k:(round(v["ms_total_per_step"],2), round(v.get("frac",0),3)) for k,v in st.items()}, "dets/cpi", d["detections_per_cpi_mean"], "status", d["status_or"])
PY
tail -3 gpurun_out/bench_q.err
