#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "range_stage or end_to_end or sharded_cpi_one" 2>&1 | tail -3
for c in 0 1; do
RDP_K1_CLUSTER=$c timeout 300 python bench.py --shape large --steps 5 --warmup 3 --no-cpu --no-e2e --no-stream > gpurun_out/bench_large_cl$c.json 2> gpurun_out/bench_large_cl$c.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_large_cl$c.json")); st=d["roofline"]["stages"]
    print("cluster=$c CPI/s %.0f" % d["value"], {k:(round(v["ms_total_per_step"],3), round(v.get("frac",0),3)) for k,v in st.items()})
except Exception as e: print("fail", e, open("gpurun_out/bench_large_cl$c.err").read()[-500:])
PY
done
