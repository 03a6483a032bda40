#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for fmt in int16 complex64; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-stream --fmt $fmt > gpurun_out/bench_ab_$fmt.json 2> gpurun_out/bench_ab_$fmt.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab_$fmt.json")); st=d["roofline"]["stages"]
print("$fmt CPI/s %.0f" % d["value"], {k:(round(v["ms_total_per_step"],2), round(v.get("frac",0),3)) for k,v in st.items()}, "dets/cpi", d["detections_per_cpi_mean"])
PY
done
