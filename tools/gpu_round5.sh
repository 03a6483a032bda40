#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r5.json 2> gpurun_out/bench_r5.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r5.json"))
st=d["roofline"]["stages"]
print("chunk", d["config"]["chunk_cpis"], "CPI/s %.0f" % d["value"], "e2e", d["e2e"] and round(d["e2e"]["value"]), {k:(round(v["ms_total_per_step"],2), round(v.get("frac",0),3)) for k,v in st.items()}, "dets/cpi", d["detections_per_cpi_mean"], "status", d["status_or"])
PY
CMD="python bench.py --steps 1 --warmup 1 --cpis 64 --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k1_range|k2_doppler" -s 4 -c 2 -o gpurun_out/prof_k12 $CMD > gpurun_out/ncu_k12.log 2>&1
echo "ncu exit $?"
