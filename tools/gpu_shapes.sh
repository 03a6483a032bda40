#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-stream "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); st=d["roofline"]["stages"]
    print("$name", d["config"]["workload"], "CPI/s %.0f" % d["value"], "e2e", d["e2e"] and round(d["e2e"]["value"]), {k:(round(v["ms_total_per_step"],3), round(v.get("frac",0),3)) for k,v in st.items()}, "dets/cpi", d["detections_per_cpi_mean"], "status", d["status_or"])
except Exception as e: print("fail", e); print(open("gpurun_out/bench_$name.err").read()[-1500:])
PY
}
run c64 --fmt complex64
run rtl_native --shape rtl_native --cpis 4096 --e2e-cpis 1024
run large --shape large --fmt complex64 --cpis 64 --e2e-cpis 16 --unique 2 --max-dets 262144
