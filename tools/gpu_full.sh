#!/bin/bash
# full check on one B200: parity tests, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default.json")); st=d["roofline"]["stages"]
print("CPI/s %.0f" % d["value"], "e2e", d["e2e"] and round(d["e2e"]["value"]), "stream", d["streaming"] and round(d["streaming"]["value"]), {k:(round(v["ms_total_per_step"],2), round(v.get("frac",0),3)) for k,v in st.items()},
      "dets/cpi", d["detections_per_cpi_mean"], d.get("detections_per_cpi_max"), "status", d["status_or"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "clocks", d["clocks"])
PY
tail -3 gpurun_out/bench_default.err
if [ -n "$WITH_REF" ]; then timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 700 gpurun_out/bench_ref.json; fi
