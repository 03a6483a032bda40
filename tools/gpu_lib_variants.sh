#!/bin/bash
# A/B of library build variants (tools/exp/librdp_*.so) against the in-tree library: stage timing of a short bench
mkdir -p gpurun_out
cp fpga-fmcw-radar-processor_b200/librdp.so /tmp/librdp_base.so
for v in base $(ls tools/exp/ | sed 's/librdp_//; s/.so//'); do
  if [ "$v" = base ]; then cp /tmp/librdp_base.so fpga-fmcw-radar-processor_b200/librdp.so; else cp tools/exp/librdp_$v.so fpga-fmcw-radar-processor_b200/librdp.so; fi
  for fmt in int16 complex64; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-stream --fmt $fmt > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_v.json")); st=d["roofline"]["stages"]
    print("$v $fmt CPI/s %.0f" % d["value"], {k:round(v["ms_total_per_step"],2) for k,v in st.items()})
except Exception as e: print("$v fail", e, open("gpurun_out/bench_v.err").read()[-400:])
PY
  done
done
cp /tmp/librdp_base.so fpga-fmcw-radar-processor_b200/librdp.so
