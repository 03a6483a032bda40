#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-stream > gpurun_out/bench_ov_$tag.json 2> gpurun_out/bench_ov_$tag.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_ov_$tag.json")); print("$tag", "$*", "CPI/s %.0f  ms/step %.2f" % (d["value"], d["ms_per_step"]), "dets", d["detections_per_cpi_mean"])
except Exception as e: print("$tag fail", e, open("gpurun_out/bench_ov_$tag.err").read()[-600:])
PY
}
run base RDP_OVERLAP=0
run ov RDP_OVERLAP=1
run ov_k3_2_k12_1 RDP_OVERLAP=1 RDP_K3_CTAS=2 RDP_K12_CTAS=1
run ov_k3_3_k12_1 RDP_OVERLAP=1 RDP_K3_CTAS=3 RDP_K12_CTAS=1
run ov_k3_2_k12_2 RDP_OVERLAP=1 RDP_K3_CTAS=2 RDP_K12_CTAS=2
run ov_k3_3 RDP_OVERLAP=1 RDP_K3_CTAS=3
