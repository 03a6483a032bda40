#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; ch=$2; shift; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-stream --chunk $ch > gpurun_out/bench_ov_$tag.json 2> gpurun_out/bench_ov_$tag.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_ov_$tag.json")); print("$tag chunk $ch", "$*", "CPI/s %.0f  ms/step %.2f" % (d["value"], d["ms_per_step"]))
except Exception as e: print("$tag fail", e, open("gpurun_out/bench_ov_$tag.err").read()[-600:])
PY
}
for ch in 16 24 32 48 64; do run ovc$ch $ch RDP_OVERLAP=1; done
run base32 32 RDP_OVERLAP=0
