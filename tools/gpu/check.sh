#!/bin/bash
# GPU check used after every kernel change: parity tests, then short device-resident bench runs (A/B knobs via env).
# usage: gpurun -- 'bash tools/gpu/check.sh [pytest-args]'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q "$@" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-stream"
for v in ${VARIANTS:-0}; do
  for fmt in int16 complex64; do
    RDP_K3_VARIANT=$v $B --fmt $fmt > gpurun_out/bench_v${v}_$fmt.json 2> gpurun_out/bench_v${v}_$fmt.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_v${v}_$fmt.json"))
    st = d["roofline"]["stages"]
    print("variant $v $fmt: %.0f CPI/s  ms/step %.2f | K1 %.2f K2 %.2f K3 %.2f K4 %.2f ms | dets/cpi %.0f" % (d["value"], d["ms_per_step"],
          st["range"]["ms_total_per_step"], st["doppler"]["ms_total_per_step"], st["cfar"]["ms_total_per_step"], st["compact"]["ms_total_per_step"], d["detections_per_cpi_mean"]))
except Exception as e:
    print("variant $v $fmt failed:", e)
PY
  done
done
