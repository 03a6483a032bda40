#!/bin/bash
# GPU check used after every kernel change: parity tests, then short device-resident bench runs.
# usage: gpurun -- 'CONFIGS="base: k3v1:RDP_K3_VARIANT=1" [FMTS="int16 complex64"] [PYTEST_ARGS=..] bash tools/gpu/check.sh'
#   each CONFIGS entry is name:ENV=VAL,ENV=VAL (environment knobs of librdp, read at rdp_create)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
fi
B="python bench.py --steps ${STEPS:-5} --warmup 3 --no-e2e --no-cpu --no-stream ${BENCH_ARGS:-}"
for cfg in ${CONFIGS:-base:}; do
  name=${cfg%%:*}; envs=${cfg#*:}
  for fmt in ${FMTS:-int16 complex64}; do
    env $(echo $envs | tr ',' ' ') $B --fmt $fmt > gpurun_out/bench_${name}_$fmt.json 2> gpurun_out/bench_${name}_$fmt.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${name}_$fmt.json"))
    st = d["roofline"]["stages"]
    print("%-14s %-9s %7.0f CPI/s  ms/step %6.2f | K1 %.2f K2 %.2f K3 %.2f K4 %.2f ms | dets/cpi %.0f" % ("$name", "$fmt", d["value"], d["ms_per_step"],
          st["range"]["ms_total_per_step"], st["doppler"]["ms_total_per_step"], st["cfar"]["ms_total_per_step"], st["compact"]["ms_total_per_step"], d["detections_per_cpi_mean"]))
except Exception as e:
    print("$name $fmt failed:", e)
PY
  done
done
