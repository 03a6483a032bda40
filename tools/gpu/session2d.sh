#!/bin/bash
# full parity suite + bench lines: default, coherent beams
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-stream"
$B > gpurun_out/bench_q_base.json 2> gpurun_out/bench_q_base.err; echo "base $?"
$B --combine beams > gpurun_out/bench_q_beams.json 2> gpurun_out/bench_q_beams.err; echo "beams $?"
python - <<'PY'
import json
for f in ("base", "beams"):
    try:
        d = json.loads(open(f"gpurun_out/bench_q_{f}.json").read().strip().splitlines()[-1])
        st = d["roofline"]["stages"]
        print("%-8s %8.0f CPI/s" % (f, d["value"]), {k: round(v["ms_total_per_step"], 3) for k, v in st.items()}, "dets/cpi %.0f" % d["detections_per_cpi_mean"], "spot", d.get("parity_spot"))
    except Exception as e:
        print(f, "failed", e)
PY
