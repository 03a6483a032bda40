#!/bin/bash
# round 2, session 2, first GPU pass: full parity suite, smoke, default bench line, fixed-point chain timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-stream --arith fixed --cpis 256 > gpurun_out/bench_fixed.json 2> gpurun_out/bench_fixed.err; echo "fixed exit $?"
python - <<'PY'
import json
for f in ("bench_default", "bench_fixed"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        st = d["roofline"]["stages"]
        print(f, "%.0f CPI/s" % d["value"], "e2e", (d.get("e2e") or {}).get("value"), {k: round(v["ms_total_per_step"], 3) for k, v in st.items()}, "status_or", d["status_or"], "spot", (d.get("parity_spot") or {}).get("ok"))
    except Exception as e:
        print(f, "failed", e)
PY
