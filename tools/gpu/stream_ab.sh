#!/bin/bash
# streaming ingest (BASELINE config 5) with and without the CUDA-graph replay of the kernel chain
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "streaming or c_abi" > gpurun_out/pytest_stream.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_stream.log
for g in 1 0; do
  RDP_STREAM_GRAPH=$g python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_streamgraph$g.json 2> gpurun_out/bench_streamgraph$g.err; echo "graph=$g exit $?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_streamgraph$g.json").read().strip().splitlines()[-1])
s = d["streaming"]
print("graph=$g  pipelined %.0f CPI/s p50 %.3f p99 %.3f ms | one in flight p50 %.3f p99 %.3f ms | e2e %.0f" % (s["value"], s["pipelined"]["latency_ms_p50"], s["pipelined"]["latency_ms_p99"], s["one_in_flight"]["latency_ms_p50"], s["one_in_flight"]["latency_ms_p99"], d["e2e"]["value"]))
PY
done
