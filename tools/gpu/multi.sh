#!/bin/bash
# the driver's multi-GPU launch of bench.py (N = $1), one JSON line into gpurun_out/bench_${N}gpu.json
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N ${BENCH_ARGS:-} > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()[-1])
    e = d["e2e"]; c4 = d.get("config4") or {}
    print("N=%d  %.0f CPI/s  e2e %.0f CPI/s (%.1f of %.1f GB/s H2D ceiling)  config4 %s CPI/s parity %s nvlink %s" % (d["n_gpus"], d["value"], e["value"], e["h2d_achieved_gbs"], e["h2d_ceiling_gbs"], c4.get("value"), c4.get("parity_vs_single_gpu"), (c4.get("all_to_all") or {}).get("nvlink_gbs_per_gpu_during_k1")))
except Exception as ex:
    print("failed", ex)
PY
tail -3 gpurun_out/bench_${N}gpu.err
