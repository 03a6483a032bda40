"""BASELINE config 4: one large CPI sharded over the GPUs of a box.  Launch (N = 2, 4 or 8):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_sharded.py [--cpis B] [--steps K] [--warmup W]
1024 chirps x 4096 samples complex64 per CPI, "wide" CFAR (ref 8/8, guard 4/2).  Every rank holds its n_chirps / N
chirps of B CPIs on the device; one step = range stage with the corner turn stored into the peers' memory over
NVLink (K1 scatter) -> barrier -> Doppler stage on the rank's range lines -> barrier -> CFAR halo lines from the
neighbours -> CFAR -> detections on the host.  Prints one JSON line: CPI/s over all ranks (max over ranks of
the wall time between barriers, CUDA-synchronised on both sides) and the all-to-all rate that goes with it."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")

import torch
import torch.distributed as dist

import fpga_fmcw_radar_processor_b200 as rdp
from fpga_fmcw_radar_processor_b200 import scenario as st      # synthetic input generator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpis", type=int, default=64, help="CPIs per step (all resident on the devices)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chirps", type=int, default=1024)
    ap.add_argument("--samples", type=int, default=4096)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kw = dict(cfar_ref_r=8, cfar_ref_d=8, cfar_guard_r=4, cfar_guard_d=2)
    nc, ns, B = args.chirps, args.samples, args.cpis
    cl = nc // world
    uniq = 2
    cube, _ = st.synthetic_cube(uniq, 1, nc, ns, seed=404, dtype="complex64", n_targets=(3, 5), sigma=40.0)
    local_cube = np.ascontiguousarray(cube[:, :, rank * cl:(rank + 1) * cl])
    local_cube = np.concatenate([local_cube] * (B // uniq), axis=0) if B > uniq else local_cube[:B]
    ex = rdp.SocketExchange(port_offset=7)       # IPC handles + barriers over TCP sockets (the product package is torch-free)
    sh = rdp.ShardedCpi(n_chirps=nc, n_samples=ns, world=world, rank=rank, n_channels=1, n_cpi=B, input_format="complex64",
                        device=local, max_dets_per_cpi=1 << 14, **kw)
    d_in = sh.upload_local(local_cube)

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()

    for _ in range(max(args.warmup, 3)):
        dets = sh.run_device(d_in, ex)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dets = sh.run_device(d_in, ex)
    sync_all()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    # per-phase host wall times of one more step, rank 0 (phases are separated by barriers)
    ph = {}
    def timed(name, fn):
        torch.cuda.synchronize(); a = time.perf_counter(); fn(); torch.cuda.synchronize(); ph[name] = (time.perf_counter() - a) * 1e3
    timed("range_scatter", lambda: sh.range_scatter_device(d_in)); ex.barrier()
    timed("doppler", sh.doppler); ex.barrier()
    timed("halo", sh.halo_exchange)
    timed("cfar", sh.cfar); ex.barrier()
    n_dets = ex.all_gather(int(len(dets)))
    if rank == 0:
        cpis = B * args.steps / dt
        a2a_bytes = nc * ns * 8 * (world - 1) / world               # bytes of one CPI that cross NVLink in the corner turn
        line = {"metric": "rd_map_os_cfar_cpis_per_sec", "value": cpis, "unit": "CPI/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"large_{nc}x{ns}x1ch_complex64_sharded_cpi", "cpis_per_step": B, "cfar": "ref 8/8 guard 4/2 (N_REF 480)",
                           "sharding": "chirps -> (K1 peer stores over NVLink, CUDA IPC) -> range lines; CFAR halo lines copied from the neighbours",
                           "barriers": "TCP sockets (exchange.SocketExchange), host-driven, 4 per step; torch.distributed only for this script's own timing"},
                "all_to_all": {"bytes_per_cpi": a2a_bytes, "gbs_sustained": a2a_bytes * cpis / 1e9,
                               "gbs_during_range_stage_per_gpu": (a2a_bytes / world) * B / (ph["range_scatter"] * 1e-3) / 1e9},
                "phase_ms_rank0": ph, "detections_per_step": int(sum(n_dets))}
        print(json.dumps(line), flush=True)
    d_in.free()
    sh.close()
    ex.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
