"""Worker for the multi-GPU test of the sharded large-CPI mode (BASELINE config 4).  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/dist_sharded_worker.py [n_chirps n_samples]
Every rank holds n_chirps / N chirps of the same synthetic CPIs; the corner turn is K1's peer stores over NVLink
(CUDA IPC); the IPC handles and the barriers go over the package's own socket rendezvous (no torch in the product:
torchrun is only the launcher here).  Rank 0 checks the gathered result
against the single-GPU path on the whole CPI: RD map identical, detection list identical."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import fpga_fmcw_radar_processor_b200 as rdp
from oracle import stimulus as st


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n_chirps = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    kw = dict(cfar_ref_r=8, cfar_ref_d=8, cfar_guard_r=4, cfar_guard_d=2)            # config 4 "wide" CFAR
    n_cpi, n_ch = 2, 1
    cube, _ = st.synthetic_cube(n_cpi, n_ch, n_chirps, n_samples, seed=404, dtype="complex64", n_targets=(3, 5), sigma=40.0)
    cl = n_chirps // world
    ex = rdp.SocketExchange()
    sh = rdp.ShardedCpi(n_chirps=n_chirps, n_samples=n_samples, world=world, rank=rank, n_channels=n_ch, n_cpi=n_cpi,
                        input_format="complex64", device=local, max_dets_per_cpi=1 << 18, **kw)
    dets, local_map = sh.run(cube[:, :, rank * cl:(rank + 1) * cl], ex)
    t0 = time.perf_counter()
    for _ in range(3):
        dets, local_map = sh.run(cube[:, :, rank * cl:(rank + 1) * cl], ex)
    dt = (time.perf_counter() - t0) / 3
    gathered = ex.all_gather((dets, local_map))
    ok = True
    if rank == 0:
        with rdp.RadarCore(n_range=n_samples, n_doppler=n_chirps, n_channels=n_ch, input_format="complex64", max_cpis=n_cpi,
                           max_dets_per_cpi=1 << 18, device=local, **kw) as c:
            ref_map, ref_dets = (x.copy() for x in c.process(cube)[:2])
        full = np.concatenate([g[1] for g in gathered], axis=2)
        merged = rdp.merge_detections([g[0] for g in gathered], [0] * world)
        ok = np.array_equal(full, ref_map) and len(merged) == len(ref_dets) and len(ref_dets) > 0 and all(
            np.array_equal(merged[f], ref_dets[f]) for f in ("cpi", "channel", "range", "doppler", "mag"))
        print(f"sharded {n_chirps}x{n_samples} over {world} GPUs: map identical={np.array_equal(full, ref_map)} "
              f"dets {len(merged)} vs {len(ref_dets)}  {'PASS' if ok else 'FAIL'}  ({dt * 1e3:.1f} ms per {n_cpi}-CPI run, "
              f"host-driven phases)", flush=True)
    sh.close()
    ex.barrier()
    ex.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
