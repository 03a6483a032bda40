"""Host-side logic: the reference's `r d mag` text records, CPI sharding across ranks, and a
world_size-2 gloo run of the multi-GPU plumbing (the data path itself has no collective)."""
import os
import socket
import sys

import numpy as np
import pytest

import fpga_fmcw_radar_processor_b200 as rdp
from oracle import float_path as fp
from oracle import stimulus as st


def _dets(rows):
    d = np.zeros(len(rows), dtype=rdp.DET_DTYPE)
    for i, r in enumerate(rows):
        d[i] = r
    return d


def test_detection_text_roundtrip(tmp_path):
    d = _dets([(0, 0, 100, 5, 8000.4), (0, 0, 500, 118, 5000.6), (0, 1, 7, 7, 12.0), (1, 0, 3, 2, 9.0)])
    p = tmp_path / "detections.txt"
    assert rdp.write_detections(p, d, cpi=0, channel=0) == 2
    assert open(p).read() == "100 5 8000\n500 118 5001\n"          # tb_radar_core.vhd:153-159 format
    assert rdp.read_detections(p).tolist() == [[100, 5, 8000], [500, 118, 5001]]
    with open(p, "a") as f:
        f.write("# junk line\nTRK 1 R=3\n")                      # the parser skips non-records
    assert rdp.read_detections(p).shape == (2, 3)


@pytest.mark.parametrize("n,w", [(1024, 8), (10, 4), (3, 8), (0, 2), (7, 1)])
def test_shard_cpis_partitions(n, w):
    blocks = [rdp.shard_cpis(n, w, r) for r in range(w)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
    sizes = [b - a for a, b in blocks]
    assert max(sizes) - min(sizes) <= 1


def test_merge_detections_canonical_order():
    a = _dets([(0, 0, 5, 5, 1.0), (1, 0, 2, 2, 2.0)])
    b = _dets([(0, 1, 9, 9, 3.0), (0, 0, 1, 1, 4.0)])
    m = rdp.merge_detections([a, b], [0, 2])
    assert m["cpi"].tolist() == [0, 1, 2, 2]
    assert m["channel"].tolist() == [0, 0, 0, 1]
    assert m["mag"].tolist() == [1.0, 2.0, 4.0, 3.0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_cpi, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        # every rank builds the same synthetic batch, keeps its shard, runs the path (here: the
        # oracle stands in for the GPU, which this container lacks), then lists are gathered
        cube, _ = st.synthetic_cube(n_cpi, 1, 32, 64, seed=9, sigma=20.0, n_targets=(1, 2))
        a, b = rdp.shard_cpis(n_cpi, world, rank)
        _, dets, counts = fp.process_batch(cube[a:b]) if b > a else (None, np.zeros(0, rdp.DET_DTYPE), np.zeros(0, np.int32))
        dets = dets.astype(rdp.DET_DTYPE)
        t = torch.tensor([1.0 + rank])                       # stand-in for the per-rank device time
        dist.all_reduce(t, op=dist.ReduceOp.MAX)              # max-over-ranks timing, as bench.py does
        gathered = [None] * world
        dist.all_gather_object(gathered, (a, dets))
        if rank == 0:
            merged = rdp.merge_detections([g[1] for g in gathered], [g[0] for g in gathered])
            q.put((float(t.item()), merged))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_merge():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, n_cpi, world = _free_port(), 5, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cpi, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    cube, _ = st.synthetic_cube(n_cpi, 1, 32, 64, seed=9, sigma=20.0, n_targets=(1, 2))
    _, ref, _ = fp.process_batch(cube)
    assert len(merged) == len(ref) and len(ref) > 0
    for f in ("cpi", "channel", "range", "doppler"):
        assert np.array_equal(merged[f], ref[f])
    assert np.array_equal(merged["mag"], ref["mag"])


def test_affinity_helpers_are_safe_without_a_gpu():
    """affinity.py is best-effort: parsing works, and with no GPU / sysfs entry it does nothing and returns None."""
    from fpga_fmcw_radar_processor_b200 import affinity
    assert affinity._cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert affinity.bind_to_gpu_numa_node(0) is None or isinstance(affinity.bind_to_gpu_numa_node(0), dict)
    assert os.sched_getaffinity(0) <= before


def _socket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ex = rdp.SocketExchange(rank=rank, world=world)
    try:
        got = ex.all_gather((rank, bytes([rank]) * 64))              # the 64-byte IPC handles of the sharded mode
        ex.barrier()
        again = ex.all_gather(rank * 10)
        q.put((rank, got, again))
        ex.barrier()
    finally:
        ex.close()


def test_socket_exchange_three_ranks():
    """The torch-free rendezvous of the sharded mode (exchange.SocketExchange): all_gather + barrier, world 3."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, world = _free_port(), 3
    procs = [ctx.Process(target=_socket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, again in res:
        assert got == [(r, bytes([r]) * 64) for r in range(world)] and again == [0, 10, 20]
    one = rdp.SocketExchange(rank=0, world=1)
    assert one.all_gather("x") == ["x"]
    one.barrier()


def test_track_log_formats_round_trip(tmp_path):
    """Both track-log formats of the reference (quick: tb_radar_core.vhd:163-181; full with VR= / S=:
    tb_tactical.vhd:350-363) read back the way model/ADR_visualize.py:124-168 parses them."""
    trk = np.zeros(3, dtype=rdp.TRACK_DTYPE)
    trk[0] = (0, 0, 4, 400, -12, 3, 0, 7, 2)
    trk[1] = (1, 0, 4, 412, -12, 3, 0, 9, 2)
    trk[2] = (1, 0, 5, 90, 40, -2, 1, 2, 3)
    active = np.array([[1], [2]], np.int32)
    for full in (False, True):
        n = rdp.write_tracks(tmp_path / "t.txt", trk, active, full=full)
        assert n == 5
        lines = open(tmp_path / "t.txt").read().split("\n")
        assert lines[0] == ("TRK 4 R=400 D=-12 VR=3 Q=7 S=10" if full else "TRK 4 R=400 D=-12 Q=7")
        assert lines[1] == "SCAN_END ACTIVE=1" and lines[4] == "SCAN_END ACTIVE=2"
        tracks, counts = rdp.read_tracks(tmp_path / "t.txt")
        assert counts == [1, 2] and tracks[4]["range_q2"] == [400, 412] and tracks[4]["scan"] == [0, 1]
        assert tracks[5]["quality"] == [2] and tracks[5]["doppler_q2"] == [40]


def test_tactical_scenario_generator():
    """The scenario generator (tb_tactical.vhd:28-40,149-171 analogue): moving targets show up where `truth` says."""
    sc = rdp.TacticalScenario(n_chirps=32, n_samples=128, targets=((40, 5, 6000.0, 2.0, 0.0), (90, 20, 4000.0, -1.0, 1.0)),
                              noise_sigma=20.0, clutter_amp=300.0, seed=7)
    from oracle import float_path as fp
    for scan in range(3):
        cube, truth = sc.next_cube()
        assert cube.shape == (1, 1, 32, 128, 2) and cube.dtype == np.int16
        m = fp.rd_map(cube[0])[0]
        for (r, d, a) in truth:
            r, d = int(round(r)) % 128, int(round(d)) % 32
            assert m[r, d] > 20 * np.median(m)
    assert sc.scan == 3 and sc.truth(0)[0][0] == 40 and sc.truth(2)[0][0] == 44
