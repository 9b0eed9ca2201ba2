"""Multi-GPU parity of the keyframe-sharded ScanContext search with the in-kernel peer-memory result
exchange (nrs_sc_query_dev_sharded): one process per GPU, the merged (dist, idx, shift) must equal the
oracle's single sequential scan on every rank -- including exact ties across shards (lowest global
index wins), empty shards and ranges, and hundreds of back-to-back queries (slot reuse by parity).
Needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_sharded_peer_gpu.py -m gpu`;
skipped on a single-GPU box."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q, binary=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    nrs = importlib.import_module("navtech-radar-slam_b200")
    synth = importlib.import_module("navtech-radar-slam_b200.synth")
    from oracle import pyoracle as O
    sc = nrs.scancontext
    ctx = nrs.Context(rank)
    db, rev = synth.sc_database(3001, 200, seed=17, binary=bool(binary))
    if binary == 2:
        # ONE general-valued keyframe: the rank that owns it searches with the f32-filter kernel, the others with the
        # bit-mask kernel -- the exchange (one epoch per query group, one slot per query) must not care
        gen, _ = synth.sc_database(4, 0, seed=5, binary=False)
        db[1501] = gen[1]
    db[2500] = db[7]
    db[2601] = db[7]          # exact ties across shards: lowest global index must win on every rank
    sdb = sc.ShardedSCDatabase(rank, world, ctx=ctx, exchange="peer")
    assert sdb.exchange == "peer"
    for a in range(0, 3001, 137):
        sdb.append(db[a:a + 137])
    results = []
    queries = [(db[rev[0][0]], 0, rev[0][0] - 30), (db[7], 8, 3001), (db[7], 0, 3001), (db[100], 0, 0),
               (db[rev[5][0]], 3, 2999), (np.zeros(1200), 0, 3001), (db[3000], 0, 3000), (db[1], 1, 2)]
    queries += [(db[j], 0, j - 30) for j, _, _ in rev[10:60]]
    for q, a, b in queries:
        got = sdb.query(q, a, b)
        _, _, exp = O.sc_search(q, db, a, b, want_all=False)
        results.append((got, exp))
    # back-to-back device-side queries without host synchronisation in between (slot parity reuse)
    qd = torch.from_numpy(np.stack([db[j] for j, _, _ in rev[:64]])).to(dev)
    out = torch.zeros(64 * 16, dtype=torch.uint8, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    lo, hi = sc.local_range(0, 2900, rank, world)
    with torch.cuda.stream(stream):
        for rep in range(4):
            for k in range(64):
                sdb.db.query_dev_sharded(qd[k].data_ptr(), lo, hi, out.data_ptr() + 16 * k)
    stream.synchronize()
    recs = out.cpu().numpy().view([("dist", "<f8"), ("idx", "<i4"), ("shift", "<i4")])
    for k in range(64):
        _, _, exp = O.sc_search(db[rev[k][0]], db, 0, 2900, want_all=False)
        results.append(((float(recs["dist"][k]), int(recs["idx"][k]), int(recs["shift"][k])), exp))
    # batched: Q queries, one launch, ONE exchange (per-query records travel, merged once per launch)
    qs = np.stack([db[j] for j, _, _ in rev[:70]] + [db[7], np.zeros(1200)])
    for a, b in ((0, 2900), (8, 3001), (0, 0), (2, 3)):
        got = sdb.query_batch(qs, a, b)
        for k in range(qs.shape[0]):
            _, _, exp = O.sc_search(qs[k], db, a, b, want_all=False)
            results.append((got[k], exp))
    # back-to-back batched launches of different sizes without host synchronisation (slot reuse by parity)
    outb = torch.zeros(6 * 64 * 16, dtype=torch.uint8, device=dev)
    sizes = [64, 1, 17, 64, 3, 64]
    with torch.cuda.stream(stream):
        for i, n in enumerate(sizes):
            sdb.db.query_batch_dev_sharded(qd.data_ptr(), n, lo, hi, outb.data_ptr() + i * 64 * 16)
    stream.synchronize()
    recb = outb.cpu().numpy().view([("dist", "<f8"), ("idx", "<i4"), ("shift", "<i4")]).reshape(6, 64)
    for i, n in enumerate(sizes):
        for k in range(n):
            _, _, exp = O.sc_search(db[rev[k][0]], db, 0, 2900, want_all=False)
            results.append(((float(recb["dist"][i, k]), int(recb["idx"][i, k]), int(recb["shift"][i, k])), exp))
    out_q.put((rank, results))
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    sdb.db.close()
    ctx.close()
    out_q.close()
    out_q.join_thread()
    os._exit(0)          # results are delivered; skip interpreter finalisation of the CUDA/NCCL runtimes


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("binary", [0, 1, 2])
def test_peer_exchange_matches_single_scan(binary):
    """binary=0: general-valued descriptors (f32-filter kernel); 1: {0,3.0} descriptors (bit-mask kernel, batched
    launches with one exchange); 2: {0,3.0} except one keyframe, so the ranks run different kernels."""
    world = min(torch.cuda.device_count(), 4)
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q, binary)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, results in got:
        for res, exp in results:
            assert res == exp, (rank, res, exp)
    assert all(g[1] == got[0][1] for g in got)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_cpp_sharded_manager(mode, tmp_path):
    """The C++ multi-GPU host (navtech-radar-slam_b200/cpp/sharded_check.cpp: ShardedSCManagerT, one process per
    GPU, handles / NCCL id exchanged through files): detectLoopClosureID and batched searches identical to the
    single-GPU SCManager on every rank, with the in-kernel peer exchange and with the two NCCL min all-reduces."""
    import subprocess
    world = min(torch.cuda.device_count(), 4)
    exe = os.path.join(ROOT, "navtech-radar-slam_b200", "cpp", "sharded_check")
    args = [str(world), str(tmp_path)] + (["nccl"] if mode == "nccl" else [])
    procs = [subprocess.Popen([exe, str(r)] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{o}"
        assert "sharded manager ok" in o and (f"({mode} exchange" in o), o
