"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: keyframe sharding, the one
all_gather per query, the exact lowest-index merge.  The per-shard search is a test double backed
by the oracle (the product's local search is the CUDA library; it is injected here because this
container has no GPU)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nrs = importlib.import_module("navtech-radar-slam_b200")
    synth = importlib.import_module("navtech-radar-slam_b200.synth")
    from oracle import pyoracle as O
    sc = nrs.scancontext
    db, rev = synth.sc_database(301, 40, seed=17)
    db[250] = db[7]
    db[260] = db[7]          # exact ties across shards: lowest index must win on every rank

    holder = {}

    def local_search(q, lo, hi):
        store = holder["db"].local_descs()
        if hi <= lo:
            return (1e7, 0, 0)
        _, _, best = O.sc_search(q, store, lo, hi, want_all=False)
        return best

    sdb = sc.ShardedSCDatabase(rank, world, local_search=local_search)
    holder["db"] = sdb
    for a in range(0, 301, 13):                      # appended in uneven batches, like a growing map
        sdb.append(db[a:a + 13])
    assert len(sdb) == 301
    results = []
    queries = [(db[rev[0][0]], 0, rev[0][0] - 30), (db[7], 8, 301), (db[7], 0, 301), (db[100], 0, 0),
               (db[rev[5][0]], 3, 299), (np.zeros(1200), 0, 301), (db[300], 0, 300)]
    for q, a, b in queries:
        got = sdb.query(q, a, b)
        _, _, exp = O.sc_search(q, db, a, b, want_all=False)
        results.append((got, exp))
    # batched: one collective for the whole batch, per-query merge
    qs = np.stack([db[rev[k][0]] for k in range(9)] + [db[7], np.zeros(1200)])
    for a, b in ((0, 290), (8, 301), (0, 0)):
        got = sdb.query_batch(qs, a, b)
        for k in range(qs.shape[0]):
            _, _, exp = O.sc_search(qs[k], db, a, b, want_all=False)
            results.append((got[k], exp))
    out_q.put((rank, results))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_query_matches_single_scan(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, results in got:
        for res, exp in results:
            assert res == exp, (rank, res, exp)
    # identical on every rank
    assert all(g[1] == got[0][1] for g in got)


def test_shard_index_arithmetic():
    sc = importlib.import_module("navtech-radar-slam_b200.scancontext")
    for world in (1, 2, 3, 8):
        for first, last in ((0, 0), (0, 1), (5, 17), (3, 100), (99, 100)):
            seen = []
            for r in range(world):
                lo, hi = sc.local_range(first, last, r, world)
                seen += [l * world + r for l in range(lo, hi)]
            assert sorted(seen) == list(range(first, last))
    assert sc.best_of([(1e7, 0, 0), (0.3, 9, 2), (0.3, 4, 5), (0.5, 1, 1)]) == (0.3, 4, 5)
    assert sc.best_of([(1e7, 0, 0), (float("nan"), 3, 1)]) == (1e7, 0, 0)
