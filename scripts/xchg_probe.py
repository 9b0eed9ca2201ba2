"""Developer probe (torchrun, N >= 2): where does the time of a sharded query's tail go?
torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/xchg_probe.py"""
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
nrs = importlib.import_module("navtech-radar-slam_b200")
synth = importlib.import_module("navtech-radar-slam_b200.synth")
sc = nrs.scancontext
ctx = nrs.Context(lr)
db, rev = synth.sc_database(10000, 500, seed=42)
mine = np.arange(rank, 10000, world)
NREP = int(os.environ.get("PROBE_REPS", "1"))
ds = []
for _ in range(NREP):
    d = sc.SCDatabase(ctx, capacity_hint=len(mine))
    d.append(db[mine])
    ds.append(d)
assert sc.enable_peer_exchange(ds, rank, world, dev)
d = ds[0]
q = torch.from_numpy(np.stack([db[j] for j, _, _ in rev[:64]])).to(dev)
out = torch.zeros(16, dtype=torch.uint8, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
rows = []
for k in range(40):
    dist.barrier()
    torch.cuda.synchronize()
    d.query_dev_sharded(q[k % 64].data_ptr(), 0, len(mine), out.data_ptr())
    t = d.tail_times()
    rows.append([t[1] - t[0], t[2] - t[1], t[3] - t[2]])
r = np.array(rows[5:], dtype=np.float64) / 1e3
print(f"rank {rank} single (synced) us: cert {r[:,0].mean():.2f}  write+fence {r[:,1].mean():.2f}  wait {r[:,2].mean():.2f} (max {r[:,2].max():.2f})", flush=True)
# back-to-back
for n in (200, 2000):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    th = time.perf_counter()
    for k in range(n):
        ds[k % NREP].query_dev_sharded(q[k % 64].data_ptr(), 0, len(mine), out.data_ptr())
    th = time.perf_counter() - th
    with torch.cuda.stream(stream):
        e1.record()
    stream.synchronize()
    t = d.tail_times()
    print(f"rank {rank} back-to-back x{n}: {1e3 * e0.elapsed_time(e1) / n:.2f} us/step (host enqueue {1e6 * th / n:.2f} us/step); "
          f"last: cert {(t[1]-t[0])/1e3:.2f} write {(t[2]-t[1])/1e3:.2f} wait {(t[3]-t[2])/1e3:.2f}", flush=True)
    # same without exchange
    with torch.cuda.stream(stream):
        e0.record()
    for k in range(n):
        ds[k % NREP].query_dev(q[k % 64].data_ptr(), 0, len(mine), 0, out.data_ptr())
    with torch.cuda.stream(stream):
        e1.record()
    stream.synchronize()
    print(f"rank {rank} no exchange x{n}: {1e3 * e0.elapsed_time(e1) / n:.2f} us/step", flush=True)
dist.barrier()
torch.cuda.synchronize()
dist.destroy_process_group()
for d in ds:
    d.close()
ctx.close()
