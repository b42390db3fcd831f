"""Sharded Scan-Context query over NCCL (run under torchrun): parity against the CPU oracle + latency.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sc_multi_gpu_check.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402  (checker only)
from direct_stereo_slam_b200 import api, synthetic as syn  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s = api.Session(local)
n = int(os.environ.get("SC_ROWS", "20000"))
sig, key = syn.make_sc_database(n, 2024)
sig[n - 5] = sig[3]  # duplicates on different shards: lowest id must win across ranks
qs, qk, truth = syn.make_sc_queries(sig, key, 40, 77)
qs[0] = sig[3]
rows = api.shard_rows(n, world, rank)
db = api.ScanContextDB(s, len(rows) + 8)
db.add(key[rows], sig[rows], global_ids=rows)
ident = [api.ScanContextDB.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ident, src=0)
db.attach_comm(ident[0], world, rank)
mode = db.exchange_mode()
idx, diff = db.query(qs)
for _ in range(6):  # the mailbox slots come round: repeated collective queries must keep giving the same answer
    idx2, diff2 = db.query(qs)
    assert np.array_equal(idx2, idx) and np.array_equal(diff2.view(np.uint32), diff.view(np.uint32)), "repeated sharded query differs"
i1, d1 = db.query(qs[:1])
assert i1[0] == idx[0] and d1[0] == diff[0]
cand, cdist = db.search_ringkey(qk, k=3, thres=0.1)
ok = True
if rank == 0:
    o = orc.Oracle()
    for q in range(len(qs)):
        i_o, d_o = o.search_sc_dense(qs[q], sig)
        if idx[q] != i_o or diff[q] != np.float32(d_o):
            ok = False
            print("MISMATCH query", q, idx[q], i_o, diff[q], d_o)
        c_o, dd_o = o.search_ringkey(qk[q], key, k=3, thres=0.1)
        got = cand[q][cand[q] >= 0]
        if not np.array_equal(got, c_o):
            ok = False
            print("MISMATCH ringkey", q, got, c_o)
    assert idx[0] == 3
lat = []
for _ in range(30):
    dist.barrier()
    t0 = time.perf_counter()
    db.query(qs[:32])
    lat.append((time.perf_counter() - t0) * 1e3)
t = torch.tensor([float(np.median(lat)), db.last_scan_ms()], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("world %d rows %d exchange %s: parity %s, query(32) latency %.3f ms, scan kernel %.3f ms" % (world, n, mode, "OK" if ok else "FAILED", t[0].item(), t[1].item()))
dist.barrier()
db.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
