#!/usr/bin/env python3
"""BASELINE configs[4] in full: the 1 B-row table row-range sharded over N GPUs, QN and QS at several
selectivities, timed like bench.py (CUDA events on the engine's stream around K steps, barrier on both sides,
max over ranks).  Launch with torchrun; rank 0 prints one JSON object."""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import support  # noqa: E402
from bench import QUERIES, shard_of  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
pkg = support.load_pkg()
from importlib import import_module  # noqa: E402
sharding = import_module("pqps_b200.sharding")

TOTAL = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
start, n_local = shard_of(TOTAL, world, rank)
cols = sorted(set(QUERIES["QN"][1]) | set(QUERIES["QS"][1]))
eng = pkg.Engine.from_synth(TOTAL, n_rows=n_local, row_base=start, columns=cols)
stream = torch.cuda.ExternalStream(eng.stream, device=dev)
cap = int(TOTAL * 0.11 / 1) + 4096   # the largest result (10 %) may sit in ONE shard (command_id < K)
cap = min(cap, n_local + 4096)
sg = sharding.ShardGroup(pkg, eng, segment_capacity=cap, host_capacity=0, counts_device=dev) if world > 1 else None
out = {"rows": TOTAL, "n_gpus": world, "steps": STEPS, "runs": []}
for qname in ("QN", "QS"):
    sql_t, _, bpr = QUERIES[qname]
    for sel in (0.0001, 0.01, 0.1):
        sql = sql_t.format(K=max(1, int(TOTAL * sel)))

        def step():
            if sg is not None:
                return sg.select(sql, stats=False)[0]
            return eng.select_ids_device(sql, force_scan=True, stats=False)[0]

        for _ in range(3):
            n = step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.set_timing(True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(STEPS):
            n = step()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / STEPS
        tot = eng.timing_totals()
        eng.set_timing(False)
        t = torch.tensor([e0.elapsed_time(e1) / STEPS, wall, tot["scan_ms"] / max(tot["calls"], 1)], dtype=torch.float64,
                         device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
        ms = max(t[0].item(), t[1].item())
        if rank == 0:
            out["runs"].append({"query": qname, "selectivity": sel, "matches": int(n), "bytes_per_row": bpr,
                                "ms_per_step": ms, "k1f_ms_max_rank": t[2].item(), "rows_per_s": TOTAL / ms * 1e3,
                                "scan_gbs_all_gpus": (TOTAL * bpr + 4 * n) / ms / 1e6,
                                "k1f_gbs_per_gpu": (shard_of(TOTAL, world, 0)[1] * bpr) / t[2].item() / 1e6})
if rank == 0:
    print(json.dumps(out, indent=1))
if sg is not None:
    sg.close()
eng.close()
if world > 1:
    dist.destroy_process_group()
