"""Sharded SELECT at 10 M rows, two ranks (csrc/shard.cu + sharding.py): a query whose matches are spread evenly over
the shards (every rank delivers its own ids, slices straddle the shard boundary) through the device and the host result,
two queries in flight; and index-path queries whose equal keys live in BOTH shards, whose merged order must be one B+
tree's (key ASCENDING, global position DESCENDING: engine/bplus.c:282-358).  Checked against ONE engine
over the whole table and against the oracle.  Two processes share cuda:0 (the GPU test tier has one GPU)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import support

pytestmark = pytest.mark.gpu

TOTAL = 10_000_019
COLS = ["command_id", "user_id", "sudo_used", "risk_level", "exit_code"]
IDX = (("command_id", 0), ("user_id", 1), ("risk_level", 1))
UNIFORM = ["SELECT command_id FROM Commands WHERE (user_id < 1100) AND (sudo_used = FALSE OR risk_level > 3)",
           "SELECT command_id FROM Commands WHERE (exit_code != 0) AND (risk_level >= 2)",
           "SELECT command_id FROM Commands WHERE (risk_level = 5)"]
INDEXED = ["SELECT command_id FROM Commands WHERE user_id = 1001",                       # one key, rows in both shards
           "SELECT command_id FROM Commands WHERE risk_level > 4 AND exit_code = 0",     # few keys, heavy duplicates
           "SELECT command_id FROM Commands WHERE user_id <= 1003 OR (exit_code = 127)",
           "SELECT command_id FROM Commands WHERE command_id >= 4999990 AND command_id < 5000030"]  # straddles the cut


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["QPE_GPU_DEVICE"] = "0"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = support.load_pkg()
    from importlib import import_module
    sharding = import_module("pqps_b200.sharding")
    start, n = sharding.shard_range(TOTAL, world, rank)
    eng = pkg.Engine.from_synth(TOTAL, n_rows=n, row_base=start, columns=COLS, indexes=IDX)
    grp = sharding.ShardGroup(pkg, eng, segment_capacity=n + 1, host_capacity=TOTAL)
    grp.balance_links(nbytes=1 << 20, reps=2)
    scans = []
    plan = [(q, h) for q in UNIFORM for h in (False, True)]
    grp.submit(*plan[0])
    for k, (q, to_host) in enumerate(plan):
        if k + 1 < len(plan):
            grp.submit(*plan[k + 1])
        total, counts, _ = grp.wait()
        if rank == 0:
            ids = grp.host_result(total).copy() if to_host else grp.device_result(total).copy()
            scans.append((q, to_host, counts, ids))
    grp.close()
    indexed = []
    for q in INDEXED:
        ids = sharding.sharded_select(eng, q)
        if rank == 0:
            indexed.append((q, ids))
    eng.close()
    if rank == 0:
        ret.put((scans, indexed))
    dist.barrier()
    dist.destroy_process_group()


def test_uniform_matches_and_cross_shard_duplicates_at_10m_rows():
    pkg = support.load_pkg()
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    scans, indexed = ret.get(timeout=900)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    whole = pkg.Engine.from_synth(TOTAL, columns=COLS, indexes=IDX)
    oracle = support.Oracle.from_columns({c: whole.fetch_column(c) for c in COLS})
    for q, to_host, counts, ids in scans:
        want = oracle.scan(q.split("WHERE", 1)[1])
        assert sum(counts) == len(want) and min(counts) > 0, (q, counts)     # every shard contributes
        assert np.array_equal(ids, want), (q, "host" if to_host else "device")
    for q, ids in indexed:
        want, used = oracle.select_ids(q.split("WHERE", 1)[1], IDX)
        assert used, q
        assert ids.shape == want.shape and np.array_equal(ids, want), q
        got_one, st = whole.select_ids(q)
        assert st["path"] == 1 and np.array_equal(got_one, want), q
    whole.close()
