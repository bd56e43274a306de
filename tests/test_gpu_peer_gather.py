"""The fused ordered gather (K1c storing straight into the owner rank's buffer through CUDA IPC)
with TWO processes.  Both ranks use cuda:0 here (the round-end GPU tier has one GPU), so NCCL
cannot be used; the count exchange and barriers go over gloo, the ids over the IPC mapping.  On a
multi-GPU box the same code runs with one GPU per rank over NVLink (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import support

pytestmark = pytest.mark.gpu

TOTAL = 3_000_001
QUERIES = ["SELECT command_id FROM Commands WHERE (command_id < 1700000) AND (sudo_used = FALSE OR risk_level > 3)",
           "SELECT command_id FROM Commands WHERE (risk_level > 4)",
           "SELECT command_id FROM Commands WHERE (command_id > 2999990)",      # only the last shard matches
           "SELECT command_id FROM Commands WHERE (risk_level > 100)"]          # nobody matches
COLS = ["command_id", "sudo_used", "risk_level"]


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
    eng = pkg.Engine.from_synth(TOTAL, n_rows=n, row_base=start, columns=COLS)
    pg = sharding.PeerGather(pkg, segment_capacity=n + 1)
    out = []
    host = np.zeros(TOTAL, dtype=np.uint32)
    for k, q in enumerate(QUERIES):
        if k % 2 == 0:
            total, counts, st = pg.run(eng, q, torch.device("cpu"))
            ids = pg.result(total).copy() if rank == 0 else None
        else:
            total, counts, st = pg.run(eng, q, torch.device("cpu"), pack="host", host_out=host)
            ids = host[:total].copy()
        if rank == 0:
            out.append((total, counts, ids))
        # no barrier needed: segments are double-buffered (see PeerGather)
    pg.close()
    eng.close()
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_gather_equals_single_engine(world):
    pkg = support.load_pkg()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    results = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    whole = pkg.Engine.from_synth(TOTAL, columns=COLS)
    for q, (total, counts, ids) in zip(QUERIES, results):
        want, _ = whole.select_ids(q, force_scan=True)
        assert total == len(want) and len(counts) == world and sum(counts) == total
        assert np.array_equal(ids, want), q
    whole.close()


# ---- sharded SELECT with the reference's path rule (index path merged across shards) ------------
SHARDED_QUERIES = [
    'SELECT command_id FROM Commands WHERE risk_level > 3 AND shell_type != "bash"',
    'SELECT command_id FROM Commands WHERE user_id = 1001 OR (exit_code = 127)',
    'SELECT command_id FROM Commands WHERE exit_code > 126 AND risk_level < 2',
    'SELECT command_id FROM Commands WHERE command_id >= 400000 AND command_id <= 700000',
    'SELECT command_id FROM Commands WHERE (risk_level > 3) AND (shell_type = "zsh")',   # groups only: scan path
    'SELECT command_id FROM Commands WHERE user_id = 999999',
]
PROBES = [("user_id", [1001, 2450, 999999, 1500], [1001, 2452, 999999, 1499]),
          ("command_id", [0, 500000, 999990, 2000000], [9, 500000, 1000010, 2000005])]
SH_TOTAL = 1_000_003
SH_COLS = ["command_id", "sudo_used", "risk_level", "exit_code", "user_id", "shell_type"]
SH_IDX = (("command_id", 0), ("user_id", 1), ("risk_level", 1), ("exit_code", 1), ("sudo_used", 3))


def _sharded_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["QPE_GPU_DEVICE"] = "0"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = support.load_pkg()
    from importlib import import_module
    sharding = import_module("pqps_b200.sharding")
    start, n = sharding.shard_range(SH_TOTAL, world, rank)
    eng = pkg.Engine.from_synth(SH_TOTAL, n_rows=n, row_base=start, columns=SH_COLS, indexes=SH_IDX)
    out = []
    for q in SHARDED_QUERIES:
        ids = sharding.sharded_select(eng, q)
        if rank == 0:
            out.append(ids)
    # batched findRange over the sharded indexes: counts on every rank, row ids on rank 0
    for attr, lo, hi in PROBES:
        total, rows = sharding.sharded_probe(eng, attr, np.asarray(lo), np.asarray(hi), rows=True)
        if rank == 0:
            out.append((total, rows))
    eng.close()
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_select_equals_single_engine(world):
    pkg = support.load_pkg()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    results = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    whole = pkg.Engine.from_synth(SH_TOTAL, columns=SH_COLS, indexes=SH_IDX)
    for q, ids in zip(SHARDED_QUERIES, results):
        want, _ = whole.select_ids(q)
        assert np.array_equal(ids, want), q
    for (attr, lo, hi), (total, rows) in zip(PROBES, results[len(SHARDED_QUERIES):]):
        dt = np.uint64 if attr == "command_id" else np.int32
        first, count, _ = whole.probe_batch(attr, np.asarray(lo, dtype=dt), np.asarray(hi, dtype=dt))
        assert np.array_equal(total, count.astype(np.int64)), attr
        for k in range(len(lo)):
            assert np.array_equal(rows[k], whole.index_slice(attr, int(first[k]), int(count[k]))), (attr, k)
    whole.close()
