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
