"""world_size-2 and -3 CPU tests (gloo) of the multi-GPU host logic: block partition + count
exchange + ordered gather.  Each rank plays a GPU: it filters ITS row range with the oracle and the
gathered list must equal the single-engine result over the whole table."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import support
from support import CSV_2K, Oracle

pytestmark = pytest.mark.skipif(not Oracle.available(), reason="oracle not built")

WHERES = ['(command_id < 1500) AND (sudo_used = FALSE OR risk_level > 3)',
          'shell_type = "zsh" AND host_name = "labpc-01" OR base_command = "ls"',
          'command_id > 100000',           # nobody matches
          'command_id >= 1990']            # only the last shard matches


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = support.load_pkg()
    from importlib import import_module
    sharding = import_module("pqps_b200.sharding")
    o = Oracle.from_csv(CSV_2K)
    total = o.num_rows
    start, n = sharding.shard_range(total, world, rank)
    results = []
    out = None
    for w in WHERES:
        local = o.scan(w, first=start, n=n).astype(np.int64)  # the oracle already yields global positions
        tot, counts, out = sharding.ordered_gather(torch.from_numpy(local), out)
        if rank == 0:
            results.append((tot, counts, out[:tot].clone().numpy().tolist()))
    if rank == 0:
        ret.put(results)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ordered_gather_equals_single_engine(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    results = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = Oracle.from_csv(CSV_2K)
    for w, (tot, counts, ids) in zip(WHERES, results):
        want = o.scan(w).tolist()
        assert ids == want, w
        assert tot == len(want) and sum(counts) == tot and len(counts) == world


def test_shard_range_is_the_mpi_block_partition():
    support.load_pkg()
    from importlib import import_module
    sharding = import_module("pqps_b200.sharding")
    for total in (0, 1, 7, 2000, 10**9, 10**9 + 5):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                start, n = sharding.shard_range(total, world, r)
                assert start == covered
                base, rem = divmod(total, world)
                assert n == base + (1 if r < rem else 0)   # executeEngine-mpi.c:703-715
                covered += n
            assert covered == total
