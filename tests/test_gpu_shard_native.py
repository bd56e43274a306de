"""The native sharded SELECT (csrc/shard.cu): ids stored into the owner's buffer by the scan kernel over
peer memory, counts exchanged by a kernel's own stores, segments packed in partition order -- no host
collective per query.  TWO/THREE processes share cuda:0 here (the round-end GPU tier has one GPU): the
waiting kernels of one process are time-sliced against the scans of the others, so this is slow but
exercises exactly the code that runs one-rank-per-GPU over NVLink (bench.py --gpus N).  gloo is only used
to hand the IPC handles around and to return the results to the test."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import support

pytestmark = pytest.mark.gpu

TOTAL = 2_500_003
QUERIES = ["SELECT command_id FROM Commands WHERE (command_id < 1700000) AND (sudo_used = FALSE OR risk_level > 3)",
           "SELECT command_id FROM Commands WHERE (risk_level > 4)",
           "SELECT command_id FROM Commands WHERE (command_id > 2499990)",      # only the last shard matches
           "SELECT command_id FROM Commands WHERE (risk_level > 100)",          # nobody matches
           "SELECT command_id FROM Commands WHERE (command_id < 5)",            # only the first shard matches
           "SELECT command_id FROM Commands WHERE (sudo_used = TRUE) OR (risk_level < 2)"]
COLS = ["command_id", "sudo_used", "risk_level"]
DELETE_SQL = "DELETE FROM Commands WHERE (risk_level = 2) OR (command_id < 1000)"


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
    grp = sharding.ShardGroup(pkg, eng, segment_capacity=n + 1, host_capacity=TOTAL)
    # both host-result paths: the first shard streaming during its scan (2 ranks) and every rank copying its
    # 1/world of the packed result out of the owner's memory (3 ranks; automatic only from 8 ranks up)
    grp.set_multipath(1 if world == 3 else 0)
    out = []
    # every query twice in a row (device result then host result), then the whole list again: epochs and
    # the two parities of every buffer get reused with different contents
    for rep in range(2):
        for q in QUERIES:
            total, counts, _ = grp.select(q, to_host=False)
            dev_ids = grp.device_result(total).copy() if rank == 0 else None
            total_h, counts_h, _ = grp.select(q, to_host=True)
            if rank == 0:
                out.append((total, counts, dev_ids, total_h, counts_h, grp.host_ids[:total_h].copy()))
    # sharded DELETE (local compaction + renumbering through the comm blocks), then the same queries again:
    # global row ids must be positions in the table AFTER the delete
    deleted, left = grp.delete(DELETE_SQL)
    after = []
    for q in QUERIES:
        total, counts, _ = grp.select(q, to_host=False)
        ids = grp.device_result(total).copy() if rank == 0 else None
        after.append((total, ids))
    grp.close()
    eng.close()
    if rank == 0:
        ret.put((out, deleted, left, after))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_native_shard_select_equals_single_engine(world):
    pkg = support.load_pkg()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    results, deleted, left, after = ret.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    whole = pkg.Engine.from_synth(TOTAL, columns=COLS)
    for k, (total, counts, dev_ids, total_h, counts_h, host_ids) in enumerate(results):
        q = QUERIES[k % len(QUERIES)]
        want, _ = whole.select_ids(q, force_scan=True)
        assert total == len(want) == total_h, q
        assert len(counts) == world and sum(counts) == total and counts == counts_h
        assert np.array_equal(dev_ids, want), q
        assert np.array_equal(host_ids, want), q
    # the same DELETE on the whole table, then the same queries
    out = whole.run(DELETE_SQL, 5)
    assert f"Rows affected: {deleted}" in out and whole.num_rows == left == TOTAL - deleted
    for q, (total, ids) in zip(QUERIES, after):
        want, _ = whole.select_ids(q, force_scan=True)
        assert total == len(want) and np.array_equal(ids, want), "after DELETE: " + q
    whole.close()
