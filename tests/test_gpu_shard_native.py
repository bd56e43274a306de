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
TINY_SQL = "SELECT command_id FROM Commands WHERE (command_id > 2499900)"


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
    # both host-result paths: every rank stages its 1/world of the result and the copy engine takes it to the host
    # (2 ranks), and the delivery kernel storing into the mapped host buffer itself (3 ranks)
    grp.set_multipath(2 if world == 3 else 1)
    if world == 3:
        grp.balance_links(nbytes=1 << 20, reps=2)    # unequal shares of a host result (whatever the links measure)
    out = []
    # every query twice in a row (device result then host result), then the whole list again: epochs and
    # the two parities of every buffer get reused with different contents
    for rep in range(2):
        for q in QUERIES:
            total, counts, _ = grp.select(q, to_host=False)
            dev_ids = grp.device_result(total).copy() if rank == 0 else None
            total_h, counts_h, _ = grp.select(q, to_host=True)
            if rank == 0:
                out.append((total, counts, dev_ids, total_h, counts_h, grp.host_result(total_h).copy()))
    # two queries in flight: submit q + 1 before waiting for q, device and host results alternating
    piped = []
    plan = [(q, k % 2 == 1) for k, q in enumerate(QUERIES + QUERIES[::-1])]
    grp.submit(*plan[0])
    for k, (q, to_host) in enumerate(plan):
        if k + 1 < len(plan):
            grp.submit(*plan[k + 1])
        total, counts, _ = grp.wait()
        if rank == 0:
            ids = grp.host_result(total).copy() if to_host else grp.device_result(total).copy()
            piped.append((q, total, counts, ids))
    # deferred completion: `wait` returns once this rank's piece is delivered, the owner takes result k - 1 (back=1)
    # after query k + 1 has been submitted and k waited for -- three id arrays keep it valid that long
    grp.set_deferred(True)
    deferred = []
    dplan = QUERIES + QUERIES[::-1] + QUERIES[:3]
    grp.submit(dplan[0], True)
    for k, q in enumerate(dplan):
        if k + 1 < len(dplan):
            grp.submit(dplan[k + 1], True)
        total, counts, _ = grp.wait()
        if rank == 0:
            deferred.append([q, total, None])
            if k >= 1:
                deferred[k - 1][2] = grp.host_result(back=1).copy()
    if rank == 0:
        deferred[-1][2] = grp.host_result(back=0).copy()
    grp.set_deferred(False)
    # sharded DELETE (local compaction + renumbering through the comm blocks), then the same queries again:
    # global row ids must be positions in the table AFTER the delete
    deleted, left = grp.delete(DELETE_SQL)
    after = []
    for q in QUERIES:
        total, counts, _ = grp.select(q, to_host=False)
        ids = grp.device_result(total).copy() if rank == 0 else None
        after.append((total, ids))
    grp.close()
    # errors are collective: a segment too small for a rank's ids fails on EVERY rank (and nothing hangs);
    # the group keeps working afterwards
    small = sharding.ShardGroup(pkg, eng, segment_capacity=1000, host_capacity=4000)
    errors = []
    for to_host in (False, True):
        try:
            small.select(QUERIES[0], to_host=to_host)
            errors.append("no error")
        except pkg.QpeError as e:
            errors.append(str(e))
    tiny_total, _, _ = small.select(TINY_SQL, to_host=True)
    all_errors = [None] * world
    dist.all_gather_object(all_errors, (errors, tiny_total))
    small.close()
    eng.close()
    if rank == 0:
        ret.put((out, deleted, left, after, piped, all_errors, deferred))
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
    results, deleted, left, after, piped, all_errors, deferred = ret.get(timeout=600)
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
    for q, total, counts, ids in piped:
        want, _ = whole.select_ids(q, force_scan=True)
        assert total == len(want) and sum(counts) == total and np.array_equal(ids, want), "two in flight: " + q
    for q, total, ids in deferred:
        want, _ = whole.select_ids(q, force_scan=True)
        assert total == len(want) and ids is not None and np.array_equal(ids, want), "deferred: " + q
    # the same DELETE on the whole table, then the same queries
    out = whole.run(DELETE_SQL, 5)
    assert f"Rows affected: {deleted}" in out and whole.num_rows == left == TOTAL - deleted
    tiny_want, _ = whole.select_ids(TINY_SQL, force_scan=True)
    for errors, tiny_total in all_errors:   # every rank saw the same failures, and the same success afterwards
        assert len(errors) == 2 and all("rc=-5" in e for e in errors), errors
        assert tiny_total == len(tiny_want) > 0
    for q, (total, ids) in zip(QUERIES, after):
        want, _ = whole.select_ids(q, force_scan=True)
        assert total == len(want) and np.array_equal(ids, want), "after DELETE: " + q
    whole.close()
