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


# ---- index path across shards: per-segment merge == one B+ tree over the whole table -----------
INDEX_QUERIES = [
    # (where, [(key column, lo, hi)] = the segments the reference generates, in order)
    ('risk_level > 3 AND shell_type != "bash"', [("risk_level", 4, 2**31 - 1)]),
    ('user_id = 1001 OR (exit_code = 127)', [("user_id", 1001, 1001)]),
    ('exit_code > 126 AND risk_level < 2', [("exit_code", 127, 2**31 - 1), ("risk_level", -2**31, 1)]),
    ('command_id >= 500 AND command_id <= 1500', [("command_id", 500, 2**63 - 1), ("command_id", 0, 1500)]),
    ('risk_level = 5 OR user_id = 1003', [("risk_level", 5, 5), ("user_id", 1003, 1003)]),
]


def _merge_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    support.load_pkg()
    from importlib import import_module
    sharding = import_module("pqps_b200.sharding")
    o = Oracle.from_csv(CSV_2K)
    total = o.num_rows
    start, n = sharding.shard_range(total, world, rank)
    pos = np.arange(start, start + n)
    out = []
    for where, segments in INDEX_QUERIES:
        passing = np.zeros(total, dtype=bool)
        passing[o.scan(where, first=start, n=n)] = True          # this shard's rows that pass the whole WHERE
        local = []
        for col, lo, hi in segments:
            key = np.array([int(o.cell(int(p), col)) for p in pos], dtype=np.int64)
            cand = (key >= lo) & (key <= hi)
            order = np.lexsort((-pos, key))                        # (key ASC, position DESC) within the shard
            order = order[cand[order] & passing[pos[order]]]
            local.append((key[order], pos[order].astype(np.uint32)))
        merged = sharding.merge_index_segments(local)
        if rank == 0:
            out.append(merged.tolist())
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_index_path_merge_across_shards(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_merge_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    results = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = Oracle.from_csv(CSV_2K)
    for (where, _), got in zip(INDEX_QUERIES, results):
        want, used = o.select_ids(where)
        assert used
        assert got == want.tolist(), where


# ---- sharded B+ probe batches: host-side merge, with a numpy stand-in for the per-shard GPU index ---------
class _StubLib:
    def __init__(self, base):
        self._base = base

    def qpe_gpu_row_base(self, _h):
        return self._base


class _StubShardEngine:
    """What sharding.sharded_probe needs from an engine, over one shard's keys: the flattened index of
    csrc/index.cu restated in numpy ((key ASC, LOCAL position DESC) order, lower/upper bound probes)."""

    def __init__(self, keys, base):
        pos = np.arange(len(keys))
        order = np.lexsort((-pos, keys))
        self.keys = np.asarray(keys, dtype=np.int64)[order]
        self.perm = pos[order].astype(np.uint32)
        self._lib = _StubLib(base)
        self._h = None

    def probe_batch(self, attribute, lo, hi):
        first = np.searchsorted(self.keys, np.asarray(lo, dtype=np.int64), side="left")
        last = np.searchsorted(self.keys, np.asarray(hi, dtype=np.int64), side="right")
        return first.astype(np.uint32), np.maximum(last - first, 0).astype(np.uint32), {}

    def index_slice(self, attribute, first, count):
        return self.perm[first:first + count]

    def index_slice_keys(self, attribute, first, count):
        return self.keys[first:first + count]


PROBE_SETS = [("user_id", [1001, 1002, 999999, 1010], [1001, 1005, 999999, 1009]),
              ("risk_level", [0, 3, 5, 9], [1, 3, 5, 9]),
              ("command_id", [0, 700, 1995, 5000], [9, 700, 2100, 6000])]


def _probe_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    support.load_pkg()
    from importlib import import_module
    sharding = import_module("pqps_b200.sharding")
    o = Oracle.from_csv(CSV_2K)
    start, n = sharding.shard_range(o.num_rows, world, rank)
    out = []
    for attr, lo, hi in PROBE_SETS:
        keys = np.array([int(o.cell(p, attr)) for p in range(start, start + n)], dtype=np.int64)
        eng = _StubShardEngine(keys, start)
        total, rows = sharding.sharded_probe(eng, attr, np.asarray(lo), np.asarray(hi), rows=True)
        if rank == 0:
            out.append((total.tolist(), [r.tolist() for r in rows]))
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_probe_merge_equals_one_tree(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_probe_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    results = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = Oracle.from_csv(CSV_2K)
    for (attr, lo, hi), (total, rows) in zip(PROBE_SETS, results):
        perm = o.index_order(attr)                                   # one B+ tree over the whole table: leaf-chain order
        keys = np.array([int(o.cell(int(p), attr)) for p in perm], dtype=np.int64)
        for k in range(len(lo)):
            want = perm[(keys >= lo[k]) & (keys <= hi[k])].tolist()  # findRange(lo, hi), bplus.c:282-314
            assert total[k] == len(want) and rows[k] == want, (attr, k)
