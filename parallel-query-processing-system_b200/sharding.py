"""Row-range sharding of the table over the GPUs of one box and the ordered result gather.

Mirrors the reference's MPI mode: contiguous row ranges with the block partition of
engine/mpi/executeEngine-mpi.c:703-715 (base = N / size, the first N % size ranks get one extra
row), per-rank counts exchanged (there: MPI_Allreduce + MPI_Allgather, :745-754; here one
all-gather over NCCL/NVLink) and the per-rank pieces concatenated in partition order (there:
MPI_Allgatherv of delete flags, :765-766; here the matching row ids, sent to rank 0).

One process per GPU; `torch.distributed` is only the plumbing (NCCL on GPUs, gloo in the CPU
tests).  On the full-scan path concatenation in partition order IS table order, so the gathered
list is bit-identical to what one engine over the whole table returns.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """(first global row, row count) of `rank`'s contiguous shard."""
    base, rem = divmod(int(total_rows), int(world))
    n = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, n


_count_bufs = {}


def exchange_counts(local_count: int, device, group=None) -> List[int]:
    """every rank learns every rank's match count (one 8-byte all-gather; NCCL over NVLink on GPUs)"""
    world = dist.get_world_size(group)
    device = torch.device(device)
    if device.type == "cuda":
        key = (device, world, id(group))
        if key not in _count_bufs:
            _count_bufs[key] = (torch.zeros(1, dtype=torch.int64, device=device),
                                torch.zeros(world, dtype=torch.int64, device=device))
        mine, every = _count_bufs[key]
        mine.fill_(int(local_count))
        dist.all_gather_into_tensor(every, mine, group=group)
        return every.tolist()
    mine = torch.tensor([int(local_count)], dtype=torch.int64)
    parts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return [int(p.item()) for p in parts]


def ordered_gather(local_ids: torch.Tensor, out: Optional[torch.Tensor] = None, dst: int = 0,
                   group=None) -> Tuple[int, List[int], Optional[torch.Tensor]]:
    """Concatenate every rank's (already global, already ordered) row ids on rank `dst`, in
    partition order.  Returns (total, per-rank counts, gathered tensor on dst / None elsewhere).
    `out` (on dst) is reused when it is large enough."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    counts = exchange_counts(local_ids.numel(), local_ids.device, group)
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    total = offs[-1]
    if rank == dst:
        if out is None or out.numel() < total:
            out = torch.empty(max(total, 1), dtype=local_ids.dtype, device=local_ids.device)
        ops = []
        for r in range(world):
            if counts[r] == 0:
                continue
            piece = out[offs[r]:offs[r + 1]]
            if r == dst:
                piece.copy_(local_ids)
            else:
                ops.append(dist.P2POp(dist.irecv, piece, r, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return total, counts, out
    if counts[rank]:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local_ids.contiguous(), dst, group)]):
            w.wait()
    return total, counts, None


def merge_index_segments(local_segments, device="cpu", dst: int = 0, group=None):
    """Cross-shard merge of an index-path SELECT.  `local_segments` = this rank's
    [(order keys int64, global ids), ...] per segment, each (key ASC, local position DESC); u64 keys arrive with
    their top bit flipped (Engine.select_segments), so the signed sort below is the tree's unsigned order.  For every
    segment the shards are concatenated from the HIGHEST rank to the lowest and sorted stably by
    key, which yields (key ASC, global position DESC): exactly the leaf-chain order of one B+ tree
    over the whole table (engine/bplus.c:282-358).  Segments are then concatenated in
    WHERE order, duplicates kept.  Returns the merged global ids on rank `dst` (numpy), None elsewhere."""
    import numpy as np
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    device = torch.device(device)
    merged = []
    for keys, ids in local_segments:
        k = torch.from_numpy(np.ascontiguousarray(keys, dtype=np.int64)).to(device)
        i = torch.from_numpy(np.ascontiguousarray(ids).astype(np.int64)).to(device)
        tot_k, counts, all_k = ordered_gather(k, None, dst, group)
        _, _, all_i = ordered_gather(i, None, dst, group)
        if rank == dst:
            offs = [0]
            for c in counts:
                offs.append(offs[-1] + c)
            # highest rank first: among equal keys rows of later shards (larger global positions) come first
            order = [r for r in range(world - 1, -1, -1)]
            kk = torch.cat([all_k[offs[r]:offs[r + 1]] for r in order]) if tot_k else all_k[:0]
            ii = torch.cat([all_i[offs[r]:offs[r + 1]] for r in order]) if tot_k else all_i[:0]
            perm = torch.sort(kk, stable=True).indices
            merged.append(ii[perm].cpu().numpy().astype(np.uint32))
    if rank != dst:
        return None
    return np.concatenate(merged) if merged else np.zeros(0, dtype=np.uint32)


def sharded_select(engine, statement: str, device="cpu", dst: int = 0, group=None):
    """SELECT on a row-range sharded table with the reference's path rule and result order:
    index path (per-segment merge) when a top-level condition names a u64/int index, else the
    full-scan path (ordered gather).  Returns the global row ids on rank `dst` (numpy), None elsewhere."""
    import numpy as np
    used, segs = engine.select_segments(statement, global_ids=True)
    if used:
        return merge_index_segments(segs, device, dst, group)
    cnt, dptr, _ = engine.select_ids_device(statement, force_scan=True, global_ids=True)
    local = engine.copy_from_device(dptr, cnt).astype(np.int64)
    total, _, out = ordered_gather(torch.from_numpy(local).to(torch.device(device)), None, dst, group)
    if dist.get_rank(group) != dst:
        return None
    return out[:total].cpu().numpy().astype(np.uint32)


def sharded_probe(engine, attribute: str, lo, hi, device="cpu", rows: bool = False, dst: int = 0, group=None):
    """Batched findRange on a row-range sharded table: the query batch is replicated, every rank
    probes the index of ITS shard (K3), the per-query counts are summed over the ranks, and -- with rows=True --
    the answers are concatenated per query from the HIGHEST rank to the lowest and sorted stably by key, which
    is (key ASC, global position DESC): the leaf-chain order of one B+ tree over the whole table
    (engine/bplus.c:282-314).
    Returns (total counts per query on every rank, list of global-row-id arrays on rank `dst` / None)."""
    import numpy as np
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    first, count, _ = engine.probe_batch(attribute, lo, hi)
    device = torch.device(device)
    total = torch.from_numpy(count.astype(np.int64)).to(device)
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    total = total.cpu().numpy()
    if not rows:
        return total, None
    # this shard's answers, flattened query by query, with global row ids
    base = int(engine._lib.qpe_gpu_row_base(engine._h))
    local = [engine.index_slice(attribute, int(f), int(c)).astype(np.int64) + base for f, c in zip(first, count)]
    flat = np.concatenate(local) if local else np.zeros(0, dtype=np.int64)
    keys = [engine.index_slice_keys(attribute, int(f), int(c)) for f, c in zip(first, count)]
    flat_keys = np.concatenate(keys) if keys else np.zeros(0, dtype=np.int64)
    if attribute == "command_id":   # u64 keys travel as bits: order them as unsigned
        flat_keys = (flat_keys.view(np.uint64) ^ np.uint64(1 << 63)).view(np.int64)
    _, flat_counts, all_flat = ordered_gather(torch.from_numpy(flat).to(device), None, dst, group)
    _, _, all_keys = ordered_gather(torch.from_numpy(flat_keys).to(device), None, dst, group)
    _, _, all_cnt = ordered_gather(torch.from_numpy(count.astype(np.int64)).to(device), None, dst, group)
    if rank != dst:
        return total, None
    q = len(count)
    all_flat = all_flat.cpu().numpy()
    all_keys = all_keys.cpu().numpy()
    per_rank_cnt = all_cnt.cpu().numpy()[:q * world].reshape(world, q)
    starts = np.zeros(world, dtype=np.int64)
    starts[1:] = np.cumsum(flat_counts)[:-1]
    offs = [np.concatenate(([0], np.cumsum(per_rank_cnt[r]))) for r in range(world)]
    out = []
    for k in range(q):
        sl = [slice(starts[r] + offs[r][k], starts[r] + offs[r][k + 1]) for r in range(world - 1, -1, -1)]
        ids_k = np.concatenate([all_flat[x] for x in sl])
        keys_k = np.concatenate([all_keys[x] for x in sl])
        out.append(ids_k[np.argsort(keys_k, kind="stable")].astype(np.uint32))
    return total, out


class PeerGather:
    """Ordered gather of a sharded full scan WITHOUT a transfer step of its own: rank `dst` owns the
    result buffer, every other rank maps it through CUDA IPC, and each rank's compaction kernel
    (K1c) stores its (global) row ids straight into its segment of that buffer -- over NVLink for
    the non-owners.  One stream synchronisation and ONE collective per query (the 8-byte count
    all-gather, which also orders the stores before the owner reads); no send/recv of ids.
    The owner then packs the segments (device-to-device, or straight into host memory)."""

    def __init__(self, pkg, segment_capacity: int, counts_device="cpu", dst: int = 0, group=None):
        self.pkg = pkg
        self.dst = dst
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        # every rank must lay the segments out identically: agree on the largest request
        self.seg_cap = max(exchange_counts(int(max(segment_capacity, 1)), counts_device, group))
        self.parity = 0
        self.buffer = None   # owner: 2 x world segments of seg_cap ids
        self.dense = None    # owner: packed result
        handle = [None]
        if self.rank == dst:
            # two sets of segments, used alternately: a rank can run at most one query ahead of the
            # owner (the next count all-gather needs the owner), so the set the owner is still
            # packing is never the one being written
            self.buffer = pkg.DeviceBuffer(4 * self.seg_cap * self.world * 2)
            self.dense = pkg.DeviceBuffer(4 * self.seg_cap * self.world)
            self.ptr = self.buffer.ptr
            handle[0] = self.buffer.export_handle()
        dist.broadcast_object_list(handle, src=dst, group=group)
        if self.rank != dst:
            self.ptr = pkg.ipc_open(handle[0])
        dist.barrier(group=group)

    def run(self, engine, statement: str, counts_device, pack: str = "device", host_out=None):
        """One sharded SELECT.  pack = "device": the owner packs the segments into self.dense;
        "host": the owner copies them into host_out (numpy uint32, ideally pinned); "none": leave them.
        Returns (total matches, per-rank counts, this rank's stats)."""
        set_off = 4 * self.seg_cap * self.world * self.parity
        self.parity ^= 1
        seg = self.ptr + set_off + 4 * self.seg_cap * self.rank
        cnt, st, _ = engine.select_ids_to(statement, seg, self.seg_cap, global_ids=True)
        counts = exchange_counts(cnt, counts_device, self.group)
        if max(counts) > self.seg_cap:
            raise RuntimeError(f"segment too small: {max(counts)} ids > capacity {self.seg_cap}")
        total = sum(counts)
        if self.rank == self.dst and pack != "none":
            lib = self.pkg.load_library()
            off = 0
            for r, c in enumerate(counts):
                if c:
                    src = self.buffer.ptr + set_off + 4 * self.seg_cap * r
                    if pack == "host":
                        lib.qpe_gpu_copy_to_host(host_out.ctypes.data + 4 * off, src, 4 * c)
                    else:
                        lib.qpe_gpu_copy_device(self.dense.ptr + 4 * off, src, 4 * c)
                off += c
        return total, counts, st

    def result(self, total: int):
        """owner: the packed result on the host (for checks)"""
        return self.dense.to_host(total)

    def close(self):
        dist.barrier(group=self.group)
        if self.rank != self.dst:
            self.pkg.ipc_close(self.ptr)
        dist.barrier(group=self.group)
        if self.buffer is not None:
            self.buffer.free()
            self.dense.free()
            self.buffer = None


class ShardGroup:
    """A row-range sharded table driven natively (csrc/shard.cu): per query NO host collective and no NCCL
    call.  Device result: every rank's scan kernel stores its global row ids into the owner's result buffer over
    NVLink peer memory, a kernel stores (epoch, count) into every rank's comm block, the owner packs the segments
    in partition order.  Host result: the ids stay in the rank's own segment, and every rank takes ITS 1/world of
    the result (read over NVLink from whichever segments it spans) to the shared host buffer over its own PCIe link.
    Up to two queries can be in flight (`submit` / `wait`).  `torch.distributed` is used once, here, to pass the
    IPC handles around."""

    def __init__(self, pkg, engine, segment_capacity: int, host_capacity: int = 0, owner: int = 0, group=None,
                 counts_device="cpu"):
        import ctypes as C
        import os
        self.pkg, self.engine, self.owner, self.group = pkg, engine, owner, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.lib = pkg.load_library()
        lib, h = self.lib, engine._h
        # every rank must lay its segments out identically: agree on the largest request
        self.seg_cap = max(exchange_counts(int(max(segment_capacity, 1)), counts_device, group))
        buf = C.create_string_buffer(64)
        engine._check(lib.qpe_shard_init(h, self.rank, self.world, self.seg_cap, buf), "qpe_shard_init")
        handles = [None] * self.world
        dist.all_gather_object(handles, buf.raw, group=group)
        engine._check(lib.qpe_shard_connect(h, b"".join(handles)), "qpe_shard_connect")
        # device result memory in the owner's HBM (dense result + one segment per rank >= 1, two parities)
        self.buffer = None
        obj = [None]
        if self.rank == owner:
            self.buffer = pkg.DeviceBuffer(4 * int(lib.qpe_shard_result_ids(self.world, self.seg_cap)))
            self.seg_ptr = self.buffer.ptr
            obj[0] = self.buffer.export_handle()
        dist.broadcast_object_list(obj, src=owner, group=group)
        if self.rank != owner:
            self.seg_ptr = pkg.ipc_open(obj[0])
        engine._check(lib.qpe_shard_set_device_result(h, owner, self.seg_ptr, self.seg_cap), "set_device_result")
        # host result: one shared-memory id buffer, every rank delivers its 1/world of a result over its own PCIe link
        self.host_cap = max(exchange_counts(int(host_capacity), counts_device, group))
        if self.host_cap > 0:
            name = [f"/qpe_shard_{os.getpid()}_{id(self) & 0xffff:x}" if self.rank == owner else None]
            dist.broadcast_object_list(name, src=owner, group=group)
            p = 1
            if self.rank == owner:
                p = lib.qpe_shard_open_host_result(h, name[0].encode(), self.host_cap, 1)
            dist.barrier(group=group)
            if self.rank != owner:
                p = lib.qpe_shard_open_host_result(h, name[0].encode(), self.host_cap, 0)
            if not p:
                raise pkg.QpeError("shared host result: " + (lib.qpe_gpu_last_error() or b"").decode())
            dist.barrier(group=group)   # every rank has placed its part of the buffer: now pin it
            engine._check(lib.qpe_shard_pin_host_result(h), "qpe_shard_pin_host_result")
        self._counts = (C.c_ulonglong * 16)()
        self._stats = pkg.ScanStats()
        self._C = C
        self._pstats = C.byref(self._stats)
        self._h = engine._h
        self._last_sql, self._last_bytes = None, None
        dist.barrier(group=group)
        if self.host_cap > 0 and self.rank == owner:
            lib.qpe_shard_unlink_host_result(h)   # every rank has it mapped: no name left in /dev/shm

    def _encode(self, statement: str) -> bytes:
        if statement is not self._last_sql:      # the same statement object again: no re-encoding
            self._last_sql, self._last_bytes = statement, statement.encode()
        return self._last_bytes

    def submit(self, statement: str, to_host: bool = False):
        """Enqueue one sharded full-scan SELECT (every rank calls it, same statement, same order); at most two may
        be in flight.  `wait` completes them in submission order."""
        rc = self.lib.qpe_sql_shard_submit(self._h, self._encode(statement), 1 if to_host else 0)
        if rc != 0:
            self.engine._check(rc, "qpe_shard_submit")

    def wait(self, stats: bool = True):
        """Complete the oldest query in flight: (total, per-rank counts, ScanStats / None)."""
        rc = self.lib.qpe_shard_wait(self._h, self._counts, self._pstats if stats else None)
        if rc != 0:
            self.engine._check(rc, "qpe_shard_wait")
        counts = self._counts[:self.world]
        return sum(counts), counts, (self._stats if stats else None)

    def wait_breakdown(self, reset: bool = True):
        """Host time `wait` has spent on host-result queries since the last reset, per query (ms): waiting for the counts,
        this rank's device->host copy, (owner) the other ranks' pieces."""
        import ctypes as C
        out = (C.c_double * 3)()
        n = C.c_longlong(0)
        self.lib.qpe_shard_wait_breakdown(self._h, out, C.byref(n), 1 if reset else 0)
        k = max(1, n.value)
        return [out[0] / k, out[1] / k, out[2] / k]

    def select(self, statement: str, to_host: bool = False, stats: bool = True):
        """One sharded full-scan SELECT, start to finish.  Returns (total, per-rank counts, ScanStats).
        to_host=False: ids packed in the owner's HBM (`device_result`); True: in the shared host buffer
        (`host_result`).  stats=False: no statistics (None); the event times stay unresolved."""
        self.submit(statement, to_host)
        return self.wait(stats)

    def balance_links(self, nbytes: int = 4 << 20, reps: int = 6):
        """Measure every rank's device->host rate into ITS part of the shared buffer with all links busy at once, and
        give each rank a share of every host result in proportion (the PCIe links of one box can differ by 2x under
        load: a result cut into equal parts waits for the slowest).  Every rank calls it; returns the rates (GB/s)."""
        import time
        lib, h = self.lib, self._h
        C = self._C
        src = self.pkg.DeviceBuffer(nbytes)
        cap = (self.host_cap + 1023) & ~1023
        dst = lib.qpe_shard_host_result(h) + 4 * (cap * self.rank // self.world)
        nbytes = min(nbytes, 4 * (cap // self.world))
        lib.qpe_gpu_copy_to_host(dst, src.ptr, nbytes)
        dist.barrier(group=self.group)
        t0 = time.perf_counter()
        for _ in range(reps):
            lib.qpe_gpu_copy_to_host(dst, src.ptr, nbytes)
        rate = reps * nbytes / (time.perf_counter() - t0) / 1e9
        src.free()
        rates = [None] * self.world
        dist.all_gather_object(rates, float(rate), group=self.group)
        # a link never gets less than a quarter of an equal share: the measurement is short
        floor = 0.25 * sum(rates) / self.world
        w = (C.c_double * self.world)(*[max(r, floor) for r in rates])
        self.engine._check(lib.qpe_shard_set_link_weights(h, w, self.world), "qpe_shard_set_link_weights")
        dist.barrier(group=self.group)
        return rates

    def set_multipath(self, mode: int):
        """host result: 1 = staging in HBM + copy engine (default), 2 = the kernel stores into host memory"""
        self.engine._check(self.lib.qpe_shard_set_multipath(self.engine._h, mode), "qpe_shard_set_multipath")

    def host_result(self, total: int = None, back: int = 0):
        """the ids of the most recent host-result query `wait` completed (back=1: of the one before it) -- a view of the
        shared buffer, valid until the third `submit` after the query's own.  With `set_deferred(True)` this is where
        the other ranks' pieces are waited for."""
        import numpy as np
        C = self._C
        n = C.c_ulonglong(0)
        p = self.lib.qpe_shard_host_result_at(self._h, back, C.byref(n))
        if total is None:
            total = n.value
        if not p and total > 0:
            raise self.pkg.QpeError("host result: " + (self.lib.qpe_gpu_last_error() or b"no such result").decode())
        if total <= 0:
            return np.zeros(0, dtype=np.uint32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(int(total),))

    def set_deferred(self, on: bool = True):
        """`wait` returns once THIS rank's piece of a host result is delivered (the owner does not wait for the other
        ranks' pieces before it submits its next query); `host_result` waits for the rest."""
        self.engine._check(self.lib.qpe_shard_set_deferred(self._h, 1 if on else 0), "qpe_shard_set_deferred")

    def delete(self, statement: str):
        """DELETE on the sharded table (every rank calls it): (rows deleted, rows left) over all shards.  The
        shards are renumbered, so later global row ids are positions in the table after the delete."""
        C = self._C
        deleted, left = C.c_ulonglong(), C.c_ulonglong()
        rc = self.lib.qpe_sql_shard_delete(self.engine._h, statement.encode(), C.byref(deleted), C.byref(left))
        self.engine._check(rc, "qpe_shard_delete")
        return int(deleted.value), int(left.value)

    def device_result_ptr(self) -> int:
        return self.lib.qpe_shard_device_result(self.engine._h)

    def device_result(self, total: int):
        """owner: the packed device result copied to the host (checks)"""
        return self.engine.copy_from_device(self.device_result_ptr(), total)

    def close(self):
        dist.barrier(group=self.group)
        self.lib.qpe_shard_close(self.engine._h)
        if self.rank != self.owner:
            self.pkg.ipc_close(self.seg_ptr)
        dist.barrier(group=self.group)
        if self.buffer is not None:
            self.buffer.free()
            self.buffer = None
