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


def exchange_counts(local_count: int, device, group=None) -> List[int]:
    """every rank learns every rank's match count (one 8-byte all-gather)"""
    world = dist.get_world_size(group)
    mine = torch.tensor([int(local_count)], dtype=torch.int64, device=device)
    parts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
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
