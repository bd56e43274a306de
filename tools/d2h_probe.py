#!/usr/bin/env python3
"""Device -> host copy rate per PCIe link when 1, some or all GPUs of the box copy at the same time (torchrun, one
rank per GPU): into the SHARED result buffer of csrc/shard.cu (POSIX shm + cudaHostRegister, this rank's part placed
on its GPU's NUMA node) and into a private cudaHostAlloc buffer.  Explains the end-to-end numbers of bench.py at N > 1.

    python -m torch.distributed.run --nproc-per-node 8 tools/d2h_probe.py [bytes_per_rank ...]
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import support  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pkg = support.load_pkg()
lib = pkg.load_library()
from importlib import import_module  # noqa: E402
sharding = import_module("pqps_b200.sharding")

sizes = [int(float(x)) for x in sys.argv[1:]] or [4_725_824, 37_806_592]
cap_ids = max(sizes) * world // 4 + 4096
eng = pkg.Engine.from_synth(1_000_000, n_rows=1_000_000 // world, row_base=rank * (1_000_000 // world),
                            columns=["command_id", "sudo_used", "risk_level"])
sg = sharding.ShardGroup(pkg, eng, segment_capacity=4096, host_capacity=cap_ids, counts_device=dev)
how = C.c_int()
node = lib.qpe_shard_numa(eng._h, C.byref(how))
shared = lib.qpe_shard_host_result(eng._h)   # parity 0 of the shared buffer
src = pkg.DeviceBuffer(max(sizes))
private = pkg.pinned_array(max(sizes), np.uint8)


def rate(dst_ptr, nbytes, active, reps=8):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps):
            lib.qpe_gpu_copy_to_host(dst_ptr, src.ptr, nbytes)
    dt = time.perf_counter() - t0
    dist.barrier()
    g = torch.tensor([reps * nbytes / dt / 1e9 if active else 0.0], dtype=torch.float64, device=dev)
    every = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(every, g)
    return [round(v.item(), 1) for v in every]


info = [None] * world
dist.all_gather_object(info, {"rank": rank, "gpu_numa_node": node, "placed_by": how.value,
                              "cpu": os.sched_getcpu() if hasattr(os, "sched_getcpu") else None})
out = {"world": world, "ranks": info, "cases": []}
groups = {"all": list(range(world)), "first_half": list(range(world // 2)), "second_half": list(range(world // 2, world)),
          "rank0_only": [0], "last_only": [world - 1]}
for nbytes in sizes:
    for gname, members in groups.items():
        mine_shared = shared + 4 * (cap_ids * rank // world)   # this rank's part of the shared buffer
        far_shared = shared + 4 * (cap_ids * ((rank + world // 2) % world) // world)   # a part placed by a rank of the other half
        for dname, ptr in (("shared_own_part", mine_shared), ("shared_other_half", far_shared), ("private_pinned", private.ctypes.data)):
            r = rate(ptr, nbytes, rank in members)
            if rank == 0:
                out["cases"].append({"bytes": nbytes, "active": gname, "dst": dname, "gbs_per_rank": r,
                                     "total_gbs": round(sum(r), 1)})
if rank == 0:
    print(json.dumps(out))
sg.close()
eng.close()
dist.destroy_process_group()
