#!/usr/bin/env python3
"""bench.py -- SELECT/WHERE scan throughput of the B200 engine (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE full-scan SELECT ... WHERE over the whole table: K1f (TMA scan + ordered compaction
with decoupled look-back fused in one launch -> row ids in table order), and at N > 1 the per-GPU
match-count exchange plus the ordered gather of the row ids to rank 0, in partition order.

Workload (config.workload): the synthetic command-log table of BASELINE.json configs[4] -- 1 B rows,
generated on the device by the counter-based generator (csrc/synth.cu, distributions of the
reference's generate_commands.py), row-range sharded over the N GPUs exactly like the reference's
MPI partition (engine/mpi/executeEngine-mpi.c:703-715), total work fixed ("scaling": "strong").
Query QN of SURVEY 8(d): compound AND/OR over command_id (u64), sudo_used (bool), risk_level
(int) = 13 B/row, ~1 % selectivity.

`value`  = rows scanned per second, whole job, result left in HBM on rank 0 (device-resident).
`e2e`    = the same through the host-facing C-ABI call (SQL text in host memory in, row ids out
           into PINNED HOST memory): per step the compiled query goes host->device and the ids
           come device->host inside the timed region.  The table itself is engine state (it is
           loaded once by initializeEngineGPU, as the reference loads its CSV once).
`roofline` = K1f's algorithmic bytes (rows x 13 B read + 4 B per match written) / K1f's own CUDA-event
           time on the engine's stream, vs the measured HBM copy bandwidth of MEASURED_PEAKS.json.
`cpu_baseline` = the reference's own linearSearchRecords (compiled unmodified under oracle/_ref)
           on a bounded sample, 1 core (its scan loop is serial in every engine).

--impl reference: the reference's serial scan run as one process per host core (query-level
parallelism, which is how QPEOMP / QPEMPI use cores), same query, bounded sample per step.
"""
import argparse
import ctypes as C
import json
import math
import os
import statistics
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

QUERIES = {
    # name: (SQL template over K = command_id bound, referenced columns, bytes per row)
    "QN": ("SELECT command_id FROM Commands WHERE (command_id < {K}) AND (sudo_used = FALSE OR risk_level > 3)",
           ["command_id", "sudo_used", "risk_level"], 13),
    "QS": ('SELECT command_id FROM Commands WHERE (command_id < {K}) AND (shell_type = "bash" OR host_name = "labpc-01")',
           ["command_id", "shell_type", "host_name"], 40),
    "QD": ('SELECT command_id FROM Commands WHERE (command_id < {K}) AND (risk_level >= 2 OR exit_code != 0) AND '
           '(user_id < 2000 OR shell_type != "sh")',
           ["command_id", "risk_level", "exit_code", "user_id", "shell_type"], 36),
}
ALL_COLUMNS = ["command_id", "raw_command", "base_command", "shell_type", "exit_code", "timestamp", "sudo_used",
               "working_directory", "user_id", "user_name", "host_name", "risk_level"]


_REAL_STDOUT = None


def emit(line):
    f = _REAL_STDOUT or sys.stdout
    f.write(line + "\n")
    f.flush()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=float, default=1e9, help="rows of the whole table (all GPUs together)")
    ap.add_argument("--query", default="QN", choices=sorted(QUERIES))
    ap.add_argument("--selectivity", type=float, default=0.01)
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="native", choices=["native", "peer", "nccl"],
                    help="N > 1: 'native' = csrc/shard.cu: ids stored into rank 0's buffer by the scan kernel over "
                         "NVLink peer memory, counts exchanged by kernel stores, no host collective per query; "
                         "'peer' = same stores, count exchange through an NCCL all-gather; 'nccl' = compact "
                         "locally, then grouped NCCL send/recv")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region through NVML (nvidia-smi's own source), every few ms
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
        0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


class DevArray:
    """device pointer -> torch tensor (zero copy) through __cuda_array_interface__"""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def shard_of(total, world, rank):
    """contiguous row ranges, the reference's MPI rule: base = N / G, the first N % G ranks get one more"""
    base, rem = divmod(total, world)
    n = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, n


# ---------------------------------------------------------------------------------------------
# CPU baseline: the compiled reference's linearSearchRecords on a bounded sample
# ---------------------------------------------------------------------------------------------
def make_sample_csv(pkg, rows, path):
    eng = pkg.Engine.from_synth(rows, columns=ALL_COLUMNS)
    eng.write_csv(path)
    eng.close()


def cpu_baseline(pkg, args, sql_template, target_s=10.0):
    import support
    rows = args.cpu_sample_rows
    sql = sql_template.format(K=max(1, int(rows * args.selectivity)))
    d = tempfile.mkdtemp(prefix="qpe_cpu_")
    csv = os.path.join(d, "sample.csv")
    make_sample_csv(pkg, rows, csv)
    if support.Ref.available():
        ref = support.Ref(csv, num_indexes=0)
        m = C.c_int()
        t1 = ref.lib().ref_time_scan(ref.h, sql.encode(), 1, C.byref(m))
        reps = max(1, min(200, int(math.ceil(target_s / max(t1, 1e-3)))))
        t = ref.lib().ref_time_scan(ref.h, sql.encode(), reps, C.byref(m))
        ref.close()
        return {"value": rows * reps / t, "unit": "rows/s", "cores": 1, "kind": "reference",
                "sample": f"{rows}-row CSV from the same generator, {args.query} at {args.selectivity:g} selectivity: "
                          f"reference linearSearchRecords x{reps} ({t:.1f} s), {m.value} matches/scan"}
    # the reference was not compiled on this machine: time the oracle port instead
    o = support.Oracle.from_csv(csv)
    where = sql.split("WHERE", 1)[1]
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < target_s:
        m = len(o.scan(where))
        reps += 1
    t = time.perf_counter() - t0
    return {"value": rows * reps / t, "unit": "rows/s", "cores": 1, "kind": "port",
            "sample": f"{rows}-row CSV, oracle port of linearSearchRecords x{reps} ({t:.1f} s), {m} matches/scan"}


def _ref_worker(conn, csv, sql):
    import support
    ref = support.Ref(csv, num_indexes=0)
    m = C.c_int()
    conn.send(("ready", ref.num_rows))
    while True:
        msg = conn.recv()
        if msg == "stop":
            break
        t = ref.lib().ref_time_scan(ref.h, sql.encode(), 1, C.byref(m))
        conn.send((t, m.value))
    ref.close()


def run_reference(args):
    """--impl reference: the reference's CPU scan on all host cores (one serial engine per core)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    import support
    if not support.Ref.available():
        emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libqpe_ref.so was not built"}))
        return 0
    pkg = support.load_pkg()
    sql_t, cols, bpr = QUERIES[args.query]
    rows = 250_000
    sql = sql_t.format(K=max(1, int(rows * args.selectivity)))
    d = tempfile.mkdtemp(prefix="qpe_ref_")
    csv = os.path.join(d, "sample.csv")
    make_sample_csv(pkg, rows, csv)
    procs = max(1, min(os.cpu_count() or 1, 64))
    ctx = mp.get_context("spawn")
    workers = []
    for _ in range(procs):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(b, csv, sql), daemon=True)
        p.start()
        workers.append((p, a))
    for _, a in workers:
        a.recv()

    def step():
        t0 = time.perf_counter()
        for _, a in workers:
            a.send("go")
        res = [a.recv() for _, a in workers]
        return time.perf_counter() - t0, res[0][1]

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    matches = 0
    for _ in range(args.steps):
        _, matches = step()
    dt = time.perf_counter() - t0
    for p, a in workers:
        a.send("stop")
    for p, _ in workers:
        p.join(timeout=10)
    value = procs * rows * args.steps / dt
    emit(json.dumps({
        "impl": "reference", "metric": "select_where_rows_per_s", "value": value, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.query} full-scan SELECT/WHERE, {args.selectivity:g} selectivity, "
                               f"reference serial engine (linearSearchRecords) x {procs} processes, "
                               f"{rows}-row sample per process per step", "query": sql, "rows_per_step": procs * rows},
        "scan_gbs": value * bpr / 1e9,
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": procs, "kind": "reference",
                         "sample": f"{procs} processes x {rows} rows x {args.steps} steps, {matches} matches/scan"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import support

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the B200 engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pkg = support.load_pkg()
    total = int(args.rows)
    start, n_local = shard_of(total, world, rank)
    sql_t, cols, bpr = QUERIES[args.query]
    sql = sql_t.format(K=max(1, int(total * args.selectivity)))
    t0 = time.perf_counter()
    eng = pkg.Engine.from_synth(total, n_rows=n_local, row_base=start, columns=cols)
    t_gen = time.perf_counter() - t0

    # result buffers: device buffer on rank 0 for the gathered ids, pinned host buffer for e2e
    from importlib import import_module
    sharding = import_module("pqps_b200.sharding")
    cnt0, _, _ = eng.select_ids_device(sql, force_scan=True)
    if world > 1:
        total_matches = sum(sharding.exchange_counts(cnt0, dev))
    else:
        total_matches = cnt0
    use_peer = world > 1 and args.gather == "peer"
    use_native = world > 1 and args.gather == "native"
    if world > 1:
        seg_cap = int(max(sharding.exchange_counts(cnt0, dev)) * 1.25) + 4096
    pg = sharding.PeerGather(pkg, segment_capacity=seg_cap, counts_device=dev) if use_peer else None
    sg = (sharding.ShardGroup(pkg, eng, segment_capacity=seg_cap, host_capacity=total_matches + 4096,
                              counts_device=dev) if use_native else None)
    gathered = (torch.empty(max(total_matches, 1), dtype=torch.int32, device=dev)
                if (rank == 0 and world > 1 and not use_peer) else None)
    pinned = torch.empty(max(total_matches, 1), dtype=torch.int32).pin_memory() if rank == 0 else None
    pinned_np = pinned.numpy().view(np.uint32) if rank == 0 else None
    launches = [0]
    scan_ms, compact_ms, kernel_ms, traces = [], [], [], []
    eng_stream = torch.cuda.ExternalStream(eng.stream, device=dev)

    def step_device(record=False, pack="device"):
        """scan + ordered compaction on every GPU, count exchange, ordered gather to rank 0 (device).
        record=False: no statistics asked for (the engine's CUDA events stay unresolved and are summed after the
        loop: Engine.timing_totals); record="trace": host-side breakdown of the call (diagnostic pass)."""
        want = bool(record)
        if use_native:
            total_n, counts, st = sg.select(sql, to_host=(pack == "host"), stats=want)
            n_launch = st.launches if want else 0
        elif use_peer:
            total_n, counts, st = pg.run(eng, sql, dev, pack=pack, host_out=pinned_np)
            n_launch = st["launches"]
        else:
            cnt, dptr, st = eng.select_ids_device(sql, force_scan=True, global_ids=(world > 1), stats=want or world > 1)
            total_n = cnt
            n_launch = st["launches"] if st else 0
            if world > 1:
                mine = torch.as_tensor(DevArray(dptr, max(cnt, 1)), device=dev)[:cnt]
                total_n, counts, _ = sharding.ordered_gather(mine, gathered)
        if want:
            launches[0] = n_launch      # kernels per step (the same every step)
        if record == "trace":
            traces.append(eng.last_trace())
        return total_n

    def step_e2e():
        """host-facing call: SQL text in host memory -> row ids in pinned host memory"""
        if world == 1:
            n, _ = eng.select_ids_into(sql, pinned_np, force_scan=True, stats=False)
            return n
        n = step_device(pack="host")
        if rank == 0:
            if use_peer or use_native:
                pass  # peer: the owner copied every segment straight into the pinned host buffer;
                      # native: every rank delivered its piece into the shared pinned buffer over its own PCIe link
            else:
                pinned[:n].copy_(gathered[:n], non_blocking=True)
                torch.cuda.current_stream().synchronize()
        return n

    def timed(fn, steps, sampler=None, **kw):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.start()
        t0 = time.perf_counter()
        ev0.record(eng_stream)      # the engine launches on its own stream: time THERE
        n = 0
        for _ in range(steps):
            n = fn(**kw)
        ev1.record(eng_stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        if world > 1:
            dist.barrier()
        ms = max(ev0.elapsed_time(ev1), 0.0)
        ms = max(ms, 0.0)
        t = torch.tensor([ms, wall * 1000.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item() / steps, t[1].item() / steps, n, clocks

    for _ in range(max(args.warmup, 3)):
        step_device(record=True)
        step_e2e()
    sampler = ClockSampler(local_rank)
    eng.set_timing(True)   # sum the CUDA-event times of every call of the timed loop, resolved after it
    ms_step, wall_step, n_matches, clocks = timed(step_device, args.steps, sampler)
    tot = eng.timing_totals()
    eng.set_timing(False)
    assert tot["calls"] == args.steps, tot
    scan_ms.append(tot["scan_ms"] / tot["calls"])
    compact_ms.append(tot["compact_ms"] / tot["calls"])
    gpu_launches = launches[0] * args.steps
    for _ in range(10):  # diagnostic pass (not timed): where a step's time goes inside the call
        step_device(record="trace")
    tr = [statistics.mean(x[k] for x in traces) for k in range(6)]
    sys.stderr.write(f"[rank {rank}] per step: wall {wall_step:.4f} ms, device events {ms_step:.4f} ms; inside the call: "
                     f"compile {tr[0]:.4f}, enqueue {tr[1]:.4f}, sync {tr[2]:.4f}, tail {tr[4]:.4f}, whole call {tr[5]:.4f} ms; device: K1f "
                     f"{statistics.mean(scan_ms):.4f}, post-scan kernel {tot['post_ms'] / tot['calls']:.4f} ms "
                     f"(means over the timed loop)\n")
    ms_e2e, wall_e2e, n_e2e, _ = timed(step_e2e, args.steps)
    assert n_e2e == n_matches

    # roofline of the dominant kernel: algorithmic bytes of a rank's shard / the kernel's own event time
    # (max over ranks).  Fused scan (default): one kernel, K1f, reads the columns and writes the ids.
    k1_ms = statistics.mean(scan_ms)
    kc_ms = statistics.mean(compact_ms)
    t = torch.tensor([k1_ms, kc_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    k1_ms, kc_ms = t[0].item(), t[1].item()
    peak, peak_src = peaks()
    shard_rows = shard_of(total, world, 0)[1]
    fused = kc_ms == 0.0
    algo_bytes = shard_rows * bpr + (4 * n_matches // world if fused else 0)
    achieved = algo_bytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("query") == args.query and tj.get("rows") and tj.get("fused", False) == fused:
                traffic = tj["dram_bytes_per_launch"] * (shard_rows / tj["rows"])
        except Exception:
            pass

    if rank == 0:
        step_s = max(ms_step, wall_step) * 1e-3  # device events and host clock agree; keep the larger
        e2e_s = max(ms_e2e, wall_e2e) * 1e-3
        out = {
            "metric": "select_where_rows_per_s", "value": total / step_s, "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_s * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"synthetic {total}-row command-log table (device generator, seed 12345), "
                                   f"row-range sharded over {world} GPU(s)"
                                   + (", ordered gather = " + {"native": "scan kernel stores into rank 0 over NVLink peer memory, "
                                                                       "counts exchanged by kernel stores (no host "
                                                                       "collective per query)",
                                                             "peer": "K1c stores into rank 0 over NVLink peer memory, "
                                                                     "NCCL count all-gather",
                                                             "nccl": "NCCL send/recv"}[args.gather]
                                      if world > 1 else "")
                                   + f"; {args.query} full-scan SELECT/WHERE, "
                                   f"{args.selectivity:g} selectivity; inputs ({shard_rows * bpr / 1e9:.1f} GB/GPU) "
                                   f"larger than L2, no flush needed",
                       "rows": total, "rows_per_gpu": shard_rows, "query": sql, "bytes_per_row": bpr,
                       "matches": int(n_matches), "tile_rows": eng.last_stats()["tile_rows"],
                       "stages": eng.last_stats()["stages"], "generate_s": round(t_gen, 2)},
            "scan_gbs": (total * bpr + 4 * n_matches) / step_s / 1e9,
            "clocks": clocks,
            "e2e": {"value": total / e2e_s, "unit": "rows/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(pkg.load_library().qpe_gpu_query_upload_bytes()) * world,
                    "d2h_bytes_per_step": int(4 * n_matches + 8 * world)},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "kernel": "scan_fused_kernel (K1f)" if fused else "scan_tma_kernel (K1)",
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "k1_ms": k1_ms, "k1c_ms": kc_ms, "algorithmic_bytes_per_launch": algo_bytes},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(pkg, args, sql_t)
            except Exception as e:  # the baseline is reported, never required
                out["cpu_baseline"] = {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference",
                                       "sample": f"failed: {e}"}
        emit(json.dumps(out))
    if pg is not None:
        pg.close()
    if sg is not None:
        sg.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
