#!/usr/bin/env python3
"""bench.py -- SELECT/WHERE scan throughput of the B200 engine (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE full-scan SELECT ... WHERE over the whole table: K1f (self-feeding TMA scan + ordered
compaction with decoupled look-back fused in one launch -> row ids in table order), and at N > 1 the
per-GPU match-count exchange plus the ordered gather of the row ids in partition order.

Workload (config.workload): the synthetic command-log table of BASELINE.json configs[4] -- 1 B rows,
generated on the device by the counter-based generator (csrc/synth.cu, distributions of the
reference's generate_commands.py), row-range sharded over the N GPUs exactly like the reference's
MPI partition (engine/mpi/executeEngine-mpi.c:703-715), total work fixed ("scaling": "strong").
Query QN of SURVEY 8(d): compound AND/OR over command_id (u64), sudo_used (bool), risk_level
(int) = 13 B/row, ~1 % selectivity.  Its matches all sit in the FIRST shard (the table is ordered by
command_id); `uniform` repeats the measurement with a query whose matches are spread evenly over the shards.

`value`  = rows scanned per second, whole job, result left packed in HBM on rank 0 (device-resident).
`e2e`    = the same through the host-facing C-ABI call (SQL text in host memory in, row ids out
           into PINNED HOST memory): per step the compiled query goes host->device and the ids
           come device->host inside the timed region.  The table itself is engine state (it is
           loaded once by initializeEngineGPU, as the reference loads its CSV once).
           At N > 1 both are measured with TWO queries in flight per rank (`config.queue_depth`: query
           q + 1 is submitted before query q is waited for -- the host's share of a query and the
           device->host copy of its ids then run beside the next scan); `sync_value` is the same with
           one query at a time.  Every step still delivers its own complete result.
`roofline` = K1f's algorithmic bytes (rows x 13 B read + 4 B per match written) / K1f's own CUDA-event
           time on the engine's stream, vs the measured HBM copy bandwidth of MEASURED_PEAKS.json.
`cpu_baseline` = the reference's own linearSearchRecords (compiled unmodified under oracle/_ref)
           on a bounded sample, 1 core (its scan loop is serial in every engine);
`cpu_baseline_omp` = the reference's QPEOMP driver (whole sample-queries.txt run, all host threads).
`extra`  (N = 1) = the other BASELINE configs under the same driver: the 100 M-row selectivity sweep
           (clustered and uniform matches), the 1 M-probe batch over 100 M keys, the drop-in API end to end
           and the index-path sample queries next to the reference.

--impl reference: the reference's serial scan run as one process per host core (query-level
parallelism, which is how QPEOMP / QPEMPI use cores), same query, bounded sample per step.  It loads
nothing of this repo's product: the sample CSV is made from the committed generator fixture.
"""
import argparse
import ctypes as C
import json
import math
import os
import re
import shutil
import statistics
import subprocess
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
# the same shapes with matches spread UNIFORMLY over the table: the bound is on user_id (drawn per row)
UNIFORM = {
    "QN": ("SELECT command_id FROM Commands WHERE (user_id < {X}) AND (sudo_used = FALSE OR risk_level > 3)",
           ["user_id", "sudo_used", "risk_level"], 9),
    "QS": ('SELECT command_id FROM Commands WHERE (user_id < {X}) AND (shell_type = "bash" OR host_name = "labpc-01")',
           ["user_id", "shell_type", "host_name"], 36),
    "QD": ('SELECT command_id FROM Commands WHERE (user_id < {X}) AND (risk_level >= 2 OR exit_code != 0) AND '
           '(command_id < 4000000000 OR shell_type != "sh")',
           ["user_id", "risk_level", "exit_code", "command_id", "shell_type"], 36),
}
ALL_COLUMNS = ["command_id", "raw_command", "base_command", "shell_type", "exit_code", "timestamp", "sudo_used",
               "working_directory", "user_id", "user_name", "host_name", "risk_level"]


def sample_queries():
    """(text of the reference's sample-queries.txt, its SELECT statements) -- tests/support.py carries the reference's
    sample-queries-FULL.txt verbatim; sample-queries.txt is that file without Sample 6 (the DELETE)"""
    import support
    text = support.SAMPLE_QUERIES_FULL.replace(
        "# -- Sample 6:\nDELETE FROM Commands WHERE command_id = 999999;\n\n", "")
    stmts = [" ".join(x.split()) for x in re.sub(r"#[^\n]*", "", text).split(";")]
    return text, [x for x in stmts if x.upper().startswith("SELECT")]


_REAL_STDOUT = None


def emit(line):
    f = _REAL_STDOUT or sys.stdout
    f.write(line + "\n")
    f.flush()


def log(msg):
    sys.stderr.write(msg + "\n")
    sys.stderr.flush()


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
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the `extra` object (configs[2], [3], API)")
    ap.add_argument("--extra-rows", type=float, default=1e8, help="rows of the `extra` sweeps / probe index")
    ap.add_argument("--only-extra", action="store_true", help="development: print only the `extra` object")
    ap.add_argument("--host-mode", type=int, default=1, choices=[1, 2],
                    help="N > 1 host result: 1 = staging + copy engine, 2 = the kernel stores into host memory")
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


def shard_of(total, world, rank):
    """contiguous row ranges, the reference's MPI rule: base = N / G, the first N % G ranks get one more"""
    base, rem = divmod(total, world)
    n = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, n


# ---------------------------------------------------------------------------------------------
# CPU side: sample CSVs and the reference's engines (oracle/_ref, compiled unmodified)
# ---------------------------------------------------------------------------------------------
def fixture_csv(rows, path):
    """A `rows`-row CSV made from the committed generator fixture (tests/golden/commands_2k.csv: 2 000 rows of the
    reference's own generate_commands.py): its rows repeated with command_id renumbered 1..rows.  Needs nothing of
    this repo's product (the reference arm must not load libqpegpu.so)."""
    src = os.path.join(ROOT, "tests", "golden", "commands_2k.csv")
    with open(src, newline="") as f:
        lines = f.read().split("\r\n")
    if len(lines) < 3:   # \n line ends
        with open(src) as f:
            lines = f.read().split("\n")
    header, body = lines[0], [ln for ln in lines[1:] if ln]
    tails = [ln[ln.index(","):] for ln in body]   # everything after command_id
    with open(path, "w", newline="") as out:
        out.write(header + "\r\n")
        k = 0
        chunk = []
        for i in range(rows):
            chunk.append(f"{i + 1}{tails[k]}")
            k = k + 1 if k + 1 < len(tails) else 0
            if len(chunk) == 65536:
                out.write("\r\n".join(chunk) + "\r\n")
                chunk = []
        if chunk:
            out.write("\r\n".join(chunk) + "\r\n")
    return path


def make_sample_csv(pkg, rows, path):
    eng = pkg.Engine.from_synth(rows, columns=ALL_COLUMNS)
    eng.write_csv(path)
    eng.close()


def cpu_baseline(csv, rows, args, sql_template, target_s=10.0):
    import support
    sql = sql_template.format(K=max(1, int(rows * args.selectivity)))
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


def _banner(text):
    vals = {}
    for key in ("Engine Initialization Time", "Query Execution Time", "Total Execution Time"):
        m = re.search(key + r"[^0-9]*([0-9.]+) seconds", text)
        if m:
            vals[key] = float(m.group(1))
    return vals


def run_cpu_driver(binary, csv, threads=None, timeout=600):
    """One run of a reference driver (QPESeq / QPEOMP) over sample-queries.txt on a private copy of `csv`."""
    import support
    exe = os.path.join(support.REF_DIR, binary)
    if not os.path.exists(exe):
        return None
    wd = tempfile.mkdtemp(prefix="qpe_drv_")
    try:
        open(os.path.join(wd, "sample-queries.txt"), "w").write(sample_queries()[0])
        data = os.path.join(wd, "data.csv")
        shutil.copyfile(csv, data)
        cmd = [exe, data] + ([str(threads)] if threads else [])
        t0 = time.perf_counter()
        r = subprocess.run(cmd, cwd=wd, capture_output=True, timeout=timeout)
        wall = time.perf_counter() - t0
        text = r.stdout.decode(errors="replace")
        return {"rc": r.returncode, "wall_s": wall, "banner": _banner(text)}
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def cpu_baseline_omp(csv, rows):
    """QPEOMP (the reference's OpenMP driver + engine, unmodified) over sample-queries.txt with every host thread,
    wall-clocked, next to QPESeq on the same file: what north_star asks to be reported beside the GPU numbers."""
    cores = os.cpu_count() or 1
    omp = run_cpu_driver("QPEOMP", csv, threads=cores)
    seq = run_cpu_driver("QPESeq", csv)
    if omp is None:
        return {"value": None, "unit": "s", "cores": cores, "kind": "reference", "sample": "oracle/_ref/QPEOMP was not built"}
    q_omp = omp["banner"].get("Query Execution Time")
    return {"value": omp["wall_s"], "unit": "s (whole run incl. ingest and index builds)", "cores": cores,
            "kind": "reference", "query_phase_s": q_omp, "rc": omp["rc"],
            "qpeseq_wall_s": seq["wall_s"] if seq else None,
            "qpeseq_query_phase_s": seq["banner"].get("Query Execution Time") if seq else None,
            "sample": f"{rows}-row CSV, the reference's sample-queries.txt (6 SELECTs + 1 INSERT), thread count argv = {cores}; "
                      f"its scan loop (linearSearchRecords) is serial, only queries / per-index probes run in parallel"}


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
    sql_t, cols, bpr = QUERIES[args.query]
    rows = 250_000
    sql = sql_t.format(K=max(1, int(rows * args.selectivity)))
    d = tempfile.mkdtemp(prefix="qpe_ref_")
    csv = fixture_csv(rows, os.path.join(d, "sample.csv"))
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
    shutil.rmtree(d, ignore_errors=True)
    value = procs * rows * args.steps / dt
    emit(json.dumps({
        "impl": "reference", "metric": "select_where_rows_per_s", "value": value, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.query} full-scan SELECT/WHERE, {args.selectivity:g} selectivity, "
                               f"reference serial engine (linearSearchRecords) x {procs} processes, "
                               f"{rows}-row sample per process per step (rows of the reference generator's fixture)",
                   "query": sql, "rows_per_step": procs * rows},
        "scan_gbs": value * bpr / 1e9,
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": procs, "kind": "reference",
                         "sample": f"{procs} processes x {rows} rows x {args.steps} steps, {matches} matches/scan"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


# ---------------------------------------------------------------------------------------------
# `extra` (N = 1): BASELINE configs[2] and [3], the drop-in API end to end, index-path latencies
# ---------------------------------------------------------------------------------------------
def best_scan(eng, sql, reps=4):
    best = None
    for _ in range(reps):
        _, _, st = eng.select_ids_device(sql, force_scan=True)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    return best


def extra_sweep(pkg, rows, peak):
    """configs[2]: 100 M-row table, QN / QS / QD at 0.01 %, 1 %, 50 %, matches clustered (command_id bound) and
    uniform (user_id bound); per point K1f's CUDA-event time (best of 4) -> algorithmic GB/s and the fraction of the
    measured copy peak."""
    import numpy as np
    cols = sorted({c for q in list(QUERIES.values()) + list(UNIFORM.values()) for c in q[1]})
    eng = pkg.Engine.from_synth(rows, columns=cols)
    # user_id bounds for the uniform variants: quantiles of the column itself (first 4 M rows)
    uid = np.sort(eng.fetch_column("user_id", 0, min(rows, 4_000_000)))
    out = []
    for sel in (0.0001, 0.01, 0.5):
        for name in ("QN", "QS", "QD"):
            for kind, (tmpl, _, bpr) in (("clustered", QUERIES[name]), ("uniform", UNIFORM[name])):
                if kind == "clustered":
                    sql = tmpl.format(K=max(1, int(rows * sel)))
                else:
                    sql = tmpl.format(X=int(uid[min(len(uid) - 1, int(len(uid) * sel))]) + 1)
                st = best_scan(eng, sql)
                gbs = st["algo_bytes"] / st["kernel_ms"] / 1e6
                out.append({"query": name, "matches": kind, "target_selectivity": sel,
                            "selectivity": st["matches"] / rows, "bytes_per_row": bpr, "kernel_ms": round(st["kernel_ms"], 4),
                            "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 4),
                            "grows_per_s": round(rows / st["kernel_ms"] / 1e6, 2),
                            "tile": f"{st['tile_rows']}x{st['stages']}"})
    eng.close()
    return {"rows": rows, "kernel": "scan_fused_kernel (K1f), CUDA events on the engine's stream, best of 4",
            "peak_gbs": peak, "points": out}


def extra_probe(pkg, rows, n_probes=1_000_000):
    """configs[3]: 1 M point / range probes over the 100 M-key command_id (unique u64) and user_id (int, heavy
    duplicates) indexes: kernel-only (device buffers), end to end with pinned host buffers, with the batch sorted
    first, and the reference's findRange on a bounded sample."""
    import numpy as np
    import support
    eng = pkg.Engine.from_synth(rows, columns=["command_id", "user_id"], indexes=(("command_id", 0), ("user_id", 1)))
    rng = np.random.default_rng(12345)
    out = {"rows": rows, "probes": n_probes, "cases": []}
    for attr, dtype, keys in (("command_id", np.uint64, rng.integers(0, int(rows * 1.1), n_probes, dtype=np.uint64)),
                              ("user_id", np.int32, rng.integers(900, 3100, n_probes).astype(np.int32))):
        lo = pkg.pinned_array(n_probes, dtype)
        hi = pkg.pinned_array(n_probes, dtype)
        first = pkg.pinned_array(n_probes, np.uint32)
        count = pkg.pinned_array(n_probes, np.uint32)
        for width in (1, 256):
            lo[:] = keys
            hi[:] = keys + dtype(width - 1)
            d_lo, d_hi = pkg.DeviceBuffer(lo.nbytes), pkg.DeviceBuffer(hi.nbytes)
            d_first, d_count = pkg.DeviceBuffer(4 * n_probes), pkg.DeviceBuffer(4 * n_probes)
            pkg.load_library().qpe_gpu_copy_to_device(d_lo.ptr, lo.ctypes.data, lo.nbytes)
            pkg.load_library().qpe_gpu_copy_to_device(d_hi.ptr, hi.ctypes.data, hi.nbytes)
            case = {"index": attr, "range_width": width}
            for label, kw, bufs in (("kernel", {}, (d_lo, d_hi, d_first, d_count)),
                                    ("kernel_sorted", {"sort": True}, (d_lo, d_hi, d_first, d_count)),
                                    ("e2e_host", {}, (lo, hi, first, count)),
                                    ("e2e_host_sorted", {"sort": True}, (lo, hi, first, count))):
                best_dev, best_wall = None, None
                for _ in range(5):
                    t0 = time.perf_counter()
                    _, _, st = eng.probe_keys(attr, bufs[0], bufs[1], first=bufs[2], count=bufs[3], **kw)
                    wall = (time.perf_counter() - t0) * 1e3
                    best_dev = st["kernel_ms"] if best_dev is None else min(best_dev, st["kernel_ms"])
                    best_wall = wall if best_wall is None else min(best_wall, wall)
                case[label] = {"device_ms": round(best_dev, 4), "wall_ms": round(best_wall, 4),
                               "gprobes_per_s": round(n_probes / (best_dev if label.startswith("kernel") else best_wall) / 1e6, 3)}
            case["found_rows"] = int(count.astype(np.int64).sum())
            # unsorted and sorted answers agree
            f2, c2, _ = eng.probe_keys(attr, lo, hi, sort=True)
            case["sorted_equals_unsorted"] = bool(np.array_equal(f2, first) and np.array_equal(c2, count))
            out["cases"].append(case)
            for b in (d_lo, d_hi, d_first, d_count):
                b.free()
    # a parity sample at full size: the row ids of a few answers against a closed form (command_id == row id)
    f, c, _ = eng.probe_keys("command_id", np.array([5, rows - 1, rows + 7], dtype=np.uint64))
    ids = [eng.index_slice("command_id", int(a), int(b)).tolist() for a, b in zip(f, c)]
    out["parity_sample"] = {"keys": [5, rows - 1, rows + 7], "row_ids": ids, "ok": ids == [[5], [rows - 1], []]}
    # K4, the sort behind the index build (csrc/radix_sort.cu), on the same two key columns: device time of the sort alone
    out["index_build"] = []
    for attr in ("command_id", "user_id"):
        keys = eng.fetch_column(attr)
        best, passes = None, 0
        for _ in range(2):
            _, _, passes, ms = pkg.sort_pairs(keys, mode=2)
            best = ms if best is None else min(best, ms)
        kb = keys.dtype.itemsize
        out["index_build"].append({"index": attr, "keys": int(keys.shape[0]), "key_bytes": kb, "digit_passes": passes,
                                   "sort_ms": round(best, 3), "gkeys_per_s": round(keys.shape[0] / best / 1e6, 2),
                                   "algo_gbs": round(keys.shape[0] * passes * (3 * kb + 8) / best / 1e6, 1)})
        del keys
    eng.close()
    # CPU: the reference's findRange on a bounded sample (it walks the leaf chain to its end: O(N) per probe)
    try:
        if support.Ref.available():
            d = tempfile.mkdtemp(prefix="qpe_probe_")
            csv = fixture_csv(200_000, os.path.join(d, "p.csv"))
            ref = support.Ref(csv, indexes=[("command_id", 0)])
            lib = ref.lib()
            lib.ref_time_probe.restype = C.c_double
            lib.ref_time_probe.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
            n = 200
            k = rng.integers(1, 200_000, n, dtype=np.uint64)
            found = C.c_longlong()
            t = lib.ref_time_probe(ref.h, 0, k.ctypes.data, k.ctypes.data, n, C.byref(found))
            ref.close()
            shutil.rmtree(d, ignore_errors=True)
            out["cpu_findRange"] = {"probes_per_s": n / t, "cores": 1, "kind": "reference",
                                    "sample": f"{n} point probes on a 200 000-key ORDER-3 B+ tree (engine/bplus.c findRange), "
                                              f"{found.value} rows found, {t:.2f} s"}
    except Exception as e:   # reported, never required
        out["cpu_findRange"] = {"probes_per_s": None, "sample": f"failed: {e}"}
    return out


def extra_api(pkg, rows=1_000_000):
    """The reference's actual entry point end to end: executeQuerySelectGPU -> resultSetS -> freeResultSet for the
    SELECTs of the reference's sample-queries.txt on a `rows`-row CSV, next to executeQuerySelectSerial +
    freeResultSet on the same file (oracle/_ref harness).  Samples 2, 3, 4, 8 take the index path (K3 + K1g), Samples 1
    and 7 the scan path (K1f); projection = K7 + K2."""
    import support
    d = tempfile.mkdtemp(prefix="qpe_api_")
    csv = os.path.join(d, "api.csv")
    make_sample_csv(pkg, rows, csv)
    out = {"rows": rows, "queries": []}
    try:
        t0 = time.perf_counter()
        eng = pkg.Engine.from_csv(csv)
        out["gpu_init_s"] = round(time.perf_counter() - t0, 3)
        ref = None
        if support.Ref.available():
            t0 = time.perf_counter()
            ref = support.Ref(csv)
            out["ref_init_s"] = round(time.perf_counter() - t0, 3)
        stmts = sample_queries()[1]
        names = [1, 2, 3, 4, 7, 8]
        for k, sql in enumerate(stmts):
            eng.select_result_time(sql)   # warm-up (pinned result pool, index build)
            reps = 3
            t0 = time.perf_counter()
            for _ in range(reps):
                n_rows = eng.select_result_time(sql)
            gpu_ms = (time.perf_counter() - t0) * 1e3 / reps
            ids, st = eng.select_ids(sql)
            q = {"sample": names[k] if k < len(names) else k + 1, "rows": n_rows, "path": "index" if st["path"] == 1 else "scan",
                 "gpu_api_ms": round(gpu_ms, 3), "gpu_match_kernel_ms": round(st["kernel_ms"], 4)}
            if ref is not None:
                m = C.c_int()
                t = ref.lib().ref_time_select(ref.h, sql.encode(), 1, C.byref(m))
                q["ref_api_ms"] = round(t * 1e3, 3)
                q["rows_equal"] = (m.value == n_rows)
                q["speedup"] = round(t * 1e3 / gpu_ms, 1)
            out["queries"].append(q)
        eng.close()
        if ref is not None:
            ref.close()
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out


def build_extra(pkg, args, peak):
    extra = {}
    rows = int(args.extra_rows)
    for name, fn in (("sweep_100m", lambda: extra_sweep(pkg, rows, peak)),
                     ("probe_100m", lambda: extra_probe(pkg, rows)),
                     ("api_1m", lambda: extra_api(pkg))):
        t0 = time.perf_counter()
        try:
            extra[name] = fn()
        except Exception as e:   # reported, never required for the headline line
            import traceback
            log(traceback.format_exc())
            extra[name] = {"failed": f"{type(e).__name__}: {e}"}
        extra[name]["wall_s"] = round(time.perf_counter() - t0, 2)
        log(f"[extra] {name}: {extra[name]['wall_s']} s")
    return extra


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
    if args.only_extra:
        emit(json.dumps({"extra": build_extra(pkg, args, peaks()[0])}))
        return 0
    total = int(args.rows)
    start, n_local = shard_of(total, world, rank)
    sql_t, cols, bpr = QUERIES[args.query]
    sql = sql_t.format(K=max(1, int(total * args.selectivity)))
    usql_t, ucols, ubpr = UNIFORM[args.query]
    all_cols = sorted(set(cols) | (set(ucols) if world > 1 else set()))
    t0 = time.perf_counter()
    eng = pkg.Engine.from_synth(total, n_rows=n_local, row_base=start, columns=all_cols)
    t_gen = time.perf_counter() - t0
    eng_stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    steps, warmup = args.steps, max(args.warmup, 3)
    peak, peak_src = peaks()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(loop, n_steps, sampler=None):
        """time `loop(n_steps)` (which runs exactly n_steps queries): CUDA events on the engine's stream and the
        host clock, max over ranks; per step"""
        sync_all()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.start()
        t_0 = time.perf_counter()
        ev0.record(eng_stream)      # the engine launches on its own stream: time THERE
        n = loop(n_steps)
        ev1.record(eng_stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t_0
        clocks = sampler.stop() if sampler else None
        if world > 1:
            dist.barrier()
        t = torch.tensor([max(ev0.elapsed_time(ev1), 0.0), wall * 1000.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return max(t[0].item(), t[1].item()) / n_steps, n, clocks

    out = None
    if world == 1:
        # ---- one GPU: the plain engine API ----
        cnt0, _, st0 = eng.select_ids_device(sql, force_scan=True)
        pinned = torch.empty(max(cnt0, 1), dtype=torch.int32).pin_memory()
        pinned_np = pinned.numpy().view(np.uint32)

        def loop_device(n):
            m = 0
            for _ in range(n):
                m, _, _ = eng.select_ids_device(sql, force_scan=True, stats=False)
            return m

        def loop_e2e(n):
            m = 0
            for _ in range(n):
                m, _ = eng.select_ids_into(sql, pinned_np, force_scan=True, stats=False)
            return m

        loop_device(warmup)
        loop_e2e(warmup)
        eng.set_timing(True)
        ms_step, n_matches, clocks = timed(loop_device, steps, ClockSampler(local_rank))
        tot = eng.timing_totals()
        eng.set_timing(False)
        assert tot["calls"] == steps, tot
        k1_ms = tot["scan_ms"] / tot["calls"]
        ms_e2e, n_e2e, _ = timed(loop_e2e, steps)
        assert n_e2e == n_matches
        launches = st0["launches"]
        sync_info = {}
        # the uniform-match query of the N > 1 runs, on one GPU (so that its scaling can be read off the per-N lines)
        uniform = None
        try:
            ueng = pkg.Engine.from_synth(total, columns=ucols)
            uid = np.sort(ueng.fetch_column("user_id", 0, 2_000_000))
            usql = usql_t.format(X=int(uid[min(len(uid) - 1, int(len(uid) * args.selectivity))]))
            ucnt, _, _ = ueng.select_ids_device(usql, force_scan=True)
            upinned = np.empty(max(ucnt, 1), dtype=np.uint32)
            upin = torch.from_numpy(upinned.view(np.int32)).pin_memory().numpy().view(np.uint32)
            ustream = torch.cuda.ExternalStream(ueng.stream, device=dev)

            def utimed(fn):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t_0 = time.perf_counter()
                e0.record(ustream)
                for _ in range(steps):
                    fn()
                e1.record(ustream)
                torch.cuda.synchronize()
                return max(e0.elapsed_time(e1), (time.perf_counter() - t_0) * 1e3) / steps

            for _ in range(warmup):
                ueng.select_ids_device(usql, force_scan=True, stats=False)
                ueng.select_ids_into(usql, upin, force_scan=True, stats=False)
            u_ms = utimed(lambda: ueng.select_ids_device(usql, force_scan=True, stats=False))
            u_e2e = utimed(lambda: ueng.select_ids_into(usql, upin, force_scan=True, stats=False))
            uniform = {"query": usql, "bytes_per_row": ubpr, "matches": int(ucnt), "value": total / (u_ms * 1e-3),
                       "ms_per_step": u_ms, "e2e": {"value": total / (u_e2e * 1e-3), "ms_per_step": u_e2e, "unit": "rows/s",
                                                    "d2h_bytes_per_step": int(4 * ucnt + 8)}}
            ueng.close()
        except Exception as e:   # reported, never required
            uniform = {"failed": f"{type(e).__name__}: {e}"}
    else:
        # ---- N GPUs: csrc/shard.cu, two queries in flight ----
        from importlib import import_module
        sharding = import_module("pqps_b200.sharding")
        usql = None
        cnt0, _, _ = eng.select_ids_device(sql, force_scan=True)
        counts0 = sharding.exchange_counts(cnt0, dev)
        # uniform query: user_id bound at the selectivity's quantile of this shard's first rows (same on every rank: rank 0's)
        uid = np.sort(eng.fetch_column("user_id", 0, min(n_local, 2_000_000)))
        x = [int(uid[min(len(uid) - 1, int(len(uid) * args.selectivity))])]
        dist.broadcast_object_list(x, src=0)
        usql = usql_t.format(X=x[0])
        ucnt0, _, _ = eng.select_ids_device(usql, force_scan=True)
        ucounts0 = sharding.exchange_counts(ucnt0, dev)
        seg_cap = int(max(max(counts0), max(ucounts0)) * 1.25) + 4096
        host_cap = max(sum(counts0), sum(ucounts0)) + 4096
        sg = sharding.ShardGroup(pkg, eng, segment_capacity=seg_cap, host_capacity=host_cap, counts_device=dev)
        sg.set_multipath(args.host_mode)
        # every rank's device->host rate with all links busy; a host result is then cut in proportion
        rates = sg.balance_links()
        link = {"d2h_gbs_per_link": [round(r, 1) for r in rates], "total_gbs": round(sum(rates), 1),
                "how": "blocking cudaMemcpy of 4 MiB from this rank's HBM into its part of the shared host buffer, x6, all "
                       "ranks at once; every rank then delivers a share of each host result proportional to its rate"}

        def make_loops(statement):
            def piped(to_host):
                def loop(n):
                    # host results: `wait` returns once this rank's piece is in host memory; the pieces of the other
                    # ranks are waited for when a result is taken (here: the last one, inside the timed region --
                    # every earlier result was complete before its id array was reused, three queries later)
                    sg.set_deferred(bool(to_host))
                    sg.submit(statement, to_host)
                    m = 0
                    for i in range(n):
                        if i + 1 < n:
                            sg.submit(statement, to_host)   # query i + 1 is in flight while query i is waited for
                        m = sg.wait(stats=False)[0]
                    if to_host:
                        sg.host_result()
                        sg.set_deferred(False)
                    return m
                return loop

            def synced(to_host):
                def loop(n):
                    m = 0
                    for _ in range(n):
                        m = sg.select(statement, to_host=to_host, stats=False)[0]
                    return m
                return loop
            return piped, synced

        piped, synced = make_loops(sql)
        _, _, st0 = sg.select(sql, to_host=False, stats=True)
        launches = st0.launches
        for to_host in (False, True):
            piped(to_host)(warmup)
        # one query at a time: K1f's own CUDA-event time (a pipelined scan starts under the previous query's
        # exchange kernel and is not timed on its own), and the sync numbers
        eng.set_timing(True)
        ms_sync, _, _ = timed(synced(False), steps)
        tot = eng.timing_totals()
        eng.set_timing(False)
        assert tot["calls"] == steps, tot
        mine = torch.tensor([tot["scan_ms"] / tot["calls"], tot["post_ms"] / tot["calls"]], dtype=torch.float64, device=dev)
        every = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(every, mine)
        k1_by_rank = [round(v[0].item(), 4) for v in every]
        post_by_rank = [round(v[1].item(), 4) for v in every]
        k1_ms = max(k1_by_rank)
        ms_e2e_sync, _, _ = timed(synced(True), steps)
        # two queries in flight
        ms_step, n_matches, clocks = timed(piped(False), steps, ClockSampler(local_rank))
        sg.wait_breakdown(reset=True)
        ms_e2e, n_e2e, _ = timed(piped(True), steps)
        assert n_e2e == n_matches
        wb = torch.tensor(sg.wait_breakdown(), dtype=torch.float64, device=dev)
        wb_all = [torch.zeros(3, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(wb_all, wb)
        wait_by_rank = [[round(x, 4) for x in v.tolist()] for v in wb_all]
        if rank == 0:   # the delivered ids are the packed device result, bit for bit
            host_ids = sg.host_result(n_e2e).copy()
        sg.select(sql, to_host=False, stats=False)
        if rank == 0:
            assert np.array_equal(host_ids, sg.device_result(n_matches)), "host result differs from the device result"
        sync_info = {"sync_ms_per_step": ms_sync, "sync_value": total / (ms_sync * 1e-3),
                     "e2e_sync_ms_per_step": ms_e2e_sync, "e2e_sync_value": total / (ms_e2e_sync * 1e-3),
                     "k1f_ms_by_rank": k1_by_rank, "post_scan_kernel_ms_by_rank": post_by_rank}
        # the same with matches spread evenly over the shards: every rank's own ids go out over its own link
        upiped, _ = make_loops(usql)
        for to_host in (False, True):
            upiped(to_host)(warmup)
        u_ms, u_matches, _ = timed(upiped(False), steps)
        sg.wait_breakdown(reset=True)
        u_e2e_ms, u_e2e_n, _ = timed(upiped(True), steps)
        assert u_e2e_n == u_matches
        uwb = torch.tensor(sg.wait_breakdown(), dtype=torch.float64, device=dev)
        uwb_all = [torch.zeros(3, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(uwb_all, uwb)
        uniform = {"query": usql, "bytes_per_row": ubpr, "matches": int(u_matches),
                   "matches_per_rank": [int(c) for c in sharding.exchange_counts(ucnt0, dev)],
                   "value": total / (u_ms * 1e-3), "ms_per_step": u_ms,
                   "e2e": {"value": total / (u_e2e_ms * 1e-3), "ms_per_step": u_e2e_ms, "unit": "rows/s",
                           "d2h_bytes_per_step": int(4 * u_matches + 8 * world),
                           "host_wait_ms_by_rank": [[round(x, 4) for x in v.tolist()] for v in uwb_all]}}
        sync_info["link"] = link
        sync_info["e2e_host_wait_ms_by_rank"] = {
            "what": "per query, host clock inside qpe_shard_wait: [waiting for the counts (scan + exchange + delivery kernel), "
                    "this rank's device->host copy, (owner) the other ranks' pieces]",
            "by_rank": wait_by_rank}

    # roofline of the dominant kernel: algorithmic bytes of a rank's shard / the kernel's own event time (max over ranks)
    shard_rows = shard_of(total, world, 0)[1]
    algo_bytes = shard_rows * bpr + 4 * n_matches // world
    achieved = algo_bytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("query") == args.query and tj.get("rows"):
                traffic = tj["dram_bytes_per_launch"] * (shard_rows / tj["rows"])
        except Exception:
            pass

    if rank == 0:
        st = eng.last_stats()
        out = {
            "metric": "select_where_rows_per_s", "value": total / (ms_step * 1e-3), "unit": "rows/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"synthetic {total}-row command-log table (device generator, seed 12345), "
                                   f"row-range sharded over {world} GPU(s)"
                                   + (", ordered gather by kernels over NVLink peer memory (csrc/shard.cu): ids stored into "
                                      "rank 0's result by the scan kernel, counts exchanged by kernel stores, no host "
                                      "collective per query; two queries in flight per rank" if world > 1 else "")
                                   + f"; {args.query} full-scan SELECT/WHERE, "
                                   f"{args.selectivity:g} selectivity; inputs ({shard_rows * bpr / 1e9:.1f} GB/GPU) "
                                   f"larger than L2, no flush needed",
                       "rows": total, "rows_per_gpu": shard_rows, "query": sql, "bytes_per_row": bpr,
                       "matches": int(n_matches), "tile_rows": st["tile_rows"], "stages": st["stages"],
                       "queue_depth": 2 if world > 1 else 1, "generate_s": round(t_gen, 2)},
            "scan_gbs": (total * bpr + 4 * n_matches) / (ms_step * 1e-3) / 1e9,
            "clocks": clocks,
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "rows/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(pkg.load_library().qpe_gpu_query_upload_bytes()) * world,
                    "d2h_bytes_per_step": int(4 * n_matches + 8 * world)},
            "gpu_launches": int(launches * steps),
            "roofline": {"bound": "hbm", "kernel": "scan_fused_kernel (K1f)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "k1_ms": k1_ms, "algorithmic_bytes_per_launch": algo_bytes},
        }
        if world > 1:
            out["e2e"]["sync_value"] = sync_info.pop("e2e_sync_value")
            out["e2e"]["sync_ms_per_step"] = sync_info.pop("e2e_sync_ms_per_step")
            out["e2e"]["host_mode"] = args.host_mode
            out.update(sync_info)
        if uniform is not None:
            out["uniform"] = uniform
    if world > 1:
        sg.close()
    eng.close()
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            d = tempfile.mkdtemp(prefix="qpe_cpu_")
            try:
                csv = os.path.join(d, "sample.csv")
                make_sample_csv(pkg, args.cpu_sample_rows, csv)
                try:
                    out["cpu_baseline"] = cpu_baseline(csv, args.cpu_sample_rows, args, sql_t)
                except Exception as e:  # the baseline is reported, never required
                    out["cpu_baseline"] = {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference",
                                           "sample": f"failed: {e}"}
                try:
                    out["cpu_baseline_omp"] = cpu_baseline_omp(csv, args.cpu_sample_rows)
                except Exception as e:
                    out["cpu_baseline_omp"] = {"value": None, "unit": "s", "cores": os.cpu_count(), "kind": "reference",
                                               "sample": f"failed: {e}"}
            finally:
                shutil.rmtree(d, ignore_errors=True)
        if not args.no_extra:
            out["extra"] = build_extra(pkg, args, peak)
    if rank == 0:
        emit(json.dumps(out))
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
