#!/usr/bin/env python3
"""Selectivity sweep (BASELINE configs[2]) and B+ probe batches (configs[3]) on one B200.
Prints one JSON object; run under gpurun and keep the output under profiles/."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
PEAK = 6538.3
try:
    PEAK = float(json.load(open(os.path.join(support.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

QS = {
    "QN": ("SELECT command_id FROM Commands WHERE (command_id < {K}) AND (sudo_used = FALSE OR risk_level > 3)", 13),
    "QS": ('SELECT command_id FROM Commands WHERE (command_id < {K}) AND (shell_type = "bash" OR host_name = "labpc-01")', 40),
    "QD": ('SELECT command_id FROM Commands WHERE (command_id < {K}) AND (risk_level >= 2 OR exit_code != 0) AND '
           '(user_id < 2000 OR shell_type != "sh")', 36),
}
out = {"rows": N, "hbm_peak_gbs": PEAK, "scan": [], "probe": []}
cols = ["command_id", "sudo_used", "risk_level", "exit_code", "user_id", "shell_type", "host_name"]
eng = pkg.Engine.from_synth(N, columns=cols, indexes=(("command_id", 0), ("user_id", 1)))
for name, (q, bpr) in QS.items():
    for frac in (0.0001, 0.001, 0.01, 0.1, 0.26, 0.52):
        sql = q.format(K=int(N * frac))
        runs = []
        for rep in range(8):
            cnt, _, st = eng.select_ids_device(sql, force_scan=True)
            if rep >= 3:
                runs.append(st)
        k1 = float(np.mean([r["scan_ms"] for r in runs]))
        kc = float(np.mean([r["compact_ms"] for r in runs]))
        tot = float(np.mean([r["kernel_ms"] for r in runs]))
        out["scan"].append({"query": name, "k_fraction": frac, "matches": cnt, "selectivity": cnt / N,
                            "bytes_per_row": bpr, "k1_ms": k1, "k1c_ms": kc, "kernels_ms": tot,
                            "k1_gbs": N * bpr / k1 / 1e6, "k1_frac_of_peak": N * bpr / k1 / 1e6 / PEAK,
                            "scan_gbs": (N * bpr + 4 * cnt) / tot / 1e6, "rows_per_s": N / tot * 1e3})
rng = np.random.default_rng(12345)
Q = 1_000_000
for attr, dtype, hi_key in (("command_id", np.uint64, int(N * 1.1)), ("user_id", np.int32, 3100)):
    for length in (0, 15, 255, 4095):
        lo = rng.integers(0 if attr == "command_id" else 990, hi_key, size=Q).astype(dtype)
        hi = (lo + dtype(length)).astype(dtype)
        best = None
        for rep in range(5):
            t0 = time.perf_counter()
            first, count, st = eng.probe_batch(attr, lo, hi)
            wall = (time.perf_counter() - t0) * 1e3
            if best is None or st["kernel_ms"] < best[0]["kernel_ms"]:
                best = (st, wall)
        st, wall = best
        out["probe"].append({"index": attr, "queries": Q, "range_len": length + 1, "kernel_ms": st["kernel_ms"],
                             "probes_per_s": Q / st["kernel_ms"] * 1e3, "algo_gbs": st["algo_bytes"] / st["kernel_ms"] / 1e6,
                             "hits": int((count > 0).sum()), "rows_answered": int(count.astype(np.int64).sum()),
                             "api_wall_ms": wall})
eng.close()
print(json.dumps(out, indent=1))
