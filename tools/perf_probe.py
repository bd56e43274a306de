#!/usr/bin/env python3
"""Quick kernel-time probe of the scan path on a synthetic table (not the bench; a dev tool)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
cols = ["command_id", "sudo_used", "risk_level", "exit_code", "user_id", "shell_type", "host_name"]
t0 = time.time()
eng = pkg.Engine.from_synth(N, columns=cols)
print(f"synth {N} rows in {time.time() - t0:.2f}s", flush=True)
QS = {
    "QN": "SELECT command_id FROM Commands WHERE (command_id < {K}) AND (sudo_used = FALSE OR risk_level > 3)",
    "QS": 'SELECT command_id FROM Commands WHERE (command_id < {K}) AND (shell_type = "bash" OR host_name = "labpc-01")',
    "QD": 'SELECT command_id FROM Commands WHERE (command_id < {K}) AND (risk_level >= 2 OR exit_code != 0) AND (user_id < 2000 OR shell_type != "sh")',
}
tiles = [(0, 0)]
if len(sys.argv) > 2:
    tiles = [tuple(int(x) for x in a.split("x")) for a in sys.argv[2].split(",")]
for tr, stg in tiles:
    eng.set_tile(tr, stg)
    for name, q in QS.items():
        for frac in (0.0001, 0.01, 0.1, 0.5):
            sql = q.format(K=int(N * frac))
            best = None
            for rep in range(4):
                cnt, dptr, st = eng.select_ids_device(sql, force_scan=True)
                if best is None or st["kernel_ms"] < best["kernel_ms"]:
                    best = st
            gbs = best["algo_bytes"] / best["kernel_ms"] / 1e6
            print(f"{name} sel={frac:<7} tile={best['tile_rows']}x{best['stages']} grid={best['grid']} "
                  f"M={best['matches']:>10} kernel={best['kernel_ms']:.3f} ms  {gbs:8.1f} GB/s  "
                  f"{N / best['kernel_ms'] / 1e6:.2f} Grows/s total_ms={best['total_ms']:.3f}", flush=True)
eng.close()
