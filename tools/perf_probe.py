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
pipe = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # 0 = fused K1f, 1 = K1 then K1c
only = sys.argv[4].split(",") if len(sys.argv) > 4 else list(QS)
fracs = [float(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0.0001, 0.01, 0.1, 0.5]
eng.set_pipeline(pipe)
for tr, stg in tiles:
    try:
        eng.set_tile(tr, stg)
    except Exception as e:
        print(f"tile {tr}x{stg}: {e}")
        continue
    for name, q in QS.items():
        if name not in only:
            continue
        for frac in fracs:
            sql = q.format(K=int(N * frac))
            best = None
            for rep in range(4):
                try:
                    cnt, dptr, st = eng.select_ids_device(sql, force_scan=True)
                except Exception as e:
                    print(f"{name} tile {tr}x{stg}: {e}")
                    break
                if best is None or st["kernel_ms"] < best["kernel_ms"]:
                    best = st
            if best is None:
                continue
            gbs = best["algo_bytes"] / best["kernel_ms"] / 1e6
            print(f"pipe={pipe} {name} sel={frac:<7} tile={best['tile_rows']}x{best['stages']} grid={best['grid']} "
                  f"M={best['matches']:>10} kernel={best['kernel_ms']:.3f} ms  {gbs:8.1f} GB/s  "
                  f"{N / best['kernel_ms'] / 1e6:.2f} Grows/s total_ms={best['total_ms']:.3f}", flush=True)
eng.close()
