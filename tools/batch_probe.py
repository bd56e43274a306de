#!/usr/bin/env python3
"""Query batch (K9) against the same queries run one by one, on one B200."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
cols = ["command_id", "sudo_used", "risk_level", "exit_code", "user_id", "shell_type", "host_name"]
eng = pkg.Engine.from_synth(N, columns=cols)
SETS = {
    "8 x QN-like (13 B/row union)": [
        f"SELECT command_id FROM Commands WHERE (command_id < {int(N * f)}) AND (sudo_used = {b} OR risk_level > {r})"
        for f, b, r in [(0.01, "FALSE", 3), (0.02, "TRUE", 2), (0.005, "FALSE", 1), (0.03, "TRUE", 4),
                        (0.015, "FALSE", 2), (0.001, "TRUE", 3), (0.025, "FALSE", 4), (0.008, "TRUE", 1)]],
    "8 x one text equality (16 B/row, HBM-bound alone)": [
        f'SELECT command_id FROM Commands WHERE (host_name = "labpc-0{k}")' for k in range(1, 9)],
    "8 x two text columns (32 B/row)": [
        f'SELECT command_id FROM Commands WHERE (host_name = "labpc-0{k}") AND (shell_type != "{sh}")'
        for k, sh in zip(range(1, 9), ["bash", "zsh", "sh", "fish", "bash", "zsh", "sh", "fish"])],
    "8 x mixed numeric + text (different column sets: no sharing)": [
        f"SELECT command_id FROM Commands WHERE (command_id < {int(N * 0.01)}) AND (sudo_used = FALSE OR risk_level > 3)",
        f'SELECT command_id FROM Commands WHERE (command_id < {int(N * 0.01)}) AND (shell_type = "bash" OR host_name = "labpc-01")',
        f'SELECT command_id FROM Commands WHERE (command_id < {int(N * 0.02)}) AND (risk_level >= 2 OR exit_code != 0) AND (user_id < 2000 OR shell_type != "sh")',
        "SELECT command_id FROM Commands WHERE (exit_code = 127)",
        'SELECT command_id FROM Commands WHERE (host_name = "labpc-07") AND (risk_level > 2)',
        "SELECT command_id FROM Commands WHERE (user_id = 2450)",
        f'SELECT command_id FROM Commands WHERE (shell_type = "zsh") AND (command_id > {int(N * 0.99)})',
        "SELECT command_id FROM Commands WHERE (risk_level = 5) AND (sudo_used = TRUE) AND (exit_code != 0)"],
}
out = {"rows": N, "sets": []}
for name, stmts in SETS.items():
    best_single = best_batch = None
    for rep in range(4):
        t0 = time.perf_counter()
        k_ms = 0.0
        singles = []
        for s in stmts:
            ids, st = eng.select_ids(s, force_scan=True)
            singles.append(len(ids))
            k_ms += st["kernel_ms"]
        t_single = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        res, st = eng.select_ids_batch(stmts)
        t_batch = (time.perf_counter() - t0) * 1e3
        assert [len(r) for r in res] == singles
        if best_single is None or t_single < best_single[0]:
            best_single = (t_single, k_ms)
        if best_batch is None or t_batch < best_batch[0]:
            best_batch = (t_batch, st["kernel_ms"], st["launches"], st["tile_rows"], st["stages"])
    out["sets"].append({"set": name, "queries": len(stmts), "matches": singles,
                        "one_by_one_wall_ms": best_single[0], "one_by_one_kernels_ms": best_single[1],
                        "batch_wall_ms": best_batch[0], "batch_kernels_ms": best_batch[1], "batch_launches": best_batch[2],
                        "batch_tile": f"{best_batch[3]}x{best_batch[4]}",
                        "kernel_speedup": best_single[1] / best_batch[1]})
eng.close()
print(json.dumps(out, indent=1))
