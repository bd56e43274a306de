#!/usr/bin/env python3
"""One QN scan at a chosen selectivity on N rows (for ncu): ncu_scan_sel.py N selectivity reps"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
sel = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
eng = pkg.Engine.from_synth(N, columns=["command_id", "sudo_used", "risk_level"])
q = f"SELECT command_id FROM Commands WHERE (command_id < {int(N * sel)}) AND (sudo_used = FALSE OR risk_level > 3)"
for _ in range(reps):
    cnt, _, st = eng.select_ids_device(q, force_scan=True)
    print(cnt, st["kernel_ms"], st["algo_bytes"] / st["kernel_ms"] / 1e6, "GB/s")
eng.close()
