#!/usr/bin/env python3
"""Smallest program that launches the scan kernel (K1) on the bench workload: used under ncu."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = pkg.Engine.from_synth(N, columns=["command_id", "sudo_used", "risk_level", "shell_type", "host_name"])
QN = f"SELECT command_id FROM Commands WHERE (command_id < {N // 100}) AND (sudo_used = FALSE OR risk_level > 3)"
QS = f'SELECT command_id FROM Commands WHERE (command_id < {N // 100}) AND (shell_type = "bash" OR host_name = "labpc-01")'
for q in (QN, QS):
    for _ in range(reps):
        cnt, _, st = eng.select_ids_device(q, force_scan=True)
        print(cnt, st["kernel_ms"], st["algo_bytes"] / st["kernel_ms"] / 1e6, "GB/s")
eng.close()
