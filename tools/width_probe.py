#!/usr/bin/env python3
"""K1f bandwidth against the set of columns a WHERE references (1 B rows): which column mixes reach what."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
eng = pkg.Engine.from_synth(N, columns=["command_id", "sudo_used", "risk_level", "exit_code", "user_id"])
K = N // 100
QS = [
    ("u64", f"(command_id < {K})"),
    ("u64+i32", f"(command_id < {K}) AND (risk_level > 3)"),
    ("u64+u8", f"(command_id < {K}) AND (sudo_used = FALSE)"),
    ("u64+u8+i32 (QN)", f"(command_id < {K}) AND (sudo_used = FALSE OR risk_level > 3)"),
    ("i32", "(risk_level > 4)"),
    ("i32+i32+i32", "(risk_level > 4) AND (exit_code = 127 OR user_id < 1100)"),
    ("u8", "(sudo_used = TRUE) AND (sudo_used = FALSE)"),
    ("u64+3xi32+u8", f"(command_id < {K}) AND (risk_level > 3 OR exit_code != 0) AND (user_id < 2000 OR sudo_used = TRUE)"),
]
for name, w in QS:
    best = None
    for rep in range(4):
        cnt, _, st = eng.select_ids_device(f"SELECT command_id FROM Commands WHERE {w}", force_scan=True)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    bpr = (best["algo_bytes"] - 4 * best["matches"]) / N
    print(f"{name:<18} {bpr:4.0f} B/row tile={best['tile_rows']}x{best['stages']} M={best['matches']:>10} "
          f"kernel={best['kernel_ms']:.3f} ms {best['algo_bytes'] / best['kernel_ms'] / 1e6:7.1f} GB/s "
          f"{N / best['kernel_ms'] / 1e6:6.1f} Grows/s", flush=True)
eng.close()
