#!/usr/bin/env python3
"""Where a fused scan's time goes on a SMALL table: kernel time against row count (fixed cost + slope) and the
per-CTA time stamps of one launch (QPE_FUSE_TRACE=1): start skew, time to the first tile, finish skew."""
import os
import sys

os.environ["QPE_FUSE_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np  # noqa: E402
import support  # noqa: E402

pkg = support.load_pkg()
Q = "SELECT command_id FROM Commands WHERE (command_id < {K}) AND (sudo_used = FALSE OR risk_level > 3)"
sizes = [int(float(x)) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["25e6", "50e6", "125e6", "250e6", "500e6"])]
sel = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
rows = []
for n in sizes:
    eng = pkg.Engine.from_synth(n, columns=["command_id", "sudo_used", "risk_level"])
    sql = Q.format(K=max(1, int(n * sel)))
    best = None
    for rep in range(6):
        cnt, dptr, st = eng.select_ids_device(sql, force_scan=True)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
            tr = eng.fused_trace().astype(np.int64)
    t0 = tr[:, 0].min()
    start = (tr[:, 0] - t0) / 1e3
    first = (tr[:, 1] - tr[:, 0]) / 1e3
    evald = (tr[:, 2] - t0) / 1e3
    end = (tr[:, 3] - t0) / 1e3
    print(f"rows={n:>11} kernel={best['kernel_ms'] * 1e3:8.1f} us  {n * 13 / best['kernel_ms'] / 1e6:7.1f} GB/s  grid={best['grid']} "
          f"tile={best['tile_rows']}x{best['stages']} | CTA start skew max {start.max():5.1f} us | first tile after "
          f"{np.median(first):5.1f} (max {first.max():5.1f}) us | evaluators done: min {evald.min():7.1f} median "
          f"{np.median(evald):7.1f} max {evald.max():7.1f} us | CTA end: min {end.min():7.1f} median {np.median(end):7.1f} "
          f"max {end.max():7.1f} us | chunks/CTA {tr[:, 5].min()}..{tr[:, 5].max()}", flush=True)
    nk = np.maximum(tr[:, 5], 1)
    print(f"      per chunk (median over CTAs, us): evaluators wait for a free chunk buffer {np.median(tr[:, 4] / nk) / 1e3:.2f} | "
          f"compaction: waits for a full chunk {np.median((tr[:, 7] & 0xffffffff) / nk) / 1e3:.2f}, look-back "
          f"{np.median(tr[:, 6] / nk) / 1e3:.2f}, expansion + stores {np.median((tr[:, 7] >> 32) / nk) / 1e3:.2f}", flush=True)
    rows.append((n, best["kernel_ms"]))
    eng.close()
if len(rows) >= 2:
    x = np.array([r[0] for r in rows], dtype=np.float64)
    y = np.array([r[1] for r in rows], dtype=np.float64) * 1e3
    a, b = np.polyfit(x, y, 1)
    print(f"fit: kernel_us = {b:.1f} + rows * {a * 1e6:.4f} us/Mrow  (steady state {13 / (a * 1e6) * 1e3 / 1e3:.2f} TB/s at 13 B/row)")
