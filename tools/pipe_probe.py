#!/usr/bin/env python3
"""Pipelined scan probe (dev tool): K1c of table segment i beside K1 of segment i+1, ids stored to HBM
or straight into pinned host memory.  usage: pipe_probe.py ROWS [P,P,...] [selectivity,...]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
PS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 0]
SELS = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.01]
Q = "SELECT command_id FROM Commands WHERE (command_id < {K}) AND (sudo_used = FALSE OR risk_level > 3)"
eng = pkg.Engine.from_synth(N, columns=["command_id", "sudo_used", "risk_level"])
pinned = torch.empty(int(N * max(SELS)) + 1024, dtype=torch.int32).pin_memory()
pin = pinned.numpy().view(np.uint32)
pageable = np.empty(pin.size, dtype=np.uint32)
TILES = [tuple(int(v) for v in x.split("x")) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [(0, 0)]
ref = {}
for sel in SELS:
    sql = Q.format(K=int(N * sel))
    for P, (tr, stg) in [(P, t) for P in PS for t in TILES]:
        eng.set_pipeline(P)
        eng.set_tile(tr, stg)
        for dst in ("hbm", "pinned", "pageable"):
            best_wall, best = 1e9, None
            for rep in range(6):
                t0 = time.perf_counter()
                if dst == "hbm":
                    n, _, st = eng.select_ids_device(sql, force_scan=True)
                else:
                    n, st = eng.select_ids_into(sql, pin if dst == "pinned" else pageable, force_scan=True)
                wall = (time.perf_counter() - t0) * 1e3
                if rep >= 2 and wall < best_wall:
                    best_wall, best = wall, st
            if dst == "pinned":
                h = (int(n), int(pin[:n].astype(np.uint64).sum()), int((pin[:n].astype(np.uint64) * np.arange(1, n + 1, dtype=np.uint64)).sum() & 0xffffffffffff))
                if sel not in ref:
                    ref[sel] = h
                ok = "same" if ref[sel] == h else f"DIFFERENT {h} vs {ref[sel]}"
            else:
                ok = ""
            print(f"sel={sel:<6} P={P:<2} dst={dst:<8} M={n:>10} tile={best['tile_rows']}x{best['stages']} wall={best_wall:7.3f} ms  "
                  f"kernels={best['kernel_ms']:7.3f} (K1 {best['scan_ms']:.3f} + tail {best['compact_ms']:.3f})  "
                  f"{N / best_wall / 1e6:7.1f} Grows/s e2e  K1 {N * 13 / best['scan_ms'] / 1e6:7.1f} GB/s {ok}", flush=True)
eng.close()
