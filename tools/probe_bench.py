#!/usr/bin/env python3
"""B+ probe batches only (BASELINE configs[3]): 1 M point / range probes over N keys; prints JSON."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
eng = pkg.Engine.from_synth(N, columns=["command_id", "user_id"], indexes=(("command_id", 0), ("user_id", 1)))
rng = np.random.default_rng(12345)
Q = 1_000_000
out = {"keys": N, "probe": []}
for attr, dtype, hi_key in (("command_id", np.uint64, int(N * 1.1)), ("user_id", np.int32, 3100)):
    for length in (0, 15, 255, 4095):
        lo = rng.integers(0 if attr == "command_id" else 990, hi_key, size=Q).astype(dtype)
        hi = (lo + dtype(length)).astype(dtype)
        best = None
        for rep in range(5):
            first, count, st = eng.probe_batch(attr, lo, hi)
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
        # closed form for command_id (keys are 0 .. N-1, unique): first = min(lo, N), count = overlap with [0, N)
        if attr == "command_id":
            lo64, hi64 = lo.astype(np.int64), hi.astype(np.int64)
            want_first = np.minimum(lo64, N)
            want_count = np.clip(np.minimum(hi64, N - 1) - lo64 + 1, 0, None)
            assert np.array_equal(first.astype(np.int64), want_first) and np.array_equal(count.astype(np.int64), want_count)
        out["probe"].append({"index": attr, "queries": Q, "range_len": length + 1, "kernel_ms": best["kernel_ms"],
                             "probes_per_s": Q / best["kernel_ms"] * 1e3,
                             "algo_gbs": best["algo_bytes"] / best["kernel_ms"] / 1e6,
                             "hits": int((count > 0).sum())})
eng.close()
print(json.dumps(out, indent=1))
