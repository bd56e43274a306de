"""Times K4 (the engine's own radix sort) alone and through the index build.  python tools/time_index_build.py [rows]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
out = {"rows": n}
rng = np.random.default_rng(1)
for name, keys in (("u64 row ids (4 varying bytes)", rng.permutation(n).astype(np.uint64)),
                   ("u64 random 64-bit", rng.integers(0, 1 << 64, size=n, dtype=np.uint64)),
                   ("int 0..4999 (user_id)", rng.integers(0, 5000, size=n, dtype=np.int64).astype(np.int32)),
                   ("int random 32-bit", rng.integers(-(1 << 31), 1 << 31, size=n, dtype=np.int64).astype(np.int32))):
    best = None
    for _ in range(3):
        _, _, passes, ms = pkg.sort_pairs(keys, mode=2)
        best = ms if best is None else min(best, ms)
    kb = keys.dtype.itemsize
    out[name] = {"passes": passes, "ms": round(best, 3), "gkeys_per_s": round(n / best / 1e6, 2),
                 "algo_gbs": round(n * passes * (3 * kb + 8) / best / 1e6, 1)}
del keys
t0 = time.perf_counter()
eng = pkg.Engine.from_synth(n, columns=["command_id", "user_id", "risk_level", "exit_code"],
                            indexes=(("command_id", 0), ("user_id", 1), ("risk_level", 1), ("exit_code", 1)))
ids, st = eng.select_ids("SELECT command_id FROM Commands WHERE user_id = 1001")   # builds what is still dirty
out["engine_with_4_indexes_s"] = round(time.perf_counter() - t0, 3)
eng.close()
print(json.dumps(out))
