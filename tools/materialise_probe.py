#!/usr/bin/env python3
"""Time the full SELECT (match + projection into a resultSetS of C strings) through the drop-in entry
point, for result sizes where the projection dominates.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
lib = pkg.load_library()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
eng = pkg.Engine.from_synth(N, columns=pkg.COLUMNS)
out = {"rows": N, "runs": []}
for sql in ("SELECT * FROM Commands WHERE (risk_level >= 0)",
            "SELECT command_id, user_id, risk_level FROM Commands WHERE (risk_level >= 0)",
            "SELECT command_id, raw_command FROM Commands WHERE (risk_level >= 3)",
            "SELECT * FROM Commands WHERE (command_id < 1000)"):
    best = None
    for rep in range(4):
        t0 = time.perf_counter()
        res = lib.qpe_sql_select(eng._h, sql.encode())
        t1 = time.perf_counter()
        r = res.contents
        m, c, qt = r.numRecords, r.numColumns, r.queryTime
        first = [r.data[0][j].decode(errors="replace") for j in range(c)] if m else []
        last = [r.data[m - 1][j].decode(errors="replace") for j in range(c)] if m else []
        lib.freeResultSet(res)
        t2 = time.perf_counter()
        cur = {"sql": sql, "matches": m, "columns": c, "select_ms": (t1 - t0) * 1e3, "match_ms": qt * 1e3,
               "free_ms": (t2 - t1) * 1e3, "cells_per_s": m * c / (t1 - t0), "first_row": first, "last_row": last}
        if best is None or cur["select_ms"] < best["select_ms"]:
            best = cur
    out["runs"].append(best)
eng.close()
print(json.dumps(out, indent=1))
