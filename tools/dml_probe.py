#!/usr/bin/env python3
"""Time INSERT / DELETE (with their CSV side effects) and index (re)builds on a CSV-backed table, next to the
compiled reference when oracle/_ref is present.  Prints one JSON object."""
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else 1_000_000
d = tempfile.mkdtemp(prefix="qpe_dml_")
master = os.path.join(d, "master.csv")
gen = pkg.Engine.from_synth(N)
gen.write_csv(master)
gen.close()
STMTS = [
    ("insert", 'INSERT INTO Commands VALUES (99999999, "ls -la", "ls", "bash", 0, "2026-01-01T00:00:00.000Z", 1, "/tmp", 4242, "student4242", "labpc-01", 2)'),
    ("delete_few", "DELETE FROM Commands WHERE command_id = 99999999"),
    ("delete_1pct", f"DELETE FROM Commands WHERE command_id < {N // 100}"),
    ("select_after", "SELECT command_id, user_id FROM Commands WHERE user_id = 2450"),
]
out = {"rows": N, "csv_bytes": os.path.getsize(master), "ours": {}, "reference": {}}

csv = os.path.join(d, "ours.csv")
shutil.copy(master, csv)
t0 = time.perf_counter()
eng = pkg.Engine.from_csv(csv)
out["ours"]["load_and_index_s"] = time.perf_counter() - t0
for name, sql in STMTS:
    t0 = time.perf_counter()
    text = eng.run(sql, 5)
    out["ours"][name + "_s"] = time.perf_counter() - t0
    out["ours"][name + "_says"] = text.strip().splitlines()[-1][:80] if text.strip() else ""
eng.close()
out["ours"]["csv_bytes_after"] = os.path.getsize(csv)

# the reference's DELETE of 1 % of a 1 M-row table takes ~9 MINUTES (per-row B+ deletes): only on request
if support.Ref.available() and "--with-reference" in sys.argv:
    csv_r = os.path.join(d, "ref.csv")
    shutil.copy(master, csv_r)
    t0 = time.perf_counter()
    ref = support.Ref(csv_r, num_indexes=5)
    out["reference"]["load_and_index_s"] = time.perf_counter() - t0
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    for name, sql in STMTS:
        sys.stdout.flush()
        os.dup2(devnull, 1)
        t0 = time.perf_counter()
        ref.run(sql, 5)
        dt = time.perf_counter() - t0
        os.dup2(saved, 1)
        out["reference"][name + "_s"] = dt
    ref.close()
    out["reference"]["csv_bytes_after"] = os.path.getsize(csv_r)
    out["csv_identical_after"] = open(csv, "rb").read() == open(csv_r, "rb").read()
shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out, indent=1))
