#!/usr/bin/env python3
"""Golden digests for BASELINE configs[0]'s size: tests/golden/commands_50k.csv.gz is 50 000 rows of the reference's
OWN generator (tools/gen_reference_csv.py 50000 ...: seeded, frozen clock, unmodified script); this script runs the
COMPILED, UNMODIFIED reference (oracle/_ref) over it -- the QPESeq loop with every row printed over
sample-queries-FULL, and the probe WHERE set -- and commits digests of what it returned, so that the GPU box (no
/root/reference, no generator) can check the CUDA path against the reference at that size.

    make -C oracle ref && python tools/make_golden_50k.py
"""
import gzip
import hashlib
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402
from support import GOLDEN, PROBE_WHERES, Ref, SAMPLE_QUERIES_FULL, normalise  # noqa: E402


def main():
    d = tempfile.mkdtemp(prefix="golden50k_")
    csv = os.path.join(d, "commands_50k.csv")
    with gzip.open(os.path.join(GOLDEN, "commands_50k.csv.gz"), "rb") as f, open(csv, "wb") as out:
        out.write(f.read())
    ref = Ref(support.scratch_copy(csv, tempfile.mkdtemp(prefix="golden50k_")))
    probes = []
    for w in PROBE_WHERES:
        sql = f"SELECT command_id FROM Commands WHERE {w}"
        pos = ref.scan_positions(sql)
        _, rows = ref.select(sql)
        ids = [int(r[0]) for r in rows]
        probes.append({"where": w, "n_select": len(ids), "n_scan": len(pos),
                       "select_sha256": hashlib.sha256(",".join(map(str, ids)).encode()).hexdigest(),
                       "scan_sha256": hashlib.sha256(",".join(map(str, pos)).encode()).hexdigest()})
    ref.close()
    qf = os.path.join(d, "q.txt")
    open(qf, "w").write(SAMPLE_QUERIES_FULL)
    run_csv = support.scratch_copy(csv, tempfile.mkdtemp(prefix="golden50k_"))
    text = normalise(support.ref_dump(run_csv, qf, max_rows=0))
    out = {"rows": 50000, "probes": probes,
           "sample_full": {"bytes": len(text), "lines": text.count("\n"), "sha256": hashlib.sha256(text.encode()).hexdigest(),
                           "head": text[:600]},
           "csv_after_sha256": hashlib.sha256(open(run_csv, "rb").read()).hexdigest()}
    json.dump(out, open(os.path.join(GOLDEN, "golden_50k.json"), "w"), indent=1)
    print("probes:", len(probes), "sample output:", out["sample_full"]["bytes"], "bytes")


if __name__ == "__main__":
    main()
