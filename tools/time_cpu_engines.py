#!/usr/bin/env python3
"""Times the reference's own CPU engines (QPESeq, QPEOMP; compiled unmodified under oracle/_ref)
and QPEGPU on the same CSV, on this host.  Reported baseline, not a target.

    python tools/time_cpu_engines.py [rows=1000000]
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402

SAMPLE = support.SAMPLE_QUERIES_FULL.replace(
    "# -- Sample 6:\nDELETE FROM Commands WHERE command_id = 999999;\n\n", "")  # = sample-queries.txt


def banner(text):
    vals = {}
    for key in ("Engine Initialization Time", "Query Loading Time", "Query Execution Time", "Total Execution Time"):
        m = re.search(key + r"[^0-9]*([0-9.]+) seconds", text)
        if m:
            vals[key] = float(m.group(1))
    return vals


def main():
    rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    pkg = support.load_pkg()
    d = tempfile.mkdtemp(prefix="qpe_cpu_")
    master = os.path.join(d, "master.csv")
    eng = pkg.Engine.from_synth(rows)
    eng.write_csv(master)
    eng.close()
    out = {"rows": rows, "host_cores": os.cpu_count(), "csv_bytes": os.path.getsize(master), "engines": {}}
    runs = [("QPESeq", [os.path.join(support.REF_DIR, "QPESeq")], SAMPLE),
            ("QPEOMP_1", [os.path.join(support.REF_DIR, "QPEOMP")], SAMPLE),
            ("QPEOMP_all", [os.path.join(support.REF_DIR, "QPEOMP")], SAMPLE),
            ("QPEGPU", [os.path.join(support.PKG_DIR, "QPEGPU")], SAMPLE)]
    for name, cmd, queries in runs:
        if not os.path.exists(cmd[0]):
            continue
        wd = os.path.join(d, name)
        os.makedirs(wd)
        open(os.path.join(wd, "sample-queries.txt"), "w").write(queries)
        csv = os.path.join(wd, "data.csv")
        shutil.copyfile(master, csv)
        args = cmd + [csv]
        if name == "QPEOMP_1":
            args.append("1")
        elif name == "QPEOMP_all":
            args.append(str(os.cpu_count()))
        elif name == "QPEGPU":
            args += [os.path.join(wd, "sample-queries.txt"), "20"]
        t0 = time.perf_counter()
        r = subprocess.run(args, cwd=wd, capture_output=True, timeout=3600)
        wall = time.perf_counter() - t0
        text = r.stdout.decode(errors="replace")
        out["engines"][name] = {"rc": r.returncode, "wall_s": wall, "banner": banner(text),
                                "query_times": [float(x) for x in re.findall(r"Query Time: ([0-9.]+) seconds", text)]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
