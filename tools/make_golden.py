#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ from the COMPILED, UNMODIFIED reference
(oracle/_ref, built from /root/reference by oracle/Makefile).  Run in the authoring container:

    make -C oracle ref && python tools/make_golden.py

Outputs (committed; the GPU box has no /root/reference and only reads these):
  tests/golden/probe_results_2k.json   per probe WHERE: SELECT result (command_id order) through
                                       executeQuerySelectSerial with the 5 default indexes, and
                                       the linearSearchRecords positions of the same WHERE
  tests/golden/where_text.json         the reference tokenizer+parser+convert_conditions output
                                       for a list of statements, incl. malformed ones
  tests/golden/sample_full_2k.out      normalised stdout of the QPESeq loop over sample-queries-FULL
  tests/golden/commands_2k_after_sample_full.csv   the CSV after that run (INSERT + DELETE side effects)
  tests/golden/index_order_2k.json     leaf-chain order (full-range SELECT) per probe-able index
"""
import ctypes as C
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import support  # noqa: E402
from support import CSV_2K, GOLDEN, PROBE_WHERES, Ref, SAMPLE_QUERIES_FULL  # noqa: E402

WHERE_TEXT_STATEMENTS = [
    f"SELECT command_id FROM Commands WHERE {w}" for w in PROBE_WHERES
] + [
    "SELECT * FROM Commands",
    "DELETE FROM Commands WHERE command_id = 999999",
    "select command_id from Commands where risk_level = 1 and user_id = 1001",
    "SELECT a FROM t WHERE a = 1 or b = 2",
    "SELECT a FROM t WHERE a=1 AND b=2 AND c=3 AND d=4 AND e=5",
    "SELECT a FROM t WHERE a=1 AND b=2 AND c=3 AND d=4 AND e=5 AND f=6",
    "SELECT a FROM t WHERE (a=1 AND (b=2 OR c=3)) OR d != 'x y'",
    "SELECT a, b FROM t WHERE 12abc = 3 AND x >= -5",
    "# -- Sample 9:\nSELECT x FROM Commands -- trailing\nWHERE risk_level > 3",
    "SELECT a FROM t WHERE a > 1 ORDER BY a DESC",
    "INSERT INTO Commands VALUES (1, 'x')",
    "DESCRIBE Commands",
    "FROM Commands",
    "hello world",
]


def main():
    lib = Ref.lib()
    lib.ref_where_text.restype = C.c_void_p
    lib.ref_where_text.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    lib.ref_free.argtypes = [C.c_void_p]

    d = tempfile.mkdtemp(prefix="golden_")
    ref = Ref(support.scratch_copy(CSV_2K, d))
    probes = []
    for w in PROBE_WHERES:
        sql = f"SELECT command_id FROM Commands WHERE {w}"
        _, rows = ref.select(sql)
        probes.append({"where": w, "select_command_ids": [int(r[0]) for r in rows],
                       "scan_positions": ref.scan_positions(sql)})
    json.dump(probes, open(os.path.join(GOLDEN, "probe_results_2k.json"), "w"))

    order = {}
    for attr in ("command_id", "user_id", "risk_level", "exit_code"):
        _, rows = ref.select(f"SELECT command_id FROM Commands WHERE {attr} >= 0")
        order[attr] = [int(r[0]) for r in rows]
    json.dump(order, open(os.path.join(GOLDEN, "index_order_2k.json"), "w"))
    ref.close()

    wt = []
    for s in WHERE_TEXT_STATEMENTS:
        cmd = C.c_int()
        p = lib.ref_where_text(s.encode(), C.byref(cmd))
        wt.append({"statement": s, "command": cmd.value, "where": C.string_at(p).decode()})
        lib.ref_free(p)
    json.dump(wt, open(os.path.join(GOLDEN, "where_text.json"), "w"), indent=1)

    qf = os.path.join(d, "q.txt")
    open(qf, "w").write(SAMPLE_QUERIES_FULL)
    d2 = tempfile.mkdtemp(prefix="golden_")
    out = support.ref_dump(support.scratch_copy(CSV_2K, d2), qf, max_rows=20)
    open(os.path.join(GOLDEN, "sample_full_2k.out"), "w").write(support.normalise(out))
    csv_after = open(os.path.join(d2, os.path.basename(CSV_2K)), "rb").read()
    open(os.path.join(GOLDEN, "commands_2k_after_sample_full.csv"), "wb").write(csv_after)
    print("golden vectors written to", GOLDEN)


if __name__ == "__main__":
    main()
