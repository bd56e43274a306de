"""CPU tests of the ORACLE (oracle/liboracle.so, the plain-C restatement of the reference path):
pinned against (1) the known-answer tests of the reference's own test programs, (2) golden vectors
produced by the compiled unmodified reference (tools/make_golden.py), and (3) -- when
oracle/_ref exists on this machine -- the compiled reference itself, live."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import support
from support import CSV_2K, GOLDEN, Oracle, PROBE_WHERES, Ref

pytestmark = pytest.mark.skipif(not Oracle.available(), reason="oracle/liboracle.so not built (run __graft_entry__.build())")


def _table(rows):
    """rows: list of dicts with the numeric fields; text fields default to 'x'."""
    n = len(rows)

    def text(key, width=16):
        a = np.zeros((n, width), dtype=np.uint8)
        for i, r in enumerate(rows):
            b = r.get(key, "x").encode()
            a[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
        return a

    cols = {
        "command_id": np.array([r.get("command_id", i) for i, r in enumerate(rows)], dtype=np.uint64),
        "exit_code": np.array([r.get("exit_code", 0) for r in rows], dtype=np.int32),
        "user_id": np.array([r.get("user_id", 0) for r in rows], dtype=np.int32),
        "risk_level": np.array([r.get("risk_level", 0) for r in rows], dtype=np.int32),
        "sudo_used": np.array([1 if r.get("sudo_used") else 0 for r in rows], dtype=np.uint8),
    }
    for k in ("raw_command", "base_command", "shell_type", "timestamp", "working_directory", "user_name", "host_name"):
        cols[k] = text(k, 32)
    return Oracle.from_columns(cols)


# ---- (1) known-answer tests restated from the reference's tests/ ----------------------------
def test_kat_evaluate_where_clause():
    """tests/executeEngine-serial-test.c:9-115 -- one record (risk_level 5, user_id 10), three clauses, all true"""
    o = _table([{"risk_level": 5, "user_id": 10}])
    assert o.scan("risk_level > 3").tolist() == [0]                               # :39
    assert o.scan("(risk_level > 3 AND user_id = 10)").tolist() == [0]            # :70
    assert o.scan("(risk_level > 10) OR (user_id = 10)").tolist() == [0]          # :113
    assert o.scan("(risk_level > 10) AND (user_id = 10)").tolist() == []


def test_kat_duplicate_keys():
    """tests/duplicate-test.c:8-61 -- 4 rows, risk_level keys 1,1,1,2: find_rows(1) -> 3 rows, find_rows(2) -> 1"""
    o = _table([{"command_id": 1, "risk_level": 1}, {"command_id": 2, "risk_level": 1},
                {"command_id": 3, "risk_level": 1}, {"command_id": 4, "risk_level": 2}])
    idx = [("risk_level", 1)]
    ids, used = o.select_ids("risk_level = 1", idx)
    assert used and len(ids) == 3
    assert ids.tolist() == [2, 1, 0]  # leaf chain: newest duplicate first
    ids, used = o.select_ids("risk_level = 2", idx)
    assert used and ids.tolist() == [3]


def test_kat_bplus_range():
    """tests/bplus-serial-test.c:15-46 -- keys 5,15,25,35,45: range 10..30 -> 15,25 ; 5..45 -> all five"""
    o = _table([{"command_id": k} for k in (5, 15, 25, 35, 45)])
    idx = [("command_id", 0)]
    ids, _ = o.select_ids("command_id >= 10 AND command_id <= 30", idx)
    # two segments (>=10: 15..45, <=30: 5..25) filtered by the whole WHERE: 15,25 then 15,25 (duplicates kept)
    assert [int(o.cell(i, "command_id")) for i in ids] == [15, 25, 15, 25]
    ids, _ = o.select_ids("command_id >= 5", idx)
    assert [int(o.cell(i, "command_id")) for i in ids] == [5, 15, 25, 35, 45]


def test_kat_delete_mask():
    """tests/delete-test.c:17-110 -- 3 rows, DELETE command_id = 2 matches exactly one row; key gone afterwards"""
    o = _table([{"command_id": 1}, {"command_id": 2}, {"command_id": 3}])
    assert o.scan("command_id = 2").tolist() == [1]
    keep = [i for i in range(3) if i not in o.scan("command_id = 2").tolist()]
    o2 = _table([{"command_id": c} for c in (1, 3)])
    assert keep == [0, 2]
    ids, _ = o2.select_ids("command_id = 2", [("command_id", 0)])
    assert len(ids) == 0


def test_index_order_closed_form_equals_insertion_replay():
    rng = np.random.default_rng(7)
    rows = [{"command_id": int(rng.integers(0, 50)), "risk_level": int(rng.integers(1, 6)),
             "exit_code": int(rng.integers(-3, 3))} for _ in range(3000)]
    o = _table(rows)
    for attr in ("command_id", "risk_level", "exit_code"):
        assert o.index_order(attr).tolist() == o.index_order(attr, by_insertion=True).tolist()


# ---- (2) golden vectors from the compiled reference ------------------------------------------
@pytest.fixture(scope="module")
def oracle_2k():
    o = Oracle.from_csv(CSV_2K)
    yield o
    o.close()


def _golden(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def test_loader_row_count(oracle_2k):
    assert oracle_2k.num_rows == 2000
    assert oracle_2k.cell(13, "raw_command") == 'grep -R "ERROR" /home/student1012/projects/cs202'  # quoted CSV field
    assert oracle_2k.cell(0, "sudo_used") == "false"


@pytest.mark.parametrize("entry", _golden("probe_results_2k.json"), ids=lambda e: e["where"][:50])
def test_select_matches_golden(oracle_2k, entry):
    ids, _ = oracle_2k.select_ids(entry["where"])
    got = [int(oracle_2k.cell(i, "command_id")) for i in ids]
    assert got == entry["select_command_ids"]
    assert oracle_2k.scan(entry["where"]).tolist() == entry["scan_positions"]


def test_index_order_matches_golden(oracle_2k):
    want = _golden("index_order_2k.json")
    for attr, order in want.items():
        perm = oracle_2k.index_order(attr)
        assert [int(oracle_2k.cell(i, "command_id")) for i in perm] == order


def test_where_tree_builder_matches_reference_parser():
    """the tests' own WHERE-tree builder renders the same list as the reference's tokenizer+parser"""
    gold = {e["statement"]: e["where"] for e in _golden("where_text.json")}
    for w in PROBE_WHERES:
        assert support.render_where_tree(support.parse_where(w)) == gold[f"SELECT command_id FROM Commands WHERE {w}"]


# ---- (3) live differential against the compiled reference, when present ----------------------
@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built on this machine")
def test_oracle_vs_compiled_reference_live(oracle_2k, tmp_path):
    r = Ref(support.scratch_copy(CSV_2K, tmp_path))
    cols = ["command_id", "raw_command", "user_name", "sudo_used", "exit_code", "timestamp"]
    for w in PROBE_WHERES:
        sql = f"SELECT {', '.join(cols)} FROM Commands WHERE {w}"
        _, rows = r.select(sql)
        ids, _ = oracle_2k.select_ids(w)
        assert oracle_2k.rows(ids, cols) == rows, w
    r.close()


# ---- (4) the same differential at BASELINE configs[0]'s size: 50 000 rows from the reference's OWN generator ----
@pytest.mark.skipif(not (Ref.available() and os.path.exists("/root/reference/data-generation/generate_commands.py")),
                    reason="needs the reference tree and oracle/_ref (authoring container only)")
def test_oracle_vs_compiled_reference_50k(tmp_path):
    """commands_50k.csv is a git-LFS pointer in the reference, so the file is regenerated with the reference's
    unmodified generator (seeded, tools/gen_reference_csv.py); the compiled QPESeq engine and the oracle must
    then agree on every probe WHERE: same rows, same order (index path included), same cell text."""
    import subprocess
    import sys
    csv = str(tmp_path / "commands_50k.csv")
    gen = os.path.join(support.ROOT, "tools", "gen_reference_csv.py")
    subprocess.run([sys.executable, gen, "50000", csv], check=True, capture_output=True, timeout=600)
    o = Oracle.from_csv(csv)
    (tmp_path / "ref").mkdir()
    r = Ref(support.scratch_copy(csv, tmp_path / "ref"))
    try:
        assert o.num_rows == r.num_rows == 50000
        cols = ["command_id", "user_name", "sudo_used", "exit_code"]
        for w in PROBE_WHERES:
            _, rows = r.select(f"SELECT {', '.join(cols)} FROM Commands WHERE {w}")
            ids, _ = o.select_ids(w)
            assert len(ids) == len(rows), w
            assert [int(x[0]) for x in rows] == [int(o.cell(int(i), "command_id")) for i in ids], w   # same rows, same order
            step = max(1, len(ids) // 500)                                                            # cell text on a sample
            assert o.rows(ids[::step], cols) == rows[::step], w
    finally:
        r.close()
        o.close()
