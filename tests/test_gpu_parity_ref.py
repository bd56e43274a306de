"""GPU engine vs the compiled, UNMODIFIED reference serial engine (oracle/_ref/libqpe_ref.so)
on the golden 2 000-row CSV made by the reference's own generator: same rows, same order,
same cell text, same printed output -- through the C-ABI of libqpegpu.so."""
import os

import numpy as np
import pytest

import support
from support import CSV_2K, PROBE_WHERES, Ref, SAMPLE_QUERIES_FULL, normalise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return support.load_pkg()


@pytest.fixture(scope="module")
def engines(pkg, tmp_path_factory):
    if not Ref.available():
        pytest.skip("oracle/_ref/libqpe_ref.so not built")
    d = tmp_path_factory.mktemp("parity")
    (d / "gpu").mkdir()
    (d / "ref").mkdir()
    g = pkg.Engine.from_csv(support.scratch_copy(CSV_2K, d / "gpu"))
    r = Ref(support.scratch_copy(CSV_2K, d / "ref"))
    yield g, r
    g.close()
    r.close()


def test_row_count(engines):
    g, r = engines
    assert g.num_rows == r.num_rows == 2000


@pytest.mark.parametrize("where", PROBE_WHERES)
def test_select_rows_identical(engines, where):
    g, r = engines
    sql = f"SELECT command_id, user_name, shell_type, risk_level, sudo_used, exit_code FROM Commands WHERE {where}"
    names_g, rows_g, _ = g.select(sql)
    names_r, rows_r = r.select(sql)
    assert names_g == names_r
    assert len(rows_g) == len(rows_r)
    assert rows_g == rows_r


@pytest.mark.parametrize("where", PROBE_WHERES)
def test_forced_scan_positions(engines, where):
    """full-scan path == linearSearchRecords over the table, whatever the indexes"""
    g, r = engines
    sql = f"SELECT command_id FROM Commands WHERE {where}"
    ids, st = g.select_ids(sql, force_scan=True)
    assert ids.tolist() == r.scan_positions(sql)
    assert st["path"] == 0
    bits, cnt = g.match_mask(sql)
    assert cnt == len(ids)
    assert np.flatnonzero(bits).tolist() == ids.tolist()


def test_select_star_and_unknown_column(engines):
    g, r = engines
    for sql in ["SELECT * FROM Commands WHERE risk_level = 5",
                "SELECT * FROM Commands WHERE command_id < 25",
                "SELECT bogus, command_id, bogus FROM Commands WHERE command_id < 5",
                "SELECT * FROM Commands",
                "SELECT raw_command FROM Commands WHERE command_id > 100000"]:
        ng, rg, _ = g.select(sql)
        nr, rr = r.select(sql)
        assert ng == nr and rg == rr, sql


def test_printed_output_sample_queries(pkg, tmp_path):
    """identical printed output for sample-queries-FULL (INSERT + DELETE included)"""
    if not os.path.exists(support.REF_DUMP):
        pytest.skip("oracle/_ref/qpe_ref_dump not built")
    qf = tmp_path / "queries.txt"
    qf.write_text(SAMPLE_QUERIES_FULL)
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    csv_ref = support.scratch_copy(CSV_2K, tmp_path / "a")
    csv_gpu = support.scratch_copy(CSV_2K, tmp_path / "b")
    want = support.ref_dump(csv_ref, str(qf), max_rows=20)
    g = pkg.Engine.from_csv(csv_gpu)
    got = ""
    for stmt in SAMPLE_QUERIES_FULL.split(";"):
        stmt = stmt.lstrip()
        if stmt:
            got += g.run(stmt, 20)
    g.close()
    assert normalise(got) == normalise(want)
    # the CSV side effects (append, then full rewrite) are byte-identical too
    assert open(csv_gpu, "rb").read() == open(csv_ref, "rb").read()


def test_insert_delete_sequence(pkg, tmp_path):
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    g = pkg.Engine.from_csv(support.scratch_copy(CSV_2K, tmp_path / "a"))
    r = Ref(support.scratch_copy(CSV_2K, tmp_path / "b"))
    stmts = [
        'DELETE FROM Commands WHERE risk_level = 3',
        'INSERT INTO Commands VALUES (777, "a very long command %s", "longcmd", "bash", 2, "2026-01-01T00:00:00.000Z", "TRUE", "/opt/some/where/deep/down/the/tree/of/directories/x", 1001, "student1001", "labpc-01", 5)' % ("x" * 150),
        'INSERT INTO Commands VALUES (5, "dup id", "dup", "zsh", 0, "2026-01-02T00:00:00.000Z", "1", "/tmp", 1001, "student1001", "labpc-02", 5)',
        'DELETE FROM Commands WHERE user_id = 1002 OR (shell_type = "fish" AND exit_code != 0)',
        'INSERT INTO Commands VALUES (0, "rejected", "r", "sh", 0, "t", "0", "/", 1, "u", "h", 1)',
    ]
    checks = ['SELECT * FROM Commands WHERE risk_level = 5',
              'SELECT command_id, raw_command FROM Commands WHERE user_id = 1001',
              'SELECT command_id FROM Commands WHERE command_id <= 800 AND command_id >= 1900',
              'SELECT command_id, working_directory FROM Commands WHERE (risk_level > 3)']
    import io, contextlib
    for s in stmts:
        out = g.run(s, 20)
        r.run(s, 20)
        assert g.num_rows == r.num_rows, s
        for q in checks:
            ng, rg, _ = g.select(q)
            nr, rr = r.select(q)
            assert rg == rr, (s, q)
    g.close()
    r.close()
