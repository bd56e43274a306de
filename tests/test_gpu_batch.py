"""K9 (query batch): n SELECTs whose full-scan members share passes over the table must give, result by
result, exactly what n single SELECTs give (which the other parity tests pin to the reference / oracle):
same path rule (index-path queries keep the index order, duplicates included), same row order."""
import numpy as np
import pytest

import support

pytestmark = pytest.mark.gpu

N = 2_000_003
COLS = ["command_id", "sudo_used", "risk_level", "exit_code", "user_id", "shell_type", "host_name", "base_command"]
IDX = (("command_id", 0), ("user_id", 1), ("risk_level", 1))
STATEMENTS = [
    "SELECT command_id FROM Commands WHERE (command_id < 700000) AND (sudo_used = FALSE OR risk_level > 3)",
    'SELECT command_id FROM Commands WHERE (shell_type = "bash" OR host_name = "labpc-01")',
    "SELECT command_id FROM Commands WHERE risk_level > 3 AND exit_code != 0",             # index path (risk_level)
    "SELECT command_id FROM Commands WHERE (exit_code = 127)",
    "SELECT command_id FROM Commands",                                                     # no WHERE: every row
    "SELECT command_id FROM Commands WHERE (risk_level > 100)",                            # nobody
    'SELECT command_id FROM Commands WHERE (base_command >= "git") AND (base_command < "ls")',
    "SELECT command_id FROM Commands WHERE user_id = 2450 OR (exit_code = 1)",            # index path (user_id)
    "SELECT command_id FROM Commands WHERE (sudo_used = TRUE)",
    'SELECT command_id FROM Commands WHERE (host_name != "labpc-01") AND ((risk_level >= 2) OR (exit_code != 0))',
    "SELECT command_id FROM Commands WHERE (command_id >= 1999990)",
    "SELECT command_id FROM Commands WHERE (nosuchcolumn = 3) OR (risk_level = 5)",
    "SELECT command_id FROM Commands WHERE (user_id < 1500) AND (shell_type != \"sh\")",
]


@pytest.fixture(scope="module")
def eng():
    pkg = support.load_pkg()
    e = pkg.Engine.from_synth(N, columns=COLS, indexes=IDX)
    yield e
    e.close()


@pytest.fixture(scope="module")
def oracle(eng):
    """the CPU restatement over the same columns: the checker, never the thing under test"""
    return support.Oracle.from_columns({c: eng.fetch_column(c) for c in COLS})


def test_batch_equals_the_oracle(eng, oracle):
    """every result of the batch against Oracle.select_ids (path rule, order, duplicates), not against the engine"""
    batch, st = eng.select_ids_batch(STATEMENTS)
    assert len(batch) == len(STATEMENTS)
    total = 0
    for s, got in zip(STATEMENTS, batch):
        where = s.split("WHERE", 1)[1] if "WHERE" in s else None
        if where is None:
            want = np.arange(N, dtype=np.uint32)
        else:
            want, _ = oracle.select_ids(where, IDX)
        assert got.shape == want.shape and np.array_equal(got, want), s
        total += len(want)
    assert st["matches"] == total


def test_batch_equals_single_queries(eng):
    singles = [eng.select_ids(s)[0] for s in STATEMENTS]
    batch, st = eng.select_ids_batch(STATEMENTS)
    assert len(batch) == len(STATEMENTS)
    for s, a, b in zip(STATEMENTS, singles, batch):
        assert np.array_equal(a, b), s
    assert st["matches"] == sum(len(a) for a in singles)


def test_same_columns_share_a_pass(eng, oracle):
    stmts = [f"SELECT command_id FROM Commands WHERE (exit_code = {k}) AND (sudo_used = {b})"
             for k, b in zip([0, 1, 2, 126, 127, 130, 137, 255, 1, 0, 2], ["TRUE", "FALSE"] * 6)]
    batch, st = eng.select_ids_batch(stmts)
    for s, b in zip(stmts, batch):
        assert np.array_equal(oracle.scan(s.split("WHERE", 1)[1]), b), s
    # 11 queries over the same two columns: 2 scan passes (8 + 3 programs) + at most one compaction per query
    assert st["launches"] <= 2 + 11


def test_batch_edge_cases(eng):
    assert eng.select_ids_batch([])[0] == []
    one, _ = eng.select_ids_batch([STATEMENTS[0]])
    assert np.array_equal(one[0], eng.select_ids(STATEMENTS[0])[0])
    same, _ = eng.select_ids_batch([STATEMENTS[3]] * 9)          # more than one pass of the same program
    want = eng.select_ids(STATEMENTS[3])[0]
    assert all(np.array_equal(x, want) for x in same)
