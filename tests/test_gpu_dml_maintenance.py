"""INSERT / DELETE maintenance without a re-sort (csrc/index.cu, csrc/engine.cu): the reference inserts / deletes ONE
entry per index (engine/bplus.c:723-740, :1022-1051, called from executeEngine-serial.c:599-614, :661-665); here an
INSERT shifts the index tail by one slot behind a K3 probe and a DELETE filters + renumbers the entries in one ordered
pass.  After every statement of a random sequence on a 1 M-row table the columns must equal a numpy restatement of the
statement (stable compaction / append), every index must equal the closed form of SURVEY App. A.3 -- row ids by
(key ASCENDING, table position DESCENDING) -- computed from the column itself, and an index-path SELECT must equal the
oracle's.  (The same statements against the compiled reference at 2 000 rows: test_gpu_parity_ref.py.)"""
import numpy as np
import pytest

import support

pytestmark = pytest.mark.gpu

N = 1_000_000
INDEXES = (("command_id", 0), ("user_id", 1), ("risk_level", 1), ("exit_code", 1), ("sudo_used", 3))
NUM = ["command_id", "user_id", "risk_level", "exit_code", "sudo_used"]


@pytest.fixture(scope="module")
def pkg():
    return support.load_pkg()


def expected_order(keys):
    """row ids by (key ASC, position DESC): a stable sort of the reversed column"""
    n = len(keys)
    rev = keys[::-1]
    order = np.argsort(rev, kind="stable")
    return (n - 1 - order).astype(np.uint32)


def check_state(eng, cols, what):
    n = eng.num_rows
    assert n == len(cols["command_id"]), what
    for c in NUM:
        got = eng.fetch_column(c)
        assert np.array_equal(got, cols[c]), (what, c)
    for attr in ("command_id", "user_id", "risk_level", "exit_code"):
        perm = eng.index_slice(attr, 0, n)
        assert np.array_equal(perm, expected_order(cols[attr])), (what, attr)


def test_random_dml_sequence_keeps_columns_and_indexes(pkg):
    rng = np.random.default_rng(20261018)
    eng = pkg.Engine.from_synth(N, indexes=INDEXES)      # every column resident (INSERT needs them all)
    cols = {c: eng.fetch_column(c).copy() for c in NUM}
    # probe every index once: built (not dirty) before the first statement, so the statements MAINTAIN them
    for attr in ("command_id", "user_id", "risk_level", "exit_code"):
        eng.index_slice(attr, 0, 1)
    check_state(eng, cols, "initial")
    next_id = int(cols["command_id"].max()) + 1
    for step in range(8):
        if step % 2 == 0:
            # DELETE with a random predicate over numeric columns (full scan, stable compaction)
            uid = int(rng.integers(1000, 3000))
            risk = int(rng.integers(1, 6))
            lo = int(rng.integers(0, N))
            sql = f"DELETE FROM Commands WHERE (user_id = {uid}) OR (risk_level = {risk} AND command_id > {lo})"
            keep = ~((cols["user_id"] == uid) | ((cols["risk_level"] == risk) & (cols["command_id"] > lo)))
            out = eng.run(sql, 5)
            deleted = int((~keep).sum())
            assert f"Rows affected: {deleted}" in out, (sql, out)
            cols = {c: v[keep] for c, v in cols.items()}
        else:
            # a few INSERTs: new ids, duplicate ids, duplicate user / risk / exit keys, extreme keys
            for k in range(3):
                cid = next_id if k != 1 else int(cols["command_id"][int(rng.integers(0, len(cols["command_id"])))])
                next_id += 1
                uid = int(rng.choice([1001, 999999, int(cols["user_id"][0])]))
                risk = int(rng.integers(1, 6))
                code = int(rng.choice([0, 1, 127, 255]))
                sudo = bool(rng.integers(0, 2))
                sql = (f'INSERT INTO Commands VALUES ({cid}, "cmd {step} {k}", "cmd", "bash", {code}, '
                       f'"2026-01-01T00:00:00.000Z", "{"TRUE" if sudo else "FALSE"}", "/tmp", {uid}, "u{uid}", "h", {risk})')
                out = eng.run(sql, 5)
                assert "Insert successful" in out, (sql, out)
                row = {"command_id": cid, "user_id": uid, "risk_level": risk, "exit_code": code, "sudo_used": int(sudo)}
                cols = {c: np.append(v, np.array([row[c]], dtype=v.dtype)) for c, v in cols.items()}
        check_state(eng, cols, f"after step {step}")
    # an index-path SELECT after all that, against the oracle over the final columns
    host = {c: eng.fetch_column(c) for c in NUM}
    o = support.Oracle.from_columns(host)
    for where in ("user_id = 1001", "risk_level > 4 AND exit_code = 0", "command_id <= 50 OR user_id = 999999"):
        ids, st = eng.select_ids(f"SELECT command_id FROM Commands WHERE {where}")
        want, used = o.select_ids(where, INDEXES)
        assert used and st["path"] == 1 and np.array_equal(ids, want), where
    eng.close()


def test_delete_everything_then_insert(pkg):
    eng = pkg.Engine.from_synth(50_000, indexes=INDEXES)
    eng.index_slice("user_id", 0, 1)
    out = eng.run("DELETE FROM Commands WHERE command_id >= 0", 5)
    assert "Rows affected: 50000" in out and eng.num_rows == 0
    ids, _ = eng.select_ids("SELECT command_id FROM Commands WHERE user_id = 1001")
    assert len(ids) == 0
    out = eng.run('INSERT INTO Commands VALUES (7, "x", "x", "sh", 0, "t", "FALSE", "/", 1001, "u", "h", 1)', 5)
    assert "Insert successful" in out and eng.num_rows == 1
    ids, st = eng.select_ids("SELECT command_id FROM Commands WHERE user_id = 1001")
    assert st["path"] == 1 and ids.tolist() == [0]
    eng.close()
