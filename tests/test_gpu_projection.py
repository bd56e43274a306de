"""K7 (csrc/format.cu): the SELECT projection rendered on the device into fixed-width text slots must give
exactly the cell text of get_attribute_string_value (engine/serial/executeEngine-serial.c:216-248):
"%llu" / "%d" / "true"|"false" / the string itself / "NULL" for an unknown column.  Checked against the
oracle's cell renderer and -- when it was compiled -- the reference itself, on CSV text with extreme
numbers (UINT64_MAX, INT_MIN, overflowing literals), empty strings and cells that fill their column."""
import os

import numpy as np
import pytest

import support
from support import Oracle
from test_gpu_ingest import nasty_rows, write_csv

pytestmark = pytest.mark.gpu

SELECTS = [
    ("SELECT * FROM Commands", None),
    ("SELECT command_id, exit_code, sudo_used, user_id, risk_level FROM Commands WHERE (exit_code != 77)", None),
    ("SELECT raw_command, command_id, raw_command, nosuchcolumn, host_name FROM Commands WHERE (sudo_used = FALSE)", None),
    ("SELECT risk_level FROM Commands WHERE (risk_level < 0)", None),
    ("SELECT user_name, timestamp FROM Commands WHERE (command_id > 18446744073709551614)", None),
]


# sizeof the text members of `record` (include/logType.h:11-24).  A CSV field at least that long is stored
# UNTERMINATED by the reference's strncpy (buildEngine-serial.c:175-214) and every later read runs on into
# the following members (undefined behaviour); the engine and the oracle keep the first sizeof-1 bytes.
FIELD_BYTES = {"raw_command": 512, "base_command": 100, "shell_type": 20, "timestamp": 30, "working_directory": 200,
               "user_name": 50, "host_name": 100}


def _assert_same_cells(got, want, names, sql, who):
    assert len(got) == len(want), (sql, who)
    for i, (a, b) in enumerate(zip(got, want)):
        if a != b:
            bad = [k for k in range(len(names)) if a[k] != b[k] and not (
                who == "reference" and len(a[k]) == FIELD_BYTES.get(names[k], 0) - 1 and b[k].startswith(a[k]))]
            if not bad:
                continue
            j = bad[0]
            raise AssertionError(f"{sql}: row {i} column {names[j]}: ours {a[j]!r} (len {len(a[j])}) != {who} "
                                 f"{b[j]!r} (len {len(b[j])})")


@pytest.fixture(scope="module")
def table(tmp_path_factory):
    d = tmp_path_factory.mktemp("proj")
    csv = str(d / "nasty.csv")
    rows = nasty_rows(seed=11, n=700)
    rows += ["18446744073709551615,max id,max,bash,-2147483648,2026-01-01T00:00:00.000Z,true,/tmp,2147483647,u,h,-1",
             "0,zero,zero,sh,0,2026-01-01T00:00:00.000Z,false,/,0,,,0"]
    write_csv(csv, rows)
    return csv


def test_projection_text_equals_oracle_and_reference(table):
    pkg = support.load_pkg()
    eng = pkg.Engine.from_csv(table, indexes=())
    o = Oracle.from_csv(table)
    ref = support.Ref(table, num_indexes=0) if support.Ref.available() else None
    try:
        for sql, _ in SELECTS:
            names, rows, _ = eng.select(sql)
            ids, _ = eng.select_ids(sql, force_scan=True)
            assert len(rows) == len(ids)
            want = [[o.cell(int(i), a) if a in pkg.COLUMNS else "NULL" for a in names] for i in ids]
            _assert_same_cells(rows, want, names, sql, "oracle")
            if ref is not None:
                rnames, rrows = ref.select(sql)
                assert names == rnames
                _assert_same_cells(rows, rrows, names, sql, "reference")
    finally:
        eng.close()
        o.close()
        if ref is not None:
            ref.close()


def test_projection_large_result_multi_chunk_pointers():
    """1.2 M rows x 12 columns: the row-pointer layout is split over host threads and every cell must still
    be the text of its own (row, column)."""
    pkg = support.load_pkg()
    n = 1_200_000
    eng = pkg.Engine.from_synth(n, columns=pkg.COLUMNS)
    lib = pkg.load_library()
    res = lib.qpe_sql_select(eng._h, b"SELECT * FROM Commands WHERE (risk_level >= 0)")
    try:
        r = res.contents
        assert r.success and r.numRecords == n and r.numColumns == 12
        cid = eng.fetch_column("command_id")
        uid = eng.fetch_column("user_id")
        sudo = eng.fetch_column("sudo_used")
        host = eng.fetch_column("host_name")
        rng = np.random.default_rng(5)
        for i in [0, 1, n - 1, n // 2] + list(rng.integers(0, n, size=2000)):
            i = int(i)
            row = r.data[i]
            assert row[0] == str(int(cid[i])).encode()
            assert row[8] == str(int(uid[i])).encode()
            assert row[6] == (b"true" if sudo[i] else b"false")
            assert row[10] == bytes(host[i]).split(b"\0", 1)[0]
    finally:
        lib.freeResultSet(res)
        eng.close()


# ---- K8: the data file rewritten from the device columns after DELETE -------------------------------------
def _expected_csv(eng, pkg):
    cols = {c: eng.fetch_column(c) for c in pkg.COLUMNS}
    n = eng.num_rows

    def text(c, i):
        return bytes(cols[c][i]).split(b"\0", 1)[0]

    out = []
    for i in range(n):
        out.append(b",".join([
            str(int(cols["command_id"][i])).encode(), text("raw_command", i), text("base_command", i),
            text("shell_type", i), str(int(cols["exit_code"][i])).encode(), text("timestamp", i),
            b"1" if cols["sudo_used"][i] else b"0", text("working_directory", i), str(int(cols["user_id"][i])).encode(),
            text("user_name", i), text("host_name", i), str(int(cols["risk_level"][i])).encode()]) + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("long_rows", [False, True])
def test_delete_rewrites_csv_from_device_text(tmp_path, long_rows):
    """executeQueryDeleteSerial rewrites the whole file with one fprintf per row (:683-706); K8 renders the same
    bytes on the device.  long_rows: 300 rows of ~700 bytes each, so whole CTAs overflow the shared-memory
    stage and take the direct-store path."""
    pkg = support.load_pkg()
    csv = str(tmp_path / "t.csv")
    rows = nasty_rows(seed=23, n=900)
    if long_rows:
        rows = [f'{i + 1},"{"r" * 480} {i}",{"b" * 90},zsh,{i - 150},2026-03-01T00:00:00.000Z,true,/{"w" * 150},{i},'
                f'user{i},host-{i},{i % 5}' for i in range(300)] + rows
    write_csv(csv, rows)
    eng = pkg.Engine.from_csv(csv, indexes=())
    try:
        n0 = eng.num_rows
        out = eng.run("DELETE FROM Commands WHERE command_id = 4000000000000", 5)      # matches nothing
        assert "Rows affected: 0" in out
        assert open(csv, "rb").read() == _expected_csv(eng, pkg)
        out = eng.run("DELETE FROM Commands WHERE (risk_level > 0) AND (sudo_used = TRUE)", 5)
        assert eng.num_rows < n0
        assert open(csv, "rb").read() == _expected_csv(eng, pkg)
    finally:
        eng.close()
