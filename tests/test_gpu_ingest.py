"""K6 (CSV parsed on the device) against the two CPU statements of the serial loader's rules
(engine/serial/buildEngine-serial.c:70-221): the host loader of the library (QPE_INGEST=host) and
the oracle's loader, on deliberately nasty CSV text."""
import os
import random

import numpy as np
import pytest

import support
from support import Oracle

pytestmark = pytest.mark.gpu

HEADER = "command_id,raw_command,base_command,shell_type,exit_code,timestamp,sudo_used,working_directory,user_id,user_name,host_name,risk_level"


def nasty_rows(seed=7, n=400):
    rng = random.Random(seed)
    ids = ["0", "1", "42", " 17", "+9", "-3", "18446744073709551615", "99999999999999999999999", "12abc", "", "007"]
    ints = ["0", "1", "-1", " 5", "+7", "2147483647", "-2147483648", "4294967296", "99999999999999999999", "x", "", "3.9"]
    bools = ["true", "TRUE", "True", "false", "1", "0", "yes", "", "tRuE", "11"]
    texts = ["ls -la", '"echo ""Hello, world"""', '"a,b,c"', '"quoted"tail', "", "plain text", '"unterminated, quote',
             "x" * 130, '""', '"""', "tab\there", "semi;colon", "ünïcödé", '"multi ""q"" , commas"']
    rows = []
    for i in range(n):
        kind = rng.random()
        if kind < 0.03:
            rows.append("")                                   # blank line -> all-zero row
            continue
        f = [rng.choice(ids), rng.choice(texts), rng.choice(texts)[:90], rng.choice(["bash", "zsh", '"fish"', "sh", ""]),
             rng.choice(ints), "2026-0%d-1%dT00:00:0%d.000Z" % (rng.randint(1, 9), rng.randint(0, 9), rng.randint(0, 9)),
             rng.choice(bools), rng.choice(["/tmp", '"/home/a,b"', "/var/log", ""]), rng.choice(ints),
             rng.choice(["student1001", '"stu""dent"', ""]), rng.choice(["labpc-01", "cs-lab-02", "h" * 120]), rng.choice(ints)]
        if kind < 0.10:
            f = f[:rng.randint(1, 11)]                        # missing trailing fields
        rows.append(",".join(f))
    return rows


def write_csv(path, rows, eol="\r\n", final_newline=True):
    text = HEADER + eol + eol.join(rows) + (eol if final_newline else "")
    open(path, "wb").write(text.encode("utf-8"))


def columns_of(eng, pkg):
    return {c: eng.fetch_column(c) for c in pkg.COLUMNS}


def assert_same_table(pkg, a, b):
    assert a.num_rows == b.num_rows
    for c in pkg.COLUMNS:
        x, y = a.fetch_column(c), b.fetch_column(c)
        if x.ndim == 2:
            w = min(x.shape[1], y.shape[1])
            assert not x[:, w:].any() and not y[:, w:].any(), c
            x, y = x[:, :w], y[:, :w]
        assert np.array_equal(x, y), c


def assert_matches_oracle(pkg, eng, csv):
    o = Oracle.from_csv(csv)
    assert eng.num_rows == o.num_rows
    n = eng.num_rows
    for c in pkg.COLUMNS:
        col = eng.fetch_column(c)
        for i in range(n):
            want = o.cell(i, c)
            if col.ndim == 2:
                got = bytes(col[i]).split(b"\0", 1)[0].decode("utf-8", errors="replace")
            elif c == "sudo_used":
                got = "true" if col[i] else "false"
            else:
                got = str(int(col[i]))
            assert got == want, (c, i, got, want)
    o.close()


@pytest.mark.parametrize("eol,final_newline", [("\r\n", True), ("\n", True), ("\n", False)])
def test_device_ingest_equals_host_and_oracle(tmp_path, monkeypatch, eol, final_newline):
    pkg = support.load_pkg()
    csv = str(tmp_path / "nasty.csv")
    write_csv(csv, nasty_rows(), eol, final_newline)
    monkeypatch.delenv("QPE_INGEST", raising=False)
    dev = pkg.Engine.from_csv(csv, indexes=())
    monkeypatch.setenv("QPE_INGEST", "host")
    host = pkg.Engine.from_csv(csv, indexes=())
    monkeypatch.delenv("QPE_INGEST", raising=False)
    assert_same_table(pkg, dev, host)
    assert_matches_oracle(pkg, dev, csv)
    dev.close()
    host.close()


def test_embedded_nul_and_header_only(tmp_path):
    pkg = support.load_pkg()
    csv = str(tmp_path / "nul.csv")
    open(csv, "wb").write((HEADER + "\n").encode() + b"5,abc\0def,ls,bash,1,t,true,/tmp,7,u,h,2\n6,x,y,z,3,t,1,/,8,u,h,4\n")
    eng = pkg.Engine.from_csv(csv, indexes=())
    assert_matches_oracle(pkg, eng, csv)
    eng.close()
    csv2 = str(tmp_path / "empty.csv")
    open(csv2, "w").write(HEADER + "\n")
    eng = pkg.Engine.from_csv(csv2)
    assert eng.num_rows == 0
    assert eng.select("SELECT * FROM Commands WHERE risk_level = 5")[1] == []
    eng.close()


def test_overlong_line_falls_back_to_the_exact_chunking(tmp_path):
    """a physical line of >= 1023 characters is several fgets(1024) chunks = several rows in the reference"""
    pkg = support.load_pkg()
    csv = str(tmp_path / "long.csv")
    rows = ["1,ls,ls,bash,0,t,true,/tmp,1,u,h,1", "2," + "y" * 1500 + ",ls,bash,0,t,false,/tmp,2,u,h,2",
            "3,pwd,pwd,zsh,0,t,1,/,3,u,h,3"]
    write_csv(csv, rows, "\n", True)
    eng = pkg.Engine.from_csv(csv, indexes=())
    assert_matches_oracle(pkg, eng, csv)
    assert eng.num_rows == 4   # the long line became two "rows"
    eng.close()


def test_device_ingest_on_the_golden_csv(tmp_path):
    pkg = support.load_pkg()
    eng = pkg.Engine.from_csv(support.scratch_copy(support.CSV_2K, tmp_path))
    assert_matches_oracle(pkg, eng, support.CSV_2K)
    eng.close()
