"""GPU parity at scale and at the edges, through the C-ABI:
  * synthetic tables vs the oracle (bit-exact row ids, both paths),
  * golden vectors of the compiled reference (no /root/reference needed),
  * CSV round trip (device generator -> CSV writer -> loader) incl. quoted fields,
  * batched B+ probes vs the closed-form leaf-chain order,
  * size-independent properties at 100 M rows."""
import json
import os

import numpy as np
import pytest

import support
from support import CSV_2K, GOLDEN, Oracle

pytestmark = pytest.mark.gpu

IDX = (("command_id", 0), ("user_id", 1), ("risk_level", 1), ("exit_code", 1), ("sudo_used", 3))
NUM_COLS = ["command_id", "sudo_used", "risk_level", "exit_code", "user_id"]
SCAN_COLS = NUM_COLS + ["shell_type", "host_name", "base_command", "user_name", "timestamp", "working_directory"]

WHERES = [
    '(command_id < {h}) AND (sudo_used = FALSE OR risk_level > 3)',
    '(command_id < {q}) AND (shell_type = "bash" OR host_name = "labpc-01")',
    '(command_id < {h}) AND (risk_level >= 2 OR exit_code != 0) AND (user_id < 2000 OR shell_type != "sh")',
    'risk_level > 3 AND shell_type != "bash"',
    'user_id = 1001 OR (exit_code = 127)',
    'exit_code > 126 AND risk_level < 2',
    'risk_level = 5 OR user_id = 1003',
    'command_id >= {q} AND command_id <= {h}',
    'timestamp > "2026-06" AND user_name <= "student1100" AND working_directory != "/tmp"',
    '((risk_level = 1 OR risk_level = 2) AND (shell_type = "zsh")) OR (exit_code = 130)',
    'base_command = "rm" AND (sudo_used = TRUE OR (risk_level = 5 AND host_name > "labpc-05"))',
    'host_name = "personal-laptop-with-a-name-longer-than-the-column"',
    'host_name < "personal-laptop-with-a-name-longer-than-the-column"',
    'shell_type = "bash"',
    'sudo_used != FALSE',
    'nosuch = 1 OR risk_level = 4',
]


@pytest.fixture(scope="module")
def pkg():
    return support.load_pkg()


@pytest.fixture(scope="module")
def table_2m(pkg):
    n = 2_000_003
    eng = pkg.Engine.from_synth(n, columns=SCAN_COLS, indexes=IDX)
    host = {c: eng.fetch_column(c) for c in SCAN_COLS}
    o = Oracle.from_columns(host)
    yield eng, o, n
    eng.close()


@pytest.mark.parametrize("where", WHERES)
def test_select_ids_equal_oracle_2m(table_2m, where):
    eng, o, n = table_2m
    w = where.format(h=n // 2, q=n // 4)
    sql = f"SELECT command_id FROM Commands WHERE {w}"
    ids, st = eng.select_ids(sql)
    want, used_index = o.select_ids(w, IDX)
    assert st["path"] == (1 if used_index else 0)
    assert np.array_equal(ids, want)
    forced, st2 = eng.select_ids(sql, force_scan=True)
    assert st2["path"] == 0 and np.array_equal(forced, o.scan(w))
    cnt, _, _ = eng.select_ids_device(sql, force_scan=True, count_only=True)
    assert cnt == len(forced)


def test_select_cells_equal_oracle(table_2m):
    eng, o, n = table_2m
    cols = ["command_id", "shell_type", "sudo_used", "exit_code", "timestamp", "nosuch"]
    w = 'risk_level = 5 AND host_name = "cs-lab-02"'
    names, rows, _ = eng.select(f"SELECT {', '.join(cols)} FROM Commands WHERE {w}")
    ids, _ = o.select_ids(w, IDX)
    assert names == cols
    assert rows == o.rows(ids, cols)


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 255, 256, 257, 4095, 4096, 4097, 65535, 65536, 65537, 131073, 600001])
def test_ragged_sizes(pkg, n):
    eng = pkg.Engine.from_synth(max(n, 1), n_rows=n, columns=NUM_COLS + ["shell_type"])
    if n:
        o = Oracle.from_columns({c: eng.fetch_column(c) for c in NUM_COLS + ["shell_type"]})
    for w in ['risk_level >= 1', '(command_id < 5) OR (risk_level > 2 AND shell_type = "bash")', 'risk_level > 99']:
        ids, _ = eng.select_ids(f"SELECT command_id FROM Commands WHERE {w}", force_scan=True)
        want = o.scan(w) if n else np.zeros(0, dtype=np.uint32)
        assert np.array_equal(ids, want), (n, w)
    ids, _ = eng.select_ids("SELECT command_id FROM Commands")  # NULL WHERE: every row, in order
    assert np.array_equal(ids, np.arange(n, dtype=np.uint32))
    eng.close()


@pytest.mark.parametrize("tile,stages", [(256, 2), (256, 16), (512, 4), (512, 1), (1024, 3), (1024, 5)])
def test_tile_geometries(table_2m, tile, stages):
    eng, o, n = table_2m
    w = '(command_id < 1500000) AND (risk_level >= 2 OR exit_code != 0) AND (user_id < 2000 OR shell_type != "sh")'
    eng.set_tile(tile, stages)
    try:
        ids, st = eng.select_ids(f"SELECT command_id FROM Commands WHERE {w}", force_scan=True)
        assert st["tile_rows"] == tile and st["stages"] == stages
        assert np.array_equal(ids, o.scan(w))
    finally:
        eng.set_tile(0, 0)


def test_golden_vectors_of_the_compiled_reference(pkg, tmp_path):
    """no /root/reference and no oracle/_ref needed: outputs of the reference, committed as fixtures"""
    eng = pkg.Engine.from_csv(support.scratch_copy(CSV_2K, tmp_path))
    for e in json.load(open(os.path.join(GOLDEN, "probe_results_2k.json"))):
        sql = f"SELECT command_id FROM Commands WHERE {e['where']}"
        _, rows, _ = eng.select(sql)
        assert [int(r[0]) for r in rows] == e["select_command_ids"], e["where"]
        ids, _ = eng.select_ids(sql, force_scan=True)
        assert ids.tolist() == e["scan_positions"], e["where"]
    order = json.load(open(os.path.join(GOLDEN, "index_order_2k.json")))
    cid = eng.fetch_column("command_id")
    for attr, want in order.items():
        perm = eng.index_slice(attr, 0, eng.num_rows)
        assert cid[perm].tolist() == want, attr
    got = "".join(eng.run(s.lstrip(), 20) for s in support.SAMPLE_QUERIES_FULL.split(";") if s.strip())
    assert support.normalise(got) == open(os.path.join(GOLDEN, "sample_full_2k.out")).read()
    eng.close()


def test_csv_round_trip_with_quoted_fields(pkg, tmp_path):
    """device generator -> qpe_gpu_write_csv (QUOTE_MINIMAL, \\r\\n) -> serial-rules loader -> same columns"""
    n = 50_000
    a = pkg.Engine.from_synth(n, indexes=IDX)
    csv = str(tmp_path / "synth_50k.csv")
    a.write_csv(csv)
    text = open(csv, "rb").read()
    assert b'"echo ""Hello, world"""' in text  # comma + quotes inside a field: exercised
    b = pkg.Engine.from_csv(csv)
    assert b.num_rows == n
    for c in pkg.COLUMNS:
        x, y = a.fetch_column(c), b.fetch_column(c)
        if x.ndim == 2:  # text: widths may differ (the loader sizes columns from the data)
            w = min(x.shape[1], y.shape[1])
            assert not x[:, w:].any() and not y[:, w:].any()
            x, y = x[:, :w], y[:, :w]
        assert np.array_equal(x, y), c
    # and the compiled reference reads the same file the same way
    if support.Ref.available():
        r = support.Ref(csv)
        for w in ['raw_command = "echo \\"Hello, world\\""', 'risk_level = 5 AND shell_type = "fish"', 'user_id = 1002']:
            sql = f"SELECT command_id, raw_command, timestamp FROM Commands WHERE {w}"
            assert b.select(sql)[1] == r.select(sql)[1], w
        r.close()
    a.close()
    b.close()


def test_probe_batch_matches_closed_form(table_2m):
    eng, o, n = table_2m
    rng = np.random.default_rng(12345)
    # unique keys: command_id
    lo = rng.integers(0, int(n * 1.1), size=200_000).astype(np.uint64)
    ln = rng.choice([0, 1, 16, 256, 4096], size=lo.size).astype(np.uint64)
    first, count, st = eng.probe_batch("command_id", lo, lo + ln)
    want_first = np.minimum(lo, n)
    want_count = np.clip(np.minimum(lo + ln + 1, n).astype(np.int64) - want_first.astype(np.int64), 0, None)
    assert np.array_equal(first, want_first.astype(np.uint32))
    assert np.array_equal(count, want_count.astype(np.uint32))
    # heavy duplicates: user_id, against the oracle's (key ASC, position DESC) order
    perm = o.index_order("user_id")
    keys = eng.fetch_column("user_id")[perm]
    qlo = rng.integers(990, 3010, size=100_000).astype(np.int32)
    qhi = qlo + rng.choice([0, 0, 1, 5, 100], size=qlo.size).astype(np.int32)
    first, count, _ = eng.probe_batch("user_id", qlo, qhi)
    wf = np.searchsorted(keys, qlo, side="left")
    wc = np.searchsorted(keys, qhi, side="right") - wf
    assert np.array_equal(first, wf.astype(np.uint32)) and np.array_equal(count, wc.astype(np.uint32))
    for q in range(0, 2000, 97):  # the slices hold the reference's row order
        assert np.array_equal(eng.index_slice("user_id", int(first[q]), int(count[q])),
                              perm[first[q]:first[q] + count[q]])
    # inverted range and extremes
    f2, c2, _ = eng.probe_batch("user_id", np.array([1500, -2**31, 2**31 - 1], dtype=np.int32),
                                np.array([1400, 2**31 - 1, 2**31 - 1], dtype=np.int32))
    assert c2.tolist() == [0, n, 0] and f2[1] == 0


def test_properties_at_100m_rows(pkg):
    """size-independent checks at a size the oracle cannot finish quickly"""
    n = 100_000_000
    eng = pkg.Engine.from_synth(n, columns=NUM_COLS)
    k = 37_123_457
    ids, st = eng.select_ids(f"SELECT command_id FROM Commands WHERE (command_id < {k})", force_scan=True)
    assert len(ids) == k and ids[0] == 0 and ids[-1] == k - 1 and np.array_equal(ids[::1001], np.arange(0, k, 1001))
    ca, _, _ = eng.select_ids_device("SELECT command_id FROM Commands WHERE (risk_level > 2)", force_scan=True, count_only=True)
    cb, _, _ = eng.select_ids_device("SELECT command_id FROM Commands WHERE (risk_level <= 2)", force_scan=True, count_only=True)
    assert ca + cb == n and 0 < ca < n
    w = '(command_id >= 50000000) AND (sudo_used = TRUE OR exit_code != 0)'
    ids, st = eng.select_ids(f"SELECT command_id FROM Commands WHERE {w}", force_scan=True)
    assert np.all(np.diff(ids.astype(np.int64)) > 0) and ids[0] >= 50_000_000      # strictly increasing = table order
    cnt, _, _ = eng.select_ids_device(f"SELECT command_id FROM Commands WHERE {w}", force_scan=True, count_only=True)
    assert cnt == len(ids)
    # a 1 M-row window of the same query, bit-exact against the oracle
    a = 73_000_000
    win = {c: eng.fetch_column(c, a, 1_000_000) for c in NUM_COLS}
    o = Oracle.from_columns(win)
    want = o.scan(w).astype(np.int64) + a
    got = ids[(ids >= a) & (ids < a + 1_000_000)]
    assert np.array_equal(got.astype(np.int64), want)
    # De Morgan-style identity on counts: |A or B| = |A| + |B| - |A and B|
    cA, _, _ = eng.select_ids_device("SELECT c FROM t WHERE (sudo_used = TRUE)", force_scan=True, count_only=True)
    cB, _, _ = eng.select_ids_device("SELECT c FROM t WHERE (exit_code != 0)", force_scan=True, count_only=True)
    cAB, _, _ = eng.select_ids_device("SELECT c FROM t WHERE (sudo_used = TRUE) AND (exit_code != 0)", force_scan=True, count_only=True)
    cAoB, _, _ = eng.select_ids_device("SELECT c FROM t WHERE (sudo_used = TRUE) OR (exit_code != 0)", force_scan=True, count_only=True)
    assert cAoB == cA + cB - cAB
    eng.close()


def test_properties_at_1b_rows(pkg):
    """BASELINE configs[4] at its full size (1 B rows, the bench's table): size-independent properties plus
    1 M-row windows checked bit-exactly against the oracle; the dense query runs twice so that the second scan
    uses the 8-compaction-warp kernel (chosen from the first scan's selectivity)."""
    n = 1_000_000_000
    cols = ["command_id", "sudo_used", "risk_level"]
    eng = pkg.Engine.from_synth(n, columns=cols)
    try:
        k = n // 100
        w = f"(command_id < {k}) AND (sudo_used = FALSE OR risk_level > 3)"     # the bench's QN at 1 %
        ids, st = eng.select_ids(f"SELECT command_id FROM Commands WHERE {w}", force_scan=True)
        cnt, _, _ = eng.select_ids_device(f"SELECT command_id FROM Commands WHERE {w}", force_scan=True, count_only=True)
        assert cnt == len(ids) and st["path"] == 0 and st["launches"] == 1
        assert np.all(np.diff(ids.astype(np.int64)) > 0) and ids[-1] < k
        for a in (0, k - 1_000_000):
            win = {c: eng.fetch_column(c, a, 1_000_000) for c in cols}
            want = Oracle.from_columns(win).scan(w).astype(np.int64) + a
            got = ids[(ids >= a) & (ids < a + 1_000_000)]
            assert np.array_equal(got.astype(np.int64), want), a
        # complement counts over the whole table
        ca, _, _ = eng.select_ids_device("SELECT c FROM t WHERE (risk_level > 2)", force_scan=True, count_only=True)
        cb, _, _ = eng.select_ids_device("SELECT c FROM t WHERE (risk_level <= 2)", force_scan=True, count_only=True)
        assert ca + cb == n
        # a dense result (600 M consecutive ids): count in closed form, ids an arithmetic sequence at both ends and
        # across a chunk boundary in the middle
        lo = 400_000_000
        dense = f"SELECT command_id FROM Commands WHERE (command_id >= {lo})"
        for rep in range(2):
            cnt, dptr, st = eng.select_ids_device(dense, force_scan=True)
            assert cnt == n - lo
            for off in (0, 300_000_000 - 12_345, cnt - 1_000_000):
                got = eng.copy_from_device(dptr + 4 * off, 1_000_000)
                assert np.array_equal(got.astype(np.int64), np.arange(lo + off, lo + off + 1_000_000)), (rep, off)
    finally:
        eng.close()


def test_insert_delete_on_synthetic_table_vs_oracle(pkg, tmp_path):
    """DELETE mask + stable compaction + index rebuild, then INSERT, checked against the oracle"""
    n = 300_000
    eng = pkg.Engine.from_synth(n, indexes=IDX)
    cols = list(pkg.COLUMNS)
    w_del = '(risk_level > 3) OR (shell_type = "fish" AND exit_code != 0)'
    o0 = Oracle.from_columns({c: eng.fetch_column(c) for c in cols})
    gone = o0.scan(w_del)
    out = eng.run(f"DELETE FROM Commands WHERE {w_del}")
    assert f"Rows affected: {len(gone)}." in out
    assert eng.num_rows == n - len(gone)
    keep = np.setdiff1d(np.arange(n, dtype=np.uint32), gone)
    cid_after = eng.fetch_column("command_id")
    assert np.array_equal(cid_after, o0._keep[0][keep])  # stable: survivors keep their relative order
    o1 = Oracle.from_columns({c: eng.fetch_column(c) for c in cols})
    for w in ['user_id = 1001', 'exit_code >= 126 AND sudo_used = TRUE', 'risk_level = 3']:
        ids, _ = eng.select_ids(f"SELECT command_id FROM Commands WHERE {w}")
        assert np.array_equal(ids, o1.select_ids(w, IDX)[0]), w
    out = eng.run('INSERT INTO Commands VALUES (424242, "ls -la /", "ls", "bash", 0, "2026-01-01T00:00:00.000Z", "TRUE", "/", 1001, "student1001", "labpc-01", 3)')
    assert out.startswith("Executing Query:") and "Insert successful." in out
    assert eng.num_rows == n - len(gone) + 1
    ids, _ = eng.select_ids("SELECT command_id FROM Commands WHERE user_id = 1001")
    assert ids[0] == eng.num_rows - 1   # newest row first among equal keys
    eng.close()


def test_two_engines_run_side_by_side(pkg):
    """calls on DIFFERENT engines of one process are not serialised against each other (an engine has its own lock,
    stream and scratch): two threads, one engine each -- the QPEOMP pattern (QPEOMP.c:234-335) with an engine per
    thread -- must each get exactly their own results"""
    import threading
    sizes = (700_001, 1_300_003)
    cols = NUM_COLS + ["shell_type"]
    engines = [pkg.Engine.from_synth(n, columns=cols, indexes=(("user_id", 1),)) for n in sizes]
    oracles = [Oracle.from_columns({c: e.fetch_column(c) for c in cols}) for e in engines]
    wheres = ['(command_id < 300000) AND (sudo_used = FALSE OR risk_level > 3)', 'user_id = 1001 OR (exit_code = 127)',
              '(shell_type = "zsh") AND (risk_level >= 4)', '(exit_code != 0)']
    want = [[o.select_ids(w, (("user_id", 1),))[0] for w in wheres] for o in oracles]
    errors = []

    def work(k):
        try:
            for rep in range(25):
                for w, expect in zip(wheres, want[k]):
                    ids, _ = engines[k].select_ids(f"SELECT command_id FROM Commands WHERE {w}")
                    if not np.array_equal(ids, expect):
                        errors.append((k, rep, w))
                        return
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in engines:
        e.close()
    assert not errors, errors[:3]


@pytest.mark.parametrize("n", [1, 15, 16, 17, 255, 256, 257, 4095, 4096, 4097, 65_537, 1_048_577])
def test_probe_node_boundaries(pkg, n):
    """K3 answers a probe with ONE descent when the range ends in the leaf node the lower bound falls in (or right behind
    it) and with a second descent otherwise: index sizes around the fan-out's powers, every kind of key run (unique ids,
    5 values, ~2 000 values), point probes on every node's first and last key, ranges of all widths, keys outside the
    index, inverted ranges -- against numpy's searchsorted on the sorted column"""
    eng = pkg.Engine.from_synth(n, columns=["command_id", "user_id", "risk_level"],
                                indexes=(("command_id", 0), ("user_id", 1), ("risk_level", 1)))
    rng = np.random.default_rng(n)
    for attr, dt in (("command_id", np.uint64), ("user_id", np.int32), ("risk_level", np.int32)):
        keys = np.sort(eng.fetch_column(attr), kind="stable")
        edge = np.concatenate([keys[::16], keys[15::16], keys[-1:], keys[:1]]).astype(np.int64)
        lo = np.concatenate([edge, edge - 1, edge + 1, rng.integers(int(keys[0]) - 3, int(keys[-1]) + 4, 500)])
        width = rng.choice([0, 0, 0, 1, 2, 15, 16, 17, 300, 70_000], size=lo.size)
        hi = lo + width
        hi[::37] = lo[::37] - 1                      # inverted: nothing
        if dt == np.uint64:
            lo, hi = np.clip(lo, 0, None), np.clip(hi, 0, None)
        lo, hi = lo.astype(dt), hi.astype(dt)
        first, count, _ = eng.probe_keys(attr, lo, hi)
        wf = np.searchsorted(keys, lo, side="left")
        wc = np.clip(np.searchsorted(keys, hi, side="right") - wf, 0, None)
        assert np.array_equal(first, wf.astype(np.uint32)), (n, attr)
        assert np.array_equal(count, wc.astype(np.uint32)), (n, attr)
        f1, c1, _ = eng.probe_keys(attr, lo)         # point form (hi = lo)
        w1 = np.searchsorted(keys, lo, side="right") - wf
        assert np.array_equal(f1, wf.astype(np.uint32)) and np.array_equal(c1, w1.astype(np.uint32)), (n, attr)
    eng.close()


def test_index_at_1b_rows(pkg):
    """BASELINE's largest table with indexes (size-independent properties): K4 sorts 1 B u64 row ids (4 digit passes,
    163 k tiles) and 1 B int user ids (2 passes); probes answer with the closed form / the scan's count, and an indexed
    SELECT returns the key's rows in DESCENDING position order -- the reference's leaf order"""
    n = 1_000_000_000
    eng = pkg.Engine.from_synth(n, columns=["command_id", "user_id"], indexes=(("command_id", 0), ("user_id", 1)))
    lo = np.array([0, 5, 123_456_789, n - 1, n, n + 5], dtype=np.uint64)
    first, count, _ = eng.probe_keys("command_id", lo)
    assert first.tolist() == [0, 5, 123_456_789, n - 1, n, n] and count.tolist() == [1, 1, 1, 1, 0, 0]
    first, count, _ = eng.probe_keys("command_id", np.array([10, n - 3], dtype=np.uint64), np.array([1_000_009, n + 9], dtype=np.uint64))
    assert first.tolist() == [10, n - 3] and count.tolist() == [1_000_000, 3]
    assert eng.index_slice("command_id", 999_999_990, 5).tolist() == list(range(999_999_990, 999_999_995))
    ids, st = eng.select_ids("SELECT command_id FROM Commands WHERE user_id = 1001")
    assert st["path"] == 1
    cnt, _, _ = eng.select_ids_device("SELECT command_id FROM Commands WHERE user_id = 1001", force_scan=True, count_only=True)
    assert len(ids) == cnt > 0
    assert np.all(np.diff(ids.astype(np.int64)) < 0)          # newest row first among equal keys
    assert np.all(eng.fetch_column("user_id", int(ids[-1]), 1) == 1001)
    # the index's key array is sorted: sample slices across it
    for start in (0, 333_333_333, 999_000_000):
        k = eng.index_slice_keys("user_id", start, 100_000)
        assert np.all(k[1:] >= k[:-1])
    eng.close()
