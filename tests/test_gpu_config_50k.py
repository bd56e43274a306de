"""BASELINE configs[0]'s data at its own size: 50 000 rows of the reference's OWN generator
(tests/golden/commands_50k.csv.gz, made by tools/gen_reference_csv.py), the CUDA path against digests of what the
COMPILED, UNMODIFIED reference returned for it (tests/golden/golden_50k.json, made by tools/make_golden_50k.py in the
authoring container): row ids of 36 probe WHEREs on both paths, the printed output of sample-queries-FULL with every
row (1.4 MB), and the data file after its INSERT + DELETE.  Needs neither /root/reference nor oracle/_ref."""
import gzip
import hashlib
import json
import os
import subprocess

import pytest

import support
from support import GOLDEN, SAMPLE_QUERIES_FULL, normalise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def csv_50k(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg50k")
    p = str(d / "commands_50k.csv")
    with gzip.open(os.path.join(GOLDEN, "commands_50k.csv.gz"), "rb") as f, open(p, "wb") as out:
        out.write(f.read())
    return p


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(GOLDEN, "golden_50k.json")))


def sha(values):
    return hashlib.sha256(",".join(map(str, values)).encode()).hexdigest()


def test_probe_wheres_match_the_reference_digests(csv_50k, golden, tmp_path):
    pkg = support.load_pkg()
    eng = pkg.Engine.from_csv(support.scratch_copy(csv_50k, tmp_path))
    cid = eng.fetch_column("command_id")
    for p in golden["probes"]:
        sql = f"SELECT command_id FROM Commands WHERE {p['where']}"
        ids, _ = eng.select_ids(sql)                      # the reference's path rule (index / scan) and order
        assert len(ids) == p["n_select"], p["where"]
        assert sha(cid[ids].tolist()) == p["select_sha256"], p["where"]
        pos, _ = eng.select_ids(sql, force_scan=True)     # linearSearchRecords: table positions in table order
        assert len(pos) == p["n_scan"] and sha(pos.tolist()) == p["scan_sha256"], p["where"]
    eng.close()


def test_sample_queries_full_output_and_side_effects(csv_50k, golden, tmp_path):
    run_csv = support.scratch_copy(csv_50k, tmp_path)
    qf = tmp_path / "sample-queries-FULL.txt"
    qf.write_text(SAMPLE_QUERIES_FULL)
    exe = os.path.join(support.PKG_DIR, "QPEGPU")
    r = subprocess.run([exe, run_csv, str(qf), "0"], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    got = r.stdout.decode(errors="replace")
    got = normalise(got[:got.index("\x1b[36m=======")])
    want = golden["sample_full"]
    assert got[:600] == want["head"]
    assert len(got) == want["bytes"] and got.count("\n") == want["lines"]
    assert hashlib.sha256(got.encode()).hexdigest() == want["sha256"]
    assert hashlib.sha256(open(run_csv, "rb").read()).hexdigest() == golden["csv_after_sha256"]
