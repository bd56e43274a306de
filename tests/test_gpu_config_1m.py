"""BASELINE configs[1]: a 1 M-row CSV with sample-queries-FULL, bit-exact against the reference's
serial engine (QPESeq's loop, full dump: every result row printed, not just 20).  The CSV comes
from the device generator (the reference's LFS fixtures are absent and its Python generator makes
12 k rows/s); both engines ingest the same file."""
import os
import subprocess
import time

import pytest

import support
from support import SAMPLE_QUERIES_FULL, normalise

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(support.REF_DUMP), reason="oracle/_ref/qpe_ref_dump not built")
def test_sample_queries_full_on_1m_rows(tmp_path):
    pkg = support.load_pkg()
    n = 1_000_000
    gen = pkg.Engine.from_synth(n)
    master = str(tmp_path / "commands_1m.csv")
    gen.write_csv(master)
    gen.close()
    qf = tmp_path / "sample-queries-FULL.txt"
    qf.write_text(SAMPLE_QUERIES_FULL)
    (tmp_path / "ref").mkdir()
    (tmp_path / "gpu").mkdir()
    csv_ref = support.scratch_copy(master, tmp_path / "ref")
    csv_gpu = support.scratch_copy(master, tmp_path / "gpu")

    t0 = time.perf_counter()
    want = support.ref_dump(csv_ref, str(qf), max_rows=0)
    t_ref = time.perf_counter() - t0

    exe = os.path.join(support.PKG_DIR, "QPEGPU")
    t0 = time.perf_counter()
    r = subprocess.run([exe, csv_gpu, str(qf), "0"], capture_output=True, timeout=900)
    t_gpu = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    got = r.stdout.decode(errors="replace")
    got = got[:got.index("\x1b[36m=======")]
    assert normalise(got) == normalise(want)
    assert open(csv_gpu, "rb").read() == open(csv_ref, "rb").read()
    print(f"1M rows, sample-queries-FULL, full dump: reference {t_ref:.2f} s, QPEGPU {t_gpu:.2f} s "
          f"({len(want) / 1e6:.1f} MB of output)")
