"""Drop-in proof: the reference's UNMODIFIED QPESeq.c + connectEngine.c + tokenizer.c +
printHelper.c, compiled with the *Serial entry points renamed to *GPU and linked against
libqpegpu.so (oracle/Makefile `refdriver`), must print exactly what the reference's own QPESeq
prints on the same CSV and sample-queries file -- and leave the same CSV behind."""
import os
import re
import subprocess

import pytest

import support
from support import CSV_2K, SAMPLE_QUERIES_FULL, normalise

pytestmark = pytest.mark.gpu

REFDRIVER = os.path.join(support.REF_DIR, "QPEGPU_refdriver")
QPESEQ = os.path.join(support.REF_DIR, "QPESeq")


def _norm(text):
    text = normalise(text)
    text = re.sub(r"[0-9.]+ seconds", "T seconds", text)
    return text


@pytest.mark.skipif(not (os.path.exists(REFDRIVER) and os.path.exists(QPESEQ)), reason="oracle/_ref drivers not built")
def test_reference_front_end_over_gpu_engine(tmp_path):
    outs = {}
    for name, exe in (("ref", QPESEQ), ("gpu", REFDRIVER)):
        d = tmp_path / name
        d.mkdir()
        (d / "sample-queries.txt").write_text(SAMPLE_QUERIES_FULL)  # QPESeq.c:40 hard-codes this name
        csv = support.scratch_copy(CSV_2K, d)
        r = subprocess.run([exe, csv], cwd=d, capture_output=True, timeout=600)
        assert r.returncode == 0, r.stderr.decode(errors="replace")
        outs[name] = (_norm(r.stdout.decode(errors="replace")), open(csv, "rb").read())
    assert outs["gpu"][0] == outs["ref"][0]
    assert outs["gpu"][1] == outs["ref"][1]


def test_qpegpu_driver_matches_golden_output(tmp_path):
    """our own plain-C driver (QPEGPU.c over the C-ABI) against the golden output of the reference"""
    exe = os.path.join(support.PKG_DIR, "QPEGPU")
    if not os.path.exists(exe):
        pytest.skip("QPEGPU not built")
    qf = tmp_path / "q.txt"
    qf.write_text(SAMPLE_QUERIES_FULL)
    csv = support.scratch_copy(CSV_2K, tmp_path)
    r = subprocess.run([exe, csv, str(qf), "20"], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    got = normalise(r.stdout.decode(errors="replace"))
    got = got[:got.index("\x1b[36m=======")]  # drop the timing banner
    want = open(os.path.join(support.GOLDEN, "sample_full_2k.out")).read()
    assert got == want
    assert open(csv, "rb").read() == open(os.path.join(support.GOLDEN, "commands_2k_after_sample_full.csv"), "rb").read()
