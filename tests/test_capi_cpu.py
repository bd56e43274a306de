"""CPU-only checks of the product library: it loads, exports every symbol include/*.h declares,
refuses to run without a GPU (no CPU fallback), and its SQL front end builds the same WHERE lists
as the reference's tokenizer + parser + bridge (golden vectors from the compiled reference)."""
import ctypes as C
import json
import os
import re

import pytest

import support

INCLUDE = os.path.join(support.ROOT, "include")


def _declared_symbols():
    names = set()
    proto = re.compile(r"^[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", re.M)
    for h in ("executeEngine-gpu.h", "buildEngine-gpu.h", "qpe_gpu.h"):
        text = open(os.path.join(INCLUDE, h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
        for m in proto.finditer(text):
            names.add(m.group(1))
    return sorted(names)


@pytest.fixture(scope="module")
def pkg():
    if not os.path.exists(os.path.join(support.PKG_DIR, "libqpegpu.so")):
        pytest.skip("libqpegpu.so not built (run __graft_entry__.build())")
    return support.load_pkg()


def test_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 30
    for must in ("initializeEngineGPU", "destroyEngineGPU", "executeQuerySelectGPU", "executeQueryDeleteGPU",
                 "executeQueryInsertGPU", "addAttributeIndexGPU", "freeResultSet", "isAttributeIndexed", "makeIndexGPU",
                 "getAllRecordsFromFileGPU", "qpe_gpu_select_ids", "qpe_gpu_probe_batch", "qpe_gpu_match_mask"):
        assert must in syms
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback(pkg):
    """without a CUDA device the engine refuses to exist (on the GPU box this test is a no-op)"""
    if pkg.gpu_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.QpeError):
        pkg.Engine.from_csv(support.CSV_2K)
    with pytest.raises(pkg.QpeError):
        pkg.Engine.from_synth(1000)


def test_sql_front_end_matches_reference_parser(pkg):
    gold = json.load(open(os.path.join(support.GOLDEN, "where_text.json")))
    assert len(gold) > 40
    for e in gold:
        if e["command"] in (2, 4):  # CMD_SELECT, CMD_DELETE
            assert pkg.where_text(e["statement"]) == e["where"], e["statement"]
        else:
            assert pkg.where_text(e["statement"]) == ""


def test_csv_loader_matches_oracle_loader(pkg):
    """getAllRecordsFromFileGPU (host-only helper) == the oracle's restatement of the serial loader"""
    if not support.Oracle.available():
        pytest.skip("oracle not built")
    lib = pkg.load_library()
    lib.getAllRecordsFromFileGPU.restype = C.POINTER(C.c_void_p)
    lib.getAllRecordsFromFileGPU.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    n = C.c_int()
    rows = lib.getAllRecordsFromFileGPU(support.CSV_2K.encode(), C.byref(n))
    assert n.value == 2000
    olib = support.Oracle.lib()
    on = C.c_longlong()
    orows = olib.oracle_load_csv(support.CSV_2K.encode(), C.byref(on))
    assert on.value == 2000
    for i in range(2000):
        a = C.string_at(rows[i], 1040)
        b = C.string_at(orows + i * 1040, 1040)
        assert a == b, i


def test_csv_loader_matches_oracle_loader_on_nasty_text(pkg, tmp_path):
    """the same comparison on deliberately nasty CSV text (quotes, "" escapes, text after a closing quote, absent
    fields, blank lines, overflowing numbers, over-long fields; the rows of tests/test_gpu_ingest.py) and on
    both line endings -- the host loader is the exact path the device parser (K6) falls back to"""
    if not support.Oracle.available():
        pytest.skip("oracle not built")
    from test_gpu_ingest import nasty_rows, write_csv
    lib = pkg.load_library()
    lib.getAllRecordsFromFileGPU.restype = C.POINTER(C.c_void_p)
    lib.getAllRecordsFromFileGPU.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    olib = support.Oracle.lib()
    for seed, eol, final_newline in ((7, "\r\n", True), (8, "\n", True), (9, "\n", False)):
        path = str(tmp_path / f"nasty_{seed}.csv")
        write_csv(path, nasty_rows(seed=seed, n=600), eol=eol, final_newline=final_newline)
        n = C.c_int()
        rows = lib.getAllRecordsFromFileGPU(path.encode(), C.byref(n))
        on = C.c_longlong()
        orows = olib.oracle_load_csv(path.encode(), C.byref(on))
        assert n.value == on.value > 500
        for i in range(n.value):
            assert C.string_at(rows[i], 1040) == C.string_at(orows + i * 1040, 1040), (seed, i)
