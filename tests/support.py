"""Shared test plumbing: loads the product package (hyphenated directory name), the compiled
UNMODIFIED reference (oracle/_ref/libqpe_ref.so, test infrastructure) and the C oracle."""
import ctypes as C
import importlib.util
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "parallel-query-processing-system_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
CSV_2K = os.path.join(GOLDEN, "commands_2k.csv")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_LIB = os.path.join(REF_DIR, "libqpe_ref.so")
REF_DUMP = os.path.join(REF_DIR, "qpe_ref_dump")


def load_pkg():
    name = "pqps_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def scratch_copy(csv_path, tmpdir=None):
    """INSERT / DELETE mutate the CSV: always work on a copy."""
    d = tmpdir or tempfile.mkdtemp(prefix="qpe_")
    dst = os.path.join(str(d), os.path.basename(csv_path))
    shutil.copyfile(csv_path, dst)
    return dst


class Ref:
    """The compiled reference serial engine behind oracle/ref_harness.c."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = C.CDLL(REF_LIB)
            lib.ref_open.restype = C.c_void_p
            lib.ref_open.argtypes = [C.c_char_p, C.c_int]
            lib.ref_open_idx.restype = C.c_void_p
            lib.ref_open_idx.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int)]
            lib.ref_close.argtypes = [C.c_void_p]
            lib.ref_num_records.argtypes = [C.c_void_p]
            lib.ref_select.restype = C.c_void_p
            lib.ref_select.argtypes = [C.c_void_p, C.c_char_p]
            lib.ref_delete.argtypes = [C.c_void_p, C.c_char_p]
            lib.ref_run.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
            lib.ref_free_result.argtypes = [C.c_void_p]
            lib.ref_result_rows.argtypes = [C.c_void_p]
            lib.ref_result_cols.argtypes = [C.c_void_p]
            lib.ref_result_colname.restype = C.c_char_p
            lib.ref_result_colname.argtypes = [C.c_void_p, C.c_int]
            lib.ref_result_cell.restype = C.c_char_p
            lib.ref_result_cell.argtypes = [C.c_void_p, C.c_int, C.c_int]
            lib.ref_scan_positions.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.c_int]
            lib.ref_time_scan.restype = C.c_double
            lib.ref_time_scan.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
            lib.ref_time_select.restype = C.c_double
            lib.ref_time_select.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
            cls._lib = lib
        return cls._lib

    @staticmethod
    def available():
        return os.path.exists(REF_LIB)

    def __init__(self, csv, num_indexes=5, indexes=None):
        lib = self.lib()
        if indexes is None:
            self.h = lib.ref_open(csv.encode(), num_indexes)
        else:
            n = len(indexes)
            names = (C.c_char_p * max(n, 1))(*[a.encode() for a, _ in indexes])
            types = (C.c_int * max(n, 1))(*[t for _, t in indexes])
            self.h = lib.ref_open_idx(csv.encode(), n, names, types)
        assert self.h

    def close(self):
        if self.h:
            self.lib().ref_close(self.h)
            self.h = None

    @property
    def num_rows(self):
        return self.lib().ref_num_records(self.h)

    def select(self, sql):
        lib = self.lib()
        r = lib.ref_select(self.h, sql.encode())
        assert r, "reference did not parse a SELECT: " + sql
        try:
            nr, nc = lib.ref_result_rows(r), lib.ref_result_cols(r)
            names = [lib.ref_result_colname(r, j).decode() for j in range(nc)]
            rows = [[lib.ref_result_cell(r, i, j).decode(errors="replace") for j in range(nc)] for i in range(nr)]
            return names, rows
        finally:
            lib.ref_free_result(r)

    def scan_positions(self, sql):
        lib = self.lib()
        n = self.num_rows
        buf = (C.c_int * max(n, 1))()
        m = lib.ref_scan_positions(self.h, sql.encode(), buf, n)
        return list(buf[:m])

    def delete(self, sql):
        return self.lib().ref_delete(self.h, sql.encode())

    def run(self, sql, max_rows=20):
        self.lib().ref_run(self.h, sql.encode(), max_rows)


def ref_dump(csv, query_file, max_rows=0, num_indexes=5, cwd=None):
    """stdout of the reference's QPESeq loop over a query file (oracle/_ref/qpe_ref_dump)."""
    out = subprocess.run([REF_DUMP, csv, query_file, str(max_rows), str(num_indexes)], capture_output=True, cwd=cwd,
                         timeout=3600)
    assert out.returncode == 0, out.stderr.decode(errors="replace")
    return out.stdout.decode(errors="replace")


import re

_TIME_PATTERNS = [
    (re.compile(r"Query Time: [0-9.]+ seconds"), "Query Time: T seconds"),
    (re.compile(r"Execution Time: [0-9.]+"), "Execution Time: T"),
]


def normalise(text):
    """Blank the only non-deterministic tokens of the printed protocol."""
    for pat, rep in _TIME_PATTERNS:
        text = pat.sub(rep, text)
    return text


SAMPLE_QUERIES_FULL = """# -- SAMPLE SQL QUERIES --
# -- Sample 1:
SELECT command_id, base_command, sudo_used, user_name, timestamp
FROM Commands
WHERE sudo_used = FALSE AND user_name = "student1030";

# -- Sample 2:
SELECT command_id, raw_command, user_name, risk_level, timestamp
FROM Commands
WHERE sudo_used = TRUE AND risk_level > 2;

# -- Sample 3:
SELECT raw_command, exit_code, timestamp, sudo_used, user_name, risk_level
FROM Commands
WHERE risk_level > 3;

# -- Sample 4:
SELECT *
FROM Commands
WHERE risk_level = 5;

# -- Sample 5:
INSERT INTO Commands VALUES (999999, "echo 'test insert'", "echo", "bash", 0, "2025-12-01T12:00:00.000Z", "FALSE", "/home/test", 1000, "testuser", "test-host", 1);

# -- Sample 6:
DELETE FROM Commands WHERE command_id = 999999;

# -- Sample 7:
SELECT command_id, raw_command, risk_level, exit_code
FROM Commands
WHERE sudo_used = TRUE OR (risk_level = 5 AND shell_type = "bash");

# -- Sample 8:
SELECT user_name, working_directory, base_command
FROM Commands
WHERE user_id = 1001 OR (user_name = "student1002" AND shell_type = "zsh");
"""

# WHERE clauses exercising both paths, precedence, strings, duplicates, OR-miss
PROBE_WHERES = [
    'sudo_used = FALSE AND user_name = "student1030"',
    'sudo_used = TRUE AND risk_level > 2',
    'risk_level > 3',
    'risk_level = 5',
    'sudo_used = TRUE OR (risk_level = 5 AND shell_type = "bash")',
    'user_id = 1001 OR (user_name = "student1002" AND shell_type = "zsh")',
    'risk_level = 5 OR user_id = 1001',
    'command_id < 10',
    'command_id >= 1990 AND shell_type != "bash"',
    'shell_type = "zsh" AND host_name = "labpc-01" OR base_command = "ls"',
    'base_command = "ls" OR shell_type = "zsh" AND host_name = "labpc-01"',
    'timestamp > "2026-10-1" AND user_name <= "student1005"',
    'exit_code = -1 AND host_name = "cs-lab-02"',
    '(risk_level >= 2 OR exit_code != 0) AND (user_id < 1020 OR shell_type != "sh") AND (sudo_used = 1)',
    'raw_command = "ls -la"',
    'exit_code > 126 AND risk_level < 2',
    '(command_id < 1000) AND (sudo_used = FALSE OR risk_level > 3)',
    '(command_id < 500) AND (shell_type = "bash" OR host_name = "labpc-01")',
    'working_directory >= "/home/student1010" AND working_directory < "/tmp"',
    'host_name != "labpc-01" AND (base_command = "git" OR (risk_level > 2 AND exit_code = 0))',
    'user_name = "student1030"',
    'user_name > "student1030xyz"',
    'nosuch_column = 5',
    'nosuch_column = 5 OR risk_level = 4',
    'sudo_used > TRUE',
    'sudo_used != TRUE',
    'command_id != 7',
    'user_id <= 1003',
    'exit_code >= 127',
    'risk_level > 3 AND exit_code = 130 AND sudo_used = FALSE AND shell_type = "bash"',
    '((risk_level = 1 OR risk_level = 2) AND (shell_type = "zsh")) OR (exit_code = 130)',
    'command_id = 1234',
    'command_id > 1995',
    'command_id <= 3',
    'raw_command > "z"',
    'timestamp < "2025"',
]
