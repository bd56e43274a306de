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


# ---------------------------------------------------------------------------------------------
# C oracle (oracle/liboracle.so) -- the CPU restatement, used ONLY as a checker
# ---------------------------------------------------------------------------------------------
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")
COLUMNS = ("command_id", "raw_command", "base_command", "shell_type", "exit_code", "timestamp", "sudo_used",
           "working_directory", "user_id", "user_name", "host_name", "risk_level")
DEFAULT_INDEXES = (("command_id", 0), ("user_id", 1), ("risk_level", 1), ("exit_code", 1), ("sudo_used", 3))


class WhereNode(C.Structure):
    """struct whereClauseS (include/executeEngine-serial.h:48-56)"""


WhereNode._fields_ = [("attribute", C.c_char_p), ("operator", C.c_char_p), ("value", C.c_char_p),
                      ("value_type", C.c_int), ("next", C.POINTER(WhereNode)), ("logical_op", C.c_char_p),
                      ("sub", C.POINTER(WhereNode))]


class OracleTable(C.Structure):
    _fields_ = [("n", C.c_longlong), ("col", C.c_void_p * 12), ("width", C.c_uint * 12)]


_WTOK = re.compile(r'\s*(?:(\()|(\))|(AND\b)|([oO][rR]\b)|(>=|<=|!=|=|>|<)|"([^"]*)"|\'([^\']*)\'|(-?\d+)|(\w+))')


def parse_where(text):
    """WHERE text (well-formed subset of the reference grammar) -> nested list form:
    [item, op, item, op, ...] with item = (col, op, value) or a nested list.  Follows the
    tokenizer's literal rules (tokenizer.c:47-107): quotes dropped, digits only, TRUE/FALSE upper."""
    pos = 0
    toks = []
    while pos < len(text):
        if text[pos:].strip() == "":
            break
        m = _WTOK.match(text, pos)
        assert m, f"cannot tokenise WHERE at: {text[pos:]!r}"
        pos = m.end()
        if m.group(1): toks.append(("(", None))
        elif m.group(2): toks.append((")", None))
        elif m.group(3): toks.append(("AND", None))
        elif m.group(4): toks.append(("OR", None))
        elif m.group(5): toks.append(("op", m.group(5)))
        elif m.group(6) is not None: toks.append(("val", m.group(6)))
        elif m.group(7) is not None: toks.append(("val", m.group(7)))
        elif m.group(8): toks.append(("val", m.group(8).lstrip("-")))
        else:
            w = m.group(9)
            toks.append(("val", w.upper()) if w.upper() in ("TRUE", "FALSE") else ("id", w))
    i = 0

    def level():
        nonlocal i
        items = []
        while i < len(toks) and toks[i][0] != ")":
            if toks[i][0] == "(":
                i += 1
                items.append(level())
                assert toks[i][0] == ")"
                i += 1
            else:
                assert toks[i][0] == "id" and toks[i + 1][0] == "op" and toks[i + 2][0] == "val", toks[i:i + 3]
                items.append((toks[i][1], toks[i + 1][1], toks[i + 2][1]))
                i += 3
            if i < len(toks) and toks[i][0] in ("AND", "OR"):
                items.append(toks[i][0])
                i += 1
        return items

    return level()


def render_where_tree(tree):
    """same rendering as ref_where_text / qpe_sql_where_to_text"""
    out = []
    for it in tree:
        if isinstance(it, str):
            out.append(it)
        elif isinstance(it, list):
            out.append("( " + render_where_tree(it) + " )")
        else:
            out.append(f"{it[0]} {it[1]} {it[2]}")
    return " ".join(out)


def build_where_nodes(tree, keep):
    """nested list form -> whereClauseS linked list (convert_conditions, connectEngine.c:65-113)"""
    items = [it for it in tree if not isinstance(it, str)]
    ops = [it for it in tree if isinstance(it, str)]
    head = None
    prev = None
    for k, it in enumerate(items):
        n = WhereNode()
        keep.append(n)
        if isinstance(it, list):
            sub = build_where_nodes(it, keep)
            if sub is not None:
                n.sub = C.pointer(sub)
        else:
            n.attribute, n.operator, n.value = it[0].encode(), it[1].encode(), it[2].encode()
        if k < len(items) - 1:
            n.logical_op = (ops[k] if k < len(ops) else "AND").encode()
        if head is None:
            head = n
        else:
            prev.next = C.pointer(n)
        prev = n
    return head


class Oracle:
    """CPU restatement over a columnar table built from a CSV (oracle loader) or numpy columns."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = C.CDLL(ORACLE_LIB)
            pt = C.POINTER(OracleTable)
            pw = C.POINTER(WhereNode)
            lib.oracle_load_csv.restype = C.c_void_p
            lib.oracle_load_csv.argtypes = [C.c_char_p, C.POINTER(C.c_longlong)]
            lib.oracle_table_from_records.restype = pt
            lib.oracle_table_from_records.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
            lib.oracle_table_free.argtypes = [pt]
            lib.oracle_free.argtypes = [C.c_void_p]
            lib.oracle_eval_row.argtypes = [pt, C.c_longlong, pw]
            lib.oracle_scan.restype = C.c_longlong
            lib.oracle_scan.argtypes = [pt, pw, C.c_void_p]
            lib.oracle_scan_range.restype = C.c_longlong
            lib.oracle_scan_range.argtypes = [pt, C.c_longlong, C.c_longlong, pw, C.c_void_p]
            lib.oracle_index_order.argtypes = [pt, C.c_int, C.c_void_p]
            lib.oracle_index_order_by_insertion.argtypes = [pt, C.c_int, C.c_void_p]
            lib.oracle_select.restype = C.c_longlong
            lib.oracle_select.argtypes = [pt, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), pw,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
            lib.oracle_cell_text.argtypes = [pt, C.c_longlong, C.c_char_p, C.c_char_p, C.c_size_t]
            lib.oracle_col_by_name.argtypes = [C.c_char_p]
            cls._lib = lib
        return cls._lib

    @staticmethod
    def available():
        return os.path.exists(ORACLE_LIB)

    def __init__(self, table_ptr=None, owned=True, keep=None):
        self.t = table_ptr
        self.owned = owned
        self._keep = keep

    @classmethod
    def from_csv(cls, path):
        lib = cls.lib()
        n = C.c_longlong()
        rows = lib.oracle_load_csv(path.encode(), C.byref(n))
        assert rows, "cannot read " + path
        t = lib.oracle_table_from_records(rows, n.value, None)
        lib.oracle_free(rows)
        return cls(t, True)

    @classmethod
    def from_columns(cls, cols):
        """cols: {name: numpy array}; numeric arrays 1-D of the column dtype, text arrays (n, width) uint8."""
        import numpy as np
        tab = OracleTable()
        keep = []
        n = None
        for c, name in enumerate(COLUMNS):
            if name not in cols:
                continue
            a = np.ascontiguousarray(cols[name])
            keep.append(a)
            n = a.shape[0] if n is None else n
            assert a.shape[0] == n
            tab.col[c] = a.ctypes.data
            tab.width[c] = a.shape[1] if a.ndim == 2 else a.dtype.itemsize
        tab.n = n or 0
        keep.append(tab)
        return cls(C.pointer(tab), False, keep)

    def close(self):
        if self.t is not None and self.owned:
            self.lib().oracle_table_free(self.t)
        self.t = None

    @property
    def num_rows(self):
        return self.t.contents.n

    @staticmethod
    def where(where_text_or_tree):
        tree = parse_where(where_text_or_tree) if isinstance(where_text_or_tree, str) else where_text_or_tree
        keep = []
        head = build_where_nodes(tree, keep)
        return (C.pointer(head) if head is not None else None), keep

    def scan(self, where, first=None, n=None):
        import numpy as np
        wc, keep = self.where(where) if where is not None else (None, None)
        total = self.num_rows
        first = 0 if first is None else first
        n = total - first if n is None else n
        ids = np.zeros(max(n, 1), dtype=np.uint32)
        m = self.lib().oracle_scan_range(self.t, first, n, wc, ids.ctypes.data)
        return ids[:m]

    def select_ids(self, where, indexes=DEFAULT_INDEXES):
        import numpy as np
        wc, keep = self.where(where) if where is not None else (None, None)
        nidx = len(indexes)
        names = (C.c_char_p * max(nidx, 1))(*[a.encode() for a, _ in indexes])
        types = (C.c_int * max(nidx, 1))(*[t for _, t in indexes])
        out = C.c_void_p()
        used = C.c_int()
        m = self.lib().oracle_select(self.t, nidx, names, types, wc, C.byref(out), C.byref(used))
        assert m >= 0
        ids = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint32)), shape=(max(m, 1),))[:m].copy()
        self.lib().oracle_free(out)
        return ids, bool(used.value)

    def index_order(self, attribute, by_insertion=False):
        import numpy as np
        perm = np.zeros(max(self.num_rows, 1), dtype=np.uint32)
        c = self.lib().oracle_col_by_name(attribute.encode())
        fn = self.lib().oracle_index_order_by_insertion if by_insertion else self.lib().oracle_index_order
        assert fn(self.t, c, perm.ctypes.data) == 0
        return perm[:self.num_rows]

    def cell(self, row, attribute):
        buf = C.create_string_buffer(1024)
        self.lib().oracle_cell_text(self.t, row, attribute.encode(), buf, 1024)
        return buf.value.decode(errors="replace")

    def rows(self, ids, attributes):
        return [[self.cell(int(i), a) for a in attributes] for i in ids]
