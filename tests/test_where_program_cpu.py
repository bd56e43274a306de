"""The WHERE compiler (csrc/where_compile.cpp) on the CPU: the predicate program the engine uploads to the scan
kernel is interpreted here, row by row in numpy, and must select exactly the rows the oracle's restatement of
evaluateWhereClause / checkCondition selects (engine/serial/executeEngine-serial.c:251-316): right-recursive,
precedence-free AND/OR lists, nested groups, unknown attributes and missing comparators as constant false, libc
literal conversions.  No device is needed: qpe_sql_compile_program returns the raw `struct Program`."""
import ctypes as C
import struct

import numpy as np
import pytest

import support
from support import CSV_2K, Oracle

pytestmark = pytest.mark.skipif(not Oracle.available(), reason="oracle not built")

T_U64, T_I32, T_STR, T_BOOL = 0, 1, 2, 3
COL_TYPES = {"command_id": T_U64, "exit_code": T_I32, "user_id": T_I32, "risk_level": T_I32, "sudo_used": T_BOOL}
P_LEAF_SET, P_LEAF_AND, P_LEAF_OR, P_PUSH, P_POP_AND, P_POP_OR, P_CONST, P_NOT = range(8)
MAX_INSTR, MAX_LEAVES, LIT_POOL = 96, 48, 2048
OFF_INSTR, OFF_LEAF = 16, 16 + 2 * MAX_INSTR
OFF_LIT = OFF_LEAF + 24 * MAX_LEAVES
PROGRAM_BYTES = OFF_LIT + LIT_POOL

WHERES = [
    '(command_id < 1500) AND (sudo_used = FALSE OR risk_level > 3)',
    'shell_type = "zsh" AND host_name = "labpc-01" OR base_command = "ls"',
    'risk_level > 3 AND shell_type != "bash"',
    'user_id = 1001 OR (exit_code = 127)',
    'working_directory >= "/home/student1010" AND working_directory < "/tmp"',
    'host_name != "labpc-01" AND (base_command = "git" OR (risk_level > 2 AND exit_code = 0))',
    '(risk_level >= 2 OR exit_code != 0) AND (user_id < 2000 OR shell_type != "sh") AND (command_id < 1000)',
    'nosuchcolumn = 3 OR risk_level = 5',
    'nosuchcolumn = 3 AND risk_level = 5',
    'sudo_used < TRUE OR exit_code = 0',                 # bool has no "<" comparator: constant false (:207-212)
    'sudo_used = 1 AND risk_level <= 1',
    'command_id != 7',
    'exit_code >= -1 AND exit_code <= 1',
    '((risk_level = 5))',
    'risk_level = 1 OR risk_level = 2 AND sudo_used = TRUE OR exit_code = 1',   # no precedence: a OR (b AND (c OR d))
    'timestamp > "2026" OR user_name = "student1001"',
    'raw_command = "ls -la"',
]


@pytest.fixture(scope="module")
def table():
    pkg = support.load_pkg()
    o = Oracle.from_csv(CSV_2K)
    n = o.num_rows
    cols = {}
    widths = []
    for name in pkg.COLUMNS:
        cells = [o.cell(i, name) for i in range(n)]
        t = COL_TYPES.get(name, T_STR)
        if t == T_U64:
            cols[name] = np.array([int(c) for c in cells], dtype=np.uint64)
            widths.append(8)
        elif t == T_I32:
            cols[name] = np.array([int(c) for c in cells], dtype=np.int64)
            widths.append(4)
        elif t == T_BOOL:
            cols[name] = np.array([c == "true" for c in cells], dtype=bool)
            widths.append(1)
        else:
            raw = [c.encode("utf-8", errors="surrogateescape") for c in cells]
            w = (max(len(r) for r in raw) + 1 + 15) // 16 * 16
            cols[name] = [r.ljust(w, b"\0") for r in raw]
            widths.append(w)
    yield pkg, o, cols, widths, n
    o.close()


def _three_way_result(lt, eq, tt):
    gt = ~lt & ~eq
    return (lt & bool(tt & 1)) | (eq & bool(tt & 2)) | (gt & bool(tt & 4))


def _run_program(blob, pkg, cols, widths, n):
    n_instr, n_leaves, col_mask, _ = struct.unpack_from("<iiII", blob, 0)
    acc = np.ones(n, dtype=bool)
    stack = {}
    for i in range(n_instr):
        op, arg = struct.unpack_from("<BB", blob, OFF_INSTR + 2 * i)
        if op <= P_LEAF_OR:
            col, typ, tt, _nch, lit_off, lit_u64, lit_i32, _so = struct.unpack_from("<BBBBIQiI", blob, OFF_LEAF + 24 * arg)
            name = pkg.COLUMNS[col]
            assert col_mask & (1 << col)
            data = cols[name]
            if typ == T_U64:
                v = _three_way_result(data < np.uint64(lit_u64), data == np.uint64(lit_u64), tt)
            elif typ == T_I32:
                v = _three_way_result(data < lit_i32, data == lit_i32, tt)
            elif typ == T_BOOL:
                want = (tt == 0b010) == bool(lit_i32 & 1)       # only = and != exist
                v = data == want
            else:
                w = widths[col]
                lit = bytes(blob[OFF_LIT + lit_off: OFF_LIT + lit_off + w])
                v = _three_way_result(np.array([c < lit for c in data]), np.array([c == lit for c in data]), tt)
            acc = v if op == P_LEAF_SET else (acc & v) if op == P_LEAF_AND else (acc | v)
        elif op == P_PUSH:
            stack[arg] = acc.copy()
        elif op in (P_POP_AND, P_POP_OR):
            acc = (acc & stack[arg]) if op == P_POP_AND else (acc | stack[arg])
        elif op == P_CONST:
            acc = np.full(n, bool(arg))
        elif op == P_NOT:
            acc = ~acc
        else:
            raise AssertionError(f"unknown op {op}")
    return np.nonzero(acc)[0]


@pytest.mark.parametrize("where", WHERES)
def test_compiled_program_selects_the_oracles_rows(table, where):
    pkg, o, cols, widths, n = table
    lib = pkg.load_library()
    buf = C.create_string_buffer(PROGRAM_BYTES)
    w = (C.c_uint * 12)(*widths)
    rc = lib.qpe_sql_compile_program(f"SELECT command_id FROM Commands WHERE {where}".encode(), w, buf, PROGRAM_BYTES)
    assert rc == PROGRAM_BYTES, rc
    got = _run_program(buf.raw, pkg, cols, widths, n)
    want = o.scan(where)
    assert np.array_equal(got, np.asarray(want, dtype=np.int64)), where


def test_program_without_where_matches_every_row(table):
    pkg, o, cols, widths, n = table
    lib = pkg.load_library()
    buf = C.create_string_buffer(PROGRAM_BYTES)
    rc = lib.qpe_sql_compile_program(b"SELECT * FROM Commands", (C.c_uint * 12)(*widths), buf, PROGRAM_BYTES)
    assert rc == PROGRAM_BYTES
    assert len(_run_program(buf.raw, pkg, cols, widths, n)) == n
    assert lib.qpe_sql_compile_program(b"this is not sql", (C.c_uint * 12)(*widths), buf, PROGRAM_BYTES) == -7
    assert lib.qpe_sql_compile_program(b"SELECT * FROM t WHERE a = 1", (C.c_uint * 12)(*widths), buf, 16) == -5
