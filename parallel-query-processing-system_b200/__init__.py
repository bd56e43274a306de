"""Host-side mirror of the B200 SELECT/WHERE engine (libqpegpu.so) for Python callers.

The product is the C-ABI shared library next to this file (``libqpegpu.so``: sm_100a CUDA
kernels behind the reference's ``executeEngine-*.h`` entry-point shape, see ``include/``).  This
module is a thin ``ctypes`` binding used by the tests, ``bench.py`` and ``__graft_entry__.py``;
it adds no compute of its own and has NO fallback: if the library is missing, or no CUDA device
is usable, it raises.

Names follow the reference (engine, SELECT/DELETE/INSERT, indexes, row ids):
  reference entry point (include/executeEngine-serial.h:69-151)   here
  initializeEngineSerial                                          Engine.from_csv
  executeQuerySelectSerial                                        Engine.select / Engine.select_ids
  executeQueryDeleteSerial / executeQueryInsertSerial             Engine.run("DELETE ..." / "INSERT ...")
  run_test_query (connectEngine.c:125)                            Engine.run
  findRange / find_rows (engine/bplus.c:282,361)                  Engine.probe_batch / Engine.index_slice
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QPE_LIB_PATH") or os.path.join(_HERE, "libqpegpu.so")  # override: A/B of two builds

COLUMNS = ("command_id", "raw_command", "base_command", "shell_type", "exit_code", "timestamp", "sudo_used",
           "working_directory", "user_id", "user_name", "host_name", "risk_level")
# the reference's default index set (connectEngine.c:48-62): attribute -> type code 0=u64 1=int 2=string 3=bool
DEFAULT_INDEXES = (("command_id", 0), ("user_id", 1), ("risk_level", 1), ("exit_code", 1), ("sudo_used", 3))
NUMERIC_DTYPES = {"command_id": np.uint64, "exit_code": np.int32, "user_id": np.int32, "risk_level": np.int32,
                  "sudo_used": np.uint8}

SCAN_FORCE = 1
SCAN_COUNT_ONLY = 2
SCAN_GLOBAL_IDS = 4


class QpeError(RuntimeError):
    pass


def pinned_array(n: int, dtype) -> np.ndarray:
    """numpy array over pinned host memory from qpe_gpu_host_alloc (copied to / from the device in place, without a
    bounce buffer); released when the array is garbage collected"""
    lib = load_library()
    dt = np.dtype(dtype)
    nbytes = max(int(n) * dt.itemsize, 16)
    p = lib.qpe_gpu_host_alloc(nbytes)
    if not p:
        raise QpeError("qpe_gpu_host_alloc failed: " + (lib.qpe_gpu_last_error() or b"").decode())

    class _Owner:
        def __del__(self, p=p, lib=lib):
            lib.qpe_gpu_host_free(p)

    buf = (C.c_uint8 * nbytes).from_address(p)
    buf._owner = _Owner()
    return np.frombuffer(buf, dtype=dt, count=int(n))


class ScanStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("rows_scanned", C.c_longlong),
                ("candidates", C.c_longlong), ("matches", C.c_longlong), ("algo_bytes", C.c_longlong),
                ("path", C.c_int), ("launches", C.c_int), ("tile_rows", C.c_int), ("stages", C.c_int),
                ("grid", C.c_int), ("reserved", C.c_int), ("scan_ms", C.c_double), ("compact_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class KeyT(C.Structure):
    """KEY_T of include/bplus.h:22-30 (16-byte tagged union)."""

    class _V(C.Union):
        _fields_ = [("u64", C.c_uint64), ("i32", C.c_int), ("b", C.c_bool), ("str", C.c_char_p)]

    _fields_ = [("type", C.c_int), ("v", _V)]


class ResultSet(C.Structure):
    """struct resultSetS (include/executeEngine-serial.h:30-38)."""
    _fields_ = [("numRecords", C.c_int), ("numColumns", C.c_int), ("columnNames", C.POINTER(C.c_char_p)),
                ("columnTypes", C.POINTER(C.c_int)), ("data", C.POINTER(C.POINTER(C.c_char_p))),
                ("queryTime", C.c_double), ("success", C.c_bool)]


_lib = None


def load_library() -> C.CDLL:
    """dlopen libqpegpu.so (built in-tree by ``__graft_entry__.build()`` / ``make``). No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QpeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, cp, i, ll, ull, sz = C.c_void_p, C.c_char_p, C.c_int, C.c_longlong, C.c_ulonglong, C.c_size_t
    pstats = C.POINTER(ScanStats)
    sig = {
        "qpe_gpu_available": (i, []),
        "qpe_gpu_last_error": (cp, []),
        "qpe_gpu_free": (None, [vp]),
        "initializeEngineGPU": (vp, [i, C.POINTER(cp), C.POINTER(i), cp, cp]),
        "destroyEngineGPU": (None, [vp]),
        "freeResultSet": (None, [vp]),
        "qpe_gpu_engine_synth": (vp, [ull, ull, ull, ull, C.c_uint, i, C.POINTER(cp), C.POINTER(i)]),
        "qpe_gpu_engine_from_records": (vp, [vp, ll, i, C.POINTER(cp), C.POINTER(i), cp, cp]),
        "qpe_gpu_num_rows": (ll, [vp]),
        "qpe_gpu_row_base": (ull, [vp]),
        "qpe_gpu_probe_batch": (i, [vp, cp, C.POINTER(KeyT), C.POINTER(KeyT), sz, vp, vp, pstats]),
        "qpe_gpu_probe_keys": (i, [vp, cp, vp, vp, sz, vp, vp, i, pstats]),
        "qpe_gpu_sort_pairs": (i, [vp, vp, sz, i, i, i, vp, vp, C.POINTER(i), C.POINTER(C.c_double)]),
        "qpe_gpu_host_alloc": (vp, [sz]),
        "qpe_gpu_host_free": (None, [vp]),
        "qpe_gpu_index_slice": (i, [vp, cp, C.c_uint, C.c_uint, vp]),
        "qpe_gpu_index_slice_keys": (i, [vp, cp, C.c_uint, C.c_uint, vp]),
        "qpe_gpu_fetch_column": (i, [vp, cp, ll, ll, vp, C.POINTER(C.c_uint)]),
        "qpe_gpu_set_tile": (i, [vp, i, i]),
        "qpe_gpu_set_pipeline": (i, [vp, i]),
        "qpe_gpu_stream": (vp, [vp]),
        "qpe_gpu_query_upload_bytes": (C.c_uint, []),
        "qpe_gpu_last_trace": (i, [vp, C.POINTER(C.c_double)]),
        "qpe_gpu_fused_trace": (i, [vp, vp, i]),
        "qpe_gpu_set_timing": (i, [vp, i]),
        "qpe_gpu_timing_totals": (i, [vp, C.POINTER(C.c_double), C.POINTER(ll)]),
        "qpe_gpu_copy_from_device": (i, [vp, vp, sz]),
        "qpe_gpu_last_stats": (i, [vp, pstats]),
        "qpe_gpu_write_csv": (i, [vp, cp]),
        "qpe_sql_run": (None, [vp, cp, i, vp]),
        "qpe_sql_run_to_text": (vp, [vp, cp, i]),
        "qpe_sql_select_ids": (i, [vp, cp, i, C.POINTER(vp), C.POINTER(sz), pstats]),
        "qpe_sql_select_ids_batch": (i, [vp, C.POINTER(cp), i, C.POINTER(vp), C.POINTER(sz), pstats]),
        "qpe_sql_select_ids_device": (i, [vp, cp, i, C.POINTER(ull), C.POINTER(vp), pstats]),
        "qpe_sql_select_ids_into": (i, [vp, cp, i, vp, sz, C.POINTER(sz), pstats]),
        "qpe_sql_match_mask": (i, [vp, cp, vp, sz, C.POINTER(ull), pstats]),
        "qpe_sql_scan_count": (i, [vp, cp, C.POINTER(ull), pstats]),
        "qpe_sql_select_segments": (i, [vp, cp, i, C.POINTER(i), C.POINTER(i), C.POINTER(sz), C.POINTER(vp),
                                        C.POINTER(vp)]),
        "qpe_gpu_compact_to": (i, [vp, vp, i, pstats]),
        "qpe_sql_select_ids_to": (i, [vp, cp, vp, ull, i, C.POINTER(ull), pstats]),
        "qpe_gpu_copy_device": (i, [vp, vp, sz]),
        "qpe_gpu_device_alloc": (vp, [sz]),
        "qpe_gpu_device_free": (None, [vp]),
        "qpe_gpu_ipc_export": (i, [vp, C.c_char_p]),
        "qpe_gpu_ipc_open": (vp, [C.c_char_p]),
        "qpe_gpu_ipc_close": (None, [vp]),
        "qpe_gpu_copy_to_host": (i, [vp, vp, sz]),
        "qpe_gpu_copy_to_device": (i, [vp, vp, sz]),
        "qpe_shard_init": (i, [vp, i, i, ull, C.c_char_p]),
        "qpe_shard_connect": (i, [vp, C.c_char_p]),
        "qpe_shard_set_device_result": (i, [vp, i, vp, ull]),
        "qpe_shard_result_ids": (ull, [i, ull]),
        "qpe_shard_open_host_result": (vp, [vp, cp, ull, i]),
        "qpe_shard_device_result": (vp, [vp]),
        "qpe_shard_host_result": (vp, [vp]),
        "qpe_shard_host_result_at": (vp, [vp, i, C.POINTER(ull)]),
        "qpe_shard_set_deferred": (i, [vp, i]),
        "qpe_shard_pin_host_result": (i, [vp]),
        "qpe_shard_numa": (i, [vp, C.POINTER(i)]),
        "qpe_shard_wait_breakdown": (i, [vp, C.POINTER(C.c_double), C.POINTER(ll), i]),
        "qpe_shard_wait": (i, [vp, C.POINTER(ull), pstats]),
        "qpe_sql_shard_submit": (i, [vp, cp, i]),
        "qpe_shard_close": (None, [vp]),
        "qpe_shard_unlink_host_result": (i, [vp]),
        "qpe_shard_set_multipath": (i, [vp, i]),
        "qpe_shard_set_link_weights": (i, [vp, C.POINTER(C.c_double), i]),
        "qpe_sql_shard_select": (i, [vp, cp, i, C.POINTER(ull), pstats]),
        "qpe_sql_shard_delete": (i, [vp, cp, C.POINTER(ull), C.POINTER(ull)]),
        "qpe_sql_select": (C.POINTER(ResultSet), [vp, cp]),
        "qpe_sql_where_to_text": (vp, [cp]),
        "qpe_sql_compile_program": (ll, [cp, C.POINTER(C.c_uint), vp, sz]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DeviceBuffer:
    """Raw device memory of this process's GPU that the other ranks of the box can map (CUDA IPC)."""

    def __init__(self, n_bytes: int):
        self._lib = load_library()
        self.ptr = self._lib.qpe_gpu_device_alloc(n_bytes)
        if not self.ptr:
            raise QpeError("device allocation failed: " + (self._lib.qpe_gpu_last_error() or b"").decode())
        self.n_bytes = n_bytes

    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        if self._lib.qpe_gpu_ipc_export(self.ptr, buf) != 0:
            raise QpeError("cudaIpcGetMemHandle failed: " + (self._lib.qpe_gpu_last_error() or b"").decode())
        return buf.raw

    def to_host(self, n_items: int, dtype=np.uint32, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(n_items, dtype=dtype)
        if n_items and self._lib.qpe_gpu_copy_to_host(out.ctypes.data, self.ptr, n_items * out.itemsize) != 0:
            raise QpeError("copy to host failed")
        return out[:n_items]

    def free(self):
        if self.ptr:
            self._lib.qpe_gpu_device_free(self.ptr)
            self.ptr = None


def ipc_open(handle: bytes) -> int:
    """Map another rank's DeviceBuffer into this process; returns the device pointer."""
    lib = load_library()
    p = lib.qpe_gpu_ipc_open(handle)
    if not p:
        raise QpeError("cudaIpcOpenMemHandle failed: " + (lib.qpe_gpu_last_error() or b"").decode())
    return p


def ipc_close(ptr: int):
    load_library().qpe_gpu_ipc_close(ptr)


def gpu_available() -> bool:
    return bool(load_library().qpe_gpu_available())


def sort_pairs(keys, vals=None, mode: int = 0, signed: Optional[bool] = None):
    """K4 on its own (qpe_gpu_sort_pairs): stable radix sort of (key, payload) pairs on the GPU.  keys: uint64 / uint32 /
    int32 / int64 array; mode 0 = (keys[i], vals[i]), 1 = (keys[i], i), 2 = the input read backwards (keys[n-1-i], n-1-i).
    Returns (sorted keys, payloads, digit passes run, device ms)."""
    keys = np.ascontiguousarray(keys)
    if keys.dtype not in (np.uint64, np.int64, np.uint32, np.int32):
        raise QpeError("sort_pairs: keys must be 32- or 64-bit integers")
    if signed is None:
        signed = keys.dtype.kind == "i"
    n = keys.shape[0]
    if mode == 0:
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        if vals.shape[0] != n:
            raise QpeError("sort_pairs: keys and vals differ in length")
    keys_out = np.empty_like(keys)
    vals_out = np.empty(n, dtype=np.uint32)
    passes, ms = C.c_int(0), C.c_double(0)
    rc = load_library().qpe_gpu_sort_pairs(keys.ctypes.data if n else None, vals.ctypes.data if (mode == 0 and n) else None, n,
                                           keys.dtype.itemsize, 1 if signed else 0, mode,
                                           keys_out.ctypes.data if n else None, vals_out.ctypes.data if n else None,
                                           C.byref(passes), C.byref(ms))
    if rc != 0:
        raise QpeError(f"sort_pairs failed (rc={rc}): " + (load_library().qpe_gpu_last_error() or b"").decode())
    return keys_out, vals_out, passes.value, ms.value


def _index_args(indexes):
    n = len(indexes)
    names = (C.c_char_p * max(n, 1))(*[a.encode() for a, _ in indexes])
    types = (C.c_int * max(n, 1))(*[t for _, t in indexes])
    return n, names, types


def column_mask(columns: Sequence[str]) -> int:
    m = 0
    for c in columns:
        m |= 1 << COLUMNS.index(c)
    return m


def where_text(statement: str) -> str:
    """The WHERE list our front end builds for `statement` (parity aid for the tokenizer restatement)."""
    lib = load_library()
    p = lib.qpe_sql_where_to_text(statement.encode())
    try:
        return C.string_at(p).decode()
    finally:
        lib.qpe_gpu_free(p)


class Engine:
    """One table resident in HBM as columns, plus its flattened indexes (struct engineS* underneath)."""

    def __init__(self, handle, lib):
        if not handle:
            raise QpeError("engine creation failed: " + (lib.qpe_gpu_last_error() or b"").decode())
        self._h = handle
        self._lib = lib

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def from_csv(cls, path: str, indexes=DEFAULT_INDEXES, table: str = "commands") -> "Engine":
        lib = load_library()
        n, names, types = _index_args(indexes)
        return cls(lib.initializeEngineGPU(n, names, types, path.encode(), table.encode()), lib)

    @classmethod
    def from_synth(cls, total_rows: int, n_rows: Optional[int] = None, row_base: int = 0, seed: int = 12345,
                   columns: Sequence[str] = COLUMNS, indexes=()) -> "Engine":
        lib = load_library()
        n, names, types = _index_args(indexes)
        if n_rows is None:
            n_rows = total_rows
        return cls(lib.qpe_gpu_engine_synth(total_rows, row_base, n_rows, seed, column_mask(columns), n, names, types),
                   lib)

    def close(self):
        if self._h:
            self._lib.destroyEngineGPU(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers --------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise QpeError(f"{what} failed (rc={rc}): " + (self._lib.qpe_gpu_last_error() or b"").decode())

    @property
    def num_rows(self) -> int:
        return int(self._lib.qpe_gpu_num_rows(self._h))

    def set_tile(self, tile_rows: int = 0, stages: int = 0):
        self._check(self._lib.qpe_gpu_set_tile(self._h, tile_rows, stages), "set_tile")

    def set_pipeline(self, segments: int = 0):
        """0 = automatic, 1 = one K1 + one K1c launch, n = n table segments with K1c(i) beside K1(i+1)"""
        self._check(self._lib.qpe_gpu_set_pipeline(self._h, segments), "set_pipeline")

    @property
    def stream(self) -> int:
        """cudaStream_t of the engine (all its kernels run there)"""
        return int(self._lib.qpe_gpu_stream(self._h))

    def set_timing(self, accumulate: bool):
        """start (and reset) / stop summing the device time of every match phase (see timing_totals)"""
        self._check(self._lib.qpe_gpu_set_timing(self._h, 1 if accumulate else 0), "set_timing")

    def timing_totals(self) -> dict:
        """CUDA-event times summed over the calls since set_timing(True), resolved after the fact"""
        t = (C.c_double * 4)()
        n = C.c_longlong()
        self._check(self._lib.qpe_gpu_timing_totals(self._h, t, C.byref(n)), "timing_totals")
        return {"kernel_ms": t[0], "scan_ms": t[1], "compact_ms": t[2], "post_ms": t[3], "calls": int(n.value)}

    def last_trace(self) -> list:
        """[compile, enqueue, sync, post-kernel] ms of the most recent match phase"""
        out = (C.c_double * 8)()
        self._check(self._lib.qpe_gpu_last_trace(self._h, out), "last_trace")
        return list(out)

    def fused_trace(self) -> np.ndarray:
        """per-CTA time stamps of the most recent fused scan (QPE_FUSE_TRACE=1), shape (ctas, 8), ns"""
        buf = np.zeros((1024, 8), dtype=np.uint64)
        n = self._lib.qpe_gpu_fused_trace(self._h, buf.ctypes.data, 1024)
        return buf[:n]

    def last_stats(self) -> dict:
        st = ScanStats()
        self._check(self._lib.qpe_gpu_last_stats(self._h, C.byref(st)), "last_stats")
        return st.as_dict()

    # ---- SQL surface ----------------------------------------------------------------------
    def run(self, statement: str, max_rows: int = 20) -> str:
        """run_test_query: execute one statement, return exactly what the reference would print."""
        p = self._lib.qpe_sql_run_to_text(self._h, statement.encode(), max_rows)
        if not p:
            raise QpeError("qpe_sql_run_to_text failed")
        try:
            return C.string_at(p).decode(errors="replace")
        finally:
            self._lib.qpe_gpu_free(p)

    def select(self, statement: str) -> Tuple[List[str], List[List[str]], float]:
        """executeQuerySelectGPU: (column names, rows of cell text, match-phase seconds)."""
        res = self._lib.qpe_sql_select(self._h, statement.encode())
        if not res:
            raise QpeError("not a SELECT statement: " + statement)
        try:
            r = res.contents
            if not r.success:
                raise QpeError("SELECT failed: " + (self._lib.qpe_gpu_last_error() or b"").decode())
            names = [r.columnNames[j].decode() for j in range(r.numColumns)]
            rows = [[r.data[i][j].decode(errors="replace") for j in range(r.numColumns)] for i in range(r.numRecords)]
            return names, rows, r.queryTime
        finally:
            self._lib.freeResultSet(res)

    def select_result_time(self, statement: str) -> int:
        """executeQuerySelectGPU -> resultSetS -> freeResultSet without touching the cells (timing aid): rows returned"""
        res = self._lib.qpe_sql_select(self._h, statement.encode())
        if not res:
            raise QpeError("not a SELECT statement: " + statement)
        try:
            if not res.contents.success:
                raise QpeError("SELECT failed: " + (self._lib.qpe_gpu_last_error() or b"").decode())
            return int(res.contents.numRecords)
        finally:
            self._lib.freeResultSet(res)

    def select_ids(self, statement: str, force_scan: bool = False) -> Tuple[np.ndarray, dict]:
        """Match phase only: row ids (table positions) in the reference's result order."""
        ids = C.c_void_p()
        n = C.c_size_t()
        st = ScanStats()
        rc = self._lib.qpe_sql_select_ids(self._h, statement.encode(), SCAN_FORCE if force_scan else 0, C.byref(ids),
                                          C.byref(n), C.byref(st))
        self._check(rc, "select_ids")
        try:
            out = np.ctypeslib.as_array(C.cast(ids, C.POINTER(C.c_uint32)), shape=(max(n.value, 1),))[:n.value].copy()
        finally:
            self._lib.qpe_gpu_free(ids)
        return out, st.as_dict()

    def select_ids_batch(self, statements: Sequence[str]) -> Tuple[List[np.ndarray], dict]:
        """Match phase of many SELECTs; each result equals select_ids(statement).  Full-scan queries share
        passes over the table (8 WHERE programs per pass, K9); index-path queries run one by one."""
        n = len(statements)
        arr = (C.c_char_p * max(n, 1))(*[s.encode() for s in statements])
        ids = (C.c_void_p * max(n, 1))()
        cnt = (C.c_size_t * max(n, 1))()
        st = ScanStats()
        rc = self._lib.qpe_sql_select_ids_batch(self._h, arr, n, ids, cnt, C.byref(st))
        self._check(rc, "select_ids_batch")
        out = []
        try:
            for q in range(n):
                m = cnt[q]
                out.append(np.ctypeslib.as_array(C.cast(ids[q], C.POINTER(C.c_uint32)), shape=(max(m, 1),))[:m].copy())
        finally:
            for q in range(n):
                self._lib.qpe_gpu_free(ids[q])
        return out, st.as_dict()

    def select_ids_into(self, statement: str, out: np.ndarray, force_scan: bool = False,
                        stats: bool = True) -> Tuple[int, dict]:
        """Match phase with the ids copied into a caller-owned (ideally pinned) uint32 buffer."""
        n = C.c_size_t()
        st = ScanStats() if stats else None
        rc = self._lib.qpe_sql_select_ids_into(self._h, statement.encode(), SCAN_FORCE if force_scan else 0,
                                               out.ctypes.data, out.size, C.byref(n), C.byref(st) if stats else None)
        self._check(rc, "select_ids_into")
        return int(n.value), (st.as_dict() if stats else None)

    def select_ids_device(self, statement: str, force_scan: bool = False, count_only: bool = False,
                          global_ids: bool = False, stats: bool = True):
        """Match phase, result left in HBM: (count, device pointer, stats).  stats=False skips the statistics
        (the event times then stay unresolved: see set_timing / timing_totals) and returns None for them."""
        cnt = C.c_ulonglong()
        dptr = C.c_void_p()
        st = ScanStats() if stats else None
        flags = ((SCAN_FORCE if force_scan else 0) | (SCAN_COUNT_ONLY if count_only else 0) |
                 (SCAN_GLOBAL_IDS if global_ids else 0))
        rc = self._lib.qpe_sql_select_ids_device(self._h, statement.encode(), flags, C.byref(cnt), C.byref(dptr),
                                                 C.byref(st) if stats else None)
        self._check(rc, "select_ids_device")
        return int(cnt.value), dptr.value, (st.as_dict() if stats else None)

    def select_segments(self, statement: str, global_ids: bool = True):
        """Index path of one shard, per segment: (used_index, [(keys int64, ids uint32), ...]).
        Used by sharding.sharded_select to merge shards into the reference's (key ASC, position DESC) order."""
        used = C.c_int()
        nseg = C.c_int()
        counts = (C.c_size_t * 32)()
        keys = C.c_void_p()
        ids = C.c_void_p()
        rc = self._lib.qpe_sql_select_segments(self._h, statement.encode(), 1 if global_ids else 0, C.byref(used),
                                               C.byref(nseg), counts, C.byref(keys), C.byref(ids))
        self._check(rc, "select_segments")
        try:
            total = sum(counts[s] for s in range(nseg.value))
            k = np.ctypeslib.as_array(C.cast(keys, C.POINTER(C.c_int64)), shape=(max(total, 1),))[:total].copy()
            d = np.ctypeslib.as_array(C.cast(ids, C.POINTER(C.c_uint32)), shape=(max(total, 1),))[:total].copy()
        finally:
            self._lib.qpe_gpu_free(keys)
            self._lib.qpe_gpu_free(ids)
        out, o = [], 0
        for s in range(nseg.value):
            out.append((k[o:o + counts[s]], d[o:o + counts[s]]))
            o += counts[s]
        return bool(used.value), out

    def scan_count(self, statement: str) -> Tuple[int, dict]:
        """First half of a split full scan: K1 only -> match count (the bitmap stays in HBM)."""
        cnt = C.c_ulonglong()
        st = ScanStats()
        self._check(self._lib.qpe_sql_scan_count(self._h, statement.encode(), C.byref(cnt), C.byref(st)), "scan_count")
        return int(cnt.value), st.as_dict()

    def select_ids_to(self, statement: str, dst_device_ptr: int, capacity: int, global_ids: bool = False):
        """Full scan (K1 + K1c) with the ids stored at `dst_device_ptr` (own or peer memory, `capacity` ids).
        Returns (count, stats, fits): when the result does not fit nothing is stored past the capacity."""
        cnt = C.c_ulonglong()
        st = ScanStats()
        rc = self._lib.qpe_sql_select_ids_to(self._h, statement.encode(), dst_device_ptr, capacity,
                                             1 if global_ids else 0, C.byref(cnt), C.byref(st))
        if rc not in (0, -5):
            self._check(rc, "select_ids_to")
        return int(cnt.value), st.as_dict(), rc == 0

    def compact_to(self, dst_device_ptr: int, global_ids: bool = False) -> dict:
        """Second half: K1c writes the row ids of the last scan_count to `dst_device_ptr`, which may be
        another GPU's buffer mapped through CUDA IPC (the ids then cross NVLink as the kernel's stores)."""
        st = ScanStats()
        self._check(self._lib.qpe_gpu_compact_to(self._h, dst_device_ptr, 1 if global_ids else 0, C.byref(st)),
                    "compact_to")
        return st.as_dict()

    def copy_from_device(self, dptr: int, n_items: int, dtype=np.uint32) -> np.ndarray:
        out = np.empty(n_items, dtype=dtype)
        if n_items:
            self._check(self._lib.qpe_gpu_copy_from_device(out.ctypes.data, dptr, out.nbytes), "copy_from_device")
        return out

    def match_mask(self, statement: str) -> Tuple[np.ndarray, int]:
        """DELETE's match mask (bit r = row r matches) without deleting."""
        n = self.num_rows
        words = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
        cnt = C.c_ulonglong()
        st = ScanStats()
        rc = self._lib.qpe_sql_match_mask(self._h, statement.encode(), words.ctypes.data, words.size, C.byref(cnt),
                                          C.byref(st))
        self._check(rc, "match_mask")
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)
        return bits, int(cnt.value)

    # ---- index surface --------------------------------------------------------------------
    def probe_batch(self, attribute: str, lo: np.ndarray, hi: np.ndarray) -> Tuple[np.ndarray, np.ndarray, dict]:
        """Batched findRange: inclusive [lo[q], hi[q]] -> (first, count) into the index order."""
        q = len(lo)
        is_u64 = attribute == "command_id"
        keys_lo = (KeyT * max(q, 1))()
        keys_hi = (KeyT * max(q, 1))()
        # bulk-fill through numpy views of the 16-byte structs
        dt = np.dtype([("type", np.int32), ("pad", np.int32), ("v", np.uint64)])
        vlo = np.frombuffer(keys_lo, dtype=dt, count=q)
        vhi = np.frombuffer(keys_hi, dtype=dt, count=q)
        if is_u64:
            vlo["type"] = 1
            vhi["type"] = 1
            vlo["v"] = np.asarray(lo, dtype=np.uint64)
            vhi["v"] = np.asarray(hi, dtype=np.uint64)
        else:
            vlo["type"] = 0
            vhi["type"] = 0
            vlo["v"] = np.asarray(lo, dtype=np.int32).astype(np.int64).astype(np.uint64) & np.uint64(0xffffffff)
            vhi["v"] = np.asarray(hi, dtype=np.int32).astype(np.int64).astype(np.uint64) & np.uint64(0xffffffff)
        first = np.zeros(max(q, 1), dtype=np.uint32)
        count = np.zeros(max(q, 1), dtype=np.uint32)
        st = ScanStats()
        rc = self._lib.qpe_gpu_probe_batch(self._h, attribute.encode(), keys_lo, keys_hi, q, first.ctypes.data,
                                           count.ctypes.data, C.byref(st))
        self._check(rc, "probe_batch")
        return first[:q], count[:q], st.as_dict()

    def probe_keys(self, attribute: str, lo, hi=None, sort: bool = False, first=None, count=None):
        """Batched probes with plain key arrays (qpe_gpu_probe_keys).  lo / hi: numpy arrays (uint64 for a u64 index,
        int32 for an int index; pinned ones from `pinned_array` are copied in place) or DeviceBuffer; hi=None: point
        probes.  first / count: numpy uint32 arrays or DeviceBuffer to receive the answers (allocated when None).
        Returns (first, count, stats dict)."""
        def ptr(x):
            return x.ptr if isinstance(x, DeviceBuffer) else x.ctypes.data
        q = (lo.n_bytes // (8 if attribute == "command_id" else 4)) if isinstance(lo, DeviceBuffer) else len(lo)
        if first is None:
            first = np.empty(q, dtype=np.uint32)
        if count is None:
            count = np.empty(q, dtype=np.uint32)
        st = ScanStats()
        rc = self._lib.qpe_gpu_probe_keys(self._h, attribute.encode(), ptr(lo), ptr(hi) if hi is not None else None, q,
                                          ptr(first), ptr(count), 1 if sort else 0, C.byref(st))
        self._check(rc, "probe_keys")
        return first, count, st.as_dict()

    def index_slice(self, attribute: str, first: int, count: int) -> np.ndarray:
        out = np.zeros(max(count, 1), dtype=np.uint32)
        self._check(self._lib.qpe_gpu_index_slice(self._h, attribute.encode(), first, count, out.ctypes.data),
                    "index_slice")
        return out[:count]

    def index_slice_keys(self, attribute: str, first: int, count: int) -> np.ndarray:
        """keys of index entries [first, first + count) (int64; u64 keys as bits)"""
        out = np.zeros(max(count, 1), dtype=np.int64)
        self._check(self._lib.qpe_gpu_index_slice_keys(self._h, attribute.encode(), first, count, out.ctypes.data),
                    "index_slice_keys")
        return out[:count]

    # ---- column access --------------------------------------------------------------------
    def column_width(self, attribute: str) -> int:
        w = C.c_uint()
        self._check(self._lib.qpe_gpu_fetch_column(self._h, attribute.encode(), 0, 0, None, C.byref(w)), "fetch_column")
        return int(w.value)

    def fetch_column(self, attribute: str, first_row: int = 0, n_rows: Optional[int] = None) -> np.ndarray:
        """Device column -> numpy: numeric columns as their dtype, text columns as (n, width) uint8."""
        if n_rows is None:
            n_rows = self.num_rows - first_row
        w = self.column_width(attribute)
        raw = np.zeros(max(n_rows * w, 1), dtype=np.uint8)
        self._check(self._lib.qpe_gpu_fetch_column(self._h, attribute.encode(), first_row, n_rows, raw.ctypes.data,
                                                   None), "fetch_column")
        raw = raw[:n_rows * w]
        if attribute in NUMERIC_DTYPES:
            return raw.view(NUMERIC_DTYPES[attribute])
        return raw.reshape(n_rows, w)

    def write_csv(self, path: str):
        self._check(self._lib.qpe_gpu_write_csv(self._h, path.encode()), "write_csv")
