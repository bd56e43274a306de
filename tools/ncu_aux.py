#!/usr/bin/env python3
"""Smallest program that launches every kernel beside K1f once or twice on realistic sizes (used under ncu):
K3 probe_kernel (1 M probes over 100 M keys), K3s + K1g (indexed SELECT), K1 + K1c (count / mask / two-kernel scan),
K9 (query batch), K7 + K2 (projection), K6 (CSV ingest), K8 (CSV rendering after DELETE), index maintenance."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np  # noqa: E402
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
IDX = (("command_id", 0), ("user_id", 1), ("risk_level", 1))
eng = pkg.Engine.from_synth(N, columns=["command_id", "user_id", "risk_level", "exit_code", "sudo_used", "shell_type", "host_name"],
                            indexes=IDX)
rng = np.random.default_rng(1)
q = 1_000_000
for attr, dt, keys in (("command_id", np.uint64, rng.integers(0, int(N * 1.1), q, dtype=np.uint64)),
                       ("user_id", np.int32, rng.integers(900, 3100, q).astype(np.int32))):
    lo = pkg.pinned_array(q, dt)
    lo[:] = keys
    d_lo, d_f, d_c = pkg.DeviceBuffer(lo.nbytes), pkg.DeviceBuffer(4 * q), pkg.DeviceBuffer(4 * q)
    pkg.load_library().qpe_gpu_copy_to_device(d_lo.ptr, lo.ctypes.data, lo.nbytes)
    for _ in range(2):
        _, _, st = eng.probe_keys(attr, d_lo, None, first=d_f, count=d_c)
    print("K3", attr, st["kernel_ms"])
for w in ("risk_level > 4 AND exit_code = 0", "user_id = 1001 OR (exit_code = 127)"):
    for _ in range(2):
        ids, st = eng.select_ids(f"SELECT command_id FROM Commands WHERE {w}")
    print("K3s+K1g", w, len(ids), st["kernel_ms"])
sql = "SELECT command_id FROM Commands WHERE (command_id < %d) AND (sudo_used = FALSE OR risk_level > 3)" % (N // 100)
print("K1 count", eng.scan_count(sql)[0])
eng.set_pipeline(1)
cnt, _, st = eng.select_ids_device(sql, force_scan=True)
print("K1 + K1c", cnt, st["kernel_ms"])
eng.set_pipeline(0)
batch = [f"SELECT command_id FROM Commands WHERE (exit_code = {k}) AND (sudo_used = FALSE)" for k in (0, 1, 2, 126, 127, 130, 137, 255)]
res, st = eng.select_ids_batch(batch)
print("K9", sum(len(r) for r in res), st["kernel_ms"])
names, rows, _ = eng.select("SELECT command_id, risk_level, sudo_used, host_name FROM Commands WHERE (command_id < 200000)")
print("K7 + K2", len(rows))
eng.close()
# K6 / K8 / index maintenance on a 1 M-row CSV
d = tempfile.mkdtemp(prefix="ncu_aux_")
csv = os.path.join(d, "t.csv")
gen = pkg.Engine.from_synth(1_000_000)
gen.write_csv(csv)
gen.close()
e2 = pkg.Engine.from_csv(csv)
print("K6 rows", e2.num_rows)
print(e2.run('INSERT INTO Commands VALUES (5000000, "echo x", "echo", "bash", 0, "2025-12-01T12:00:00.000Z", "FALSE", "/home/test", 1000, "testuser", "test-host", 1)', 5).strip().splitlines()[-1])
print(e2.run("DELETE FROM Commands WHERE risk_level = 3", 5).strip().splitlines()[-1])
e2.close()
