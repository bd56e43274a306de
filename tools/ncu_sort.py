#!/usr/bin/env python3
"""Smallest program that runs K4 (csrc/radix_sort.cu) once on a realistic column (used under ncu): N u64 row ids in
random order, read backwards with positions as payload -- the index build of a command_id column."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np  # noqa: E402
import support  # noqa: E402

pkg = support.load_pkg()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
keys = np.random.default_rng(1).permutation(N).astype(np.uint64)
k, v, passes, ms = pkg.sort_pairs(keys, mode=2)
print(f"K4: {N} u64 keys, {passes} passes, {ms:.3f} ms")
