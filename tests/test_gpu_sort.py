"""K4, the engine's own radix sort (csrc/radix_sort.cu), against numpy's stable sort -- bit-exact keys AND payloads.
The index order it produces through the index build is covered against the reference's B+ tree elsewhere
(test_gpu_index*.py, golden index_order_2k.json); here the kernel is driven directly through qpe_gpu_sort_pairs with the
shapes an index build never sees: ragged sizes around the tile sizes (6144 / 8192 pairs), keys that use all 64 bits,
negative ints, byte positions that never vary (skipped passes), one key value only."""
import numpy as np
import pytest

import support

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    p = support.load_pkg()
    if not p.gpu_available():
        raise RuntimeError("no CUDA device: the B200 engine has no CPU fallback")
    return p


def _expect(keys, mode, vals=None):
    n = len(keys)
    if mode == 2:
        k, v = keys[::-1], np.arange(n - 1, -1, -1, dtype=np.uint32)
    elif mode == 1:
        k, v = keys, np.arange(n, dtype=np.uint32)
    else:
        k, v = keys, vals
    order = np.argsort(k, kind="stable")
    return k[order], v[order]


SIZES = [0, 1, 31, 33, 6143, 6144, 6145, 8191, 8192, 8193, 100_003, 1_000_001]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("dtype", [np.uint64, np.int32])
def test_sort_ragged_sizes_all_modes(pkg, n, dtype):
    rng = np.random.default_rng(n + 7)
    if dtype == np.uint64:
        keys = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
        keys[: n // 3] &= np.uint64(0xffff)           # many duplicates among small keys
    else:
        keys = rng.integers(-(1 << 31), 1 << 31, size=n, dtype=np.int64).astype(np.int32)
        keys[: n // 3] = keys[: n // 3] % 7 - 3       # duplicates around zero, both signs
    vals = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    for mode in (0, 1, 2):
        got_k, got_v, passes, _ = pkg.sort_pairs(keys, vals if mode == 0 else None, mode=mode)
        want_k, want_v = _expect(keys, mode, vals)
        assert np.array_equal(got_k, want_k), (n, dtype, mode)
        assert np.array_equal(got_v, want_v), (n, dtype, mode)
        if n > 1000:
            assert passes == keys.dtype.itemsize


def test_sort_skips_constant_bytes(pkg):
    """row ids below 2^24 stored as u64 (a command_id column): 3 digit passes, not 8; all keys equal: one pass that
    keeps (mode 1) or reverses (mode 2) the order"""
    rng = np.random.default_rng(3)
    n = 300_007
    keys = rng.integers(0, 1 << 24, size=n, dtype=np.uint64)
    k, v, passes, _ = pkg.sort_pairs(keys, mode=2)
    want_k, want_v = _expect(keys, 2)
    assert passes == 3 and np.array_equal(k, want_k) and np.array_equal(v, want_v)
    # only the top byte and byte 1 vary
    keys = (rng.integers(0, 256, size=n, dtype=np.uint64) << np.uint64(56)) | (rng.integers(0, 256, size=n, dtype=np.uint64) << np.uint64(8)) | np.uint64(0x11)
    k, v, passes, _ = pkg.sort_pairs(keys, mode=1)
    want_k, want_v = _expect(keys, 1)
    assert passes == 2 and np.array_equal(k, want_k) and np.array_equal(v, want_v)
    same = np.full(50_001, 42, dtype=np.int32)
    k, v, passes, _ = pkg.sort_pairs(same, mode=2)
    assert passes == 1 and np.array_equal(k, same) and np.array_equal(v, np.arange(50_000, -1, -1, dtype=np.uint32))
    k, v, passes, _ = pkg.sort_pairs(same, mode=1)
    assert np.array_equal(v, np.arange(50_001, dtype=np.uint32))


def test_sort_skewed_digits(pkg):
    """two key values only (a boolean-like column), and a run of one value inside random keys: the warp-aggregated
    counters see every lane of a warp on one digit"""
    rng = np.random.default_rng(11)
    n = 2_000_003
    keys = rng.integers(0, 2, size=n, dtype=np.int64).astype(np.int32) * 1000 - 1
    k, v, _, _ = pkg.sort_pairs(keys, mode=2)
    want_k, want_v = _expect(keys, 2)
    assert np.array_equal(k, want_k) and np.array_equal(v, want_v)
    keys = rng.integers(0, 1 << 40, size=n, dtype=np.uint64)
    keys[500_000:1_200_000] = np.uint64(123456789)
    k, v, _, _ = pkg.sort_pairs(keys, mode=2)
    want_k, want_v = _expect(keys, 2)
    assert np.array_equal(k, want_k) and np.array_equal(v, want_v)


def test_sort_large_is_sorted_and_a_permutation(pkg):
    """64 M pairs (size-independent properties): keys non-decreasing, payloads a permutation, keys[payload] == sorted
    keys, equal keys in descending position order (mode 2)"""
    n = 64_000_000
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 1 << 20, size=n, dtype=np.uint64)     # ~61 duplicates per key
    k, v, passes, ms = pkg.sort_pairs(keys, mode=2)
    assert passes == 3
    assert np.all(k[1:] >= k[:-1])
    assert np.array_equal(keys[v], k)
    same = k[1:] == k[:-1]
    assert np.all(v[1:][same] < v[:-1][same])
    seen = np.zeros(n, dtype=np.bool_)
    seen[v] = True
    assert seen.all()
    print(f"K4: {n} u64 pairs, {passes} passes, {ms:.2f} ms")
