// index.cu -- flattened B+ tree: build (K4) and batched point/range probe (K3)
//
// The reference's tree (engine/bplus.c, ORDER 3) is only ever read through its leaf chain:
// findRange (:282-314) walks leaves from findLeaf(key_start) (:317-358) and emits every entry with
// key <= key_end.  The leaf chain holds entries by (key ASCENDING, table position DESCENDING):
// insert (:723-740) descends to the leftmost leaf that may hold the key (strict '>' at :340-342) and
// places the new entry before all equal keys (:475-477, :512-517); the index is built by inserting
// rows in table order (buildEngine-serial.c:46-53).  So the whole tree is equivalent to
//     perm  = row ids sorted by (key ASC, position DESC)          keys = key[perm]
// and findRange(lo, hi) == perm[lower_bound(keys, lo) .. upper_bound(keys, hi)).
//
// Layout in HBM: `keys` and `perm` (n entries each) plus implicit separator levels: level above an
// array = every kFanout-th element of it (the first key of each group of kFanout), stored
// contiguously, coarsest level <= kFanout entries.  A node is kFanout consecutive separators =
// one 128-byte line for u64 keys (64 bytes for int keys).
//
// Build: one stable LSD radix sort of the keys taken in REVERSE table order (stable + reversed
// input == position-descending among equal keys).  The sort is cub::DeviceRadixSort (library
// primitive); gather-in-reverse and the level construction are ours.

#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include "index.cuh"

namespace qpe {

constexpr int kFanout = 16;

template <typename K>
__global__ void reverse_keys_kernel(const K *__restrict__ col, long long n, K *__restrict__ keys_rev,
                                    uint32_t *__restrict__ pos_rev) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long src = n - 1 - i;
        keys_rev[i] = col[src];
        pos_rev[i] = static_cast<uint32_t>(src);
    }
}

template <typename K>
__global__ void sample_level_kernel(const K *__restrict__ below, long long n_below, K *__restrict__ level,
                                    long long n_level) {
    for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < n_level;
         g += static_cast<long long>(gridDim.x) * blockDim.x)
        level[g] = below[g * kFanout];
}

static int grid_for(long long work, int threads) {
    long long g = (work + threads - 1) / threads;
    if (g > 148ll * 16) g = 148ll * 16;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

void index_free(DevIndex *ix) {
    if (ix->keys) cudaFree(ix->keys);
    if (ix->perm) cudaFree(ix->perm);
    for (int l = 0; l < kMaxIndexLevels; ++l)
        if (ix->level[l]) cudaFree(ix->level[l]);
    ix->keys = nullptr;
    ix->perm = nullptr;
    for (int l = 0; l < kMaxIndexLevels; ++l) ix->level[l] = nullptr;
    ix->n_levels = 0;
    ix->cap = 0;
    ix->n = 0;
}

template <typename K>
static cudaError_t build_typed(DevIndex *ix, const K *col, long long n, cudaStream_t stream, int *launches) {
    cudaError_t e;
    // (re)allocate
    if (ix->cap < n || ix->keys == nullptr) {
        index_free(ix);
        long long cap = n + n / 8 + 1024;
        if ((e = cudaMalloc(&ix->keys, static_cast<size_t>(cap) * sizeof(K))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&ix->perm, static_cast<size_t>(cap) * sizeof(uint32_t))) != cudaSuccess) return e;
        ix->cap = cap;
    }
    for (int l = 0; l < kMaxIndexLevels; ++l)
        if (ix->level[l]) {
            cudaFree(ix->level[l]);
            ix->level[l] = nullptr;
        }
    ix->n_levels = 0;
    ix->n = n;
    ix->fanout = kFanout;
    if (n == 0) return cudaSuccess;

    K *keys_rev = nullptr;
    uint32_t *pos_rev = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    if ((e = cudaMalloc(&keys_rev, static_cast<size_t>(n) * sizeof(K))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&pos_rev, static_cast<size_t>(n) * sizeof(uint32_t))) != cudaSuccess) return e;
    reverse_keys_kernel<K><<<grid_for(n, 256), 256, 0, stream>>>(col, n, keys_rev, pos_rev);
    ++*launches;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_rev, static_cast<K *>(ix->keys), pos_rev, ix->perm, n, 0,
                                    static_cast<int>(sizeof(K) * 8), stream);
    if ((e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16)) != cudaSuccess) return e;
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_rev, static_cast<K *>(ix->keys), pos_rev, ix->perm, n, 0,
                                        static_cast<int>(sizeof(K) * 8), stream);
    if (e != cudaSuccess) return e;

    // separator levels, finest first, then reversed so level[0] is the coarsest
    void *lv[kMaxIndexLevels];
    long long cnt[kMaxIndexLevels];
    int nl = 0;
    const K *below = static_cast<const K *>(ix->keys);
    long long n_below = n;
    while (n_below > kFanout && nl < kMaxIndexLevels) {
        const long long n_level = (n_below + kFanout - 1) / kFanout;
        K *lvl = nullptr;
        if ((e = cudaMalloc(&lvl, static_cast<size_t>(n_level) * sizeof(K))) != cudaSuccess) return e;
        sample_level_kernel<K><<<grid_for(n_level, 256), 256, 0, stream>>>(below, n_below, lvl, n_level);
        ++*launches;
        lv[nl] = lvl;
        cnt[nl] = n_level;
        ++nl;
        below = lvl;
        n_below = n_level;
    }
    for (int l = 0; l < nl; ++l) {
        ix->level[l] = lv[nl - 1 - l];
        ix->level_cnt[l] = cnt[nl - 1 - l];
    }
    ix->n_levels = nl;
    e = cudaStreamSynchronize(stream);
    cudaFree(keys_rev);
    cudaFree(pos_rev);
    cudaFree(tmp);
    return e;
}

cudaError_t index_build(DevIndex *ix, const DevTable &t, cudaStream_t stream, int *launches) {
    int dummy = 0;
    if (!launches) launches = &dummy;
    if (!ix->usable) {
        ix->dirty = false;
        return cudaSuccess;
    }
    const DevColumn &c = t.col[ix->col];
    if (!c.d && t.n > 0) return cudaErrorInvalidValue;
    cudaError_t e;
    if (ix->type == T_U64)
        e = build_typed<unsigned long long>(ix, reinterpret_cast<const unsigned long long *>(c.d), t.n, stream,
                                            launches);
    else
        e = build_typed<int>(ix, reinterpret_cast<const int *>(c.d), t.n, stream, launches);
    if (e == cudaSuccess) ix->dirty = false;
    return e;
}

// ------------------------------------------------------------------------------------------
// K3: batched probe.  One warp per query: lanes 0-15 run lower_bound(lo), lanes 16-31 run
// upper_bound(hi); each 16-lane group reads one separator node per level (one coalesced line).
// ------------------------------------------------------------------------------------------
struct ProbeParams {
    const void *keys;
    long long n;
    int n_levels;
    const void *level[kMaxIndexLevels];
    long long level_cnt[kMaxIndexLevels];
    const void *lo;   // Q keys
    const void *hi;   // Q keys
    long long q;
    uint32_t *first;  // Q
    uint32_t *count;  // Q
};

template <typename K>
__global__ void __launch_bounds__(256) probe_kernel(const __grid_constant__ ProbeParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t sub = lane & 15u;
    const bool upper = lane >= 16u;  // second half-warp searches the upper bound
    const uint32_t half_mask = upper ? 0xffff0000u : 0x0000ffffu;
    const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const K *lo = static_cast<const K *>(p.lo);
    const K *hi = static_cast<const K *>(p.hi);
    for (long long qi = warp_global; qi < p.q; qi += n_warps) {
        const K key = upper ? hi[qi] : lo[qi];
        long long group = 0;  // node index within the current level
        for (int l = 0; l <= p.n_levels; ++l) {
            const bool leaf = (l == p.n_levels);
            const K *arr = static_cast<const K *>(leaf ? p.keys : p.level[l]);
            const long long cnt = leaf ? p.n : p.level_cnt[l];
            const long long base = group * kFanout;
            const long long idx = base + sub;
            bool less = false;
            if (idx < cnt) {
                const K v = __ldg(arr + idx);
                less = upper ? (v <= key) : (v < key);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, less) & half_mask;
            const int rank = __popc(bal);  // entries of this node ordered before the key
            if (leaf) {
                group = base + rank;  // final position
            } else {
                // descend into the last group whose first key is ordered before the key
                group = base + (rank > 0 ? rank - 1 : 0);
            }
        }
        const long long pos_lo = __shfl_sync(0xffffffffu, group, 0);
        const long long pos_hi = __shfl_sync(0xffffffffu, group, 16);
        if (lane == 0) {
            p.first[qi] = static_cast<uint32_t>(pos_lo);
            p.count[qi] = pos_hi > pos_lo ? static_cast<uint32_t>(pos_hi - pos_lo) : 0u;
        }
    }
}

cudaError_t index_probe(const DevIndex &ix, const void *d_lo, const void *d_hi, long long q, uint32_t *d_first,
                        uint32_t *d_count, cudaStream_t stream) {
    if (q <= 0) return cudaSuccess;
    ProbeParams p{};
    p.keys = ix.keys;
    p.n = ix.n;
    p.n_levels = ix.n_levels;
    for (int l = 0; l < ix.n_levels; ++l) {
        p.level[l] = ix.level[l];
        p.level_cnt[l] = ix.level_cnt[l];
    }
    p.lo = d_lo;
    p.hi = d_hi;
    p.q = q;
    p.first = d_first;
    p.count = d_count;
    const int threads = 256;
    long long warps = q;
    long long blocks = (warps * 32 + threads - 1) / threads;
    if (blocks > 148ll * 8) blocks = 148ll * 8;
    if (ix.type == T_U64)
        probe_kernel<unsigned long long><<<static_cast<int>(blocks), threads, 0, stream>>>(p);
    else
        probe_kernel<int><<<static_cast<int>(blocks), threads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace qpe
