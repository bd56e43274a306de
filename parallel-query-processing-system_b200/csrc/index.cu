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

// Latency-bound by construction: a descent is n_levels + 1 DEPENDENT node reads (~0.6 us each once the lower
// levels miss L1), so throughput = (descents in flight) / latency.  Each half-warp therefore owns kProbeU
// queries per iteration and walks the lower-bound and the upper-bound descent of all of them together:
// 2 * kProbeU independent loads in flight per lane, 4 * kProbeU descents per warp (the first version had one
// query per warp, lanes 0-15 on the lower bound and 16-31 on the upper: 2.2 G probes/s over 100 M keys).
constexpr int kProbeU = 2;

template <typename K>
__global__ void __launch_bounds__(256) probe_kernel(const __grid_constant__ ProbeParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t sub = lane & 15u;
    const uint32_t half = lane >> 4;  // which of the warp's two query slots this lane serves
    const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const K *lo = static_cast<const K *>(p.lo);
    const K *hi = static_cast<const K *>(p.hi);
    constexpr int kPerWarp = 2 * kProbeU;
    for (long long q0 = warp_global * kPerWarp; q0 < p.q; q0 += n_warps * kPerWarp) {
        long long qi[kProbeU];
        K key_lo[kProbeU], key_hi[kProbeU];
        long long g_lo[kProbeU], g_hi[kProbeU];  // node index within the current level, per descent
#pragma unroll
        for (int u = 0; u < kProbeU; ++u) {
            qi[u] = q0 + 2 * u + half;
            const bool live = qi[u] < p.q;
            key_lo[u] = live ? lo[qi[u]] : K(0);
            key_hi[u] = live ? hi[qi[u]] : K(0);
            g_lo[u] = 0;
            g_hi[u] = 0;
        }
        for (int l = 0; l <= p.n_levels; ++l) {  // warp-uniform trip count: every lane takes part in the ballots
            const bool leaf = (l == p.n_levels);
            const K *arr = static_cast<const K *>(leaf ? p.keys : p.level[l]);
            const long long cnt = leaf ? p.n : p.level_cnt[l];
            bool less_lo[kProbeU], less_hi[kProbeU];
#pragma unroll
            for (int u = 0; u < kProbeU; ++u) {  // all loads first: 2 * kProbeU independent reads in flight
                const long long i_lo = g_lo[u] * kFanout + sub, i_hi = g_hi[u] * kFanout + sub;
                const K v_lo = i_lo < cnt ? __ldg(arr + i_lo) : K(0);
                const K v_hi = i_hi < cnt ? __ldg(arr + i_hi) : K(0);
                less_lo[u] = i_lo < cnt && v_lo < key_lo[u];
                less_hi[u] = i_hi < cnt && v_hi <= key_hi[u];
            }
#pragma unroll
            for (int u = 0; u < kProbeU; ++u) {
                const uint32_t b_lo = (__ballot_sync(0xffffffffu, less_lo[u]) >> (16u * half)) & 0xffffu;
                const uint32_t b_hi = (__ballot_sync(0xffffffffu, less_hi[u]) >> (16u * half)) & 0xffffu;
                const int r_lo = __popc(b_lo), r_hi = __popc(b_hi);  // entries of the node ordered before the key
                // leaf: final position; above: descend into the last group whose first key is ordered before the key
                g_lo[u] = g_lo[u] * kFanout + (leaf ? r_lo : (r_lo > 0 ? r_lo - 1 : 0));
                g_hi[u] = g_hi[u] * kFanout + (leaf ? r_hi : (r_hi > 0 ? r_hi - 1 : 0));
            }
        }
        if (sub == 0) {
#pragma unroll
            for (int u = 0; u < kProbeU; ++u)
                if (qi[u] < p.q) {
                    p.first[qi[u]] = static_cast<uint32_t>(g_lo[u]);
                    p.count[qi[u]] = g_hi[u] > g_lo[u] ? static_cast<uint32_t>(g_hi[u] - g_lo[u]) : 0u;
                }
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
    long long warps = (q + 2 * kProbeU - 1) / (2 * kProbeU);
    long long blocks = (warps * 32 + threads - 1) / threads;
    // one resident wave: SM count x the occupancy of this instantiation (asked once)
    static int wave_u64 = 0, wave_i32 = 0;
    int &wave = ix.type == T_U64 ? wave_u64 : wave_i32;
    if (wave == 0) {
        int dev = 0, n_sm = 148, per_sm = 4;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (ix.type == T_U64)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, probe_kernel<unsigned long long>, threads, 0);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, probe_kernel<int>, threads, 0);
        wave = n_sm * (per_sm > 0 ? per_sm : 1);
    }
    if (blocks > wave) blocks = wave;
    if (ix.type == T_U64)
        probe_kernel<unsigned long long><<<static_cast<int>(blocks), threads, 0, stream>>>(p);
    else
        probe_kernel<int><<<static_cast<int>(blocks), threads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace qpe
