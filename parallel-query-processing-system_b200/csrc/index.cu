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
// Build (K4): one stable LSD radix sort of the keys taken in REVERSE table order (stable + reversed
// input == position-descending among equal keys) -- our own sort (radix_sort.cu): its first pass reads the table
// column backwards with the position as payload, byte positions that are the same in every key are skipped --
// then the separator levels are sampled.

#include <cuda_runtime.h>

#include <type_traits>

#include "index.cuh"
#include "radix_sort.cuh"
#include "scan_kernels.cuh"

namespace qpe {

constexpr int kFanout = 16;

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
        if (ix->level_buf[l]) cudaFree(ix->level_buf[l]);
    ix->keys = nullptr;
    ix->perm = nullptr;
    for (int l = 0; l < kMaxIndexLevels; ++l) {
        ix->level[l] = nullptr;
        ix->level_buf[l] = nullptr;
        ix->level_cap[l] = 0;
        ix->level_cnt[l] = 0;
    }
    ix->n_levels = 0;
    ix->cap = 0;
    ix->n = 0;
}

// keys / perm for `cap` entries and every separator level such a tree can have (height h holds every 16^(h+1)-th key)
template <typename K>
static cudaError_t alloc_index(DevIndex *ix, long long n) {
    if (ix->cap >= n && ix->keys != nullptr) return cudaSuccess;
    index_free(ix);
    const long long cap = n + n / 8 + 1024;
    cudaError_t e;
    if ((e = cudaMalloc(&ix->keys, static_cast<size_t>(cap) * sizeof(K))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&ix->perm, static_cast<size_t>(cap) * sizeof(uint32_t))) != cudaSuccess) return e;
    long long below = cap;
    for (int h = 0; h < kMaxIndexLevels && below > kFanout; ++h) {
        const long long cnt = (below + kFanout - 1) / kFanout;
        if ((e = cudaMalloc(&ix->level_buf[h], static_cast<size_t>(cnt) * sizeof(K))) != cudaSuccess) return e;
        ix->level_cap[h] = cnt;
        below = cnt;
    }
    ix->cap = cap;
    return cudaSuccess;
}

// separator levels for the current ix->n keys, finest first into level_buf[], then viewed coarsest first
template <typename K>
static cudaError_t build_levels(DevIndex *ix, cudaStream_t stream, int *launches) {
    const K *below = static_cast<const K *>(ix->keys);
    long long n_below = ix->n;
    int nl = 0;
    long long cnt[kMaxIndexLevels];
    while (n_below > kFanout && nl < kMaxIndexLevels) {
        const long long n_level = (n_below + kFanout - 1) / kFanout;
        if (!ix->level_buf[nl] || n_level > ix->level_cap[nl]) return cudaErrorInvalidValue;
        K *lvl = static_cast<K *>(ix->level_buf[nl]);
        sample_level_kernel<K><<<grid_for(n_level, 256), 256, 0, stream>>>(below, n_below, lvl, n_level);
        ++*launches;
        cnt[nl] = n_level;
        ++nl;
        below = lvl;
        n_below = n_level;
    }
    for (int l = 0; l < kMaxIndexLevels; ++l) {
        ix->level[l] = nullptr;
        ix->level_cnt[l] = 0;
    }
    for (int l = 0; l < nl; ++l) {
        ix->level[l] = ix->level_buf[nl - 1 - l];
        ix->level_cnt[l] = cnt[nl - 1 - l];
    }
    ix->n_levels = nl;
    ix->fanout = kFanout;
    return cudaGetLastError();
}

template <typename K>
static cudaError_t build_typed(DevIndex *ix, const K *col, long long n, cudaStream_t stream, int *launches) {
    cudaError_t e;
    if ((e = alloc_index<K>(ix, n)) != cudaSuccess) return e;
    ix->n = n;
    ix->n_levels = 0;
    ix->fanout = kFanout;
    if (n == 0) return cudaSuccess;

    // int keys order as signed values; the sort itself works on the unsigned bit patterns
    using U = typename std::conditional<sizeof(K) == 8, unsigned long long, uint32_t>::type;
    void *scratch = nullptr;
    const size_t scratch_bytes = radix_sort_scratch_bytes(n, static_cast<int>(sizeof(K)));
    if ((e = cudaMalloc(&scratch, scratch_bytes)) != cudaSuccess) return e;
    e = radix_sort_pairs<U>(reinterpret_cast<const U *>(col), nullptr, kSortReverseIota, std::is_signed<K>::value,
                            static_cast<U *>(ix->keys), ix->perm, n, scratch, scratch_bytes, stream, launches);
    if (e == cudaSuccess) e = build_levels<K>(ix, stream, launches);
    const cudaError_t es = cudaStreamSynchronize(stream);
    cudaFree(scratch);
    return e != cudaSuccess ? e : es;
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
// Index maintenance without a re-sort (the reference's insert / delete, engine/bplus.c:723-740,
// :1022-1051, called per row from executeEngine-serial.c:599-614, :661-665).
//
// INSERT: the new row is the table's LAST row, so among equal keys it goes to the FRONT of its key run
// (position DESC): slot p = lower_bound(keys, key).  K3 finds p (it stays on the device), the tail [p, n) moves one
// slot to the right through a scratch copy, the entry is written, the separator levels are re-sampled.
// DELETE: the table's stable compaction renumbers row r to remap[r] (monotonic), so the (key ASC, position DESC) order
// of the surviving entries is unchanged: one ordered filter pass (ballot / popc + decoupled look-back per 2 Ki entries)
// drops the deleted rows' entries and writes the new row ids.
// Both are O(n) streaming passes per index and allocate nothing.
// ------------------------------------------------------------------------------------------
template <typename K>
__global__ void tail_to_scratch_kernel(const K *__restrict__ keys, const uint32_t *__restrict__ perm, long long n,
                                       const uint32_t *__restrict__ p_ptr, K *__restrict__ s_keys,
                                       uint32_t *__restrict__ s_perm) {
    const long long p = *p_ptr;
    for (long long i = p + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        s_keys[i - p] = keys[i];
        s_perm[i - p] = perm[i];
    }
}
template <typename K>
__global__ void insert_entry_kernel(K *__restrict__ keys, uint32_t *__restrict__ perm, long long n,
                                    const uint32_t *__restrict__ p_ptr, const K *__restrict__ s_keys,
                                    const uint32_t *__restrict__ s_perm, const K *__restrict__ key_ptr, uint32_t row) {
    const long long p = *p_ptr;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        keys[p] = *key_ptr;
        perm[p] = row;
    }
    for (long long i = p + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        keys[i + 1] = s_keys[i - p];
        perm[i + 1] = s_perm[i - p];
    }
}

size_t index_scratch_bytes(const DevIndex &ix, long long n_entries) {
    const size_t ksz = ix.type == T_U64 ? 8 : 4;
    return ((static_cast<size_t>(n_entries) * ksz + 255) & ~size_t(255)) + ((static_cast<size_t>(n_entries) * 4 + 255) & ~size_t(255)) + 256;
}

template <typename K>
static cudaError_t insert_typed(DevIndex *ix, const K *col, long long row, void *scratch, uint32_t *d_first,
                                uint32_t *d_count, cudaStream_t stream, int *launches) {
    const long long n = ix->n;  // entries before the insert
    K *keys = static_cast<K *>(ix->keys);
    const K *key_ptr = col + row;  // the new row's key, read on the device
    cudaError_t e = index_probe(*ix, key_ptr, key_ptr, 1, d_first, d_count, stream);  // first = lower_bound(key)
    if (e != cudaSuccess) return e;
    ++*launches;
    K *s_keys = static_cast<K *>(scratch);
    uint32_t *s_perm = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(scratch) + ((static_cast<size_t>(n) * sizeof(K) + 255) & ~size_t(255)));
    if (n > 0) {
        tail_to_scratch_kernel<K><<<grid_for(n, 256), 256, 0, stream>>>(keys, ix->perm, n, d_first, s_keys, s_perm);
        ++*launches;
    }
    insert_entry_kernel<K><<<grid_for(n > 0 ? n : 1, 256), 256, 0, stream>>>(keys, ix->perm, n, d_first, s_keys, s_perm, key_ptr,
                                                                         static_cast<uint32_t>(row));
    ++*launches;
    ix->n = n + 1;
    return build_levels<K>(ix, stream, launches);
}

cudaError_t index_insert_row(DevIndex *ix, const DevTable &t, long long row, void *scratch, uint32_t *d_first,
                             uint32_t *d_count, cudaStream_t stream, int *launches) {
    int dummy = 0;
    if (!launches) launches = &dummy;
    if (!ix->usable || ix->dirty) return cudaSuccess;   // nothing to maintain: built on next use
    if (ix->n + 1 > ix->cap) {                          // out of head-room: rebuild (allocates)
        ix->dirty = true;
        return cudaSuccess;
    }
    const DevColumn &c = t.col[ix->col];
    if (!c.d) return cudaErrorInvalidValue;
    if (ix->type == T_U64)
        return insert_typed<unsigned long long>(ix, reinterpret_cast<const unsigned long long *>(c.d), row, scratch, d_first,
                                                d_count, stream, launches);
    return insert_typed<int>(ix, reinterpret_cast<const int *>(c.d), row, scratch, d_first, d_count, stream, launches);
}

// remap[old row] = new row, 0xffffffff for a deleted row: scatter of the keep list (remap pre-set to 0xff)
__global__ void remap_kernel(const uint32_t *__restrict__ keep, long long n_keep, uint32_t *__restrict__ remap) {
    for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < n_keep;
         j += static_cast<long long>(gridDim.x) * blockDim.x)
        remap[keep[j]] = static_cast<uint32_t>(j);
}
cudaError_t index_build_remap(const uint32_t *d_keep, long long n_keep, long long n_old, uint32_t *d_remap,
                              cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_remap, 0xff, static_cast<size_t>(n_old) * 4, stream);
    if (e != cudaSuccess) return e;
    if (n_keep > 0) remap_kernel<<<grid_for(n_keep, 256), 256, 0, stream>>>(d_keep, n_keep, d_remap);
    return cudaGetLastError();
}

constexpr int kIdxFilterThreads = 256;
constexpr int kIdxFilterItems = 8;
constexpr int kIdxFilterTile = kIdxFilterThreads * kIdxFilterItems;  // 2048 entries per CTA

__device__ __forceinline__ unsigned long long ix_ld_desc(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ix_st_desc(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ordered filter of the index entries: tiles are claimed in order (atomic counter), so every predecessor of a tile is
// resident or finished and the look-back cannot deadlock.  desc[tile] = epoch << 34 | state << 32 | value.
template <typename K>
__global__ void __launch_bounds__(kIdxFilterThreads)
    index_filter_kernel(const K *__restrict__ keys, const uint32_t *__restrict__ perm, long long n,
                        const uint32_t *__restrict__ remap, K *__restrict__ out_keys, uint32_t *__restrict__ out_perm,
                        unsigned long long *__restrict__ desc, unsigned int *__restrict__ tile_counter,
                        unsigned long long *__restrict__ total_out, uint32_t epoch) {
    __shared__ long long s_tile;
    __shared__ uint32_t s_wcnt[kIdxFilterItems][kIdxFilterThreads / 32];
    __shared__ uint32_t s_woff[kIdxFilterItems][kIdxFilterThreads / 32];
    __shared__ uint32_t s_excl;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr int kWarps = kIdxFilterThreads / 32;
    const long long n_tiles = (n + kIdxFilterTile - 1) / kIdxFilterTile;
    if (tid == 0) s_tile = static_cast<long long>(atomicAdd(tile_counter, 1u));
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= n_tiles) return;
    K key[kIdxFilterItems];
    uint32_t nid[kIdxFilterItems];
    uint32_t bal[kIdxFilterItems];
#pragma unroll
    for (int it = 0; it < kIdxFilterItems; ++it) {
        const long long i = tile * kIdxFilterTile + it * kIdxFilterThreads + tid;
        uint32_t v = 0xffffffffu;
        key[it] = K(0);
        if (i < n) {
            key[it] = keys[i];
            v = __ldg(remap + perm[i]);
        }
        nid[it] = v;
        bal[it] = __ballot_sync(0xffffffffu, v != 0xffffffffu);
        if (lane == 0) s_wcnt[it][warp] = __popc(bal[it]);
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive offsets of the (item, warp) groups in entry order, tile total, look-back
        uint32_t run = 0;
        for (int g0 = 0; g0 < kIdxFilterItems * kWarps; g0 += 32) {
            const int gi = g0 + static_cast<int>(lane);
            const uint32_t c = gi < kIdxFilterItems * kWarps ? s_wcnt[gi / kWarps][gi % kWarps] : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= static_cast<uint32_t>(d)) inc += t;
            }
            if (gi < kIdxFilterItems * kWarps) s_woff[gi / kWarps][gi % kWarps] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        const uint32_t total = run;
        const unsigned long long tag = static_cast<unsigned long long>(epoch) << 34;
        if (lane == 0) ix_st_desc(desc + tile, tag | (static_cast<unsigned long long>(tile == 0 ? 2u : 1u) << 32) | total);
        uint32_t excl = 0;
        long long idx = tile - 1;
        while (idx >= 0) {
            const long long mine = idx - static_cast<long long>(lane);
            uint32_t state = 2u, val = 0;
            if (mine >= 0) {
                unsigned long long d = ix_ld_desc(desc + mine);
                while ((d >> 34) != epoch) {
                    __nanosleep(64);
                    d = ix_ld_desc(desc + mine);
                }
                state = static_cast<uint32_t>(d >> 32) & 3u;
                val = static_cast<uint32_t>(d);
            }
            const uint32_t pm = __ballot_sync(0xffffffffu, state == 2u);
            uint32_t contrib = val;
            if (pm) {
                const uint32_t first = static_cast<uint32_t>(__ffs(pm) - 1);
                if (lane > first) contrib = 0;
            }
            excl += __reduce_add_sync(0xffffffffu, contrib);
            if (pm) break;
            idx -= 32;
        }
        if (lane == 0) {
            ix_st_desc(desc + tile, tag | (2ull << 32) | (excl + total));
            s_excl = excl;
            if (tile == n_tiles - 1) *total_out = static_cast<unsigned long long>(excl) + total;
        }
    }
    __syncthreads();
    const uint32_t excl = s_excl;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < kIdxFilterItems; ++it)
        if ((bal[it] >> lane) & 1u) {
            const uint32_t o = excl + s_woff[it][warp] + __popc(bal[it] & lt);
            out_keys[o] = key[it];
            out_perm[o] = nid[it];
        }
}

template <typename K>
static cudaError_t delete_typed(DevIndex *ix, const uint32_t *d_remap, long long n_new, void *scratch,
                                unsigned long long *desc, unsigned int *d_counter, unsigned long long *d_total,
                                uint32_t epoch, cudaStream_t stream, int *launches) {
    const long long n = ix->n;
    K *keys = static_cast<K *>(ix->keys);
    K *s_keys = static_cast<K *>(scratch);
    uint32_t *s_perm = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(scratch) + ((static_cast<size_t>(n) * sizeof(K) + 255) & ~size_t(255)));
    cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    const long long n_tiles = (n + kIdxFilterTile - 1) / kIdxFilterTile;
    if (n_tiles > 0) {
        index_filter_kernel<K><<<static_cast<unsigned>(n_tiles), kIdxFilterThreads, 0, stream>>>(
            keys, ix->perm, n, d_remap, s_keys, s_perm, desc, d_counter, d_total, epoch);
        ++*launches;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        // every surviving entry belongs to a surviving row: the count is the new table size
        if ((e = cudaMemcpyAsync(keys, s_keys, static_cast<size_t>(n_new) * sizeof(K), cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(ix->perm, s_perm, static_cast<size_t>(n_new) * 4, cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
    }
    ix->n = n_new;
    return build_levels<K>(ix, stream, launches);
}

long long index_filter_tiles(long long n) { return (n + kIdxFilterTile - 1) / kIdxFilterTile; }

cudaError_t index_apply_delete(DevIndex *ix, const uint32_t *d_remap, long long n_new, void *scratch,
                               unsigned long long *desc, unsigned int *d_counter, unsigned long long *d_total,
                               uint32_t epoch, cudaStream_t stream, int *launches) {
    int dummy = 0;
    if (!launches) launches = &dummy;
    if (!ix->usable || ix->dirty) return cudaSuccess;
    if (ix->type == T_U64)
        return delete_typed<unsigned long long>(ix, d_remap, n_new, scratch, desc, d_counter, d_total, epoch, stream, launches);
    return delete_typed<int>(ix, d_remap, n_new, scratch, desc, d_counter, d_total, epoch, stream, launches);
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
// (Round 2 tried ONE descent per probe: the lower-bound descent ends in the leaf node with the first key >= lo, and when a
// key of that node, or the next node's first key -- a shuffle away in the last separator node -- is ordered after hi, the
// upper bound is a ballot away; a second descent only for the probes that need one, 4 probes per half-warp.  Point probes on
// unique keys 3.25 vs 3.03 G probes/s, but ranges 2.57 vs 3.04 and a heavy-duplicate index 3.4 vs 4.0: the second descent
// then runs AFTER the first instead of beside it, and the kernel is bound by latency, not by loads.  Dropped;
// tests/test_gpu_scale.py::test_probe_node_boundaries, written for it, stayed.)
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

// ------------------------------------------------------------------------------------------
// K3s: the probes of ONE indexed SELECT -- one warp per (top-level condition x index) segment, possibly over
// different indexes and key types, keys passed as kernel parameters -- leaving the candidate-segment table
// (first entry and running total per segment: executeEngine-serial.c:358-459 concatenates the findRange results
// in this order) in device memory for K1g.  The host never sees first / count: no synchronisation between
// the probes and the filter.
// ------------------------------------------------------------------------------------------
struct SegProbeOne {
    const void *keys;
    const uint32_t *perm;
    long long n;
    const void *level[kMaxIndexLevels];
    long long level_cnt[kMaxIndexLevels];
    unsigned long long lo, hi;   // int keys: the value in the low 32 bits
    int n_levels;
    int is_u64;
};
struct SegProbeParams {
    int n_seg;
    int pad;
    CandSegments *out;
    SegProbeOne seg[kMaxSegments];
};

template <typename K>
__device__ __forceinline__ long long descend_half(const SegProbeOne &sg, K key, bool upper, uint32_t sub, uint32_t half) {
    long long g = 0;
    for (int l = 0; l <= sg.n_levels; ++l) {  // warp-uniform trip count
        const bool leaf = (l == sg.n_levels);
        const K *arr = static_cast<const K *>(leaf ? sg.keys : sg.level[l]);
        const long long cnt = leaf ? sg.n : sg.level_cnt[l];
        const long long i = g * kFanout + sub;
        const K v = i < cnt ? __ldg(arr + i) : K(0);
        const bool less = i < cnt && (upper ? v <= key : v < key);
        const uint32_t b = (__ballot_sync(0xffffffffu, less) >> (16u * half)) & 0xffffu;
        const int r = __popc(b);
        g = g * kFanout + (leaf ? r : (r > 0 ? r - 1 : 0));
    }
    return g;
}

__global__ void __launch_bounds__(32 * kMaxSegments) probe_segments_kernel(const __grid_constant__ SegProbeParams p) {
    __shared__ long long s_first[kMaxSegments], s_count[kMaxSegments];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t sub = lane & 15u, half = lane >> 4;   // lanes 0-15: lower_bound(lo), lanes 16-31: upper_bound(hi)
    if (warp < static_cast<uint32_t>(p.n_seg)) {
        const SegProbeOne &sg = p.seg[warp];
        long long pos;
        if (sg.is_u64)
            pos = descend_half<unsigned long long>(sg, half ? sg.hi : sg.lo, half != 0, sub, half);
        else
            pos = descend_half<int>(sg, static_cast<int>(static_cast<uint32_t>(half ? sg.hi : sg.lo)), half != 0, sub, half);
        const long long lo_pos = __shfl_sync(0xffffffffu, pos, 0), hi_pos = __shfl_sync(0xffffffffu, pos, 16);
        if (lane == 0) {
            s_first[warp] = lo_pos;
            s_count[warp] = hi_pos > lo_pos ? hi_pos - lo_pos : 0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        CandSegments *o = p.out;
        o->n_seg = p.n_seg;
        long long run = 0;
        o->vstart[0] = 0;
        for (int s = 0; s < p.n_seg; ++s) {
            o->perm[s] = p.seg[s].perm;
            o->first[s] = s_first[s];
            run += s_count[s];
            o->vstart[s + 1] = run;
        }
    }
}

cudaError_t index_probe_segments(const DevIndex *const *ix, const unsigned long long *lo, const unsigned long long *hi,
                                 int n_seg, CandSegments *d_out, cudaStream_t stream) {
    if (n_seg < 1 || n_seg > kMaxSegments) return cudaErrorInvalidValue;
    SegProbeParams p{};
    p.n_seg = n_seg;
    p.out = d_out;
    for (int s = 0; s < n_seg; ++s) {
        const DevIndex &x = *ix[s];
        SegProbeOne &o = p.seg[s];
        o.keys = x.keys;
        o.perm = x.perm;
        o.n = x.n;
        o.n_levels = x.n_levels;
        for (int l = 0; l < x.n_levels; ++l) {
            o.level[l] = x.level[l];
            o.level_cnt[l] = x.level_cnt[l];
        }
        o.lo = lo[s];
        o.hi = hi[s];
        o.is_u64 = x.type == T_U64 ? 1 : 0;
    }
    probe_segments_kernel<<<1, 32 * n_seg, 0, stream>>>(p);
    return cudaGetLastError();
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
