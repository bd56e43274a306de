// radix_sort.cu -- K4: our own stable LSD radix sort of (key, 32-bit payload) pairs, 8 bits per pass
//
// Used by the index build (index.cu: the flattened B+ tree is the table's keys sorted by (key ASC, position DESC),
// which a STABLE sort of the table read backwards gives; the reference builds the same order by inserting row after
// row, buildEngine-serial.c:46-53 -> engine/bplus.c:723-740) and by the sorted probe batch (probe_batch.cu).
//
// One pass over digit d = three launches:
//   radix_count_kernel    per tile of 512 x ITEMS keys: histogram of the digit (shared-memory adds; a warp whose keys
//                         share the digit adds once) -> tile_hist[digit][tile]
//   radix_scan_kernel     one CTA per digit: exclusive scan of its row over the tiles, total -> digit_total[digit]
//   radix_scatter_kernel  per tile: every warp ranks its keys in index order (same-digit lane mask through a shared-memory word + a warp-private
//                         counter row, no atomics), the tile is put in digit order in shared memory and leaves as runs of
//                         consecutive addresses: out = prefix(digit_total)[d] + tile_hist[d][tile] + rank within the tile
// HBM traffic per pass and pair: 2 x key read + payload read + key and payload write = 3 x sizeof(K) + 8 bytes.
// Byte positions in which ALL keys agree are skipped (key_bits_kernel: OR and AND of the keys, 16 bytes to the host):
// row ids below 2^32 stored as u64, small ints -- the index build of a u64 command_id column runs 4 passes, not 8.
// The first pass reads the table column itself (optionally backwards, payload = position): no key copy, no iota.

#include "radix_sort.cuh"

#include <type_traits>

namespace qpe {
namespace {

constexpr int kBins = 256;
constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;

template <typename K>
struct SortCfg;
template <>
struct SortCfg<unsigned long long> {
    static constexpr int kItems = 12;  // 6144 pairs per tile: 72 KB of stage + 34 KB of counters and masks, two CTAs per SM
};
template <>
struct SortCfg<uint32_t> {
    static constexpr int kItems = 16;  // 8192 pairs per tile: 64 KB + 34 KB
};

template <typename K>
__global__ void key_bits_kernel(const K *__restrict__ keys, long long n, unsigned long long *__restrict__ or_and) {
    K o = 0, a = ~static_cast<K>(0);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const K k = keys[i];
        o |= k;
        a &= k;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        o |= __shfl_xor_sync(0xffffffffu, o, off);
        a &= __shfl_xor_sync(0xffffffffu, a, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicOr(&or_and[0], static_cast<unsigned long long>(o));
        // a narrower key leaves the upper bits of the AND word set; the host masks the difference to the key's width
        atomicAnd(&or_and[1], static_cast<unsigned long long>(a) | ~static_cast<unsigned long long>(~static_cast<K>(0)));
    }
}

template <typename K>
__device__ __forceinline__ uint32_t digit_of(K key, K flip, int shift) {
    return static_cast<uint32_t>((key ^ flip) >> shift) & (kBins - 1);
}

// exclusive scan of one value per thread over the CTA (any number of warps <= 32); every thread must call it
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_scan, uint32_t *total_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    uint32_t x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < n_warps ? s_scan[lane] : 0;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += y;
        }
        s_scan[lane] = w;
    }
    __syncthreads();
    const uint32_t before = warp > 0 ? s_scan[warp - 1] : 0;
    if (total_out) *total_out = s_scan[n_warps - 1];
    __syncthreads();  // s_scan may be reused by the caller's next scan
    return before + x - v;
}

template <typename K, int ITEMS>
__global__ void __launch_bounds__(kSortThreads)
    radix_count_kernel(const K *__restrict__ keys_in, long long n, int reversed, int shift, K flip,
                       uint32_t *__restrict__ tile_hist, long long tiles) {
    __shared__ uint32_t hist[kBins];
    const long long tile = blockIdx.x;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < kBins) hist[threadIdx.x] = 0;
    __syncthreads();
    const long long base = tile * static_cast<long long>(kSortThreads * ITEMS);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const long long idx = base + i * kSortThreads + threadIdx.x;
        const bool valid = idx < n;
        uint32_t d = 0;
        if (valid) d = digit_of<K>(keys_in[reversed ? n - 1 - idx : idx], flip, shift);
        // a warp whose 32 keys share the digit (a byte that rarely changes) adds once; otherwise plain shared-memory adds
        const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
        if (__all_sync(0xffffffffu, valid && d == d0)) {
            if (lane == 0) atomicAdd(&hist[d0], 32u);
        } else if (valid) {
            atomicAdd(&hist[d], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < kBins) tile_hist[threadIdx.x * tiles + tile] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t *__restrict__ tile_hist, long long tiles,
                                                          uint32_t *__restrict__ digit_total) {
    __shared__ uint32_t s_scan[32];
    uint32_t *row = tile_hist + blockIdx.x * tiles;
    uint32_t running = 0;
    for (long long base = 0; base < tiles; base += 1024) {
        const long long i = base + threadIdx.x;
        const uint32_t v = i < tiles ? row[i] : 0;
        uint32_t total = 0;
        const uint32_t before = block_exclusive_scan(v, s_scan, &total);
        if (i < tiles) row[i] = running + before;
        running += total;
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = running;
}

// MODE: how the input is read (SortInput); FULL: every element of the tile exists (all tiles but possibly the last one:
// no bounds checks in the ranking loop); tile0: first tile of this launch.
template <typename K, int ITEMS, int MODE, bool FULL>
__global__ void __launch_bounds__(kSortThreads, 2)
    radix_scatter_kernel(const K *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, long long n, long long tile0,
                         int shift, K flip, const uint32_t *__restrict__ tile_hist, long long tiles,
                         const uint32_t *__restrict__ digit_total, K *__restrict__ keys_out,
                         uint32_t *__restrict__ vals_out) {
    constexpr int mode = MODE;
    constexpr int kTile = kSortThreads * ITEMS;
    extern __shared__ __align__(16) unsigned char sort_smem[];
    K *s_keys = reinterpret_cast<K *>(sort_smem);
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(s_keys + kTile);
    uint32_t *s_whist = s_vals + kTile;               // [warp][digit]: keys of the digit in the warps before
    uint32_t *s_gbase = s_whist + kSortWarps * kBins;  // where this tile's run of the digit starts in the output
    uint32_t *s_dstart = s_gbase + kBins;              // where the digit starts in the tile's staged order
    uint32_t *s_scan = s_dstart + kBins;               // 32 words
    uint32_t *s_wmask = s_scan + 32;                   // [warp][digit]: lanes of the warp whose current key has the digit

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long tile = tile0 + blockIdx.x;
    for (int i = threadIdx.x; i < kSortWarps * kBins; i += kSortThreads) {
        s_whist[i] = 0;
        s_wmask[i] = 0;
    }
    {
        const uint32_t tot = threadIdx.x < kBins ? digit_total[threadIdx.x] : 0;
        const uint32_t before = block_exclusive_scan(tot, s_scan, nullptr);
        if (threadIdx.x < kBins) s_gbase[threadIdx.x] = before + tile_hist[threadIdx.x * tiles + tile];
    }
    __syncthreads();

    // a warp owns ITEMS x 32 consecutive elements, taken 32 at a time: (i, lane) order == index order
    const long long wbase = tile * static_cast<long long>(kTile) + warp * (ITEMS * 32);
    uint32_t *wh = s_whist + warp * kBins;
    uint32_t *wm = s_wmask + warp * kBins;
    K key[ITEMS];
    uint32_t val[ITEMS];
    uint32_t rank2[(ITEMS + 1) / 2] = {0};  // ranks within the warp's part (< 512), two per register
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const long long idx = wbase + i * 32 + lane;
        const bool valid = FULL || idx < n;
        key[i] = 0;
        val[i] = 0;
        if (valid) {
            if (mode == kSortReverseIota) {
                key[i] = keys_in[n - 1 - idx];
                val[i] = static_cast<uint32_t>(n - 1 - idx);
            } else {
                key[i] = keys_in[idx];
                val[i] = mode == kSortIota ? static_cast<uint32_t>(idx) : vals_in[idx];
            }
        }
        // elements past the end (last tile only) rank behind everything: digit 255, highest indices
        const uint32_t d = valid ? digit_of<K>(key[i], flip, shift) : kBins - 1;
        // lanes with the same digit: every lane sets its bit in the warp's mask word of its digit (one shared-memory
        // atomic), reads the word back, and the first of them clears it for the next element.  (match.any gives the same
        // mask in one instruction, but at ~64 cycles per warp on sm_100: 7.2 ms per 100 M u64 row ids; eight ballots, one
        // per digit bit: ~48 instructions per element, 5.3 ms; this: ~8 instructions.)
        atomicOr(&wm[d], 1u << lane);
        __syncwarp();
        const uint32_t peers = wm[d];
        __syncwarp();
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (lane == leader) {
            wm[d] = 0;
            prev = wh[d];
            wh[d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        __syncwarp();
        rank2[i >> 1] |= (prev + __popc(peers & ((1u << lane) - 1u))) << (16 * (i & 1));
    }
    __syncthreads();
    {
        uint32_t cnt = 0;
        if (threadIdx.x < kBins) {
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                const uint32_t t = s_whist[w * kBins + threadIdx.x];
                s_whist[w * kBins + threadIdx.x] = cnt;
                cnt += t;
            }
        }
        const uint32_t before = block_exclusive_scan(cnt, s_scan, nullptr);
        if (threadIdx.x < kBins) s_dstart[threadIdx.x] = before;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const bool valid = FULL || wbase + i * 32 + lane < n;
        const uint32_t d = valid ? digit_of<K>(key[i], flip, shift) : kBins - 1;
        const uint32_t pos = s_dstart[d] + wh[d] + ((rank2[i >> 1] >> (16 * (i & 1))) & 0xffffu);
        s_keys[pos] = key[i];
        s_vals[pos] = val[i];
    }
    __syncthreads();
    const long long left = n - tile * static_cast<long long>(kTile);
    const int n_valid = (FULL || left >= kTile) ? kTile : static_cast<int>(left);
    for (int j = threadIdx.x; j < n_valid; j += kSortThreads) {
        const K k = s_keys[j];
        const uint32_t d = digit_of<K>(k, flip, shift);
        const size_t out = static_cast<size_t>(s_gbase[d]) + (j - s_dstart[d]);
        keys_out[out] = k;
        vals_out[out] = s_vals[j];
    }
}

constexpr size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename K>
struct SortScratch {
    unsigned long long *or_and;
    uint32_t *digit_total;
    uint32_t *tile_hist;
    K *alt_keys;
    uint32_t *alt_vals;
    long long tiles;
    size_t bytes;
};

template <typename K>
SortScratch<K> carve(void *scratch, long long n) {
    constexpr int kTile = kSortThreads * SortCfg<K>::kItems;
    SortScratch<K> s;
    s.tiles = (n + kTile - 1) / kTile;
    if (s.tiles < 1) s.tiles = 1;
    unsigned char *p = static_cast<unsigned char *>(scratch);
    size_t off = 0;
    s.or_and = reinterpret_cast<unsigned long long *>(p + off);
    off += 256;
    s.digit_total = reinterpret_cast<uint32_t *>(p + off);
    off += kBins * sizeof(uint32_t);
    s.tile_hist = reinterpret_cast<uint32_t *>(p + off);
    off = align_up(off + static_cast<size_t>(kBins) * s.tiles * sizeof(uint32_t), 256);
    s.alt_keys = reinterpret_cast<K *>(p + off);
    off = align_up(off + static_cast<size_t>(n) * sizeof(K), 256);
    s.alt_vals = reinterpret_cast<uint32_t *>(p + off);
    off = align_up(off + static_cast<size_t>(n) * sizeof(uint32_t), 256);
    s.bytes = off;
    return s;
}

}  // namespace

size_t radix_sort_scratch_bytes(long long n, int key_bytes) {
    if (n < 0) n = 0;
    return key_bytes == 8 ? carve<unsigned long long>(nullptr, n).bytes : carve<uint32_t>(nullptr, n).bytes;
}

template <typename K>
cudaError_t radix_sort_pairs(const K *keys_in, const uint32_t *vals_in, SortInput mode, bool signed_keys, K *keys_out,
                             uint32_t *vals_out, long long n, void *scratch, size_t scratch_bytes, cudaStream_t stream,
                             int *launches, int *passes_out) {
    constexpr int ITEMS = SortCfg<K>::kItems;
    constexpr int kTile = kSortThreads * ITEMS;
    constexpr int kKeyBytes = static_cast<int>(sizeof(K));
    int dummy = 0;
    if (!launches) launches = &dummy;
    if (passes_out) *passes_out = 0;
    if (n <= 0) return cudaSuccess;
    if (n > 0xffffffffll) return cudaErrorInvalidValue;  // positions and counters are 32-bit
    const SortScratch<K> s = carve<K>(scratch, n);
    if (!scratch || scratch_bytes < s.bytes) return cudaErrorInvalidValue;
    if (mode == kSortPairs && !vals_in) return cudaErrorInvalidValue;

    // which byte positions differ between any two keys
    cudaError_t e;
    const unsigned long long init[2] = {0ull, ~0ull};
    if ((e = cudaMemcpyAsync(s.or_and, init, sizeof(init), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
    long long rb = (n + 255) / 256;
    if (rb > 148 * 8) rb = 148 * 8;
    key_bits_kernel<K><<<static_cast<unsigned int>(rb), 256, 0, stream>>>(keys_in, n, s.or_and);
    ++*launches;
    unsigned long long got[2] = {0, 0};
    if ((e = cudaMemcpyAsync(got, s.or_and, sizeof(got), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    unsigned long long differ = got[0] ^ got[1];
    differ &= kKeyBytes < 8 ? (1ull << (8 * (kKeyBytes & 7))) - 1 : ~0ull;
    int shifts[8], n_pass = 0;
    for (int b = 0; b < kKeyBytes; ++b)
        if ((differ >> (8 * b)) & 0xff) shifts[n_pass++] = 8 * b;
    if (n_pass == 0) shifts[n_pass++] = 0;  // all keys equal: one pass still moves the pairs to the output, in order
    if (passes_out) *passes_out = n_pass;

    const K flip = signed_keys ? static_cast<K>(1) << (8 * kKeyBytes - 1) : 0;
    const size_t smem = static_cast<size_t>(kTile) * (sizeof(K) + sizeof(uint32_t)) +
                        (2 * static_cast<size_t>(kSortWarps) * kBins + 2 * kBins + 32) * sizeof(uint32_t);
    const unsigned int grid = static_cast<unsigned int>(s.tiles);
    const K *src_k = keys_in;
    const uint32_t *src_v = vals_in;
    int src_mode = mode;
    for (int k = 0; k < n_pass; ++k) {
        const bool to_out = ((n_pass - 1 - k) & 1) == 0;  // the last pass lands in keys_out / vals_out
        K *dst_k = to_out ? keys_out : s.alt_keys;
        uint32_t *dst_v = to_out ? vals_out : s.alt_vals;
        radix_count_kernel<K, ITEMS><<<grid, kSortThreads, 0, stream>>>(src_k, n, src_mode == kSortReverseIota ? 1 : 0,
                                                                        shifts[k], flip, s.tile_hist, s.tiles);
        radix_scan_kernel<<<kBins, 1024, 0, stream>>>(s.tile_hist, s.tiles, s.digit_total);
        // all full tiles in one launch (no bounds checks), the last, partial tile in one of its own
        const long long full_tiles = n / kTile;
        auto scatter = [&](auto mode_tag) -> cudaError_t {
            constexpr int M = decltype(mode_tag)::value;
            cudaError_t se;
            if (full_tiles > 0) {
                if ((se = cudaFuncSetAttribute(radix_scatter_kernel<K, ITEMS, M, true>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess)
                    return se;
                radix_scatter_kernel<K, ITEMS, M, true><<<static_cast<unsigned int>(full_tiles), kSortThreads, smem, stream>>>(
                    src_k, src_v, n, 0, shifts[k], flip, s.tile_hist, s.tiles, s.digit_total, dst_k, dst_v);
                ++*launches;
            }
            if (full_tiles < s.tiles) {
                if ((se = cudaFuncSetAttribute(radix_scatter_kernel<K, ITEMS, M, false>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess)
                    return se;
                radix_scatter_kernel<K, ITEMS, M, false><<<static_cast<unsigned int>(s.tiles - full_tiles), kSortThreads, smem, stream>>>(
                    src_k, src_v, n, full_tiles, shifts[k], flip, s.tile_hist, s.tiles, s.digit_total, dst_k, dst_v);
                ++*launches;
            }
            return cudaSuccess;
        };
        if (src_mode == kSortReverseIota)
            e = scatter(std::integral_constant<int, kSortReverseIota>());
        else if (src_mode == kSortIota)
            e = scatter(std::integral_constant<int, kSortIota>());
        else
            e = scatter(std::integral_constant<int, kSortPairs>());
        if (e != cudaSuccess) return e;
        *launches += 2;
        src_k = dst_k;
        src_v = dst_v;
        src_mode = kSortPairs;
    }
    return cudaGetLastError();
}

template cudaError_t radix_sort_pairs<unsigned long long>(const unsigned long long *, const uint32_t *, SortInput, bool,
                                                          unsigned long long *, uint32_t *, long long, void *, size_t,
                                                          cudaStream_t, int *, int *);
template cudaError_t radix_sort_pairs<uint32_t>(const uint32_t *, const uint32_t *, SortInput, bool, uint32_t *,
                                                uint32_t *, long long, void *, size_t, cudaStream_t, int *, int *);

}  // namespace qpe
