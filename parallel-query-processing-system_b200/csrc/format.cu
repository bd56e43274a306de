// format.cu -- K7: SELECT projection rendered on the device
//
// Replaces the per-cell get_attribute_string_value + strdup loop of executeQuerySelectSerial
// (engine/serial/executeEngine-serial.c:216-248, :504-515; 1.9 s for SELECT * over 1 M rows in the
// reference).  For every matching row id and every projected NUMERIC column one thread gathers the value
// and writes its text -- "%llu" / "%d" / "true" | "false", NUL-terminated, exactly what sprintf produces --
// into a FIXED-WIDTH slot (24 / 16 / 8 bytes).  Text columns need no rendering at all: a device text cell
// is NUL-padded to its fixed width, so the plain gather (K2) of the column already is an array of C
// strings.  The host therefore receives, per projected column, one dense array of slots and builds
// data[i][j] = slots_j + i * slot_width_j by arithmetic: no per-cell strlen / memcpy / malloc anywhere.
// Bound: HBM random access for the gather (one 4-128 B cell per id), then PCIe for the slots.

#include "scan_kernels.cuh"

namespace qpe {

namespace {

// decimal digits of v into buf (most significant first); returns the length
__device__ __forceinline__ int render_u64(unsigned long long v, char *buf) {
    char tmp[20];
    int n = 0;
    do {
        const unsigned long long q = v / 10ull;
        tmp[n++] = static_cast<char>('0' + static_cast<int>(v - q * 10ull));
        v = q;
    } while (v);
    for (int i = 0; i < n; ++i) buf[i] = tmp[n - 1 - i];
    return n;
}

__global__ void __launch_bounds__(256) format_u64_kernel(const unsigned long long *__restrict__ col,
                                                         const uint32_t *__restrict__ ids, long long n,
                                                         uint2 *__restrict__ out) {  // 24 B per slot = 3 x uint2
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        alignas(8) char s[24] = {0};
        render_u64(__ldg(col + __ldg(ids + i)), s);  // <= 20 digits: s[20..23] stay NUL
        const uint2 *w = reinterpret_cast<const uint2 *>(s);
        out[3 * i] = w[0];
        out[3 * i + 1] = w[1];
        out[3 * i + 2] = w[2];
    }
}

__global__ void __launch_bounds__(256) format_i32_kernel(const int *__restrict__ col, const uint32_t *__restrict__ ids,
                                                         long long n, uint4 *__restrict__ out) {  // 16 B per slot
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        alignas(16) char s[16] = {0};
        const int v = __ldg(col + __ldg(ids + i));
        int p = 0;
        unsigned int mag = static_cast<unsigned int>(v);
        if (v < 0) {
            s[p++] = '-';
            mag = 0u - mag;  // INT_MIN safe
        }
        render_u64(mag, s + p);  // '-' + 10 digits + NUL = 12 <= 16
        out[i] = *reinterpret_cast<const uint4 *>(s);
    }
}

__global__ void __launch_bounds__(256) format_bool_kernel(const uint8_t *__restrict__ col, const uint32_t *__restrict__ ids,
                                                          long long n, uint2 *__restrict__ out) {  // 8 B per slot
    // little endian: "true\0\0\0\0" / "false\0\0\0"
    const uint2 t = make_uint2(0x65757274u, 0u);           // 't','r','u','e'
    const uint2 f = make_uint2(0x736c6166u, 0x00000065u);  // 'f','a','l','s' | 'e'
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = __ldg(col + __ldg(ids + i)) ? t : f;
}

int grid_of(long long work) {
    long long g = (work + 255) / 256;
    if (g > 148ll * 16) g = 148ll * 16;
    return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace

uint32_t format_slot_width(int col_type, uint32_t cell_width) {
    switch (col_type) {
        case T_U64: return 24;
        case T_I32: return 16;
        case T_BOOL: return 8;
        default: return cell_width;  // text: the NUL-padded cell itself
    }
}

cudaError_t format_launch(const uint8_t *col, int col_type, uint32_t cell_width, const uint32_t *ids, int64_t n,
                          uint8_t *out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    switch (col_type) {
        case T_U64:
            format_u64_kernel<<<grid_of(n), 256, 0, stream>>>(reinterpret_cast<const unsigned long long *>(col), ids, n,
                                                              reinterpret_cast<uint2 *>(out));
            break;
        case T_I32:
            format_i32_kernel<<<grid_of(n), 256, 0, stream>>>(reinterpret_cast<const int *>(col), ids, n,
                                                              reinterpret_cast<uint4 *>(out));
            break;
        case T_BOOL:
            format_bool_kernel<<<grid_of(n), 256, 0, stream>>>(col, ids, n, reinterpret_cast<uint2 *>(out));
            break;
        default:
            return gather_launch(col, cell_width, ids, n, out, stream);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K8: the table rendered as CSV text on the device
//
// executeQueryDeleteSerial rewrites the WHOLE data file after every DELETE, one fprintf per row
// (engine/serial/executeEngine-serial.c:683-706): "%llu,%s,%s,%s,%d,%s,%d,%s,%d,%s,%s,%d\n", no header, no
// quoting, sudo_used as 0/1.  Here: csv_len_kernel (bytes of every row) -> exclusive scan (our own three launches:
// sums of 4096-entry blocks, one CTA scanning the sums, blocks scanned from their sums) -> csv_write_kernel (a CTA renders
// its 128 rows into shared memory at their offsets and copies the contiguous text out with 16-byte stores);
// the host then writes the file with ONE write().  Bound: HBM (all columns read twice) + PCIe + the file write.
// ------------------------------------------------------------------------------------------
namespace {

struct CsvCols {
    const uint8_t *col[NUM_COLS];
    uint32_t width[NUM_COLS];
};

__device__ __forceinline__ int dec_len_u64(unsigned long long v) {
    int n = 1;
    while (v >= 10ull) {
        v /= 10ull;
        ++n;
    }
    return n;
}

// length of the NUL-padded text cell (the cell always holds at least one NUL)
__device__ __forceinline__ int text_len(const uint8_t *cell, uint32_t w) {
    const uint4 *q = reinterpret_cast<const uint4 *>(cell);
    for (uint32_t k = 0; k < (w >> 4); ++k) {
        const uint4 v = __ldg(q + k);
        const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t x = words[i];
            const uint32_t z = (x - 0x01010101u) & ~x & 0x80808080u;  // a zero byte in x
            if (z) return static_cast<int>(k * 16 + i * 4 + ((__ffs(z) - 1) >> 3));
        }
    }
    return static_cast<int>(w);
}

__device__ __forceinline__ char *put_text(char *dst, const uint8_t *cell, uint32_t w) {
    const uint4 *q = reinterpret_cast<const uint4 *>(cell);
    for (uint32_t k = 0; k < (w >> 4); ++k) {
        const uint4 v = __ldg(q + k);
        const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const char c = static_cast<char>((words[i] >> (8 * b)) & 0xffu);
                if (c == 0) return dst;
                *dst++ = c;
            }
        }
    }
    return dst;
}

__device__ __forceinline__ char *put_i32(char *dst, int v) {
    unsigned int mag = static_cast<unsigned int>(v);
    if (v < 0) {
        *dst++ = '-';
        mag = 0u - mag;
    }
    return dst + render_u64(mag, dst);
}

__device__ __forceinline__ unsigned int row_len(const CsvCols &t, long long r) {
    unsigned int n = 12;  // 11 commas + '\n'
#pragma unroll
    for (int c = 0; c < NUM_COLS; ++c) {
        const uint8_t *cell = t.col[c] + static_cast<size_t>(r) * t.width[c];
        if (c == C_COMMAND_ID) {
            n += dec_len_u64(*reinterpret_cast<const unsigned long long *>(cell));
        } else if (c == C_EXIT_CODE || c == C_USER_ID || c == C_RISK_LEVEL) {
            const int v = *reinterpret_cast<const int *>(cell);
            n += (v < 0 ? 1 : 0) + dec_len_u64(v < 0 ? 0u - static_cast<unsigned int>(v) : static_cast<unsigned int>(v));
        } else if (c == C_SUDO_USED) {
            n += 1;
        } else {
            n += text_len(cell, t.width[c]);
        }
    }
    return n;
}

__device__ __forceinline__ void render_row(const CsvCols &t, long long r, char *dst) {
#pragma unroll
    for (int c = 0; c < NUM_COLS; ++c) {
        const uint8_t *cell = t.col[c] + static_cast<size_t>(r) * t.width[c];
        if (c == C_COMMAND_ID)
            dst += render_u64(*reinterpret_cast<const unsigned long long *>(cell), dst);
        else if (c == C_EXIT_CODE || c == C_USER_ID || c == C_RISK_LEVEL)
            dst = put_i32(dst, *reinterpret_cast<const int *>(cell));
        else if (c == C_SUDO_USED)
            *dst++ = cell[0] ? '1' : '0';
        else
            dst = put_text(dst, cell, t.width[c]);
        *dst++ = (c == NUM_COLS - 1) ? '\n' : ',';
    }
}

constexpr int kCsvRows = 128;           // rows (= threads) per CTA
constexpr int kCsvStage = 40 * 1024;    // shared-memory stage; a CTA whose text is longer writes straight to HBM

__global__ void __launch_bounds__(kCsvRows) csv_len_kernel(const CsvCols t, long long n, unsigned long long *lens) {
    const long long r = blockIdx.x * static_cast<long long>(kCsvRows) + threadIdx.x;
    if (r < n) lens[r] = row_len(t, r);
    if (r == n) lens[r] = 0;  // slot n: the scan leaves the total there
}

// rows [row_begin, n) of the table; out[0] is the byte at file offset base_off (= offs[row_begin])
__global__ void __launch_bounds__(kCsvRows) csv_write_kernel(const CsvCols t, long long row_begin, long long n,
                                                             const unsigned long long *__restrict__ offs,
                                                             unsigned long long base_off, char *out_chunk) {
    __shared__ __align__(16) char stage[kCsvStage];
    char *out = out_chunk - base_off;  // only dereferenced at offsets >= base_off
    const long long r0 = row_begin + blockIdx.x * static_cast<long long>(kCsvRows);
    const long long r1 = (r0 + kCsvRows < n) ? r0 + kCsvRows : n;
    const long long r = r0 + threadIdx.x;
    const unsigned long long begin = offs[r0], end = offs[r1];
    const unsigned long long bytes = end - begin;
    const unsigned int pad = static_cast<unsigned int>(reinterpret_cast<uintptr_t>(out + begin) & 15u);
    if (pad + bytes > kCsvStage) {  // unusually long rows: no staging
        if (r < r1) render_row(t, r, out + offs[r]);
        return;
    }
    // stage[pad + i] <-> out[begin + i]: shared and global addresses agree modulo 16
    if (r < r1) render_row(t, r, stage + pad + (offs[r] - begin));
    __syncthreads();
    char *gbase = out + begin - pad;  // 16-byte aligned
    const unsigned int lo = pad, hi = pad + static_cast<unsigned int>(bytes);
    const unsigned int s0 = (lo + 15u) & ~15u, e0 = hi & ~15u;
    if (s0 >= e0) {
        for (unsigned int i = lo + threadIdx.x; i < hi; i += kCsvRows) gbase[i] = stage[i];
        return;
    }
    for (unsigned int i = lo + threadIdx.x; i < s0; i += kCsvRows) gbase[i] = stage[i];
    for (unsigned int i = s0 / 16 + threadIdx.x; i < e0 / 16; i += kCsvRows)
        reinterpret_cast<uint4 *>(gbase)[i] = reinterpret_cast<const uint4 *>(stage)[i];
    for (unsigned int i = e0 + threadIdx.x; i < hi; i += kCsvRows) gbase[i] = stage[i];
}

}  // namespace

// ---- exclusive scan of 64-bit values, in place: block sums -> scan of the sums (one CTA) -> blocks from their sums ----
namespace {
constexpr int kScanThreads = 512, kScanItems = 8, kScanBlock = kScanThreads * kScanItems;

// exclusive scan of one 64-bit value per thread over the CTA; *total = the CTA's sum.  Every thread calls it.
__device__ __forceinline__ unsigned long long cta_exclusive_scan_u64(unsigned long long v, unsigned long long *s_warp,
                                                                     unsigned long long *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < n_warps ? s_warp[lane] : 0ull;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += y;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    const unsigned long long before = warp > 0 ? s_warp[warp - 1] : 0ull;
    *total = s_warp[n_warps - 1];
    __syncthreads();
    return before + x - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(const unsigned long long *__restrict__ data,
                                                                       long long n, unsigned long long *__restrict__ sums) {
    __shared__ unsigned long long s_warp[32];
    const long long base = blockIdx.x * static_cast<long long>(kScanBlock);
    unsigned long long v = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const long long idx = base + i * kScanThreads + threadIdx.x;
        if (idx < n) v += data[idx];
    }
    unsigned long long total = 0;
    cta_exclusive_scan_u64(v, s_warp, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned long long *__restrict__ sums, long long n_blocks) {
    __shared__ unsigned long long s_warp[32];
    unsigned long long running = 0;
    for (long long base = 0; base < n_blocks; base += 1024) {
        const long long i = base + threadIdx.x;
        const unsigned long long v = i < n_blocks ? sums[i] : 0ull;
        unsigned long long total = 0;
        const unsigned long long before = cta_exclusive_scan_u64(v, s_warp, &total);
        if (i < n_blocks) sums[i] = running + before;
        running += total;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(unsigned long long *__restrict__ data, long long n,
                                                                  const unsigned long long *__restrict__ sums) {
    __shared__ unsigned long long s_warp[32];
    // a thread owns kScanItems consecutive entries
    const long long first = blockIdx.x * static_cast<long long>(kScanBlock) + threadIdx.x * kScanItems;
    unsigned long long v[kScanItems];
    unsigned long long mine = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = first + i < n ? data[first + i] : 0ull;
        mine += v[i];
    }
    unsigned long long total = 0;
    unsigned long long run = sums[blockIdx.x] + cta_exclusive_scan_u64(mine, s_warp, &total);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (first + i < n) data[first + i] = run;
        run += v[i];
    }
}

cudaError_t exclusive_scan_u64(unsigned long long *d_data, long long n, unsigned long long *d_sums, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const long long n_blocks = (n + kScanBlock - 1) / kScanBlock;
    scan_block_sums_kernel<<<static_cast<unsigned int>(n_blocks), kScanThreads, 0, stream>>>(d_data, n, d_sums);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(d_sums, n_blocks);
    scan_apply_kernel<<<static_cast<unsigned int>(n_blocks), kScanThreads, 0, stream>>>(d_data, n, d_sums);
    return cudaGetLastError();
}
}  // namespace

// Phase 1: row lengths + scan.  d_offs holds n + 1 entries; after the call d_offs[i] = byte offset of row i and
// d_offs[n] = total bytes.  scan_tmp / scan_tmp_bytes: scratch of the scan (call with scan_tmp == nullptr to size it).
cudaError_t csv_measure(const DevTable &t, unsigned long long *d_offs, void *scan_tmp, size_t *scan_tmp_bytes,
                        cudaStream_t stream) {
    const size_t need = static_cast<size_t>((t.n + 1 + kScanBlock - 1) / kScanBlock) * sizeof(unsigned long long) + 256;
    if (!scan_tmp) {
        *scan_tmp_bytes = need;
        return cudaSuccess;
    }
    if (*scan_tmp_bytes < need) return cudaErrorInvalidValue;
    CsvCols cc;
    for (int c = 0; c < NUM_COLS; ++c) {
        cc.col[c] = t.col[c].d;
        cc.width[c] = t.col[c].width;
    }
    const long long blocks = (t.n + 1 + kCsvRows - 1) / kCsvRows;
    csv_len_kernel<<<static_cast<unsigned int>(blocks), kCsvRows, 0, stream>>>(cc, t.n, d_offs);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return exclusive_scan_u64(d_offs, t.n + 1, static_cast<unsigned long long *>(scan_tmp), stream);
}

// Phase 2: the text of rows [r0, r1) into d_out, whose first byte is file offset base_off = d_offs[r0]
cudaError_t csv_write(const DevTable &t, const unsigned long long *d_offs, long long r0, long long r1,
                      unsigned long long base_off, char *d_out, cudaStream_t stream) {
    if (r1 <= r0) return cudaSuccess;
    CsvCols cc;
    for (int c = 0; c < NUM_COLS; ++c) {
        cc.col[c] = t.col[c].d;
        cc.width[c] = t.col[c].width;
    }
    const long long blocks = (r1 - r0 + kCsvRows - 1) / kCsvRows;
    csv_write_kernel<<<static_cast<unsigned int>(blocks), kCsvRows, 0, stream>>>(cc, r0, r1, d_offs, base_off, d_out);
    return cudaGetLastError();
}

}  // namespace qpe
