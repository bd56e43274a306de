// ingest_gpu.cu -- K6: CSV text -> columnar device table, parsed ON THE DEVICE.
//
// "Next" row 1 of the scope table: ingest is ~2/3 of QPESeq's wall time
// (engine/serial/buildEngine-serial.c:70-221: fgets + 12 x malloc(1024) + calloc per row).
// Here the file is copied to HBM once and four small kernels do the rest:
//   count_newlines   one count per 4 KiB block of text
//   scan_counts      exclusive scan of the block counts (single CTA; <= a few 10^4 blocks)
//   mark_lines       start offset of every line, in file order
//   measure_rows     one thread per row: field rules of csv_rules.h, max unescaped length per text
//                    column (sizes the fixed-width cells) and the longest line
//   parse_rows       one thread per row: same rules, cells written straight into the columns
// Row = one fgets(line, 1024) chunk after the header chunk.  For lines shorter than 1023 characters
// a chunk IS a line; if any line is longer the reference splits it into several "rows" -- that
// pathological case is detected (longest line) and handed to the host loader, which restates the
// chunking exactly (ingest.cpp).  Blank lines become all-zero rows, as in the reference.

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "csv_rules.h"
#include "engine.cuh"

namespace qpe {

namespace {

constexpr int kBlockBytes = 4096;
constexpr int kLineThreads = 256;  // 16 bytes per thread

__global__ void __launch_bounds__(kLineThreads) count_newlines_kernel(const char *__restrict__ text, long long n,
                                                                      unsigned int *__restrict__ block_counts) {
    const long long base = static_cast<long long>(blockIdx.x) * kBlockBytes + threadIdx.x * 16;
    unsigned int c = 0;
    if (base + 16 <= n) {
        const uint4 v = *reinterpret_cast<const uint4 *>(text + base);
        const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) c += ((w[k] >> (8 * b)) & 0xffu) == '\n';
    } else {
        for (long long i = base; i < n && i < base + 16; ++i) c += text[i] == '\n';
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ unsigned int s[kLineThreads / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int i = 0; i < kLineThreads / 32; ++i) t += s[i];
        block_counts[blockIdx.x] = t;
    }
}

// exclusive scan of n counts in place (single CTA), total -> *total_out
__global__ void __launch_bounds__(1024) scan_counts_kernel(unsigned int *counts, long long n, unsigned long long *offsets,
                                                           unsigned long long *total_out) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        const unsigned long long v = i < n ? counts[i] : 0;
        unsigned long long inc = v;
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if ((threadIdx.x & 31) >= d) inc += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long w = s_warp[threadIdx.x];
            unsigned long long winc = w;
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, d);
                if (threadIdx.x >= d) winc += t;
            }
            s_warp[threadIdx.x] = winc - w;
        }
        __syncthreads();
        const unsigned long long excl = s_carry + s_warp[threadIdx.x >> 5] + inc - v;
        if (i < n) offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}

// line_start[k] = offset of the first byte of line k (line 0 starts at 0)
__global__ void __launch_bounds__(kLineThreads) mark_lines_kernel(const char *__restrict__ text, long long n,
                                                                  const unsigned long long *__restrict__ block_off,
                                                                  unsigned long long *__restrict__ line_start) {
    const long long base = static_cast<long long>(blockIdx.x) * kBlockBytes + threadIdx.x * 16;
    unsigned int mask = 0;  // bit i: byte base+i is a newline
    for (int i = 0; i < 16; ++i)
        if (base + i < n && text[base + i] == '\n') mask |= 1u << i;
    const unsigned int c = __popc(mask);
    unsigned int inc = c;
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
        if ((threadIdx.x & 31) >= d) inc += t;
    }
    __shared__ unsigned int s[kLineThreads / 32];
    if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned int before = inc - c;
    for (unsigned int w = 0; w < (threadIdx.x >> 5); ++w) before += s[w];
    unsigned long long k = block_off[blockIdx.x] + before + 1;  // the newline ending line j starts line j+1
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        line_start[k++] = static_cast<unsigned long long>(base + b + 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
}

struct IngestParams {
    const char *text;
    long long n_bytes;
    const unsigned long long *line_start;  // n_lines + 1 entries (sentinel = n_bytes)
    long long n_rows;                      // rows = lines after the header
    unsigned int *max_len;                 // [NUM_COLS] max unescaped text length, [NUM_COLS] = longest line
    uint8_t *col[NUM_COLS];
    uint32_t width[NUM_COLS];
};

__device__ __forceinline__ void row_span(const IngestParams &p, long long row, const char **b, const char **e) {
    const unsigned long long s = p.line_start[row + 1];  // line 0 is the header
    const unsigned long long t = p.line_start[row + 2];
    *b = p.text + s;
    *e = p.text + t;
}

__global__ void __launch_bounds__(256) measure_rows_kernel(const __grid_constant__ IngestParams p) {
    unsigned int mx[NUM_COLS];
    for (int c = 0; c < NUM_COLS; ++c) mx[c] = 0;
    unsigned int longest = 0;
    for (long long row = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; row < p.n_rows;
         row += static_cast<long long>(gridDim.x) * blockDim.x) {
        const char *cur, *end;
        row_span(p, row, &cur, &end);
        const unsigned int ll = static_cast<unsigned int>(end - cur);
        longest = ll > longest ? ll : longest;
        for (int c = 0; c < NUM_COLS; ++c) {
            int len = 0;
            if (!csv::next_field(cur, end, nullptr, 0, &len)) break;  // absent: this and all later fields stay zero
            if (csv::col_type(c) == T_STR) {
                const unsigned int cap = csv::col_field_bytes(c) - 1;
                const unsigned int l = static_cast<unsigned int>(len) > cap ? cap : static_cast<unsigned int>(len);
                mx[c] = l > mx[c] ? l : mx[c];
            }
        }
    }
    for (int c = 0; c < NUM_COLS; ++c) {
        const unsigned int w = __reduce_max_sync(0xffffffffu, mx[c]);
        if ((threadIdx.x & 31) == 0 && w) atomicMax(&p.max_len[c], w);
    }
    longest = __reduce_max_sync(0xffffffffu, longest);
    if ((threadIdx.x & 31) == 0) atomicMax(&p.max_len[NUM_COLS], longest);
}

__global__ void __launch_bounds__(256) parse_rows_kernel(const __grid_constant__ IngestParams p) {
    for (long long row = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; row < p.n_rows;
         row += static_cast<long long>(gridDim.x) * blockDim.x) {
        const char *cur, *end;
        row_span(p, row, &cur, &end);
        char num[64];
        for (int c = 0; c < NUM_COLS; ++c) {
            int len = 0;
            if (csv::col_type(c) == T_STR) {
                // text goes straight into its (pre-zeroed) cell; at most field_bytes - 1 characters are kept
                char *cell = reinterpret_cast<char *>(p.col[c]) + static_cast<size_t>(row) * p.width[c];
                const int cap = static_cast<int>(csv::col_field_bytes(c)) - 1;
                const int room = static_cast<int>(p.width[c]) - 1;
                if (!csv::next_field(cur, end, cell, cap < room ? cap : room, &len)) break;
            } else {
                if (!csv::next_field(cur, end, num, 63, &len)) break;
                const int n = len < 63 ? len : 63;
                if (csv::col_type(c) == T_U64)
                    reinterpret_cast<unsigned long long *>(p.col[c])[row] = csv::parse_u64(num, n);
                else if (csv::col_type(c) == T_I32)
                    reinterpret_cast<int *>(p.col[c])[row] = csv::parse_i32(num, n);
                else
                    p.col[c][row] = csv::parse_bool(num, len) ? 1 : 0;
            }
        }
    }
}

struct Mapped {
    const char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = static_cast<size_t>(st.st_size);
        if (n == 0) {
            p = "";
            return true;
        }
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (m == MAP_FAILED) return false;
        p = static_cast<const char *>(m);
        return true;
    }
    ~Mapped() {
        if (p && n) munmap(const_cast<char *>(p), n);
        if (fd >= 0) ::close(fd);
    }
};

}  // namespace

// 1 = table built on the device; 0 = the caller must use the host loader (file unreadable, a line of
// >= 1023 characters, embedded NUL bytes before a newline are fine); -1 = CUDA error (message set)
int ingest_csv_gpu(GpuEngine *g, const char *path, int *launches_out) {
    cudaSetDevice(g->device);
    Mapped f;
    if (!f.open(path)) return 0;
    const long long n = static_cast<long long>(f.n);
    if (n == 0) return 0;
    char *d_text = nullptr;
    unsigned int *d_counts = nullptr, *d_max = nullptr;
    unsigned long long *d_off = nullptr, *d_total = nullptr, *d_lines = nullptr;
    const long long n_blocks = (n + kBlockBytes - 1) / kBlockBytes;
    int launches = 0;
    int rc = -1;
    unsigned long long n_newlines = 0;
    unsigned int h_max[NUM_COLS + 1];
    IngestParams p{};
    long long n_lines = 0, n_rows = 0;
    cudaStream_t st = g->stream;
    auto fail = [&]() {
        if (d_text) cudaFree(d_text);
        if (d_counts) cudaFree(d_counts);
        if (d_max) cudaFree(d_max);
        if (d_off) cudaFree(d_off);
        if (d_total) cudaFree(d_total);
        if (d_lines) cudaFree(d_lines);
        return rc;
    };
    if (!cuda_ok(cudaMalloc(&d_text, static_cast<size_t>(n) + 64), "cudaMalloc csv text") ||
        !cuda_ok(cudaMalloc(&d_counts, static_cast<size_t>(n_blocks) * 4), "cudaMalloc") ||
        !cuda_ok(cudaMalloc(&d_off, static_cast<size_t>(n_blocks) * 8), "cudaMalloc") ||
        !cuda_ok(cudaMalloc(&d_total, 8), "cudaMalloc") ||
        !cuda_ok(cudaMalloc(&d_max, sizeof(unsigned int) * (NUM_COLS + 1)), "cudaMalloc") ||
        !cuda_ok(cudaMemcpyAsync(d_text, f.p, static_cast<size_t>(n), cudaMemcpyHostToDevice, st), "upload csv text") ||
        !cuda_ok(cudaMemsetAsync(d_text + n, 0, 64, st), "memset") ||
        !cuda_ok(cudaMemsetAsync(d_max, 0, sizeof(unsigned int) * (NUM_COLS + 1), st), "memset"))
        return fail();
    count_newlines_kernel<<<static_cast<unsigned>(n_blocks), kLineThreads, 0, st>>>(d_text, n, d_counts);
    scan_counts_kernel<<<1, 1024, 0, st>>>(d_counts, n_blocks, d_off, d_total);
    launches += 2;
    if (!cuda_ok(cudaGetLastError(), "line count kernels") ||
        !cuda_ok(cudaMemcpyAsync(&n_newlines, d_total, 8, cudaMemcpyDeviceToHost, st), "download line count") ||
        !cuda_ok(cudaStreamSynchronize(st), "line count sync"))
        return fail();
    // lines = newline-terminated lines (+ one unterminated tail if the file does not end with '\n')
    n_lines = static_cast<long long>(n_newlines) + (f.p[n - 1] != '\n' ? 1 : 0);
    n_rows = n_lines > 0 ? n_lines - 1 : 0;  // first chunk = header (buildEngine-serial.c:86-89)
    if (!cuda_ok(cudaMalloc(&d_lines, static_cast<size_t>(n_lines + 2) * 8), "cudaMalloc line index")) return fail();
    mark_lines_kernel<<<static_cast<unsigned>(n_blocks), kLineThreads, 0, st>>>(d_text, n, d_off, d_lines);
    ++launches;
    {
        // sentinel(s): the end of the last line is the end of the file
        const unsigned long long endv = static_cast<unsigned long long>(n);
        const unsigned long long tail[2] = {endv, endv};
        // when the file ends with '\n' mark_lines already wrote line_start[n_lines] = n
        if (!cuda_ok(cudaMemcpyAsync(d_lines + n_lines, tail, (f.p[n - 1] != '\n') ? 16 : 8, cudaMemcpyHostToDevice, st),
                     "line sentinel"))
            return fail();
    }
    p.text = d_text;
    p.n_bytes = n;
    p.line_start = d_lines;
    p.n_rows = n_rows;
    p.max_len = d_max;
    {
        const int grid = n_rows > 0 ? static_cast<int>(std::min<long long>((n_rows + 255) / 256, 148 * 16)) : 1;
        measure_rows_kernel<<<grid, 256, 0, st>>>(p);
        ++launches;
        if (!cuda_ok(cudaGetLastError(), "measure kernel") ||
            !cuda_ok(cudaMemcpyAsync(h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, st), "download widths") ||
            !cuda_ok(cudaStreamSynchronize(st), "measure sync"))
            return fail();
    }
    if (h_max[NUM_COLS] >= 1023u) {  // a line the reference's fgets(1024) would split: exact host path instead
        rc = 0;
        return fail();
    }
    // allocate the columns at their data-driven widths and parse
    for (int c = 0; c < NUM_COLS; ++c) {
        if (g->table.col[c].d) cudaFree(g->table.col[c].d);
        g->table.col[c] = DevColumn();
    }
    {
        int64_t cap = n_rows + n_rows / 16 + 1;
        cap = (cap + kRowPad - 1) / kRowPad * kRowPad + kRowPad;
        for (int c = 0; c < NUM_COLS; ++c) {
            uint32_t w;
            switch (kCols[c].type) {
                case T_U64: w = 8; break;
                case T_I32: w = 4; break;
                case T_BOOL: w = 1; break;
                default: w = round_up16(h_max[c] + 1); break;
            }
            if (!column_alloc(&g->table.col[c], w, cap, st)) return fail();
            p.col[c] = g->table.col[c].d;
            p.width[c] = w;
        }
    }
    if (n_rows > 0) {
        const int grid = static_cast<int>(std::min<long long>((n_rows + 255) / 256, 148 * 16));
        parse_rows_kernel<<<grid, 256, 0, st>>>(p);
        ++launches;
        if (!cuda_ok(cudaGetLastError(), "parse kernel")) return fail();
    }
    if (!cuda_ok(cudaStreamSynchronize(st), "parse sync")) return fail();
    g->table.n = n_rows;
    g->table.row_base = 0;
    g->head.num_records = static_cast<int>(n_rows);
    for (auto &ix : g->idx) ix.dirty = true;
    if (launches_out) *launches_out = launches;
    rc = 1;
    return fail();  // releases the scratch buffers only; the columns stay
}

}  // namespace qpe
