// capi.cu -- the C-ABI of libqpegpu.so: include/executeEngine-gpu.h (drop-in surface, mirrors
// include/executeEngine-serial.h:69-151 of the reference), include/buildEngine-gpu.h and the
// low-level row-id interface of include/qpe_gpu.h.

#include <fcntl.h>
#include <unistd.h>

#include <climits>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "engine.cuh"
#include "executeEngine-gpu.h"
#include "buildEngine-gpu.h"
#include "qpe_gpu.h"

using namespace qpe;

namespace {


// Results produced by this library keep everything -- data[] (row pointers), the cell pointers and the cell
// text -- in ONE block; freeResultSet recognises them.  Large blocks are PINNED host memory (the projection's
// device->host copies then run at PCIe speed, asynchronously) and go back to a small pool instead of being
// released: a fresh 450 MB allocation costs ~150 ms of page faults / page locking, which is what a
// SELECT * of 1 M rows spent most of its time on.  The pool holds at most kPoolBlocks blocks / kPoolBytes.
struct ResultArena {
    char *block = nullptr;
    size_t bytes = 0;
    bool pinned = false;
    char *text = nullptr;    // inside block
    char **rows = nullptr;   // inside block
};
std::mutex g_result_mutex;
std::unordered_map<const struct resultSetS *, ResultArena> g_results;
std::vector<ResultArena> g_pool;
size_t g_pool_bytes = 0;
constexpr size_t kPoolBlocks = 6;
constexpr size_t kPoolBytes = size_t(3) << 30;
constexpr size_t kPinThreshold = size_t(1) << 20;

ResultArena arena_acquire(size_t need) {
    {
        std::lock_guard<std::mutex> lk(g_result_mutex);
        int best = -1;
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i].bytes >= need && (best < 0 || g_pool[i].bytes < g_pool[best].bytes)) best = static_cast<int>(i);
        if (best >= 0 && g_pool[best].bytes <= 8 * need + (size_t(64) << 20)) {
            ResultArena a = g_pool[best];
            g_pool.erase(g_pool.begin() + best);
            g_pool_bytes -= a.bytes;
            return a;
        }
    }
    ResultArena a;
    if (need >= kPinThreshold) {
        const size_t want = (need + need / 4 + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);  // head-room for reuse
        void *p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) == cudaSuccess) {
            a.block = static_cast<char *>(p);
            a.bytes = want;
            a.pinned = true;
            return a;
        }
        cudaGetLastError();  // no pinned memory left: plain memory works too, only slower
    }
    a.block = static_cast<char *>(std::malloc(need ? need : 1));
    a.bytes = need;
    a.pinned = false;
    return a;
}

void arena_pool_clear() {
    std::vector<ResultArena> blocks;
    {
        std::lock_guard<std::mutex> lk(g_result_mutex);
        blocks.swap(g_pool);
        g_pool_bytes = 0;
    }
    for (ResultArena &a : blocks)
        if (a.pinned)
            cudaFreeHost(a.block);
        else
            std::free(a.block);
}

void arena_release(ResultArena a) {
    if (!a.block) return;
    if (a.pinned) {
        std::lock_guard<std::mutex> lk(g_result_mutex);
        if (g_pool.size() < kPoolBlocks && g_pool_bytes + a.bytes <= kPoolBytes) {
            g_pool.push_back(a);
            g_pool_bytes += a.bytes;
            return;
        }
    }
    if (a.pinned)
        cudaFreeHost(a.block);
    else
        std::free(a.block);
}

const char *const kAllColumns[NUM_COLS] = {"command_id", "raw_command", "base_command", "shell_type",
                                           "exit_code",  "timestamp",   "sudo_used",    "working_directory",
                                           "user_id",    "user_name",   "host_name",    "risk_level"};

struct resultSetS *new_result() {
    struct resultSetS *r = static_cast<struct resultSetS *>(std::malloc(sizeof(struct resultSetS)));
    if (!r) {
        std::perror("Failed to allocate memory for result set");
        std::exit(EXIT_FAILURE);
    }
    r->numRecords = 0;
    r->numColumns = 0;
    r->columnNames = nullptr;
    r->columnTypes = nullptr;
    r->data = nullptr;
    r->queryTime = 0.0;
    r->success = false;
    return r;
}

void fill_stats(GpuEngine *g, qpe_scan_stats *s) {
    if (!s) return;
    engine_resolve_timing(g);
    const ScanStats &a = g->last;
    s->kernel_ms = a.kernel_ms;
    s->total_ms = a.total_ms;
    s->rows_scanned = a.rows_scanned;
    s->candidates = a.candidates;
    s->matches = a.matches;
    s->algo_bytes = a.algo_bytes;
    s->path = a.path;
    s->launches = a.launches;
    s->tile_rows = a.tile_rows;
    s->stages = a.stages;
    s->grid = a.grid;
    s->reserved = 0;
    s->scan_ms = a.scan_ms;
    s->compact_ms = a.compact_ms;
}

// SELECT projection: get_attribute_string_value (executeEngine-serial.c:216-248) for every
// (match, selected column), from gathered device columns.
bool materialise(GpuEngine *g, const char **selectItems, int numItems, int64_t m, struct resultSetS *res) {
    if (selectItems == nullptr || numItems == 0) {  // SELECT * (:490-493)
        selectItems = const_cast<const char **>(kAllColumns);
        numItems = NUM_COLS;
    }
    res->numRecords = static_cast<int>(m);
    res->numColumns = numItems;
    res->columnNames = static_cast<char **>(std::malloc(sizeof(char *) * (numItems > 0 ? numItems : 1)));
    for (int j = 0; j < numItems; ++j) res->columnNames[j] = strdup(selectItems[j]);

    // K7: every distinct known column is rendered ON THE DEVICE into fixed-width NUL-terminated slots
    // (format.cu) and lands in its own region of one arena; a cell pointer is then pure arithmetic.
    std::vector<int> colid(numItems);
    uint32_t slot[NUM_COLS] = {0};
    size_t region[NUM_COLS] = {0};
    bool have[NUM_COLS] = {false};
    size_t text_bytes = 16;  // [0..5) = "NULL" for unknown columns (:245-247)
    size_t row_slots = 0;
    for (int j = 0; j < numItems; ++j) {
        const int c = col_by_name(selectItems[j]);
        colid[j] = c;
        if (c >= 0 && !have[c]) {
            if (m > 0 && !g->table.col[c].d) {
                set_error(std::string("column '") + kCols[c].name + "' is not resident on the device");
                res->numRecords = 0;
                return false;
            }
            have[c] = true;
            slot[c] = format_slot_width(kCols[c].type, g->table.col[c].width);
            region[c] = text_bytes;
            text_bytes += static_cast<size_t>(m) * slot[c];
            row_slots += slot[c];
        }
    }
    // one block: data[m] | cell pointers [m x numItems] | text.  data is never NULL, also for zero matches
    // (printTable then prints an empty framed table, printHelper.c:38-41)
    const size_t data_bytes = (sizeof(char **) * static_cast<size_t>(m) + 63) & ~size_t(63);
    const size_t rows_bytes = (sizeof(char *) * (static_cast<size_t>(m) * numItems + 1) + 63) & ~size_t(63);
    cudaSetDevice(g->device);
    ResultArena arena = arena_acquire(data_bytes + rows_bytes + text_bytes + 64);
    if (!arena.block) {
        std::fprintf(stderr, "Memory allocation failed\n");
        std::exit(EXIT_FAILURE);
    }
    res->data = reinterpret_cast<char ***>(arena.block);
    arena.rows = reinterpret_cast<char **>(arena.block + data_bytes + (data_bytes ? 0 : 64));
    arena.text = arena.block + (data_bytes ? data_bytes : 64) + rows_bytes;
    std::memcpy(arena.text, "NULL", 5);

    // device side, in row chunks that keep the scratch under 1 GiB: render, then copy each column's slots
    // to its arena region; the copies of a chunk run while the host lays out the row pointers
    bool ok = true;
    int64_t chunk_rows = m;
    if (row_slots > 0 && m > 0) {
        const size_t kScratch = size_t(1) << 30;
        if (static_cast<size_t>(m) * row_slots > kScratch) chunk_rows = static_cast<int64_t>(kScratch / row_slots);
        size_t need = 0;  // every column's slots start on a 256-byte boundary of the scratch
        for (int c = 0; c < NUM_COLS; ++c)
            if (have[c]) need += (static_cast<size_t>(chunk_rows) * slot[c] + 255) & ~size_t(255);
        if (need > g->fmt_cap) {
            if (g->d_fmt) cudaFree(g->d_fmt);
            g->d_fmt = nullptr;
            g->fmt_cap = 0;
            ok = cuda_ok(cudaMalloc(&g->d_fmt, need + 256), "cudaMalloc projection scratch");
            if (ok) g->fmt_cap = need;
        }
    }
    auto enqueue_chunk = [&](int64_t r0, int64_t r1) {
        size_t off = 0;
        for (int c = 0; c < NUM_COLS && ok; ++c) {
            if (!have[c]) continue;
            const size_t bytes = static_cast<size_t>(r1 - r0) * slot[c];
            ok = cuda_ok(format_launch(g->table.col[c].d, kCols[c].type, g->table.col[c].width, g->d_ids + r0, r1 - r0,
                                       g->d_fmt + off, g->stream),
                         "projection kernel launch") &&
                 cuda_ok(cudaMemcpyAsync(arena.text + region[c] + static_cast<size_t>(r0) * slot[c], g->d_fmt + off, bytes,
                                         cudaMemcpyDeviceToHost, g->stream),
                         "download projection");
            off += (bytes + 255) & ~size_t(255);
            ++g->last.launches;
        }
    };
    if (ok && row_slots > 0 && m > 0) enqueue_chunk(0, chunk_rows < m ? chunk_rows : m);

    // row pointers (the reference's char ***data): arithmetic only, split over a few threads when large
    auto lay_rows = [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i) {
            char **row = arena.rows + static_cast<size_t>(i) * numItems;
            res->data[i] = row;
            for (int j = 0; j < numItems; ++j) {
                const int c = colid[j];
                row[j] = c < 0 ? arena.text : arena.text + region[c] + static_cast<size_t>(i) * slot[c];
            }
        }
    };
    const int64_t cells = m * numItems;
    int n_thr = cells > (int64_t(1) << 20) ? 8 : 1;
    const unsigned hw = std::thread::hardware_concurrency();
    if (hw && static_cast<unsigned>(n_thr) > hw) n_thr = static_cast<int>(hw);
    if (n_thr <= 1) {
        lay_rows(0, m);
    } else {
        std::vector<std::thread> pool;
        const int64_t per = (m + n_thr - 1) / n_thr;
        for (int t = 0; t < n_thr; ++t) {
            const int64_t a = t * per, b = (a + per < m) ? a + per : m;
            if (a < b) pool.emplace_back(lay_rows, a, b);
        }
        for (auto &th : pool) th.join();
    }
    ok = cuda_ok(cudaStreamSynchronize(g->stream), "projection sync") && ok;
    for (int64_t r0 = chunk_rows; ok && r0 < m; r0 += chunk_rows) {
        enqueue_chunk(r0, r0 + chunk_rows < m ? r0 + chunk_rows : m);
        ok = cuda_ok(cudaStreamSynchronize(g->stream), "projection sync") && ok;
    }
    if (!ok) {
        arena_release(arena);
        res->data = nullptr;
        res->numRecords = 0;
        return false;
    }
    res->columnTypes = static_cast<FieldType *>(std::calloc(numItems > 0 ? numItems : 1, sizeof(FieldType)));  // :524-525
    {
        std::lock_guard<std::mutex> lk(g_result_mutex);
        g_results[res] = arena;
    }
    return true;
}

bool engine_load_indexes(GpuEngine *g, int num_indexes, const char *indexed_attributes[], const int attribute_types[]) {
    for (int i = 0; i < num_indexes; ++i) {
        if (!engine_add_index(g, indexed_attributes[i], attribute_types ? attribute_types[i] : -1))
            std::fprintf(stderr, "Failed to create index for attribute: %s\n",
                         indexed_attributes[i] ? indexed_attributes[i] : "(null)");
    }
    return true;
}

// K8: the data file rewritten from the device columns (executeEngine-serial.c:683-706): the text is rendered on
// the GPU in chunks of <= 4 Mi rows, copied into a pooled pinned block and written with one write() per chunk.
// Returns false (error set) if anything fails; the caller then falls back to nothing -- there is no host path.
bool persist_csv_device(GpuEngine *g, const char *path) {
    cudaSetDevice(g->device);
    const DevTable &t = g->table;
    for (int c = 0; c < NUM_COLS; ++c)
        if (t.n > 0 && !t.col[c].d) {
            set_error(std::string("cannot rewrite the data file: column '") + kCols[c].name + "' is not resident");
            return false;
        }
    static const bool trace = std::getenv("QPE_TRACE_DML") != nullptr;
    const double t_open0 = now_ms();
    const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) {
        set_error(std::string("cannot open ") + path + " for writing");
        return false;
    }
    double t_open = now_ms() - t_open0, t_dev = 0, t_write = 0;
    if (t.n == 0) {
        ::close(fd);
        return true;
    }
    unsigned long long *d_offs = nullptr;
    void *d_tmp = nullptr;
    char *d_text = nullptr;
    size_t tmp_bytes = 0, text_cap = 0;
    ResultArena pin;
    bool ok = cuda_ok(cudaMalloc(&d_offs, sizeof(unsigned long long) * (t.n + 1)), "cudaMalloc csv offsets") &&
              cuda_ok(csv_measure(t, d_offs, nullptr, &tmp_bytes, g->stream), "csv scan sizing") &&
              cuda_ok(cudaMalloc(&d_tmp, tmp_bytes + 16), "cudaMalloc csv scan") &&
              cuda_ok(csv_measure(t, d_offs, d_tmp, &tmp_bytes, g->stream), "csv measure kernels");
    const int64_t kChunkRows = int64_t(4) << 20;
    for (int64_t r0 = 0; ok && r0 < t.n; r0 += kChunkRows) {
        const double t_c0 = now_ms();
        const int64_t r1 = (r0 + kChunkRows < t.n) ? r0 + kChunkRows : t.n;
        unsigned long long edge[2] = {0, 0};
        ok = cuda_ok(cudaMemcpyAsync(&edge[0], d_offs + r0, 8, cudaMemcpyDeviceToHost, g->stream), "csv offsets") &&
             cuda_ok(cudaMemcpyAsync(&edge[1], d_offs + r1, 8, cudaMemcpyDeviceToHost, g->stream), "csv offsets") &&
             cuda_ok(cudaStreamSynchronize(g->stream), "csv offsets");
        if (!ok) break;
        const size_t bytes = static_cast<size_t>(edge[1] - edge[0]);
        if (bytes + 32 > text_cap) {
            if (d_text) cudaFree(d_text);
            d_text = nullptr;
            text_cap = 0;
            ok = cuda_ok(cudaMalloc(&d_text, bytes + bytes / 8 + 64), "cudaMalloc csv text");
            if (!ok) break;
            text_cap = bytes + bytes / 8 + 64;
        }
        if (pin.bytes < bytes) {
            arena_release(pin);
            pin = arena_acquire(bytes + bytes / 8);
            if (!pin.block) {
                set_error("out of host memory for the CSV text");
                ok = false;
                break;
            }
        }
        g->last.launches += 1;
        ok = cuda_ok(csv_write(t, d_offs, r0, r1, edge[0], d_text, g->stream), "csv write kernel") &&
             cuda_ok(cudaMemcpyAsync(pin.block, d_text, bytes, cudaMemcpyDeviceToHost, g->stream), "download csv text") &&
             cuda_ok(cudaStreamSynchronize(g->stream), "csv text sync");
        const double t_c1 = now_ms();
        t_dev += t_c1 - t_c0;
        for (size_t done = 0; ok && done < bytes;) {
            const ssize_t w = ::write(fd, pin.block + done, bytes - done);
            if (w <= 0) {
                set_error(std::string("write failed on ") + path);
                ok = false;
                break;
            }
            done += static_cast<size_t>(w);
        }
        t_write += now_ms() - t_c1;
    }
    const double t_cl0 = now_ms();
    ::close(fd);
    if (trace)
        std::fprintf(stderr, "libqpegpu: csv rewrite of %lld rows: open+truncate %.2f ms, device render + copy %.2f ms, "
                             "write() %.2f ms, close %.2f ms\n",
                     static_cast<long long>(t.n), t_open, t_dev, t_write, now_ms() - t_cl0);
    arena_release(pin);
    if (d_text) cudaFree(d_text);
    if (d_tmp) cudaFree(d_tmp);
    if (d_offs) cudaFree(d_offs);
    return ok;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------
// include/qpe_gpu.h : utilities
// ------------------------------------------------------------------------------------------
int qpe_gpu_available(void) { return device_count() > 0 ? 1 : 0; }
const char *qpe_gpu_last_error(void) { return last_error_cstr(); }
void qpe_gpu_free(void *p) { std::free(p); }

// ------------------------------------------------------------------------------------------
// include/executeEngine-gpu.h
// ------------------------------------------------------------------------------------------
struct engineS *initializeEngineGPU(int num_indexes, const char *indexed_attributes[], const int attribute_types[],
                                    const char *datafile, const char *tableName) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!datafile) datafile = "../data/commands_50k.csv";  // the reference's default (executeEngine-serial.c:757)
    GpuEngine *g = engine_create(tableName ? tableName : "", datafile, num_indexes);
    if (!g) return nullptr;
    // ingest: parsed on the device (K6); the host loader handles what K6 declines (unreadable or
    // empty file, a line the reference's fgets(1024) would split) and QPE_INGEST=host forces it
    const char *mode = std::getenv("QPE_INGEST");
    int r = (mode && std::strcmp(mode, "host") == 0) ? 0 : ingest_csv_gpu(g, datafile, nullptr);
    if (r < 0) {
        engine_destroy(g);
        return nullptr;
    }
    if (r == 0) {
        HostColumns hc;
        if (load_csv_columns(datafile, &hc) < 0) hc.init_widths_minimal();  // unreadable file: empty table
        if (!engine_upload(g, hc)) {
            engine_destroy(g);
            return nullptr;
        }
    }
    engine_load_indexes(g, num_indexes, indexed_attributes, attribute_types);
    return &g->head;
}

struct engineS *qpe_gpu_engine_from_records(const record *rows, long long n_rows, int num_indexes,
                                            const char *indexed_attributes[], const int attribute_types[],
                                            const char *datafile, const char *tableName) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    GpuEngine *g = engine_create(tableName ? tableName : "", datafile, num_indexes);
    if (!g) return nullptr;
    HostColumns hc;
    hc.init_widths_minimal();
    hc.reserve_rows(n_rows);
    for (long long i = 0; i < n_rows; ++i) hc.append_record(rows[i]);
    if (!engine_upload(g, hc)) {
        engine_destroy(g);
        return nullptr;
    }
    engine_load_indexes(g, num_indexes, indexed_attributes, attribute_types);
    return &g->head;
}

void destroyEngineGPU(struct engineS *engine) {
    std::lock_guard<std::mutex> lk(g_api_mutex);  // (the engine's own mutex dies with it; no call may be running on it)
    if (engine == nullptr) {
        std::fprintf(stderr, "Attempted to destroy a NULL engine pointer\n");  // executeEngine-serial.c:811
        return;
    }
    GpuEngine *g = as_engine(engine);
    if (!g) {
        std::fprintf(stderr, "libqpegpu: %s\n", qpe_gpu_last_error());
        return;
    }
    engine_destroy(g);
    if (engine_live_count() == 0) arena_pool_clear();  // the last engine is gone: give the pinned result blocks back
}

struct resultSetS *executeQuerySelectGPU(struct engineS *engine, const char **selectItems, int numSelectItems,
                                         const char *tableName, struct whereClauseS *whereClause) {
    (void)tableName;  // ignored by every reference engine
    EngineLock lk(engine);
    struct resultSetS *res = new_result();
    GpuEngine *g = as_engine(engine);
    if (!g) {
        std::fprintf(stderr, "libqpegpu: executeQuerySelectGPU: %s\n", qpe_gpu_last_error());
        return res;  // success == false, data == NULL -> printTable prints "No data found."
    }
    uint64_t m = 0;
    const auto t0 = std::chrono::steady_clock::now();
    if (!engine_match(g, whereClause, false, false, false, false, &m)) {
        std::fprintf(stderr, "libqpegpu: executeQuerySelectGPU: %s\n", qpe_gpu_last_error());
        return res;
    }
    const double match_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (m > static_cast<uint64_t>(INT_MAX)) {
        set_error("result has more than INT_MAX rows");
        return res;
    }
    if (!materialise(g, selectItems, numSelectItems, static_cast<int64_t>(m), res)) {
        std::fprintf(stderr, "libqpegpu: executeQuerySelectGPU: %s\n", qpe_gpu_last_error());
        return res;
    }
    res->queryTime = match_s;  // the reference reports the match phase only (:355,:475-476)
    res->success = true;
    return res;
}

struct resultSetS *executeQueryDeleteGPU(struct engineS *engine, const char *tableName,
                                         struct whereClauseS *whereClause) {
    (void)tableName;
    EngineLock lk(engine);
    struct resultSetS *res = new_result();
    GpuEngine *g = as_engine(engine);
    if (!g) {
        std::fprintf(stderr, "libqpegpu: executeQueryDeleteGPU: %s\n", qpe_gpu_last_error());
        return res;
    }
    const auto t0 = std::chrono::steady_clock::now();
    int64_t deleted = 0;
    if (!engine_delete(g, whereClause, &deleted)) {
        std::fprintf(stderr, "libqpegpu: executeQueryDeleteGPU: %s\n", qpe_gpu_last_error());
        return res;
    }
    // the reference rewrites the whole CSV after every DELETE, matched rows or not (:683-706)
    if (g->head.datafile && !persist_csv_device(g, g->head.datafile))
        std::fprintf(stderr, "libqpegpu: executeQueryDeleteGPU: %s\n", qpe_gpu_last_error());  // the reference ignores fopen failures too (:684)
    res->numRecords = static_cast<int>(deleted);
    res->queryTime = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    res->success = true;
    return res;
}

bool executeQueryInsertGPU(struct engineS *engine, const char *tableName, const record *r) {
    (void)tableName;
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g || !r) return false;
    // validation of executeEngine-serial.c:544-551
    if (r->command_id == 0 || r->raw_command[0] == '\0' || r->base_command[0] == '\0' || r->shell_type[0] == '\0' ||
        r->timestamp[0] == '\0' || r->working_directory[0] == '\0' || r->user_name[0] == '\0' ||
        r->host_name[0] == '\0')
        return false;
    if (g->head.datafile) {
        FILE *f = std::fopen(g->head.datafile, "a");
        if (!f) return false;
        std::fprintf(f, "%llu,%s,%s,%s,%d,%s,%d,%s,%d,%s,%s,%d\n", r->command_id, r->raw_command, r->base_command,
                     r->shell_type, r->exit_code, r->timestamp, r->sudo_used, r->working_directory, r->user_id,
                     r->user_name, r->host_name, r->risk_level);
        std::fclose(f);
    }
    if (!engine_append(g, *r)) {
        std::fprintf(stderr, "libqpegpu: executeQueryInsertGPU: %s\n", qpe_gpu_last_error());
        return false;
    }
    return true;
}

bool addAttributeIndexGPU(struct engineS *engine, const char *tableName, const char *attributeName,
                          int attributeType) {
    (void)tableName;
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return true;
    // the reference returns FALSE when the index was created (inverted test, :833-840)
    return !engine_add_index(g, attributeName, attributeType);
}

bool makeIndexGPU(struct engineS *engine, const char *indexName, int attributeType) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return false;
    return engine_add_index(g, indexName, attributeType);
}

FieldType mapAttributeTypeGPU(int attributeType) {
    switch (attributeType) {
        case 0: return FIELD_UINT64;
        case 1: return FIELD_INT;
        case 2: return FIELD_STRING;
        case 3: return FIELD_BOOL;
        default: return static_cast<FieldType>(-1);
    }
}

int isAttributeIndexed(struct engineS *engine, const char *attributeName) {
    if (!engine || !attributeName) return -1;
    for (int i = 0; i < engine->num_indexes; ++i)
        if (std::strcmp(engine->indexed_attributes[i], attributeName) == 0) return i;
    return -1;
}

void freeResultSet(struct resultSetS *result) {
    if (!result) return;
    ResultArena arena;
    bool ours = false;
    {
        std::lock_guard<std::mutex> lk(g_result_mutex);
        auto it = g_results.find(result);
        if (it != g_results.end()) {
            arena = it->second;
            g_results.erase(it);
            ours = true;
        }
    }
    if (result->columnNames) {
        for (int i = 0; i < result->numColumns; ++i) std::free(result->columnNames[i]);
        std::free(result->columnNames);
    }
    std::free(result->columnTypes);
    if (ours) {
        arena_release(arena);  // data[], the cell pointers and the text are one block
    } else if (result->data) {
        // a result built by someone else, cell by cell: the reference's destructor (:881-908)
        for (int i = 0; i < result->numRecords; ++i) {
            if (result->data[i]) {
                for (int j = 0; j < result->numColumns; ++j) std::free(result->data[i][j]);
                std::free(result->data[i]);
            }
        }
        std::free(result->data);
    }
    std::free(result);
}

// ------------------------------------------------------------------------------------------
// include/qpe_gpu.h : row-id interface
// ------------------------------------------------------------------------------------------
unsigned long long qpe_gpu_row_base(const struct engineS *engine) {
    GpuEngine *g = as_engine(const_cast<struct engineS *>(engine));
    return g ? g->table.row_base : 0ull;
}

long long qpe_gpu_num_rows(const struct engineS *engine) {
    GpuEngine *g = as_engine(const_cast<struct engineS *>(engine));
    return g ? g->table.n : -1;
}

int qpe_gpu_select_ids_device(struct engineS *engine, struct whereClauseS *whereClause, int flags,
                              unsigned long long *count_out, const unsigned int **d_ids_out, qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    uint64_t m = 0;
    const bool count_only = (flags & QPE_SCAN_COUNT_ONLY) != 0;
    if (!engine_match(g, whereClause, (flags & QPE_SCAN_FORCE) != 0, false, count_only, false, &m)) return -2;
    if ((flags & QPE_SCAN_GLOBAL_IDS) && !count_only && m > 0 && g->table.row_base != 0) {
        // sharded table: local position -> position in the whole table (this shard starts at row_base)
        if (g->table.row_base + static_cast<uint64_t>(g->table.n) > 0xffffffffull) {
            set_error("global row ids do not fit 32 bits");
            return -5;
        }
        if (!cuda_ok(add_base_launch(g->d_ids, static_cast<int64_t>(m), static_cast<uint32_t>(g->table.row_base),
                                     g->stream),
                     "add_base kernel launch") ||
            !cuda_ok(cudaStreamSynchronize(g->stream), "add_base sync"))
            return -4;
        g->last.launches += 1;
    }
    if (count_out) *count_out = m;
    if (d_ids_out) *d_ids_out = count_only ? nullptr : g->d_ids;
    fill_stats(g, stats);
    return 0;
}

/* K9: the match phase of n SELECTs.  Result by result identical to n calls of qpe_gpu_select_ids (same path
 * rule, same order): queries the reference would send down the index path run one by one through it; all the
 * full-scan queries that reference the same set of columns are evaluated together, up to 8 programs per pass. */
int qpe_gpu_select_ids_batch(struct engineS *engine, struct whereClauseS *const *whereClauses, int n_queries,
                             unsigned int **ids_out, size_t *n_out, qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    if (n_queries < 0 || (n_queries > 0 && (!whereClauses || !ids_out || !n_out))) {
        set_error("qpe_gpu_select_ids_batch: bad arguments");
        return -5;
    }
    for (int q = 0; q < n_queries; ++q) {
        ids_out[q] = nullptr;
        n_out[q] = 0;
    }
    auto fail = [&](int rc) {
        for (int q = 0; q < n_queries; ++q) {
            std::free(ids_out[q]);
            ids_out[q] = nullptr;
        }
        return rc;
    };
    auto download = [&](int q, const uint32_t *d_src, uint64_t m) -> bool {
        ids_out[q] = static_cast<unsigned int *>(std::malloc(sizeof(unsigned int) * (m ? m : 1)));
        if (!ids_out[q]) {
            set_error("out of host memory");
            return false;
        }
        n_out[q] = static_cast<size_t>(m);
        return m == 0 || cuda_ok(cudaMemcpyAsync(ids_out[q], d_src, m * 4, cudaMemcpyDeviceToHost, g->stream), "download ids");
    };
    qpe_scan_stats total{};
    auto add_stats = [&]() {
        qpe_scan_stats s1{};
        fill_stats(g, &s1);
        total.kernel_ms += s1.kernel_ms;
        total.total_ms += s1.total_ms;
        total.rows_scanned += s1.rows_scanned;
        total.candidates += s1.candidates;
        total.matches += s1.matches;
        total.algo_bytes += s1.algo_bytes;
        total.launches += s1.launches;
        total.scan_ms += s1.scan_ms;
        total.compact_ms += s1.compact_ms;
        total.tile_rows = s1.tile_rows;
        total.stages = s1.stages;
        total.grid = s1.grid;
    };
    // Queries that reference the SAME set of columns share a pass: the staged tile is then exactly what each of
    // them would stage alone, so the batch costs one scan's HBM traffic and no tile geometry is given up.
    // (A union of different column sets shrinks the tile: measured slower than running the queries one by
    // one, profiles/r1_batch_probe.json.)  Everything else takes the single-query path (K1f / index).
    std::vector<std::pair<uint32_t, std::vector<int>>> groups;
    auto run_single = [&](int q) -> int {
        uint64_t m = 0;
        if (!engine_match(g, whereClauses[q], false, false, false, false, &m)) return -2;
        if (!download(q, g->d_ids, m) || !cuda_ok(cudaStreamSynchronize(g->stream), "download ids")) return -4;
        if (stats) add_stats();
        return 0;
    };
    for (int q = 0; q < n_queries; ++q) {
        if (engine_uses_index(g, whereClauses[q])) {
            const int rc = run_single(q);
            if (rc) return fail(rc);
            total.path = 1;
            continue;
        }
        uint32_t mask = 0;
        if (!engine_where_columns(g, whereClauses[q], &mask)) return fail(-2);
        size_t k = 0;
        while (k < groups.size() && groups[k].first != mask) ++k;
        if (k == groups.size()) groups.emplace_back(mask, std::vector<int>());
        groups[k].second.push_back(q);
    }
    for (const auto &grp : groups) {
        const std::vector<int> &qs = grp.second;
        if (qs.size() == 1 || grp.first == 0u) {  // alone, or no column referenced at all
            for (int q : qs) {
                const int rc = run_single(q);
                if (rc) return fail(rc);
            }
            continue;
        }
        for (size_t b = 0; b < qs.size(); b += kMaxBatch) {
            const int nb = static_cast<int>(qs.size() - b < static_cast<size_t>(kMaxBatch) ? qs.size() - b : kMaxBatch);
            const struct whereClauseS *wcs[kMaxBatch];
            for (int k = 0; k < nb; ++k) wcs[k] = whereClauses[qs[b + k]];
            uint64_t offs[kMaxBatch + 1];
            if (!engine_match_batch(g, wcs, nb, offs)) return fail(-2);
            for (int k = 0; k < nb; ++k)
                if (!download(qs[b + k], g->d_ids + offs[k], offs[k + 1] - offs[k])) return fail(-4);
            if (!cuda_ok(cudaStreamSynchronize(g->stream), "download ids")) return fail(-4);
            if (stats) add_stats();
        }
    }
    if (stats) *stats = total;
    return 0;
}

int qpe_gpu_select_ids(struct engineS *engine, struct whereClauseS *whereClause, unsigned int **ids_out, size_t *n_out,
                       qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    const double t0 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    uint64_t m = 0;
    if (!engine_match(g, whereClause, false, false, false, false, &m)) return -2;
    unsigned int *ids = static_cast<unsigned int *>(std::malloc(sizeof(unsigned int) * (m ? m : 1)));
    if (!ids) {
        set_error("out of host memory");
        return -3;
    }
    if (m && !cuda_ok(cudaMemcpy(ids, g->d_ids, sizeof(unsigned int) * m, cudaMemcpyDeviceToHost), "download ids")) {
        std::free(ids);
        return -4;
    }
    if (ids_out)
        *ids_out = ids;
    else
        std::free(ids);
    if (n_out) *n_out = static_cast<size_t>(m);
    fill_stats(g, stats);
    if (stats)
        stats->total_ms =
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count() - t0;
    return 0;
}

// Device-visible alias of a host pointer, or nullptr: pinned (cudaHostAlloc / cudaHostRegister)
// memory is mapped into the device's address space, so a kernel can store to it directly.
static unsigned int *device_alias_of_host(void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return static_cast<unsigned int *>(a.devicePointer);
}

// same as qpe_gpu_select_ids but into a caller-provided host buffer of `cap` ids (pinned memory makes
// the copies asynchronous).  On the full-scan path the fused kernel K1f reports every finished table
// segment through a progress word in mapped host memory and the segment's ids are copied out by the
// copy engine while the scan of the following segments is still running.
int qpe_gpu_select_ids_into(struct engineS *engine, struct whereClauseS *whereClause, int flags, unsigned int *ids,
                            size_t cap, size_t *n_out, qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    const double t0 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    uint64_t m = 0;
    cudaSetDevice(g->device);
    // K1 then K1c (pipeline setting 1) with a pinned destination: K1c stores straight into `ids`
    unsigned int *alias = (ids && cap && g->pipe_segments == 1) ? device_alias_of_host(ids) : nullptr;
    if (alias) {
        g->out_override = alias;
        g->out_override_cap = cap;
        g->id_base_override = 0;
    } else {
        g->host_out = ids;   // fused scan: finished table segments are copied out during the scan
        g->host_out_cap = cap;
    }
    const bool ok = engine_match(g, whereClause, (flags & QPE_SCAN_FORCE) != 0, false, false, false, &m);
    const bool direct = (alias && g->last.path == 0 && g->last.tile_rows > 0) || g->host_out_done;
    g->out_override = nullptr;
    g->out_override_cap = 0;
    g->host_out = nullptr;
    g->host_out_cap = 0;
    if (!ok) return -2;
    if (n_out) *n_out = static_cast<size_t>(m);
    if (m > cap) {
        set_error("id buffer too small");
        return -5;
    }
    if (!direct) {
        if (m && !cuda_ok(cudaMemcpyAsync(ids, g->d_ids, sizeof(unsigned int) * m, cudaMemcpyDeviceToHost, g->stream),
                          "download ids"))
            return -4;
        if (!cuda_ok(cudaStreamSynchronize(g->stream), "download ids")) return -4;
    }
    fill_stats(g, stats);
    if (stats)
        stats->total_ms =
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count() - t0;
    return 0;
}

int qpe_gpu_match_mask(struct engineS *engine, struct whereClauseS *whereClause, unsigned int *bitmap, size_t n_words,
                       unsigned long long *count_out, qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    const size_t need = static_cast<size_t>((g->table.n + 31) / 32);
    if (n_words < need) {
        set_error("bitmap too small");
        return -5;
    }
    uint64_t m = 0;
    if (!engine_match(g, whereClause, true, false, true, true, &m)) return -2;
    if (need && !cuda_ok(cudaMemcpy(bitmap, g->d_bitmap, need * 4, cudaMemcpyDeviceToHost), "download bitmap")) return -4;
    if (count_out) *count_out = m;
    fill_stats(g, stats);
    return 0;
}

int qpe_gpu_probe_batch(struct engineS *engine, const char *attribute, const KEY_T *lo, const KEY_T *hi,
                        size_t n_queries, unsigned int *first, unsigned int *count, qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    cudaSetDevice(g->device);
    const int slot = isAttributeIndexed(engine, attribute);
    if (slot < 0 || !g->idx[slot].usable) {
        set_error("attribute has no probe-able (u64 / int) index");
        return -6;
    }
    DevIndex &ix = g->idx[slot];
    int launches = 0;
    if (ix.dirty && !cuda_ok(index_build(&ix, g->table, g->stream, &launches), "index rebuild")) return -2;
    const size_t ksz = ix.type == T_U64 ? 8 : 4;
    std::vector<uint8_t> hlo(n_queries * ksz), hhi(n_queries * ksz);
    for (size_t q = 0; q < n_queries; ++q) {
        if (ix.type == T_U64) {
            const unsigned long long a = lo[q].v.u64, b = hi[q].v.u64;
            std::memcpy(&hlo[q * 8], &a, 8);
            std::memcpy(&hhi[q * 8], &b, 8);
        } else {
            const int a = lo[q].v.i32, b = hi[q].v.i32;
            std::memcpy(&hlo[q * 4], &a, 4);
            std::memcpy(&hhi[q * 4], &b, 4);
        }
    }
    void *dlo = nullptr, *dhi = nullptr;
    uint32_t *dfirst = nullptr, *dcount = nullptr;
    bool ok = cuda_ok(cudaMalloc(&dlo, n_queries * ksz + 16), "cudaMalloc probe") &&
              cuda_ok(cudaMalloc(&dhi, n_queries * ksz + 16), "cudaMalloc probe") &&
              cuda_ok(cudaMalloc(&dfirst, n_queries * 4 + 16), "cudaMalloc probe") &&
              cuda_ok(cudaMalloc(&dcount, n_queries * 4 + 16), "cudaMalloc probe");
    const auto t0 = std::chrono::steady_clock::now();
    ok = ok && cuda_ok(cudaMemcpyAsync(dlo, hlo.data(), n_queries * ksz, cudaMemcpyHostToDevice, g->stream), "probe h2d");
    ok = ok && cuda_ok(cudaMemcpyAsync(dhi, hhi.data(), n_queries * ksz, cudaMemcpyHostToDevice, g->stream), "probe h2d");
    engine_resolve_timing(g);  // ev0 / ev1 belong to the last match phase's slot: settle it before reuse
    if (ok) cudaEventRecord(g->ev0, g->stream);
    ok = ok && cuda_ok(index_probe(ix, dlo, dhi, static_cast<long long>(n_queries), dfirst, dcount, g->stream),
                       "probe kernel launch");
    if (ok) cudaEventRecord(g->ev1, g->stream);
    ok = ok && cuda_ok(cudaMemcpyAsync(first, dfirst, n_queries * 4, cudaMemcpyDeviceToHost, g->stream), "probe d2h");
    ok = ok && cuda_ok(cudaMemcpyAsync(count, dcount, n_queries * 4, cudaMemcpyDeviceToHost, g->stream), "probe d2h");
    ok = cuda_ok(cudaStreamSynchronize(g->stream), "probe sync") && ok;
    if (ok && stats) {
        std::memset(stats, 0, sizeof(*stats));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g->ev0, g->ev1);
        stats->kernel_ms = ms;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        stats->rows_scanned = static_cast<long long>(n_queries);
        stats->path = 1;
        stats->launches = 1 + launches;
        // per probe: two descents, each reading one node (fanout keys) per level incl. the leaf array
        stats->algo_bytes = static_cast<long long>(n_queries) * 2 * (ix.n_levels + 1) * ix.fanout * static_cast<long long>(ksz) +
                            static_cast<long long>(n_queries) * (2 * ksz + 8);
    }
    cudaFree(dlo);
    cudaFree(dhi);
    cudaFree(dfirst);
    cudaFree(dcount);
    return ok ? 0 : -4;
}

int qpe_gpu_index_slice(struct engineS *engine, const char *attribute, unsigned int first, unsigned int count,
                        unsigned int *row_ids_out) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    cudaSetDevice(g->device);
    const int slot = isAttributeIndexed(engine, attribute);
    if (slot < 0 || !g->idx[slot].usable) {
        set_error("attribute has no probe-able (u64 / int) index");
        return -6;
    }
    DevIndex &ix = g->idx[slot];
    int launches = 0;
    if (ix.dirty && !cuda_ok(index_build(&ix, g->table, g->stream, &launches), "index rebuild")) return -2;
    if (static_cast<long long>(first) + count > ix.n) {
        set_error("index slice out of range");
        return -5;
    }
    if (count && !cuda_ok(cudaMemcpy(row_ids_out, ix.perm + first, static_cast<size_t>(count) * 4, cudaMemcpyDeviceToHost),
                          "download slice"))
        return -4;
    return 0;
}

int qpe_gpu_index_slice_keys(struct engineS *engine, const char *attribute, unsigned int first, unsigned int count,
                             long long *keys_out) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    cudaSetDevice(g->device);
    const int slot = isAttributeIndexed(engine, attribute);
    if (slot < 0 || !g->idx[slot].usable) {
        set_error("attribute has no probe-able (u64 / int) index");
        return -6;
    }
    DevIndex &ix = g->idx[slot];
    int launches = 0;
    if (ix.dirty && !cuda_ok(index_build(&ix, g->table, g->stream, &launches), "index rebuild")) return -2;
    if (static_cast<long long>(first) + count > ix.n) {
        set_error("index slice out of range");
        return -5;
    }
    if (count == 0) return 0;
    if (ix.type == T_U64)
        return cuda_ok(cudaMemcpy(keys_out, static_cast<const unsigned long long *>(ix.keys) + first,
                                  static_cast<size_t>(count) * 8, cudaMemcpyDeviceToHost), "download keys") ? 0 : -4;
    std::vector<int> tmp(count);
    if (!cuda_ok(cudaMemcpy(tmp.data(), static_cast<const int *>(ix.keys) + first, static_cast<size_t>(count) * 4,
                            cudaMemcpyDeviceToHost), "download keys"))
        return -4;
    for (unsigned int i = 0; i < count; ++i) keys_out[i] = tmp[i];
    return 0;
}

int qpe_gpu_fetch_column(struct engineS *engine, const char *attribute, long long first_row, long long n_rows, void *out,
                         unsigned int *width_out) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    cudaSetDevice(g->device);
    const int c = col_by_name(attribute);
    if (c < 0) {
        set_error("unknown column");
        return -6;
    }
    const DevColumn &dc = g->table.col[c];
    if (width_out) *width_out = dc.width;
    if (!out) return 0;
    if (!dc.d) {
        set_error("column is not resident on the device");
        return -6;
    }
    if (first_row < 0 || n_rows < 0 || first_row + n_rows > g->table.n) {
        set_error("row range out of bounds");
        return -5;
    }
    if (n_rows && !cuda_ok(cudaMemcpy(out, dc.d + static_cast<size_t>(first_row) * dc.width,
                                      static_cast<size_t>(n_rows) * dc.width, cudaMemcpyDeviceToHost),
                           "download column"))
        return -4;
    return 0;
}

// ---- split scan: count first, compaction later to a caller-chosen (possibly peer) destination ----
int qpe_gpu_scan_count(struct engineS *engine, struct whereClauseS *whereClause, unsigned long long *count_out,
                       qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    uint64_t m = 0;
    if (!engine_match(g, whereClause, true, false, true, true, &m)) return -2;
    if (count_out) *count_out = m;
    fill_stats(g, stats);
    return 0;
}

// K1 + K1c in one call with the ids written to a caller-chosen destination (this GPU or a peer)
int qpe_gpu_select_ids_to(struct engineS *engine, struct whereClauseS *whereClause, unsigned int *dst_device,
                          unsigned long long dst_capacity, int global_ids, unsigned long long *count_out,
                          qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    uint32_t base = 0;
    if (global_ids) {
        if (g->table.row_base + static_cast<uint64_t>(g->table.n) > 0xffffffffull) {
            set_error("global row ids do not fit 32 bits");
            return -5;
        }
        base = static_cast<uint32_t>(g->table.row_base);
    }
    g->out_override = dst_device;
    g->out_override_cap = dst_capacity;
    g->id_base_override = base;
    uint64_t m = 0;
    const bool ok = engine_match(g, whereClause, true, false, false, false, &m);
    g->out_override = nullptr;
    g->out_override_cap = 0;
    g->id_base_override = 0;
    if (!ok) return -2;
    if (count_out) *count_out = m;
    fill_stats(g, stats);
    if (m > dst_capacity) {
        set_error("destination too small for the result (nothing was written past it)");
        return -5;
    }
    return 0;
}

int qpe_gpu_compact_to(struct engineS *engine, unsigned int *dst_device, int global_ids, qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    uint32_t base = 0;
    if (global_ids) {
        if (g->table.row_base + static_cast<uint64_t>(g->table.n) > 0xffffffffull) {
            set_error("global row ids do not fit 32 bits");
            return -5;
        }
        base = static_cast<uint32_t>(g->table.row_base);
    }
    if (!engine_compact_to(g, dst_device, base, ~0ull)) return -2;
    fill_stats(g, stats);
    return 0;
}

// ---- index path of a sharded table: per-segment (key, id) lists for the cross-shard merge ----
int qpe_gpu_select_segments(struct engineS *engine, struct whereClauseS *whereClause, int global_ids,
                            int *used_index_out, int *n_segments_out, size_t seg_counts_out[32], long long **keys_out,
                            unsigned int **ids_out) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    std::vector<SegmentResult> segs;
    bool used = false;
    if (!engine_match_segments(g, whereClause, &segs, &used)) return -2;
    if (used_index_out) *used_index_out = used ? 1 : 0;
    if (n_segments_out) *n_segments_out = static_cast<int>(segs.size());
    size_t total = 0;
    for (size_t s = 0; s < segs.size(); ++s) {
        if (seg_counts_out) seg_counts_out[s] = segs[s].ids.size();
        total += segs[s].ids.size();
    }
    long long *keys = static_cast<long long *>(std::malloc(sizeof(long long) * (total ? total : 1)));
    unsigned int *ids = static_cast<unsigned int *>(std::malloc(sizeof(unsigned int) * (total ? total : 1)));
    if (!keys || !ids) {
        std::free(keys);
        std::free(ids);
        set_error("out of host memory");
        return -3;
    }
    const uint32_t base = global_ids ? static_cast<uint32_t>(g->table.row_base) : 0u;
    size_t o = 0;
    for (const SegmentResult &r : segs) {
        // keys are handed out as ORDER keys: a u64 key has its top bit flipped, so that the caller's signed 64-bit
        // sort orders them as the B+ tree's unsigned compare does (recordSchema.c:88-127), keys >= 2^63 included
        const bool u64 = r.key_col >= 0 && kCols[r.key_col].type == T_U64;
        for (size_t k = 0; k < r.ids.size(); ++k, ++o) {
            keys[o] = u64 ? static_cast<long long>(static_cast<unsigned long long>(r.keys[k]) ^ 0x8000000000000000ull) : r.keys[k];
            ids[o] = r.ids[k] + base;
        }
    }
    if (keys_out) *keys_out = keys; else std::free(keys);
    if (ids_out) *ids_out = ids; else std::free(ids);
    return 0;
}

// ---- raw device buffers that can be shared with the other ranks of the box (CUDA IPC) ----
void *qpe_gpu_device_alloc(size_t bytes) {
    void *p = nullptr;
    if (!cuda_ok(cudaMalloc(&p, bytes ? bytes : 16), "cudaMalloc")) return nullptr;
    return p;
}
void qpe_gpu_device_free(void *p) {
    if (p) cudaFree(p);
}
int qpe_gpu_ipc_export(void *device_ptr, unsigned char handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    if (!cuda_ok(cudaIpcGetMemHandle(&h, device_ptr), "cudaIpcGetMemHandle")) return -4;
    std::memcpy(handle_out, &h, 64);
    return 0;
}
void *qpe_gpu_ipc_open(const unsigned char handle[64]) {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void *p = nullptr;
    if (!cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return nullptr;
    return p;
}
void qpe_gpu_ipc_close(void *mapped_ptr) {
    if (mapped_ptr) cudaIpcCloseMemHandle(mapped_ptr);
}
int qpe_gpu_copy_device(void *dst_device, const void *src_device, size_t bytes) {
    if (bytes == 0) return 0;
    return cuda_ok(cudaMemcpy(dst_device, src_device, bytes, cudaMemcpyDeviceToDevice), "device copy") ? 0 : -4;
}
int qpe_gpu_copy_to_device(void *dst_device, const void *src_host, size_t bytes) {
    if (bytes == 0) return 0;
    return cuda_ok(cudaMemcpy(dst_device, src_host, bytes, cudaMemcpyHostToDevice), "copy to device") ? 0 : -4;
}
int qpe_gpu_copy_to_host(void *dst_host, const void *src_device, size_t bytes) {
    if (bytes == 0) return 0;
    return cuda_ok(cudaMemcpy(dst_host, src_device, bytes, cudaMemcpyDeviceToHost), "copy to host") ? 0 : -4;
}

int qpe_gpu_copy_from_device(void *dst_host, const void *src_device, size_t bytes) {
    if (bytes == 0) return 0;
    return cuda_ok(cudaMemcpy(dst_host, src_device, bytes, cudaMemcpyDeviceToHost), "copy from device") ? 0 : -4;
}

int qpe_gpu_last_stats(struct engineS *engine, qpe_scan_stats *stats) {
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    fill_stats(g, stats);
    return 0;
}

int qpe_gpu_set_timing(struct engineS *engine, int accumulate) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    engine_resolve_all(g);
    g->accumulate_timing = accumulate != 0;
    g->acc_kernel_ms = g->acc_scan_ms = g->acc_compact_ms = g->acc_post_ms = 0;
    g->acc_calls = 0;
    return 0;
}

int qpe_gpu_timing_totals(struct engineS *engine, double totals_ms[4], long long *calls_out) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    engine_resolve_all(g);
    totals_ms[0] = g->acc_kernel_ms;
    totals_ms[1] = g->acc_scan_ms;
    totals_ms[2] = g->acc_compact_ms;
    totals_ms[3] = g->acc_post_ms;
    if (calls_out) *calls_out = g->acc_calls;
    return 0;
}

void qpe_gpu_trace_put(struct engineS *engine, int slot, double ms) {
    GpuEngine *g = as_engine(engine);
    if (g && slot >= 0 && slot < 8) g->trace[slot] = ms;
}

int qpe_gpu_last_trace(struct engineS *engine, double out[8]) {
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    engine_resolve_timing(g);
    for (int k = 0; k < 8; ++k) out[k] = g->trace[k];
    return 0;
}

int qpe_gpu_fused_trace(struct engineS *engine, unsigned long long *out, int max_ctas) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g || !out || !g->d_trace || g->last.grid <= 0) return 0;
    cudaSetDevice(g->device);
    const int n = g->last.grid < max_ctas ? g->last.grid : max_ctas;
    if (!cuda_ok(cudaMemcpy(out, g->d_trace, sizeof(unsigned long long) * 8 * n, cudaMemcpyDeviceToHost), "download trace"))
        return 0;
    return n;
}

int qpe_gpu_write_csv(struct engineS *engine, const char *path) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    HostColumns hc;
    if (!engine_download_all(g, &hc)) return -4;
    FILE *f = std::fopen(path, "w");
    if (!f) {
        set_error(std::string("cannot open ") + path);
        return -8;
    }
    std::vector<char> iobuf(1 << 20);
    std::setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
    std::fputs("command_id,raw_command,base_command,shell_type,exit_code,timestamp,sudo_used,working_directory,"
               "user_id,user_name,host_name,risk_level\r\n", f);
    auto put_text = [&](int c, int64_t i) {
        const uint8_t *cell = hc.data[c].data() + static_cast<size_t>(i) * hc.width[c];
        const size_t len = strnlen(reinterpret_cast<const char *>(cell), hc.width[c]);
        bool quote = false;
        for (size_t k = 0; k < len; ++k)
            if (cell[k] == ',' || cell[k] == '"' || cell[k] == '\n' || cell[k] == '\r') quote = true;
        if (!quote) {
            std::fwrite(cell, 1, len, f);
            return;
        }
        std::fputc('"', f);
        for (size_t k = 0; k < len; ++k) {
            if (cell[k] == '"') std::fputc('"', f);
            std::fputc(cell[k], f);
        }
        std::fputc('"', f);
    };
    for (int64_t i = 0; i < hc.n; ++i) {
        for (int c = 0; c < NUM_COLS; ++c) {
            if (c) std::fputc(',', f);
            const uint8_t *cell = hc.data[c].data() + static_cast<size_t>(i) * hc.width[c];
            switch (kCols[c].type) {
                case T_U64: {
                    unsigned long long v;
                    std::memcpy(&v, cell, 8);
                    std::fprintf(f, "%llu", v);
                    break;
                }
                case T_I32: {
                    int v;
                    std::memcpy(&v, cell, 4);
                    std::fprintf(f, "%d", v);
                    break;
                }
                case T_BOOL: std::fputs(cell[0] ? "true" : "false", f); break;
                default: put_text(c, i); break;
            }
        }
        std::fputs("\r\n", f);
    }
    std::fclose(f);
    return 0;
}

unsigned int qpe_gpu_query_upload_bytes(void) { return static_cast<unsigned int>(fused_param_bytes()); }

void *qpe_gpu_stream(struct engineS *engine) {
    GpuEngine *g = as_engine(engine);
    return g ? static_cast<void *>(g->stream) : nullptr;
}

int qpe_gpu_set_pipeline(struct engineS *engine, int segments) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    if (segments < 0 || segments > kMaxPipeSegments) {
        set_error("pipeline setting must be 0 (fused K1f), 1 (K1 then K1c) or 2..16 (pipelined table segments)");
        return -5;
    }
    g->pipe_segments = segments;
    return 0;
}

int qpe_gpu_set_tile(struct engineS *engine, int tile_rows, int stages) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    g->force_tile_rows = tile_rows;
    g->force_stages = stages;
    return 0;
}

}  // extern "C"
