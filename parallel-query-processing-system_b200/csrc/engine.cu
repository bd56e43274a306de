// engine.cu -- the engine object behind `struct engineS *`: device table management, the match
// phase (scan path / index path), projection download, DELETE / INSERT maintenance.
//
// Mirrors the control flow of the reference's serial engine, with the per-row loops replaced by
// kernel launches:
//   executeQuerySelectSerial  engine/serial/executeEngine-serial.c:328-528
//       :358-459 candidate generation per (top-level condition x index)  -> plan_segments + K3 probe
//       :464-474 linearSearchRecords on the table / on the candidates    -> K1 scan / K1g filter
//       :504-515 projection                                              -> K2 gather + host format
//   executeQueryDeleteSerial  :627-715   mask + stable compaction + CSV rewrite
//   executeQueryInsertSerial  :538-617   validation, CSV append, row append, index update
//
// There is NO CPU fallback anywhere in this file: without a usable CUDA device engine_create fails.

#include "engine.cuh"

#include <atomic>

#include <climits>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>

namespace qpe {

std::mutex g_api_mutex;
static std::atomic<int> g_live_engines{0};
int engine_live_count() { return g_live_engines.load(); }


// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }
const char *last_error_cstr() { return g_last_error.c_str(); }

bool cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    cudaGetLastError();  // reported here: a non-sticky error must not resurface at the next launch check
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    std::fprintf(stderr, "libqpegpu: %s: %s\n", what, cudaGetErrorString(e));
    return false;
}

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

GpuEngine *as_engine(struct engineS *e) {
    if (!e) {
        set_error("NULL engine");
        return nullptr;
    }
    GpuEngine *g = reinterpret_cast<GpuEngine *>(e);
    if (g->magic != kEngineMagic) {
        set_error("engine handle was not created by initializeEngineGPU");
        return nullptr;
    }
    return g;
}

int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------------
// creation / destruction
// ------------------------------------------------------------------------------------------
static int pick_device() {
    // one process per GPU: honour LOCAL_RANK (torchrun) unless QPE_GPU_DEVICE overrides it
    const char *s = std::getenv("QPE_GPU_DEVICE");
    if (!s) s = std::getenv("LOCAL_RANK");
    int d = s ? std::atoi(s) : 0;
    const int n = device_count();
    if (n > 0) d %= n;
    return d;
}

GpuEngine *engine_create(const char *tableName, const char *datafile, int index_slots) {
    if (device_count() <= 0) {
        set_error("no CUDA device is usable by this process (the B200 engine has no CPU fallback)");
        std::fprintf(stderr, "libqpegpu: %s\n", g_last_error.c_str());
        return nullptr;
    }
    GpuEngine *g = new GpuEngine();
    ++g_live_engines;
    g->device = pick_device();
    if (!cuda_ok(cudaSetDevice(g->device), "cudaSetDevice")) {
        delete g;
        return nullptr;
    }
    std::memset(&g->head, 0, sizeof(g->head));
    g->head.tableName = strdup(tableName ? tableName : "");
    g->head.datafile = datafile ? strdup(datafile) : nullptr;
    if (index_slots < 8) index_slots = 8;
    g->idx_slots = index_slots;
    g->head.bplus_tree_roots = static_cast<node **>(std::calloc(index_slots, sizeof(node *)));
    g->head.indexed_attributes = static_cast<char **>(std::calloc(index_slots, sizeof(char *)));
    g->head.attribute_types = static_cast<FieldType *>(std::calloc(index_slots, sizeof(FieldType)));
    g->head.num_indexes = 0;
    g->head.all_records = nullptr;
    g->head.record_block = nullptr;
    g->head.num_records = 0;

    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    bool ok = cuda_ok(cudaStreamCreateWithPriority(&g->stream, cudaStreamNonBlocking, prio_greatest), "cudaStreamCreate");
    ok = ok && cuda_ok(cudaStreamCreateWithPriority(&g->stream2, cudaStreamNonBlocking, prio_least), "cudaStreamCreate");
    for (int i = 0; ok && i < kMaxPipeSegments; ++i)
        ok = cuda_ok(cudaEventCreateWithFlags(&g->ev_seg[i], cudaEventDisableTiming), "cudaEventCreate");
    if (const char *e = std::getenv("QPE_PIPE_SEGMENTS")) g->pipe_segments = std::atoi(e);
    ok = ok && cuda_ok(cudaHostAlloc(&g->h_progress, sizeof(unsigned long long) * (kMaxProgressSegments + 1),
                                     cudaHostAllocMapped),
                       "cudaHostAlloc progress");
    if (ok) {
        std::memset(g->h_progress, 0, sizeof(unsigned long long) * (kMaxProgressSegments + 1));
        ok = cuda_ok(cudaHostGetDevicePointer(&g->d_progress, g->h_progress, 0), "cudaHostGetDevicePointer");
    }
    for (int i = 0; i < GpuEngine::kTimingRing && ok; ++i)
        for (int k = 0; k < 4 && ok; ++k) ok = cuda_ok(cudaEventCreate(&g->ring[i].ev[k]), "cudaEventCreate");
    if (ok) {
        g->ev0 = g->ring[0].ev[0];
        g->ev_mid = g->ring[0].ev[1];
        g->ev1 = g->ring[0].ev[2];
        g->ev_post = g->ring[0].ev[3];
    }
    ok = ok && cuda_ok(cudaMalloc(&g->d_ctl, sizeof(QueryCtl)), "cudaMalloc ctl");
    ok = ok && cuda_ok(cudaMalloc(&g->d_fctl, sizeof(FusedCtl)), "cudaMalloc ctl") &&
         cuda_ok(cudaMemset(g->d_fctl, 0, sizeof(FusedCtl)), "cudaMemset ctl");
    ok = ok && cuda_ok(cudaMallocHost(&g->h_ctl, sizeof(QueryCtl)), "cudaMallocHost ctl");
    ok = ok && cuda_ok(cudaMalloc(&g->d_probe_lo, sizeof(unsigned long long) * kMaxSegments), "cudaMalloc probe");
    ok = ok && cuda_ok(cudaMalloc(&g->d_probe_hi, sizeof(unsigned long long) * kMaxSegments), "cudaMalloc probe");
    ok = ok && cuda_ok(cudaMalloc(&g->d_probe_first, sizeof(uint32_t) * kMaxSegments), "cudaMalloc probe");
    ok = ok && cuda_ok(cudaMalloc(&g->d_probe_count, sizeof(uint32_t) * kMaxSegments), "cudaMalloc probe");
    ok = ok && cuda_ok(cudaMallocHost(&g->h_probe_keys, sizeof(unsigned long long) * 2 * kMaxSegments),
                       "cudaMallocHost probe");
    ok = ok && cuda_ok(cudaMallocHost(&g->h_probe_out, sizeof(uint32_t) * 2 * kMaxSegments), "cudaMallocHost probe");
    if (!ok) {
        engine_destroy(g);
        return nullptr;
    }
    return g;
}

static void free_table(DevTable *t) {
    for (int c = 0; c < NUM_COLS; ++c) {
        if (t->col[c].d) cudaFree(t->col[c].d);
        t->col[c] = DevColumn();
    }
    t->n = 0;
}

void engine_destroy(GpuEngine *g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    if (g->stream2) cudaStreamSynchronize(g->stream2);
    shard_destroy(g);
    free_table(&g->table);
    for (auto &ix : g->idx) index_free(&ix);
    if (g->d_ctl) cudaFree(g->d_ctl);
    if (g->d_fctl) cudaFree(g->d_fctl);
    if (g->d_trace) cudaFree(g->d_trace);
    if (g->h_ctl) cudaFreeHost(g->h_ctl);
    if (g->h_progress) cudaFreeHost(g->h_progress);
    if (g->d_tile_desc) cudaFree(g->d_tile_desc);
    if (g->d_ids) cudaFree(g->d_ids);
    if (g->d_bitmap) cudaFree(g->d_bitmap);
    if (g->d_fmt) cudaFree(g->d_fmt);
    if (g->h_bprogs) cudaFreeHost(g->h_bprogs);
    if (g->d_bprogs) cudaFree(g->d_bprogs);
    if (g->d_bctl) cudaFree(g->d_bctl);
    if (g->d_bcounts) cudaFree(g->d_bcounts);
    if (g->h_bcounts) cudaFreeHost(g->h_bcounts);
    if (g->d_bbitmaps) cudaFree(g->d_bbitmaps);
    if (g->d_probe_lo) cudaFree(g->d_probe_lo);
    if (g->d_probe_hi) cudaFree(g->d_probe_hi);
    if (g->d_probe_first) cudaFree(g->d_probe_first);
    if (g->d_probe_count) cudaFree(g->d_probe_count);
    if (g->h_probe_keys) cudaFreeHost(g->h_probe_keys);
    if (g->h_probe_out) cudaFreeHost(g->h_probe_out);
    if (g->d_probe_scratch) cudaFree(g->d_probe_scratch);
    if (g->d_dml_scratch) cudaFree(g->d_dml_scratch);
    if (g->d_segs) cudaFree(g->d_segs);
    if (g->h_row_stage) cudaFreeHost(g->h_row_stage);
    if (g->h_probe_bounce) cudaFreeHost(g->h_probe_bounce);
    for (int i = 0; i < GpuEngine::kTimingRing; ++i)
        for (int k = 0; k < 4; ++k)
            if (g->ring[i].ev[k]) cudaEventDestroy(g->ring[i].ev[k]);
    for (int i = 0; i < kMaxPipeSegments; ++i)
        if (g->ev_seg[i]) cudaEventDestroy(g->ev_seg[i]);
    if (g->stream2) cudaStreamDestroy(g->stream2);
    if (g->stream) cudaStreamDestroy(g->stream);
    for (int i = 0; i < g->head.num_indexes; ++i) std::free(g->head.indexed_attributes[i]);
    std::free(g->head.indexed_attributes);
    std::free(g->head.attribute_types);
    std::free(g->head.bplus_tree_roots);
    std::free(g->head.tableName);
    std::free(g->head.datafile);
    g->magic = 0;
    delete g;
    --g_live_engines;
}

// rows to allocate for a table of n rows: head-room for INSERTs plus one full tile of slack so
// a bulk copy of the last (partial) tile never leaves the allocation
static int64_t cap_for(int64_t n) {
    int64_t want = n + n / 16 + 1;
    want = (want + kRowPad - 1) / kRowPad * kRowPad;
    return want + kRowPad;
}

bool column_alloc(DevColumn *col, uint32_t width, int64_t cap_rows, cudaStream_t stream) {
    col->d = nullptr;
    col->width = width;
    col->cap = cap_rows;
    const size_t bytes = static_cast<size_t>(cap_rows) * width;
    if (!cuda_ok(cudaMalloc(&col->d, bytes), "cudaMalloc column")) return false;
    return cuda_ok(cudaMemsetAsync(col->d, 0, bytes, stream), "cudaMemset column");
}

bool engine_upload(GpuEngine *g, const HostColumns &hc) {
    cudaSetDevice(g->device);
    free_table(&g->table);
    const int64_t cap = cap_for(hc.n);
    for (int c = 0; c < NUM_COLS; ++c) {
        if (!column_alloc(&g->table.col[c], hc.width[c], cap, g->stream)) return false;
        if (hc.n > 0 &&
            !cuda_ok(cudaMemcpyAsync(g->table.col[c].d, hc.data[c].data(), static_cast<size_t>(hc.n) * hc.width[c],
                                     cudaMemcpyHostToDevice, g->stream),
                     "upload column"))
            return false;
    }
    if (!cuda_ok(cudaStreamSynchronize(g->stream), "upload sync")) return false;
    g->table.n = hc.n;
    g->table.row_base = 0;
    g->head.num_records = static_cast<int>(hc.n);
    for (auto &ix : g->idx) ix.dirty = true;
    return true;
}

// makeIndexSerial (buildEngine-serial.c:13-31): the attribute is appended to the engine's index
// arrays whatever its type; only u64 / int indexes are ever probed (executeEngine-serial.c:425-429).
bool engine_add_index(GpuEngine *g, const char *name, int attributeType) {
    if (!name) return false;
    if (g->head.num_indexes >= g->idx_slots) {
        const int ns = g->idx_slots * 2;
        g->head.bplus_tree_roots = static_cast<node **>(std::realloc(g->head.bplus_tree_roots, ns * sizeof(node *)));
        g->head.indexed_attributes =
            static_cast<char **>(std::realloc(g->head.indexed_attributes, ns * sizeof(char *)));
        g->head.attribute_types =
            static_cast<FieldType *>(std::realloc(g->head.attribute_types, ns * sizeof(FieldType)));
        for (int i = g->idx_slots; i < ns; ++i) {
            g->head.bplus_tree_roots[i] = nullptr;
            g->head.indexed_attributes[i] = nullptr;
        }
        g->idx_slots = ns;
    }
    const int slot = g->head.num_indexes;
    g->head.bplus_tree_roots[slot] = nullptr;
    g->head.indexed_attributes[slot] = strdup(name);
    g->head.attribute_types[slot] =
        (attributeType >= 0 && attributeType <= 3) ? static_cast<FieldType>(attributeType) : static_cast<FieldType>(-1);
    g->head.num_indexes = slot + 1;

    DevIndex ix;
    ix.col = col_by_name(name);
    ix.type = (attributeType == 0) ? T_U64 : T_I32;
    // probed only when the declared type is u64/int AND it is the column's real type
    ix.usable = ix.col >= 0 && (attributeType == 0 || attributeType == 1) &&
                static_cast<int>(kCols[ix.col].type) == attributeType && g->table.resident(ix.col);
    ix.dirty = true;
    g->idx.push_back(ix);
    DevIndex &ref = g->idx.back();
    if (ref.usable) {
        int launches = 0;
        cudaSetDevice(g->device);
        if (!cuda_ok(index_build(&ref, g->table, g->stream, &launches), "index build")) return false;
    }
    return true;
}

static bool ensure_index(GpuEngine *g, DevIndex *ix, int *launches) {
    if (!ix->usable || !ix->dirty) return true;
    return cuda_ok(index_build(ix, g->table, g->stream, launches), "index rebuild");
}

bool index_ready(GpuEngine *g, DevIndex *ix, int *launches) { return ensure_index(g, ix, launches); }

bool engine_ensure_ids(GpuEngine *g, int64_t n) {
    if (n <= g->ids_cap && g->d_ids) return true;
    if (g->d_ids) cudaFree(g->d_ids);
    g->d_ids = nullptr;
    g->ids_cap = 0;
    const int64_t cap = n + n / 8 + 1024;
    if (!cuda_ok(cudaMalloc(&g->d_ids, static_cast<size_t>(cap) * sizeof(uint32_t)), "cudaMalloc ids")) return false;
    g->ids_cap = cap;
    return true;
}

static bool ensure_desc(GpuEngine *g, int64_t n_tiles) {
    if (n_tiles <= g->desc_cap && g->d_tile_desc) return true;
    if (g->d_tile_desc) cudaFree(g->d_tile_desc);
    g->d_tile_desc = nullptr;
    const int64_t cap = n_tiles + n_tiles / 4 + 1024;
    if (!cuda_ok(cudaMalloc(&g->d_tile_desc, static_cast<size_t>(cap) * 8), "cudaMalloc descriptors")) return false;
    if (!cuda_ok(cudaMemsetAsync(g->d_tile_desc, 0, static_cast<size_t>(cap) * 8, g->stream), "memset descriptors"))
        return false;
    g->desc_cap = cap;
    return true;
}

static bool ensure_bitmap(GpuEngine *g, int64_t words) {
    if (words <= g->bitmap_cap_words && g->d_bitmap) return true;
    if (g->d_bitmap) cudaFree(g->d_bitmap);
    g->d_bitmap = nullptr;
    const int64_t cap = words + 1024;
    if (!cuda_ok(cudaMalloc(&g->d_bitmap, static_cast<size_t>(cap) * 4), "cudaMalloc bitmap")) return false;
    g->bitmap_cap_words = cap;
    return true;
}

static uint32_t next_epoch(GpuEngine *g) {
    // descriptors carry (epoch << 2 | state) in their upper word: 30 usable bits, 0 = never written
    if (g->epoch >= 0x3ffffff0u) {
        if (g->d_tile_desc) cudaMemsetAsync(g->d_tile_desc, 0, static_cast<size_t>(g->desc_cap) * 8, g->stream);
        g->epoch = 0;
    }
    return ++g->epoch;
}

// ------------------------------------------------------------------------------------------
// index-path planning: executeEngine-serial.c:358-459
// ------------------------------------------------------------------------------------------
static int plan_segments(const GpuEngine *g, const struct whereClauseS *wc, SegmentPlan *out, int max_out,
                         bool *too_many) {
    int n = 0;
    *too_many = false;
    for (const struct whereClauseS *w = wc; w; w = w->next) {
        if (w->attribute == nullptr) continue;  // parenthesised groups are skipped (:361-364)
        for (int i = 0; i < g->head.num_indexes; ++i) {
            if (std::strcmp(w->attribute, g->head.indexed_attributes[i]) != 0) continue;
            const FieldType type = g->head.attribute_types[i];
            if (type != FIELD_UINT64 && type != FIELD_INT) continue;  // bool / string: unsupported (:425-429)
            const DevIndex &ix = g->idx[i];
            if (!ix.usable) {
                // declared u64 / int, so the reference WOULD take the index path here (:366-424) and return its order;
                // this engine cannot probe it (declared type differs from the column's, or the column is not
                // resident): say so instead of silently answering in table order
                *too_many = true;
                set_error(std::string("index on '") + w->attribute + "' cannot be probed (declared type differs from the "
                          "column's type, or the column is not resident): the reference would use it for this WHERE");
                return -1;
            }
            const char *op = w->op_ ? w->op_ : "";
            const char *val = w->value ? w->value : "";
            SegmentPlan s{};
            s.index_slot = i;
            if (type == FIELD_UINT64) {
                const unsigned long long v = std::strtoull(val, nullptr, 10);
                s.is_u64 = true;
                if (!std::strcmp(op, "=")) {
                    s.lo_u64 = v; s.hi_u64 = v;
                } else if (!std::strcmp(op, ">")) {
                    s.lo_u64 = v + 1; s.hi_u64 = UINT64_MAX;  // wraps at v == MAX exactly like the reference
                } else if (!std::strcmp(op, ">=")) {
                    s.lo_u64 = v; s.hi_u64 = UINT64_MAX;
                } else if (!std::strcmp(op, "<")) {
                    s.lo_u64 = 0; s.hi_u64 = v - 1;           // wraps at v == 0 exactly like the reference
                } else if (!std::strcmp(op, "<=")) {
                    s.lo_u64 = 0; s.hi_u64 = v;
                } else {
                    s.lo_u64 = 0; s.hi_u64 = UINT64_MAX;
                }
            } else {
                const int v = std::atoi(val);
                s.is_u64 = false;
                // v + 1 / v - 1 overflow is UB in the reference; wrap-around is what gcc -O2 emits
                const int vp = static_cast<int>(static_cast<unsigned>(v) + 1u);
                const int vm = static_cast<int>(static_cast<unsigned>(v) - 1u);
                if (!std::strcmp(op, "=")) {
                    s.lo_i32 = v; s.hi_i32 = v;
                } else if (!std::strcmp(op, ">")) {
                    s.lo_i32 = vp; s.hi_i32 = INT_MAX;
                } else if (!std::strcmp(op, ">=")) {
                    s.lo_i32 = v; s.hi_i32 = INT_MAX;
                } else if (!std::strcmp(op, "<")) {
                    s.lo_i32 = INT_MIN; s.hi_i32 = vm;
                } else if (!std::strcmp(op, "<=")) {
                    s.lo_i32 = INT_MIN; s.hi_i32 = v;
                } else {
                    s.lo_i32 = INT_MIN; s.hi_i32 = INT_MAX;
                }
            }
            if (n >= max_out) {
                *too_many = true;
                return n;
            }
            out[n++] = s;
        }
    }
    return n;
}

// ------------------------------------------------------------------------------------------
// lazy event timing
// ------------------------------------------------------------------------------------------
static void resolve_slot(GpuEngine *g, GpuEngine::TimingSlot *slot, ScanStats *into) {
    float ms = 0.f, a = 0.f, b = 0.f, pm = 0.f;
    cudaEventElapsedTime(&ms, slot->ev[0], slot->ev[2]);
    if (slot->staged) {
        cudaEventElapsedTime(&a, slot->ev[0], slot->ev[1]);
        if (slot->two_kernels) cudaEventElapsedTime(&b, slot->ev[1], slot->ev[2]);
    }
    if (slot->has_post) cudaEventElapsedTime(&pm, slot->ev[2], slot->ev[3]);
    slot->pending = false;
    if (into) {
        into->kernel_ms = ms;
        into->scan_ms = a;
        into->compact_ms = b;
        g->trace[3] = pm;
    }
    if (g->accumulate_timing) {
        g->acc_kernel_ms += ms;
        g->acc_scan_ms += a;
        g->acc_compact_ms += b;
        g->acc_post_ms += pm;
        g->acc_calls += 1;
    }
}

void engine_resolve_timing(GpuEngine *g) {
    if (g->cur_slot && g->cur_slot->pending) {
        cudaSetDevice(g->device);
        resolve_slot(g, g->cur_slot, &g->last);
    }
}

void engine_resolve_all(GpuEngine *g) {
    cudaSetDevice(g->device);
    for (int i = 0; i < GpuEngine::kTimingRing; ++i) {
        GpuEngine::TimingSlot *slot = &g->ring[i];
        if (slot->pending) resolve_slot(g, slot, slot == g->cur_slot ? &g->last : nullptr);
    }
}

// ------------------------------------------------------------------------------------------
// match phase
// ------------------------------------------------------------------------------------------
constexpr int kFusedMaxStages = 16;

bool engine_match(GpuEngine *g, const struct whereClauseS *wc, bool force_scan, bool invert, bool count_only,
                  bool want_bitmap, uint64_t *count) {
    cudaSetDevice(g->device);
    const double t_begin = now_ms();
    ScanStats st;
    const DevTable &t = g->table;
    g->last_bm_words = 0;
    bool post_done = false;
    bool index_synced = false;  // the index path has already synchronised (count and candidates are on the host)
    const unsigned long long *count_src = g->count_mapped;  // where the match count arrives without a download (or null)
    g->count_dev = &g->d_ctl->out_count;                    // (K1f: its own control block, set below)
    {
        GpuEngine::TimingSlot *slot = &g->ring[g->ring_head];
        g->ring_head = (g->ring_head + 1) % GpuEngine::kTimingRing;
        if (slot->pending && g->accumulate_timing) resolve_slot(g, slot, nullptr);
        slot->pending = false;
        if (g->cur_slot && g->cur_slot->pending && !g->accumulate_timing) g->cur_slot->pending = false;  // never asked for
        g->cur_slot = slot;
        g->ev0 = slot->ev[0];
        g->ev_mid = slot->ev[1];
        g->ev1 = slot->ev[2];
        g->ev_post = slot->ev[3];
    }

    uint32_t widths[NUM_COLS];
    for (int c = 0; c < NUM_COLS; ++c) widths[c] = t.col[c].width;
    QueryCtl *hc = g->h_ctl;
    const std::string err = compile_where(wc, widths, &hc->prog, invert);
    if (!err.empty()) {
        set_error(err);
        return false;
    }
    for (int c = 0; c < NUM_COLS; ++c)
        if ((hc->prog.col_mask & (1u << c)) && !t.resident(c)) {
            set_error(std::string("WHERE references column '") + kCols[c].name + "' which is not resident on the device");
            return false;
        }
    const double t_compiled = now_ms();
    hc->tile_counter = 0;
    hc->chunk_counter = 0;
    hc->out_count = 0;
    std::memset(hc->seg_stored, 0, sizeof(hc->seg_stored));
    g->host_out_done = false;

    // ---- path rule of the reference ----
    SegmentPlan segs[kMaxSegments];
    int n_seg = 0;
    if (!force_scan && !invert) {
        bool too_many = false;
        n_seg = plan_segments(g, wc, segs, kMaxSegments, &too_many);
        if (n_seg < 0) return false;  // an index the reference would use cannot be probed (error set)
        if (too_many) {
            set_error("too many (condition x index) segments in one WHERE clause");
            return false;
        }
    }

    int64_t bytes_per_row = 0;
    for (int c = 0; c < NUM_COLS; ++c)
        if (hc->prog.col_mask & (1u << c)) bytes_per_row += t.col[c].width;

    // the compiled query goes up right before the first kernel of whichever path runs (all host-side planning
    // first: the device then never idles between this copy and the launch that waits for it)
    auto upload_query = [&]() {
        return cuda_ok(cudaMemcpyAsync(g->d_ctl, hc, sizeof(QueryCtl), cudaMemcpyHostToDevice, g->stream), "upload query");
    };

    if (n_seg == 0) {
        // ===== full-scan path: linearSearchRecords over the table (:464-467) =====
        st.path = 0;
        st.rows_scanned = t.n;
        ScanGeometry geo{};
        const char *why = nullptr;
        bool staged = scan_plan(t, hc->prog, g->force_tile_rows, g->force_stages, kFusedMaxStages, &geo, &why);
        // Pipelined scan: the table is cut into P segments of whole 64 Ki-row chunks; K1 of segment
        // i+1 (stream, high priority) runs while K1c of segment i (stream2, low priority) compacts
        // and stores its ids -- possibly straight into pinned host memory or a peer GPU -- so the
        // ordered output leaves the GPU during the scan instead of after it.  K1 then runs with at
        // most 3 stages so that a K1c CTA (33 KB of shared memory) fits beside it on every SM.
        // Default for a SELECT's id list: K1f, scan + ordered compaction fused in one launch.
        bool fused_done = false;
        if (staged && !count_only && !want_bitmap && t.n > 0 && g->pipe_segments == 0) {
            ScanGeometry fg{};
            const char *fwhy = nullptr;
            static const int pinned_cw = [] {
                const char *e = std::getenv("QPE_FUSE_CW");
                const int v = e ? std::atoi(e) : 0;
                return (v == 4 || v == 8) ? v : 0;
            }();
            const int want_cw = pinned_cw ? pinned_cw : g->fuse_cw;
            // up to 8 stages: a narrow WHERE (1-8 B/row) is bound by ROWS in flight per SM (stages x tile rows over the
            // ~3 us a stage takes to be refilled and consumed), not by bytes
            bool planned = scan_plan(t, hc->prog, g->force_tile_rows, g->force_stages, kFusedMaxStages, &fg, &fwhy, 4);
            if (planned && want_cw == 8) {
                // eight compaction warps need 16 KB more shared memory: only when that costs no tile rows and at most a quarter of the stages
                ScanGeometry g8{};
                const char *why8 = nullptr;
                if (scan_plan(t, hc->prog, g->force_tile_rows, g->force_stages, kFusedMaxStages, &g8, &why8, 8) &&
                    g8.tile_rows == fg.tile_rows && g8.stages * 4 >= fg.stages * 3)
                    fg = g8;
            }
            if (planned) {
                if (!ensure_desc(g, fg.n_chunks)) return false;
                if (!g->out_override && !engine_ensure_ids(g, t.n)) return false;
                FusedLaunch F{};
                F.scan.table = &t;
                F.scan.d_ctl = g->d_ctl;
                F.scan.h_prog = &hc->prog;
                F.scan.out_bitmap = nullptr;
                F.desc = g->d_tile_desc;
                F.epoch = next_epoch(g);
                F.id_base = (g->out_override || g->id_base_always) ? g->id_base_override : 0u;
                F.out_ids = g->out_override ? g->out_override : g->d_ids;
                F.out_cap = g->out_override ? g->out_override_cap : static_cast<unsigned long long>(g->ids_cap);
                // progress segments only when the ids are wanted on the host: ~16 Mi rows each, at most 16
                int n_prog = 0;
                if (g->host_out && !g->out_override) {
                    n_prog = static_cast<int>(t.n / (16ll << 20));
                    if (n_prog > kMaxProgressSegments) n_prog = kMaxProgressSegments;
                    if (n_prog < 1) n_prog = 1;
                    if (n_prog > fg.n_chunks) n_prog = static_cast<int>(fg.n_chunks);
                    F.seg_chunks = (fg.n_chunks + n_prog - 1) / n_prog;
                    n_prog = static_cast<int>((fg.n_chunks + F.seg_chunks - 1) / F.seg_chunks);
                    F.progress = g->d_progress;
                }
                // no upload: K1f takes the program as a kernel parameter and resets its own control words
                F.d_fctl = g->d_fctl;
                F.l2_stream = g->fuse_l2;
                static const bool want_trace = std::getenv("QPE_FUSE_TRACE") != nullptr;
                if (want_trace && !g->d_trace) cudaMalloc(&g->d_trace, 8 * 8 * 1024);
                F.trace = want_trace ? g->d_trace : nullptr;
                g->count_dev = &g->d_fctl->final_count;
                // a plain (unsharded) scan: the kernel's last CTA writes the count into mapped host memory
                if (!g->count_mapped) {
                    F.host_count = g->d_progress + kMaxProgressSegments;
                    count_src = g->h_progress + kMaxProgressSegments;
                }
                cudaEventRecord(g->ev0, g->stream);
                if (!cuda_ok(fused_launch(F, fg, g->stream), "fused scan kernel launch")) return false;
                cudaEventRecord(g->ev_mid, g->stream);
                cudaEventRecord(g->ev1, g->stream);
                if (g->post_match) {
                    // sharded table: the count exchange follows the scan at once (before the host starts
                    // copying finished segments out), so the other ranks never wait for this rank's copies
                    if (!g->post_match()) return false;
                    cudaEventRecord(g->ev_post, g->stream);
                    post_done = true;
                }
                st.launches = 1;
                st.tile_rows = fg.tile_rows;
                st.stages = fg.stages;
                st.grid = fg.grid;
                if (n_prog > 0) {
                    // copy each finished table segment's ids to the host while the scan goes on
                    volatile unsigned long long *hp = g->h_progress;
                    uint64_t prev = 0;
                    bool kernel_over = false;
                    for (int sgi = 0; sgi < n_prog; ++sgi) {
                        uint32_t spins = 0;
                        while ((hp[sgi] >> 32) != F.epoch) {
                            if (kernel_over) {
                                set_error("fused scan finished without publishing its progress");
                                return false;
                            }
                            if ((++spins & 0x3ffu) == 0) {
                                const cudaError_t q = cudaStreamQuery(g->stream);
                                if (q == cudaSuccess)
                                    kernel_over = true;  // one more look at the word, then give up
                                else if (q != cudaErrorNotReady)
                                    return cuda_ok(q, "fused scan kernel");
                            }
                        }
                        uint64_t cum = hp[sgi] & 0xffffffffull;
                        if (cum > g->host_out_cap) cum = g->host_out_cap;  // the caller reports the overflow
                        if (cum > prev) {
                            if (!cuda_ok(cudaMemcpyAsync(g->host_out + prev, g->d_ids + prev, (cum - prev) * 4,
                                                         cudaMemcpyDeviceToHost, g->stream2),
                                         "download ids"))
                                return false;
                            prev = cum;
                        }
                    }
                    cudaEventRecord(g->ev_seg[0], g->stream2);
                    cudaStreamWaitEvent(g->stream, g->ev_seg[0], 0);
                    g->host_out_done = true;
                }
                fused_done = true;
            }
        }
        int P = 1;
        if (staged && !fused_done && !count_only && !want_bitmap && t.n > 0) {
            const int64_t chunks = compact_chunks(geo.n_tiles * (geo.tile_rows / 32));
            P = g->pipe_segments;
            if (P > chunks) P = static_cast<int>(chunks);
            if (P < 1) P = 1;
            if (P > 1 && !g->force_stages && geo.stages > 3) {
                ScanGeometry g3{};
                if (scan_plan(t, hc->prog, geo.tile_rows, 0, 3, &g3, &why)) geo = g3;
            }
        }
        if (fused_done) {
            g->last_bm_words = 0;
        } else if (staged) {
            const int64_t bm_words = geo.n_tiles * (geo.tile_rows / 32);
            const int64_t n_chunks = compact_chunks(bm_words);
            if (!count_only || want_bitmap) {
                if (!ensure_bitmap(g, bm_words)) return false;
            }
            if (!count_only && !ensure_desc(g, n_chunks)) return false;
            if (!count_only && !g->out_override && !engine_ensure_ids(g, t.n)) return false;
            ScanLaunch L{};
            L.table = &t;
            L.d_ctl = g->d_ctl;
            L.h_prog = &hc->prog;
            L.out_bitmap = (!count_only || want_bitmap) ? g->d_bitmap : nullptr;
            L.force_tile_rows = g->force_tile_rows;
            L.force_stages = g->force_stages;
            uint32_t *dst = g->out_override ? g->out_override : g->d_ids;
            const unsigned long long cap =
                g->out_override ? g->out_override_cap : static_cast<unsigned long long>(g->ids_cap);
            const uint32_t id_base = (g->out_override || g->id_base_always) ? g->id_base_override : 0u;
            if (!upload_query()) return false;
            cudaEventRecord(g->ev0, g->stream);
            if (P == 1) {
                if (!cuda_ok(scan_launch(L, geo, g->stream), "scan kernel launch")) return false;
                st.launches = 1;
                cudaEventRecord(g->ev_mid, g->stream);
                if (!count_only && t.n > 0) {
                    // K1c: ordered compaction of the match bitmap (decoupled look-back per 64 Ki rows)
                    if (!cuda_ok(compact_launch(g->d_bitmap, bm_words, g->d_ctl, g->d_tile_desc, next_epoch(g), dst,
                                                id_base, cap, g->stream),
                                 "compaction kernel launch"))
                        return false;
                    st.launches = 2;
                }
                cudaEventRecord(g->ev1, g->stream);
            } else {
                const uint32_t epoch = next_epoch(g);
                const int64_t tiles_per_chunk = kCompactChunkRows / geo.tile_rows;
                const int64_t seg_chunks = (n_chunks + P - 1) / P;
                int launched = 0;
                for (int i = 0; i < P; ++i) {
                    const int64_t c0 = static_cast<int64_t>(i) * seg_chunks;
                    const int64_t c1 = (c0 + seg_chunks < n_chunks) ? c0 + seg_chunks : n_chunks;
                    if (c0 >= c1) break;
                    L.tile_begin = c0 * tiles_per_chunk;
                    L.tile_end = (c1 * tiles_per_chunk < geo.n_tiles) ? c1 * tiles_per_chunk : geo.n_tiles;
                    if (!cuda_ok(scan_launch(L, geo, g->stream), "scan kernel launch")) return false;
                    cudaEventRecord(g->ev_seg[i], g->stream);
                    cudaStreamWaitEvent(g->stream2, g->ev_seg[i], 0);
                    if (!cuda_ok(compact_launch(g->d_bitmap, bm_words, g->d_ctl, g->d_tile_desc, epoch, dst, id_base, cap,
                                                g->stream2, c1 - c0),
                                 "compaction kernel launch"))
                        return false;
                    launched += 2;
                }
                cudaEventRecord(g->ev_mid, g->stream);   // the last K1 is done
                cudaEventRecord(g->ev1, g->stream2);     // the last K1c is done
                cudaStreamWaitEvent(g->stream, g->ev1, 0);
                st.launches = launched;
            }
            st.tile_rows = geo.tile_rows;
            st.stages = geo.stages;
            st.grid = geo.grid;
            g->last_bm_words = L.out_bitmap ? bm_words : 0;
        } else {
            // a tile of this WHERE's columns does not fit shared memory: evaluate with gathered
            // loads over the identity candidate list (still on the device, still ordered)
            if (want_bitmap) {
                set_error(std::string("match mask unavailable: ") + (why ? why : "scan cannot be staged"));
                return false;
            }
            CandSegments cs{};
            cs.n_seg = 1;
            cs.perm[0] = nullptr;
            cs.first[0] = 0;
            cs.vstart[0] = 0;
            cs.vstart[1] = t.n;
            if (!ensure_desc(g, filter_tiles(t.n))) return false;
            if (!count_only && !engine_ensure_ids(g, t.n)) return false;
            if (!upload_query()) return false;
            cudaEventRecord(g->ev0, g->stream);
            if (!cuda_ok(filter_launch(t, g->d_ctl, cs, g->d_tile_desc, next_epoch(g), count_only ? nullptr : g->d_ids,
                                       g->stream),
                         "filter kernel launch"))
                return false;
            cudaEventRecord(g->ev1, g->stream);
            st.launches = t.n > 0 ? 1 : 0;
        }
    } else {
        // ===== index path: candidates from findRange per segment, then the whole WHERE (:469-474) =====
        st.path = 1;
        if (want_bitmap) {
            set_error("match mask is only defined on the full-scan path");
            return false;
        }
        int launches = 0;
        for (int s = 0; s < n_seg; ++s)
            if (!ensure_index(g, &g->idx[segs[s].index_slot], &launches)) return false;
        // K3s (all segments' probes, one launch, keys as kernel parameters) -> segment table in device memory -> K1g
        // (persistent, reads the table there): ONE synchronisation per indexed SELECT.  The host only knows an upper
        // bound of the candidates (the sum of the probed indexes' sizes); the id buffer holds a table's worth of ids,
        // as the reference's candidate array does (:342), and a result that does not fit is run again with room.
        const DevIndex *ixs[kMaxSegments];
        unsigned long long lo[kMaxSegments], hi[kMaxSegments];
        long long bound = 0;
        for (int s = 0; s < n_seg; ++s) {
            ixs[s] = &g->idx[segs[s].index_slot];
            bound += ixs[s]->n;
            if (segs[s].is_u64) {
                lo[s] = segs[s].lo_u64;
                hi[s] = segs[s].hi_u64;
            } else {
                lo[s] = static_cast<uint32_t>(segs[s].lo_i32);
                hi[s] = static_cast<uint32_t>(segs[s].hi_i32);
            }
        }
        if (!g->d_segs && !cuda_ok(cudaMalloc(&g->d_segs, sizeof(CandSegments)), "cudaMalloc segment table")) return false;
        if (!ensure_desc(g, filter_tiles(bound))) return false;
        if (!count_only && !engine_ensure_ids(g, t.n > 0 ? t.n : 1)) return false;
        long long n_cand = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (!upload_query()) return false;
            cudaEventRecord(g->ev0, g->stream);
            if (!cuda_ok(index_probe_segments(ixs, lo, hi, n_seg, g->d_segs, g->stream), "probe kernel launch")) return false;
            if (!cuda_ok(filter_launch(t, g->d_ctl, CandSegments{}, g->d_tile_desc, next_epoch(g), count_only ? nullptr : g->d_ids,
                                       g->stream, g->d_segs, bound, static_cast<unsigned long long>(g->ids_cap)),
                         "filter kernel launch"))
                return false;
            cudaEventRecord(g->ev1, g->stream);
            launches += 2;
            if (!cuda_ok(cudaMemcpyAsync(&g->h_probe_out[0], &g->d_segs->vstart[n_seg], sizeof(long long), cudaMemcpyDeviceToHost, g->stream),
                         "download candidates") ||
                !cuda_ok(cudaMemcpyAsync(&hc->out_count, &g->d_ctl->out_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, g->stream),
                         "download count") ||
                !cuda_ok(cudaStreamSynchronize(g->stream), "match sync"))
                return false;
            std::memcpy(&n_cand, &g->h_probe_out[0], sizeof(long long));
            if (count_only || static_cast<int64_t>(hc->out_count) <= g->ids_cap) break;
            // more survivors than a table's worth of ids (duplicated segments): make room, run once more
            if (attempt == 1 || !engine_ensure_ids(g, static_cast<int64_t>(hc->out_count))) {
                set_error("indexed SELECT: result does not fit the id buffer");
                return false;
            }
            hc->tile_counter = 0;
            hc->out_count = 0;
        }
        st.candidates = n_cand;
        st.rows_scanned = n_cand;
        st.launches = launches;
        index_synced = true;
    }

    if (g->post_match && !post_done) {
        if (!g->post_match()) return false;
        cudaEventRecord(g->ev_post, g->stream);
    }
    if (!index_synced && !count_src &&
        !cuda_ok(cudaMemcpyAsync(&hc->out_count, g->count_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, g->stream),
                 "download count"))
        return false;
    const double t_enq = now_ms();
    if (!index_synced && !cuda_ok(cudaStreamSynchronize(g->stream), "match sync")) return false;
    if (count_src && !index_synced) hc->out_count = *static_cast<const volatile unsigned long long *>(count_src);
    const double t_sync = now_ms();
    g->trace[0] = t_compiled - t_begin;
    g->trace[1] = t_enq - t_compiled;
    g->trace[2] = t_sync - t_enq;
    g->trace[3] = 0;
    // kernel_ms / scan_ms / compact_ms: resolved lazily from the slot's events (engine_resolve_timing)
    g->cur_slot->pending = true;
    g->cur_slot->staged = st.path == 0 && st.tile_rows > 0;
    g->cur_slot->two_kernels = st.launches > 1;
    g->cur_slot->has_post = static_cast<bool>(g->post_match);
    st.matches = static_cast<int64_t>(hc->out_count);
    // next full scan: see engine_adapt_fused
    if (st.path == 0 && st.rows_scanned > 0) engine_adapt_fused(g, st.matches, st.rows_scanned);
    st.algo_bytes = st.rows_scanned * bytes_per_row + (count_only ? 0 : 4 * st.matches) +
                    (st.path == 1 ? 4 * st.candidates : 0);
    st.total_ms = now_ms() - t_begin;
    g->trace[4] = now_ms() - t_sync;  // tail of the call: event queries + statistics
    g->last = st;
    g->last_bm_count = hc->out_count;
    if (count) *count = hc->out_count;
    return true;
}

// ------------------------------------------------------------------------------------------
// K1f without the synchronisation: compile + plan + launch on the engine's stream and return.  The sharded
// SELECT (shard.cu) enqueues its count exchange / delivery kernels right behind and may enqueue the NEXT
// query before it waits for this one.  The match count ends up in g->d_fctl->final_count (device).
// ------------------------------------------------------------------------------------------
static GpuEngine::TimingSlot *take_timing_slot(GpuEngine *g) {
    GpuEngine::TimingSlot *slot = &g->ring[g->ring_head];
    g->ring_head = (g->ring_head + 1) % GpuEngine::kTimingRing;
    if (slot->pending && g->accumulate_timing) resolve_slot(g, slot, nullptr);
    slot->pending = false;
    if (g->cur_slot && g->cur_slot->pending && !g->accumulate_timing) g->cur_slot->pending = false;  // never asked for
    return slot;
}

// How the NEXT full scan of this engine runs, from what this one matched (the same statement usually comes again: a
// driver's loop, the ranks of a sharded table).  8 compaction warps from 1/16 of the rows on, and between 1/16 and 1/4 the
// table is streamed through L2 with the evict_first policy, so that the ids the scan writes stay in the 126 MB L2 instead
// of competing with the reads for HBM.  Measured with the final K1f (tickets, aggregates from the evaluators) on a
// 125 M-row shard whose first 10 M rows match at 94 % -- rank 0 of an 8-GPU QN: 273 us with 4 warps, 268 with 8, 267.6
// with 4 + the policy, 260.4 with 8 + the policy (a shard without matches: 246-250).  At 1 % the policy costs 1-3 %
// (1.803 vs 1.788 ms on 1 B rows), at 50 % 1-2 % (the ids no longer fit L2).  (Round 1 measured the 7.6 % shard faster with 4
// warps, 0.302 vs 0.310 ms: that was the look-back convoy, not the warps.)
void engine_adapt_fused(GpuEngine *g, int64_t matches, int64_t rows) {
    g->fuse_cw = (matches * 16 > rows) ? 8 : 4;
    g->fuse_l2 = matches * 16 > rows && matches * 4 < rows;
}

bool engine_fused_enqueue(GpuEngine *g, const struct whereClauseS *wc, uint32_t *out_ids, uint64_t out_cap,
                          uint32_t id_base, FusedEnqueue *fe, bool overlap_previous, bool l2_stream) {
    cudaSetDevice(g->device);
    const DevTable &t = g->table;
    fe->slot = nullptr;
    fe->launched = false;
    fe->geo = ScanGeometry{};
    fe->bytes_per_row = 0;
    uint32_t widths[NUM_COLS];
    for (int c = 0; c < NUM_COLS; ++c) widths[c] = t.col[c].width;
    Program *prog = &g->h_ctl->prog;  // copied into the kernel's parameter block by the launch: reusable at once
    const std::string err = compile_where(wc, widths, prog, false);
    if (!err.empty()) {
        set_error(err);
        return false;
    }
    for (int c = 0; c < NUM_COLS; ++c)
        if (prog->col_mask & (1u << c)) {
            if (!t.resident(c)) {
                set_error(std::string("WHERE references column '") + kCols[c].name + "' which is not resident on the device");
                return false;
            }
            fe->bytes_per_row += t.col[c].width;
        }
    // an overlapped scan is not timed: an event between it and the stream's previous kernel would serialise the two
    GpuEngine::TimingSlot *slot = overlap_previous ? nullptr : take_timing_slot(g);
    fe->slot = slot;
    if (t.n == 0) {
        // an empty shard still takes part: its count is 0
        if (slot) cudaEventRecord(slot->ev[0], g->stream);
        if (!cuda_ok(cudaMemsetAsync(&g->d_fctl->final_count, 0, sizeof(unsigned long long), g->stream), "reset count"))
            return false;
        if (slot) {
            cudaEventRecord(slot->ev[1], g->stream);
            cudaEventRecord(slot->ev[2], g->stream);
        }
        return true;
    }
    const char *why = nullptr;
    ScanGeometry fg{};
    if (!scan_plan(t, *prog, g->force_tile_rows, g->force_stages, kFusedMaxStages, &fg, &why, 4)) {
        set_error(std::string("this WHERE cannot be staged by the scan kernel: ") + (why ? why : "row too wide"));
        return false;
    }
    if (g->fuse_cw == 8) {
        ScanGeometry g8{};
        const char *why8 = nullptr;
        if (scan_plan(t, *prog, g->force_tile_rows, g->force_stages, kFusedMaxStages, &g8, &why8, 8) &&
            g8.tile_rows == fg.tile_rows && g8.stages * 4 >= fg.stages * 3)
            fg = g8;
    }
    if (!ensure_desc(g, fg.n_chunks)) return false;
    FusedLaunch F{};
    F.scan.table = &t;
    F.scan.d_ctl = g->d_ctl;
    F.scan.h_prog = prog;
    F.desc = g->d_tile_desc;
    F.epoch = next_epoch(g);
    F.id_base = id_base;
    F.out_ids = out_ids;
    F.out_cap = out_cap;
    F.d_fctl = g->d_fctl;
    F.pdl = overlap_previous;
    F.l2_stream = l2_stream || g->fuse_l2;
    if (slot) cudaEventRecord(slot->ev[0], g->stream);
    if (!cuda_ok(fused_launch(F, fg, g->stream), "fused scan kernel launch")) return false;
    if (slot) {
        cudaEventRecord(slot->ev[1], g->stream);
        cudaEventRecord(slot->ev[2], g->stream);
    }
    fe->geo = fg;
    fe->launched = true;
    return true;
}

// bookkeeping once the match count of an enqueued fused scan is known (statistics, compaction-warp choice)
void engine_fused_finish(GpuEngine *g, const FusedEnqueue &fe, uint64_t matches, int extra_launches, double t_begin_ms) {
    ScanStats st;
    st.path = 0;
    st.rows_scanned = g->table.n;
    st.matches = static_cast<int64_t>(matches);
    st.launches = (fe.launched ? 1 : 0) + extra_launches;
    st.tile_rows = fe.geo.tile_rows;
    st.stages = fe.geo.stages;
    st.grid = fe.geo.grid;
    st.algo_bytes = st.rows_scanned * fe.bytes_per_row + 4 * st.matches;
    st.total_ms = now_ms() - t_begin_ms;
    if (st.rows_scanned > 0) engine_adapt_fused(g, st.matches, st.rows_scanned);
    g->last = st;
    if (!fe.slot) g->cur_slot = nullptr;  // an overlapped (untimed) scan: no event times to resolve for it
    if (fe.slot) {
        g->cur_slot = fe.slot;
        fe.slot->pending = true;
        fe.slot->staged = fe.launched;
        fe.slot->two_kernels = false;
        fe.slot->has_post = true;
        g->ev0 = fe.slot->ev[0];
        g->ev_mid = fe.slot->ev[1];
        g->ev1 = fe.slot->ev[2];
        g->ev_post = fe.slot->ev[3];
    }
}

// ------------------------------------------------------------------------------------------
// K9: query batch (the reference's query-level parallelism, QPEOMP.c:234-335)
// ------------------------------------------------------------------------------------------
bool engine_uses_index(GpuEngine *g, const struct whereClauseS *wc) {
    SegmentPlan segs[kMaxSegments];
    bool too_many = false;
    return plan_segments(g, wc, segs, kMaxSegments, &too_many) > 0 || too_many;
}

bool engine_where_columns(GpuEngine *g, const struct whereClauseS *wc, uint32_t *mask) {
    uint32_t widths[NUM_COLS];
    for (int c = 0; c < NUM_COLS; ++c) widths[c] = g->table.col[c].width;
    static thread_local Program tmp;
    const std::string err = compile_where(wc, widths, &tmp, false);
    if (!err.empty()) {
        set_error(err);
        return false;
    }
    *mask = tmp.col_mask;
    return true;
}

bool engine_match_batch(GpuEngine *g, const struct whereClauseS *const *wcs, int nq, uint64_t *offsets) {
    cudaSetDevice(g->device);
    if (nq < 1 || nq > kMaxBatch) {
        set_error("query batch: between 1 and 8 queries per pass");
        return false;
    }
    const double t_begin = now_ms();
    const DevTable &t = g->table;
    g->last_bm_words = 0;
    if (!g->h_bprogs) {
        bool ok = cuda_ok(cudaHostAlloc(&g->h_bprogs, sizeof(Program) * kMaxBatch, cudaHostAllocDefault), "cudaHostAlloc batch") &&
                  cuda_ok(cudaMalloc(&g->d_bprogs, sizeof(Program) * kMaxBatch), "cudaMalloc batch") &&
                  cuda_ok(cudaMalloc(&g->d_bctl, sizeof(QueryCtl) * kMaxBatch), "cudaMalloc batch") &&
                  cuda_ok(cudaMalloc(&g->d_bcounts, sizeof(unsigned long long) * kMaxBatch), "cudaMalloc batch") &&
                  cuda_ok(cudaHostAlloc(&g->h_bcounts, sizeof(unsigned long long) * kMaxBatch, cudaHostAllocDefault),
                          "cudaHostAlloc batch");
        if (!ok) return false;
    }
    uint32_t widths[NUM_COLS];
    for (int c = 0; c < NUM_COLS; ++c) widths[c] = t.col[c].width;
    Program uni{};  // only its column mask is used: the union of what the queries reference
    for (int q = 0; q < nq; ++q) {
        const std::string err = compile_where(wcs[q], widths, &g->h_bprogs[q], false);
        if (!err.empty()) {
            set_error(err);
            return false;
        }
        uni.col_mask |= g->h_bprogs[q].col_mask;
    }
    int64_t bytes_per_row = 0;
    for (int c = 0; c < NUM_COLS; ++c)
        if (uni.col_mask & (1u << c)) {
            if (!t.resident(c)) {
                set_error(std::string("WHERE references column '") + kCols[c].name + "' which is not resident on the device");
                return false;
            }
            bytes_per_row += t.col[c].width;
        }
    ScanGeometry geo{};
    const char *why = nullptr;
    if (!scan_plan(t, uni, g->force_tile_rows, g->force_stages, kFusedMaxStages, &geo, &why, 0, batch_smem_bytes(nq))) {
        set_error(std::string("query batch cannot be staged: ") + (why ? why : "row too wide"));
        return false;
    }
    const int64_t bm_words = geo.n_tiles * (geo.tile_rows / 32);
    const int64_t n_chunks = compact_chunks(bm_words);
    const int64_t bm_stride = (bm_words + 63) & ~int64_t(63);
    if (bm_stride * nq > g->bbitmap_cap_words || !g->d_bbitmaps) {
        if (g->d_bbitmaps) cudaFree(g->d_bbitmaps);
        g->d_bbitmaps = nullptr;
        g->bbitmap_cap_words = 0;
        if (!cuda_ok(cudaMalloc(&g->d_bbitmaps, static_cast<size_t>(bm_stride * kMaxBatch + 1024) * 4), "cudaMalloc batch bitmaps"))
            return false;
        g->bbitmap_cap_words = bm_stride * kMaxBatch;
    }
    if (!ensure_desc(g, n_chunks)) return false;
    {
        GpuEngine::TimingSlot *slot = &g->ring[g->ring_head];
        g->ring_head = (g->ring_head + 1) % GpuEngine::kTimingRing;
        if (slot->pending && g->accumulate_timing) resolve_slot(g, slot, nullptr);
        slot->pending = false;
        if (g->cur_slot && g->cur_slot->pending && !g->accumulate_timing) g->cur_slot->pending = false;
        g->cur_slot = slot;
        g->ev0 = slot->ev[0];
        g->ev_mid = slot->ev[1];
        g->ev1 = slot->ev[2];
        g->ev_post = slot->ev[3];
    }
    uint32_t *bitmaps[kMaxBatch] = {nullptr};
    for (int q = 0; q < nq; ++q) bitmaps[q] = g->d_bbitmaps + static_cast<size_t>(q) * bm_stride;
    ScanLaunch L{};
    L.table = &t;
    L.d_ctl = g->d_ctl;
    L.h_prog = &uni;
    bool ok = cuda_ok(cudaMemcpyAsync(g->d_bprogs, g->h_bprogs, sizeof(Program) * nq, cudaMemcpyHostToDevice, g->stream),
                      "upload batch programs") &&
              cuda_ok(cudaMemsetAsync(g->d_bcounts, 0, sizeof(unsigned long long) * kMaxBatch, g->stream), "reset batch counts") &&
              cuda_ok(cudaMemsetAsync(g->d_bctl, 0, sizeof(QueryCtl) * kMaxBatch, g->stream), "reset batch control");
    if (!ok) return false;
    cudaEventRecord(g->ev0, g->stream);
    if (t.n > 0 && !cuda_ok(batch_launch(L, geo, nq, g->d_bprogs, bitmaps, g->d_bcounts, g->stream), "batch scan kernel launch"))
        return false;
    cudaEventRecord(g->ev_mid, g->stream);
    if (!cuda_ok(cudaMemcpyAsync(g->h_bcounts, g->d_bcounts, sizeof(unsigned long long) * nq, cudaMemcpyDeviceToHost, g->stream),
                 "download batch counts") ||
        !cuda_ok(cudaStreamSynchronize(g->stream), "batch scan sync"))
        return false;
    uint64_t total = 0;
    for (int q = 0; q < nq; ++q) {
        offsets[q] = total;
        total += g->h_bcounts[q];
    }
    offsets[nq] = total;
    if (!engine_ensure_ids(g, static_cast<int64_t>(total))) return false;
    int launches = t.n > 0 ? 1 : 0;
    for (int q = 0; q < nq; ++q) {
        const uint64_t m = g->h_bcounts[q];
        if (m == 0 || t.n == 0) continue;
        if (!cuda_ok(compact_launch(bitmaps[q], bm_words, &g->d_bctl[q], g->d_tile_desc, next_epoch(g), g->d_ids + offsets[q],
                                    0u, m, g->stream),
                     "compaction kernel launch"))
            return false;
        ++launches;
    }
    cudaEventRecord(g->ev1, g->stream);
    if (!cuda_ok(cudaStreamSynchronize(g->stream), "batch compaction sync")) return false;
    ScanStats st;
    st.path = 0;
    st.rows_scanned = t.n;
    st.matches = static_cast<int64_t>(total);
    st.algo_bytes = t.n * bytes_per_row + 4 * static_cast<int64_t>(total);
    st.launches = launches;
    st.tile_rows = geo.tile_rows;
    st.stages = geo.stages;
    st.grid = geo.grid;
    st.total_ms = now_ms() - t_begin;
    g->last = st;
    g->cur_slot->pending = true;
    g->cur_slot->staged = true;
    g->cur_slot->two_kernels = true;
    g->cur_slot->has_post = false;
    return true;
}

// ------------------------------------------------------------------------------------------
// K1c on its own: the second half of a split scan (count first, then compaction to a caller
// chosen destination -- possibly another GPU's buffer, so the ordered gather of a sharded
// table is done by the compaction kernel's own stores over NVLink)
// ------------------------------------------------------------------------------------------
bool engine_compact_to(GpuEngine *g, uint32_t *dst, uint32_t id_base, uint64_t cap) {
    cudaSetDevice(g->device);
    engine_resolve_timing(g);  // this call re-records two of the slot's events
    if (g->last_bm_words <= 0 && g->table.n > 0) {
        set_error("no match bitmap to compact: run a full-scan match with a bitmap first");
        return false;
    }
    if (g->table.n == 0 || g->last_bm_count == 0) {
        g->last.compact_ms = 0;
        return true;
    }
    const int64_t n_chunks = compact_chunks(g->last_bm_words);
    if (!ensure_desc(g, n_chunks)) return false;
    // the chunk claim counter lives in the control block: reset it for this launch
    if (!cuda_ok(cudaMemsetAsync(&g->d_ctl->chunk_counter, 0, sizeof(unsigned int), g->stream), "reset chunk counter"))
        return false;
    cudaEventRecord(g->ev_mid, g->stream);
    if (!cuda_ok(compact_launch(g->d_bitmap, g->last_bm_words, g->d_ctl, g->d_tile_desc, next_epoch(g), dst, id_base,
                                cap, g->stream),
                 "compaction kernel launch"))
        return false;
    cudaEventRecord(g->ev1, g->stream);
    if (!cuda_ok(cudaStreamSynchronize(g->stream), "compaction sync")) return false;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g->ev_mid, g->ev1);
    g->last.compact_ms = ms;
    g->last.kernel_ms = g->last.scan_ms + ms;
    g->last.launches += 1;
    g->last.algo_bytes += 4 * static_cast<int64_t>(g->last_bm_count);
    return true;
}

// ------------------------------------------------------------------------------------------
// index path, one segment at a time, with the keys of the survivors: what a SHARDED table needs.
// The global result of segment S(w, i) is ordered (key ASC, GLOBAL position DESC); a shard can
// only produce (key ASC, local position DESC).  Handing out (key, id) per segment lets the caller
// merge the shards: concatenate them from the highest rank to the lowest and sort stably by key
// (sharding.py).  Single-GPU queries never come here.
// ------------------------------------------------------------------------------------------
bool engine_match_segments(GpuEngine *g, const struct whereClauseS *wc, std::vector<SegmentResult> *out,
                           bool *used_index) {
    cudaSetDevice(g->device);
    out->clear();
    const DevTable &t = g->table;
    uint32_t widths[NUM_COLS];
    for (int c = 0; c < NUM_COLS; ++c) widths[c] = t.col[c].width;
    QueryCtl *hc = g->h_ctl;
    const std::string err = compile_where(wc, widths, &hc->prog, false);
    if (!err.empty()) {
        set_error(err);
        return false;
    }
    for (int c = 0; c < NUM_COLS; ++c)
        if ((hc->prog.col_mask & (1u << c)) && !t.resident(c)) {
            set_error(std::string("WHERE references column '") + kCols[c].name + "' which is not resident on the device");
            return false;
        }
    SegmentPlan segs[kMaxSegments];
    bool too_many = false;
    const int n_seg = plan_segments(g, wc, segs, kMaxSegments, &too_many);
    if (n_seg < 0) return false;
    if (too_many) {
        set_error("too many (condition x index) segments in one WHERE clause");
        return false;
    }
    *used_index = n_seg > 0;
    if (n_seg == 0) return true;
    int launches = 0;
    for (int s = 0; s < n_seg; ++s) {
        DevIndex &ix = g->idx[segs[s].index_slot];
        if (!ensure_index(g, &ix, &launches)) return false;
        // probe this segment
        unsigned long long lo = 0, hi = 0;
        if (segs[s].is_u64) {
            lo = segs[s].lo_u64;
            hi = segs[s].hi_u64;
        } else {
            std::memcpy(&lo, &segs[s].lo_i32, 4);
            std::memcpy(&hi, &segs[s].hi_i32, 4);
        }
        g->h_probe_keys[0] = lo;
        g->h_probe_keys[kMaxSegments] = hi;
        if (!cuda_ok(cudaMemcpyAsync(g->d_probe_lo, g->h_probe_keys, 8, cudaMemcpyHostToDevice, g->stream), "probe h2d") ||
            !cuda_ok(cudaMemcpyAsync(g->d_probe_hi, g->h_probe_keys + kMaxSegments, 8, cudaMemcpyHostToDevice, g->stream),
                     "probe h2d") ||
            !cuda_ok(index_probe(ix, g->d_probe_lo, g->d_probe_hi, 1, g->d_probe_first, g->d_probe_count, g->stream),
                     "probe kernel launch") ||
            !cuda_ok(cudaMemcpyAsync(g->h_probe_out, g->d_probe_first, 4, cudaMemcpyDeviceToHost, g->stream), "probe d2h") ||
            !cuda_ok(cudaMemcpyAsync(g->h_probe_out + kMaxSegments, g->d_probe_count, 4, cudaMemcpyDeviceToHost, g->stream),
                     "probe d2h") ||
            !cuda_ok(cudaStreamSynchronize(g->stream), "probe sync"))
            return false;
        CandSegments cs{};
        cs.n_seg = 1;
        cs.perm[0] = ix.perm;
        cs.first[0] = g->h_probe_out[0];
        cs.vstart[0] = 0;
        cs.vstart[1] = g->h_probe_out[kMaxSegments];
        const long long n_cand = cs.vstart[1];
        SegmentResult r;
        r.key_col = ix.col;
        if (n_cand > 0) {
            hc->tile_counter = 0;
            hc->chunk_counter = 0;
            hc->out_count = 0;
            if (!cuda_ok(cudaMemcpyAsync(g->d_ctl, hc, sizeof(QueryCtl), cudaMemcpyHostToDevice, g->stream), "upload query"))
                return false;
            if (!ensure_desc(g, filter_tiles(n_cand)) || !engine_ensure_ids(g, n_cand)) return false;
            if (!cuda_ok(filter_launch(t, g->d_ctl, cs, g->d_tile_desc, next_epoch(g), g->d_ids, g->stream),
                         "filter kernel launch") ||
                !cuda_ok(cudaMemcpyAsync(&hc->out_count, &g->d_ctl->out_count, 8, cudaMemcpyDeviceToHost, g->stream),
                         "download count") ||
                !cuda_ok(cudaStreamSynchronize(g->stream), "filter sync"))
                return false;
            const int64_t m = static_cast<int64_t>(hc->out_count);
            r.ids.resize(static_cast<size_t>(m));
            if (m > 0) {
                if (!cuda_ok(cudaMemcpy(r.ids.data(), g->d_ids, static_cast<size_t>(m) * 4, cudaMemcpyDeviceToHost),
                             "download ids"))
                    return false;
                std::vector<uint8_t> raw;
                if (!engine_fetch_rows(g, ix.col, g->d_ids, m, &raw)) return false;
                r.keys.resize(static_cast<size_t>(m));
                for (int64_t k = 0; k < m; ++k) {
                    if (ix.type == T_U64) {
                        unsigned long long v;
                        std::memcpy(&v, raw.data() + 8 * k, 8);
                        r.keys[k] = static_cast<long long>(v);
                    } else {
                        int v;
                        std::memcpy(&v, raw.data() + 4 * k, 4);
                        r.keys[k] = v;
                    }
                }
            }
        }
        out->push_back(std::move(r));
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// downloads
// ------------------------------------------------------------------------------------------
bool engine_fetch_rows(GpuEngine *g, int col, const uint32_t *d_ids, int64_t n, std::vector<uint8_t> *out) {
    cudaSetDevice(g->device);
    const DevColumn &c = g->table.col[col];
    out->assign(static_cast<size_t>(n) * c.width, 0);
    if (n == 0) return true;
    if (!c.d) {
        set_error(std::string("column '") + kCols[col].name + "' is not resident on the device");
        return false;
    }
    uint8_t *d_tmp = nullptr;
    if (!cuda_ok(cudaMalloc(&d_tmp, static_cast<size_t>(n) * c.width), "cudaMalloc gather")) return false;
    bool ok = cuda_ok(gather_launch(c.d, c.width, d_ids, n, d_tmp, g->stream), "gather kernel launch");
    ok = ok && cuda_ok(cudaMemcpyAsync(out->data(), d_tmp, out->size(), cudaMemcpyDeviceToHost, g->stream),
                       "download gather");
    ok = cuda_ok(cudaStreamSynchronize(g->stream), "gather sync") && ok;
    cudaFree(d_tmp);
    return ok;
}

bool engine_download_all(GpuEngine *g, HostColumns *out) {
    cudaSetDevice(g->device);
    out->n = g->table.n;
    for (int c = 0; c < NUM_COLS; ++c) {
        const DevColumn &dc = g->table.col[c];
        out->width[c] = dc.width;
        out->data[c].assign(static_cast<size_t>(g->table.n) * dc.width, 0);
        if (g->table.n == 0) continue;
        if (!dc.d) {
            set_error(std::string("column '") + kCols[c].name + "' is not resident on the device");
            return false;
        }
        if (!cuda_ok(cudaMemcpyAsync(out->data[c].data(), dc.d, out->data[c].size(), cudaMemcpyDeviceToHost, g->stream),
                     "download column"))
            return false;
    }
    return cuda_ok(cudaStreamSynchronize(g->stream), "download sync");
}

// ------------------------------------------------------------------------------------------
// DELETE: stable compaction of every column by the keep list (K1 with the inverted program)
// ------------------------------------------------------------------------------------------
static bool ensure_dml_scratch(GpuEngine *g, size_t bytes) {
    if (bytes <= g->dml_scratch_bytes && g->d_dml_scratch) return true;
    if (g->d_dml_scratch) cudaFree(g->d_dml_scratch);
    g->d_dml_scratch = nullptr;
    g->dml_scratch_bytes = 0;
    const size_t cap = bytes + bytes / 8 + 4096;
    if (!cuda_ok(cudaMalloc(&g->d_dml_scratch, cap), "cudaMalloc DML scratch")) return false;
    g->dml_scratch_bytes = cap;
    return true;
}

bool engine_delete(GpuEngine *g, const struct whereClauseS *wc, int64_t *deleted) {
    cudaSetDevice(g->device);
    uint64_t kept = 0;
    // executeQueryDeleteSerial never uses an index (:646-677): force the scan path; invert = keep list
    if (!engine_match(g, wc, true, true, false, false, &kept)) return false;
    const int64_t n_old = g->table.n;
    *deleted = n_old - static_cast<int64_t>(kept);
    if (*deleted == 0) return true;
    // K5: stable compaction of every column by the keep list (ids ascending: ids[k] >= k), IN PLACE, a block of output
    // rows at a time through a scratch buffer -- the block's source rows all lie at or behind the block's first row,
    // where nothing has been overwritten yet.  No allocation, no per-column synchronisation.
    constexpr int64_t kBlockRows = int64_t(4) << 20;
    uint32_t max_w = 1;
    for (int c = 0; c < NUM_COLS; ++c)
        if (g->table.col[c].d && g->table.col[c].width > max_w) max_w = g->table.col[c].width;
    // indexes that are up to date are maintained (entries of deleted rows dropped, the others renumbered) instead of
    // being re-sorted: remap[old row] = new row
    bool any_index = false;
    size_t idx_scratch = 0;
    for (auto &ix : g->idx)
        if (ix.usable && !ix.dirty) {
            any_index = true;
            const size_t b = index_scratch_bytes(ix, ix.n);
            if (b > idx_scratch) idx_scratch = b;
        }
    const size_t col_scratch =
        ((static_cast<size_t>(kBlockRows < static_cast<int64_t>(kept) ? kBlockRows : static_cast<int64_t>(kept)) * max_w + 255) & ~size_t(255)) + 256;
    const size_t remap_bytes = any_index ? ((static_cast<size_t>(n_old) * 4 + 255) & ~size_t(255)) : 0;
    const size_t need = remap_bytes + (idx_scratch > col_scratch ? idx_scratch : col_scratch) + 512;
    if (!ensure_dml_scratch(g, need)) return false;
    uint8_t *scratch = static_cast<uint8_t *>(g->d_dml_scratch);
    uint32_t *d_remap = reinterpret_cast<uint32_t *>(scratch);
    uint8_t *work = scratch + remap_bytes;
    unsigned int *d_counter = reinterpret_cast<unsigned int *>(work + (need - remap_bytes - 256));   // last 256 bytes
    unsigned long long *d_total = reinterpret_cast<unsigned long long *>(d_counter + 2);
    if (any_index) {
        if (!cuda_ok(index_build_remap(g->d_ids, static_cast<long long>(kept), n_old, d_remap, g->stream), "remap kernel launch"))
            return false;
        for (auto &ix : g->idx) {
            if (!ix.usable || ix.dirty) continue;
            if (!ensure_desc(g, index_filter_tiles(ix.n))) return false;
            int launches = 0;
            if (!cuda_ok(index_apply_delete(&ix, d_remap, static_cast<long long>(kept), work, g->d_tile_desc, d_counter, d_total,
                                            next_epoch(g), g->stream, &launches),
                         "index maintenance (DELETE)"))
                return false;
        }
    }
    for (int c = 0; c < NUM_COLS; ++c) {
        DevColumn &dc = g->table.col[c];
        if (!dc.d) continue;
        for (int64_t r0 = 0; r0 < static_cast<int64_t>(kept); r0 += kBlockRows) {
            const int64_t r1 = (r0 + kBlockRows < static_cast<int64_t>(kept)) ? r0 + kBlockRows : static_cast<int64_t>(kept);
            if (!cuda_ok(gather_launch(dc.d, dc.width, g->d_ids + r0, r1 - r0, work, g->stream), "compaction kernel launch") ||
                !cuda_ok(cudaMemcpyAsync(dc.d + static_cast<size_t>(r0) * dc.width, work, static_cast<size_t>(r1 - r0) * dc.width,
                                         cudaMemcpyDeviceToDevice, g->stream),
                         "compaction copy"))
                return false;
        }
        // rows behind the new end stay zero (a tile may be read past the last row)
        const int64_t tail = n_old - static_cast<int64_t>(kept);
        if (!cuda_ok(cudaMemsetAsync(dc.d + static_cast<size_t>(kept) * dc.width, 0, static_cast<size_t>(tail) * dc.width, g->stream),
                     "compaction tail"))
            return false;
    }
    if (!cuda_ok(cudaStreamSynchronize(g->stream), "compaction sync")) return false;
    g->table.n = static_cast<int64_t>(kept);
    g->head.num_records = static_cast<int>(kept);
    return true;
}

// ------------------------------------------------------------------------------------------
// INSERT: append one row (widening string columns when the new text needs it)
// ------------------------------------------------------------------------------------------
static bool grow_column(GpuEngine *g, int c, uint32_t new_width, int64_t new_cap) {
    DevColumn &dc = g->table.col[c];
    DevColumn nc;
    if (!column_alloc(&nc, new_width, new_cap, g->stream)) return false;
    if (g->table.n > 0) {
        if (new_width == dc.width) {
            if (!cuda_ok(cudaMemcpyAsync(nc.d, dc.d, static_cast<size_t>(g->table.n) * dc.width,
                                         cudaMemcpyDeviceToDevice, g->stream),
                         "grow column"))
                return false;
        } else if (!cuda_ok(restride_launch(dc.d, dc.width, nc.d, new_width, g->table.n, g->stream),
                            "restride kernel launch"))
            return false;
    }
    if (!cuda_ok(cudaStreamSynchronize(g->stream), "grow sync")) return false;
    cudaFree(dc.d);
    dc = nc;
    return true;
}

bool engine_append(GpuEngine *g, const record &r) {
    cudaSetDevice(g->device);
    const uint8_t *rb = reinterpret_cast<const uint8_t *>(&r);
    const int64_t n = g->table.n;
    // the new row's cells, one after the other in a pinned staging block; ONE synchronisation for the whole row
    constexpr size_t kCellSlot = 512 + 16;
    if (!g->h_row_stage && !cuda_ok(cudaMallocHost(&g->h_row_stage, kCellSlot * NUM_COLS), "cudaMallocHost row stage")) return false;
    std::memset(g->h_row_stage, 0, kCellSlot * NUM_COLS);
    for (int c = 0; c < NUM_COLS; ++c) {
        DevColumn &dc = g->table.col[c];
        if (!dc.d) {
            set_error("INSERT needs every column resident on the device");
            return false;
        }
        uint32_t need_w = dc.width;
        size_t len = 0;
        if (kCols[c].type == T_STR) {
            len = strnlen(reinterpret_cast<const char *>(rb + kCols[c].rec_offset), kCols[c].field_bytes - 1);
            const uint32_t w = round_up16(static_cast<uint32_t>(len) + 1);
            if (w > need_w) need_w = w;
        }
        // keep one full tile of slack behind the last row (bulk copies read whole tiles)
        const bool need_cap = n + 1 + kRowPad > dc.cap;
        if (need_w != dc.width || need_cap) {
            if (!grow_column(g, c, need_w, need_cap ? cap_for(n + 1) : dc.cap)) return false;
        }
        uint8_t *cell = g->h_row_stage + kCellSlot * c;
        if (kCols[c].type == T_STR)
            std::memcpy(cell, rb + kCols[c].rec_offset, len);
        else if (kCols[c].type == T_BOOL)
            cell[0] = r.sudo_used ? 1 : 0;
        else
            std::memcpy(cell, rb + kCols[c].rec_offset, dc.width);
        if (!cuda_ok(cudaMemcpyAsync(g->table.col[c].d + static_cast<size_t>(n) * g->table.col[c].width, cell,
                                     g->table.col[c].width, cudaMemcpyHostToDevice, g->stream),
                     "append cell"))
            return false;
    }
    g->table.n = n + 1;
    g->head.num_records = static_cast<int>(n + 1);
    // insert() of the reference (engine/bplus.c:723-740), per index: the new row enters at the FRONT of its key run.
    // One K3 probe + one shift of the tail, no re-sort (an index that is not built yet stays "dirty" and is built on use).
    size_t idx_scratch = 0;
    for (auto &ix : g->idx)
        if (ix.usable && !ix.dirty) {
            const size_t b = index_scratch_bytes(ix, ix.n);
            if (b > idx_scratch) idx_scratch = b;
        }
    bool ok = true;
    if (idx_scratch && !ensure_dml_scratch(g, idx_scratch + 256)) ok = false;
    for (auto &ix : g->idx) {
        if (!ok) break;
        int launches = 0;
        ok = cuda_ok(index_insert_row(&ix, g->table, n, g->d_dml_scratch, g->d_probe_first, g->d_probe_count, g->stream, &launches),
                     "index maintenance (INSERT)");
    }
    return cuda_ok(cudaStreamSynchronize(g->stream), "append sync") && ok;
}

}  // namespace qpe
