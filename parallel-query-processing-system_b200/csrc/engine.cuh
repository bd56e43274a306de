// engine.cuh -- the engine object behind `struct engineS *` and its internal operations
#pragma once

#include <cuda_runtime.h>

#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "qpe_internal.h"
#include "scan_kernels.cuh"
#include "index.cuh"
#include "qpe_gpu.h"

namespace qpe {

constexpr int kMaxPipeSegments = 16;
constexpr uint64_t kEngineMagic = 0x5150454750553031ull;  // "QPEGPU01"
constexpr uint64_t kResultMagic = 0x5150455245533031ull;  // "QPERES01"

struct GpuEngine {
    struct engineS head;  // MUST stay first: callers hold `struct engineS *`
    uint64_t magic = kEngineMagic;
    std::mutex mu;        // serialises the calls on THIS engine (EngineLock)
    int device = 0;
    cudaStream_t stream = nullptr;   // K1 and everything else (highest priority)
    cudaStream_t stream2 = nullptr;  // K1c of a pipelined scan (lowest priority: fills the SMs beside K1)
    // Events of the CURRENT match phase: handles borrowed from `ring` (one slot per call).  Elapsed times are
    // resolved lazily -- when statistics are asked for (qpe_scan_stats / qpe_gpu_last_stats), or, with
    // accumulate_timing, when the slot comes round again / qpe_gpu_timing_totals is called -- because four
    // cudaEventElapsedTime calls cost ~11 us of host time per query, which a sharded 125 M-row scan notices.
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;
    struct TimingSlot {
        cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // ev0, ev_mid, ev1, ev_post
        bool pending = false, staged = false, two_kernels = false, has_post = false;
    };
    static constexpr int kTimingRing = 512;
    TimingSlot ring[kTimingRing];
    int ring_head = 0;
    TimingSlot *cur_slot = nullptr;      // slot of the most recent match phase (pending = not resolved yet)
    bool accumulate_timing = false;
    double acc_kernel_ms = 0, acc_scan_ms = 0, acc_compact_ms = 0, acc_post_ms = 0;
    long long acc_calls = 0;
    cudaEvent_t ev_seg[kMaxPipeSegments] = {nullptr};  // K1 of segment i done -> K1c of segment i may start
    int pipe_segments = 0;           // 0 = fused K1f (default), 1 = K1 then K1c, n >= 2 = n pipelined table segments
    // K1f progress words (mapped pinned host memory): the kernel publishes "ids of table segment s are
    // complete in HBM" so the host can copy that segment out while the scan is still running
    unsigned long long *h_progress = nullptr;
    unsigned long long *d_progress = nullptr;  // device alias of h_progress
    // host destination for the ids of the next full-scan match (qpe_gpu_select_ids_into): copied segment
    // by segment on stream2 during the scan
    uint32_t *host_out = nullptr;
    uint64_t host_out_cap = 0;
    bool host_out_done = false;      // the match phase delivered the ids to host_out

    DevTable table;
    std::vector<DevIndex> idx;  // parallel to head.indexed_attributes
    int idx_slots = 0;          // capacity of the head's index arrays

    // per-query scratch
    FusedCtl *d_fctl = nullptr;  // K1f's self-resetting control words
    unsigned long long *d_trace = nullptr;  // QPE_FUSE_TRACE=1: per-CTA time stamps of the last K1f (diagnostics)
    const unsigned long long *count_dev = nullptr;  // device word holding the match count of the scan just enqueued
    QueryCtl *d_ctl = nullptr;
    QueryCtl *h_ctl = nullptr;  // pinned
    unsigned long long *d_tile_desc = nullptr;
    int64_t desc_cap = 0;
    uint32_t epoch = 0;
    uint32_t *d_ids = nullptr;
    int64_t ids_cap = 0;
    uint32_t *d_bitmap = nullptr;
    int64_t bitmap_cap_words = 0;
    // K9 query batch: programs (pinned host + device), per-query control blocks (K1c chunk counters), counts,
    // and one bitmap per query in a single allocation
    Program *h_bprogs = nullptr, *d_bprogs = nullptr;
    QueryCtl *d_bctl = nullptr;
    unsigned long long *d_bcounts = nullptr, *h_bcounts = nullptr;
    uint32_t *d_bbitmaps = nullptr;
    int64_t bbitmap_cap_words = 0;
    uint8_t *d_fmt = nullptr;        // K7 scratch: rendered projection slots of one row chunk (grow-only)
    size_t fmt_cap = 0;
    // destination override for the next full-scan match (qpe_gpu_select_ids_to): ids go to out_override
    // (this GPU's or a peer's memory, out_override_cap ids at most) with id_base_override added
    uint32_t *out_override = nullptr;
    uint64_t out_override_cap = 0;
    uint32_t id_base_override = 0;
    bool id_base_always = false;     // add id_base_override even when the ids go to d_ids (sharded host result)
    // sharded table (shard.cu): kernels enqueued on `stream` right after the match kernels, before the
    // single stream synchronisation of engine_match (count exchange + pack over peer memory)
    std::function<bool()> post_match;
    const unsigned long long *count_mapped = nullptr;  // set: post_match's kernel writes the match count here (mapped pinned)
    void *shard = nullptr;           // ShardState of shard.cu
    int64_t last_bm_words = 0;   // words of the bitmap the last full-scan match left in d_bitmap (0 = none)
    uint64_t last_bm_count = 0;  // its match count
    // probe scratch (device + pinned host), kMaxSegments entries each
    unsigned long long *d_probe_lo = nullptr, *d_probe_hi = nullptr;
    uint32_t *d_probe_first = nullptr, *d_probe_count = nullptr;
    unsigned long long *h_probe_keys = nullptr;  // pinned: lo[kMaxSegments], hi[kMaxSegments]
    uint32_t *h_probe_out = nullptr;             // pinned: first[kMaxSegments], count[kMaxSegments]
    CandSegments *d_segs = nullptr;              // the candidate-segment table K3s leaves for K1g (index path)
    // batched probes (probe_batch.cu): grow-only device scratch and pinned bounce buffer
    void *d_probe_scratch = nullptr;
    size_t probe_scratch_bytes = 0;
    void *h_probe_bounce = nullptr;
    size_t probe_bounce_bytes = 0;
    // DML (engine_delete / engine_append): grow-only device scratch (column block / index tail / remap), pinned row stage
    void *d_dml_scratch = nullptr;
    size_t dml_scratch_bytes = 0;
    uint8_t *h_row_stage = nullptr;

    // host-side breakdown of the last match phase (ms): [0] parse/compile, [1] enqueue (copies + launches),
    // [2] stream synchronisation, [3] device time ev1 -> end of the post-match kernels, [4] tail of the
    // match phase after the synchronisation, [5] whole C-ABI call of a sharded SELECT (parse included)
    double trace[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev_post = nullptr;
    // K1f compaction warps for the next full scan: 4, or 8 once a scan matched more than 1/8 of its rows
    // (the evaluators then wait for the compaction warps); QPE_FUSE_CW=4|8 pins it
    int fuse_cw = 4;
    bool fuse_l2 = false;             // next full scan streams the table through L2 with the evict_first policy (see below)
    int force_tile_rows = 0, force_stages = 0;
    ScanStats last;
};

extern std::mutex g_api_mutex;  // engine creation / destruction, and calls on a handle that is not an engine
// An engine is not re-entrant (one stream, one scratch set): calls on ONE engine are serialised by its own mutex,
// calls on different engines of a process run side by side (a QPEOMP-style driver with one engine per thread).
struct EngineLock {
    std::unique_lock<std::mutex> lk;
    explicit EngineLock(const struct engineS *e) {
        const GpuEngine *g = reinterpret_cast<const GpuEngine *>(e);
        if (g && g->magic == kEngineMagic)
            lk = std::unique_lock<std::mutex>(const_cast<GpuEngine *>(g)->mu);
        else
            lk = std::unique_lock<std::mutex>(g_api_mutex);
    }
};
GpuEngine *as_engine(struct engineS *e);  // nullptr (and error set) if e is not one of ours
void set_error(const std::string &msg);
double now_ms();
bool cuda_ok(cudaError_t e, const char *what);

// create an engine with an empty table; nullptr if no device
GpuEngine *engine_create(const char *tableName, const char *datafile, int index_slots);
void engine_destroy(GpuEngine *g);
int engine_live_count();  // engines alive in this process
void shard_destroy(GpuEngine *g);  // shard.cu: releases the peer mappings / shared host buffer, if any
int device_count();
const char *last_error_cstr();
bool column_alloc(DevColumn *col, uint32_t width, int64_t cap_rows, cudaStream_t stream);
// DELETE (stable compaction of every column) and INSERT (append one row); indexes go dirty
bool engine_delete(GpuEngine *g, const struct whereClauseS *wc, int64_t *deleted);
bool engine_append(GpuEngine *g, const record &r);
bool engine_upload(GpuEngine *g, const HostColumns &hc);  // replaces the table
bool engine_add_index(GpuEngine *g, const char *name, int attributeType);
bool engine_ensure_ids(GpuEngine *g, int64_t n);
// bring a (usable) index up to date with the table before it is probed
bool index_ready(GpuEngine *g, DevIndex *ix, int *launches);
// resolve the event times of the most recent match phase into g->last (no-op when already done)
void engine_resolve_timing(GpuEngine *g);
// resolve every pending slot into the accumulators (accumulate_timing)
void engine_resolve_all(GpuEngine *g);

// match phase. On success the ids are in g->d_ids[0 .. *count) (unless count_only).
bool engine_match(GpuEngine *g, const struct whereClauseS *wc, bool force_scan, bool invert, bool count_only,
                  bool want_bitmap, uint64_t *count);

// K1f enqueued WITHOUT a synchronisation (sharded SELECT, shard.cu): compile + plan + launch, the count stays in
// g->d_fctl->final_count.  engine_fused_finish does the bookkeeping once the caller knows the count.
struct FusedEnqueue {
    ScanGeometry geo;
    int64_t bytes_per_row = 0;
    GpuEngine::TimingSlot *slot = nullptr;  // its ev[3] is free for the caller's post-scan kernels
    bool launched = false;                   // false: empty shard (count 0), nothing launched
};
// overlap_previous: the scan may start while the stream's previous kernel is still running (it must not depend on it)
void engine_adapt_fused(GpuEngine *g, int64_t matches, int64_t rows);
bool engine_fused_enqueue(GpuEngine *g, const struct whereClauseS *wc, uint32_t *out_ids, uint64_t out_cap,
                          uint32_t id_base, FusedEnqueue *fe, bool overlap_previous, bool l2_stream = false);
void engine_fused_finish(GpuEngine *g, const FusedEnqueue &fe, uint64_t matches, int extra_launches, double t_begin_ms);

// K9: the match phase of up to kMaxBatch full-scan queries in ONE pass over the columns.  On success the ids of
// query q are g->d_ids[offsets[q] .. offsets[q + 1]) in table order.  Every query must be a full-scan query
// (engine_uses_index(...) == false); returns false with an error if the union of columns cannot be staged.
bool engine_match_batch(GpuEngine *g, const struct whereClauseS *const *wcs, int nq, uint64_t *offsets);
// true when the reference's path rule sends this WHERE down the index path (a top-level condition on a u64/int index)
bool engine_uses_index(GpuEngine *g, const struct whereClauseS *wc);
// bit c set = the WHERE references schema column c (0 and an error if it does not compile)
bool engine_where_columns(GpuEngine *g, const struct whereClauseS *wc, uint32_t *mask);

// K1c alone: compact the bitmap left by the last count-only full-scan match into `dst` (device
// memory of this GPU or a peer mapping), adding id_base to every id
bool engine_compact_to(GpuEngine *g, uint32_t *dst, uint32_t id_base, uint64_t cap);

// index path per segment with keys (sharded tables; see engine.cu)
struct SegmentResult {
    int key_col = -1;
    std::vector<long long> keys;  // key of every surviving row (u64 keys reinterpreted; int keys sign-extended)
    std::vector<uint32_t> ids;    // local row ids, (key ASC, local position DESC)
};
bool engine_match_segments(GpuEngine *g, const struct whereClauseS *wc, std::vector<SegmentResult> *out,
                           bool *used_index);

// download helpers
bool engine_fetch_rows(GpuEngine *g, int col, const uint32_t *d_ids, int64_t n, std::vector<uint8_t> *out);
bool engine_download_all(GpuEngine *g, HostColumns *out);

int64_t load_csv_columns(const char *path, HostColumns *out);
// K6: parse the CSV on the device straight into columns. 1 = done, 0 = use the host loader, -1 = CUDA error
int ingest_csv_gpu(GpuEngine *g, const char *path, int *launches_out);
void parse_csv_chunk(const char *chunk, size_t len, record *r);

}  // namespace qpe
