/*
 * qpe_gpu.h -- low-level C-ABI of libqpegpu.so (new; nothing like it exists in the reference).
 *
 * executeEngine-gpu.h is the drop-in surface (resultSetS of strings, as the reference's bridge
 * expects).  This header exposes the same hot path WITHOUT string materialisation, for
 * benchmarks, parity tests and host languages that want row ids: the match phase of
 * executeQuerySelectSerial (engine/serial/executeEngine-serial.c:355-476), the DELETE match
 * mask (:646-677), and batched findRange / find_rows lookups (engine/bplus.c:282-314, :361-411).
 *
 * Conventions: every function returns 0 on success and a negative value on failure, with a
 * message retrievable through qpe_gpu_last_error().  Buffers returned through an out-pointer
 * are malloc-owned by the caller and released with qpe_gpu_free().  Row ids are positions in
 * table order (index into the reference's all_records[]), local to the engine's shard.
 */
#ifndef QPE_GPU_H
#define QPE_GPU_H

#include "qpe_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qpe_scan_stats {
    double kernel_ms;       /* device time of the match phase (CUDA events on the engine's stream) */
    double total_ms;        /* host wall time of the call */
    long long rows_scanned; /* rows the match phase evaluated (N on the scan path, candidates on the index path) */
    long long candidates;   /* index path: sum of segment lengths; 0 on the scan path */
    long long matches;
    long long algo_bytes;   /* N * sum(width of distinct WHERE columns) + 4 * matches   */
    int path;               /* 0 = full scan (K1), 1 = index (K3 + K1g) */
    int launches;           /* kernels launched by this call */
    int tile_rows;
    int stages;
    int grid;
    int reserved;
    double scan_ms;         /* K1 (scan_tma_kernel) alone; 0 on the index path */
    double compact_ms;      /* K1c (compact_kernel) alone; 0 when not launched */
} qpe_scan_stats;

/* 1 when a CUDA device is usable by this process, else 0 (then every engine call fails loudly). */
int qpe_gpu_available(void);
const char *qpe_gpu_last_error(void);
void qpe_gpu_free(void *p);

/* Engine over rows already in host memory (array-of-structs), instead of a CSV file.
 * datafile may be NULL (then INSERT/DELETE have no file side effect). */
struct engineS *qpe_gpu_engine_from_records(const record *rows, long long n_rows, int num_indexes,
                                            const char *indexed_attributes[], const int attribute_types[],
                                            const char *datafile, const char *tableName);

/* Engine over a synthetic table generated ON THE DEVICE (distributions of the reference's
 * data-generation/generate_commands.py, restated for a counter-based generator; see
 * csrc/synth.cuh).  Rows [row_base, row_base + n_rows) of the virtual table of `total_rows`
 * rows; column_mask selects which columns are materialised (bit c = schema column c). */
struct engineS *qpe_gpu_engine_synth(unsigned long long total_rows, unsigned long long row_base,
                                     unsigned long long n_rows, unsigned long long seed,
                                     unsigned int column_mask, int num_indexes,
                                     const char *indexed_attributes[], const int attribute_types[]);

long long qpe_gpu_num_rows(const struct engineS *engine);
/* Global row id of this engine's row 0 (non-zero for a shard of a row-range sharded table). */
unsigned long long qpe_gpu_row_base(const struct engineS *engine);

/* Match phase of SELECT with the reference's path rule (index path iff a top-level condition
 * names a u64/int index; else full scan).  *ids_out receives the matching row ids in the
 * reference's result order. */
int qpe_gpu_select_ids(struct engineS *engine, struct whereClauseS *whereClause, unsigned int **ids_out,
                       size_t *n_out, qpe_scan_stats *stats);

/* Query batch (the GPU analogue of QPEOMP's query-level parallelism, QPEOMP.c:234-335): the match phase of
 * n_queries SELECTs, result by result identical to n_queries calls of qpe_gpu_select_ids.  Queries on the
 * index path run one by one; full-scan queries that reference the SAME set of columns are evaluated together --
 * up to 8 WHERE programs per pass, so those columns are read once per 8 queries.  ids_out[q] is
 * malloc'ed (qpe_gpu_free), n_out[q] its length.  stats (optional) = sums over the passes. */
int qpe_gpu_select_ids_batch(struct engineS *engine, struct whereClauseS *const *whereClauses, int n_queries,
                             unsigned int **ids_out, size_t *n_out, qpe_scan_stats *stats);
int qpe_sql_select_ids_batch(struct engineS *engine, const char *const *statements, int n_queries,
                             unsigned int **ids_out, size_t *n_out, qpe_scan_stats *stats);

/* Same match phase, result left in HBM (device pointer valid until the next call on this
 * engine).  flags: bit0 = force the full-scan path even if an index applies;
 * bit1 = count only (no ids written); bit2 = global row ids (see QPE_SCAN_GLOBAL_IDS). */
#define QPE_SCAN_FORCE 1
#define QPE_SCAN_COUNT_ONLY 2
#define QPE_SCAN_GLOBAL_IDS 4 /* sharded table: add the shard's first global row to every id */
int qpe_gpu_select_ids_device(struct engineS *engine, struct whereClauseS *whereClause, int flags,
                              unsigned long long *count_out, const unsigned int **d_ids_out,
                              qpe_scan_stats *stats);

/* Same as qpe_gpu_select_ids, into a caller-provided (ideally pinned) host buffer of `cap` ids.
 * *n_out always receives the match count; returns -5 when it exceeds cap. flags: QPE_SCAN_FORCE. */
int qpe_gpu_select_ids_into(struct engineS *engine, struct whereClauseS *whereClause, int flags, unsigned int *ids,
                            size_t cap, size_t *n_out, qpe_scan_stats *stats);

/* Split full scan for sharded tables.  qpe_gpu_scan_count runs K1 only (match bitmap + count,
 * always the full-scan path); qpe_gpu_compact_to then runs K1c on that bitmap and writes the row
 * ids to dst_device, which may be memory of ANOTHER GPU mapped with qpe_gpu_ipc_open: the
 * ordered gather of a sharded table is then the compaction kernel's own coalesced stores over
 * NVLink (each rank writes at the exclusive prefix of the lower ranks' counts).  global_ids != 0
 * adds the shard's first global row to every id. */
int qpe_gpu_scan_count(struct engineS *engine, struct whereClauseS *whereClause, unsigned long long *count_out,
                       qpe_scan_stats *stats);
int qpe_gpu_compact_to(struct engineS *engine, unsigned int *dst_device, int global_ids, qpe_scan_stats *stats);
/* K1 + K1c in one call, ids written to dst_device (this GPU's memory or a peer mapping) which holds
 * dst_capacity ids.  *count_out always receives the match count; if it exceeds dst_capacity the call
 * returns -5 and nothing was stored at or beyond the capacity. */
int qpe_gpu_select_ids_to(struct engineS *engine, struct whereClauseS *whereClause, unsigned int *dst_device,
                          unsigned long long dst_capacity, int global_ids, unsigned long long *count_out,
                          qpe_scan_stats *stats);
int qpe_sql_select_ids_to(struct engineS *engine, const char *statement, unsigned int *dst_device,
                          unsigned long long dst_capacity, int global_ids, unsigned long long *count_out,
                          qpe_scan_stats *stats);
int qpe_sql_scan_count(struct engineS *engine, const char *statement, unsigned long long *count_out,
                       qpe_scan_stats *stats);

/* Index path of a SHARDED table.  For every (top-level condition x index) segment of the WHERE, in the
 * reference's generation order (executeEngine-serial.c:358-459), the rows of THIS shard that pass the
 * whole WHERE, in (key ASC, local position DESC) order, with their ORDER keys (u64 keys with the top bit
 * flipped, so that a signed 64-bit sort orders them as the B+ tree's unsigned compare does; int keys
 * sign-extended).  The caller merges the shards per segment: concatenate from the highest
 * rank to the lowest and sort stably by key = (key ASC, global position DESC), the reference's order.
 * *used_index_out = 0 when no index applies (use the scan path).  keys_out / ids_out are malloc'ed
 * (qpe_gpu_free), concatenated over the segments; seg_counts_out[s] = rows of segment s (<= 32). */
int qpe_gpu_select_segments(struct engineS *engine, struct whereClauseS *whereClause, int global_ids,
                            int *used_index_out, int *n_segments_out, size_t seg_counts_out[32], long long **keys_out,
                            unsigned int **ids_out);
int qpe_sql_select_segments(struct engineS *engine, const char *statement, int global_ids, int *used_index_out,
                            int *n_segments_out, size_t seg_counts_out[32], long long **keys_out,
                            unsigned int **ids_out);

/* Raw device buffers shareable between the processes (one per GPU) of one box through CUDA IPC. */
void *qpe_gpu_device_alloc(size_t bytes);
void qpe_gpu_device_free(void *p);
int qpe_gpu_ipc_export(void *device_ptr, unsigned char handle_out[64]);
void *qpe_gpu_ipc_open(const unsigned char handle[64]);
void qpe_gpu_ipc_close(void *mapped_ptr);
int qpe_gpu_copy_to_device(void *dst_device, const void *src_host, size_t bytes);
int qpe_gpu_copy_to_host(void *dst_host, const void *src_device, size_t bytes);
int qpe_gpu_copy_device(void *dst_device, const void *src_device, size_t bytes);

/* ---- Row-range sharded table, one process per GPU of one box (csrc/shard.cu) ----------------------
 * What the reference's MPI mode does with MPI_Allreduce / MPI_Allgather(v) (engine/mpi/executeEngine-mpi.c:
 * 703-770), done by the kernels themselves over NVLink peer memory: every rank's scan stores its (global)
 * row ids into its segment of the owner's result buffer (device result) or into its own segment (host result),
 * a kernel stores (epoch, count) into every rank's comm block, and the owner packs the segments in partition
 * order = table order / every rank takes its 1/world of the result to the host.  No host collective and no
 * NCCL call per query; the caller only passes the 64-byte IPC handles around once.
 *   qpe_shard_init(engine, rank, world, cap, handle_out) allocate this rank's comm block + its own id segments
 *                                                        (2 x cap ids), export the allocation
 *   qpe_shard_connect(engine, all_handles)               world x 64 bytes in rank order
 *   qpe_shard_set_device_result(engine, owner, segments, cap)  segments = qpe_shard_result_ids(world, cap) ids
 *                                                        in the OWNER's memory (own pointer / qpe_gpu_ipc_open
 *                                                        mapping); the first shard stores straight into the
 *                                                        dense result (its offset is always 0)
 *   qpe_shard_open_host_result(engine, name, cap, create) POSIX shared-memory id buffer (3 x cap ids) all ranks
 *                                                        write their 1/world of a result into over their own
 *                                                        PCIe link; this rank's parts go to its GPU's NUMA node
 *   qpe_shard_pin_host_result(engine)                    after every rank has opened it: register it with CUDA
 *   qpe_shard_submit / qpe_shard_wait                    one full-scan SELECT, enqueued / completed; every rank
 *                                                        calls them with the same statement in the same order;
 *                                                        at most two queries in flight
 *   qpe_shard_select / qpe_sql_shard_select              submit + wait */
int qpe_shard_init(struct engineS *engine, int rank, int world, unsigned long long segment_capacity,
                   unsigned char comm_handle_out[64]);
int qpe_shard_connect(struct engineS *engine, const unsigned char *all_handles);
unsigned long long qpe_shard_result_ids(int world, unsigned long long segment_capacity);
int qpe_shard_set_device_result(struct engineS *engine, int owner_rank, unsigned int *segments,
                                unsigned long long segment_capacity);
unsigned int *qpe_shard_open_host_result(struct engineS *engine, const char *name, unsigned long long capacity,
                                         int create);
int qpe_shard_pin_host_result(struct engineS *engine);
/* NUMA node of this rank's GPU (-1 unknown); *how_out: bit 0 = mbind accepted, bit 1 = first touch under that node's CPUs */
int qpe_shard_numa(struct engineS *engine, int *how_out);
/* diagnostics: host time qpe_shard_wait has spent on host-result queries since the last reset (ms, summed over *n_out
 * queries): [0] waiting for the counts, [1] this rank's device->host copy, [2] (owner) the other ranks' pieces */
int qpe_shard_wait_breakdown(struct engineS *engine, double out_ms[3], long long *n_out, int reset);
/* creator, after every rank has opened the buffer: drop its /dev/shm name (mappings stay valid) */
int qpe_shard_unlink_host_result(struct engineS *engine);
/* the packed ids of the most recent device-result / host-result query that qpe_shard_wait completed.  A host result
 * stays valid until the third qpe_shard_submit after its own. */
const unsigned int *qpe_shard_device_result(struct engineS *engine);
const unsigned int *qpe_shard_host_result(struct engineS *engine);
/* the same for the most recent (back = 0) or the previous (back = 1) host-result query, with its number of ids; and
 * deferred completion: with `on`, qpe_shard_wait returns once the counts are known and THIS rank's device->host copy
 * is queued -- it waits neither for that copy nor (owner) for the other ranks' pieces, so the next query is submitted
 * at once; qpe_shard_host_result / _at wait for every piece.  The shared buffer holds three id arrays (epoch % 3), so a
 * host result stays valid until the third qpe_shard_submit after its own: a consumer takes result q after submitting
 * q + 2.  Every rank of a group should choose the same. */
const unsigned int *qpe_shard_host_result_at(struct engineS *engine, int back, unsigned long long *total_out);
int qpe_shard_set_deferred(struct engineS *engine, int on);
/* Host result path: 1 (default) = every rank reads its 1/world of the result out of the ranks' segments over NVLink
 * into its own HBM and the copy engine takes it to the host over this GPU's PCIe link; 2 = the delivery kernel
 * stores into the mapped host buffer itself.  Every rank must choose the same. */
int qpe_shard_set_multipath(struct engineS *engine, int mode);
/* shares of a host result per rank, proportional to `weights` (measured link rates); same numbers on every rank */
int qpe_shard_set_link_weights(struct engineS *engine, const double *weights, int n);
void qpe_shard_close(struct engineS *engine);
int qpe_shard_submit(struct engineS *engine, struct whereClauseS *whereClause, int to_host);
int qpe_shard_wait(struct engineS *engine, unsigned long long *counts_out, qpe_scan_stats *stats);
int qpe_shard_select(struct engineS *engine, struct whereClauseS *whereClause, int to_host,
                     unsigned long long *counts_out, qpe_scan_stats *stats);
int qpe_sql_shard_submit(struct engineS *engine, const char *statement, int to_host);
int qpe_sql_shard_select(struct engineS *engine, const char *statement, int to_host,
                         unsigned long long *counts_out, qpe_scan_stats *stats);
/* DELETE on the sharded table: every rank deletes its shard's matches, the new shard sizes are all-gathered
 * through the comm blocks and the shards are renumbered so that global row ids stay positions in the whole
 * table (engine/mpi/executeEngine-mpi.c:703-770).  Outputs: rows deleted / rows left, over all shards. */
int qpe_shard_delete(struct engineS *engine, struct whereClauseS *whereClause, unsigned long long *deleted_total_out,
                     unsigned long long *rows_total_out);
int qpe_sql_shard_delete(struct engineS *engine, const char *statement, unsigned long long *deleted_total_out,
                         unsigned long long *rows_total_out);

/* cudaMemcpy device -> host for pointers handed out by the *_device calls. 0 on success. */
int qpe_gpu_copy_from_device(void *dst_host, const void *src_device, size_t bytes);

/* Statistics of the engine's most recent match phase. */
int qpe_gpu_last_stats(struct engineS *engine, qpe_scan_stats *stats);

/* Device timing is taken with CUDA events on the engine's stream around every match phase and resolved
 * lazily (only when a qpe_scan_stats is asked for).  qpe_gpu_set_timing(engine, 1) additionally sums the
 * times of EVERY call since then (up to 512 calls are held unresolved, older ones are folded in as their slot
 * is reused); qpe_gpu_timing_totals returns {match kernels, scan kernel, compaction kernel, post-scan kernel}
 * ms and the number of calls, so a benchmark can time a loop without paying for event queries inside it. */
int qpe_gpu_set_timing(struct engineS *engine, int accumulate);
int qpe_gpu_timing_totals(struct engineS *engine, double totals_ms[4], long long *calls_out);

/* Host-side breakdown of the most recent match phase, in ms: [0] WHERE compile, [1] enqueue (copies +
 * launches), [2] stream synchronisation, [3] device time of the post-scan kernel of a sharded SELECT,
 * [4] tail of the match phase after the synchronisation, [5] whole qpe_sql_shard_select call, [6..7] 0. */
int qpe_gpu_last_trace(struct engineS *engine, double out[8]);

/* Diagnostics (only with QPE_FUSE_TRACE=1 in the environment): per-CTA %globaltimer stamps of the most recent
 * fused scan, 8 words per CTA: [0] CTA start, [1] its first tile landed, [2] evaluators done, [3] CTA end,
 * [5] chunks it processed.  Returns the number of CTAs written (0 when tracing is off). */
int qpe_gpu_fused_trace(struct engineS *engine, unsigned long long *out, int max_ctas);

/* Write the whole table as a CSV in the data generator's format (header line, QUOTE_MINIMAL
 * quoting, \r\n line ends, sudo_used as true/false: data-generation/generate_commands.py:812-816)
 * so that the reference's loader and ours can both ingest it.  Every column must be resident. */
int qpe_gpu_write_csv(struct engineS *engine, const char *path);

/* DELETE's match mask without deleting: bit r of the bitmap = row r matches.  bitmap must hold
 * (num_rows + 31) / 32 words. */
int qpe_gpu_match_mask(struct engineS *engine, struct whereClauseS *whereClause, unsigned int *bitmap,
                       size_t n_words, unsigned long long *count_out, qpe_scan_stats *stats);

/* Batched probes with PLAIN key arrays (unsigned long long for a u64 index, int for an int index) -- the end-to-end
 * form of findRange / find_rows (engine/bplus.c:282-314, :361-411) for many probes at once.  hi == NULL: point
 * probes (hi = lo).  lo / hi and first / count may be host pointers (ideally from qpe_gpu_host_alloc: pinned memory is
 * copied in place, chunk by chunk, beside the probe kernel; pageable memory goes through a bounce buffer) or device
 * pointers (no copies).  flags: QPE_PROBE_SORT = sort the batch by its lower keys on the device first, probe in key
 * order and scatter the answers back.  stats->kernel_ms is the device time of the whole batch, copies included. */
#define QPE_PROBE_SORT 1
int qpe_gpu_probe_keys(struct engineS *engine, const char *attribute, const void *lo, const void *hi, size_t n_queries,
                       unsigned int *first, unsigned int *count, int flags, qpe_scan_stats *stats);
/* K4 on its own: the stable LSD radix sort behind the index build (the reference builds the same (key ASC, position
 * DESC) leaf order by inserting row after row: engine/serial/buildEngine-serial.c:46-53 -> engine/bplus.c:723-740) and
 * behind QPE_PROBE_SORT.  keys: n keys of key_bytes (4 or 8) in host memory, ordered as unsigned or (signed_keys) two's
 * complement values.  mode 0: pairs (keys[i], vals[i]); 1: (keys[i], i); 2: the input read backwards, (keys[n-1-i],
 * n-1-i) -- equal keys then come out in DESCENDING position order.  passes_out: digit passes run (byte positions equal
 * in all keys are skipped); kernel_ms_out: device time of the sort. */
int qpe_gpu_sort_pairs(const void *keys, const unsigned int *vals, size_t n, int key_bytes, int signed_keys, int mode,
                       void *keys_out, unsigned int *vals_out, int *passes_out, double *kernel_ms_out);
/* pinned host memory for probe batches and id lists (release with qpe_gpu_host_free) */
void *qpe_gpu_host_alloc(size_t bytes);
void qpe_gpu_host_free(void *p);

/* Batched inclusive-range probes [lo[q], hi[q]] on the index of `attribute` (point lookup:
 * lo == hi).  first[q] / count[q] delimit the answer inside the index order
 * (key ascending, table position descending == the reference's leaf chain). */
int qpe_gpu_probe_batch(struct engineS *engine, const char *attribute, const KEY_T *lo, const KEY_T *hi,
                        size_t n_queries, unsigned int *first, unsigned int *count, qpe_scan_stats *stats);
/* Row ids of index entries [first, first + count) of `attribute`'s index. */
int qpe_gpu_index_slice(struct engineS *engine, const char *attribute, unsigned int first, unsigned int count,
                        unsigned int *row_ids_out);
/* Keys of the same index entries (u64 keys as long long bits, int keys sign-extended). */
int qpe_gpu_index_slice_keys(struct engineS *engine, const char *attribute, unsigned int first, unsigned int count,
                             long long *keys_out);

/* Copy n_rows values of a column to the host in device layout (u64 / int32 / u8 / fixed-width
 * NUL-padded text); *width_out = bytes per row.  out must hold n_rows * width bytes; pass
 * out == NULL to query the width only. */
int qpe_gpu_fetch_column(struct engineS *engine, const char *attribute, long long first_row, long long n_rows,
                         void *out, unsigned int *width_out);

/* Bytes that go host -> device per full-scan query: the scan kernel's parameter block, which carries the
 * compiled WHERE program (no separate upload). */
unsigned int qpe_gpu_query_upload_bytes(void);

/* The CUDA stream (cudaStream_t) every kernel of this engine is launched on, so a caller can bracket
 * calls with its own CUDA events for device-side timing. */
void *qpe_gpu_stream(struct engineS *engine);

/* Tuning override for the scan kernel (0 = automatic). */
int qpe_gpu_set_tile(struct engineS *engine, int tile_rows, int stages);
/* Pipelined full scan: the table is cut into `segments` pieces of whole 64 Ki-row chunks and K1c of
 * piece i runs beside K1 of piece i+1 (two streams), so the ordered ids leave the GPU during the
 * scan.  0 = automatic (one piece per 32 Mi rows, at most 8), 1 = never pipeline, up to 16. */
int qpe_gpu_set_pipeline(struct engineS *engine, int segments);

/* SQL front end (same grammar and quirks as the reference's tokenizer/src/tokenizer.c +
 * connectEngine.c bridge).  qpe_sql_run executes one statement and writes what the
 * reference's run_test_query prints (connectEngine.c:125-245) to `out` (stdout if NULL). */
void qpe_sql_run(struct engineS *engine, const char *statement, int max_rows, void *out_FILE);
/* Same, into a malloc'ed NUL-terminated buffer. */
char *qpe_sql_run_to_text(struct engineS *engine, const char *statement, int max_rows);
/* Parse `statement`, run only the match phase of its WHERE (SELECT or DELETE text). */
int qpe_sql_select_ids(struct engineS *engine, const char *statement, int flags, unsigned int **ids_out,
                       size_t *n_out, qpe_scan_stats *stats);
int qpe_sql_select_ids_device(struct engineS *engine, const char *statement, int flags,
                              unsigned long long *count_out, const unsigned int **d_ids_out,
                              qpe_scan_stats *stats);
int qpe_sql_select_ids_into(struct engineS *engine, const char *statement, int flags, unsigned int *ids, size_t cap,
                            size_t *n_out, qpe_scan_stats *stats);
int qpe_sql_match_mask(struct engineS *engine, const char *statement, unsigned int *bitmap, size_t n_words,
                       unsigned long long *count_out, qpe_scan_stats *stats);
/* Full SELECT through executeQuerySelectGPU; NULL if `statement` is not a SELECT.  Release with
 * freeResultSet (executeEngine-gpu.h). */
struct resultSetS *qpe_sql_select(struct engineS *engine, const char *statement);
/* The predicate program the engine compiles from `statement`'s WHERE (struct Program of csrc/qpe_internal.h, raw
 * bytes) for the given cell widths of the 12 schema columns.  Needs no device: the CPU tests interpret it row by row
 * against the oracle.  Returns the bytes written, or -7 (parse), -2 (compile), -5 (cap too small). */
long long qpe_sql_compile_program(const char *statement, const unsigned int widths[12], void *out, size_t cap);
/* The whereClauseS list the front end builds for `statement`, rendered as text (malloc'ed). */
char *qpe_sql_where_to_text(const char *statement);

#ifdef __cplusplus
}
#endif

#endif /* QPE_GPU_H */
