// scan_kernels.cuh -- launch interface of the SELECT/WHERE kernels (sm_100a)
#pragma once

#include <cuda_runtime.h>
#include "qpe_internal.h"

namespace qpe {

constexpr int kMaxProgressSegments = 16;

// Device-resident control block of one query: the compiled predicate plus the words the
// kernels need zeroed at launch.  Uploaded with ONE cudaMemcpyAsync per query.
struct QueryCtl {
    Program prog;
    unsigned int tile_counter;      // dynamic tile claim (K1, K1g)
    unsigned int chunk_counter;     // ordered chunk claim (K1c)
    unsigned long long out_count;   // total matches (written by the kernel)
    unsigned int seg_stored[kMaxProgressSegments];  // (unused since K1f has its own control block)
};

// K1f keeps its few control words in a block of its own that it RESETS ITSELF (the last CTA to finish hands the
// count over and zeroes the rest), and takes the compiled program as a kernel parameter: a fused scan needs no
// host->device copy before its launch, so nothing sits between two queries' kernels but the launch itself.
struct FusedCtl {
    unsigned long long out_count;    // running sum of the CTAs' counts; zero between queries
    unsigned long long final_count;  // the last query's match count (read by the sharded post-scan kernel)
    unsigned int ctas_done;          // zero between queries
    unsigned int next_chunk;         // tickets handed out by the dynamic chunk assignment (ticket t = chunk grid + t); zero between queries
    unsigned int seg_stored[kMaxProgressSegments];  // chunks of table segment s whose ids are stored; zero between queries
};

// K1f (fused scan + compaction): chunk = consecutive tiles, at most this many rows; the kernel keeps
// two chunk bitmaps and one id stage in shared memory beside the TMA stages
constexpr int kFuseMaxChunkRows = 65536;
// shared memory K1f sets aside beside the TMA stages: two chunk bitmaps + the id stage of its compaction warps
constexpr size_t fuse_reserve_bytes(int compact_warps) {
    return 2 * (kFuseMaxChunkRows / 32) * 4 + static_cast<size_t>(32 * 32 * compact_warps) * 4;
}

struct ScanLaunch {
    const DevTable *table;
    const QueryCtl *d_ctl;          // device copy (prog already uploaded)
    const Program *h_prog;          // host copy (for col_mask / sizing)
    uint32_t *out_bitmap;           // device, may be null (count only); one bit per row, tile-padded
    int force_tile_rows;            // 0 = choose
    int force_stages;               // 0 = choose
    long long tile_begin;           // tiles [tile_begin, tile_end) of the table; tile_end 0 = to the last tile
    long long tile_end;
};

struct ScanGeometry {
    int tile_rows;
    int stages;
    int grid;
    size_t smem_bytes;
    int64_t n_tiles;
    int64_t bytes_per_row;
    int chunk_tiles;      // K1f only
    int64_t n_chunks;     // K1f only
    int compact_warps;    // K1f only: 4 or 8
};

// K1: TMA-staged predicate evaluation over the whole table -> match bitmap + count (ctl->out_count).
// Returns false (and sets *why) if the query cannot be staged (row too wide for shared memory):
// the caller then uses the gather path below with an identity candidate list.
// max_stages (1..4) caps the pipeline depth: 3 leaves room for a K1c CTA beside K1's on every SM.
// fused_cw != 0: plan for K1f (larger header, fuse_reserve_bytes(cw) of shared memory set aside; geo->chunk_tiles and
// geo->n_chunks are filled in).
// fused_cw: 0 = plan for K1 / K9, 4 or 8 = plan for K1f with that many compaction warps
bool scan_plan(const DevTable &t, const Program &prog, int force_tile_rows, int force_stages, int max_stages,
               ScanGeometry *geo, const char **why, int fused_cw = 0, size_t extra_reserve = 0);

// K9: up to kMaxBatch programs over one pass of the union of their columns -> one bitmap + count per query
constexpr int kMaxBatch = 8;
size_t batch_smem_bytes(int n_prog);
cudaError_t batch_launch(const ScanLaunch &L, const ScanGeometry &geo, int n_prog, const Program *d_progs,
                         uint32_t *const *d_bitmaps, unsigned long long *d_counts, cudaStream_t stream);

// K1f: scan + ordered compaction in one launch.  desc needs geo.n_chunks descriptors; progress (may be
// null) points to mapped pinned host memory with one word per table segment of seg_chunks chunks.
struct FusedLaunch {
    ScanLaunch scan;
    unsigned long long *desc;
    uint32_t epoch;
    uint32_t id_base;
    uint32_t *out_ids;
    unsigned long long out_cap;
    long long seg_chunks;
    unsigned long long *progress;
    unsigned long long *host_count;  // mapped pinned host word (device alias) that receives the match count, or null
    FusedCtl *d_fctl;                // device, zeroed once at engine creation
    unsigned long long *trace;       // diagnostics: 8 words per CTA (device), or null
    bool pdl;                        // programmatic dependent launch: may overlap the previous kernel of the stream
    bool l2_stream;                  // fetch the table with the L2 evict_first policy (what the query writes then stays in L2)
};
size_t fused_param_bytes();          // bytes of K1f's kernel parameter block (carries the compiled program)
cudaError_t fused_launch(const FusedLaunch &L, const ScanGeometry &geo, cudaStream_t stream);
cudaError_t scan_launch(const ScanLaunch &L, const ScanGeometry &geo, cudaStream_t stream);

// K1c: order-preserving compaction of a match bitmap (bit b of word w = row 32 w + b) into row
// ids, single pass with decoupled look-back over 64 Ki-row chunks.  desc needs
// compact_chunks(n_words) descriptors; epoch is stamped into them (never 0).  The match count is
// NOT produced here (K1 already accumulated it into ctl->out_count).
int64_t compact_chunks(long long n_words);
// out_ids may point into ANOTHER GPU's memory (peer mapping): the ids then travel over NVLink as
// the coalesced stores of the compaction itself; id_base is added to every id (shard -> table).
// launch_chunks > 0: launch only the next launch_chunks chunks (pipelined scan; see compact_launch).
constexpr int kCompactChunkRows = 65536;
cudaError_t compact_launch(const uint32_t *bitmap, long long n_words, const QueryCtl *d_ctl, unsigned long long *desc,
                           uint32_t epoch, uint32_t *out_ids, uint32_t id_base, unsigned long long out_cap,
                           cudaStream_t stream, long long launch_chunks = 0);

// K1g: evaluate the predicate on a list of candidate rows (concatenated index segments, or the
// identity list when perm == nullptr) and compact the survivors in list order.
constexpr int kMaxSegments = 32;
struct CandSegments {
    int n_seg;
    const uint32_t *perm[kMaxSegments];  // nullptr => identity
    long long first[kMaxSegments];       // first entry of the slice within perm
    long long vstart[kMaxSegments + 1];  // prefix of slice lengths (virtual candidate index)
};
// d_segs != nullptr: the segment table is read from device memory (written by index_probe_segments just before, no
// host round trip); max_candidates then bounds the grid, and segs is ignored.  Nothing is stored at or beyond
// out_cap (the count in ctl->out_count is still exact).
cudaError_t filter_launch(const DevTable &t, const QueryCtl *d_ctl, const CandSegments &segs,
                          unsigned long long *tile_desc, uint32_t epoch, uint32_t *out_ids, cudaStream_t stream,
                          const CandSegments *d_segs = nullptr, long long max_candidates = 0,
                          unsigned long long out_cap = ~0ull);
int64_t filter_tiles(long long n_candidates);

// K2: projection gather  out[k] = column[ids[k]]  (width bytes per row)
cudaError_t gather_launch(const uint8_t *col, uint32_t width, const uint32_t *ids, int64_t n_ids,
                          uint8_t *out, cudaStream_t stream);

// K7 (format.cu): projection of one column for a list of row ids, rendered as fixed-width NUL-terminated text
// slots (numeric columns: 24 / 16 / 8 bytes; text columns: the cell itself, i.e. K2).  out holds n slots.
uint32_t format_slot_width(int col_type, uint32_t cell_width);
cudaError_t format_launch(const uint8_t *col, int col_type, uint32_t cell_width, const uint32_t *ids, int64_t n,
                          uint8_t *out, cudaStream_t stream);

// K8 (format.cu): the whole table as CSV text in the format of the reference's DELETE rewrite / INSERT append
// (executeEngine-serial.c:562-575, :687-700).  csv_measure: d_offs[i] = byte offset of row i, d_offs[n] = total
// (d_offs holds n + 1 entries; scan_tmp == nullptr sizes the scan scratch); csv_write: the text of rows [r0, r1)
// into d_out, whose first byte is file offset base_off = d_offs[r0].
cudaError_t csv_measure(const DevTable &t, unsigned long long *d_offs, void *scan_tmp, size_t *scan_tmp_bytes,
                        cudaStream_t stream);
cudaError_t csv_write(const DevTable &t, const unsigned long long *d_offs, long long r0, long long r1,
                      unsigned long long base_off, char *d_out, cudaStream_t stream);
// same, but ids come from a bitmap-free "keep list" and output goes to a new column (DELETE)
// -- identical kernel; alias kept for readability at call sites.

// widen / re-stride a fixed-width column (INSERT of a longer string): dst width >= src width
cudaError_t restride_launch(const uint8_t *src, uint32_t src_w, uint8_t *dst, uint32_t dst_w, int64_t n,
                            cudaStream_t stream);

// ids[k] += base (sharded tables: local -> global row id), 64-bit output optional
cudaError_t add_base_launch(uint32_t *ids, int64_t n, uint32_t base, cudaStream_t stream);

}  // namespace qpe
