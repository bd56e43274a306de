// index.cuh -- flattened index build / probe interface (see index.cu)
#pragma once

#include <cuda_runtime.h>
#include "qpe_internal.h"

namespace qpe {

// (re)build ix from table column ix->col; synchronises the stream; clears ix->dirty
cudaError_t index_build(DevIndex *ix, const DevTable &t, cudaStream_t stream, int *launches);
void index_free(DevIndex *ix);

// batched inclusive-range probe: for q in [0,Q): first[q] = lower_bound(lo[q]),
// count[q] = upper_bound(hi[q]) - first[q] (0 when lo > hi).  Keys are u64 or int per ix.type;
// all pointers are device pointers.
cudaError_t index_probe(const DevIndex &ix, const void *d_lo, const void *d_hi, long long q, uint32_t *d_first,
                        uint32_t *d_count, cudaStream_t stream);

// K3s: all probes of one indexed SELECT in ONE launch (one warp per segment; ix[s] / lo[s] / hi[s] per segment, int
// keys in the low 32 bits), leaving the candidate-segment table for K1g in device memory (*d_out)
struct CandSegments;
cudaError_t index_probe_segments(const DevIndex *const *ix, const unsigned long long *lo, const unsigned long long *hi,
                                 int n_seg, CandSegments *d_out, cudaStream_t stream);

// ---- maintenance without a re-sort (no allocation; `scratch` holds index_scratch_bytes(ix, ix.n) bytes) ----
size_t index_scratch_bytes(const DevIndex &ix, long long n_entries);
// INSERT: table row `row` (the table's last row) enters the index at the front of its key run.  d_first / d_count:
// two device words of scratch.  A dirty or unusable index is left alone; one without head-room is marked dirty.
cudaError_t index_insert_row(DevIndex *ix, const DevTable &t, long long row, void *scratch, uint32_t *d_first,
                             uint32_t *d_count, cudaStream_t stream, int *launches);
// DELETE: remap[old row] = new row / 0xffffffff (index_build_remap from the keep list); the entries of deleted rows
// are dropped, the others renumbered, order kept.  desc: index_filter_tiles(ix.n) look-back descriptors (epoch-tagged,
// never 0), d_counter / d_total: device scratch words.
cudaError_t index_build_remap(const uint32_t *d_keep, long long n_keep, long long n_old, uint32_t *d_remap,
                              cudaStream_t stream);
long long index_filter_tiles(long long n);
cudaError_t index_apply_delete(DevIndex *ix, const uint32_t *d_remap, long long n_new, void *scratch,
                               unsigned long long *desc, unsigned int *d_counter, unsigned long long *d_total,
                               uint32_t epoch, cudaStream_t stream, int *launches);

}  // namespace qpe
