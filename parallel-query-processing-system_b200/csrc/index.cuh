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

}  // namespace qpe
