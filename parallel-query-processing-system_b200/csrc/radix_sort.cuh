// radix_sort.cuh -- K4: stable LSD radix sort of (key, 32-bit payload) pairs (see radix_sort.cu)
#pragma once

#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace qpe {

// How the n input pairs are read.
enum SortInput {
    kSortPairs = 0,        // element i = (keys_in[i], vals_in[i])
    kSortIota = 1,         // element i = (keys_in[i], i)                      (vals_in ignored)
    kSortReverseIota = 2,  // element i = (keys_in[n-1-i], n-1-i): the table read backwards -- a stable sort of it puts
                           // equal keys in DESCENDING position order, the leaf-chain order of the reference's tree
};

// bytes of device scratch a sort of n pairs with KEY_BYTES-byte keys needs (4 or 8)
size_t radix_sort_scratch_bytes(long long n, int key_bytes);

// Sorts by key ascending (signed_keys: as two's-complement values), stable.  keys_in is only read; keys_out / vals_out
// (n entries each) receive the result; they must not overlap the input or the scratch.  Byte positions in which all
// keys agree cost nothing: one reduction pass finds them (this needs ONE stream synchronisation at the start).
// K = unsigned long long or uint32_t.
template <typename K>
cudaError_t radix_sort_pairs(const K *keys_in, const uint32_t *vals_in, SortInput mode, bool signed_keys, K *keys_out,
                             uint32_t *vals_out, long long n, void *scratch, size_t scratch_bytes, cudaStream_t stream,
                             int *launches, int *passes_out = nullptr);

}  // namespace qpe
