// probe_batch.cu -- batched B+-tree probes end to end: plain key arrays in, (first, count) out
//
// The reference answers one probe at a time: findRange (engine/bplus.c:282-314) descends from the root (findLeaf,
// :317-358) and walks the leaf chain; find_rows (:361-411) is its point-lookup sibling.  Here a BATCH of probes is
// answered by K3 (csrc/index.cu) over the flattened index.  This file is the host side that keeps the PCIe link and
// the kernel busy together:
//
//   qpe_gpu_probe_keys   keys as plain u64 / int arrays (no 16-byte KEY_T unions to convert), engine-owned device
//                        scratch (no cudaMalloc per call), the batch cut into chunks that alternate between the
//                        engine's two streams -- H2D of chunk i + 1, K3 of chunk i and D2H of chunk i - 1 overlap.
//                        Caller buffers from qpe_gpu_host_alloc (pinned) are copied in place; pageable ones go through
//                        a pinned bounce buffer.  Device pointers are taken as they are (no copies at all).
//   QPE_PROBE_SORT       the batch is sorted by its lower keys on the device first (radix sort of (key, slot) pairs),
//                        probed in key order -- neighbouring probes then walk the same separator nodes, and the leaf
//                        level is touched front to back like a merge -- and the answers are scattered back to the
//                        caller's order.

#include <chrono>
#include <cstring>

#include "engine.cuh"
#include "radix_sort.cuh"
#include "executeEngine-gpu.h"

namespace qpe {

template <typename K>
__global__ void gather_keys_kernel(const K *__restrict__ src, const uint32_t *__restrict__ slot, K *__restrict__ dst,
                                   long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        dst[i] = src[slot[i]];
}
__global__ void scatter_answers_kernel(const uint32_t *__restrict__ first_s, const uint32_t *__restrict__ count_s,
                                       const uint32_t *__restrict__ slot, uint32_t *__restrict__ first,
                                       uint32_t *__restrict__ count, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint32_t s = slot[i];
        first[s] = first_s[i];
        count[s] = count_s[i];
    }
}

static int grid_for(long long work, int threads) {
    long long g = (work + threads - 1) / threads;
    if (g > 148ll * 8) g = 148ll * 8;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

static bool ensure_probe_scratch(GpuEngine *g, size_t bytes) {
    if (bytes <= g->probe_scratch_bytes && g->d_probe_scratch) return true;
    if (g->d_probe_scratch) cudaFree(g->d_probe_scratch);
    g->d_probe_scratch = nullptr;
    g->probe_scratch_bytes = 0;
    const size_t cap = bytes + bytes / 4 + 4096;
    if (!cuda_ok(cudaMalloc(&g->d_probe_scratch, cap), "cudaMalloc probe scratch")) return false;
    g->probe_scratch_bytes = cap;
    return true;
}
static bool ensure_probe_bounce(GpuEngine *g, size_t bytes) {
    if (bytes <= g->probe_bounce_bytes && g->h_probe_bounce) return true;
    if (g->h_probe_bounce) cudaFreeHost(g->h_probe_bounce);
    g->h_probe_bounce = nullptr;
    g->probe_bounce_bytes = 0;
    const size_t cap = bytes + bytes / 4 + 4096;
    if (!cuda_ok(cudaHostAlloc(&g->h_probe_bounce, cap, cudaHostAllocDefault), "cudaHostAlloc probe bounce")) return false;
    g->probe_bounce_bytes = cap;
    return true;
}

enum PtrKind { kPageable, kPinned, kDevice };
static PtrKind kind_of(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return kPageable;
    }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return kDevice;
    if (a.type == cudaMemoryTypeHost) return kPinned;
    return kPageable;
}

}  // namespace qpe

using namespace qpe;

extern "C" {

void *qpe_gpu_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (!cuda_ok(cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable), "cudaHostAlloc")) return nullptr;
    return p;
}
void qpe_gpu_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int qpe_gpu_probe_keys(struct engineS *engine, const char *attribute, const void *lo, const void *hi, size_t n_queries,
                       unsigned int *first, unsigned int *count, int flags, qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    cudaSetDevice(g->device);
    const int slot = isAttributeIndexed(engine, attribute);
    if (slot < 0 || !g->idx[slot].usable) {
        set_error("attribute has no probe-able (u64 / int) index");
        return -6;
    }
    if (!lo || !first || !count) {
        set_error("qpe_gpu_probe_keys: lo, first and count are required");
        return -5;
    }
    if (n_queries > 0xffffffffull) {
        set_error("qpe_gpu_probe_keys: at most 2^32 - 1 probes per batch");
        return -5;
    }
    DevIndex &ix = g->idx[slot];
    int launches = 0;
    if (!index_ready(g, &ix, &launches)) return -2;
    const long long Q = static_cast<long long>(n_queries);
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (Q == 0) return 0;
    const size_t ksz = ix.type == T_U64 ? 8 : 4;
    const bool sorted = (flags & QPE_PROBE_SORT) != 0;
    const bool point = hi == nullptr;
    const PtrKind k_lo = kind_of(lo), k_hi = point ? k_lo : kind_of(hi), k_first = kind_of(first), k_count = kind_of(count);
    const bool dev_in = k_lo == kDevice, dev_out = k_first == kDevice;
    if ((k_hi == kDevice) != dev_in || (k_count == kDevice) != dev_out) {
        set_error("qpe_gpu_probe_keys: lo / hi (and first / count) must both be host or both be device pointers");
        return -5;
    }
    const auto t0 = std::chrono::steady_clock::now();
    // device scratch: lo | hi | first | count | (sort) keys_sorted | hi_sorted | slot | slot_sorted | first_s | count_s | tmp
    const size_t kq = (static_cast<size_t>(Q) * ksz + 255) & ~size_t(255);
    const size_t wq = (static_cast<size_t>(Q) * 4 + 255) & ~size_t(255);
    const size_t tmp_bytes = sorted ? radix_sort_scratch_bytes(Q, static_cast<int>(ksz)) : 0;
    const size_t need = 2 * kq + 2 * wq + (sorted ? 2 * kq + 4 * wq + tmp_bytes + 256 : 0);
    if (!ensure_probe_scratch(g, need)) return -4;
    uint8_t *sp = static_cast<uint8_t *>(g->d_probe_scratch);
    void *d_lo = sp;
    void *d_hi = sp + kq;
    uint32_t *d_first = reinterpret_cast<uint32_t *>(sp + 2 * kq);
    uint32_t *d_count = reinterpret_cast<uint32_t *>(sp + 2 * kq + wq);
    uint8_t *sort_base = sp + 2 * kq + 2 * wq;
    // pageable caller memory goes through a pinned bounce buffer (one memcpy per array)
    const bool bounce_in = !dev_in && (k_lo == kPageable || (!point && k_hi == kPageable));
    const bool bounce_out = !dev_out && (k_first == kPageable || k_count == kPageable);
    const size_t in_bytes = static_cast<size_t>(Q) * ksz, out_bytes = static_cast<size_t>(Q) * 4;
    if ((bounce_in || bounce_out) && !ensure_probe_bounce(g, 2 * kq + 2 * wq)) return -4;
    uint8_t *hb = static_cast<uint8_t *>(g->h_probe_bounce);
    const uint8_t *h_lo = static_cast<const uint8_t *>(lo), *h_hi = static_cast<const uint8_t *>(point ? lo : hi);
    if (bounce_in) {
        std::memcpy(hb, lo, in_bytes);
        h_lo = hb;
        if (!point) {
            std::memcpy(hb + kq, hi, in_bytes);
            h_hi = hb + kq;
        } else {
            h_hi = hb;
        }
    }
    uint8_t *h_first = bounce_out ? hb + 2 * kq : reinterpret_cast<uint8_t *>(first);
    uint8_t *h_count = bounce_out ? hb + 2 * kq + wq : reinterpret_cast<uint8_t *>(count);
    if (dev_in) {
        d_lo = const_cast<void *>(lo);
        d_hi = const_cast<void *>(point ? lo : hi);
    }
    uint32_t *o_first = dev_out ? first : d_first, *o_count = dev_out ? count : d_count;

    engine_resolve_timing(g);  // ev0 / ev1 belong to the last match phase's slot: settle it before reuse
    bool ok = true;
    cudaEventRecord(g->ev0, g->stream);
    if (!sorted) {
        // chunks alternate between the two streams: copy in, probe, copy out
        long long n_chunks = (Q + (1 << 17) - 1) >> 17;
        if (n_chunks > 8) n_chunks = 8;
        if (n_chunks < 1 || (dev_in && dev_out)) n_chunks = 1;
        const long long per = ((Q + n_chunks - 1) / n_chunks + 63) & ~63ll;
        cudaEventRecord(g->ev_seg[0], g->stream);
        cudaStreamWaitEvent(g->stream2, g->ev_seg[0], 0);
        for (long long c = 0; c < n_chunks && ok; ++c) {
            const long long q0 = c * per, q1 = (q0 + per < Q) ? q0 + per : Q;
            if (q0 >= q1) break;
            cudaStream_t st = (c & 1) ? g->stream2 : g->stream;
            const size_t kb = static_cast<size_t>(q1 - q0) * ksz, ko = static_cast<size_t>(q0) * ksz;
            if (!dev_in) {
                ok = ok && cuda_ok(cudaMemcpyAsync(static_cast<uint8_t *>(d_lo) + ko, h_lo + ko, kb, cudaMemcpyHostToDevice, st), "probe h2d");
                if (!point)
                    ok = ok && cuda_ok(cudaMemcpyAsync(static_cast<uint8_t *>(d_hi) + ko, h_hi + ko, kb, cudaMemcpyHostToDevice, st), "probe h2d");
            }
            const void *c_lo = static_cast<const uint8_t *>(d_lo) + ko;
            const void *c_hi = (point && !dev_in) ? c_lo : static_cast<const uint8_t *>(d_hi) + ko;
            ok = ok && cuda_ok(index_probe(ix, c_lo, c_hi, q1 - q0, o_first + q0, o_count + q0, st), "probe kernel launch");
            ++launches;
            if (!dev_out) {
                ok = ok && cuda_ok(cudaMemcpyAsync(h_first + q0 * 4, o_first + q0, static_cast<size_t>(q1 - q0) * 4, cudaMemcpyDeviceToHost, st), "probe d2h");
                ok = ok && cuda_ok(cudaMemcpyAsync(h_count + q0 * 4, o_count + q0, static_cast<size_t>(q1 - q0) * 4, cudaMemcpyDeviceToHost, st), "probe d2h");
            }
        }
        cudaEventRecord(g->ev_seg[1], g->stream2);
        cudaStreamWaitEvent(g->stream, g->ev_seg[1], 0);
    } else {
        if (!dev_in) {
            ok = ok && cuda_ok(cudaMemcpyAsync(d_lo, h_lo, in_bytes, cudaMemcpyHostToDevice, g->stream), "probe h2d");
            if (!point) ok = ok && cuda_ok(cudaMemcpyAsync(d_hi, h_hi, in_bytes, cudaMemcpyHostToDevice, g->stream), "probe h2d");
        }
        void *keys_s = sort_base;
        void *hi_s = sort_base + kq;
        uint32_t *slot_in = reinterpret_cast<uint32_t *>(sort_base + 2 * kq);
        uint32_t *slot_s = reinterpret_cast<uint32_t *>(sort_base + 2 * kq + wq);
        uint32_t *first_s = reinterpret_cast<uint32_t *>(sort_base + 2 * kq + 2 * wq);
        uint32_t *count_s = reinterpret_cast<uint32_t *>(sort_base + 2 * kq + 3 * wq);
        void *tmp = sort_base + 2 * kq + 4 * wq;
        // K4 (radix_sort.cu): by lower key, payload = the probe's slot; int keys order as signed values
        (void)slot_in;
        if (ix.type == T_U64) {
            ok = ok && cuda_ok(radix_sort_pairs<unsigned long long>(static_cast<const unsigned long long *>(d_lo), nullptr, kSortIota,
                                                                    false, static_cast<unsigned long long *>(keys_s), slot_s, Q, tmp,
                                                                    tmp_bytes, g->stream, &launches),
                               "probe sort");
            if (!point)
                gather_keys_kernel<unsigned long long><<<grid_for(Q, 256), 256, 0, g->stream>>>(
                    static_cast<const unsigned long long *>(d_hi), slot_s, static_cast<unsigned long long *>(hi_s), Q);
        } else {
            ok = ok && cuda_ok(radix_sort_pairs<uint32_t>(static_cast<const uint32_t *>(d_lo), nullptr, kSortIota, true,
                                                          static_cast<uint32_t *>(keys_s), slot_s, Q, tmp, tmp_bytes, g->stream,
                                                          &launches),
                               "probe sort");
            if (!point)
                gather_keys_kernel<int><<<grid_for(Q, 256), 256, 0, g->stream>>>(static_cast<const int *>(d_hi), slot_s,
                                                                              static_cast<int *>(hi_s), Q);
        }
        ok = ok && cuda_ok(index_probe(ix, keys_s, point ? keys_s : hi_s, Q, first_s, count_s, g->stream), "probe kernel launch");
        scatter_answers_kernel<<<grid_for(Q, 256), 256, 0, g->stream>>>(first_s, count_s, slot_s, o_first, o_count, Q);
        ok = ok && cuda_ok(cudaGetLastError(), "probe sort kernels");
        launches += point ? 2 : 3;
        if (!dev_out) {
            ok = ok && cuda_ok(cudaMemcpyAsync(h_first, o_first, out_bytes, cudaMemcpyDeviceToHost, g->stream), "probe d2h");
            ok = ok && cuda_ok(cudaMemcpyAsync(h_count, o_count, out_bytes, cudaMemcpyDeviceToHost, g->stream), "probe d2h");
        }
    }
    cudaEventRecord(g->ev1, g->stream);
    ok = cuda_ok(cudaStreamSynchronize(g->stream), "probe sync") && ok;
    if (ok && bounce_out) {
        std::memcpy(first, h_first, out_bytes);
        std::memcpy(count, h_count, out_bytes);
    }
    if (ok && stats) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g->ev0, g->ev1);
        stats->kernel_ms = ms;  // device time of the whole batch, copies included
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        stats->rows_scanned = Q;
        stats->path = 1;
        stats->launches = launches;
        // per probe: two descents, each reading one node (fanout keys) per level incl. the leaf array
        stats->algo_bytes = Q * 2 * (ix.n_levels + 1) * ix.fanout * static_cast<long long>(ksz) +
                            Q * static_cast<long long>(2 * ksz + 8);
    }
    return ok ? 0 : -4;
}

/* K4 on its own (csrc/radix_sort.cu): n (key, payload) pairs from host arrays, sorted by key ascending, stable.
 * See qpe_gpu.h.  Runs on the current device with buffers of its own -- no engine involved. */
int qpe_gpu_sort_pairs(const void *keys, const unsigned int *vals, size_t n, int key_bytes, int signed_keys, int mode,
                       void *keys_out, unsigned int *vals_out, int *passes_out, double *kernel_ms_out) {
    if ((key_bytes != 4 && key_bytes != 8) || mode < 0 || mode > 2 || (n && (!keys || !keys_out || !vals_out)) ||
        (mode == kSortPairs && n && !vals) || n > 0xffffffffull) {
        set_error("qpe_gpu_sort_pairs: bad arguments");
        return -1;
    }
    if (passes_out) *passes_out = 0;
    if (kernel_ms_out) *kernel_ms_out = 0;
    if (n == 0) return 0;
    const long long N = static_cast<long long>(n);
    const size_t kb = n * static_cast<size_t>(key_bytes), vb = n * 4;
    const size_t scratch_bytes = radix_sort_scratch_bytes(N, key_bytes);
    void *d_kin = nullptr, *d_kout = nullptr, *d_scratch = nullptr;
    uint32_t *d_vin = nullptr, *d_vout = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int launches = 0, passes = 0;
    bool ok = cuda_ok(cudaMalloc(&d_kin, kb), "cudaMalloc sort keys") && cuda_ok(cudaMalloc(&d_kout, kb), "cudaMalloc sort keys") &&
              cuda_ok(cudaMalloc(&d_vout, vb), "cudaMalloc sort payload") &&
              cuda_ok(cudaMalloc(&d_scratch, scratch_bytes), "cudaMalloc sort scratch") &&
              cuda_ok(cudaMemcpy(d_kin, keys, kb, cudaMemcpyHostToDevice), "sort h2d") &&
              cuda_ok(cudaEventCreate(&e0), "event") && cuda_ok(cudaEventCreate(&e1), "event");
    if (ok && mode == kSortPairs)
        ok = cuda_ok(cudaMalloc(&d_vin, vb), "cudaMalloc sort payload") &&
             cuda_ok(cudaMemcpy(d_vin, vals, vb, cudaMemcpyHostToDevice), "sort h2d");
    if (ok) {
        cudaEventRecord(e0, nullptr);
        const cudaError_t e =
            key_bytes == 8
                ? radix_sort_pairs<unsigned long long>(static_cast<const unsigned long long *>(d_kin), d_vin,
                                                       static_cast<SortInput>(mode), signed_keys != 0,
                                                       static_cast<unsigned long long *>(d_kout), d_vout, N, d_scratch,
                                                       scratch_bytes, nullptr, &launches, &passes)
                : radix_sort_pairs<uint32_t>(static_cast<const uint32_t *>(d_kin), d_vin, static_cast<SortInput>(mode),
                                             signed_keys != 0, static_cast<uint32_t *>(d_kout), d_vout, N, d_scratch,
                                             scratch_bytes, nullptr, &launches, &passes);
        cudaEventRecord(e1, nullptr);
        ok = cuda_ok(e, "radix sort") && cuda_ok(cudaDeviceSynchronize(), "radix sort") &&
             cuda_ok(cudaMemcpy(keys_out, d_kout, kb, cudaMemcpyDeviceToHost), "sort d2h") &&
             cuda_ok(cudaMemcpy(vals_out, d_vout, vb, cudaMemcpyDeviceToHost), "sort d2h");
        if (ok && kernel_ms_out) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            *kernel_ms_out = ms;
        }
        if (passes_out) *passes_out = passes;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_kin);
    cudaFree(d_kout);
    cudaFree(d_vin);
    cudaFree(d_vout);
    cudaFree(d_scratch);
    return ok ? 0 : -4;
}

}  // extern "C"
