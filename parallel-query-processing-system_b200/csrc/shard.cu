// shard.cu -- a row-range sharded table: one process per GPU, full-scan SELECT with the ordered
// gather done by kernels over peer memory, and up to TWO queries in flight per rank.
//
// What the reference does in MPI mode (engine/mpi/executeEngine-mpi.c:703-770): block partition of
// the rows, every rank evaluates its slice, MPI_Allreduce / MPI_Allgather of the per-rank counts,
// MPI_Allgatherv of the per-rank pieces in partition order.  Here, per query (no host collective,
// no NCCL call, no send/recv on the data path):
//
//   every rank      K1f scans its shard; the ids are already global (+ first row of the shard).
//
//   DEVICE RESULT   (ids packed in the owner's HBM).  The compaction warps store the ids into the rank's segment of
//                   the OWNER's result memory -- peer memory mapped through CUDA IPC, so the ids cross NVLink as the
//                   scan kernel's own coalesced stores, during the scan; the first shard's offset is always 0, so ITS
//                   scan stores straight into the dense result.  post_kernel: one system-scope release store of
//                   (epoch, match count) into EVERY rank's comm block = the count exchange; every rank waits for
//                   all counts of the epoch (acquire loads of its own block); the owner moves the segments of ranks
//                   1.. behind rank 0's ids.
//
//   HOST RESULT     (ids in a host buffer shared by all ranks: POSIX shared memory, registered with CUDA in every
//                   process, each rank's part of it placed on the NUMA node of that rank's GPU).  The ids stay in the
//                   rank's OWN segment (own HBM, mapped by every other rank).  deliver_kernel: the same count
//                   exchange, then rank j takes the j-th 1/world of the RESULT -- whichever ranks' segments it spans
//                   -- reading it over NVLink straight from those segments (no pack, no flag, no second hop) into a
//                   local staging buffer, from where the copy engine takes it to the host over this GPU's own PCIe
//                   link (or, mode 2, the kernel stores it into the mapped host buffer itself).  Every link carries
//                   1/world of the result whatever its distribution over the shards.
//
//   TWO IN FLIGHT   qpe_shard_submit enqueues scan + exchange / delivery kernel and returns; qpe_shard_wait waits for
//                   the counts (a mapped host word written by the kernel: no stream synchronisation), finishes the
//                   host copy and returns the per-rank counts.  A caller that submits query q + 1 before it waits for
//                   q hides the host's share of a query (launch, wake-up) behind the scans and lets the host copy of q
//                   run beside the scan of q + 1.  Every buffer a query touches exists twice (epoch parity).
//
//   ERRORS          are collective: a rank that cannot run its scan (WHERE does not compile, row ids beyond 32 bits,
//                   row too wide to stage) still publishes a failure marker instead of a count, and every decision
//                   that follows (overflow of a segment / of the host buffer) is taken from the exchanged counts, so
//                   all ranks take the same branch and nobody is left waiting.  A wait that times out (~20 s) sets a
//                   status word instead of trapping.
//
//   DELETE          qpe_shard_delete: local mask + stable compaction per shard, the new shard sizes
//                   all-gathered through the same comm blocks, shards renumbered.
//
// Why the buffers of parity e & 1 are free again when the scan of query e + 2 starts: that scan is released by this rank's
// post-scan kernel of e + 1 only AFTER every rank's count of e + 1 has arrived (griddepcontrol.launch_dependents
// follows the exchange), a rank publishes its count of e + 1 only after its scan of e + 1 has completed, and that scan
// -- a programmatic dependent of the post-scan kernel of e -- ends with griddepcontrol.wait, i.e. not before that kernel
// (the last reader of the other ranks' segments of e) has completed.  Releasing the scan at the START of the post-scan
// kernel, as the first version did, let a fast rank overwrite a segment a slow rank was still delivering (caught by
// the 3-processes-on-one-GPU test, where the time slicing makes ranks arbitrarily slow).
// NCCL / torch.distributed is only used by the caller to hand the IPC handles around at start-up.

#include "engine.cuh"

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace qpe {

constexpr int kMaxRanks = 16;
constexpr unsigned long long kCountFailed = 0xffffffffull;   // published instead of a count: this rank could not scan
constexpr int kHostSlots = 3;                                 // id arrays of the shared host buffer: a host result of epoch e lives in slot e % 3
constexpr int kHostWords = kMaxRanks + 2;                     // per parity: counts, [kMaxRanks] = epoch flag, [+1] = status

struct ShardComm {                      // device memory of one rank, mapped by all the others; followed by its segments
    unsigned long long count[2][kMaxRanks];  // [epoch parity][source rank] = epoch << 32 | match count
    unsigned int ctas_done[2];               // [epoch parity] CTAs of the post-scan kernel that have finished
    unsigned int pad_[2];
};
constexpr size_t kCommBytes = (sizeof(ShardComm) + 255) & ~size_t(255);

struct ShardHostHeader {                // start of the shared host buffer
    volatile unsigned long long done[kMaxRanks][8];  // [rank][0] = epoch whose ids rank has delivered (64 B apart)
};
constexpr size_t kHostHeaderBytes = (sizeof(ShardHostHeader) + 4095) & ~size_t(4095);

struct ShardPending {
    bool active = false;
    bool local_ok = false;              // this rank's scan was enqueued (else it published the failure marker)
    int to_host = 0;
    uint32_t epoch = 0;
    double t_begin = 0;
    FusedEnqueue fe;
    std::string local_error;
};

struct ShardState {
    int rank = 0, world = 1;
    ShardComm *comm[kMaxRanks] = {nullptr};  // [rank] = own allocation, others = IPC mappings
    uint32_t *seg[kMaxRanks] = {nullptr};    // [r] = rank r's own segments: 2 parities x seg_cap ids (behind its comm block)
    uint64_t seg_cap = 0;
    uint32_t epoch = 0;
    // device result: memory in the owner's HBM (own pointer on the owner, IPC mapping elsewhere), per epoch parity:
    //   [ dense result: world x dev_cap ids | segments of ranks 1 .. world-1: dev_cap ids each ]
    int owner = 0;
    uint32_t *dev_base = nullptr;
    uint64_t dev_cap = 0;
    uint32_t last_parity = 0;           // parity of the most recent device-result query
    int host_mode = 1;                  // host result: 1 = staging + copy engine, 2 = the kernel stores into host memory
    // host result: shared mapping = header, then 2 parities x host_cap ids
    void *host_map = nullptr;
    size_t host_bytes = 0;
    uint64_t host_cap = 0;
    uint32_t *host_ids = nullptr;       // first parity's ids (host pointer)
    uint32_t *host_ids_dev = nullptr;   // device alias (after qpe_shard_pin_host_result)
    bool host_pinned = false;
    char host_name[96] = {0};
    bool host_creator = false;
    // Deferred completion (qpe_shard_set_deferred): the owner's qpe_shard_wait returns once ITS piece of a host result is
    // delivered; the other ranks' pieces are waited for when the result is asked for (qpe_shard_host_result*).
    bool deferred = false;
    double link_gbs = 0;                // this rank's device->host rate as given to qpe_shard_set_link_weights (GB/s), 0 = unknown
    uint64_t last_bpr = 0;              // bytes per row the previous query's scan read
    uint32_t copy_pending[2] = {0, 0};  // [parity] epoch of a device->host copy that is queued but not yet known complete
    uint32_t host_epoch[2] = {0, 0};    // epochs of the last two host-result queries waited: [0] the most recent
    uint64_t host_total[2] = {0, 0};
    uint32_t *staging[2] = {nullptr, nullptr};   // this rank's slice of the result before it goes to the host
    uint64_t staging_cap = 0;
    cudaEvent_t ev_post[2] = {nullptr, nullptr};  // post-scan kernel of parity p has finished (stream2 waits for it)
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    // per-query words written by the post-scan kernel (mapped pinned host memory), per parity
    unsigned long long *h_words = nullptr;
    unsigned long long *d_words = nullptr;
    ShardPending pending[2];
    int n_pending = 0;
    bool last_was_post = false;         // the engine stream's most recent kernel is a post-scan kernel of this file
    // share of a host result each rank delivers, as cumulative fractions of 2^20 (equal shares unless
    // qpe_shard_set_link_weights: PCIe links of one box can differ by 2x when all of them copy at once)
    uint32_t slice_cum[kMaxRanks + 1] = {0};
    int numa_node = -1, numa_how = 0;   // where this rank's part of the host buffer was placed, and by what (1 mbind, 2 affinity)
    // where qpe_shard_wait spends its time on host-result queries (host clock, summed): waiting for the counts, for this
    // rank's device->host copy, (owner) for the other ranks' pieces
    double wait_ms[3] = {0, 0, 0};
    long long wait_n = 0;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long shard_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// wait until the slot carries `epoch`; bounded by time (~20 s): then *timed_out is set and the stale word returned
__device__ __forceinline__ unsigned long long wait_slot(const unsigned long long *slot, uint32_t epoch, int *timed_out) {
    unsigned long long v = ld_acquire_sys(slot);
    if (static_cast<uint32_t>(v >> 32) == epoch) return v;
    const unsigned long long t0 = shard_now_ns();
    uint32_t spins = 0;
    while (static_cast<uint32_t>(v >> 32) != epoch) {
        __nanosleep(100);
        if ((++spins & 0x3ffu) == 0 && shard_now_ns() - t0 > 20000000000ull) {
            *timed_out = 1;
            return v;
        }
        v = ld_acquire_sys(slot);
    }
    return v;
}

struct PeerPtrs {
    ShardComm *comm[kMaxRanks];
};
struct PeerSegs {
    const uint32_t *seg[kMaxRanks];   // every rank's segment of THIS query's parity
};
struct SliceCum {
    uint32_t cum[kMaxRanks + 1];      // rank j delivers result positions [total * cum[j] >> 20, total * cum[j + 1] >> 20)
};
__host__ __device__ inline unsigned long long slice_bound(unsigned long long total, uint32_t cum) {
    return (total * cum) >> 20;       // total < 2^32, cum <= 2^20: no overflow
}

// ids per parity set of the device result: the dense area, then one segment per rank >= 1
__host__ __device__ inline unsigned long long set_ids(int world, unsigned long long seg_cap) {
    return static_cast<unsigned long long>(2 * world - 1) * seg_cap;
}

// The count exchange, common to both post-scan kernels.  Thread r < world of CTA 0 stores (epoch, count) into rank r's
// comm block; thread r < world of EVERY CTA waits for rank r's count in this rank's own block.  On return s_cnt[r]
// holds the counts (kCountFailed = that rank failed / timed out) -- after a __syncthreads().
__device__ __forceinline__ void exchange_counts_dev(unsigned long long my, const PeerPtrs &peers, int rank, int world,
                                                    uint32_t epoch, unsigned long long *s_cnt, int *s_timeout) {
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();  // this rank's id stores (peer memory included) are visible before the count is
        st_release_sys(&peers.comm[threadIdx.x]->count[epoch & 1u][rank],
                       (static_cast<unsigned long long>(epoch) << 32) | (my & 0xffffffffull));
    }
    if (threadIdx.x < world) {
        const int r = threadIdx.x;
        int to = 0;
        unsigned long long c = my & 0xffffffffull;
        if (r != rank) {
            c = wait_slot(&peers.comm[rank]->count[epoch & 1u][r], epoch, &to) & 0xffffffffull;
            if (to) {
                c = kCountFailed;
                *s_timeout = 1;
            }
        }
        s_cnt[r] = c;
    }
}

// copy n ids, 4 independent loads in flight per thread (a plain grid-stride copy over NVLink / within HBM is
// latency-bound: 78 us -> 17 us for 9.4 M ids)
__device__ __forceinline__ void copy_ids(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, unsigned long long n) {
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const uint32_t a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride),
                       d = __ldcs(src + i + 3 * stride);
        dst[i] = a;
        dst[i + stride] = b;
        dst[i + 2 * stride] = c;
        dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = src[i];
}

// hand the counts to the host (mapped pinned memory): words[r] = count, words[kMaxRanks + 1] = status, and LAST
// words[kMaxRanks] = epoch, which is what the host waits for
__device__ __forceinline__ void publish_to_host(unsigned long long *words, const unsigned long long *s_cnt, int world,
                                                uint32_t epoch, int timed_out) {
    for (int r = 0; r < world; ++r) words[r] = s_cnt[r];
    words[kMaxRanks + 1] = timed_out ? 1ull : 0ull;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(words + kMaxRanks) = static_cast<unsigned long long>(epoch);
}

// DEVICE RESULT: the ONE kernel that follows the scan on every rank (same stream):
//   1. count exchange;
//   2. pack != 0 (owner): the segments of ranks 1.. are moved to their offsets in the dense result;
//   3. the last CTA to finish hands the counts to the host.
// count == nullptr: this rank could not scan and publishes the failure marker.
__global__ void __launch_bounds__(256) post_kernel(const unsigned long long *count, PeerPtrs peers, int rank, int world,
                                                   uint32_t epoch, unsigned long long *host_words, int pack,
                                                   uint32_t *set, unsigned long long seg_cap) {
    __shared__ unsigned long long s_cnt[kMaxRanks];
    __shared__ unsigned long long s_off[kMaxRanks + 1];
    __shared__ int s_timeout, s_ok;
    if (threadIdx.x == 0) s_timeout = 0;
    __syncthreads();
    const unsigned long long my = count ? *reinterpret_cast<const volatile unsigned long long *>(count) : kCountFailed;
    exchange_counts_dev(my, peers, rank, world, epoch, s_cnt, &s_timeout);
    __syncthreads();
    // The next query's scan (a programmatic dependent launch) needs nothing from this kernel and may start now -- but
    // not earlier: it overwrites the segments of the query BEFORE this one (same parity), which the other ranks'
    // post-scan kernels of that query read.  Every rank's count of THIS query has arrived, so every rank has finished
    // that kernel (its stream order); and the count word of this rank's own scan has been read.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        unsigned long long o = 0;
        int ok = 1;
        for (int r = 0; r < world; ++r) {
            const unsigned long long c = s_cnt[r];
            if (c == kCountFailed || c > seg_cap) ok = 0;
            s_off[r] = o;
            o += c;
        }
        s_off[world] = o;
        s_ok = ok;
    }
    __syncthreads();
    // nothing to move (one CTA, or every other rank's count is 0: all matches in the first shard): CTA 0 hands the
    // counts over at once, nobody counts CTAs
    const bool moving = pack && s_ok && gridDim.x > 1 && s_off[world] > s_off[1];
    if (!moving) {
        if (blockIdx.x == 0 && threadIdx.x == 0) publish_to_host(host_words, s_cnt, world, epoch, s_timeout);
        return;
    }
    for (int r = 1; r < world; ++r) {
        const unsigned long long n = s_off[r + 1] - s_off[r];
        copy_ids(set + (static_cast<unsigned long long>(world) + (r - 1)) * seg_cap, set + s_off[r], n);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int *done = &peers.comm[rank]->ctas_done[epoch & 1u];
        if (atomicAdd(done, 1u) == gridDim.x - 1u) {
            *done = 0u;
            __threadfence();
            publish_to_host(host_words, s_cnt, world, epoch, s_timeout);
        }
    }
}

// HOST RESULT: count exchange, then this rank's 1/world of the RESULT is read from whichever ranks' segments it spans
// (NVLink peer loads) into `dst` = local staging (mode 1: the host then queues the copy engine) or the mapped host
// buffer itself (mode 2).  The counts go to the host when the last CTA has stored its ids, so the host needs no
// event between this kernel and the next query's scan to know that the slice is complete.
__global__ void __launch_bounds__(256) deliver_kernel(const unsigned long long *count, PeerPtrs peers, PeerSegs segs, int rank,
                                                      int world, uint32_t epoch, unsigned long long *host_words,
                                                      uint32_t *dst, unsigned long long seg_cap,
                                                      unsigned long long host_cap, unsigned long long dst_cap,
                                                      int direct, SliceCum sc) {
    __shared__ unsigned long long s_cnt[kMaxRanks];
    __shared__ unsigned long long s_off[kMaxRanks + 1];
    __shared__ int s_timeout, s_ok;
    if (threadIdx.x == 0) s_timeout = 0;
    __syncthreads();
    const unsigned long long my = count ? *reinterpret_cast<const volatile unsigned long long *>(count) : kCountFailed;
    exchange_counts_dev(my, peers, rank, world, epoch, s_cnt, &s_timeout);
    __syncthreads();
    // The next query's scan (a programmatic dependent launch) needs nothing from this kernel and may start now -- but
    // not earlier: it overwrites the segments of the query BEFORE this one (same parity), which the other ranks'
    // post-scan kernels of that query read.  Every rank's count of THIS query has arrived, so every rank has finished
    // that kernel (its stream order); and the count word of this rank's own scan has been read.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        unsigned long long o = 0;
        int ok = 1;
        for (int r = 0; r < world; ++r) {
            const unsigned long long c = s_cnt[r];
            if (c == kCountFailed || c > seg_cap) ok = 0;
            s_off[r] = o;
            o += c;
        }
        s_off[world] = o;
        if (o > host_cap) ok = 0;
        s_ok = ok;
    }
    __syncthreads();
    if (s_ok) {
        const unsigned long long total = s_off[world];
        const unsigned long long lo = slice_bound(total, sc.cum[rank]), hi = slice_bound(total, sc.cum[rank + 1]);
        // staging holds the slice from its first id on; the host array is addressed by result position
        uint32_t *d0 = direct ? dst : dst - lo;
        if (direct || hi - lo <= dst_cap)
            for (int r = 0; r < world; ++r) {
                const unsigned long long a = s_off[r] > lo ? s_off[r] : lo;
                const unsigned long long b = s_off[r + 1] < hi ? s_off[r + 1] : hi;
                if (a < b) copy_ids(segs.seg[r] + (a - s_off[r]), d0 + a, b - a);
            }
    }
    // the last CTA to finish hands the counts to the host: the slice is then complete in `dst`
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        unsigned int *done = &peers.comm[rank]->ctas_done[epoch & 1u];
        if (atomicAdd(done, 1u) == gridDim.x - 1u) {
            *done = 0u;
            __threadfence();
            publish_to_host(host_words, s_cnt, world, epoch, s_timeout);
        }
    }
}

// A plain all-gather of one 32-bit value per rank through the comm blocks (same slots, same epoch protocol as
// the post-scan kernels): thread r stores (epoch, value) into rank r's block, then waits for rank r's value.
__global__ void exchange_kernel(PeerPtrs peers, int rank, int world, uint32_t epoch, unsigned long long value,
                                unsigned long long *host_values) {
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    st_release_sys(&peers.comm[r]->count[epoch & 1u][rank],
                   (static_cast<unsigned long long>(epoch) << 32) | (value & 0xffffffffull));
    int to = 0;
    host_values[r] = (r == rank) ? (value & 0xffffffffull)
                                 : (wait_slot(&peers.comm[rank]->count[epoch & 1u][r], epoch, &to) & 0xffffffffull);
    if (to) host_values[kMaxRanks + 1] = 1ull;
}

static ShardState *shard_of(GpuEngine *g) { return static_cast<ShardState *>(g->shard); }

static void release_host_result(ShardState *s) {
    if (!s->host_map) return;
    if (s->host_pinned) cudaHostUnregister(s->host_map);
    munmap(s->host_map, s->host_bytes);
    if (s->host_creator) shm_unlink(s->host_name);
    for (int p = 0; p < 2; ++p) {
        if (s->staging[p]) cudaFree(s->staging[p]);
        s->staging[p] = nullptr;
    }
    s->staging_cap = 0;
    s->host_map = nullptr;
    s->host_epoch[0] = s->host_epoch[1] = 0;
    s->host_ids = nullptr;
    s->host_ids_dev = nullptr;
    s->host_pinned = false;
    s->host_bytes = 0;
    s->host_cap = 0;
    s->host_creator = false;
}

static bool complete_own_copies(ShardState *s);

void shard_destroy(GpuEngine *g) {
    ShardState *s = shard_of(g);
    if (!s) return;
    cudaSetDevice(g->device);
    cudaStreamSynchronize(g->stream);
    cudaStreamSynchronize(g->stream2);
    complete_own_copies(s);  // (deferred mode) announce what this rank still owes the others
    for (int r = 0; r < s->world; ++r) {
        if (!s->comm[r]) continue;
        if (r == s->rank)
            cudaFree(s->comm[r]);
        else
            cudaIpcCloseMemHandle(s->comm[r]);
    }
    if (s->h_words) cudaFreeHost(s->h_words);
    release_host_result(s);
    for (int p = 0; p < 2; ++p) {
        if (s->ev_post[p]) cudaEventDestroy(s->ev_post[p]);
        if (s->ev_copy[p]) cudaEventDestroy(s->ev_copy[p]);
    }
    delete s;
    g->shard = nullptr;
}

// NUMA node of this process's GPU (from sysfs), -1 when it cannot be told
static int gpu_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *c = bus; *c; ++c)
        if (*c >= 'A' && *c <= 'Z') *c = static_cast<char>(*c - 'A' + 'a');
    char path[128];
    std::snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = std::fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}

// Place [p, p + bytes) on `node`: first touch by this thread while it is pinned to that node's CPUs (the default
// first-touch policy then allocates there; mbind is tried as well, but containers usually filter that system call).
// Best effort: on a single-node box, or when the node cannot be told, the pages land wherever the first touch puts them.
static int place_on_node(void *p, size_t bytes, int node) {
    if (bytes == 0) return 0;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) & ~uintptr_t(4095);
    const uintptr_t e = (reinterpret_cast<uintptr_t>(p) + bytes + 4095) & ~uintptr_t(4095);
    int how = 0;
    cpu_set_t old_set, node_set;
    bool pinned = false;
    if (node >= 0) {
#ifdef SYS_mbind
        if (node < 64) {
            unsigned long mask = 1ul << node;
            if (syscall(SYS_mbind, reinterpret_cast<void *>(a), static_cast<unsigned long>(e - a), 1 /* MPOL_PREFERRED */, &mask,
                        sizeof(mask) * 8 + 1, 0u) == 0)
                how |= 1;
        }
#endif
        char path[96];
        std::snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
        FILE *f = std::fopen(path, "r");
        if (f) {
            CPU_ZERO(&node_set);
            int lo = 0, hi = 0, n = 0;
            char sep = 0;
            while (std::fscanf(f, "%d", &lo) == 1) {
                hi = lo;
                const int c = std::fgetc(f);
                if (c == '-') {
                    if (std::fscanf(f, "%d", &hi) != 1) hi = lo;
                    sep = static_cast<char>(std::fgetc(f));
                } else {
                    sep = static_cast<char>(c);
                }
                for (int k = lo; k <= hi && k < CPU_SETSIZE; ++k) {
                    CPU_SET(k, &node_set);
                    ++n;
                }
                if (sep != ',') break;
            }
            std::fclose(f);
            if (n > 0 && sched_getaffinity(0, sizeof old_set, &old_set) == 0 &&
                sched_setaffinity(0, sizeof node_set, &node_set) == 0) {
                pinned = true;
                how |= 2;
            }
        }
    }
    madvise(reinterpret_cast<void *>(a), e - a, MADV_HUGEPAGE);  // helps only where shmem huge pages are "advise"
    volatile char *c = reinterpret_cast<volatile char *>(a);
    for (uintptr_t o = 0; o < e - a; o += 4096) c[o] = 0;
    if (pinned) sched_setaffinity(0, sizeof old_set, &old_set);
    return how;
}

// wait (spinning on mapped host memory) until the post-scan kernel of `epoch` has handed its counts over
static int wait_host_words(GpuEngine *g, ShardState *s, uint32_t epoch) {
    volatile unsigned long long *w = s->h_words + static_cast<size_t>(epoch & 1u) * kHostWords;
    unsigned long long spins = 0;
    while (w[kMaxRanks] != epoch) {
        if ((++spins & 0xfffu) == 0) {
            const cudaError_t q = cudaStreamQuery(g->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) {
                cuda_ok(q, "sharded SELECT kernels");
                return -4;
            }
            if (q == cudaSuccess && w[kMaxRanks] != epoch) {
                set_error("sharded SELECT: the post-scan kernel finished without handing the counts over");
                return -4;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (w[kMaxRanks + 1] != 0) {
        set_error("sharded SELECT: a rank never published its count (timed out after ~20 s)");
        return -4;
    }
    return 0;
}

// Deferred mode: this rank's queued device->host copies are waited for (oldest first) and announced to the other ranks.
static bool complete_own_copies(ShardState *s) {
    bool ok = true;
    for (int k = 0; k < 2; ++k) {
        int p = -1;  // the parity with the older pending epoch
        for (int q = 0; q < 2; ++q)
            if (s->copy_pending[q] && (p < 0 || static_cast<int32_t>(s->copy_pending[q] - s->copy_pending[p]) < 0)) p = q;
        if (p < 0) break;
        ok = cuda_ok(cudaEventSynchronize(s->ev_copy[p]), "download ids") && ok;
        ShardHostHeader *hh = static_cast<ShardHostHeader *>(s->host_map);
        if (hh) __atomic_store_n(&hh->done[s->rank][0], static_cast<unsigned long long>(s->copy_pending[p]), __ATOMIC_RELEASE);
        s->copy_pending[p] = 0;
    }
    return ok;
}

// every rank has delivered its piece of the host result of `epoch` (done[] only grows: later results imply earlier ones)
static bool wait_all_delivered(ShardState *s, uint32_t epoch) {
    ShardHostHeader *hh = static_cast<ShardHostHeader *>(s->host_map);
    for (int r = 0; r < s->world; ++r) {
        unsigned long long spins = 0;
        while (static_cast<int32_t>(static_cast<uint32_t>(__atomic_load_n(&hh->done[r][0], __ATOMIC_ACQUIRE)) - epoch) < 0) {
            if (++spins > 4000000000ull) {
                set_error("sharded SELECT: a rank never delivered its ids");
                return false;
            }
        }
    }
    return true;
}

}  // namespace qpe

using namespace qpe;

extern "C" {

/* Allocates this rank's comm block and its OWN id segments (2 parities x segment_capacity ids, one allocation) and
 * returns the CUDA IPC handle of it for the other ranks.  segment_capacity must be the same on every rank. */
int qpe_shard_init(struct engineS *engine, int rank, int world, unsigned long long segment_capacity,
                   unsigned char comm_handle_out[64]) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || segment_capacity == 0) {
        set_error("qpe_shard_init: rank / world out of range (at most 16 ranks) or no segment capacity");
        return -5;
    }
    if (g->shard) shard_destroy(g);
    cudaSetDevice(g->device);
    ShardState *s = new ShardState();
    s->rank = rank;
    s->world = world;
    s->seg_cap = segment_capacity;
    for (int r = 0; r <= world; ++r) s->slice_cum[r] = static_cast<uint32_t>((static_cast<unsigned long long>(r) << 20) / world);
    const size_t bytes = kCommBytes + 2 * static_cast<size_t>(segment_capacity) * sizeof(uint32_t) + 256;
    void *block = nullptr;
    bool ok = cuda_ok(cudaMalloc(&block, bytes), "cudaMalloc comm + segments") &&
              cuda_ok(cudaMemset(block, 0, kCommBytes), "cudaMemset comm") &&
              cuda_ok(cudaHostAlloc(&s->h_words, sizeof(unsigned long long) * 2 * kHostWords, cudaHostAllocMapped),
                      "cudaHostAlloc counts") &&
              cuda_ok(cudaHostGetDevicePointer(&s->d_words, s->h_words, 0), "cudaHostGetDevicePointer");
    if (ok) std::memset(s->h_words, 0, sizeof(unsigned long long) * 2 * kHostWords);
    s->comm[rank] = static_cast<ShardComm *>(block);
    if (block) s->seg[rank] = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(block) + kCommBytes);
    for (int p = 0; p < 2 && ok; ++p)
        ok = cuda_ok(cudaEventCreateWithFlags(&s->ev_post[p], cudaEventDisableTiming), "cudaEventCreate") &&
             cuda_ok(cudaEventCreateWithFlags(&s->ev_copy[p], cudaEventDisableTiming), "cudaEventCreate");
    cudaIpcMemHandle_t h;
    ok = ok && cuda_ok(cudaIpcGetMemHandle(&h, block), "cudaIpcGetMemHandle");
    g->shard = s;
    if (!ok) {
        shard_destroy(g);
        return -4;
    }
    std::memcpy(comm_handle_out, &h, 64);
    return 0;
}

/* all_handles: world x 64 bytes, in rank order (what every rank's qpe_shard_init returned) */
int qpe_shard_connect(struct engineS *engine, const unsigned char *all_handles) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s) {
        set_error("qpe_shard_connect: call qpe_shard_init first");
        return -1;
    }
    cudaSetDevice(g->device);
    for (int r = 0; r < s->world; ++r) {
        if (r == s->rank) continue;
        if (s->comm[r]) {  // connected before: drop the old mapping
            cudaIpcCloseMemHandle(s->comm[r]);
            s->comm[r] = nullptr;
            s->seg[r] = nullptr;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, all_handles + 64 * r, 64);
        void *p = nullptr;
        if (!cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle comm")) return -4;
        s->comm[r] = static_cast<ShardComm *>(p);
        s->seg[r] = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(p) + kCommBytes);
    }
    return 0;
}

/* Device result memory, in the OWNER's HBM (the owner passes its own allocation, the other ranks their IPC
 * mapping of it): qpe_shard_result_ids(world, segment_capacity) ids.  Per epoch parity it holds the dense
 * result (world x segment_capacity ids; the first shard's scan stores straight into it, at offset 0) and
 * one segment of segment_capacity ids for every other rank. */
unsigned long long qpe_shard_result_ids(int world, unsigned long long segment_capacity) {
    return 2ull * set_ids(world, segment_capacity) + 16;
}

int qpe_shard_set_device_result(struct engineS *engine, int owner_rank, unsigned int *segments,
                                unsigned long long segment_capacity) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s) {
        set_error("qpe_shard_set_device_result: call qpe_shard_init first");
        return -1;
    }
    s->owner = owner_rank;
    s->dev_base = segments;
    s->dev_cap = segment_capacity;
    return 0;
}

/* Host result: a buffer of 2 x `capacity` ids (two queries can be in flight) shared by all ranks of the box.
 * create != 0 on exactly one rank (it creates /dev/shm/<name>), the others open it afterwards.  This maps the buffer
 * and places THIS rank's parts of it on the NUMA node of its GPU; qpe_shard_pin_host_result, called once every
 * rank has opened it, registers it with CUDA.  Returns the host pointer of the first id array (NULL on failure);
 * three of them follow each other, `capacity` (rounded up to 1 Ki) ids apart: a host result of epoch e uses array e % 3. */
unsigned int *qpe_shard_open_host_result(struct engineS *engine, const char *name, unsigned long long capacity,
                                         int create) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s || !name || std::strlen(name) >= sizeof(s->host_name) || capacity == 0) {
        set_error("qpe_shard_open_host_result: call qpe_shard_init first / bad name / no capacity");
        return nullptr;
    }
    cudaSetDevice(g->device);
    if (s->n_pending) {
        set_error("qpe_shard_open_host_result: queries are still in flight");
        return nullptr;
    }
    if (s->host_map) {  // opened before: release the old buffer first
        cudaStreamSynchronize(g->stream);
        cudaStreamSynchronize(g->stream2);
        complete_own_copies(s);
        release_host_result(s);
    }
    const uint64_t cap = (capacity + 1023) & ~uint64_t(1023);  // parity 1 starts page aligned
    const size_t bytes = kHostHeaderBytes + sizeof(uint32_t) * (kHostSlots * cap + 16);
    const int fd = shm_open(name, create ? (O_CREAT | O_RDWR | O_TRUNC) : O_RDWR, 0600);
    if (fd < 0) {
        set_error(std::string("shm_open failed for ") + name);
        return nullptr;
    }
    if (create && ftruncate(fd, static_cast<off_t>(bytes)) != 0) {
        close(fd);
        set_error("ftruncate failed on the shared result buffer");
        return nullptr;
    }
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        set_error("mmap failed on the shared result buffer");
        return nullptr;
    }
    if (create) std::memset(p, 0, sizeof(ShardHostHeader));
    s->host_map = p;
    s->host_bytes = bytes;
    s->host_cap = cap;
    s->host_creator = create != 0;
    s->host_ids = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(p) + kHostHeaderBytes);
    std::strncpy(s->host_name, name, sizeof(s->host_name) - 1);
    // rank j delivers the j-th 1/world of every result: put that part of both id arrays next to its GPU
    const int node = gpu_numa_node(g->device);
    s->numa_node = node;
    for (int slot = 0; slot < kHostSlots; ++slot) {
        const uint64_t lo = cap * s->rank / s->world, hi = cap * (s->rank + 1) / s->world;
        s->numa_how = place_on_node(s->host_ids + slot * cap + lo, (hi - lo) * sizeof(uint32_t), node);
    }
    // staging for this rank's slice (mode 1)
    s->staging_cap = cap;  // any share of a result fits (qpe_shard_set_link_weights may give a fast link most of it)
    for (int par = 0; par < 2; ++par)
        if (!cuda_ok(cudaMalloc(&s->staging[par], s->staging_cap * sizeof(uint32_t)), "cudaMalloc staging")) {
            release_host_result(s);
            return nullptr;
        }
    return s->host_ids;
}

/* Once every rank has opened (and placed) the shared buffer: register it with CUDA in this process. */
int qpe_shard_pin_host_result(struct engineS *engine) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s || !s->host_map) {
        set_error("qpe_shard_pin_host_result: open the host result first");
        return -1;
    }
    if (s->host_pinned) return 0;
    cudaSetDevice(g->device);
    if (!cuda_ok(cudaHostRegister(s->host_map, s->host_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped),
                 "cudaHostRegister shared result"))
        return -4;
    s->host_pinned = true;
    void *d = nullptr;
    if (!cuda_ok(cudaHostGetDevicePointer(&d, s->host_ids, 0), "cudaHostGetDevicePointer shared result")) return -4;
    s->host_ids_dev = static_cast<uint32_t *>(d);
    return 0;
}

/* Host result path: 1 = every rank stages its 1/world of the result in its HBM and the copy engine takes it to the
 * host (default), 2 = the delivery kernel stores into the mapped host buffer itself.  Every rank must choose the same. */
int qpe_shard_set_multipath(struct engineS *engine, int mode) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s || (mode != 1 && mode != 2) || s->n_pending) {
        set_error("qpe_shard_set_multipath: call qpe_shard_init first; mode is 1 or 2; no query may be in flight");
        return -1;
    }
    s->host_mode = mode;
    return 0;
}

/* Shares of a host result per rank, proportional to `weights` (e.g. the device->host rate of every rank's PCIe link
 * measured with all links busy).  Every rank must pass the same numbers.  Equal shares by default. */
int qpe_shard_set_link_weights(struct engineS *engine, const double *weights, int n) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s || n != s->world || s->n_pending) {
        set_error("qpe_shard_set_link_weights: one weight per rank, no query in flight");
        return -1;
    }
    double sum = 0;
    for (int r = 0; r < n; ++r) {
        if (!(weights[r] > 0)) {
            set_error("qpe_shard_set_link_weights: weights must be positive");
            return -5;
        }
        sum += weights[r];
    }
    double run = 0;
    s->slice_cum[0] = 0;
    for (int r = 0; r < n; ++r) {
        run += weights[r];
        s->slice_cum[r + 1] = static_cast<uint32_t>(run / sum * 1048576.0 + 0.5);
    }
    s->slice_cum[n] = 1u << 20;
    s->link_gbs = weights[s->rank];
    return 0;
}

/* Creator only, once every rank has opened the shared host buffer: remove its name from /dev/shm (the mappings
 * stay valid), so that nothing is left behind if a process dies. */
int qpe_shard_unlink_host_result(struct engineS *engine) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s) return -1;
    if (s->host_map && s->host_creator) {
        shm_unlink(s->host_name);
        s->host_creator = false;
    }
    return 0;
}

const unsigned int *qpe_shard_device_result(struct engineS *engine) {
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    return (s && s->dev_base) ? s->dev_base + s->last_parity * set_ids(s->world, s->dev_cap) : nullptr;
}

/* the id array of the most recent host-result query qpe_shard_wait completed (back = 0) or of the one before it
 * (back = 1); *total_out = its number of ids.  Blocks until every rank has delivered its piece (a no-op unless
 * qpe_shard_set_deferred is on).  NULL if there is no such result or a rank never delivered. */
const unsigned int *qpe_shard_host_result_at(struct engineS *engine, int back, unsigned long long *total_out) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s || !s->host_ids || back < 0 || back > 1 || s->host_epoch[back] == 0) return nullptr;
    const uint32_t epoch = s->host_epoch[back];
    cudaSetDevice(g->device);
    if (!complete_own_copies(s) || !wait_all_delivered(s, epoch)) return nullptr;
    if (total_out) *total_out = s->host_total[back];
    return s->host_ids + static_cast<size_t>(epoch % kHostSlots) * s->host_cap;
}
const unsigned int *qpe_shard_host_result(struct engineS *engine) {
    {
        EngineLock lk(engine);
        GpuEngine *g = as_engine(engine);
        ShardState *s = g ? shard_of(g) : nullptr;
        if (s && s->host_ids && s->host_epoch[0] == 0) return s->host_ids;  // no host result yet: the first id array
    }
    return qpe_shard_host_result_at(engine, 0, nullptr);
}

/* Deferred completion of host results.  on != 0: qpe_shard_wait returns as soon as THIS rank's piece of a host result
 * is in host memory -- on the owner too, which otherwise also waits for every other rank's piece before it may submit its
 * next query (the owner's scan then starts late, and with it every rank's count exchange).  The complete result is
 * asked for with qpe_shard_host_result / qpe_shard_host_result_at, which wait for the missing pieces; with three id
 * arrays a result stays valid until the third qpe_shard_submit after its own, so a consumer can take result q after
 * submitting q + 2.  Every rank may choose for itself (only the owner waits for other ranks at all). */
int qpe_shard_set_deferred(struct engineS *engine, int on) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s) return -1;
    cudaSetDevice(g->device);
    const bool ok = complete_own_copies(s);
    s->deferred = on != 0;
    return ok ? 0 : -4;
}

/* NUMA node of this rank's GPU (-1: unknown) and how its part of the host buffer was placed there: bit 0 = mbind
 * accepted, bit 1 = first touch under that node's CPU affinity. */
int qpe_shard_numa(struct engineS *engine, int *how_out) {
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (how_out) *how_out = s ? s->numa_how : 0;
    return s ? s->numa_node : -1;
}

void qpe_shard_close(struct engineS *engine) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    if (g) shard_destroy(g);
}

/* Enqueue one sharded full-scan SELECT and return; every rank of the group calls it with the same statement, in the
 * same order, and qpe_shard_wait once per submit, in the same order.  At most two queries may be in flight.
 * to_host == 0: the ids end up packed in the owner's HBM (qpe_shard_device_result); to_host != 0: in the shared host
 * buffer (qpe_shard_host_result).  A rank whose scan cannot be enqueued still takes part (it publishes a failure
 * marker) and learns the error from qpe_shard_wait, like every other rank. */
int qpe_shard_submit(struct engineS *engine, struct whereClauseS *whereClause, int to_host) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s) {
        set_error("qpe_shard_submit: call qpe_shard_init / qpe_shard_connect first");
        return -1;
    }
    if (to_host ? !(s->host_map && s->host_pinned) : !s->dev_base) {
        set_error(to_host ? "qpe_shard_submit: no (pinned) host result buffer" : "qpe_shard_submit: no device result memory");
        return -1;
    }
    if (s->n_pending >= 2) {
        set_error("qpe_shard_submit: two queries are already in flight (call qpe_shard_wait)");
        return -1;
    }
    cudaSetDevice(g->device);
    if (++s->epoch == 0) s->epoch = 1;
    const uint32_t epoch = s->epoch, par = epoch & 1u;
    ShardPending &pq = s->pending[par];
    pq = ShardPending();
    pq.active = true;
    pq.to_host = to_host;
    pq.epoch = epoch;
    pq.t_begin = now_ms();
    ++s->n_pending;
    PeerPtrs peers{};
    for (int r = 0; r < s->world; ++r) peers.comm[r] = s->comm[r];

    // where this rank's scan stores its ids
    uint32_t *out = nullptr;
    uint64_t out_cap = 0;
    uint32_t *set = nullptr;
    if (to_host) {
        out = s->seg[s->rank] + static_cast<size_t>(par) * s->seg_cap;  // own segment, read by the other ranks
        out_cap = s->seg_cap;
    } else {
        set = s->dev_base + par * set_ids(s->world, s->dev_cap);
        // the first shard's offset in the result is always 0: its scan writes the dense result in place
        out = s->rank == 0 ? set : set + (static_cast<size_t>(s->world) + (s->rank - 1)) * s->dev_cap;
        out_cap = s->dev_cap;
    }
    bool local_ok = true;
    if (g->table.row_base + static_cast<uint64_t>(g->table.n) > 0xffffffffull) {
        set_error("global row ids do not fit 32 bits");
        local_ok = false;
    }
    // With a query already in flight the stream's last kernel is ITS count exchange / delivery kernel, which hands
    // nothing to this scan (other parity of every buffer): the scan is launched as a programmatic dependent, so its
    // CTAs move in while that kernel is still waiting for the other ranks / moving ids.
    static const bool no_pdl = std::getenv("QPE_SHARD_NO_PDL") != nullptr;
    const bool overlap = s->n_pending >= 2 && s->last_was_post && !no_pdl;
    if (local_ok)
    {
        // Host result: the table is streamed through L2 with the evict_first policy, so that the staging slice the copy
        // engine reads right after the delivery kernel is still in L2 while the NEXT scan saturates HBM -- when the copy
        // is what the pipeline waits for.  The policy costs the scan 1-3 %, so it is used only if this rank's copy (its
        // share of a result like the last one, at the link rate given to qpe_shard_set_link_weights) takes more than
        // 0.6 of its scan (rows x bytes per row at ~6.5 TB/s): 8 GPUs of this pool (3 MB over a 13 GB/s link against a
        // 0.25 ms scan) yes, 2 or 4 GPUs (19 / 9 MB over 46 GB/s against 0.92 / 0.49 ms) no.  Unknown rate: from 8 ranks on.
        bool stream_l2 = false;
        if (to_host) {
            const double share = static_cast<double>(s->slice_cum[s->rank + 1] - s->slice_cum[s->rank]) / 1048576.0;
            const double copy_s = s->link_gbs > 0 ? static_cast<double>(s->host_total[0]) * 4.0 * share / (s->link_gbs * 1e9) : 0.0;
            const double scan_s = static_cast<double>(g->table.n) * static_cast<double>(s->last_bpr ? s->last_bpr : 13) / 6.5e12;
            stream_l2 = s->link_gbs > 0 && s->host_epoch[0] ? copy_s > 0.6 * scan_s : s->world >= 8;
        }
        local_ok = engine_fused_enqueue(g, whereClause, out, out_cap, static_cast<uint32_t>(g->table.row_base), &pq.fe, overlap,
                                        stream_l2);
        if (local_ok) s->last_bpr = pq.fe.bytes_per_row;
    }
    if (!local_ok) pq.local_error = last_error_cstr();
    pq.local_ok = local_ok;
    const unsigned long long *count = local_ok ? &g->d_fctl->final_count : nullptr;
    unsigned long long *words = s->d_words + static_cast<size_t>(par) * kHostWords;
    if (to_host) {
        PeerSegs segs{};
        for (int r = 0; r < s->world; ++r) segs.seg[r] = s->seg[r] + static_cast<size_t>(par) * s->seg_cap;
        const bool direct = s->host_mode == 2;
        uint32_t *dst = direct ? s->host_ids_dev + static_cast<size_t>(epoch % kHostSlots) * s->host_cap : s->staging[par];
        SliceCum sc{};
        for (int r = 0; r <= s->world; ++r) sc.cum[r] = s->slice_cum[r];
        // (deferred mode) the copy that last read this staging buffer may still be queued: the kernel waits for it
        if (!direct && s->copy_pending[par]) cudaStreamWaitEvent(g->stream, s->ev_copy[par], 0);
        deliver_kernel<<<148 * 4, 256, 0, g->stream>>>(count, peers, segs, s->rank, s->world, epoch, words, dst, s->seg_cap,
                                                       s->host_cap, s->staging_cap, direct ? 1 : 0, sc);
    } else {
        const int pack = (s->rank == s->owner && s->world > 1) ? 1 : 0;
        post_kernel<<<pack ? 148 * 8 : 1, 256, 0, g->stream>>>(count, peers, s->rank, s->world, epoch, words, pack, set,
                                                               s->dev_cap);
    }
    const bool launched = cuda_ok(cudaGetLastError(), "shard post-scan kernel launch");
    s->last_was_post = launched;
    // (no event may sit between this kernel and the next query's scan if that scan is to overlap it)
    if (pq.fe.slot && !overlap) cudaEventRecord(pq.fe.slot->ev[3], g->stream);
    if (!launched) {
        // nothing was published: the other ranks will time out; report it here at once
        pq.active = false;
        --s->n_pending;
        return -4;
    }
    return 0;
}

/* Wait for the OLDEST query in flight.  counts_out[world] = per-rank match counts (partition order); the result is
 * their concatenation = table order.  Returns 0, -5 if a rank's ids did not fit its segment or the result did not fit
 * the host buffer (every rank returns it; nothing was written past a buffer), -2 / -6 if a rank could not run its scan
 * (every rank returns it), -4 on a CUDA error or a rank that never arrived. */
int qpe_shard_wait(struct engineS *engine, unsigned long long *counts_out, qpe_scan_stats *stats) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s || s->n_pending == 0) {
        set_error("qpe_shard_wait: no query in flight");
        return -1;
    }
    cudaSetDevice(g->device);
    // the oldest pending query: with two in flight it is the one with the smaller epoch
    ShardPending *pq = nullptr;
    for (int p = 0; p < 2; ++p)
        if (s->pending[p].active && (!pq || static_cast<int32_t>(s->pending[p].epoch - pq->epoch) < 0)) pq = &s->pending[p];
    const uint32_t epoch = pq->epoch, par = epoch & 1u;
    const double tw0 = now_ms();
    // (deferred mode) the copy of the query before this one: it is ahead of this query's copy on the copy stream anyway
    const bool older_ok = complete_own_copies(s);
    int rc = wait_host_words(g, s, epoch);
    if (!older_ok && rc == 0) rc = -4;
    const double tw1 = now_ms();
    pq->active = false;
    --s->n_pending;
    if (rc != 0) return rc;
    const volatile unsigned long long *w = s->h_words + static_cast<size_t>(par) * kHostWords;
    unsigned long long total = 0;
    bool failed = false, overflow = false;
    const unsigned long long cap = pq->to_host ? s->seg_cap : s->dev_cap;
    for (int r = 0; r < s->world; ++r) {
        const unsigned long long c = w[r];
        if (c == kCountFailed) {
            failed = true;
            if (counts_out) counts_out[r] = 0;
            continue;
        }
        if (counts_out) counts_out[r] = c;
        if (c > cap) overflow = true;
        total += c;
    }
    if (pq->to_host && total > s->host_cap) overflow = true;
    const unsigned long long mine = pq->local_ok ? w[s->rank] : 0;
    int extra = 1;
    if (pq->to_host) {
        ShardHostHeader *hh = static_cast<ShardHostHeader *>(s->host_map);
        uint32_t *host_ids = s->host_ids + static_cast<size_t>(epoch % kHostSlots) * s->host_cap;
        if (!failed && !overflow && s->host_mode == 1) {
            // this rank's 1/world of the result: local staging -> (its own PCIe link) -> host, on the copy stream, so
            // that the next query's scan (already enqueued on the main stream) runs beside it
            const unsigned long long lo = slice_bound(total, s->slice_cum[s->rank]), hi = slice_bound(total, s->slice_cum[s->rank + 1]);
            if (hi > lo && hi - lo <= s->staging_cap) {
                if (!cuda_ok(cudaMemcpyAsync(host_ids + lo, s->staging[par], (hi - lo) * sizeof(uint32_t),
                                             cudaMemcpyDeviceToHost, g->stream2),
                             "download ids") ||
                    !cuda_ok(cudaEventRecord(s->ev_copy[par], g->stream2), "download ids"))
                    rc = -4;
                else if (s->deferred)
                    s->copy_pending[par] = epoch;  // completed and announced by the next wait / when the result is taken
                else if (!cuda_ok(cudaEventSynchronize(s->ev_copy[par]), "download ids"))
                    rc = -4;
            }
        }
        // (mode 2: the kernel stored the ids and fenced before it handed the counts over)
        if (!s->copy_pending[par])
            __atomic_store_n(&hh->done[s->rank][0], static_cast<unsigned long long>(epoch), __ATOMIC_RELEASE);
        const double tw2 = now_ms();
        // the result is complete when every rank has delivered its piece (deferred: asked for with the result)
        if (s->rank == s->owner && rc == 0 && !s->deferred && !wait_all_delivered(s, epoch)) return -4;
        s->host_epoch[1] = s->host_epoch[0];
        s->host_total[1] = s->host_total[0];
        s->host_epoch[0] = epoch;
        s->host_total[0] = (failed || overflow) ? 0 : total;
        s->wait_ms[0] += tw1 - tw0;
        s->wait_ms[1] += tw2 - tw1;
        s->wait_ms[2] += now_ms() - tw2;
        ++s->wait_n;
    } else {
        s->last_parity = par;
    }
    engine_fused_finish(g, pq->fe, mine, extra, pq->t_begin);
    if (rc != 0) return rc;
    if (failed) {
        if (pq->local_ok)
            set_error("qpe_shard_select: another rank could not run its scan");
        else
            set_error(pq->local_error);
        return pq->local_ok ? -2 : -6;
    }
    if (overflow) {
        set_error("a rank's ids did not fit its segment / the result buffer (nothing was written past it)");
        return -5;
    }
    if (stats) qpe_gpu_last_stats(engine, stats);
    return 0;
}

/* Diagnostics: where qpe_shard_wait has spent its time on host-result queries since the last reset (host clock, ms,
 * summed over *n_out queries): [0] waiting for the counts (scan + exchange + delivery kernel), [1] this rank's
 * device->host copy, [2] (owner) the other ranks' pieces. */
int qpe_shard_wait_breakdown(struct engineS *engine, double out_ms[3], long long *n_out, int reset) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s) return -1;
    for (int i = 0; i < 3; ++i) {
        if (out_ms) out_ms[i] = s->wait_ms[i];
        if (reset) s->wait_ms[i] = 0;
    }
    if (n_out) *n_out = s->wait_n;
    if (reset) s->wait_n = 0;
    return 0;
}

/* One sharded full-scan SELECT, start to finish (submit + wait). */
int qpe_shard_select(struct engineS *engine, struct whereClauseS *whereClause, int to_host,
                     unsigned long long *counts_out, qpe_scan_stats *stats) {
    const int rc = qpe_shard_submit(engine, whereClause, to_host);
    if (rc != 0) return rc;
    return qpe_shard_wait(engine, counts_out, stats);
}

/* DELETE on a sharded table (the MPI engine's executeQueryDeleteMPI, engine/mpi/executeEngine-mpi.c:703-770:
 * local match flags, MPI_Allreduce of the count, renumbering): every rank deletes the matching rows of ITS shard
 * (K1 with the inverted program + stable column compaction, as executeQueryDeleteGPU), then the new shard sizes
 * and the deleted counts are all-gathered through the comm blocks and every shard's first global row becomes
 * the sum of the lower shards' new sizes -- global row ids stay positions in the whole table, in table order.
 * Every rank calls it with the same statement.  No data file is touched (a sharded table has none). */
int qpe_shard_delete(struct engineS *engine, struct whereClauseS *whereClause, unsigned long long *deleted_total_out,
                     unsigned long long *rows_total_out) {
    EngineLock lk(engine);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s) {
        set_error("qpe_shard_delete: call qpe_shard_init / qpe_shard_connect first");
        return -1;
    }
    cudaSetDevice(g->device);
    complete_own_copies(s);
    if (s->n_pending) {
        set_error("qpe_shard_delete: queries are still in flight");
        return -1;
    }
    cudaSetDevice(g->device);
    int64_t deleted = 0;
    s->last_was_post = false;
    // a rank whose DELETE fails still takes part in both exchanges (it publishes its old size and 0 deleted rows)
    const bool local_ok = engine_delete(g, whereClause, &deleted);
    const std::string local_error = local_ok ? std::string() : std::string(last_error_cstr());
    PeerPtrs peers{};
    for (int r = 0; r < s->world; ++r) peers.comm[r] = s->comm[r];
    unsigned long long rows_before = 0, rows_total = 0, deleted_total = 0;
    const unsigned long long mine[2] = {static_cast<unsigned long long>(g->table.n),
                                        static_cast<unsigned long long>(local_ok ? deleted : 0)};
    for (int pass = 0; pass < 2; ++pass) {
        if (++s->epoch == 0) s->epoch = 1;
        unsigned long long *words = s->d_words + static_cast<size_t>(s->epoch & 1u) * kHostWords;
        volatile unsigned long long *hw = s->h_words + static_cast<size_t>(s->epoch & 1u) * kHostWords;
        hw[kMaxRanks + 1] = 0;
        exchange_kernel<<<1, 32, 0, g->stream>>>(peers, s->rank, s->world, s->epoch, mine[pass], words);
        if (!cuda_ok(cudaGetLastError(), "shard exchange kernel launch") ||
            !cuda_ok(cudaStreamSynchronize(g->stream), "shard exchange"))
            return -4;
        if (hw[kMaxRanks + 1] != 0) {
            set_error("qpe_shard_delete: a rank never published its value (timed out)");
            return -4;
        }
        for (int r = 0; r < s->world; ++r) {
            const unsigned long long v = hw[r];
            if (pass == 0) {
                rows_total += v;
                if (r < s->rank) rows_before += v;
            } else {
                deleted_total += v;
            }
        }
    }
    g->table.row_base = rows_before;
    if (deleted_total_out) *deleted_total_out = deleted_total;
    if (rows_total_out) *rows_total_out = rows_total;
    if (!local_ok) {
        set_error(local_error);
        return -2;
    }
    return 0;
}

}  // extern "C"
