// shard.cu -- a row-range sharded table: one process per GPU, full-scan SELECT with the ordered
// gather fused into the kernels over peer memory.
//
// What the reference does in MPI mode (engine/mpi/executeEngine-mpi.c:703-770): block partition of
// the rows, every rank evaluates its slice, MPI_Allreduce / MPI_Allgather of the per-rank counts,
// MPI_Allgatherv of the per-rank pieces in partition order.  Here, per query (no host collective,
// no NCCL call, no send/recv on the data path):
//
//   every rank      K1f scans its shard; the ids (already global: + first row of the shard) are
//                   stored by the compaction warps either into the rank's segment of the OWNER's
//                   result buffer -- peer memory mapped through CUDA IPC, so the ids cross NVLink
//                   as the kernel's own coalesced stores, during the scan -- or into the rank's own
//                   HBM (host result);
//   every rank      post_kernel: one system-scope release store of (epoch, match count) into
//                   EVERY rank's comm block (peer stores) = the count exchange;
//   owner           post_kernel: waits for all counts of this epoch (acquire loads of its own comm
//                   block), then moves the segments of ranks 1.. behind rank 0's ids -- rank 0's
//                   offset is always 0, so ITS scan stores straight into the dense result and is never
//                   moved -- giving one dense id list in partition order, which is table order;
//   host result     a host buffer shared by all ranks (POSIX shared memory, registered with CUDA in every
//                   process).  Up to 7 ranks: the first shard streams its ids out during its scan (its
//                   offset is 0), every other rank learns the counts of the lower ranks the same way and
//                   copies its ids to its exact offset over ITS OWN PCIe link.  From 8 ranks ("multipath"):
//                   the ids land in the owner's HBM as for a device result and every rank pulls 1/world of
//                   the packed list over NVLink and copies that out over its own link.
//   DELETE          qpe_shard_delete: local mask + stable compaction per shard, the new shard sizes
//                   all-gathered through the same comm blocks, shards renumbered.
//
// Slots are double buffered by epoch parity: no rank finishes query e before every rank has
// published its count of e, so a rank is never more than one query ahead of another.
// Every wait is bounded (trap, never a hung GPU).  NCCL / torch.distributed is only used by the
// caller to hand the IPC handles around at start-up.

#include "engine.cuh"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <mutex>

namespace qpe {

constexpr int kMaxRanks = 16;

struct ShardComm {                      // device memory of one rank, mapped by all the others
    unsigned long long count[2][kMaxRanks];  // [epoch parity][source rank] = epoch << 32 | match count
    unsigned long long packed[2];            // [epoch parity] = epoch whose dense result the owner has packed
};

struct ShardHostHeader {                // start of the shared host buffer
    volatile unsigned long long done[kMaxRanks][8];  // [rank][0] = epoch whose ids rank has delivered (64 B apart)
};

struct ShardState {
    int rank = 0, world = 1;
    ShardComm *comm[kMaxRanks] = {nullptr};  // [rank] = own allocation, others = IPC mappings
    uint32_t epoch = 0;
    // device result: segments in the owner's memory (own pointer on the owner, IPC mapping elsewhere)
    int owner = 0;
    // result memory (owner's; own pointer on the owner, IPC mapping elsewhere), per epoch parity:
    //   [ dense result: world x seg_cap ids | segments of ranks 1 .. world-1: seg_cap ids each ]
    uint32_t *seg_base = nullptr;
    uint64_t seg_cap = 0;
    uint32_t last_parity = 0;           // parity of the most recent device-result query
    int multipath = -1;                 // host result over every rank's PCIe link: -1 = auto (world >= 8), 0, 1
    // host result
    void *host_map = nullptr;           // shared mapping: ShardHostHeader, then the ids
    size_t host_bytes = 0;
    uint64_t host_cap = 0;              // ids
    char host_name[96] = {0};
    bool host_creator = false;
    // per-query counts as seen by this rank (mapped pinned host memory, written by post_kernel)
    unsigned long long *h_counts = nullptr;
    unsigned long long *d_counts = nullptr;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// wait until the slot carries `epoch`; bounded (about 20 s), then trap
__device__ __forceinline__ unsigned long long wait_slot(const unsigned long long *slot, uint32_t epoch) {
    unsigned long long v = ld_acquire_sys(slot);
    uint32_t spins = 0;
    while (static_cast<uint32_t>(v >> 32) != epoch) {
        __nanosleep(200);
        if (++spins == 100000000u) __trap();
        v = ld_acquire_sys(slot);
    }
    return v;
}

struct PeerPtrs {
    ShardComm *comm[kMaxRanks];
};

// ids per parity set: the dense area, then one segment per rank >= 1
__host__ __device__ inline unsigned long long set_ids(int world, unsigned long long seg_cap) {
    return static_cast<unsigned long long>(2 * world - 1) * seg_cap;
}

// The ONE kernel that follows the scan on every rank (same stream):
//   1. count exchange: thread r of CTA 0 stores (epoch, this rank's match count) into rank r's comm block;
//   2. every CTA waits (thread 0, acquire loads of this rank's own comm block) for all counts of the epoch;
//      CTA 0 also hands them to the host through mapped pinned memory;
//   3. pack != 0 (owner, device result): the segments of ranks 1.. are moved to their offsets in the dense
//      result, 4 independent loads in flight per thread (the move is latency-bound otherwise).
__global__ void __launch_bounds__(256) post_kernel(const unsigned long long *count, PeerPtrs peers, int rank, int world,
                                                   uint32_t epoch, unsigned long long *host_counts, int pack,
                                                   uint32_t *set, unsigned long long seg_cap) {
    __shared__ unsigned long long s_off[kMaxRanks + 1];
    const unsigned long long my = *reinterpret_cast<const volatile unsigned long long *>(count);
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();  // this rank's id stores (peer memory included) are visible before the count is
        st_release_sys(&peers.comm[threadIdx.x]->count[epoch & 1u][rank],
                       (static_cast<unsigned long long>(epoch) << 32) | (my & 0xffffffffull));
    }
    if (threadIdx.x < world) {
        const int r = threadIdx.x;
        const unsigned long long c =
            (r == rank) ? (my & 0xffffffffull) : (wait_slot(&peers.comm[rank]->count[epoch & 1u][r], epoch) & 0xffffffffull);
        s_off[r + 1] = c;
        if (blockIdx.x == 0) host_counts[r] = c;
    }
    __syncthreads();
    if (!pack) return;
    if (threadIdx.x == 0) {
        unsigned long long o = 0;
        for (int r = 0; r < world; ++r) {
            const unsigned long long c = s_off[r + 1];
            s_off[r] = o;
            o += c;
        }
        s_off[world] = o;
    }
    __syncthreads();
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    const unsigned long long t0 = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
    for (int r = 1; r < world; ++r) {
        const unsigned long long n = s_off[r + 1] - s_off[r];
        if (n > seg_cap || s_off[r] + n > static_cast<unsigned long long>(world) * seg_cap) continue;  // host reports it
        const uint32_t *__restrict__ src = set + (static_cast<unsigned long long>(world) + (r - 1)) * seg_cap;
        uint32_t *__restrict__ dst = set + s_off[r];
        unsigned long long i = t0;
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
}

// A plain all-gather of one 32-bit value per rank through the comm blocks (same slots, same epoch protocol as
// the post-scan kernel): thread r stores (epoch, value) into rank r's block, then waits for rank r's value.
__global__ void exchange_kernel(PeerPtrs peers, int rank, int world, uint32_t epoch, unsigned long long value,
                                unsigned long long *host_values) {
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    st_release_sys(&peers.comm[r]->count[epoch & 1u][rank],
                   (static_cast<unsigned long long>(epoch) << 32) | (value & 0xffffffffull));
    host_values[r] = (r == rank) ? (value & 0xffffffffull)
                                 : (wait_slot(&peers.comm[rank]->count[epoch & 1u][r], epoch) & 0xffffffffull);
}

// owner, after its pack: tell every rank that the dense result of this epoch is complete
__global__ void packed_flag_kernel(PeerPtrs peers, int world, uint32_t epoch) {
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    st_release_sys(&peers.comm[r]->packed[epoch & 1u], static_cast<unsigned long long>(epoch));
}
// every other rank: wait for it (bounded) before reading the owner's memory
__global__ void packed_wait_kernel(const ShardComm *mine, uint32_t epoch) {
    if (threadIdx.x != 0) return;
    uint32_t spins = 0;
    while (ld_acquire_sys(&mine->packed[epoch & 1u]) != static_cast<unsigned long long>(epoch)) {
        __nanosleep(200);
        if (++spins == 100000000u) __trap();
    }
}

// this rank's slice of the packed result: owner's HBM -> (NVLink) -> local HBM, 4 independent loads in flight per
// thread; the copy engine then takes it to the host over this GPU's own PCIe link.  (A cudaMemcpy straight from the
// peer mapping to the host measured 12 GB/s per rank.)
__global__ void __launch_bounds__(256) pull_slice_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst,
                                                         unsigned long long n) {
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

static ShardState *shard_of(GpuEngine *g) { return static_cast<ShardState *>(g->shard); }

void shard_destroy(GpuEngine *g) {
    ShardState *s = shard_of(g);
    if (!s) return;
    cudaSetDevice(g->device);
    cudaStreamSynchronize(g->stream);
    for (int r = 0; r < s->world; ++r) {
        if (!s->comm[r]) continue;
        if (r == s->rank)
            cudaFree(s->comm[r]);
        else
            cudaIpcCloseMemHandle(s->comm[r]);
    }
    if (s->h_counts) cudaFreeHost(s->h_counts);
    if (s->host_map) {
        cudaHostUnregister(s->host_map);
        munmap(s->host_map, s->host_bytes);
        if (s->host_creator) shm_unlink(s->host_name);
    }
    delete s;
    g->shard = nullptr;
}

}  // namespace qpe

using namespace qpe;

extern "C" {

int qpe_shard_init(struct engineS *engine, int rank, int world, unsigned char comm_handle_out[64]) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    GpuEngine *g = as_engine(engine);
    if (!g) return -1;
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) {
        set_error("qpe_shard_init: rank / world out of range (at most 16 ranks)");
        return -5;
    }
    if (g->shard) shard_destroy(g);
    cudaSetDevice(g->device);
    ShardState *s = new ShardState();
    s->rank = rank;
    s->world = world;
    bool ok = cuda_ok(cudaMalloc(&s->comm[rank], sizeof(ShardComm)), "cudaMalloc comm") &&
              cuda_ok(cudaMemset(s->comm[rank], 0, sizeof(ShardComm)), "cudaMemset comm") &&
              cuda_ok(cudaHostAlloc(&s->h_counts, sizeof(unsigned long long) * kMaxRanks, cudaHostAllocMapped),
                      "cudaHostAlloc counts") &&
              cuda_ok(cudaHostGetDevicePointer(&s->d_counts, s->h_counts, 0), "cudaHostGetDevicePointer");
    cudaIpcMemHandle_t h;
    ok = ok && cuda_ok(cudaIpcGetMemHandle(&h, s->comm[rank]), "cudaIpcGetMemHandle");
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
    std::lock_guard<std::mutex> lk(g_api_mutex);
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
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, all_handles + 64 * r, 64);
        void *p = nullptr;
        if (!cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle comm")) return -4;
        s->comm[r] = static_cast<ShardComm *>(p);
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
    std::lock_guard<std::mutex> lk(g_api_mutex);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s) {
        set_error("qpe_shard_set_device_result: call qpe_shard_init first");
        return -1;
    }
    s->owner = owner_rank;
    s->seg_base = segments;
    s->seg_cap = segment_capacity;
    return 0;
}

/* Host result: a buffer of `capacity` ids shared by all ranks of the box.  create != 0 on exactly one
 * rank (it creates /dev/shm/<name>), the others open it afterwards.  Returns the host pointer of the
 * id array in this process (NULL on failure). */
unsigned int *qpe_shard_open_host_result(struct engineS *engine, const char *name, unsigned long long capacity,
                                         int create) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s || !name || std::strlen(name) >= sizeof(s->host_name)) {
        set_error("qpe_shard_open_host_result: call qpe_shard_init first / bad name");
        return nullptr;
    }
    cudaSetDevice(g->device);
    if (s->host_map) {  // opened before: release the old buffer first
        cudaStreamSynchronize(g->stream);
        cudaHostUnregister(s->host_map);
        munmap(s->host_map, s->host_bytes);
        if (s->host_creator) shm_unlink(s->host_name);
        s->host_map = nullptr;
        s->host_bytes = 0;
        s->host_cap = 0;
        s->host_creator = false;
    }
    const size_t bytes = sizeof(ShardHostHeader) + sizeof(uint32_t) * (capacity + 16);
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
    if (!cuda_ok(cudaHostRegister(p, bytes, cudaHostRegisterPortable), "cudaHostRegister shared result")) {
        munmap(p, bytes);
        return nullptr;
    }
    s->host_map = p;
    s->host_bytes = bytes;
    s->host_cap = capacity;
    s->host_creator = create != 0;
    std::strncpy(s->host_name, name, sizeof(s->host_name) - 1);
    return reinterpret_cast<unsigned int *>(static_cast<uint8_t *>(p) + sizeof(ShardHostHeader));
}

/* Host result path: -1 = automatic (every rank's PCIe link from 8 ranks up), 0 = the first shard streams during its
 * scan and the others copy afterwards, 1 = always over every link.  Every rank must choose the same. */
int qpe_shard_set_multipath(struct engineS *engine, int mode) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s || mode < -1 || mode > 1) {
        set_error("qpe_shard_set_multipath: call qpe_shard_init first; mode is -1, 0 or 1");
        return -1;
    }
    s->multipath = mode;
    return 0;
}

/* Creator only, once every rank has opened the shared host buffer: remove its name from /dev/shm (the mappings
 * stay valid), so that nothing is left behind if a process dies. */
int qpe_shard_unlink_host_result(struct engineS *engine) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
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
    return (s && s->seg_base) ? s->seg_base + s->last_parity * set_ids(s->world, s->seg_cap) : nullptr;
}

void qpe_shard_close(struct engineS *engine) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    GpuEngine *g = as_engine(engine);
    if (g) shard_destroy(g);
}

/* One sharded full-scan SELECT; every rank of the group calls it with the same statement, in the same
 * order.  to_host == 0: the ids end up packed in the owner's HBM (qpe_shard_device_result);
 * to_host != 0: in the shared host buffer.  counts_out[world] = per-rank match counts (partition
 * order); the result is their concatenation = table order.  Returns -5 if a rank's ids did not fit. */
int qpe_shard_select(struct engineS *engine, struct whereClauseS *whereClause, int to_host,
                     unsigned long long *counts_out, qpe_scan_stats *stats) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s) {
        set_error("qpe_shard_select: call qpe_shard_init / qpe_shard_connect first");
        return -1;
    }
    if (to_host ? !s->host_map : !s->seg_base) {
        set_error(to_host ? "qpe_shard_select: no host result buffer" : "qpe_shard_select: no device result segments");
        return -1;
    }
    if (g->table.row_base + static_cast<uint64_t>(g->table.n) > 0xffffffffull) {
        set_error("global row ids do not fit 32 bits");
        return -5;
    }
    cudaSetDevice(g->device);
    if (++s->epoch == 0) s->epoch = 1;
    const uint32_t epoch = s->epoch;
    PeerPtrs peers{};
    for (int r = 0; r < s->world; ++r) peers.comm[r] = s->comm[r];

    // Host result over EVERY rank's PCIe link ("multipath"): the ids first land in the owner's HBM exactly as for
    // a device result, then each rank copies 1/world of the packed list (read over NVLink from the owner) to the
    // shared host buffer over its own link.  It pays when one link's copy of the whole result far outlasts a shard's
    // scan.  Measured (1 B rows, 9.45 M ids = 38 MB, all from the first shard): 8 ranks 0.845 -> 0.748 ms per query;
    // 4 ranks 0.849 -> 1.07 ms (the small per-rank copies reach only ~13-20 GB/s each, and the first shard's
    // streaming copy hides most of itself under a 0.56 ms scan), hence automatic from 8 ranks up.  Otherwise the
    // first shard streams its ids out during the scan and the others copy theirs afterwards.
    const bool multi = to_host && s->world > 1 && s->seg_base != nullptr &&
                       (s->multipath == 1 || (s->multipath < 0 && s->world >= 8));
    const bool to_segments = !to_host || multi;
    // the kernels that follow the scan on the engine's stream, before its single synchronisation
    uint32_t *set = to_segments ? s->seg_base + (epoch & 1u) * set_ids(s->world, s->seg_cap) : nullptr;
    const int pack = (to_segments && s->rank == s->owner && s->world > 1) ? 1 : 0;
    g->post_match = [&]() -> bool {
        post_kernel<<<pack ? 148 * 8 : 1, 256, 0, g->stream>>>(g->count_dev, peers, s->rank, s->world, epoch, s->d_counts,
                                                               pack, set, s->seg_cap);
        if (multi) {
            if (s->rank == s->owner)
                packed_flag_kernel<<<1, 32, 0, g->stream>>>(peers, s->world, epoch);
            else
                packed_wait_kernel<<<1, 32, 0, g->stream>>>(s->comm[s->rank], epoch);
        }
        return cuda_ok(cudaGetLastError(), "shard post-scan kernel launch");
    };
    if (to_segments) {
        // the first shard's offset in the result is always 0: its scan writes the dense result in place
        g->out_override = s->rank == 0 ? set : set + (static_cast<size_t>(s->world) + (s->rank - 1)) * s->seg_cap;
        g->out_override_cap = s->seg_cap;
        s->last_parity = epoch & 1u;
    }
    uint32_t *host_ids =
        to_host ? reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(s->host_map) + sizeof(ShardHostHeader)) : nullptr;
    if (to_host && !multi && s->rank == 0) {
        // the first shard's offset is always 0: its ids are copied out segment by segment DURING the scan
        g->host_out = host_ids;
        g->host_out_cap = s->host_cap;
    }
    g->id_base_override = static_cast<uint32_t>(g->table.row_base);
    g->id_base_always = true;
    g->count_mapped = &s->h_counts[s->rank];  // the post kernel hands the count over: no separate download
    uint64_t m = 0;
    const bool ok = engine_match(g, whereClause, true, false, false, false, &m);
    g->post_match = nullptr;
    g->count_mapped = nullptr;
    const bool delivered = g->host_out != nullptr && g->host_out_done;
    g->host_out = nullptr;
    g->host_out_cap = 0;
    g->out_override = nullptr;
    g->out_override_cap = 0;
    g->id_base_override = 0;
    g->id_base_always = false;
    if (!ok) return -2;
    if (g->last.path != 0 || (g->table.n > 0 && g->last.tile_rows == 0)) {
        // every rank takes the same branch (same statement, same column widths), so nobody is left waiting
        set_error("qpe_shard_select: this WHERE cannot be staged by the scan kernel (too wide for shared memory)");
        return -6;
    }
    g->last.launches += multi ? 2 : 1;

    int rc = 0;
    unsigned long long before = 0, total = 0;
    for (int r = 0; r < s->world; ++r) {
        const unsigned long long c = s->h_counts[r];
        if (counts_out) counts_out[r] = c;
        if (r < s->rank) before += c;
        total += c;
        if (to_segments && c > s->seg_cap) rc = -5;
    }
    if (to_host) {
        ShardHostHeader *hh = static_cast<ShardHostHeader *>(s->host_map);
        if (multi) {
            // this rank's 1/world of the packed result: owner's HBM -> (NVLink) -> this GPU -> (its PCIe) -> host
            const unsigned long long lo = total * s->rank / s->world, hi = total * (s->rank + 1) / s->world;
            if (total > s->host_cap) rc = -5;
            if (rc == 0 && hi > lo) {
                const uint32_t *src = set + lo;
                if (s->rank != s->owner) {
                    if (!engine_ensure_ids(g, static_cast<int64_t>(hi - lo))) return -4;
                    pull_slice_kernel<<<148 * 4, 256, 0, g->stream>>>(set + lo, g->d_ids, hi - lo);
                    if (!cuda_ok(cudaGetLastError(), "shard slice kernel launch")) return -4;
                    g->last.launches += 1;
                    src = g->d_ids;
                }
                if (!cuda_ok(cudaMemcpyAsync(host_ids + lo, src, (hi - lo) * 4, cudaMemcpyDeviceToHost, g->stream),
                             "download ids") ||
                    !cuda_ok(cudaStreamSynchronize(g->stream), "download ids"))
                    return -4;
            }
        } else if (before + m > s->host_cap) {
            rc = -5;
        } else if (m && !delivered) {
            if (!cuda_ok(cudaMemcpyAsync(host_ids + before, g->d_ids, m * 4, cudaMemcpyDeviceToHost, g->stream),
                         "download ids") ||
                !cuda_ok(cudaStreamSynchronize(g->stream), "download ids"))
                return -4;
        }
        __atomic_store_n(&hh->done[s->rank][0], static_cast<unsigned long long>(epoch), __ATOMIC_RELEASE);
        if (s->rank == s->owner) {
            // the result is complete when every rank has delivered its piece
            for (int r = 0; r < s->world; ++r) {
                unsigned long long spins = 0;
                while (__atomic_load_n(&hh->done[r][0], __ATOMIC_ACQUIRE) != epoch) {
                    if (++spins > 4000000000ull) {
                        set_error("qpe_shard_select: a rank never delivered its ids");
                        return -4;
                    }
                }
            }
        }
    }
    if (rc == -5) set_error("a rank's ids did not fit the result buffer (nothing was written past it)");
    if (stats) qpe_gpu_last_stats(engine, stats);
    return rc;
}

/* DELETE on a sharded table (the MPI engine's executeQueryDeleteMPI, engine/mpi/executeEngine-mpi.c:703-770:
 * local match flags, MPI_Allreduce of the count, renumbering): every rank deletes the matching rows of ITS shard
 * (K1 with the inverted program + stable column compaction, as executeQueryDeleteGPU), then the new shard sizes
 * and the deleted counts are all-gathered through the comm blocks and every shard's first global row becomes
 * the sum of the lower shards' new sizes -- global row ids stay positions in the whole table, in table order.
 * Every rank calls it with the same statement.  No data file is touched (a sharded table has none). */
int qpe_shard_delete(struct engineS *engine, struct whereClauseS *whereClause, unsigned long long *deleted_total_out,
                     unsigned long long *rows_total_out) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    GpuEngine *g = as_engine(engine);
    ShardState *s = g ? shard_of(g) : nullptr;
    if (!s) {
        set_error("qpe_shard_delete: call qpe_shard_init / qpe_shard_connect first");
        return -1;
    }
    cudaSetDevice(g->device);
    int64_t deleted = 0;
    if (!engine_delete(g, whereClause, &deleted)) return -2;
    PeerPtrs peers{};
    for (int r = 0; r < s->world; ++r) peers.comm[r] = s->comm[r];
    unsigned long long rows_before = 0, rows_total = 0, deleted_total = 0;
    const unsigned long long mine[2] = {static_cast<unsigned long long>(g->table.n), static_cast<unsigned long long>(deleted)};
    for (int pass = 0; pass < 2; ++pass) {
        if (++s->epoch == 0) s->epoch = 1;
        exchange_kernel<<<1, 32, 0, g->stream>>>(peers, s->rank, s->world, s->epoch, mine[pass], s->d_counts);
        if (!cuda_ok(cudaGetLastError(), "shard exchange kernel launch") ||
            !cuda_ok(cudaStreamSynchronize(g->stream), "shard exchange"))
            return -4;
        for (int r = 0; r < s->world; ++r) {
            const unsigned long long v = s->h_counts[r];
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
    return 0;
}

}  // extern "C"
