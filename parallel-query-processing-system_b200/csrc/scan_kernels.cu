// scan_kernels.cu -- SELECT/WHERE hot path for B200 (sm_100a)
//
// K1  scan_tma_kernel      full-table predicate evaluation + order-preserving compaction.
//       Replaces linearSearchRecords/evaluateWhereClause/checkCondition
//       (engine/serial/executeEngine-serial.c:854-878, :292-316, :251-289).
//       Warp-specialised persistent CTAs:
//         P  (1 warp, 1 lane)  claims tiles from an atomic counter and streams every referenced
//                              column's slice of the tile into shared memory with 1-D TMA bulk
//                              copies (cp.async.bulk + mbarrier complete_tx), S stages deep;
//         E  (kEvalWarps)      evaluate the compiled WHERE program on the staged tile, one
//                              R-bit match mask per lane, __ballot_sync -> tile bitmap in smem,
//                              publish the tile's match count (look-back "aggregate");
//         W  (kWriteWarps)     decoupled look-back over the tile descriptors to get the tile's
//                              global offset, then expand the bitmap into row ids with
//                              popc-ranked coalesced stores -- rows come out in table order.
//       E never waits on a look-back, so HBM streaming is not stalled by the scan chain.
// K1g filter_kernel        same program on a gathered candidate list (index path), same
//                          ordered compaction (single-pass, decoupled look-back).
// K2  gather_kernel        projection / column compaction gather.
//
// HBM-bound integer/byte work: no tensor cores by design.

#include "scan_kernels.cuh"

#include <cstdint>

namespace qpe {

// ------------------------------------------------------------------------------------------
// small PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ uint32_t lane_id() {
    uint32_t l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

constexpr uint32_t kStateAgg = 1u;
constexpr uint32_t kStatePrefix = 2u;
__device__ __forceinline__ unsigned long long make_desc(uint32_t epoch, uint32_t state, uint32_t value) {
    return (static_cast<unsigned long long>((epoch << 2) | state) << 32) | value;
}

// ------------------------------------------------------------------------------------------
// comparison helpers.  Every leaf is a 3-way compare (lt / eq / gt) looked up in the leaf's
// 3-bit truth table, which encodes the six operators of create_where_condition
// (executeEngine-serial.c:129-213) and the "no comparator => false" case.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tt_bit(uint32_t tt, bool lt, bool eq) {
    const uint32_t idx = lt ? 0u : (eq ? 1u : 2u);
    return (tt >> idx) & 1u;
}

// unsigned-byte lexicographic compare of one 16-byte chunk (strcmp order on NUL-padded data)
// r: 0 = lt, 1 = eq so far, 2 = gt
__device__ __forceinline__ uint32_t cmp_chunk(uint32_t r, const uint4 v, const uint4 l) {
    const uint32_t a0 = bswap32(v.x), b0 = bswap32(l.x);
    const uint32_t a1 = bswap32(v.y), b1 = bswap32(l.y);
    const uint32_t a2 = bswap32(v.z), b2 = bswap32(l.z);
    const uint32_t a3 = bswap32(v.w), b3 = bswap32(l.w);
    const bool d0 = a0 != b0, d1 = a1 != b1, d2 = a2 != b2, d3 = a3 != b3;
    const bool lt = d0 ? (a0 < b0) : d1 ? (a1 < b1) : d2 ? (a2 < b2) : (a3 < b3);
    const bool any = d0 | d1 | d2 | d3;
    return (r == 1u && any) ? (lt ? 0u : 2u) : r;
}

// string compare of a row (nchunks x 16 B) against the padded literal; all 32 lanes must call
template <typename RowPtr>
__device__ __forceinline__ uint32_t str_cmp3(RowPtr row, const uint4 *lit, int nchunks) {
    uint32_t r = 1u;
    for (int k = 0; k < nchunks; ++k) {
        r = cmp_chunk(r, row[k], lit[k]);
        if (__all_sync(0xffffffffu, r != 1u)) break;  // warp-uniform early exit
    }
    return r;
}

// ------------------------------------------------------------------------------------------
// K1: TMA-staged scan
// ------------------------------------------------------------------------------------------
constexpr int kEvalWarps = 8;
constexpr int kWriteWarps = 4;
constexpr int kBmSlots = kWriteWarps;  // one bitmap slot per writer warp
constexpr int kMaxStages = 8;
constexpr int kScanThreads = 32 * (1 + kWriteWarps + kEvalWarps);
constexpr int kMaxTileRows = 8192;  // == kRowPad: a full tile is always inside the allocation
constexpr int kRowsPerGroup = 32 * kEvalWarps;  // tile_rows is a multiple of this

struct ScanParams {
    const uint8_t *col[NUM_COLS];
    uint32_t width[NUM_COLS];
    uint32_t smem_off[NUM_COLS];  // byte offset of the column inside a stage
    int32_t n_ref;
    int32_t ref_col[NUM_COLS];
    uint32_t stage_bytes;
    int32_t tile_rows;
    int32_t n_stages;
    uint32_t epoch;
    long long n_rows;
    long long n_tiles;
    QueryCtl *ctl;
    unsigned long long *tile_desc;
    uint32_t *out_ids;
    uint32_t *out_bitmap;
};

struct ScanSmemHeader {
    Program prog;
    alignas(8) uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t bm_full[kBmSlots];
    uint64_t bm_empty[kBmSlots];
    long long tile_of_stage[kMaxStages];
    long long tile_of_bm[kBmSlots];
    uint32_t agg[kBmSlots];
    alignas(16) uint32_t bm[kBmSlots][kMaxTileRows / 32];
};

// evaluate one leaf for the R rows of this lane (rows lrow, lrow+32, ...) out of a staged tile
__device__ __forceinline__ uint32_t eval_leaf_tile(const PLeaf &lf, const Program *sp, const uint8_t *stage,
                                                   const ScanParams &p, int lrow, int R) {
    uint32_t m = 0;
    const uint8_t *base = stage + p.smem_off[lf.col];
    const uint32_t tt = lf.tt;
    switch (lf.type) {
        case T_I32: {
            const int32_t *c = reinterpret_cast<const int32_t *>(base) + lrow;
            const int32_t lit = lf.lit_i32;
#pragma unroll 4
            for (int j = 0; j < R; ++j) {
                const int32_t v = c[j * 32];
                m |= tt_bit(tt, v < lit, v == lit) << j;
            }
            break;
        }
        case T_U64: {
            const unsigned long long *c = reinterpret_cast<const unsigned long long *>(base) + lrow;
            const unsigned long long lit = lf.lit_u64;
#pragma unroll 4
            for (int j = 0; j < R; ++j) {
                const unsigned long long v = c[j * 32];
                m |= tt_bit(tt, v < lit, v == lit) << j;
            }
            break;
        }
        case T_BOOL: {
            const uint8_t *c = base + lrow;
            const uint32_t lit = static_cast<uint32_t>(lf.lit_i32) & 1u;
#pragma unroll 4
            for (int j = 0; j < R; ++j) {
                const uint32_t v = c[j * 32] != 0 ? 1u : 0u;
                m |= tt_bit(tt, v < lit, v == lit) << j;
            }
            break;
        }
        default: {  // T_STR
            const uint32_t w = p.width[lf.col];
            const int nch = static_cast<int>(w >> 4);
            const uint4 *lit = reinterpret_cast<const uint4 *>(sp->lit_pool + lf.lit_off);
            const uint8_t *c = base + static_cast<size_t>(lrow) * w;
            for (int j = 0; j < R; ++j) {
                const uint4 *row = reinterpret_cast<const uint4 *>(c + static_cast<size_t>(j) * 32u * w);
                const uint32_t r = str_cmp3(row, lit, nch);
                m |= ((tt >> r) & 1u) << j;
            }
            break;
        }
    }
    return m;
}

// run the compiled WHERE program; returns the R-bit match mask of this lane's rows
template <typename LeafFn>
__device__ __forceinline__ uint32_t run_program(const Program *sp, uint32_t all_mask, LeafFn leaf_fn) {
    uint32_t acc = all_mask;  // empty program == NULL where clause == every row matches
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    const int n = sp->n_instr;
    for (int i = 0; i < n; ++i) {
        const PInstr in = sp->instr[i];
        switch (in.op) {
            case P_LEAF_SET: acc = leaf_fn(sp->leaf[in.arg]); break;
            case P_LEAF_AND: acc &= leaf_fn(sp->leaf[in.arg]); break;
            case P_LEAF_OR: acc |= leaf_fn(sp->leaf[in.arg]); break;
            case P_PUSH:
                switch (in.arg) {
                    case 0: s0 = acc; break;
                    case 1: s1 = acc; break;
                    case 2: s2 = acc; break;
                    case 3: s3 = acc; break;
                    case 4: s4 = acc; break;
                    case 5: s5 = acc; break;
                    case 6: s6 = acc; break;
                    default: s7 = acc; break;
                }
                break;
            case P_POP_AND:
            case P_POP_OR: {
                uint32_t v;
                switch (in.arg) {
                    case 0: v = s0; break;
                    case 1: v = s1; break;
                    case 2: v = s2; break;
                    case 3: v = s3; break;
                    case 4: v = s4; break;
                    case 5: v = s5; break;
                    case 6: v = s6; break;
                    default: v = s7; break;
                }
                acc = (in.op == P_POP_AND) ? (acc & v) : (acc | v);
                break;
            }
            case P_CONST: acc = in.arg ? all_mask : 0u; break;
            case P_NOT: acc = ~acc & all_mask; break;
            default: break;
        }
    }
    return acc & all_mask;
}

// warp-wide decoupled look-back: exclusive prefix of `tile` (sum of the aggregates of all
// earlier tiles).  Lane l inspects tile-1-l; windows of 32 predecessors until a PREFIX is found.
__device__ __forceinline__ uint32_t warp_lookback(const unsigned long long *desc, long long tile, uint32_t epoch,
                                                  uint32_t lane) {
    uint32_t excl = 0;
    long long idx = tile - 1;
    while (true) {
        const long long mine = idx - static_cast<long long>(lane);
        uint32_t state = kStatePrefix, val = 0;
        if (mine >= 0) {
            unsigned long long d;
            uint32_t flag;
            do {
                d = ld_desc(desc + mine);
                flag = static_cast<uint32_t>(d >> 32);
            } while ((flag >> 2) != epoch);
            state = flag & 3u;
            val = static_cast<uint32_t>(d);
        }
        const uint32_t pm = __ballot_sync(0xffffffffu, state == kStatePrefix);
        uint32_t contrib = val;
        if (pm) {
            const uint32_t first = static_cast<uint32_t>(__ffs(pm) - 1);
            if (lane > first) contrib = 0;
        }
        contrib = __reduce_add_sync(0xffffffffu, contrib);
        excl += contrib;
        if (pm) break;
        idx -= 32;
    }
    return excl;
}

// expand `nwords` bitmap words (32 rows each, row id of bit b in word w = row0 + 32*w + b) into
// ids at out[excl...], in row order.  Whole warp cooperates; stores are popc-ranked.
__device__ __forceinline__ void warp_expand_bitmap(const uint32_t *words, int nwords, long long row0, uint32_t *out,
                                                   uint32_t excl, uint32_t lane) {
    uint32_t running = excl;
    for (int wb = 0; wb < nwords; wb += 32) {
        const int wi = wb + static_cast<int>(lane);
        const uint32_t word = (wi < nwords) ? words[wi] : 0u;
        const uint32_t pc = __popc(word);
        // inclusive scan of pc over lanes
        uint32_t inc = pc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= static_cast<uint32_t>(d)) inc += t;
        }
        const uint32_t woff = running + inc - pc;
        uint32_t nz = __ballot_sync(0xffffffffu, word != 0u);
        while (nz) {
            const int src = __ffs(nz) - 1;
            nz &= nz - 1;
            const uint32_t w = __shfl_sync(0xffffffffu, word, src);
            const uint32_t o = __shfl_sync(0xffffffffu, woff, src);
            if ((w >> lane) & 1u) {
                out[o + __popc(w & lanemask_lt())] =
                    static_cast<uint32_t>(row0 + 32ll * (wb + src) + static_cast<long long>(lane));
            }
        }
        running += __shfl_sync(0xffffffffu, inc, 31);
    }
}

__global__ void __launch_bounds__(kScanThreads, 1) scan_tma_kernel(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ScanSmemHeader *sh = reinterpret_cast<ScanSmemHeader *>(smem_raw);
    uint8_t *stages = smem_raw + ((sizeof(ScanSmemHeader) + 127) & ~size_t(127));

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31u;
    const int S = p.n_stages;
    const int T = p.tile_rows;
    const int R = T / kRowsPerGroup;   // rows per lane per tile (<= 32)
    const int WPT = T >> 5;            // bitmap words per tile

    // program -> shared memory (uniform reads afterwards), barrier init
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&p.ctl->prog);
        uint4 *dst = reinterpret_cast<uint4 *>(&sh->prog);
        for (uint32_t i = tid; i < sizeof(Program) / 16; i += kScanThreads) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&sh->full[s], 1);
            mbar_init(&sh->empty[s], kEvalWarps);
        }
        for (int b = 0; b < kBmSlots; ++b) {
            mbar_init(&sh->bm_full[b], kEvalWarps);
            mbar_init(&sh->bm_empty[b], 1);
            sh->agg[b] = 0;
        }
        fence_mbar_init();
    }
    __syncthreads();
    const Program *sp = &sh->prog;

    if (warp == 0) {
        // ===== P: tile claim + TMA producer =====
        if (lane == 0) {
            for (long long k = 0;; ++k) {
                const int s = static_cast<int>(k % S);
                const uint32_t u = static_cast<uint32_t>(k / S);
                mbar_wait(&sh->empty[s], (u & 1u) ^ 1u);
                const long long tile = static_cast<long long>(atomicAdd(&p.ctl->tile_counter, 1u));
                if (tile >= p.n_tiles) {
                    sh->tile_of_stage[s] = -1;
                    mbar_arrive(&sh->full[s]);
                    break;
                }
                sh->tile_of_stage[s] = tile;
                mbar_arrive_expect_tx(&sh->full[s], p.stage_bytes);
                uint8_t *dst = stages + static_cast<size_t>(s) * p.stage_bytes;
                for (int r = 0; r < p.n_ref; ++r) {
                    const int c = p.ref_col[r];
                    const uint32_t w = p.width[c];
                    tma_bulk_g2s(dst + p.smem_off[c], p.col[c] + static_cast<size_t>(tile) * T * w,
                                 static_cast<uint32_t>(T) * w, &sh->full[s]);
                }
            }
        }
    } else if (warp <= kWriteWarps) {
        // ===== W: look-back + ordered id expansion =====
        const int w = static_cast<int>(warp) - 1;  // == bitmap slot
        for (long long k = w;; k += kWriteWarps) {
            const uint32_t u = static_cast<uint32_t>(k / kBmSlots);
            mbar_wait(&sh->bm_full[w], u & 1u);
            const long long tile = sh->tile_of_bm[w];
            if (tile < 0) break;
            const uint32_t total = sh->agg[w] & 0xffffffu;
            const uint32_t excl = warp_lookback(p.tile_desc, tile, p.epoch, lane);
            if (lane == 0) {
                st_desc(p.tile_desc + tile, make_desc(p.epoch, kStatePrefix, excl + total));
                if (tile == p.n_tiles - 1) p.ctl->out_count = static_cast<unsigned long long>(excl) + total;
            }
            const uint32_t *words = sh->bm[w];
            if (p.out_bitmap) {
                uint32_t *dstw = p.out_bitmap + tile * WPT;
                for (int i = static_cast<int>(lane); i < WPT; i += 32) dstw[i] = words[i];
            }
            if (p.out_ids && total) warp_expand_bitmap(words, WPT, tile * T, p.out_ids, excl, lane);
            __syncwarp();
            if (lane == 0) {
                sh->agg[w] = 0;
                mbar_arrive(&sh->bm_empty[w]);
            }
        }
    } else {
        // ===== E: predicate evaluation =====
        const int ew = static_cast<int>(warp) - 1 - kWriteWarps;
        const uint32_t all_mask = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
        const int lrow = ew * (32 * R) + static_cast<int>(lane);  // first row of this lane inside a tile
        long long k = 0;
        for (;; ++k) {
            const int s = static_cast<int>(k % S);
            const uint32_t us = static_cast<uint32_t>(k / S);
            mbar_wait(&sh->full[s], us & 1u);
            const long long tile = sh->tile_of_stage[s];
            if (tile < 0) break;
            const int b = static_cast<int>(k % kBmSlots);
            const uint32_t ub = static_cast<uint32_t>(k / kBmSlots);
            mbar_wait(&sh->bm_empty[b], (ub & 1u) ^ 1u);

            const uint8_t *stage = stages + static_cast<size_t>(s) * p.stage_bytes;
            const uint32_t acc = run_program(sp, all_mask, [&](const PLeaf &lf) {
                return eval_leaf_tile(lf, sp, stage, p, lrow, R);
            });

            const long long row_base = tile * T + lrow;  // global row of this lane's j = 0
            uint32_t myword = 0, cnt = 0;
            for (int j = 0; j < R; ++j) {
                const bool hit = ((acc >> j) & 1u) && (row_base + 32ll * j < p.n_rows);
                const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                if (static_cast<int>(lane) == j) myword = bal;
                cnt += __popc(bal);
            }
            if (static_cast<int>(lane) < R) sh->bm[b][ew * R + lane] = myword;
            __syncwarp();
            if (lane == 0) {
                const uint32_t old = atomicAdd(&sh->agg[b], (1u << 24) | cnt);
                if ((old >> 24) == kEvalWarps - 1) {  // last evaluator of this tile: publish aggregate
                    const uint32_t total = (old & 0xffffffu) + cnt;
                    sh->tile_of_bm[b] = tile;
                    st_desc(p.tile_desc + tile, make_desc(p.epoch, tile == 0 ? kStatePrefix : kStateAgg, total));
                }
                mbar_arrive(&sh->bm_full[b]);
                mbar_arrive(&sh->empty[s]);
            }
        }
        // termination: hand one sentinel to every writer warp (next kWriteWarps slots in sequence)
        for (int t = 0; t < kWriteWarps; ++t, ++k) {
            const int b = static_cast<int>(k % kBmSlots);
            const uint32_t ub = static_cast<uint32_t>(k / kBmSlots);
            mbar_wait(&sh->bm_empty[b], (ub & 1u) ^ 1u);
            if (lane == 0) {
                if (ew == 0) sh->tile_of_bm[b] = -1;
                mbar_arrive(&sh->bm_full[b]);
            }
        }
    }
}

bool scan_plan(const DevTable &t, const Program &prog, int force_tile_rows, int force_stages, ScanGeometry *geo,
               const char **why) {
    int dev = 0;
    cudaGetDevice(&dev);
    int max_smem = 0, n_sm = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (max_smem <= 0) max_smem = 227 * 1024;
    if (n_sm <= 0) n_sm = 148;

    int64_t bpr = 0;
    for (int c = 0; c < NUM_COLS; ++c)
        if (prog.col_mask & (1u << c)) {
            if (!t.resident(c)) {
                if (why) *why = "WHERE references a column that is not resident on the device";
                return false;
            }
            bpr += t.col[c].width;
        }
    const size_t header = (sizeof(ScanSmemHeader) + 127) & ~size_t(127);
    const size_t budget = static_cast<size_t>(max_smem) - header - 1024;

    int T = force_tile_rows;
    int S = force_stages;
    if (bpr == 0) {  // no column referenced (NULL where / constants): nothing to stage
        T = T ? T : 4096;
        S = S ? S : 2;
    } else {
        if (!T) {
            // aim at ~48 KB per stage, 4 stages; shrink for wide rows
            int64_t rows = (48 * 1024) / bpr;
            rows = (rows / kRowsPerGroup) * kRowsPerGroup;
            if (rows < kRowsPerGroup) rows = kRowsPerGroup;
            if (rows > kMaxTileRows) rows = kMaxTileRows;
            T = static_cast<int>(rows);
        }
        if (!S) {
            S = 4;
            while (S > 2 && static_cast<size_t>(S) * (static_cast<size_t>(T) * bpr + 128 * NUM_COLS) > budget) --S;
        }
    }
    if (T % kRowsPerGroup != 0 || T > kMaxTileRows || T <= 0 || S < 1 || S > kMaxStages) {
        if (why) *why = "invalid tile geometry";
        return false;
    }
    // stage layout: each referenced column 128-byte aligned
    size_t stage_bytes = 0;
    for (int c = 0; c < NUM_COLS; ++c)
        if (prog.col_mask & (1u << c)) stage_bytes += (static_cast<size_t>(T) * t.col[c].width + 127) & ~size_t(127);
    if (stage_bytes * S > budget || stage_bytes >= (1u << 20)) {
        if (why) *why = "row too wide to stage a tile in shared memory";
        return false;
    }
    geo->tile_rows = T;
    geo->stages = S;
    geo->n_tiles = (t.n + T - 1) / T;
    geo->bytes_per_row = bpr;
    geo->smem_bytes = header + stage_bytes * S + 128;
    int64_t grid = geo->n_tiles < n_sm ? geo->n_tiles : n_sm;
    if (grid < 1) grid = 1;
    geo->grid = static_cast<int>(grid);
    return true;
}

cudaError_t scan_launch(const ScanLaunch &L, const ScanGeometry &geo, cudaStream_t stream) {
    ScanParams p{};
    const DevTable &t = *L.table;
    size_t off = 0;
    p.n_ref = 0;
    for (int c = 0; c < NUM_COLS; ++c) {
        p.col[c] = t.col[c].d;
        p.width[c] = t.col[c].width;
        p.smem_off[c] = 0;
        if (L.h_prog->col_mask & (1u << c)) {
            p.ref_col[p.n_ref++] = c;
            p.smem_off[c] = static_cast<uint32_t>(off);
            off += (static_cast<size_t>(geo.tile_rows) * t.col[c].width + 127) & ~size_t(127);
        }
    }
    // expect_tx counts the bytes actually copied (unpadded)
    uint32_t tx = 0;
    for (int r = 0; r < p.n_ref; ++r) tx += static_cast<uint32_t>(geo.tile_rows) * p.width[p.ref_col[r]];
    // stage stride must cover the padded layout; tx is what the barrier waits for
    p.stage_bytes = static_cast<uint32_t>(off);
    p.tile_rows = geo.tile_rows;
    p.n_stages = geo.stages;
    p.epoch = L.epoch;
    p.n_rows = t.n;
    p.n_tiles = geo.n_tiles;
    p.ctl = const_cast<QueryCtl *>(L.d_ctl);
    p.tile_desc = L.tile_desc;
    p.out_ids = L.out_ids;
    p.out_bitmap = L.out_bitmap;
    // the kernel arms each barrier with stage_bytes: make them equal by construction
    if (tx != p.stage_bytes) {
        // padded layout differs from copied bytes: pass tx through a second field
        // (columns are 128-byte multiples whenever tile_rows*width is, which holds for
        // tile_rows % 256 == 0 and width in {1,4,8,16k}) -- so this cannot happen.
        return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaFuncSetAttribute(scan_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(geo.smem_bytes));
    if (e != cudaSuccess) return e;
    scan_tma_kernel<<<geo.grid, kScanThreads, geo.smem_bytes, stream>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K1g: candidate filter (index path) -- gathered rows, ordered single-pass compaction
// ------------------------------------------------------------------------------------------
constexpr int kFilterThreads = 256;
constexpr int kFilterItems = 4;
constexpr int kFilterTile = kFilterThreads * kFilterItems;  // 1024 candidates per tile

int64_t filter_tiles(long long n) { return (n + kFilterTile - 1) / kFilterTile; }

struct FilterParams {
    const uint8_t *col[NUM_COLS];
    uint32_t width[NUM_COLS];
    CandSegments segs;
    long long n_cand;
    long long n_tiles;
    QueryCtl *ctl;
    unsigned long long *tile_desc;
    uint32_t epoch;
    uint32_t *out_ids;
};

__device__ __forceinline__ uint32_t eval_leaf_row(const PLeaf &lf, const Program *sp, const FilterParams &p,
                                                  uint32_t row) {
    const uint32_t tt = lf.tt;
    const uint8_t *base = p.col[lf.col];
    switch (lf.type) {
        case T_I32: {
            const int32_t v = __ldg(reinterpret_cast<const int32_t *>(base) + row);
            return tt_bit(tt, v < lf.lit_i32, v == lf.lit_i32);
        }
        case T_U64: {
            const unsigned long long v = __ldg(reinterpret_cast<const unsigned long long *>(base) + row);
            return tt_bit(tt, v < lf.lit_u64, v == lf.lit_u64);
        }
        case T_BOOL: {
            const uint32_t v = __ldg(base + row) != 0 ? 1u : 0u;
            const uint32_t lit = static_cast<uint32_t>(lf.lit_i32) & 1u;
            return tt_bit(tt, v < lit, v == lit);
        }
        default: {
            const uint32_t w = p.width[lf.col];
            const uint4 *rp = reinterpret_cast<const uint4 *>(base + static_cast<size_t>(row) * w);
            const uint4 *lit = reinterpret_cast<const uint4 *>(sp->lit_pool + lf.lit_off);
            uint32_t r = 1u;
            const int nch = static_cast<int>(w >> 4);
            for (int k = 0; k < nch; ++k) {
                r = cmp_chunk(r, __ldg(rp + k), lit[k]);
                if (__all_sync(0xffffffffu, r != 1u)) break;
            }
            return (tt >> r) & 1u;
        }
    }
}

__global__ void __launch_bounds__(kFilterThreads) filter_kernel(const __grid_constant__ FilterParams p) {
    __shared__ Program s_prog;
    __shared__ long long s_tile;
    __shared__ uint32_t s_words[kFilterItems * (kFilterThreads / 32)];
    __shared__ uint32_t s_woff[kFilterItems * (kFilterThreads / 32)];
    __shared__ uint32_t s_total, s_excl;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&p.ctl->prog);
        uint4 *dst = reinterpret_cast<uint4 *>(&s_prog);
        for (uint32_t i = tid; i < sizeof(Program) / 16; i += kFilterThreads) dst[i] = src[i];
    }
    if (tid == 0) s_tile = static_cast<long long>(atomicAdd(&p.ctl->tile_counter, 1u));
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= p.n_tiles) return;  // cannot happen with grid == n_tiles; kept for safety
    const Program *sp = &s_prog;

    constexpr int kWarps = kFilterThreads / 32;
    uint32_t rows[kFilterItems];
#pragma unroll
    for (int it = 0; it < kFilterItems; ++it) {
        const long long i = tile * kFilterTile + it * kFilterThreads + tid;
        uint32_t row = 0;
        bool valid = i < p.n_cand;
        if (valid) {
            int sg = 0;
            while (sg + 1 < p.segs.n_seg && i >= p.segs.vstart[sg + 1]) ++sg;
            const long long k = p.segs.first[sg] + (i - p.segs.vstart[sg]);
            row = p.segs.perm[sg] ? __ldg(p.segs.perm[sg] + k) : static_cast<uint32_t>(k);
        }
        rows[it] = row;
        const uint32_t acc = run_program(sp, 1u, [&](const PLeaf &lf) { return eval_leaf_row(lf, sp, p, row); });
        const uint32_t bal = __ballot_sync(0xffffffffu, valid && (acc & 1u));
        if (lane == 0) s_words[it * kWarps + warp] = bal;
    }
    __syncthreads();
    // block scan over the 32 words by warp 0, then publish + look back
    if (warp == 0) {
        const uint32_t word = s_words[lane];
        const uint32_t pc = __popc(word);
        uint32_t inc = pc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t tmp = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= static_cast<uint32_t>(d)) inc += tmp;
        }
        s_woff[lane] = inc - pc;
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        if (lane == 0) st_desc(p.tile_desc + tile, make_desc(p.epoch, tile == 0 ? kStatePrefix : kStateAgg, total));
        const uint32_t excl = warp_lookback(p.tile_desc, tile, p.epoch, lane);
        if (lane == 0) {
            st_desc(p.tile_desc + tile, make_desc(p.epoch, kStatePrefix, excl + total));
            if (tile == p.n_tiles - 1) p.ctl->out_count = static_cast<unsigned long long>(excl) + total;
            s_total = total;
            s_excl = excl;
        }
    }
    __syncthreads();
    if (p.out_ids == nullptr || s_total == 0) return;
    const uint32_t excl = s_excl;
#pragma unroll
    for (int it = 0; it < kFilterItems; ++it) {
        const uint32_t w = s_words[it * kWarps + warp];
        if ((w >> lane) & 1u) p.out_ids[excl + s_woff[it * kWarps + warp] + __popc(w & lanemask_lt())] = rows[it];
    }
}

cudaError_t filter_launch(const DevTable &t, const QueryCtl *d_ctl, const CandSegments &segs,
                          unsigned long long *tile_desc, uint32_t epoch, uint32_t *out_ids, cudaStream_t stream) {
    FilterParams p{};
    for (int c = 0; c < NUM_COLS; ++c) {
        p.col[c] = t.col[c].d;
        p.width[c] = t.col[c].width;
    }
    p.segs = segs;
    p.n_cand = segs.vstart[segs.n_seg];
    p.n_tiles = filter_tiles(p.n_cand);
    p.ctl = const_cast<QueryCtl *>(d_ctl);
    p.tile_desc = tile_desc;
    p.epoch = epoch;
    p.out_ids = out_ids;
    if (p.n_tiles == 0) return cudaSuccess;
    filter_kernel<<<static_cast<unsigned>(p.n_tiles), kFilterThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K2: gather
// ------------------------------------------------------------------------------------------
template <typename V>
__global__ void gather_kernel(const V *__restrict__ col, int vec_per_row, const uint32_t *__restrict__ ids,
                              long long n_ids, V *__restrict__ out) {
    const long long total = n_ids * vec_per_row;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long k = i / vec_per_row;
        const int v = static_cast<int>(i - k * vec_per_row);
        out[i] = __ldg(col + static_cast<size_t>(__ldg(ids + k)) * vec_per_row + v);
    }
}

static int grid_for(long long work, int threads) {
    long long g = (work + threads - 1) / threads;
    const long long cap = 148ll * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

cudaError_t gather_launch(const uint8_t *col, uint32_t width, const uint32_t *ids, int64_t n_ids, uint8_t *out,
                          cudaStream_t stream) {
    if (n_ids <= 0) return cudaSuccess;
    const int threads = 256;
    if (width % 16 == 0) {
        const int vpr = static_cast<int>(width / 16);
        gather_kernel<uint4><<<grid_for(n_ids * vpr, threads), threads, 0, stream>>>(
            reinterpret_cast<const uint4 *>(col), vpr, ids, n_ids, reinterpret_cast<uint4 *>(out));
    } else if (width == 8) {
        gather_kernel<unsigned long long><<<grid_for(n_ids, threads), threads, 0, stream>>>(
            reinterpret_cast<const unsigned long long *>(col), 1, ids, n_ids,
            reinterpret_cast<unsigned long long *>(out));
    } else if (width == 4) {
        gather_kernel<uint32_t><<<grid_for(n_ids, threads), threads, 0, stream>>>(
            reinterpret_cast<const uint32_t *>(col), 1, ids, n_ids, reinterpret_cast<uint32_t *>(out));
    } else if (width == 1) {
        gather_kernel<uint8_t><<<grid_for(n_ids, threads), threads, 0, stream>>>(col, 1, ids, n_ids, out);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

__global__ void restride_kernel(const uint4 *__restrict__ src, int src_v, uint4 *__restrict__ dst, int dst_v,
                                long long n) {
    const long long total = n * dst_v;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / dst_v;
        const int v = static_cast<int>(i - r * dst_v);
        dst[i] = v < src_v ? src[r * src_v + v] : make_uint4(0, 0, 0, 0);
    }
}

cudaError_t restride_launch(const uint8_t *src, uint32_t src_w, uint8_t *dst, uint32_t dst_w, int64_t n,
                            cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if (src_w % 16 || dst_w % 16 || dst_w < src_w) return cudaErrorInvalidValue;
    restride_kernel<<<grid_for(n * (dst_w / 16), 256), 256, 0, stream>>>(
        reinterpret_cast<const uint4 *>(src), static_cast<int>(src_w / 16), reinterpret_cast<uint4 *>(dst),
        static_cast<int>(dst_w / 16), n);
    return cudaGetLastError();
}

__global__ void add_base_kernel(uint32_t *ids, long long n, uint32_t base) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        ids[i] += base;
}

cudaError_t add_base_launch(uint32_t *ids, int64_t n, uint32_t base, cudaStream_t stream) {
    if (n <= 0 || base == 0) return cudaSuccess;
    add_base_kernel<<<grid_for(n, 256), 256, 0, stream>>>(ids, n, base);
    return cudaGetLastError();
}

}  // namespace qpe
