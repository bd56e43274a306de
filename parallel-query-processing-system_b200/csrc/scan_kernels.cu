// scan_kernels.cu -- SELECT/WHERE hot path for B200 (sm_100a)
//
// K1  scan_tma_kernel      full-table predicate evaluation -> match bitmap + match count.
//       Replaces linearSearchRecords/evaluateWhereClause/checkCondition
//       (engine/serial/executeEngine-serial.c:854-878, :292-316, :251-289).
//       Warp-specialised persistent CTAs, one per SM:
//         P  (1 warp, 1 lane)  streams every referenced column's slice of a tile into shared
//                              memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier
//                              complete_tx), S stages deep;
//         E  (kEvalWarps)      evaluate the compiled WHERE program on the staged tile, one R-bit
//                              match mask per lane (the operator dispatch happens once per tile,
//                              the per-row work is load / compare / predicated OR), transpose it
//                              with __ballot_sync into bitmap words (bit b of word w = row 32w+b)
//                              and store them.  The bitmap is 1 bit per row: 1/104 of the
//                              narrowest scan's input traffic.
//       Nothing in K1 waits on another CTA, so HBM streaming is never stalled by a scan chain.
// K1c compact_kernel       order-preserving stream compaction of the bitmap into row ids:
//       popc per word, block scan, single-pass DECOUPLED LOOK-BACK over 64 Ki-row chunks
//       (status flag + aggregate / inclusive prefix per chunk), ids staged in shared memory and
//       written with coalesced stores -- rows come out in exactly table order.  The same bitmap
//       is DELETE's match mask.  (A first version ran the look-back per 4 Ki-row tile inside K1:
//       at HBM speed that is >100 descriptors/us, more than a 32-wide look-back window can
//       follow at L2 latency, and the chain fell behind; see DESIGN.md.)
// K1f scan_fused_kernel    THE SELECT KERNEL: K1's producer / evaluators plus 4 (or 8) compaction warps per CTA that turn
//       each finished <= 64 Ki-row chunk of the bitmap -- kept in shared memory, never written to HBM -- into
//       row ids with the same decoupled look-back, beside the scan.  One launch per query; the program comes as a
//       kernel parameter and the last CTA hands the count over and resets the control words (banner further down).
// K9  scan_batch_kernel    up to 8 WHERE programs over one pass of the same columns (query batch).
// K1g filter_kernel        same program on a gathered candidate list (index path), ordered
//                          single-pass compaction with decoupled look-back per 1 Ki candidates.
// K2  gather_kernel        projection / column compaction gather.
//
// HBM-bound integer/byte work: no tensor cores by design.

#include "scan_kernels.cuh"

#include <cstdint>
#include <cstdlib>

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
// non-blocking probe of a barrier phase (try_wait may suspend the thread for a while; this never does)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Every spin in this file is bounded: a protocol bug (or a peer that never arrives) ends in a trapped
// kernel and a CUDA error on the host, never in a hung GPU.  2^31 polls are minutes of waiting.
constexpr uint32_t kSpinLimit = 0x7fffffffu;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins == kSpinLimit) __trap();
    }
}
// The same wait for a thread that expects to wait LONG (the TMA producer waits ~1 us per tile for a stage to be
// released): sleep between polls.  A tight poll loop is 6 instructions per poll; ncu showed the single producer
// lane issuing 12 % of the whole kernel's instructions that way, on a scheduler it shares with four evaluator warps.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity, uint32_t sleep_ns) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(sleep_ns);
        if (++spins == kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
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
constexpr int kEvalWarps = 16;                   // evaluator warps per CTA (4 per scheduler: latency hiding)
constexpr int kEvalWarpsWide = 8;                // variant for very wide rows (256-row tiles)
constexpr int kMaxStages = 8;
constexpr int kRowsPerGroup = 32 * kEvalWarps;   // rows one "row group" covers: a tile is R groups
constexpr int kMaxR = 8;                         // rows per lane per tile: 1, 2, 4 or 8
constexpr int kMaxTileRows = kRowsPerGroup * kMaxR;  // 4096 (<= kRowPad: a full tile is always inside the allocation)

struct ScanParams {
    const uint8_t *col[NUM_COLS];
    uint32_t width[NUM_COLS];
    uint32_t smem_off[NUM_COLS];  // byte offset of the column inside a stage
    int32_t n_ref;
    int32_t ref_col[NUM_COLS];
    uint32_t stage_bytes;
    int32_t tile_rows;
    int32_t n_stages;
    int32_t dynamic_tiles;  // 1: claim tiles from the global counter; 0: tile = cta + k * grid
    long long n_rows;
    long long tile_begin;   // this launch covers tiles [tile_begin, n_tiles): a pipelined scan launches
    long long n_tiles;      // K1 once per table segment so that K1c of segment i overlaps K1 of segment i+1
    QueryCtl *ctl;
    uint32_t *out_bitmap;   // n_tiles * tile_rows / 32 words, or null (count only)
};

struct ScanSmemHeader {
    Program prog;
    alignas(8) uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    long long tile_of_stage[kMaxStages];
    unsigned long long cta_count;
};

// R-bit mask of a per-row predicate, fully unrolled so every shift is an immediate
template <int R, typename F>
__device__ __forceinline__ uint32_t rows_mask(F pred) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) m |= pred(j) ? (1u << j) : 0u;
    return m;
}

// numeric leaf: the operator is warp-uniform, so it is dispatched ONCE per evaluation and each case is
// a straight run of RB (load, compare, predicated OR) triples; ld(j) loads the value of this lane's j-th row
template <int RB, typename T, typename Ld>
__device__ __forceinline__ uint32_t cmp_numeric(Ld ld, const T lit, const uint32_t tt) {
    switch (tt) {
        case 0b010: return rows_mask<RB>([&](int j) { return ld(j) == lit; });
        case 0b101: return rows_mask<RB>([&](int j) { return ld(j) != lit; });
        case 0b100: return rows_mask<RB>([&](int j) { return ld(j) > lit; });
        case 0b001: return rows_mask<RB>([&](int j) { return ld(j) < lit; });
        case 0b110: return rows_mask<RB>([&](int j) { return ld(j) >= lit; });
        case 0b011: return rows_mask<RB>([&](int j) { return ld(j) <= lit; });
        case 0b111: return (RB >= 32) ? 0xffffffffu : ((1u << RB) - 1u);
        default: return 0u;
    }
}

__device__ __forceinline__ uint32_t diff16(const uint4 v, const uint4 l) {
    return (v.x ^ l.x) | (v.y ^ l.y) | (v.z ^ l.z) | (v.w ^ l.w);
}

// After the program has been copied into shared memory: give every leaf its column's stage offset and width, so
// that evaluating a leaf needs nothing but the leaf record (call between two CTA-wide barriers).
__device__ __forceinline__ void patch_leaves(Program *sp, const ScanParams &p, uint32_t tid, uint32_t n_threads) {
    for (uint32_t k = tid; k < static_cast<uint32_t>(sp->n_leaves); k += n_threads) {
        const int c = sp->leaf[k].col;
        sp->leaf[k].smem_off = p.smem_off[c];
        sp->leaf[k].nch = static_cast<uint8_t>(p.width[c] >> 4);
    }
}

// Evaluate one leaf for this lane's rows of NT staged tiles at once (NT = 1: rows lrow, lrow+32, ... of the
// tile at `stage`; NT = 2: also the same rows of the tile at `stage + dB`, which become mask bits R .. 2R-1).
// Two tiles per call halve the per-tile cost of interpreting the program (~60 % of the evaluators'
// instructions for a 3-leaf WHERE), so the fused kernel takes two whenever the next stage has already landed.
template <int R, int NT>
__device__ __forceinline__ uint32_t eval_leaf_tile(const PLeaf &lf, const Program *sp, const uint8_t *stage,
                                                   const int dB, const ScanParams &p, int lrow) {
    constexpr int RB = R * NT;
    const uint8_t *base = stage + lf.smem_off;  // patched by patch_leaves
    const uint32_t tt = lf.tt;
    constexpr uint32_t kAll = (RB >= 32) ? 0xffffffffu : ((1u << RB) - 1u);
    // byte address of this lane's j-th row of a column with `w`-byte cells (j is a constant after unrolling)
    auto row_ptr = [&](const uint8_t *c0, int j, uint32_t w) -> const uint8_t * {
        return (j < R) ? c0 + static_cast<size_t>(j) * 32u * w : c0 + dB + static_cast<size_t>(j - R) * 32u * w;
    };
    switch (lf.type) {
        case T_I32: {
            const uint8_t *c = base + static_cast<size_t>(lrow) * 4u;
            return cmp_numeric<RB, int32_t>([&](int j) { return *reinterpret_cast<const int32_t *>(row_ptr(c, j, 4)); },
                                            lf.lit_i32, tt);
        }
        case T_U64: {
            const uint8_t *c = base + static_cast<size_t>(lrow) * 8u;
            return cmp_numeric<RB, unsigned long long>(
                [&](int j) { return *reinterpret_cast<const unsigned long long *>(row_ptr(c, j, 8)); }, lf.lit_u64, tt);
        }
        case T_BOOL: {
            // only = and != exist; (cell != 0) == want
            const uint8_t *c = base + lrow;
            const bool want = ((tt == 0b010u) == ((lf.lit_i32 & 1) != 0));
            const uint32_t nz = rows_mask<RB>([&](int j) { return *row_ptr(c, j, 1) != 0; });
            return want ? nz : (~nz & kAll);
        }
        default: {  // T_STR: strcmp order == unsigned byte order over the NUL-padded cell
            const int nch = lf.nch;
            const uint32_t w = static_cast<uint32_t>(nch) << 4;
            const uint4 *lit = reinterpret_cast<const uint4 *>(sp->lit_pool + lf.lit_off);
            const uint8_t *c = base + static_cast<size_t>(lrow) * w;
            auto cell = [&](int j) { return reinterpret_cast<const uint4 *>(row_ptr(c, j, w)); };
            if (tt == 0b010u || tt == 0b101u) {
                // equality only: OR of XORs, no byte swapping
                uint32_t ne = 0;
                if (nch == 1) {
                    const uint4 l0 = lit[0];
                    ne = rows_mask<RB>([&](int j) { return diff16(*cell(j), l0) != 0u; });
                } else {
                    for (int k = 0; k < nch; ++k) {
                        const uint4 lk = lit[k];
                        ne |= rows_mask<RB>([&](int j) { return diff16(cell(j)[k], lk) != 0u; });
                        if (__all_sync(0xffffffffu, ne == kAll)) break;  // warp-uniform early exit
                    }
                }
                return (tt == 0b010u) ? (~ne & kAll) : ne;
            }
            // ordering: three-way compare chunk by chunk, first differing chunk decides
            uint32_t lt = 0, decided = 0;
            for (int k = 0; k < nch; ++k) {
                const uint4 lk = lit[k];
                const uint32_t b0 = bswap32(lk.x), b1 = bswap32(lk.y), b2 = bswap32(lk.z), b3 = bswap32(lk.w);
                uint32_t dk = 0, ltk = 0;
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const uint4 v = cell(j)[k];
                    const uint32_t a0 = bswap32(v.x), a1 = bswap32(v.y), a2 = bswap32(v.z), a3 = bswap32(v.w);
                    const bool d0 = a0 != b0, d1 = a1 != b1, d2 = a2 != b2, d3 = a3 != b3;
                    const bool l = d0 ? (a0 < b0) : d1 ? (a1 < b1) : d2 ? (a2 < b2) : (a3 < b3);
                    dk |= (d0 | d1 | d2 | d3) ? (1u << j) : 0u;
                    ltk |= l ? (1u << j) : 0u;
                }
                const uint32_t fresh = dk & ~decided;
                lt |= ltk & fresh;
                decided |= dk;
                if (__all_sync(0xffffffffu, decided == kAll)) break;
            }
            const uint32_t eq = ~decided & kAll;
            const uint32_t gt = decided & ~lt;
            return ((tt & 1u) ? lt : 0u) | ((tt & 2u) ? eq : 0u) | ((tt & 4u) ? gt : 0u);
        }
    }
}

// Transpose this warp's R-bit row masks into R bitmap words (word j = rows [32 j, 32 j + 32) of the warp's slice
// of a tile) and store them with ONE lane: every ballot result is warp-uniform, so lane 0 holds all R words and
// writes them as 16-byte vectors (dst is R*4-byte aligned).  The earlier form -- lane j keeps word j, R lanes store
// one word each -- spent 8 compares + selects per tile on picking the word.
template <int R>
__device__ __forceinline__ void store_mask_words(uint32_t acc, uint32_t lane, uint32_t *dst, uint32_t *dst2) {
    uint32_t w[R];
#pragma unroll
    for (int j = 0; j < R; ++j) w[j] = __ballot_sync(0xffffffffu, (acc & (1u << j)) != 0u);
    if (lane == 0) {
        if constexpr (R % 4 == 0) {
#pragma unroll
            for (int j = 0; j < R; j += 4) {
                const uint4 v = make_uint4(w[j], w[j + 1], w[j + 2], w[j + 3]);
                *reinterpret_cast<uint4 *>(dst + j) = v;
                if (dst2) *reinterpret_cast<uint4 *>(dst2 + j) = v;
            }
        } else if constexpr (R == 2) {
            const uint2 v = make_uint2(w[0], w[1]);
            *reinterpret_cast<uint2 *>(dst) = v;
            if (dst2) *reinterpret_cast<uint2 *>(dst2) = v;
        } else {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                dst[j] = w[j];
                if (dst2) dst2[j] = w[j];
            }
        }
    }
}

// run the compiled WHERE program; returns the R-bit match mask of this lane's rows
template <typename LeafFn>
__device__ __forceinline__ uint32_t run_program(const Program *sp, uint32_t all_mask, LeafFn leaf_fn) {
    uint32_t acc = all_mask;  // empty program == NULL where clause == every row matches
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    const int n = sp->n_instr;
    for (int i = 0; i < n; ++i) {
        const PInstr in = sp->instr[i];
        if (in.op <= P_LEAF_OR) {
            // one inlined copy of the (large, unrolled) leaf evaluator for all three leaf ops
            const uint32_t v = leaf_fn(sp->leaf[in.arg]);
            acc = (in.op == P_LEAF_SET) ? v : (in.op == P_LEAF_AND) ? (acc & v) : (acc | v);
            continue;
        }
        switch (in.op) {
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
            unsigned long long d = ld_desc(desc + mine);
            uint32_t flag = static_cast<uint32_t>(d >> 32);
            uint32_t spins = 0;
            while ((flag >> 2) != epoch) {
                __nanosleep(64);  // predecessor not published yet: back off instead of hammering L2
                if (++spins == (kSpinLimit >> 6)) __trap();
                d = ld_desc(desc + mine);
                flag = static_cast<uint32_t>(d >> 32);
            }
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

template <int EW, int R>
__global__ void __launch_bounds__(32 * (1 + EW), 1) scan_tma_kernel(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ScanSmemHeader *sh = reinterpret_cast<ScanSmemHeader *>(smem_raw);
    uint8_t *stages = smem_raw + ((sizeof(ScanSmemHeader) + 127) & ~size_t(127));

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31u;
    const int S = p.n_stages;
    constexpr int T = 32 * EW * R;        // rows per tile
    constexpr int WPT = T >> 5;           // bitmap words per tile
    constexpr uint32_t kThreads = 32 * (1 + EW);

    // program -> shared memory (uniform reads afterwards), barrier init
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&p.ctl->prog);
        uint4 *dst = reinterpret_cast<uint4 *>(&sh->prog);
        for (uint32_t i = tid; i < sizeof(Program) / 16; i += kThreads) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&sh->full[s], 1);
            mbar_init(&sh->empty[s], EW);
        }
        sh->cta_count = 0;
        fence_mbar_init();
    }
    __syncthreads();
    patch_leaves(&sh->prog, p, tid, kThreads);
    __syncthreads();
    const Program *sp = &sh->prog;

    if (warp == 0) {
        // ===== P: tile claim + TMA producer =====
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (long long k = 0;; ++k) {
                mbar_wait_relaxed(&sh->empty[s], phase ^ 1u, 128);
                // static round-robin keeps the claim off the critical path (a global atomic costs a
                // full round trip per tile); every CTA is resident (grid <= SM count) and walks its
                // tiles in increasing order, so the look-back chain always makes progress.
                const long long tile = p.tile_begin +
                                       (p.dynamic_tiles ? static_cast<long long>(atomicAdd(&p.ctl->tile_counter, 1u))
                                                        : static_cast<long long>(blockIdx.x) + k * gridDim.x);
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
                if (++s == S) {
                    s = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        // ===== E: predicate evaluation =====
        const int ew = static_cast<int>(warp) - 1;
        constexpr uint32_t all_mask = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
        const int lrow = ew * (32 * R) + static_cast<int>(lane);  // first row of this lane inside a tile
        int s = 0;
        uint32_t sphase = 0;
        uint32_t my_count = 0;  // matches seen by this lane's rows (summed per CTA at the end)
        for (;;) {
            mbar_wait(&sh->full[s], sphase);
            const long long tile = sh->tile_of_stage[s];
            if (tile < 0) break;

            const uint8_t *stage = stages + static_cast<size_t>(s) * p.stage_bytes;
            uint32_t acc = run_program(sp, all_mask, [&](const PLeaf &lf) {
                return eval_leaf_tile<R, 1>(lf, sp, stage, 0, p, lrow);
            });
            // the stage can be refilled as soon as every evaluator has read it
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->empty[s]);

            const long long row_base = tile * T + lrow;  // global row of this lane's j = 0
            if (tile * T + T > p.n_rows)                 // last (partial) tile: drop the padding rows
                acc &= rows_mask<R>([&](int j) { return row_base + 32ll * j < p.n_rows; });
            my_count += static_cast<uint32_t>(__popc(acc));
            if (p.out_bitmap) store_mask_words<R>(acc, lane, p.out_bitmap + tile * WPT + ew * R, nullptr);
            if (++s == S) {
                s = 0;
                sphase ^= 1u;
            }
        }
        const uint32_t warp_total = __reduce_add_sync(0xffffffffu, my_count);
        if (lane == 0) atomicAdd(&sh->cta_count, static_cast<unsigned long long>(warp_total));
    }
    __syncthreads();
    if (tid == 0 && sh->cta_count) atomicAdd(&p.ctl->out_count, sh->cta_count);
}


// ------------------------------------------------------------------------------------------
// K9: query batch -- up to kMaxBatch WHERE programs over ONE pass of the columns
//
// The GPU analogue of QPEOMP's query-level parallelism (QPEOMP.c:234-335: one thread per query over one
// engine): the union of the columns the queries reference is staged once per tile, the evaluators run every
// program on the staged tile and leave one match bitmap + count per query; K1c then compacts each bitmap.
// HBM traffic is that of ONE scan, so a batch is bound by the evaluators' instruction rate instead.
// ------------------------------------------------------------------------------------------
struct BatchParams {
    ScanParams s;
    int32_t n_prog;
    int32_t pad;
    const Program *progs;                 // device, n_prog programs
    uint32_t *bitmap[kMaxBatch];          // device, one per query
    unsigned long long *counts;           // device, n_prog match counts (zeroed by the caller)
};

template <int EW, int R>
__global__ void __launch_bounds__(32 * (1 + EW), 1) scan_batch_kernel(const __grid_constant__ BatchParams bp) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const ScanParams &p = bp.s;
    ScanSmemHeader *sh = reinterpret_cast<ScanSmemHeader *>(smem_raw);  // its own prog slot is unused here
    __shared__ unsigned long long s_count[kMaxBatch];
    Program *progs = reinterpret_cast<Program *>(smem_raw + ((sizeof(ScanSmemHeader) + 127) & ~size_t(127)));
    const int Q = bp.n_prog;
    uint8_t *stages = reinterpret_cast<uint8_t *>(progs) + ((static_cast<size_t>(Q) * sizeof(Program) + 127) & ~size_t(127));

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31u;
    const int S = p.n_stages;
    constexpr int T = 32 * EW * R;
    constexpr int WPT = T >> 5;
    constexpr uint32_t kThreads = 32 * (1 + EW);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(bp.progs);
        uint4 *dst = reinterpret_cast<uint4 *>(progs);
        for (uint32_t i = tid; i < static_cast<uint32_t>(Q) * (sizeof(Program) / 16); i += kThreads) dst[i] = src[i];
    }
    if (tid < kMaxBatch) s_count[tid] = 0;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&sh->full[s], 1);
            mbar_init(&sh->empty[s], EW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    for (int q = 0; q < Q; ++q) patch_leaves(progs + q, p, tid, kThreads);
    __syncthreads();

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (long long k = 0;; ++k) {
                mbar_wait_relaxed(&sh->empty[s], phase ^ 1u, 128);
                const long long tile = static_cast<long long>(blockIdx.x) + k * gridDim.x;
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
                if (++s == S) {
                    s = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        const int ew = static_cast<int>(warp) - 1;
        constexpr uint32_t all_mask = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
        const int lrow = ew * (32 * R) + static_cast<int>(lane);
        int s = 0;
        uint32_t sphase = 0;
        for (;;) {
            mbar_wait(&sh->full[s], sphase);
            const long long tile = sh->tile_of_stage[s];
            if (tile < 0) break;
            const uint8_t *stage = stages + static_cast<size_t>(s) * p.stage_bytes;
            const long long row_base = tile * T + lrow;
            uint32_t tail = all_mask;
            if (tile * T + T > p.n_rows) tail = rows_mask<R>([&](int j) { return row_base + 32ll * j < p.n_rows; });
            for (int q = 0; q < Q; ++q) {
                const Program *sp = progs + q;
                const uint32_t acc = tail & run_program(sp, all_mask, [&](const PLeaf &lf) {
                                         return eval_leaf_tile<R, 1>(lf, sp, stage, 0, p, lrow);
                                     });
                const uint32_t warp_cnt = __reduce_add_sync(0xffffffffu, static_cast<uint32_t>(__popc(acc)));
                if (lane == 0 && warp_cnt) atomicAdd(&s_count[q], static_cast<unsigned long long>(warp_cnt));
                store_mask_words<R>(acc, lane, bp.bitmap[q] + tile * WPT + ew * R, nullptr);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->empty[s]);  // every program has read the stage
            if (++s == S) {
                s = 0;
                sphase ^= 1u;
            }
        }
    }
    __syncthreads();
    if (tid < static_cast<uint32_t>(Q) && s_count[tid]) atomicAdd(&bp.counts[tid], s_count[tid]);
}

// ------------------------------------------------------------------------------------------
// K1f: fused scan + ordered compaction (one launch, ids leave the SM while the scan is running)
//
// Same producer / evaluator roles as K1, plus kFuseCompactWarps COMPACTION warps per CTA.  Work is
// handed out in CHUNKS of chunk_tiles consecutive tiles (<= 64 Ki rows), chunk c to CTA c % grid, in
// increasing order.  The evaluators leave the chunk's match bitmap in shared memory (double
// buffered); the compaction warps popc/scan it, publish the chunk aggregate, run the DECOUPLED
// LOOK-BACK over the chunk descriptors of the other CTAs (all resident: grid <= SM count, every
// CTA walks its chunks in increasing order, so the smallest unfinished chunk never waits), expand
// the bits into row ids in shared memory and store them coalesced at out_ids + prefix -- while the
// evaluators are already two chunks ahead.  Compared with K1 -> K1c: no second launch, no bitmap
// round trip through HBM, no barrier-bound compaction pass after the scan (K1c: 7.5 % of a step).
// The look-back granularity stays one descriptor per chunk (a per-tile chain cannot keep up with
// HBM), but the chain now advances beside the scan instead of after it.
// Optional progress words in pinned host memory tell the host when a table segment's ids are
// complete in HBM, so it can start the device->host copy of that segment during the scan.
// ------------------------------------------------------------------------------------------
// CW = compaction warps per CTA, 4 or 8.  Four are plenty while few rows match (and measured 3.5 % faster there:
// 2.07 vs 2.15 ms on 1 B rows at 1 %); from ~10 % selectivity the evaluators wait for the compaction warps
// (ncu: 18.6 % of all samples in their wait for a free chunk buffer at 50 %), and eight cut 2.79 ms to 2.39 ms.
// The engine picks per query from the selectivity the previous full scan saw (engine.cu).
constexpr int kFuseChunkWords = kFuseMaxChunkRows / 32;              // 2048 words
static_assert(fuse_reserve_bytes(4) == 2 * kFuseChunkWords * 4 + 32 * 32 * 4 * 4, "smem reserve");
static_assert(fuse_reserve_bytes(8) == 2 * kFuseChunkWords * 4 + 32 * 32 * 8 * 4, "smem reserve");

struct FusedParams {
    ScanParams s;
    int32_t chunk_tiles;            // tiles per chunk; chunk_tiles * tile_rows <= kFuseMaxChunkRows
    uint32_t poll_ns;               // sleep between the compaction warps' polls of the chunk barrier
    long long n_chunks;
    unsigned long long *desc;       // one look-back descriptor per chunk
    uint32_t epoch;
    uint32_t id_base;
    uint32_t *out_ids;              // HBM of this GPU, a peer mapping, or mapped pinned host memory
    unsigned long long out_cap;
    long long seg_chunks;           // progress: chunks per table segment (0 = no progress words)
    unsigned long long *progress;   // mapped pinned host memory: [seg] = epoch << 32 | ids complete through seg
    unsigned long long *host_count; // mapped pinned host memory: the match count, written by the last CTA (or null)
    FusedCtl *fctl;                 // this kernel's own control words (self-resetting)
    Program prog;                   // the compiled WHERE, by value: no upload precedes the launch
};

struct FusedSmemHeader {   // sized for the 4-warp variant (16 rounds x 4 warps); the 8-warp one needs 8 x 8
    Program prog;
    alignas(8) uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t cb_full[2];
    uint64_t cb_empty[2];
    long long tile_of_stage[kMaxStages];
    unsigned long long cta_count;
    uint32_t warp_tot[64];          // [round][compaction warp]
    uint32_t round_base[16 + 1];
    uint32_t excl;
};

template <int EW, int R, int CW>
__global__ void __launch_bounds__(32 * (1 + EW + CW), 1)
    scan_fused_kernel(const __grid_constant__ FusedParams fp) {
    constexpr int kFuseCompactWarps = CW;
    constexpr int kFuseCompactThreads = 32 * CW;                           // 128 | 256
    constexpr int kFuseRounds = kFuseChunkWords / kFuseCompactThreads;     // 16 | 8 words per compaction thread
    constexpr int kFuseStageIds = 32 * kFuseCompactThreads;                // a round never holds more ids
    static_assert(kFuseRounds * CW <= 64 && kFuseRounds <= 16, "FusedSmemHeader scratch");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const ScanParams &p = fp.s;
    FusedSmemHeader *sh = reinterpret_cast<FusedSmemHeader *>(smem_raw);
    uint8_t *stages = smem_raw + ((sizeof(FusedSmemHeader) + 127) & ~size_t(127));
    const int S = p.n_stages;
    uint32_t *cbuf = reinterpret_cast<uint32_t *>(stages + static_cast<size_t>(S) * p.stage_bytes);  // [2][2048]
    uint32_t *id_stage = cbuf + 2 * kFuseChunkWords;                                                    // [4096]

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31u;
    constexpr int T = 32 * EW * R;   // rows per tile
    constexpr int WPT = T >> 5;      // bitmap words per tile
    constexpr uint32_t kThreads = 32 * (1 + EW + kFuseCompactWarps);
    constexpr bool kDual = (2 * R <= 16);  // two tiles per pass of the program (mask bits 0..2R-1)
    const int CT = fp.chunk_tiles;

    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&fp.prog);  // kernel parameter space
        uint4 *dst = reinterpret_cast<uint4 *>(&sh->prog);
        for (uint32_t i = tid; i < sizeof(Program) / 16; i += kThreads) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&sh->full[s], 1);
            mbar_init(&sh->empty[s], EW);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sh->cb_full[b], EW);
            mbar_init(&sh->cb_empty[b], 1);
        }
        sh->cta_count = 0;
        fence_mbar_init();
    }
    __syncthreads();
    patch_leaves(&sh->prog, p, tid, kThreads);
    __syncthreads();
    const Program *sp = &sh->prog;

    if (warp == 0) {
        // ===== P: TMA producer, tiles of this CTA's chunks in order =====
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (long long chunk = blockIdx.x;; chunk += gridDim.x) {
                const bool done = chunk >= fp.n_chunks;
                const long long t0 = chunk * CT;
                const int nt = done ? 1 : static_cast<int>((p.n_tiles - t0 < CT) ? (p.n_tiles - t0) : CT);
                for (int j = 0; j < nt; ++j) {
                    mbar_wait_relaxed(&sh->empty[s], phase ^ 1u, 128);
                    if (done) {
                        sh->tile_of_stage[s] = -1;
                        mbar_arrive(&sh->full[s]);
                    } else {
                        const long long tile = t0 + j;
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
                    if (++s == S) {
                        s = 0;
                        phase ^= 1u;
                    }
                }
                if (done) break;
            }
        }
    } else if (warp <= EW) {
        // ===== E: predicate evaluation, bitmap words into the chunk buffer =====
        const int ew = static_cast<int>(warp) - 1;
        constexpr uint32_t all_mask = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
        const int lrow = ew * (32 * R) + static_cast<int>(lane);
        int s = 0;
        uint32_t sphase = 0;
        uint32_t my_count = 0;
        uint32_t k = 0;  // k-th chunk of this CTA
        for (long long chunk = blockIdx.x; chunk < fp.n_chunks; chunk += gridDim.x, ++k) {
            const uint32_t buf = k & 1u;
            // the compaction warps must have read this buffer's previous chunk (k - 2)
            mbar_wait(&sh->cb_empty[buf], ((k >> 1) & 1u) ^ 1u);
            uint32_t *cb = cbuf + buf * kFuseChunkWords;
            const long long t0 = chunk * CT;
            const int nt = static_cast<int>((p.n_tiles - t0 < CT) ? (p.n_tiles - t0) : CT);
            for (int j = 0; j < nt;) {
                mbar_wait(&sh->full[s], sphase);
                const long long tile = t0 + j;
                const uint8_t *stage = stages + static_cast<size_t>(s) * p.stage_bytes;
                int s2 = s + 1;
                uint32_t phase2 = sphase;
                if (s2 == S) {
                    s2 = 0;
                    phase2 ^= 1u;
                }
                // Two tiles in one pass of the program when the NEXT stage has already landed (the scan is then
                // instruction-bound, not waiting for HBM) and a further stage stays in flight (S >= 3).
                bool dual = false;
                if (kDual && S >= 3 && j + 1 < nt) {
                    uint32_t ready = 0;
                    if (lane == 0) ready = mbar_test(&sh->full[s2], phase2) ? 1u : 0u;
                    dual = __shfl_sync(0xffffffffu, ready, 0) != 0u;
                }
                if (dual) {
                    if constexpr (kDual) {
                        mbar_wait(&sh->full[s2], phase2);  // every lane observes the phase (returns at once)
                        const int dB = (s2 - s) * static_cast<int>(p.stage_bytes);
                        constexpr uint32_t all2 = (2 * R >= 32) ? 0xffffffffu : ((1u << (2 * R)) - 1u);
                        uint32_t acc2 = run_program(sp, all2, [&](const PLeaf &lf) {
                            return eval_leaf_tile<R, 2>(lf, sp, stage, dB, p, lrow);
                        });
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(&sh->empty[s]);
                            mbar_arrive(&sh->empty[s2]);
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const long long tl = tile + h;
                            uint32_t acc = (acc2 >> (h * R)) & all_mask;
                            const long long row_base = tl * T + lrow;
                            if (tl * T + T > p.n_rows)
                                acc &= rows_mask<R>([&](int jj) { return row_base + 32ll * jj < p.n_rows; });
                            my_count += static_cast<uint32_t>(__popc(acc));
                            store_mask_words<R>(acc, lane, cb + (j + h) * WPT + ew * R,
                                                p.out_bitmap ? p.out_bitmap + tl * WPT + ew * R : nullptr);
                        }
                    }
                    j += 2;
                    s = s2 + 1;
                    sphase = phase2;
                    if (s == S) {
                        s = 0;
                        sphase ^= 1u;
                    }
                    continue;
                }
                uint32_t acc = run_program(sp, all_mask, [&](const PLeaf &lf) {
                    return eval_leaf_tile<R, 1>(lf, sp, stage, 0, p, lrow);
                });
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh->empty[s]);

                const long long row_base = tile * T + lrow;
                if (tile * T + T > p.n_rows)
                    acc &= rows_mask<R>([&](int jj) { return row_base + 32ll * jj < p.n_rows; });
                my_count += static_cast<uint32_t>(__popc(acc));
                store_mask_words<R>(acc, lane, cb + j * WPT + ew * R,
                                    p.out_bitmap ? p.out_bitmap + tile * WPT + ew * R : nullptr);
                ++j;
                if (++s == S) {
                    s = 0;
                    sphase ^= 1u;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->cb_full[buf]);  // release: this warp's words of the chunk are in cb
        }
        // consume the producer's end marker so that every barrier phase is balanced
        mbar_wait(&sh->full[s], sphase);
        const uint32_t warp_total = __reduce_add_sync(0xffffffffu, my_count);
        if (lane == 0) atomicAdd(&sh->cta_count, static_cast<unsigned long long>(warp_total));
    } else {
        // ===== C: ordered compaction of finished chunks =====
        const uint32_t ct = tid - 32u * (1 + EW);
        const uint32_t cw = ct >> 5;
        const uint32_t chunk_rows = static_cast<uint32_t>(CT) * T;
        uint32_t k = 0;
        for (long long chunk = blockIdx.x; chunk < fp.n_chunks; chunk += gridDim.x, ++k) {
            const uint32_t buf = k & 1u;
            const long long t0 = chunk * CT;
            const int nt = static_cast<int>((p.n_tiles - t0 < CT) ? (p.n_tiles - t0) : CT);
            const uint32_t nw = static_cast<uint32_t>(nt) * WPT;
            // (one polling warp + a named barrier for the others measured 1.5 % SLOWER than every warp polling)
            mbar_wait_relaxed(&sh->cb_full[buf], (k >> 1) & 1u, fp.poll_ns);
            const uint32_t *cb = cbuf + buf * kFuseChunkWords;
            // 1. words -> registers, popc, warp-inclusive scan per round
            uint32_t word[kFuseRounds], off[kFuseRounds];
#pragma unroll
            for (int r = 0; r < kFuseRounds; ++r) {
                const uint32_t wi = r * kFuseCompactThreads + ct;
                word[r] = wi < nw ? cb[wi] : 0u;
                const uint32_t pc = __popc(word[r]);
                uint32_t inc = pc;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= static_cast<uint32_t>(d)) inc += t;
                }
                off[r] = inc - pc;
                if (lane == 31) sh->warp_tot[r * kFuseCompactWarps + cw] = inc;
            }
            named_bar_sync(1, kFuseCompactThreads);
            if (ct == 0) mbar_arrive(&sh->cb_empty[buf]);  // the evaluators may refill this buffer
            // 2. round bases, chunk aggregate, look-back
            if (cw == 0) {
                uint32_t rt = 0;
                if (lane < kFuseRounds) {
#pragma unroll
                    for (int w = 0; w < kFuseCompactWarps; ++w) rt += sh->warp_tot[lane * kFuseCompactWarps + w];
                }
                uint32_t inc = rt;
#pragma unroll
                for (int d = 1; d < kFuseRounds; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= static_cast<uint32_t>(d)) inc += t;
                }
                if (lane < kFuseRounds) sh->round_base[lane] = inc - rt;
                const uint32_t total = __shfl_sync(0xffffffffu, inc, kFuseRounds - 1);
                if (lane == 0) {
                    sh->round_base[kFuseRounds] = total;
                    st_desc(fp.desc + chunk, make_desc(fp.epoch, chunk == 0 ? kStatePrefix : kStateAgg, total));
                }
                const uint32_t excl = warp_lookback(fp.desc, chunk, fp.epoch, lane);
                if (lane == 0) {
                    st_desc(fp.desc + chunk, make_desc(fp.epoch, kStatePrefix, excl + total));
                    sh->excl = excl;
                }
            }
            named_bar_sync(1, kFuseCompactThreads);
            const uint32_t total = sh->round_base[kFuseRounds];
            const uint32_t excl = sh->excl;
            // 3. expand the bits into row ids, staged in shared memory, coalesced stores
            if (total != 0 && static_cast<unsigned long long>(excl) + total <= fp.out_cap) {
                uint32_t *out = fp.out_ids + excl;
                const uint32_t row_chunk = static_cast<uint32_t>(chunk) * chunk_rows + fp.id_base;
                if (total <= kFuseStageIds) {
#pragma unroll
                    for (int r = 0; r < kFuseRounds; ++r) {
                        uint32_t w = word[r];
                        if (w) {
                            uint32_t o = sh->round_base[r] + off[r];
                            for (uint32_t q = 0; q < cw; ++q) o += sh->warp_tot[r * kFuseCompactWarps + q];
                            const uint32_t row0 = row_chunk + (r * kFuseCompactThreads + ct) * 32u;
                            while (w) {
                                const int b = __ffs(w) - 1;
                                w &= w - 1;
                                id_stage[o++] = row0 + static_cast<uint32_t>(b);
                            }
                        }
                    }
                    named_bar_sync(1, kFuseCompactThreads);
                    for (uint32_t i = ct; i < total; i += kFuseCompactThreads) out[i] = id_stage[i];
                } else {
                    // dense chunk: a round (128 words, <= 4096 ids) at a time through the stage.  (Expanding
                    // word by word with lane-parallel direct stores was measured slower: 2.4 x the
                    // instructions, and they compete with the evaluators for issue slots.)
#pragma unroll
                    for (int r = 0; r < kFuseRounds; ++r) {
                        const uint32_t base = sh->round_base[r];
                        const uint32_t cnt = sh->round_base[r + 1] - base;
                        if (cnt == 0) continue;  // uniform over the compaction warps
                        uint32_t w = word[r];
                        uint32_t o = off[r];
                        for (uint32_t q = 0; q < cw; ++q) o += sh->warp_tot[r * kFuseCompactWarps + q];
                        const uint32_t row0 = row_chunk + (r * kFuseCompactThreads + ct) * 32u;
                        while (w) {
                            const int b = __ffs(w) - 1;
                            w &= w - 1;
                            id_stage[o++] = row0 + static_cast<uint32_t>(b);
                        }
                        named_bar_sync(1, kFuseCompactThreads);
                        for (uint32_t i = ct; i < cnt; i += kFuseCompactThreads) out[base + i] = id_stage[i];
                        named_bar_sync(1, kFuseCompactThreads);
                    }
                }
            }
            named_bar_sync(1, kFuseCompactThreads);  // stage / scan scratch are free again; this chunk's stores are issued
            // 4. progress: the last chunk of a table segment to finish publishes the segment to the host
            if (fp.seg_chunks > 0 && ct == 0) {
                __threadfence();
                const long long seg = chunk / fp.seg_chunks;
                const long long first = seg * fp.seg_chunks;
                const long long last = (first + fp.seg_chunks < fp.n_chunks ? first + fp.seg_chunks : fp.n_chunks) - 1;
                const unsigned int stored = atomicAdd(&fp.fctl->seg_stored[seg], 1u) + 1u;
                if (stored == static_cast<unsigned int>(last - first + 1)) {
                    __threadfence();
                    const unsigned long long d = ld_desc(fp.desc + last);  // PREFIX: that chunk has finished
                    __threadfence_system();
                    *reinterpret_cast<volatile unsigned long long *>(fp.progress + seg) =
                        (static_cast<unsigned long long>(fp.epoch) << 32) | static_cast<uint32_t>(d);
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        FusedCtl *fc = fp.fctl;
        if (sh->cta_count) atomicAdd(&fc->out_count, sh->cta_count);
        // The last CTA to get here hands the total over (device word for the sharded post-scan kernel, mapped host
        // word for a plain scan: no count download) and leaves the control words zeroed for the next query.
        __threadfence();
        if (atomicAdd(&fc->ctas_done, 1u) == gridDim.x - 1u) {
            __threadfence();
            const unsigned long long total = atomicExch(&fc->out_count, 0ull);
            fc->final_count = total;
            fc->ctas_done = 0u;
            for (int i = 0; i < kMaxProgressSegments; ++i) fc->seg_stored[i] = 0u;
            if (fp.host_count) {
                *reinterpret_cast<volatile unsigned long long *>(fp.host_count) = total;
                __threadfence_system();
            }
        }
    }
}

size_t fused_param_bytes() { return sizeof(FusedParams); }

// function attributes are per device: cache what was set per (instantiation, device)
constexpr int kMaxDevices = 64;
static int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

bool scan_plan(const DevTable &t, const Program &prog, int force_tile_rows, int force_stages, int max_stages,
               ScanGeometry *geo, const char **why, int fused_cw, size_t extra_reserve) {
    const bool fused = fused_cw != 0;
    if (fused && fused_cw != 4 && fused_cw != 8) {
        if (why) *why = "K1f runs with 4 or 8 compaction warps";
        return false;
    }
    const size_t fuse_reserve = fused ? fuse_reserve_bytes(fused_cw) : 0;
    if (max_stages < 1 || max_stages > kMaxStages) max_stages = kMaxStages;
    // device attributes are asked once per process (one process per GPU; this runs twice per query)
    static int max_smem_by_device[kMaxDevices] = {0}, n_sm_by_device[kMaxDevices] = {0};
    const int slot = current_device_slot();
    if (max_smem_by_device[slot] == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_smem_by_device[slot], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&n_sm_by_device[slot], cudaDevAttrMultiProcessorCount, dev);
    }
    int max_smem = max_smem_by_device[slot], n_sm = n_sm_by_device[slot];
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
    const size_t header = ((fused ? sizeof(FusedSmemHeader) : sizeof(ScanSmemHeader)) + 127) & ~size_t(127);
    const size_t budget = static_cast<size_t>(max_smem) - header - 256 - fuse_reserve - extra_reserve;

    auto stage_bytes_for = [&](int T) {
        size_t sb = 0;
        for (int c = 0; c < NUM_COLS; ++c)
            if (prog.col_mask & (1u << c)) sb += (static_cast<size_t>(T) * t.col[c].width + 127) & ~size_t(127);
        return sb;
    };

    int T = force_tile_rows;
    int S = force_stages;
    if (bpr == 0) {  // no column referenced (NULL where / constants): nothing to stage
        T = T ? T : kMaxTileRows;
        S = S ? S : 2;
    } else if (!T) {
        // largest tile that still leaves >= 2 stages in flight; else the largest with 1.  Measured (100 M rows,
        // profiles/r1_perf_history.md): the per-tile fixed work (program dispatch, barrier waits, ballots) costs
        // more than a third stage buys -- QS 2048x2 6.9 TB/s vs 1024x4 5.3, QD 2048x2 5.4 TB/s vs 1024x4 3.1.
        static const int kTiles[] = {4096, 2048, 1024, 512, 256};
        int best_T = 0, best_S = 0;
        for (int want = 2; want >= 1 && !best_T; --want)
            for (int cand : kTiles) {
                const size_t sb = stage_bytes_for(cand);
                int fit = static_cast<int>(budget / (sb ? sb : 1));
                if (fit > max_stages) fit = max_stages;
                if (fit >= want) {
                    best_T = cand;
                    best_S = fit;
                    break;
                }
            }
        if (!best_T) {
            if (why) *why = "row too wide to stage a tile in shared memory";
            return false;
        }
        T = best_T;
        if (!S) S = best_S;
    } else if (!S) {
        const size_t sb = stage_bytes_for(T);
        S = static_cast<int>(budget / (sb ? sb : 1));
        if (S > max_stages) S = max_stages;
    }
    if ((T != 256 && T != 512 && T != 1024 && T != 2048 && T != 4096) || S < 1 || S > kMaxStages) {
        if (why) *why = "invalid tile geometry (tile rows must be 256, 512, 1024, 2048 or 4096)";
        return false;
    }
    const size_t stage_bytes = stage_bytes_for(T);
    if (stage_bytes * S > budget || stage_bytes >= (1u << 20)) {
        if (why) *why = "row too wide to stage a tile in shared memory";
        return false;
    }
    geo->tile_rows = T;
    geo->stages = S;
    geo->n_tiles = (t.n + T - 1) / T;
    geo->bytes_per_row = bpr;
    geo->smem_bytes = header + stage_bytes * S + 128 + fuse_reserve + extra_reserve;
    geo->compact_warps = fused_cw;
    int64_t grid = geo->n_tiles < n_sm ? geo->n_tiles : n_sm;
    if (grid < 1) grid = 1;
    geo->grid = static_cast<int>(grid);
    geo->chunk_tiles = 0;
    geo->n_chunks = 0;
    if (fused) {
        // chunk = the largest power-of-two run of tiles (<= 64 Ki rows) whose last, partly filled round of
        // CTAs costs < 3 %; halving stops at 8 Ki rows (shorter chunks publish descriptors faster than
        // a look-back window can follow at HBM speed)
        int ct = kFuseMaxChunkRows / T;
        int best_ct = ct;
        double best_loss = 1e9;
        for (; ct >= 1 && static_cast<long long>(ct) * T >= 8192; ct >>= 1) {
            const int64_t nc = (geo->n_tiles + ct - 1) / ct;
            const int64_t g = nc < n_sm ? nc : n_sm;
            const int64_t rounds = (nc + g - 1) / (g > 0 ? g : 1);
            const double loss = nc > 0 ? static_cast<double>(rounds * g) / static_cast<double>(nc) - 1.0 : 0.0;
            if (loss < best_loss - 1e-9) {
                best_loss = loss;
                best_ct = ct;
            }
            if (loss < 0.03) break;
        }
        geo->chunk_tiles = best_ct;
        geo->n_chunks = (geo->n_tiles + best_ct - 1) / best_ct;
        int64_t g = geo->n_chunks < n_sm ? geo->n_chunks : n_sm;
        if (g < 1) g = 1;
        geo->grid = static_cast<int>(g);
    }
    return true;
}

template <int EW, int R>
static cudaError_t launch_scan_r(const ScanParams &p, const ScanGeometry &geo, cudaStream_t stream) {
    // per instantiation and device: raise the dynamic shared memory limit only when it grows
    static size_t allowed_by_device[kMaxDevices] = {0};
    size_t &allowed = allowed_by_device[current_device_slot()];
    if (geo.smem_bytes > allowed) {
        const cudaError_t e = cudaFuncSetAttribute(scan_tma_kernel<EW, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(geo.smem_bytes));
        if (e != cudaSuccess) return e;
        allowed = geo.smem_bytes;
    }
    long long grid = p.n_tiles - p.tile_begin;  // tiles of this launch
    if (grid > geo.grid) grid = geo.grid;
    if (grid < 1) grid = 1;
    scan_tma_kernel<EW, R><<<static_cast<unsigned>(grid), 32 * (1 + EW), geo.smem_bytes, stream>>>(p);
    return cudaGetLastError();
}

static cudaError_t fill_scan_params(const ScanLaunch &L, const ScanGeometry &geo, ScanParams &p) {
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
    // the barrier of a stage is armed with the bytes actually copied; tile_rows is a multiple of
    // 256 and widths are 1, 4, 8 or 16k, so every column slice is already a 128-byte multiple
    uint32_t tx = 0;
    for (int r = 0; r < p.n_ref; ++r) tx += static_cast<uint32_t>(geo.tile_rows) * p.width[p.ref_col[r]];
    p.stage_bytes = static_cast<uint32_t>(off);
    if (tx != p.stage_bytes) return cudaErrorInvalidValue;
    p.tile_rows = geo.tile_rows;
    p.n_stages = geo.stages;
    {
        static const int dyn = [] {
            const char *e = std::getenv("QPE_SCAN_DYNAMIC");
            return (e && e[0] == '1') ? 1 : 0;
        }();
        p.dynamic_tiles = dyn;
    }
    p.n_rows = t.n;
    p.tile_begin = L.tile_begin;
    p.n_tiles = L.tile_end > 0 ? L.tile_end : geo.n_tiles;
    if (p.tile_begin != 0 || p.n_tiles != geo.n_tiles) p.dynamic_tiles = 0;  // segments use the static walk
    p.ctl = const_cast<QueryCtl *>(L.d_ctl);
    p.out_bitmap = L.out_bitmap;
    return cudaSuccess;
}

cudaError_t scan_launch(const ScanLaunch &L, const ScanGeometry &geo, cudaStream_t stream) {
    ScanParams p{};
    const cudaError_t e = fill_scan_params(L, geo, p);
    if (e != cudaSuccess) return e;
    switch (geo.tile_rows) {
        case 256: return launch_scan_r<kEvalWarpsWide, 1>(p, geo, stream);
        case 512: return launch_scan_r<kEvalWarps, 1>(p, geo, stream);
        case 1024: return launch_scan_r<kEvalWarps, 2>(p, geo, stream);
        case 2048: return launch_scan_r<kEvalWarps, 4>(p, geo, stream);
        case 4096: return launch_scan_r<kEvalWarps, 8>(p, geo, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <int EW, int R>
static cudaError_t launch_batch_r(const BatchParams &bp, const ScanGeometry &geo, cudaStream_t stream) {
    static size_t allowed_by_device[kMaxDevices] = {0};
    size_t &allowed = allowed_by_device[current_device_slot()];
    if (geo.smem_bytes > allowed) {
        const cudaError_t e = cudaFuncSetAttribute(scan_batch_kernel<EW, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(geo.smem_bytes));
        if (e != cudaSuccess) return e;
        allowed = geo.smem_bytes;
    }
    scan_batch_kernel<EW, R><<<geo.grid, 32 * (1 + EW), geo.smem_bytes, stream>>>(bp);
    return cudaGetLastError();
}

// geo comes from scan_plan(..., extra_reserve = batch_smem_bytes(n_prog)) over the UNION of the programs' columns
size_t batch_smem_bytes(int n_prog) { return ((static_cast<size_t>(n_prog) * sizeof(Program) + 127) & ~size_t(127)); }

cudaError_t batch_launch(const ScanLaunch &L, const ScanGeometry &geo, int n_prog, const Program *d_progs,
                         uint32_t *const *d_bitmaps, unsigned long long *d_counts, cudaStream_t stream) {
    if (n_prog < 1 || n_prog > kMaxBatch) return cudaErrorInvalidValue;
    BatchParams bp{};
    const cudaError_t e = fill_scan_params(L, geo, bp.s);
    if (e != cudaSuccess) return e;
    bp.s.dynamic_tiles = 0;
    bp.n_prog = n_prog;
    bp.progs = d_progs;
    for (int q = 0; q < n_prog; ++q) bp.bitmap[q] = d_bitmaps[q];
    bp.counts = d_counts;
    if (geo.n_tiles == 0) return cudaSuccess;
    switch (geo.tile_rows) {
        case 256: return launch_batch_r<kEvalWarpsWide, 1>(bp, geo, stream);
        case 512: return launch_batch_r<kEvalWarps, 1>(bp, geo, stream);
        case 1024: return launch_batch_r<kEvalWarps, 2>(bp, geo, stream);
        case 2048: return launch_batch_r<kEvalWarps, 4>(bp, geo, stream);
        case 4096: return launch_batch_r<kEvalWarps, 8>(bp, geo, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <int EW, int R, int CW>
static cudaError_t launch_fused_r(const FusedParams &fp, const ScanGeometry &geo, cudaStream_t stream) {
    static size_t allowed_by_device[kMaxDevices] = {0};
    size_t &allowed = allowed_by_device[current_device_slot()];
    if (geo.smem_bytes > allowed) {
        const cudaError_t e = cudaFuncSetAttribute(scan_fused_kernel<EW, R, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(geo.smem_bytes));
        if (e != cudaSuccess) return e;
        allowed = geo.smem_bytes;
    }
    scan_fused_kernel<EW, R, CW><<<geo.grid, 32 * (1 + EW + CW), geo.smem_bytes, stream>>>(fp);
    return cudaGetLastError();
}

template <int EW, int R>
static cudaError_t launch_fused_cw(const FusedParams &fp, const ScanGeometry &geo, cudaStream_t stream) {
    return geo.compact_warps == 8 ? launch_fused_r<EW, R, 8>(fp, geo, stream) : launch_fused_r<EW, R, 4>(fp, geo, stream);
}

cudaError_t fused_launch(const FusedLaunch &L, const ScanGeometry &geo, cudaStream_t stream) {
    FusedParams fp{};
    const cudaError_t e = fill_scan_params(L.scan, geo, fp.s);
    if (e != cudaSuccess) return e;
    if (geo.chunk_tiles < 1 || static_cast<long long>(geo.chunk_tiles) * geo.tile_rows > kFuseMaxChunkRows ||
        (geo.compact_warps != 4 && geo.compact_warps != 8))
        return cudaErrorInvalidValue;
    fp.s.dynamic_tiles = 0;
    fp.chunk_tiles = geo.chunk_tiles;
    fp.poll_ns = 256;  // 0 .. 1000 ns measured alike (2.075-2.092 ms on 1 B rows): any sleep that keeps the polls rare
    fp.n_chunks = geo.n_chunks;
    fp.desc = L.desc;
    fp.epoch = L.epoch;
    fp.id_base = L.id_base;
    fp.out_ids = L.out_ids;
    fp.out_cap = L.out_cap;
    fp.seg_chunks = L.progress ? L.seg_chunks : 0;
    fp.progress = L.progress;
    fp.host_count = L.host_count;
    fp.fctl = L.d_fctl;
    fp.prog = *L.scan.h_prog;
    if (!fp.fctl) return cudaErrorInvalidValue;
    if (fp.n_chunks == 0) return cudaSuccess;
    switch (geo.tile_rows) {
        case 256: return launch_fused_cw<kEvalWarpsWide, 1>(fp, geo, stream);
        case 512: return launch_fused_cw<kEvalWarps, 1>(fp, geo, stream);
        case 1024: return launch_fused_cw<kEvalWarps, 2>(fp, geo, stream);
        case 2048: return launch_fused_cw<kEvalWarps, 4>(fp, geo, stream);
        case 4096: return launch_fused_cw<kEvalWarps, 8>(fp, geo, stream);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------
// K1c: order-preserving compaction of the match bitmap into row ids
//
// One CTA per chunk of kChunkWords bitmap words (64 Ki rows), chunks claimed in order from an
// atomic counter (so a chunk's predecessors are always resident or finished: the look-back
// cannot deadlock).  Thread t owns words t, t + 256, ... of the chunk (coalesced loads); popc,
// warp shuffles and one shared-memory pass give every word its offset inside the chunk; warp 0
// publishes the chunk aggregate, looks back for the exclusive prefix and publishes the inclusive
// prefix; ids are scattered into a shared-memory stage and copied out with coalesced stores.
// ------------------------------------------------------------------------------------------
constexpr int kCompactThreads = 256;
constexpr int kCompactRounds = 8;                                    // words per thread
constexpr int kChunkWords = kCompactThreads * kCompactRounds;        // 2048 words = 65536 rows
constexpr int kStageIds = 8192;                                      // ids staged per copy-out (32 KB)
static_assert(kChunkWords * 32 == kCompactChunkRows, "chunk size is part of the launch interface");

struct CompactParams {
    const uint32_t *bitmap;
    long long n_words;
    long long n_chunks;
    QueryCtl *ctl;
    unsigned long long *desc;
    uint32_t epoch;
    uint32_t id_base;   // added to every row id (a shard's first global row); 0 for a whole table
    uint32_t *out_ids;  // may be PEER memory (another GPU's buffer mapped through CUDA IPC / NVLink)
    unsigned long long out_cap;  // ids the destination can hold: nothing is ever stored at or beyond it
};

__global__ void __launch_bounds__(kCompactThreads) compact_kernel(const __grid_constant__ CompactParams p) {
    __shared__ uint32_t s_stage[kStageIds];
    __shared__ uint32_t s_warp_tot[kCompactRounds][kCompactThreads / 32];
    __shared__ uint32_t s_round_base[kCompactRounds + 1];
    __shared__ long long s_chunk;
    __shared__ uint32_t s_excl;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr int kWarps = kCompactThreads / 32;
    if (tid == 0) s_chunk = static_cast<long long>(atomicAdd(&p.ctl->chunk_counter, 1u));
    __syncthreads();
    const long long chunk = s_chunk;
    if (chunk >= p.n_chunks) return;

    // 1. load this thread's words, popc, warp-inclusive scan per round
    uint32_t word[kCompactRounds], off[kCompactRounds];
    const long long w0 = chunk * kChunkWords;
#pragma unroll
    for (int r = 0; r < kCompactRounds; ++r) {
        const long long wi = w0 + r * kCompactThreads + tid;
        word[r] = wi < p.n_words ? __ldg(p.bitmap + wi) : 0u;
        const uint32_t pc = __popc(word[r]);
        uint32_t inc = pc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= static_cast<uint32_t>(d)) inc += t;
        }
        off[r] = inc - pc;  // exclusive within the warp
        if (lane == 31) s_warp_tot[r][warp] = inc;
    }
    __syncthreads();
    // 2. per-round bases (thread r of warp 0 sums round r), then chunk total + look-back
    if (warp == 0) {
        uint32_t rt = 0;
        if (lane < kCompactRounds) {
#pragma unroll
            for (int w = 0; w < kWarps; ++w) rt += s_warp_tot[lane][w];
        }
        uint32_t inc = rt;
#pragma unroll
        for (int d = 1; d < kCompactRounds; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= static_cast<uint32_t>(d)) inc += t;
        }
        if (lane < kCompactRounds) s_round_base[lane] = inc - rt;
        const uint32_t total = __shfl_sync(0xffffffffu, inc, kCompactRounds - 1);
        if (lane == 0) {
            s_round_base[kCompactRounds] = total;
            st_desc(p.desc + chunk, make_desc(p.epoch, chunk == 0 ? kStatePrefix : kStateAgg, total));
        }
        const uint32_t excl = warp_lookback(p.desc, chunk, p.epoch, lane);
        if (lane == 0) {
            st_desc(p.desc + chunk, make_desc(p.epoch, kStatePrefix, excl + total));
            s_excl = excl;
        }
    }
    __syncthreads();
    const uint32_t total = s_round_base[kCompactRounds];
    if (total == 0) return;
    const uint32_t excl = s_excl;
    // destination too small (the caller learns it from the match count): never store out of bounds
    if (static_cast<unsigned long long>(excl) + total > p.out_cap) return;
    // offsets inside the chunk: round base + earlier warps of the round + earlier lanes of the warp
#pragma unroll
    for (int r = 0; r < kCompactRounds; ++r) {
        uint32_t o = s_round_base[r] + off[r];
        for (uint32_t w = 0; w < warp; ++w) o += s_warp_tot[r][w];
        off[r] = o;
    }
    uint32_t *out = p.out_ids + excl;
    if (total <= kStageIds) {
        // sparse chunk: everything fits the stage -> one scatter, one coalesced copy
#pragma unroll
        for (int r = 0; r < kCompactRounds; ++r) {
            uint32_t w = word[r], o = off[r];
            const uint32_t row0 = static_cast<uint32_t>((w0 + r * kCompactThreads + tid) * 32) + p.id_base;
            while (w) {
                const int b = __ffs(w) - 1;
                w &= w - 1;
                s_stage[o++] = row0 + static_cast<uint32_t>(b);
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < total; i += kCompactThreads) out[i] = s_stage[i];
    } else {
        // dense chunk: a round (256 words, <= 8192 ids) at a time
        for (int r = 0; r < kCompactRounds; ++r) {
            const uint32_t base = s_round_base[r];
            const uint32_t cnt = (r + 1 < kCompactRounds ? s_round_base[r + 1] : total) - base;
            if (cnt == 0) continue;  // block-uniform
            uint32_t w = word[r], o = off[r] - base;
            const uint32_t row0 = static_cast<uint32_t>((w0 + r * kCompactThreads + tid) * 32) + p.id_base;
            while (w) {
                const int b = __ffs(w) - 1;
                w &= w - 1;
                s_stage[o++] = row0 + static_cast<uint32_t>(b);
            }
            __syncthreads();
            for (uint32_t i = tid; i < cnt; i += kCompactThreads) out[base + i] = s_stage[i];
            __syncthreads();
        }
    }
}

int64_t compact_chunks(long long n_words) { return (n_words + kChunkWords - 1) / kChunkWords; }

cudaError_t compact_launch(const uint32_t *bitmap, long long n_words, const QueryCtl *d_ctl, unsigned long long *desc,
                           uint32_t epoch, uint32_t *out_ids, uint32_t id_base, unsigned long long out_cap,
                           cudaStream_t stream, long long launch_chunks) {
    CompactParams p{};
    p.bitmap = bitmap;
    p.n_words = n_words;
    p.n_chunks = compact_chunks(n_words);
    p.ctl = const_cast<QueryCtl *>(d_ctl);
    p.desc = desc;
    p.epoch = epoch;
    p.id_base = id_base;
    p.out_ids = out_ids;
    p.out_cap = out_cap;
    if (p.n_chunks == 0) return cudaSuccess;
    // chunks are claimed from ctl->chunk_counter, which keeps counting across the launches of one
    // query: a pipelined scan launches the chunks of one table segment at a time (launch_chunks of
    // them), in order, on one stream -- the look-back of a later launch finds the earlier ones done
    const long long blocks = launch_chunks > 0 ? launch_chunks : p.n_chunks;
    compact_kernel<<<static_cast<unsigned>(blocks), kCompactThreads, 0, stream>>>(p);
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
