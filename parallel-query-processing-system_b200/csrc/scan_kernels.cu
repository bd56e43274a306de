// scan_kernels.cu -- SELECT/WHERE hot path for B200 (sm_100a)
//
// K1  scan_tma_kernel      full-table predicate evaluation -> match bitmap + match count.
//       Replaces linearSearchRecords/evaluateWhereClause/checkCondition
//       (engine/serial/executeEngine-serial.c:854-878, :292-316, :251-289).
//       Warp-specialised persistent CTAs, one per SM:
//         P  (1 warp, 1 lane)  streams every referenced column's slice of a tile (1024 / 512 / 256 rows) into
//                              shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx),
//                              up to 16 stages deep;
//         E  (16 warps)        each takes a WHOLE tile: a lane owns 32 / 16 / 8 consecutive rows, reads them with
//                              16-byte shared-memory loads and runs the compiled WHERE program once per tile;
//                              the lane's match mask is the bitmap word of its rows (bit b of word w = row 32w+b).
//       Nothing in K1 waits on another CTA, so HBM streaming is never stalled by a scan chain.
// K1c compact_kernel       order-preserving stream compaction of the bitmap into row ids:
//       popc per word, block scan, single-pass DECOUPLED LOOK-BACK over 64 Ki-row chunks
//       (status flag + aggregate / inclusive prefix per chunk), ids staged in shared memory and
//       written with coalesced stores -- rows come out in exactly table order.  The same bitmap
//       is DELETE's match mask.
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
#include <mutex>

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
// Every spin in this file is bounded by TIME: a protocol bug (or a peer that never arrives) ends in a trapped
// kernel and a CUDA error on the host after ~2 s, never in a hung GPU.
constexpr uint32_t kSpinLimit = 0x7fffffffu;
constexpr unsigned long long kSpinTimeoutNs = 2000000000ull;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xffu) == 0 && global_ns() - t0 > kSpinTimeoutNs) __trap();
    }
}
// The same wait for a thread that expects to wait LONG (the TMA producer waits ~1 us per tile for a stage to be
// released): sleep between polls.  A tight poll loop is 6 instructions per poll; ncu showed the single producer
// lane issuing 12 % of the whole kernel's instructions that way, on a scheduler it shares with four evaluator warps.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity, uint32_t sleep_ns) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(sleep_ns);
        if ((++spins & 0xffu) == 0 && global_ns() - t0 > kSpinTimeoutNs) __trap();
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
// the same with an L2 eviction policy: a table is read once per scan, so its lines are the first to go (evict_first)
// and what the query WRITES -- ids, the staging slice a sharded host result is copied from -- stays in L2
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                                  unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
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
// K1 / K1f / K9: TMA-staged scan with the VECTORISED evaluator
//
// Geometry.  A tile is what ONE evaluator warp consumes in one pass of the WHERE program: T = 32 * RPL
// consecutive rows (RPL = rows per lane = 32, 16 or 8; the largest that leaves >= 3 stages in flight), each lane
// owning RPL CONSECUTIVE rows.  The producer lane streams a tile's slice of every referenced column into one of
// S <= 16 shared-memory stages (1-D TMA bulk copies, one mbarrier pair per stage).  STAGE s BELONGS TO EVALUATOR
// WARP s: tile `seq` of a CTA goes to stage / warp seq % S, so every barrier is produced and consumed strictly in
// order by one producer and one warp (a stage shared by several warps that run at their own pace cannot be told
// apart by phase parity: a fast warp would take the barrier's previous phase for its own).  While a warp
// evaluates its tile the other stages are in flight or being evaluated by their warps.  (Round 1 had one 4096-row tile shared by all 16 warps, 8 rows per lane with a stride of 32 rows, one
// scalar shared-memory load per row and leaf, the match bits transposed with one ballot per row slot: ncu showed
// 1.28 warp-instructions per row, ALU pipe 58 % busy, QN at 0.88 of what the same kernel read on wide rows.)
//
// Why consecutive rows per lane: (1) a lane's match mask IS the bitmap word of its rows (bit b of word w = row
// 32 w + b) -- no transposition; (2) cells are read with 16-byte shared-memory loads (4 int / 2 u64 / 16 bool cells
// per LDS.128); (3) the program is dispatched once per 1024 rows instead of once per 256 / 512.
// Bank conflicts: lane blocks are RPL * width bytes apart, so reading "unit k" (4 rows) on every lane at the same
// time would hit the same banks 8 ways.  Lane L therefore reads its units in ROTATED order, unit (k + L) mod U in
// step k, which makes the 16-byte loads of an int column conflict-free (2-way for u64); text cells rotate by rows.
// The masks are built in that rotated "slot order" with immediate bit positions and rotated back ONCE per pass.
// Compare + bit insert is two instructions per row and leaf (ISETP, predicated OR; inline PTX, because the
// compiler's own choice was SEL + IADD3).
// ------------------------------------------------------------------------------------------
constexpr int kEvalWarps = 16;                   // evaluator warps per CTA (4 per scheduler)
constexpr int kMaxStages = 16;
constexpr int kMaxTileRows = 1024;               // 32 lanes x 32 rows (<= kRowPad: a full tile is always inside the allocation)

struct ScanParams {
    const uint8_t *col[NUM_COLS];
    uint32_t width[NUM_COLS];
    uint32_t smem_off[NUM_COLS];  // byte offset of the column inside a stage
    int32_t n_ref;
    int32_t ref_col[NUM_COLS];
    uint32_t stage_bytes;
    int32_t tile_rows;
    int32_t n_stages;
    int32_t pad_;
    long long n_rows;
    long long tile_begin;   // this launch covers tiles [tile_begin, n_tiles): a pipelined scan launches
    long long n_tiles;      // K1 once per table segment so that K1c of segment i overlaps K1 of segment i+1
    QueryCtl *ctl;
    uint32_t *out_bitmap;   // n_tiles * tile_rows / 32 words, or null (count only)
};

struct ScanSmemHeader {
    Program prog;
    alignas(8) uint64_t full[kMaxStages];
    unsigned long long cta_count;
};

// After the program has been copied into shared memory: give every leaf its column's stage offset and width, so
// that evaluating a leaf needs nothing but the leaf record (call between two CTA-wide barriers).
__device__ __forceinline__ void patch_leaves(Program *sp, const ScanParams &p, uint32_t tid, uint32_t n_threads) {
    for (uint32_t k = tid; k < static_cast<uint32_t>(sp->n_leaves); k += n_threads) {
        const int c = sp->leaf[k].col;
        sp->leaf[k].smem_off = p.smem_off[c];
        sp->leaf[k].nch = static_cast<uint8_t>(p.width[c] >> 4);
    }
}

// ---- compare + predicated OR: m |= bit when (x OP lit).  OP is the leaf's 3-bit truth table. ----
#define QPE_CMP_OR_S32(OPNAME)                                                                                  \
    asm("{\n\t.reg .pred p;\n\tsetp." OPNAME ".s32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}"                   \
        : "+r"(m)                                                                                               \
        : "r"(x), "r"(lit), "r"(bit))
#define QPE_CMP_OR_U64(OPNAME)                                                                                  \
    asm("{\n\t.reg .pred p;\n\tsetp." OPNAME ".u64 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}"                   \
        : "+r"(m)                                                                                               \
        : "l"(x), "l"(lit), "r"(bit))

template <int OP>
__device__ __forceinline__ void cmp_or(uint32_t &m, int32_t x, int32_t lit, uint32_t bit) {
    if constexpr (OP == 0b010) QPE_CMP_OR_S32("eq");
    else if constexpr (OP == 0b101) QPE_CMP_OR_S32("ne");
    else if constexpr (OP == 0b100) QPE_CMP_OR_S32("gt");
    else if constexpr (OP == 0b001) QPE_CMP_OR_S32("lt");
    else if constexpr (OP == 0b110) QPE_CMP_OR_S32("ge");
    else QPE_CMP_OR_S32("le");
}
template <int OP>
__device__ __forceinline__ void cmp_or(uint32_t &m, unsigned long long x, unsigned long long lit, uint32_t bit) {
    if constexpr (OP == 0b010) QPE_CMP_OR_U64("eq");
    else if constexpr (OP == 0b101) QPE_CMP_OR_U64("ne");
    else if constexpr (OP == 0b100) QPE_CMP_OR_U64("hi");
    else if constexpr (OP == 0b001) QPE_CMP_OR_U64("lo");
    else if constexpr (OP == 0b110) QPE_CMP_OR_U64("hs");
    else QPE_CMP_OR_U64("ls");
}
// m |= bit when x != 0
__device__ __forceinline__ void nz_or(uint32_t &m, uint32_t x, uint32_t bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(m) : "r"(x), "r"(bit));
}

__device__ __forceinline__ uint32_t diff16(const uint4 v, const uint4 l) {
    return (v.x ^ l.x) | (v.y ^ l.y) | (v.z ^ l.z) | (v.w ^ l.w);
}

template <int RPL>
__device__ __forceinline__ uint32_t rotl_rpl(uint32_t m, uint32_t s) {  // rotate left by s (0 <= s < RPL) within RPL bits
    if constexpr (RPL == 32) {
        return __funnelshift_l(m, m, s);
    } else {
        constexpr uint32_t all = (1u << RPL) - 1u;
        return ((m << s) | (m >> (RPL - s))) & all;
    }
}

// Per-lane constants of the rotated access order (computed once per kernel).
//   numeric cells: step k reads unit (k + lane) mod U (unit = 4 consecutive rows); slot bit b <-> row (b + rot4) mod RPL
//   text cells:    step j reads row  (j + lane) mod RPL;                            slot bit j <-> row (j + rotr) mod RPL
template <int RPL>
struct LaneGeom {
    static constexpr int U = RPL / 4;
    uint32_t uoff[U];   // 16 * ((k + lane) mod U): byte offset of step k's unit in an int column's lane block
    uint32_t rot4;      // 4 * (lane mod U)
    uint32_t rotr;      // lane mod RPL
    uint32_t trot;      // text slot order -> numeric slot order: rotate left by (rotr - rot4) mod RPL
    __device__ __forceinline__ void init(uint32_t lane) {
#pragma unroll
        for (int k = 0; k < U; ++k) uoff[k] = ((static_cast<uint32_t>(k) + lane) & (U - 1)) << 4;
        rot4 = (lane & (U - 1)) << 2;
        rotr = lane & (RPL - 1);
        trot = (rotr - rot4) & (RPL - 1);
    }
};

template <int RPL, int OP>
__device__ __forceinline__ uint32_t leaf_i32(const uint8_t *lb, const LaneGeom<RPL> &g, int32_t lit) {
    uint32_t m0 = 0, m1 = 0;  // two accumulators: half the length of the predicated-OR dependency chain
#pragma unroll
    for (int k = 0; k < RPL / 4; ++k) {
        const int4 v = *reinterpret_cast<const int4 *>(lb + g.uoff[k]);
        uint32_t &m = (k & 1) ? m1 : m0;
        cmp_or<OP>(m, v.x, lit, 1u << (4 * k));
        cmp_or<OP>(m, v.y, lit, 2u << (4 * k));
        cmp_or<OP>(m, v.z, lit, 4u << (4 * k));
        cmp_or<OP>(m, v.w, lit, 8u << (4 * k));
    }
    return m0 | m1;
}

template <int RPL, int OP>
__device__ __forceinline__ uint32_t leaf_u64(const uint8_t *lb, const LaneGeom<RPL> &g, unsigned long long lit) {
    uint32_t m0 = 0, m1 = 0;
#pragma unroll
    for (int k = 0; k < RPL / 4; ++k) {
        const uint8_t *a = lb + 2u * g.uoff[k];
        const ulonglong2 v0 = *reinterpret_cast<const ulonglong2 *>(a);
        const ulonglong2 v1 = *reinterpret_cast<const ulonglong2 *>(a + 16);
        uint32_t &m = (k & 1) ? m1 : m0;
        cmp_or<OP>(m, v0.x, lit, 1u << (4 * k));
        cmp_or<OP>(m, v0.y, lit, 2u << (4 * k));
        cmp_or<OP>(m, v1.x, lit, 4u << (4 * k));
        cmp_or<OP>(m, v1.y, lit, 8u << (4 * k));
    }
    return m0 | m1;
}

#define QPE_OP_SWITCH(tt, FN, ...)                                  \
    switch (tt) {                                                   \
        case 0b010: return FN<RPL, 0b010>(__VA_ARGS__);             \
        case 0b101: return FN<RPL, 0b101>(__VA_ARGS__);             \
        case 0b100: return FN<RPL, 0b100>(__VA_ARGS__);             \
        case 0b001: return FN<RPL, 0b001>(__VA_ARGS__);             \
        case 0b110: return FN<RPL, 0b110>(__VA_ARGS__);             \
        case 0b011: return FN<RPL, 0b011>(__VA_ARGS__);             \
        case 0b111: return kAll;                                    \
        default: return 0u;                                         \
    }

// four cells of 0 / 1 bytes -> their low bits as a nibble (no carries: every product term lands on its own bit)
__device__ __forceinline__ uint32_t bool4_nibble(uint32_t x) { return (x * 0x01020408u) >> 24; }

// bool column: cells are exactly 0 or 1 (every writer of a device table normalises them); rows are read in ROW order
// (the whole lane block is 32 / 16 / 8 bytes) and the mask is rotated into slot order at the end
template <int RPL>
__device__ __forceinline__ uint32_t leaf_bool(const uint8_t *lb, const LaneGeom<RPL> &g, bool want) {
    constexpr uint32_t kAll = (RPL >= 32) ? 0xffffffffu : ((1u << RPL) - 1u);
    uint32_t rows;
    if constexpr (RPL == 32) {
        const uint4 a = *reinterpret_cast<const uint4 *>(lb);
        const uint4 b = *reinterpret_cast<const uint4 *>(lb + 16);
        rows = bool4_nibble(a.x) | (bool4_nibble(a.y) << 4) | (bool4_nibble(a.z) << 8) | (bool4_nibble(a.w) << 12) |
               (bool4_nibble(b.x) << 16) | (bool4_nibble(b.y) << 20) | (bool4_nibble(b.z) << 24) | (bool4_nibble(b.w) << 28);
    } else if constexpr (RPL == 16) {
        const uint4 a = *reinterpret_cast<const uint4 *>(lb);
        rows = bool4_nibble(a.x) | (bool4_nibble(a.y) << 4) | (bool4_nibble(a.z) << 8) | (bool4_nibble(a.w) << 12);
    } else {
        const uint2 a = *reinterpret_cast<const uint2 *>(lb);
        rows = bool4_nibble(a.x) | (bool4_nibble(a.y) << 4);
    }
    if (!want) rows = ~rows & kAll;
    // row order -> slot order: slot bit b holds row (b + rot4) mod RPL, i.e. rotate RIGHT by rot4
    return rotl_rpl<RPL>(rows, (RPL - g.rot4) & (RPL - 1));
}

// text column (cell = nch x 16 bytes, NUL padded): strcmp order == unsigned byte order over the cell.
// Step j reads row (j + lane) mod RPL of the lane block (conflict-free 16-byte loads for 16-byte cells).
template <int RPL>
__device__ __forceinline__ uint32_t leaf_text(const PLeaf &lf, const Program *sp, const uint8_t *col_stage,
                                              uint32_t lane, const LaneGeom<RPL> &g) {
    constexpr uint32_t kAll = (RPL >= 32) ? 0xffffffffu : ((1u << RPL) - 1u);
    const int nch = lf.nch;
    const uint32_t w = static_cast<uint32_t>(nch) << 4;
    const uint32_t span = RPL * w;
    const uint4 *lit = reinterpret_cast<const uint4 *>(sp->lit_pool + lf.lit_off);
    const uint8_t *lb = col_stage + lane * span;
    const uint32_t start = g.rotr * w;
    const uint32_t tt = lf.tt;
    // byte offset of step j's row inside the lane block (j is a constant after unrolling)
    auto row_off = [&](int j) -> uint32_t {
        uint32_t o = start + static_cast<uint32_t>(j) * w;
        return o >= span ? o - span : o;
    };
    uint32_t res;
    if (tt == 0b010u || tt == 0b101u) {
        // equality only: OR of XORs, no byte swapping
        uint32_t ne = 0;
        if (nch == 1) {
            const uint4 l0 = lit[0];
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                const uint32_t o = (start + 16u * j) & (RPL * 16u - 1u);
                nz_or(ne, diff16(*reinterpret_cast<const uint4 *>(lb + o), l0), 1u << j);
            }
        } else {
            for (int c = 0; c < nch; ++c) {
                const uint4 lc = lit[c];
#pragma unroll
                for (int j = 0; j < RPL; ++j)
                    nz_or(ne, diff16(*reinterpret_cast<const uint4 *>(lb + row_off(j) + 16u * c), lc), 1u << j);
                if (__all_sync(0xffffffffu, ne == kAll)) break;  // warp-uniform early exit
            }
        }
        res = (tt == 0b010u) ? (~ne & kAll) : ne;
    } else {
        // ordering: three-way compare chunk by chunk, first differing chunk decides
        uint32_t lt = 0, decided = 0;
        for (int c = 0; c < nch; ++c) {
            const uint4 lc = lit[c];
            const uint32_t b0 = bswap32(lc.x), b1 = bswap32(lc.y), b2 = bswap32(lc.z), b3 = bswap32(lc.w);
            uint32_t dk = 0, ltk = 0;
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                const uint4 v = *reinterpret_cast<const uint4 *>(lb + row_off(j) + 16u * c);
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
        res = ((tt & 1u) ? lt : 0u) | ((tt & 2u) ? eq : 0u) | ((tt & 4u) ? gt : 0u);
    }
    return rotl_rpl<RPL>(res, g.trot);
}

template <int RPL>
__device__ __forceinline__ uint32_t leaf_numeric_i32(const uint8_t *lb, const LaneGeom<RPL> &g, int32_t lit, uint32_t tt) {
    constexpr uint32_t kAll = (RPL >= 32) ? 0xffffffffu : ((1u << RPL) - 1u);
    QPE_OP_SWITCH(tt, leaf_i32, lb, g, lit)
}
template <int RPL>
__device__ __forceinline__ uint32_t leaf_numeric_u64(const uint8_t *lb, const LaneGeom<RPL> &g, unsigned long long lit,
                                                     uint32_t tt) {
    constexpr uint32_t kAll = (RPL >= 32) ? 0xffffffffu : ((1u << RPL) - 1u);
    QPE_OP_SWITCH(tt, leaf_u64, lb, g, lit)
}

// One leaf over this lane's RPL rows of the staged tile; the result is in numeric SLOT order.
template <int RPL>
__device__ __forceinline__ uint32_t eval_leaf(const PLeaf &lf, const Program *sp, const uint8_t *stage, uint32_t lane,
                                              const LaneGeom<RPL> &g) {
    const uint8_t *base = stage + lf.smem_off;  // patched by patch_leaves
    switch (lf.type) {
        case T_I32: return leaf_numeric_i32<RPL>(base + lane * (RPL * 4u), g, lf.lit_i32, lf.tt);
        case T_U64: return leaf_numeric_u64<RPL>(base + lane * (RPL * 8u), g, lf.lit_u64, lf.tt);
        case T_BOOL:  // only = and != exist; (cell != 0) == want
            return leaf_bool<RPL>(base + lane * RPL, g, (lf.tt == 0b010u) == ((lf.lit_i32 & 1) != 0));
        default: return leaf_text<RPL>(lf, sp, base, lane, g);
    }
}

// run the compiled WHERE program; returns the match mask (whatever order the leaves use)
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

// The whole WHERE over one staged tile: this lane's RPL rows as a bitmap (bit b = row first_row + b), rows at or
// beyond n_rows cleared.  first_row = global row of the lane's first row.
template <int RPL>
__device__ __forceinline__ uint32_t eval_tile(const Program *sp, const uint8_t *stage, uint32_t lane,
                                              const LaneGeom<RPL> &g, long long first_row, long long n_rows) {
    constexpr uint32_t kAll = (RPL >= 32) ? 0xffffffffu : ((1u << RPL) - 1u);
    uint32_t acc = run_program(sp, kAll, [&](const PLeaf &lf) { return eval_leaf<RPL>(lf, sp, stage, lane, g); });
    acc = rotl_rpl<RPL>(acc, g.rot4);  // slot order -> row order
    const long long left = n_rows - first_row;
    if (left < RPL) acc = left <= 0 ? 0u : (acc & ((1u << static_cast<uint32_t>(left)) - 1u));
    return acc;
}

// a lane's RPL match bits into a bitmap whose first byte holds the tile's first row (shared or global memory)
template <int RPL>
__device__ __forceinline__ void store_mask(uint8_t *tile_bitmap, uint32_t lane, uint32_t mask) {
    if constexpr (RPL == 32) reinterpret_cast<uint32_t *>(tile_bitmap)[lane] = mask;
    else if constexpr (RPL == 16) reinterpret_cast<uint16_t *>(tile_bitmap)[lane] = static_cast<uint16_t>(mask);
    else tile_bitmap[lane] = static_cast<uint8_t>(mask);
}

// producer: the slices of one tile into a stage
__device__ __forceinline__ void produce_tile(const ScanParams &p, uint8_t *dst, long long tile, int T, uint64_t *bar) {
    mbar_arrive_expect_tx(bar, p.stage_bytes);
    for (int r = 0; r < p.n_ref; ++r) {
        const int c = p.ref_col[r];
        const uint32_t w = p.width[c];
        tma_bulk_g2s(dst + p.smem_off[c], p.col[c] + static_cast<size_t>(tile) * T * w, static_cast<uint32_t>(T) * w, bar);
    }
}
// (K1f) with the streaming L2 policy; policy == 0: plain copies
__device__ __forceinline__ void produce_tile_stream(const ScanParams &p, uint8_t *dst, long long tile, int T, uint64_t *bar,
                                                    unsigned long long policy) {
    if (policy == 0ull) {
        produce_tile(p, dst, tile, T, bar);
        return;
    }
    mbar_arrive_expect_tx(bar, p.stage_bytes);
    for (int r = 0; r < p.n_ref; ++r) {
        const int c = p.ref_col[r];
        const uint32_t w = p.width[c];
        tma_bulk_g2s_hint(dst + p.smem_off[c], p.col[c] + static_cast<size_t>(tile) * T * w, static_cast<uint32_t>(T) * w, bar,
                          policy);
    }
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
            unsigned long long t0 = 0;
            while ((flag >> 2) != epoch) {
                __nanosleep(64);  // predecessor not published yet: back off instead of hammering L2
                if ((++spins & 0xffu) == 0) {
                    if (t0 == 0) t0 = global_ns();
                    else if (global_ns() - t0 > kSpinTimeoutNs) __trap();
                }
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

// ------------------------------------------------------------------------------------------
// K1: scan -> match bitmap (optional) + match count.  Persistent, one CTA per SM; tile seq of CTA b is table
// tile tile_begin + b + seq * grid, taken by warp seq % S.  Nothing waits on another CTA.
//
// SELF-FEEDING WARPS.  There is no producer warp: evaluator warp w owns stage w and its mbarrier, and its lane 0
// issues the bulk copies of the warp's NEXT tile as soon as the warp has finished reading the current one.  (With
// one producer lane for the whole CTA, 1024-row tiles were bound by that lane: ~0.43 us per tile whatever its
// size.)  While a warp waits for its refill, the other warps evaluate; about S - 3 stages are in flight per SM.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int RPL>
__global__ void __launch_bounds__(32 * kEvalWarps, 1) scan_tma_kernel(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ScanSmemHeader *sh = reinterpret_cast<ScanSmemHeader *>(smem_raw);
    uint8_t *stages = smem_raw + ((sizeof(ScanSmemHeader) + 127) & ~size_t(127));
    constexpr int EW = kEvalWarps;
    constexpr int T = 32 * RPL;
    constexpr uint32_t kThreads = 32 * EW;
    const uint32_t tid = threadIdx.x, ew = tid >> 5, lane = tid & 31u;
    const uint32_t S = static_cast<uint32_t>(p.n_stages);

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) mbar_init(&sh->full[s], 1);
        sh->cta_count = 0;
        fence_mbar_init();
    }
    __syncthreads();
    // first tile of every warp: in flight while the program is copied into shared memory
    uint8_t *stage = stages + static_cast<size_t>(ew) * p.stage_bytes;
    const long long step = static_cast<long long>(S) * gridDim.x;
    long long tile = p.tile_begin + blockIdx.x + static_cast<long long>(ew) * gridDim.x;
    const bool active = ew < S;
    if (active && lane == 0 && tile < p.n_tiles) produce_tile(p, stage, tile, T, &sh->full[ew]);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&p.ctl->prog);
        uint4 *dst = reinterpret_cast<uint4 *>(&sh->prog);
        for (uint32_t i = tid; i < sizeof(Program) / 16; i += kThreads) dst[i] = src[i];
    }
    __syncthreads();
    patch_leaves(&sh->prog, p, tid, kThreads);
    __syncthreads();
    const Program *sp = &sh->prog;

    LaneGeom<RPL> g;
    g.init(lane);
    uint32_t my_count = 0;
    if (active)
        for (uint32_t it = 0; tile < p.n_tiles; tile += step, ++it) {
            mbar_wait(&sh->full[ew], it & 1u);
            const uint32_t acc = eval_tile<RPL>(sp, stage, lane, g, tile * T + static_cast<long long>(lane) * RPL, p.n_rows);
            __syncwarp();  // every lane has read the stage: refill it
            if (lane == 0 && tile + step < p.n_tiles) {
                fence_proxy_async_smem();
                produce_tile(p, stage, tile + step, T, &sh->full[ew]);
            }
            my_count += static_cast<uint32_t>(__popc(acc));
            if (p.out_bitmap) store_mask<RPL>(reinterpret_cast<uint8_t *>(p.out_bitmap) + tile * (T / 8), lane, acc);
        }
    const uint32_t warp_total = __reduce_add_sync(0xffffffffu, my_count);
    if (lane == 0 && warp_total) atomicAdd(&sh->cta_count, static_cast<unsigned long long>(warp_total));
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

template <int RPL>
__global__ void __launch_bounds__(32 * kEvalWarps, 1) scan_batch_kernel(const __grid_constant__ BatchParams bp) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const ScanParams &p = bp.s;
    ScanSmemHeader *sh = reinterpret_cast<ScanSmemHeader *>(smem_raw);  // its own prog slot is unused here
    __shared__ unsigned long long s_count[kMaxBatch];
    Program *progs = reinterpret_cast<Program *>(smem_raw + ((sizeof(ScanSmemHeader) + 127) & ~size_t(127)));
    const int Q = bp.n_prog;
    uint8_t *stages = reinterpret_cast<uint8_t *>(progs) + ((static_cast<size_t>(Q) * sizeof(Program) + 127) & ~size_t(127));
    constexpr int EW = kEvalWarps;
    constexpr int T = 32 * RPL;
    constexpr uint32_t kThreads = 32 * EW;
    const uint32_t tid = threadIdx.x, ew = tid >> 5, lane = tid & 31u;
    const uint32_t S = static_cast<uint32_t>(p.n_stages);
    if (tid < kMaxBatch) s_count[tid] = 0;
    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) mbar_init(&sh->full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint8_t *stage = stages + static_cast<size_t>(ew) * p.stage_bytes;
    const long long step = static_cast<long long>(S) * gridDim.x;
    long long tile = blockIdx.x + static_cast<long long>(ew) * gridDim.x;
    const bool active = ew < S;
    if (active && lane == 0 && tile < p.n_tiles) produce_tile(p, stage, tile, T, &sh->full[ew]);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(bp.progs);
        uint4 *dst = reinterpret_cast<uint4 *>(progs);
        for (uint32_t i = tid; i < static_cast<uint32_t>(Q) * (sizeof(Program) / 16); i += kThreads) dst[i] = src[i];
    }
    __syncthreads();
    for (int q = 0; q < Q; ++q) patch_leaves(progs + q, p, tid, kThreads);
    __syncthreads();

    LaneGeom<RPL> g;
    g.init(lane);
    if (active)
        for (uint32_t it = 0; tile < p.n_tiles; tile += step, ++it) {
            mbar_wait(&sh->full[ew], it & 1u);
            const long long first_row = tile * T + static_cast<long long>(lane) * RPL;
#pragma unroll 1
            for (int q = 0; q < Q; ++q) {
                const uint32_t a = eval_tile<RPL>(progs + q, stage, lane, g, first_row, p.n_rows);
                const uint32_t warp_cnt = __reduce_add_sync(0xffffffffu, static_cast<uint32_t>(__popc(a)));
                if (lane == 0 && warp_cnt) atomicAdd(&s_count[q], static_cast<unsigned long long>(warp_cnt));
                store_mask<RPL>(reinterpret_cast<uint8_t *>(bp.bitmap[q]) + tile * (T / 8), lane, a);
            }
            __syncwarp();  // every program has read the stage: refill it
            if (lane == 0 && tile + step < p.n_tiles) {
                fence_proxy_async_smem();
                produce_tile(p, stage, tile + step, T, &sh->full[ew]);
            }
        }
    __syncthreads();
    if (tid < static_cast<uint32_t>(Q) && s_count[tid]) atomicAdd(&bp.counts[tid], s_count[tid]);
}

// ------------------------------------------------------------------------------------------
// K1f: fused scan + ordered compaction (one launch, ids leave the SM while the scan is running)
//
// Self-feeding evaluator warps as in K1, plus CW COMPACTION warps per CTA.  Work is handed out in CHUNKS of
// chunk_tiles consecutive tiles (<= 64 Ki rows), each CTA taking the next unclaimed chunk (tickets), so in increasing order; tile j of a CTA's
// k-th chunk is its tile seq = k * chunk_tiles + j and belongs to warp / stage seq % S.  The evaluators leave
// the chunk's match bitmap in shared memory (double buffered); the compaction warps popc/scan it, publish the
// chunk aggregate, run the DECOUPLED LOOK-BACK over the chunk descriptors of the other CTAs (all resident:
// grid <= SM count, every CTA walks its chunks in increasing order, so the smallest unfinished chunk never
// waits), expand the bits into row ids in shared memory and store them coalesced at out_ids + prefix -- while the
// evaluators are already two chunks ahead.  Compared with K1 -> K1c: no second launch, no bitmap round trip
// through HBM, no barrier-bound compaction pass after the scan.
// The look-back granularity stays one descriptor per chunk (a per-tile chain cannot keep up with HBM),
// but the chain advances beside the scan instead of after it.
// Optional progress words in pinned host memory tell the host when a table segment's ids are complete in HBM, so
// it can start the device->host copy of that segment during the scan.
// CW = compaction warps per CTA, 4 or 8.  Four are plenty while few rows match; from ~10 % selectivity the
// evaluators wait for the compaction warps, and eight take over (the engine picks per query, engine.cu).
// ------------------------------------------------------------------------------------------
constexpr int kFuseChunkWords = kFuseMaxChunkRows / 32;              // 2048 words
static_assert(fuse_reserve_bytes(4) == 2 * kFuseChunkWords * 4 + 32 * 32 * 4 * 4, "smem reserve");
static_assert(fuse_reserve_bytes(8) == 2 * kFuseChunkWords * 4 + 32 * 32 * 8 * 4, "smem reserve");

struct FusedParams {
    ScanParams s;
    int32_t chunk_tiles;            // tiles per chunk; chunk_tiles * tile_rows <= kFuseMaxChunkRows
    uint32_t poll_ns;               // sleep between the compaction warps' polls of the chunk barrier
    long long n_chunks;
    uint32_t l2_stream;             // table tiles are fetched with the L2 evict_first policy
    uint32_t pad_;
    unsigned long long *desc;       // one look-back descriptor per chunk
    uint32_t epoch;
    uint32_t id_base;
    uint32_t *out_ids;              // HBM of this GPU, a peer mapping, or mapped pinned host memory
    unsigned long long out_cap;
    long long seg_chunks;           // progress: chunks per table segment (0 = no progress words)
    unsigned long long *progress;   // mapped pinned host memory: [seg] = epoch << 32 | ids complete through seg
    unsigned long long *host_count; // mapped pinned host memory: the match count, written by the last CTA (or null)
    FusedCtl *fctl;                 // this kernel's own control words (self-resetting)
    unsigned long long *trace;      // diagnostics (QPE_FUSE_TRACE=1): 8 words per CTA of %globaltimer stamps, or null
    Program prog;                   // the compiled WHERE, by value: no upload precedes the launch
};

struct FusedSmemHeader {   // sized for the 4-warp variant (16 rounds x 4 warps); the 8-warp one needs 8 x 8
    Program prog;
    alignas(8) uint64_t full[kMaxStages];
    uint64_t cb_full[2];
    uint64_t cb_empty[2];
    unsigned long long cta_count;
    uint32_t warp_tot[64];          // [round][compaction warp]
    uint32_t round_base[16 + 1];
    uint32_t excl;
    // the chunks this CTA has claimed: [k & 7] = (k + 1) << 32 | chunk index of its k-th chunk.  Evaluator warps are
    // never more than three chunks apart and the compaction warps at most two behind them, so eight slots suffice.
    alignas(8) unsigned long long chunk_ring[8];
    // per chunk buffer: matches counted by the evaluator warps that have finished the chunk, and how many have.  The
    // LAST warp to finish a chunk publishes the chunk's aggregate itself (see the evaluator loop).
    uint32_t chunk_cnt[2];
    uint32_t chunk_arr[2];
};

// A warp's walk over ITS tiles of the CTA's chunks, in order: tile j of the CTA's k-th chunk has the CTA-wide
// sequence number k * CT + j, and the tiles with seq % S == w belong to warp w.
// WHICH chunk is the CTA's k-th is decided at run time: chunk b for k = 0, after that the next unclaimed one (a ticket
// from FusedCtl::next_chunk).  SMs do not all stream at the same rate (+-3 % on a B200: r2_k1f_cta_trace.txt, static
// assignment left the fast ones idle for the last ~25 us of a 125 M-row scan); tickets keep every SM busy to the end.
// A CTA's chunks still increase with k and all CTAs are resident, so the look-back never waits for an unstarted chunk:
// the smallest unfinished chunk is either being worked on or is the next one of a CTA that only waits for smaller ones.
struct WarpTiles {
    long long chunk;   // current chunk (>= n_chunks: exhausted)
    uint32_t k;        // its index among this CTA's chunks
    uint32_t j;        // current tile within the chunk
    uint32_t nt;       // tiles in the current chunk
};
__device__ __forceinline__ uint32_t chunk_tiles_of(long long chunk, uint32_t CT, long long n_tiles) {
    const long long left = n_tiles - chunk * CT;
    return static_cast<uint32_t>(left < CT ? left : CT);
}
// first tile of warp w in the CTA's k-th chunk
__device__ __forceinline__ uint32_t first_tile_of(uint32_t w, uint32_t k, uint32_t CT, uint32_t S) {
    return (w + S - (k * CT) % S) % S;
}
// the CTA's k-th chunk (k >= 1 was claimed by the warp that fetched the claim tile of chunk k - 1, some tiles ago; the wait is
// bounded by time like every spin of this file)
__device__ __forceinline__ long long chunk_of(const unsigned long long *ring, uint32_t k) {
    const volatile unsigned long long *slot = ring + (k & 7u);
    unsigned long long v = *slot;
    if (static_cast<uint32_t>(v >> 32) != k + 1u) {
        const unsigned long long t0 = global_ns();
        do {
            __nanosleep(20);
            v = *slot;
            if (global_ns() - t0 > 2000000000ull) __trap();
        } while (static_cast<uint32_t>(v >> 32) != k + 1u);
    }
    return static_cast<long long>(static_cast<uint32_t>(v));
}
// The tile of a chunk whose fetch triggers the claim of the CTA's next chunk: two rounds of the stages before the
// chunk's end (~8 us ahead of the first warp that needs the answer, the ticket takes ~1-2 us) rather than its first
// tile -- a chunk claimed early is a chunk a faster SM cannot take at the end of the scan.
__device__ __forceinline__ uint32_t claim_tile_of(uint32_t nt, uint32_t S) { return nt > 2u * S ? nt - 2u * S : 0u; }
// one lane: take the next ticket for the CTA's k-th chunk and publish it to the CTA
__device__ __forceinline__ void claim_chunk(unsigned long long *ring, uint32_t k, FusedCtl *fc, uint32_t grid) {
    const uint32_t c = atomicAdd(&fc->next_chunk, 1u) + grid;
    *reinterpret_cast<volatile unsigned long long *>(ring + (k & 7u)) = (static_cast<unsigned long long>(k + 1u) << 32) | c;
}
// position `wt` on this warp's first tile at or after (chunk, j); false when there is none
__device__ __forceinline__ bool settle(WarpTiles &wt, uint32_t w, uint32_t CT, uint32_t S, long long n_chunks,
                                       long long n_tiles, const unsigned long long *ring) {
    while (wt.chunk < n_chunks) {
        if (wt.j < wt.nt) return true;
        ++wt.k;
        wt.chunk = chunk_of(ring, wt.k);
        if (wt.chunk >= n_chunks) break;
        wt.nt = chunk_tiles_of(wt.chunk, CT, n_tiles);
        wt.j = first_tile_of(w, wt.k, CT, S);
    }
    return false;
}

template <int RPL, int CW>
__global__ void __launch_bounds__(32 * (kEvalWarps + CW), 1)
    scan_fused_kernel(const __grid_constant__ FusedParams fp) {
    constexpr int EW = kEvalWarps;
    constexpr int kFuseCompactWarps = CW;
    constexpr int kFuseCompactThreads = 32 * CW;                           // 128 | 256
    constexpr int kFuseRounds = kFuseChunkWords / kFuseCompactThreads;     // 16 | 8 words per compaction thread
    constexpr int kFuseStageIds = 32 * kFuseCompactThreads;                // a round never holds more ids
    static_assert(kFuseRounds * CW <= 64 && kFuseRounds <= 16, "FusedSmemHeader scratch");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const ScanParams &p = fp.s;
    FusedSmemHeader *sh = reinterpret_cast<FusedSmemHeader *>(smem_raw);
    uint8_t *stages = smem_raw + ((sizeof(FusedSmemHeader) + 127) & ~size_t(127));
    const uint32_t S = static_cast<uint32_t>(p.n_stages);
    uint32_t *cbuf = reinterpret_cast<uint32_t *>(stages + static_cast<size_t>(S) * p.stage_bytes);  // [2][2048]
    uint32_t *id_stage = cbuf + 2 * kFuseChunkWords;                                                    // [32 * 32 * CW]

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31u;
    constexpr int T = 32 * RPL;      // rows per tile
    constexpr int WPT = T >> 5;      // bitmap words per tile
    constexpr uint32_t kThreads = 32 * (EW + kFuseCompactWarps);
    const uint32_t CT = static_cast<uint32_t>(fp.chunk_tiles);

    if (tid == 0) {
        if (fp.trace) fp.trace[blockIdx.x * 8 + 0] = global_ns();
        for (uint32_t s = 0; s < S; ++s) mbar_init(&sh->full[s], 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sh->cb_full[b], EW);
            mbar_init(&sh->cb_empty[b], 1);
        }
        sh->cta_count = 0;
        for (int i = 1; i < 8; ++i) sh->chunk_ring[i] = 0ull;
        sh->chunk_cnt[0] = sh->chunk_cnt[1] = 0u;
        sh->chunk_arr[0] = sh->chunk_arr[1] = 0u;
        sh->chunk_ring[0] = (1ull << 32) | blockIdx.x;   // the CTA's first chunk needs no ticket
        fence_mbar_init();
    }
    __syncthreads();
    // every evaluator warp's first tile: in flight while the program is copied into shared memory
    const bool active = warp < S;   // evaluator warp that owns a stage
    const unsigned long long l2pol = fp.l2_stream ? l2_evict_first_policy() : 0ull;
    uint8_t *stage = stages + static_cast<size_t>(warp < EW ? warp : 0) * p.stage_bytes;
    WarpTiles nxt;                  // the next tile to FETCH (one ahead of the tile being evaluated)
    nxt.chunk = blockIdx.x;
    nxt.k = 0;
    nxt.nt = nxt.chunk < fp.n_chunks ? chunk_tiles_of(nxt.chunk, CT, p.n_tiles) : 0;
    nxt.j = warp;                   // first_tile_of(warp, 0, CT, S) for warp < S
    if (active) {
        if (settle(nxt, warp, CT, S, fp.n_chunks, p.n_tiles, sh->chunk_ring)) {
            if (lane == 0) {
                produce_tile_stream(p, stage, nxt.chunk * CT + nxt.j, T, &sh->full[warp], l2pol);
                // whoever fetches a chunk's claim tile takes the CTA's next chunk
                if (nxt.j == claim_tile_of(nxt.nt, S)) claim_chunk(sh->chunk_ring, nxt.k + 1u, fp.fctl, gridDim.x);
            }
            nxt.j += S;
        }
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&fp.prog);  // kernel parameter space
        uint4 *dst = reinterpret_cast<uint4 *>(&sh->prog);
        for (uint32_t i = tid; i < sizeof(Program) / 16; i += kThreads) dst[i] = src[i];
    }
    __syncthreads();
    patch_leaves(&sh->prog, p, tid, kThreads);
    __syncthreads();
    const Program *sp = &sh->prog;

    if (warp < EW) {
        // ===== E: predicate evaluation, bitmap words into the chunk buffer =====
        LaneGeom<RPL> g;
        g.init(lane);
        uint32_t my_count = 0;
        uint32_t k = 0;   // k-th chunk of this CTA
        uint32_t it = 0;  // tiles this warp has consumed: the phase of ITS stage
        unsigned long long t_wait_empty = 0;  // diagnostics: time warp 0 waited for a free chunk buffer
        for (long long chunk = blockIdx.x; chunk < fp.n_chunks; chunk = chunk_of(sh->chunk_ring, ++k)) {
            const uint32_t buf = k & 1u;
            // the compaction warps must have read this buffer's previous chunk (k - 2)
            const unsigned long long tw0 = (fp.trace && tid == 0) ? global_ns() : 0ull;
            mbar_wait(&sh->cb_empty[buf], ((k >> 1) & 1u) ^ 1u);
            if (fp.trace && tid == 0) t_wait_empty += global_ns() - tw0;
            uint8_t *cb = reinterpret_cast<uint8_t *>(cbuf + buf * kFuseChunkWords);
            const long long t0 = chunk * CT;
            const uint32_t nt = chunk_tiles_of(chunk, CT, p.n_tiles);
            uint32_t chunk_count = 0;  // this lane's matches in this chunk
            if (active)
                for (uint32_t j = first_tile_of(warp, k, CT, S); j < nt; j += S, ++it) {
                    mbar_wait(&sh->full[warp], it & 1u);
                    if (fp.trace && it == 0 && tid == 0) fp.trace[blockIdx.x * 8 + 1] = global_ns();
                    const long long tile = t0 + j;
                    const uint32_t acc =
                        eval_tile<RPL>(sp, stage, lane, g, tile * T + static_cast<long long>(lane) * RPL, p.n_rows);
                    __syncwarp();  // every lane has read the stage: fetch this warp's next tile into it
                    if (settle(nxt, warp, CT, S, fp.n_chunks, p.n_tiles, sh->chunk_ring)) {
                        if (lane == 0) {
                            fence_proxy_async_smem();
                            produce_tile_stream(p, stage, nxt.chunk * CT + nxt.j, T, &sh->full[warp], l2pol);
                            if (nxt.j == claim_tile_of(nxt.nt, S)) claim_chunk(sh->chunk_ring, nxt.k + 1u, fp.fctl, gridDim.x);
                        }
                        nxt.j += S;
                    }
                    chunk_count += static_cast<uint32_t>(__popc(acc));
                    store_mask<RPL>(cb + j * (T / 8), lane, acc);
                }
            my_count += chunk_count;
            // The chunk's AGGREGATE is published here, by the last evaluator warp to finish the chunk, not by the
            // compaction warps: those take the CTA's chunks one after the other, so an aggregate they published had
            // to wait for the look-back of the chunk before it -- which was waiting for other CTAs' aggregates, held
            // up the same way.  Measured (tools/k1f_trace.py): 11.8 us of look-back per 17.5 us chunk, a convoy.
            const uint32_t warp_chunk = __reduce_add_sync(0xffffffffu, chunk_count);
            __syncwarp();  // every lane's bitmap words are written before lane 0 arrives
            if (lane == 0) {
                if (warp_chunk) atomicAdd(&sh->chunk_cnt[buf], warp_chunk);
                __threadfence_block();
                if (atomicAdd(&sh->chunk_arr[buf], 1u) == static_cast<uint32_t>(EW) - 1u) {
                    __threadfence_block();
                    const uint32_t total = atomicExch(&sh->chunk_cnt[buf], 0u);
                    sh->chunk_arr[buf] = 0u;
                    st_desc(fp.desc + chunk, make_desc(fp.epoch, chunk == 0 ? kStatePrefix : kStateAgg, total));
                }
                mbar_arrive(&sh->cb_full[buf]);  // release: this warp's words of the chunk are in cb
            }
        }
        const uint32_t warp_total = __reduce_add_sync(0xffffffffu, my_count);
        if (lane == 0 && warp_total) atomicAdd(&sh->cta_count, static_cast<unsigned long long>(warp_total));
        if (fp.trace && tid == 0) {
            fp.trace[blockIdx.x * 8 + 2] = global_ns();
            fp.trace[blockIdx.x * 8 + 5] = k;
            fp.trace[blockIdx.x * 8 + 4] = t_wait_empty;
        }
    } else {
        // ===== C: ordered compaction of finished chunks =====
        const uint32_t ct = tid - 32u * EW;
        const uint32_t cw = ct >> 5;
        const uint32_t chunk_rows = CT * T;
        unsigned long long t_lookback = 0, t_expand = 0, t_wait_full = 0;  // diagnostics (QPE_FUSE_TRACE)
        uint32_t k = 0;
        for (long long chunk = blockIdx.x; chunk < fp.n_chunks; chunk = chunk_of(sh->chunk_ring, ++k)) {
            const uint32_t buf = k & 1u;
            const uint32_t nt = chunk_tiles_of(chunk, CT, p.n_tiles);
            const uint32_t nw = nt * WPT;
            // (one polling warp + a named barrier for the others measured 1.5 % SLOWER than every warp polling)
            const unsigned long long tf0 = (fp.trace && ct == 0) ? global_ns() : 0ull;
            mbar_wait_relaxed(&sh->cb_full[buf], (k >> 1) & 1u, fp.poll_ns);
            if (fp.trace && ct == 0) t_wait_full += global_ns() - tf0;
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
                if (lane == 0) sh->round_base[kFuseRounds] = total;  // (the aggregate was published by the evaluators)
                const unsigned long long tl0 = (fp.trace && ct == 0) ? global_ns() : 0ull;
                const uint32_t excl = warp_lookback(fp.desc, chunk, fp.epoch, lane);
                if (fp.trace && ct == 0) t_lookback += global_ns() - tl0;
                if (lane == 0) {
                    st_desc(fp.desc + chunk, make_desc(fp.epoch, kStatePrefix, excl + total));
                    sh->excl = excl;
                }
            }
            named_bar_sync(1, kFuseCompactThreads);
            const uint32_t total = sh->round_base[kFuseRounds];
            const uint32_t excl = sh->excl;
            const unsigned long long tx0 = (fp.trace && ct == 0) ? global_ns() : 0ull;
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
                    // dense chunk: a round (128 / 256 words, <= 4096 / 8192 ids) at a time through the stage.  (Expanding
                    // word by word with lane-parallel direct stores -- one coalesced store per 32-row word, no barriers --
                    // was measured again in round 2, with the light evaluators: still slower, 2.46 vs 2.36 ms for QN at
                    // 50 % on 1 B rows, 0.309 vs 0.294 ms on 100 M.  So was the same lane-parallel expansion INTO THE
                    // STAGE (32 independent iterations per warp and round instead of a ~30-deep bit walk per thread, no
                    // bank conflicts): a 125 M-row shard with a 94 %-full first 10 M rows 281 vs 269 us with 4 compaction
                    // warps, 266 vs 265 us with 8; QN at 50 % 0.242 vs 0.248 ms.  The bit walk is not what a dense chunk
                    // waits for.)
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
            if (fp.trace && ct == 0) t_expand += global_ns() - tx0;
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
        if (fp.trace && ct == 0) {
            fp.trace[blockIdx.x * 8 + 6] = t_lookback;
            fp.trace[blockIdx.x * 8 + 7] = (t_expand << 32) | (t_wait_full & 0xffffffffull);
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (fp.trace) fp.trace[blockIdx.x * 8 + 3] = global_ns();
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
            fc->next_chunk = 0u;
            for (int i = 0; i < kMaxProgressSegments; ++i) fc->seg_stored[i] = 0u;
            if (fp.host_count) {
                *reinterpret_cast<volatile unsigned long long *>(fp.host_count) = total;
                __threadfence_system();
            }
        }
        // Launched as a programmatic dependent (sharded SELECT, two queries in flight) this grid must not COMPLETE
        // before the kernel it overlapped has: the stream's next kernel (this query's count exchange) tells the other
        // ranks that everything before it on this rank is done.  A no-op for an ordinary launch.
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
}

size_t fused_param_bytes() { return sizeof(FusedParams); }

// function attributes are per device: cache what was set per (instantiation, device); engines of different threads
// may launch at the same time, so the caches are updated under a lock
constexpr int kMaxDevices = 64;
static std::mutex g_attr_mutex;
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
        // the largest tile (rows per evaluator warp and pass) that leaves >= 8 stages = 8 evaluating warps, else
        // >= 4, >= 2, 1: a stage is held by its warp while it evaluates, the others are in flight
        static const int kTiles[] = {1024, 512, 256};
        int best_T = 0, best_S = 0;
        for (int want = 8; want >= 1 && !best_T; want >>= 1)
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
    if ((T != 256 && T != 512 && T != 1024) || S < 1 || S > kMaxStages) {
        if (why) *why = "invalid tile geometry (tile rows must be 256, 512 or 1024, at most 16 stages)";
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
        // Chunk size: 64 Ki rows unless a smaller one finishes sooner.  Per round of chunks a CTA needs the longer of
        // its scan time (rows x bytes per row at ~47 GB/s per SM) and the compaction warps' per-chunk latency
        // (~6 us: barrier, scan, look-back, expansion; measured with tools/k1f_trace.py -- a 25 M-row table cut into
        // 8 Ki-row chunks ran at 2.2 TB/s, every round waiting for the compaction).  Small tables take smaller chunks
        // only to put more SMs to work.  (With the chunks handed out by tickets the last round is shared out evenly, and
        // a model that charges (chunks per SM + 1) x chunk time prefers 32 Ki-row chunks for a 125 M-row shard and 16 Ki
        // for 40 B rows -- measured, that is SLOWER: 295 vs 258 us and 0.78 vs 0.58 ms; a chunk costs ~3 us that its
        // scan does not hide once it lasts less than ~15 us.  So the rule stays as it was.)
        int best_ct = kFuseMaxChunkRows / T;
        double best_us = 1e30;
        for (int ct = kFuseMaxChunkRows / T; ct >= 1 && static_cast<long long>(ct) * T >= 8192; ct >>= 1) {
            const int64_t nc = (geo->n_tiles + ct - 1) / ct;
            const int64_t g = nc < n_sm ? nc : n_sm;
            const int64_t rounds = (nc + g - 1) / (g > 0 ? g : 1);
            const double scan_us = static_cast<double>(ct) * T * static_cast<double>(bpr > 0 ? bpr : 1) / 47e3;
            const double us = static_cast<double>(rounds) * (scan_us > 6.0 ? scan_us : 6.0);
            if (us < best_us * 0.97) {  // a smaller chunk must win by 3 %
                best_us = us;
                best_ct = ct;
            }
        }
        // experiments: QPE_FUSE_CHUNK_ROWS=<rows> forces the chunk size (rounded down to whole tiles)
        static const char *force_rows = std::getenv("QPE_FUSE_CHUNK_ROWS");
        if (force_rows) {
            long long fr = std::atoll(force_rows) / T;
            if (fr >= 1 && fr * T <= kFuseMaxChunkRows) best_ct = static_cast<int>(fr);
        }
        geo->chunk_tiles = best_ct;
        geo->n_chunks = (geo->n_tiles + best_ct - 1) / best_ct;
        int64_t g = geo->n_chunks < n_sm ? geo->n_chunks : n_sm;
        if (g < 1) g = 1;
        geo->grid = static_cast<int>(g);
    }
    return true;
}

template <int RPL>
static cudaError_t launch_scan_r(const ScanParams &p, const ScanGeometry &geo, cudaStream_t stream) {
    // per instantiation and device: raise the dynamic shared memory limit only when it grows
    static size_t allowed_by_device[kMaxDevices] = {0};
    size_t &allowed = allowed_by_device[current_device_slot()];
    std::unique_lock<std::mutex> attr_lock(g_attr_mutex);
    if (geo.smem_bytes > allowed) {
        const cudaError_t e = cudaFuncSetAttribute(scan_tma_kernel<RPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(geo.smem_bytes));
        if (e != cudaSuccess) return e;
        allowed = geo.smem_bytes;
    }
    attr_lock.unlock();
    long long grid = p.n_tiles - p.tile_begin;  // tiles of this launch
    if (grid > geo.grid) grid = geo.grid;
    if (grid < 1) grid = 1;
    scan_tma_kernel<RPL><<<static_cast<unsigned>(grid), 32 * kEvalWarps, geo.smem_bytes, stream>>>(p);
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
    p.n_rows = t.n;
    p.tile_begin = L.tile_begin;
    p.n_tiles = L.tile_end > 0 ? L.tile_end : geo.n_tiles;
    p.ctl = const_cast<QueryCtl *>(L.d_ctl);
    p.out_bitmap = L.out_bitmap;
    return cudaSuccess;
}

cudaError_t scan_launch(const ScanLaunch &L, const ScanGeometry &geo, cudaStream_t stream) {
    ScanParams p{};
    const cudaError_t e = fill_scan_params(L, geo, p);
    if (e != cudaSuccess) return e;
    switch (geo.tile_rows) {
        case 256: return launch_scan_r<8>(p, geo, stream);
        case 512: return launch_scan_r<16>(p, geo, stream);
        case 1024: return launch_scan_r<32>(p, geo, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <int RPL>
static cudaError_t launch_batch_r(const BatchParams &bp, const ScanGeometry &geo, cudaStream_t stream) {
    static size_t allowed_by_device[kMaxDevices] = {0};
    size_t &allowed = allowed_by_device[current_device_slot()];
    std::unique_lock<std::mutex> attr_lock(g_attr_mutex);
    if (geo.smem_bytes > allowed) {
        const cudaError_t e = cudaFuncSetAttribute(scan_batch_kernel<RPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(geo.smem_bytes));
        if (e != cudaSuccess) return e;
        allowed = geo.smem_bytes;
    }
    attr_lock.unlock();
    scan_batch_kernel<RPL><<<geo.grid, 32 * kEvalWarps, geo.smem_bytes, stream>>>(bp);
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
    bp.n_prog = n_prog;
    bp.progs = d_progs;
    for (int q = 0; q < n_prog; ++q) bp.bitmap[q] = d_bitmaps[q];
    bp.counts = d_counts;
    if (geo.n_tiles == 0) return cudaSuccess;
    switch (geo.tile_rows) {
        case 256: return launch_batch_r<8>(bp, geo, stream);
        case 512: return launch_batch_r<16>(bp, geo, stream);
        case 1024: return launch_batch_r<32>(bp, geo, stream);
        default: return cudaErrorInvalidValue;
    }
}

template <int RPL, int CW>
static cudaError_t launch_fused_r(const FusedParams &fp, const ScanGeometry &geo, cudaStream_t stream, bool pdl) {
    static size_t allowed_by_device[kMaxDevices] = {0};
    size_t &allowed = allowed_by_device[current_device_slot()];
    std::unique_lock<std::mutex> attr_lock(g_attr_mutex);
    if (geo.smem_bytes > allowed) {
        const cudaError_t e = cudaFuncSetAttribute(scan_fused_kernel<RPL, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(geo.smem_bytes));
        if (e != cudaSuccess) return e;
        allowed = geo.smem_bytes;
    }
    attr_lock.unlock();
    if (pdl) {
        // programmatic dependent launch: this scan may start while the kernel before it in the stream (the previous
        // query's count exchange / delivery, which has nothing to hand to it) is still running
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(geo.grid);
        cfg.blockDim = dim3(32 * (kEvalWarps + CW));
        cfg.dynamicSmemBytes = geo.smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, scan_fused_kernel<RPL, CW>, fp);
    }
    scan_fused_kernel<RPL, CW><<<geo.grid, 32 * (kEvalWarps + CW), geo.smem_bytes, stream>>>(fp);
    return cudaGetLastError();
}

template <int RPL>
static cudaError_t launch_fused_cw(const FusedParams &fp, const ScanGeometry &geo, cudaStream_t stream, bool pdl) {
    return geo.compact_warps == 8 ? launch_fused_r<RPL, 8>(fp, geo, stream, pdl) : launch_fused_r<RPL, 4>(fp, geo, stream, pdl);
}

cudaError_t fused_launch(const FusedLaunch &L, const ScanGeometry &geo, cudaStream_t stream) {
    FusedParams fp{};
    const cudaError_t e = fill_scan_params(L.scan, geo, fp.s);
    if (e != cudaSuccess) return e;
    if (geo.chunk_tiles < 1 || static_cast<long long>(geo.chunk_tiles) * geo.tile_rows > kFuseMaxChunkRows ||
        (geo.compact_warps != 4 && geo.compact_warps != 8))
        return cudaErrorInvalidValue;
    fp.chunk_tiles = geo.chunk_tiles;
    fp.poll_ns = 256;  // 0 .. 1000 ns measured alike: any sleep that keeps the polls rare
    fp.n_chunks = geo.n_chunks;
    // QPE_SCAN_L2_HINT=0 / 1 forces the streaming L2 policy off / on (experiments); else the caller decides
    static const char *l2_env = std::getenv("QPE_SCAN_L2_HINT");
    fp.l2_stream = l2_env ? (std::atoi(l2_env) != 0 ? 1u : 0u) : (L.l2_stream ? 1u : 0u);
    fp.desc = L.desc;
    fp.epoch = L.epoch;
    fp.id_base = L.id_base;
    fp.out_ids = L.out_ids;
    fp.out_cap = L.out_cap;
    fp.seg_chunks = L.progress ? L.seg_chunks : 0;
    fp.progress = L.progress;
    fp.host_count = L.host_count;
    fp.fctl = L.d_fctl;
    fp.trace = L.trace;
    fp.prog = *L.scan.h_prog;
    if (!fp.fctl) return cudaErrorInvalidValue;
    if (fp.n_chunks == 0) return cudaSuccess;
    switch (geo.tile_rows) {
        case 256: return launch_fused_cw<8>(fp, geo, stream, L.pdl);
        case 512: return launch_fused_cw<16>(fp, geo, stream, L.pdl);
        case 1024: return launch_fused_cw<32>(fp, geo, stream, L.pdl);
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
    CandSegments segs;            // host-known segments (identity list) ...
    const CandSegments *d_segs;   // ... or the table K3s left on the device (index path: no host round trip)
    QueryCtl *ctl;
    unsigned long long *tile_desc;
    uint32_t epoch;
    uint32_t *out_ids;
    unsigned long long out_cap;   // nothing is stored at or beyond it (the count is still exact)
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

// Persistent CTAs: tiles of 1 Ki candidates are claimed IN ORDER from an atomic counter (every predecessor of a tile is
// finished or being worked on by a resident CTA, so the look-back cannot deadlock); the number of candidates is read
// from the segment table, which on the index path was written by the probe kernel just before -- the host never sees it.
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
    const CandSegments &S = p.d_segs ? *p.d_segs : p.segs;
    const int n_seg = S.n_seg;
    const long long n_cand = S.vstart[n_seg];
    const long long n_tiles = (n_cand + kFilterTile - 1) / kFilterTile;
    const Program *sp = &s_prog;
    constexpr int kWarps = kFilterThreads / 32;
    for (;;) {
        __syncthreads();  // the previous tile's shared words have been read; (first pass) the program is in place
        if (tid == 0) s_tile = static_cast<long long>(atomicAdd(&p.ctl->tile_counter, 1u));
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= n_tiles) return;

        uint32_t rows[kFilterItems];
#pragma unroll
        for (int it = 0; it < kFilterItems; ++it) {
            const long long i = tile * kFilterTile + it * kFilterThreads + tid;
            uint32_t row = 0;
            bool valid = i < n_cand;
            if (valid) {
                int sg = 0;
                while (sg + 1 < n_seg && i >= S.vstart[sg + 1]) ++sg;
                const long long k = S.first[sg] + (i - S.vstart[sg]);
                row = S.perm[sg] ? __ldg(S.perm[sg] + k) : static_cast<uint32_t>(k);
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
                if (tile == n_tiles - 1) p.ctl->out_count = static_cast<unsigned long long>(excl) + total;
                s_total = total;
                s_excl = excl;
            }
        }
        __syncthreads();
        if (p.out_ids == nullptr || s_total == 0) continue;
        const uint32_t excl = s_excl;
#pragma unroll
        for (int it = 0; it < kFilterItems; ++it) {
            const uint32_t w = s_words[it * kWarps + warp];
            if ((w >> lane) & 1u) {
                const unsigned long long o = static_cast<unsigned long long>(excl) + s_woff[it * kWarps + warp] + __popc(w & lanemask_lt());
                if (o < p.out_cap) p.out_ids[o] = rows[it];
            }
        }
    }
}

cudaError_t filter_launch(const DevTable &t, const QueryCtl *d_ctl, const CandSegments &segs,
                          unsigned long long *tile_desc, uint32_t epoch, uint32_t *out_ids, cudaStream_t stream,
                          const CandSegments *d_segs, long long max_candidates, unsigned long long out_cap) {
    FilterParams p{};
    for (int c = 0; c < NUM_COLS; ++c) {
        p.col[c] = t.col[c].d;
        p.width[c] = t.col[c].width;
    }
    p.segs = segs;
    p.d_segs = d_segs;
    p.ctl = const_cast<QueryCtl *>(d_ctl);
    p.tile_desc = tile_desc;
    p.epoch = epoch;
    p.out_ids = out_ids;
    p.out_cap = out_cap;
    // with the segment table on the device the host only knows an upper bound of the candidates
    const long long bound = d_segs ? max_candidates : segs.vstart[segs.n_seg];
    long long grid = filter_tiles(bound);
    if (grid == 0) return cudaSuccess;
    static int wave = 0;
    if (wave == 0) {
        int dev = 0, n_sm = 148, per_sm = 4;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, filter_kernel, kFilterThreads, 0);
        wave = n_sm * (per_sm > 0 ? per_sm : 1);
    }
    if (grid > wave) grid = wave;
    filter_kernel<<<static_cast<unsigned>(grid), kFilterThreads, 0, stream>>>(p);
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
