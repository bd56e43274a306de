// qpe_internal.h -- private types of libqpegpu (columnar table, predicate program, engine object)
#pragma once

#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <string>
#include <vector>

#include "qpe_abi.h"

namespace qpe {

// ---------------------------------------------------------------------------------------------
// Schema.  Column order == struct order of `record` (logType.h:11-24) == the reference's
// SELECT * order (executeEngine-serial.c:485-488) and its field table (recordSchema.c:12-25).
// ---------------------------------------------------------------------------------------------
enum ColId : int {
    C_COMMAND_ID = 0,
    C_RAW_COMMAND,
    C_BASE_COMMAND,
    C_SHELL_TYPE,
    C_EXIT_CODE,
    C_TIMESTAMP,
    C_SUDO_USED,
    C_WORKING_DIRECTORY,
    C_USER_ID,
    C_USER_NAME,
    C_HOST_NAME,
    C_RISK_LEVEL,
    NUM_COLS
};

enum ColType : int { T_U64 = 0, T_I32 = 1, T_STR = 2, T_BOOL = 3 };  // == FieldType values

struct ColInfo {
    const char *name;
    ColType type;
    uint32_t field_bytes;  // sizeof the member of `record`
    uint32_t rec_offset;   // offsetof in `record`
};
extern const ColInfo kCols[NUM_COLS];
int col_by_name(const char *name);  // -1 if unknown

static inline uint32_t round_up16(uint32_t x) { return (x + 15u) & ~15u; }
// widest device layout a string column can ever need: its struct field, rounded to 16 B
static inline uint32_t max_str_width(int c) { return round_up16(kCols[c].field_bytes); }

// Rows are processed in tiles; every column allocation is padded to a multiple of this so a
// full tile can always be read (bulk copies need 16-byte multiples, and the u8 column has
// 1-byte rows).
constexpr int64_t kRowPad = 8192;

// ---------------------------------------------------------------------------------------------
// Host-side columnar staging (what ingest produces and what DELETE/INSERT side effects read).
// Strings are fixed width, NUL padded; width[c] is a multiple of 16.
// ---------------------------------------------------------------------------------------------
struct HostColumns {
    int64_t n = 0;
    uint32_t width[NUM_COLS] = {0};
    std::vector<uint8_t> data[NUM_COLS];
    void init_widths_minimal();
    void reserve_rows(int64_t rows);
    void append_record(const record &r);   // widens string columns on demand
    void widen(int c, uint32_t new_width);
};

// ---------------------------------------------------------------------------------------------
// Device table
// ---------------------------------------------------------------------------------------------
struct DevColumn {
    uint8_t *d = nullptr;
    uint32_t width = 0;  // bytes per row (8 / 4 / 1 / 16k)
    int64_t cap = 0;     // rows allocated (multiple of kRowPad)
};

struct DevTable {
    int64_t n = 0;
    uint64_t row_base = 0;  // global id of local row 0 (sharded tables)
    DevColumn col[NUM_COLS];
    bool resident(int c) const { return col[c].d != nullptr; }
};

// ---------------------------------------------------------------------------------------------
// Predicate program (device-visible, POD).  Produced by where_compile.cpp from a whereClauseS
// list; evaluation order restates evaluateWhereClause (executeEngine-serial.c:292-316).
// ---------------------------------------------------------------------------------------------
enum POp : uint8_t {
    P_LEAF_SET = 0,  // acc  = leaf
    P_LEAF_AND = 1,  // acc  = leaf & acc
    P_LEAF_OR = 2,   // acc  = leaf | acc
    P_PUSH = 3,      // stack[arg] = acc
    P_POP_AND = 4,   // acc  = acc & stack[arg]
    P_POP_OR = 5,    // acc  = acc | stack[arg]
    P_CONST = 6,     // acc  = arg ? all : none
    P_NOT = 7        // acc  = ~acc        (extension: used for DELETE's keep-mask only)
};

constexpr int kMaxInstr = 96;
constexpr int kMaxLeaves = 48;
constexpr int kMaxStack = 8;
constexpr int kLitPoolBytes = 2048;

struct PLeaf {
    uint8_t col;    // ColId
    uint8_t type;   // ColType
    uint8_t tt;     // 3-bit truth table over {lt, eq, gt}: bit0 = result when field<lit, bit1 ==, bit2 >
    uint8_t nch;    // text column: cell width / 16.  0 on the host; patched into the kernel's shared-memory copy
    uint32_t lit_off;  // string literal: byte offset into lit_pool (16-byte aligned), length = column width
    uint64_t lit_u64;  // T_U64 literal
    int32_t lit_i32;   // T_I32 literal; T_BOOL literal in bit 0
    uint32_t smem_off; // byte offset of the column inside a staged tile.  0 on the host; patched like nch, so that a
                       // leaf is evaluated from ONE shared-memory record (no dependent parameter look-ups)
};

struct PInstr {
    uint8_t op;
    uint8_t arg;  // leaf index / stack slot / const
};

struct Program {
    int32_t n_instr;
    int32_t n_leaves;
    uint32_t col_mask;  // columns referenced
    uint32_t pad;
    PInstr instr[kMaxInstr];
    PLeaf leaf[kMaxLeaves];
    alignas(16) uint8_t lit_pool[kLitPoolBytes];
};

// Index-path plan: one entry per (top-level condition, index) pair, in the reference's
// generation order (executeEngine-serial.c:358-459).
struct SegmentPlan {
    int index_slot;
    bool is_u64;
    uint64_t lo_u64, hi_u64;
    int32_t lo_i32, hi_i32;
};

std::string compile_where(const struct whereClauseS *wc, const uint32_t width[NUM_COLS], Program *out,
                          bool invert);

// ---------------------------------------------------------------------------------------------
// Flattened index (one per indexed u64/int attribute): the row permutation sorted by
// (key ASC, position DESC) -- the leaf-chain order of the reference's B+ tree
// -- plus the keys in that order and implicit separator levels above them.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxIndexLevels = 10;
struct DevIndex {
    int col = -1;
    ColType type = T_I32;
    bool usable = false;  // bool/string indexes are registered but never probed (serial :425-429)
    bool dirty = true;    // needs (re)build before the next probe
    int64_t n = 0;
    void *keys = nullptr;      // n sorted keys (u64 or i32)
    uint32_t *perm = nullptr;  // n row ids
    int64_t cap = 0;
    // separator levels: level[0] is the coarsest. level l has cnt[l] keys, each the first key of a
    // group of `fanout` entries of the level below (the finest level groups entries of `keys`).
    int n_levels = 0;
    int fanout = 16;
    void *level[kMaxIndexLevels] = {nullptr};      // views into level_buf, coarsest first
    int64_t level_cnt[kMaxIndexLevels] = {0};
    // the level buffers themselves, by HEIGHT above the key array (0 = finest), sized once for `cap` entries so that
    // INSERT / DELETE maintenance never allocates
    void *level_buf[kMaxIndexLevels] = {nullptr};
    int64_t level_cap[kMaxIndexLevels] = {0};
};

struct ScanStats {
    double kernel_ms = 0;     // device time of the match phase (CUDA events)
    double total_ms = 0;      // wall time of the call
    int64_t rows_scanned = 0;
    int64_t candidates = 0;   // index path: sum of segment lengths
    int64_t matches = 0;
    int64_t algo_bytes = 0;   // N * sum(width of referenced cols) + 4*M
    int32_t path = 0;         // 0 = full scan, 1 = index
    int32_t launches = 0;     // kernels launched by this call
    int32_t tile_rows = 0;
    int32_t stages = 0;
    int32_t grid = 0;
    int32_t pad = 0;
    double scan_ms = 0;       // K1 alone
    double compact_ms = 0;    // K1c alone
};

}  // namespace qpe
