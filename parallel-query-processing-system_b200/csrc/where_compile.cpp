// where_compile.cpp -- whereClauseS list -> device predicate program
//
// Restates, once per query on the host, what the reference does per row per condition:
//   checkCondition          engine/serial/executeEngine-serial.c:251-289  (literal conversion by
//                           attribute NAME with strtoull / atoi / strcasecmp -- the same libc calls
//                           are made here so the conversion is exact by construction)
//   create_where_condition  :129-213  (attribute x operator -> comparator, NULL => row fails)
//   evaluateWhereClause     :292-316  (right-recursive list, no precedence, AND is the default join)
//
// Program form.  eval(list n0 op0 n1 op1 ... nk) = val(n0) op0 (val(n1) op1 (... val(nk))).
// The fold is emitted from the LAST node to the first so a single accumulator suffices:
//     acc = val(nk); acc = val(n_{k-1}) op acc; ...
// A parenthesised group in non-last position needs the accumulator saved around its own fold
// (PUSH slot / POP_AND|POP_OR slot); the slot number is the nesting depth, fixed at compile time,
// so the device keeps the stack in registers.

#include <cstdlib>
#include <cstring>
#include <strings.h>

#include "qpe_internal.h"

namespace qpe {

const ColInfo kCols[NUM_COLS] = {
    {"command_id", T_U64, 8, offsetof(record, command_id)},
    {"raw_command", T_STR, 512, offsetof(record, raw_command)},
    {"base_command", T_STR, 100, offsetof(record, base_command)},
    {"shell_type", T_STR, 20, offsetof(record, shell_type)},
    {"exit_code", T_I32, 4, offsetof(record, exit_code)},
    {"timestamp", T_STR, 30, offsetof(record, timestamp)},
    {"sudo_used", T_BOOL, 1, offsetof(record, sudo_used)},
    {"working_directory", T_STR, 200, offsetof(record, working_directory)},
    {"user_id", T_I32, 4, offsetof(record, user_id)},
    {"user_name", T_STR, 50, offsetof(record, user_name)},
    {"host_name", T_STR, 100, offsetof(record, host_name)},
    {"risk_level", T_I32, 4, offsetof(record, risk_level)},
};

int col_by_name(const char *name) {
    if (!name) return -1;
    for (int c = 0; c < NUM_COLS; ++c)
        if (std::strcmp(kCols[c].name, name) == 0) return c;
    return -1;
}

namespace {

// operator text -> truth table over {lt, eq, gt}; 0 when the reference finds no comparator
uint8_t op_truth_table(const char *op, ColType type) {
    if (!op) return 0;
    if (std::strcmp(op, "=") == 0) return 0b010;
    if (std::strcmp(op, "!=") == 0) return 0b101;
    if (type == T_BOOL) return 0;  // sudo_used has only eq / neq comparators (:207-210)
    if (std::strcmp(op, ">") == 0) return 0b100;
    if (std::strcmp(op, "<") == 0) return 0b001;
    if (std::strcmp(op, ">=") == 0) return 0b110;
    if (std::strcmp(op, "<=") == 0) return 0b011;
    return 0;
}

struct Compiler {
    Program *p;
    const uint32_t *width;
    uint32_t lit_used = 0;
    std::string err;

    bool emit(uint8_t op, uint8_t arg) {
        if (p->n_instr >= kMaxInstr) {
            err = "WHERE clause too large (instruction limit)";
            return false;
        }
        p->instr[p->n_instr].op = op;
        p->instr[p->n_instr].arg = arg;
        p->n_instr++;
        return true;
    }

    // returns leaf index, -1 => constant false leaf, -2 => error
    int make_leaf(const struct whereClauseS *n) {
        const int c = col_by_name(n->attribute);
        if (c < 0 || !n->value) return -1;  // unknown attribute: no comparator => false (:212)
        const ColType type = kCols[c].type;
        const uint8_t tt = op_truth_table(n->op_, type);
        if (tt == 0) return -1;
        if (p->n_leaves >= kMaxLeaves) {
            err = "WHERE clause too large (leaf limit)";
            return -2;
        }
        PLeaf lf{};
        lf.col = static_cast<uint8_t>(c);
        lf.type = static_cast<uint8_t>(type);
        lf.tt = tt;
        switch (type) {
            case T_U64:
                lf.lit_u64 = std::strtoull(n->value, nullptr, 10);  // :258
                break;
            case T_I32:
                lf.lit_i32 = std::atoi(n->value);  // :265
                break;
            case T_BOOL:
                lf.lit_i32 = (strcasecmp(n->value, "true") == 0 || std::strcmp(n->value, "1") == 0) ? 1 : 0;  // :270
                break;
            case T_STR: {
                // strcmp(field, value) on a NUL-padded fixed-width column == unsigned memcmp over
                // the column width against the literal NUL-padded (or, if longer, truncated: the
                // field's terminator then sorts it below the literal, exactly as strcmp would).
                const uint32_t w = width[c];
                if (w == 0 || (w & 15u)) {
                    err = "string column has no device layout";
                    return -2;
                }
                if (lit_used + w > kLitPoolBytes) {
                    err = "WHERE clause too large (literal pool)";
                    return -2;
                }
                lf.lit_off = lit_used;
                const size_t len = std::strlen(n->value);
                std::memset(p->lit_pool + lit_used, 0, w);
                std::memcpy(p->lit_pool + lit_used, n->value, len < w ? len : w);
                lit_used += w;
                break;
            }
        }
        p->col_mask |= 1u << c;
        p->leaf[p->n_leaves] = lf;
        return p->n_leaves++;
    }

    bool emit_list(const struct whereClauseS *head, int depth) {
        // collect nodes
        const struct whereClauseS *nodes[64];
        int n = 0;
        for (const struct whereClauseS *w = head; w; w = w->next) {
            if (n >= 64) {
                err = "WHERE list too long";
                return false;
            }
            nodes[n++] = w;
        }
        if (n == 0) return emit(P_CONST, 1);  // evaluateWhereClause(NULL) == true (:293)
        for (int j = n - 1; j >= 0; --j) {
            const struct whereClauseS *nd = nodes[j];
            const bool last = (j == n - 1);
            // join between node j and j+1 is node j's logical_op; "OR" => or, anything else => and (:307-315)
            const bool is_or = !last && nd->logical_op && std::strcmp(nd->logical_op, "OR") == 0;
            // a group holding exactly one plain condition evaluates like that condition
            // ("(a < 5) AND ..." == "a < 5 AND ..."): no accumulator save / restore needed
            const struct whereClauseS *single = nullptr;
            if (nd->sub != nullptr) {
                const struct whereClauseS *x = nd->sub;
                while (x->sub != nullptr && x->next == nullptr) x = x->sub;  // ((a < 5)) unwraps too
                if (x->sub == nullptr && x->next == nullptr) single = x;
            }
            if (nd->sub != nullptr && single == nullptr) {
                if (last) {
                    if (!emit_list(nd->sub, depth)) return false;
                } else {
                    if (depth >= kMaxStack) {
                        err = "WHERE clause nested too deeply";
                        return false;
                    }
                    if (!emit(P_PUSH, static_cast<uint8_t>(depth))) return false;
                    if (!emit_list(nd->sub, depth + 1)) return false;
                    if (!emit(is_or ? P_POP_OR : P_POP_AND, static_cast<uint8_t>(depth))) return false;
                }
            } else {
                const struct whereClauseS *cond = single ? single : nd;
                const int lf = cond->attribute ? make_leaf(cond) : -1;
                if (lf == -2) return false;
                if (lf == -1) {
                    // constant-false condition
                    if (last) {
                        if (!emit(P_CONST, 0)) return false;
                    } else if (!is_or) {
                        if (!emit(P_CONST, 0)) return false;  // false AND acc == false
                    }                                        // false OR acc == acc: nothing to emit
                } else {
                    if (!emit(last ? P_LEAF_SET : (is_or ? P_LEAF_OR : P_LEAF_AND), static_cast<uint8_t>(lf)))
                        return false;
                }
            }
        }
        return true;
    }
};

}  // namespace

std::string compile_where(const struct whereClauseS *wc, const uint32_t width[NUM_COLS], Program *out, bool invert) {
    std::memset(out, 0, sizeof(Program));
    Compiler c{out, width};
    if (wc != nullptr) {
        if (!c.emit_list(wc, 0)) return c.err.empty() ? std::string("WHERE compile failed") : c.err;
    }
    if (invert) {
        if (wc == nullptr && !c.emit(P_CONST, 1)) return c.err;
        if (!c.emit(P_NOT, 0)) return c.err;
    }
    return std::string();
}

}  // namespace qpe
