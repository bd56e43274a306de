/*
 * qpe_oracle.c -- TEST INFRASTRUCTURE ONLY (see qpe_oracle.h).  Plain-C CPU restatement of the
 * reference's SELECT/WHERE path over a columnar table.  Nothing here is shipped or measured as
 * the product; libqpegpu.so never references this file.
 *
 * Parity status: PINNED -- against the reference's own known-answer tests, against the compiled
 * unmodified reference (oracle/_ref) on the probe set, and against tests/golden/*.json produced
 * by that compiled reference (tests/test_oracle.py).
 *
 * Each function cites the reference code it restates (paths relative to the reference root).
 */
#define _POSIX_C_SOURCE 200809L
#include "qpe_oracle.h"

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

/* schema: include/logType.h:11-24, engine/recordSchema.c:12-25 */
enum { O_U64 = 0, O_I32 = 1, O_STR = 2, O_BOOL = 3 };
static const struct {
    const char *name;
    int type;
    unsigned field_bytes;
    size_t offset;
} kSchema[ORACLE_NUM_COLS] = {
    {"command_id", O_U64, 8, offsetof(record, command_id)},
    {"raw_command", O_STR, 512, offsetof(record, raw_command)},
    {"base_command", O_STR, 100, offsetof(record, base_command)},
    {"shell_type", O_STR, 20, offsetof(record, shell_type)},
    {"exit_code", O_I32, 4, offsetof(record, exit_code)},
    {"timestamp", O_STR, 30, offsetof(record, timestamp)},
    {"sudo_used", O_BOOL, 1, offsetof(record, sudo_used)},
    {"working_directory", O_STR, 200, offsetof(record, working_directory)},
    {"user_id", O_I32, 4, offsetof(record, user_id)},
    {"user_name", O_STR, 50, offsetof(record, user_name)},
    {"host_name", O_STR, 100, offsetof(record, host_name)},
    {"risk_level", O_I32, 4, offsetof(record, risk_level)},
};

int oracle_col_by_name(const char *name) {
    if (!name) return -1;
    for (int c = 0; c < ORACLE_NUM_COLS; c++)
        if (strcmp(kSchema[c].name, name) == 0) return c;
    return -1;
}

static const unsigned char *cell_ptr(const oracle_table *t, int c, long long row) {
    return (const unsigned char *)t->col[c] + (size_t)row * t->width[c];
}

/* operator text -> 0..5, -1 when create_where_condition finds none
 * (engine/serial/executeEngine-serial.c:129-213: "=", "!=", ">", "<", ">=", "<=") */
static int op_code(const char *op) {
    if (!op) return -1;
    if (strcmp(op, "=") == 0) return 0;
    if (strcmp(op, "!=") == 0) return 1;
    if (strcmp(op, ">") == 0) return 2;
    if (strcmp(op, "<") == 0) return 3;
    if (strcmp(op, ">=") == 0) return 4;
    if (strcmp(op, "<=") == 0) return 5;
    return -1;
}

static int apply_op(int op, int cmp /* <0, 0, >0 : field vs literal */) {
    switch (op) {
        case 0: return cmp == 0;
        case 1: return cmp != 0;
        case 2: return cmp > 0;
        case 3: return cmp < 0;
        case 4: return cmp >= 0;
        case 5: return cmp <= 0;
        default: return 0;
    }
}

/* checkCondition (executeEngine-serial.c:251-289) + the CMP_NUM / CMP_STR comparators (:18-26):
 * the literal is converted by ATTRIBUTE NAME; no comparator => false. */
static int check_condition(const oracle_table *t, long long row, const struct whereClauseS *w) {
    const char *attr = w->attribute;
    if (!attr || !w->value) return 0;
    const int op = op_code(w->operator);
    const int c = oracle_col_by_name(attr);
    if (c < 0 || op < 0) return 0; /* unknown attribute: string branch, comparator lookup fails (:212) */
    const unsigned char *p = cell_ptr(t, c, row);
    switch (kSchema[c].type) {
        case O_U64: {
            unsigned long long lit = strtoull(w->value, NULL, 10); /* :258 */
            unsigned long long v;
            memcpy(&v, p, 8);
            return apply_op(op, v < lit ? -1 : (v > lit ? 1 : 0));
        }
        case O_I32: {
            int lit = atoi(w->value); /* :265 */
            int v;
            memcpy(&v, p, 4);
            return apply_op(op, v < lit ? -1 : (v > lit ? 1 : 0));
        }
        case O_BOOL: {
            int lit = (strcasecmp(w->value, "true") == 0 || strcmp(w->value, "1") == 0); /* :270 */
            int v = p[0] != 0;
            if (op > 1) return 0; /* only eq / neq exist for sudo_used (:207-210) */
            return op == 0 ? v == lit : v != lit;
        }
        default: /* strcmp(field, value) op 0 (:24-26) */
            return apply_op(op, strcmp((const char *)p, w->value));
    }
}

/* evaluateWhereClause (executeEngine-serial.c:292-316): right-recursive, no precedence;
 * "OR" => or, "AND" / NULL / anything else => and. */
int oracle_eval_row(const oracle_table *t, long long row, const struct whereClauseS *wc) {
    if (wc == NULL) return 1;
    int cur = wc->sub ? oracle_eval_row(t, row, wc->sub) : check_condition(t, row, wc);
    if (wc->next == NULL) return cur;
    if (wc->logical_op && strcmp(wc->logical_op, "OR") == 0) return cur || oracle_eval_row(t, row, wc->next);
    return cur && oracle_eval_row(t, row, wc->next);
}

/* linearSearchRecords (executeEngine-serial.c:854-878) */
long long oracle_scan_range(const oracle_table *t, long long first, long long n, const struct whereClauseS *wc,
                            uint32_t *ids) {
    long long m = 0;
    for (long long i = first; i < first + n; i++) {
        if (wc == NULL || oracle_eval_row(t, i, wc)) {
            if (ids) ids[m] = (uint32_t)i;
            m++;
        }
    }
    return m;
}

long long oracle_scan(const oracle_table *t, const struct whereClauseS *wc, uint32_t *ids) {
    return oracle_scan_range(t, 0, t->n, wc, ids);
}

/* ---- index order ------------------------------------------------------------------------- */
/* compare_key (engine/recordSchema.c:88-127) for the two probe-able key types */
static int key_cmp(const oracle_table *t, int c, uint32_t a, uint32_t b) {
    if (kSchema[c].type == O_U64) {
        unsigned long long x, y;
        memcpy(&x, cell_ptr(t, c, a), 8);
        memcpy(&y, cell_ptr(t, c, b), 8);
        return x < y ? -1 : (x > y ? 1 : 0);
    }
    int x, y;
    memcpy(&x, cell_ptr(t, c, a), 4);
    memcpy(&y, cell_ptr(t, c, b), 4);
    return x < y ? -1 : (x > y ? 1 : 0);
}

static const oracle_table *g_sort_table;
static int g_sort_col;
static int order_cmp(const void *pa, const void *pb) {
    const uint32_t a = *(const uint32_t *)pa, b = *(const uint32_t *)pb;
    const int k = key_cmp(g_sort_table, g_sort_col, a, b);
    if (k) return k;
    return a > b ? -1 : (a < b ? 1 : 0); /* equal keys: later table position first */
}

int oracle_index_order(const oracle_table *t, int col, uint32_t *perm) {
    if (col < 0 || col >= ORACLE_NUM_COLS) return -1;
    if (kSchema[col].type != O_U64 && kSchema[col].type != O_I32) return -1;
    for (long long i = 0; i < t->n; i++) perm[i] = (uint32_t)i;
    g_sort_table = t;
    g_sort_col = col;
    qsort(perm, (size_t)t->n, sizeof(uint32_t), order_cmp);
    return 0;
}

/* Literal replay of the insertion rule: findLeaf descends to the leftmost leaf that may hold the
 * key (bplus.c:340-342) and insertIntoLeaf / insertIntoLeafAfterSplitting put the entry at the
 * first slot whose key is >= the new key (:475-477, :512-517); rows are inserted in table order
 * (buildEngine-serial.c:46-53).  On the flattened leaf chain that is "insert before the first
 * entry with key >= new key". */
int oracle_index_order_by_insertion(const oracle_table *t, int col, uint32_t *perm) {
    if (col < 0 || col >= ORACLE_NUM_COLS) return -1;
    if (kSchema[col].type != O_U64 && kSchema[col].type != O_I32) return -1;
    long long len = 0;
    for (long long i = 0; i < t->n; i++) {
        long long lo = 0, hi = len; /* first slot with key >= key(i) */
        while (lo < hi) {
            long long mid = (lo + hi) / 2;
            if (key_cmp(t, col, perm[mid], (uint32_t)i) < 0) lo = mid + 1; else hi = mid;
        }
        memmove(perm + lo + 1, perm + lo, (size_t)(len - lo) * sizeof(uint32_t));
        perm[lo] = (uint32_t)i;
        len++;
    }
    return 0;
}

static int key_vs_lit(const oracle_table *t, int c, uint32_t row, KEY_T lit) {
    if (kSchema[c].type == O_U64) {
        unsigned long long x;
        memcpy(&x, cell_ptr(t, c, row), 8);
        return x < lit.v.u64 ? -1 : (x > lit.v.u64 ? 1 : 0);
    }
    int x;
    memcpy(&x, cell_ptr(t, c, row), 4);
    return x < lit.v.i32 ? -1 : (x > lit.v.i32 ? 1 : 0);
}

/* findRange (engine/bplus.c:282-314): start at the first entry >= key_start, emit while <= key_end */
int oracle_find_range(const oracle_table *t, int col, const uint32_t *perm, KEY_T lo, KEY_T hi, long long *first,
                      long long *count) {
    if (col < 0 || col >= ORACLE_NUM_COLS) return -1;
    long long a = 0, b = t->n;
    while (a < b) { /* lower bound of lo */
        long long mid = (a + b) / 2;
        if (key_vs_lit(t, col, perm[mid], lo) < 0) a = mid + 1; else b = mid;
    }
    const long long start = a;
    long long end = start;
    while (end < t->n && key_vs_lit(t, col, perm[end], hi) <= 0) end++;
    *first = start;
    *count = end - start;
    return 0;
}

/* ---- SELECT match phase ------------------------------------------------------------------- */
/* executeQuerySelectSerial (executeEngine-serial.c:328-476) */
long long oracle_select(const oracle_table *t, int num_indexes, const char *const *index_attrs,
                        const int *index_types, const struct whereClauseS *where, uint32_t **ids_out,
                        int *used_index) {
    uint32_t *cand = NULL;
    long long n_cand = 0, cap = 0;
    int any = 0;
    uint32_t **perms = calloc((size_t)(num_indexes > 0 ? num_indexes : 1), sizeof(uint32_t *));
    for (const struct whereClauseS *w = where; w; w = w->next) {
        if (w->attribute == NULL) continue; /* nested groups are skipped (:361-364) */
        for (int i = 0; i < num_indexes; i++) {
            if (strcmp(w->attribute, index_attrs[i]) != 0) continue;
            const int type = index_types[i];
            if (type != 0 && type != 1) continue; /* bool / string indexes unsupported (:425-429) */
            const int c = oracle_col_by_name(index_attrs[i]);
            if (c < 0 || kSchema[c].type != type) continue;
            const char *op = w->operator ? w->operator : "";
            KEY_T lo, hi;
            memset(&lo, 0, sizeof lo);
            memset(&hi, 0, sizeof hi);
            if (type == 0) { /* :377-400 */
                unsigned long long v = strtoull(w->value ? w->value : "", NULL, 10);
                lo.type = hi.type = KEY_UINT64;
                if (!strcmp(op, "=")) { lo.v.u64 = v; hi.v.u64 = v; }
                else if (!strcmp(op, ">")) { lo.v.u64 = v + 1; hi.v.u64 = UINT64_MAX; }
                else if (!strcmp(op, ">=")) { lo.v.u64 = v; hi.v.u64 = UINT64_MAX; }
                else if (!strcmp(op, "<")) { lo.v.u64 = 0; hi.v.u64 = v - 1; }
                else if (!strcmp(op, "<=")) { lo.v.u64 = 0; hi.v.u64 = v; }
                else { lo.v.u64 = 0; hi.v.u64 = UINT64_MAX; }
            } else { /* :401-424 */
                int v = atoi(w->value ? w->value : "");
                lo.type = hi.type = KEY_INT;
                if (!strcmp(op, "=")) { lo.v.i32 = v; hi.v.i32 = v; }
                else if (!strcmp(op, ">")) { lo.v.i32 = (int)((unsigned)v + 1u); hi.v.i32 = INT_MAX; }
                else if (!strcmp(op, ">=")) { lo.v.i32 = v; hi.v.i32 = INT_MAX; }
                else if (!strcmp(op, "<")) { lo.v.i32 = INT_MIN; hi.v.i32 = (int)((unsigned)v - 1u); }
                else if (!strcmp(op, "<=")) { lo.v.i32 = INT_MIN; hi.v.i32 = v; }
                else { lo.v.i32 = INT_MIN; hi.v.i32 = INT_MAX; }
            }
            any = 1;
            if (!perms[i]) {
                perms[i] = malloc(sizeof(uint32_t) * (size_t)(t->n > 0 ? t->n : 1));
                oracle_index_order(t, c, perms[i]);
            }
            long long first = 0, count = 0;
            oracle_find_range(t, c, perms[i], lo, hi, &first, &count);
            if (n_cand + count > cap) {
                cap = (n_cand + count) * 2 + 16;
                cand = realloc(cand, sizeof(uint32_t) * (size_t)cap);
            }
            memcpy(cand + n_cand, perms[i] + first, sizeof(uint32_t) * (size_t)count);
            n_cand += count; /* segments are concatenated, duplicates kept (:446-448) */
        }
    }
    for (int i = 0; i < num_indexes; i++) free(perms[i]);
    free(perms);
    uint32_t *out;
    long long m = 0;
    if (!any) { /* :464-467 */
        out = malloc(sizeof(uint32_t) * (size_t)(t->n > 0 ? t->n : 1));
        m = oracle_scan(t, where, out);
    } else { /* :469-474: the whole WHERE over the candidates, in candidate order */
        out = malloc(sizeof(uint32_t) * (size_t)(n_cand > 0 ? n_cand : 1));
        for (long long k = 0; k < n_cand; k++)
            if (oracle_eval_row(t, cand[k], where)) out[m++] = cand[k];
    }
    free(cand);
    if (used_index) *used_index = any;
    if (ids_out) *ids_out = out; else free(out);
    return m;
}

/* get_attribute_string_value (executeEngine-serial.c:216-248) */
void oracle_cell_text(const oracle_table *t, long long row, const char *attribute, char *buf, size_t cap) {
    const int c = oracle_col_by_name(attribute);
    if (c < 0) {
        snprintf(buf, cap, "NULL");
        return;
    }
    const unsigned char *p = cell_ptr(t, c, row);
    switch (kSchema[c].type) {
        case O_U64: {
            unsigned long long v;
            memcpy(&v, p, 8);
            snprintf(buf, cap, "%llu", v);
            break;
        }
        case O_I32: {
            int v;
            memcpy(&v, p, 4);
            snprintf(buf, cap, "%d", v);
            break;
        }
        case O_BOOL: snprintf(buf, cap, "%s", p[0] ? "true" : "false"); break;
        default: snprintf(buf, cap, "%s", (const char *)p); break;
    }
}

/* ---- table construction -------------------------------------------------------------------- */
oracle_table *oracle_table_from_records(const record *rows, long long n, const unsigned *widths) {
    oracle_table *t = calloc(1, sizeof *t);
    t->n = n;
    for (int c = 0; c < ORACLE_NUM_COLS; c++) {
        unsigned w;
        switch (kSchema[c].type) {
            case O_U64: w = 8; break;
            case O_I32: w = 4; break;
            case O_BOOL: w = 1; break;
            default: w = widths ? widths[c] : ((kSchema[c].field_bytes + 15u) & ~15u); break;
        }
        t->width[c] = w;
        unsigned char *d = calloc((size_t)(n > 0 ? n : 1), w);
        for (long long i = 0; i < n; i++) {
            const unsigned char *src = (const unsigned char *)&rows[i] + kSchema[c].offset;
            if (kSchema[c].type == O_STR) {
                size_t len = strnlen((const char *)src, kSchema[c].field_bytes - 1);
                if (len > w - 1) len = w - 1;
                memcpy(d + (size_t)i * w, src, len);
            } else if (kSchema[c].type == O_BOOL) {
                d[i] = rows[i].sudo_used ? 1 : 0;
            } else {
                memcpy(d + (size_t)i * w, src, w);
            }
        }
        t->col[c] = d;
    }
    return t;
}

void oracle_table_free(oracle_table *t) {
    if (!t) return;
    for (int c = 0; c < ORACLE_NUM_COLS; c++) free((void *)t->col[c]);
    free(t);
}

/* ---- CSV loader ---------------------------------------------------------------------------- */
/* parseCSVField (engine/serial/buildEngine-serial.c:111-151) */
static int csv_field(char **cursor, char *out) {
    char *s = *cursor;
    if (*s == '\0' || *s == '\n' || *s == '\r') return 0;
    int i = 0, quoted = 0;
    if (*s == '"') {
        quoted = 1;
        s++;
    }
    while (*s != '\0' && *s != '\n' && *s != '\r') {
        if (quoted) {
            if (*s == '"') {
                if (s[1] == '"') {
                    out[i++] = '"';
                    s += 2;
                } else {
                    quoted = 0;
                    s++;
                }
            } else {
                out[i++] = *s++;
            }
        } else if (*s == ',') {
            s++;
            break;
        } else {
            out[i++] = *s++;
        }
    }
    out[i] = '\0';
    *cursor = s;
    return 1;
}

static void copy_text(char *dst, size_t cap, const char *src) {
    /* strncpy(dst, src, cap) of the reference, but always terminated (see ingest.cpp header) */
    size_t len = strlen(src);
    if (len > cap - 1) len = cap - 1;
    memcpy(dst, src, len);
}

/* getRecordFromLine (buildEngine-serial.c:159-221) */
static void record_from_line(char *line, record *r) {
    char tok[1100];
    char *cur = line;
    memset(r, 0, sizeof *r);
    if (csv_field(&cur, tok)) r->command_id = strtoull(tok, NULL, 10);
    if (csv_field(&cur, tok)) copy_text(r->raw_command, sizeof r->raw_command, tok);
    if (csv_field(&cur, tok)) copy_text(r->base_command, sizeof r->base_command, tok);
    if (csv_field(&cur, tok)) copy_text(r->shell_type, sizeof r->shell_type, tok);
    if (csv_field(&cur, tok)) r->exit_code = atoi(tok);
    if (csv_field(&cur, tok)) copy_text(r->timestamp, sizeof r->timestamp, tok);
    if (csv_field(&cur, tok)) r->sudo_used = (strcasecmp(tok, "true") == 0 || strcmp(tok, "1") == 0);
    if (csv_field(&cur, tok)) copy_text(r->working_directory, sizeof r->working_directory, tok);
    if (csv_field(&cur, tok)) r->user_id = atoi(tok);
    if (csv_field(&cur, tok)) copy_text(r->user_name, sizeof r->user_name, tok);
    if (csv_field(&cur, tok)) copy_text(r->host_name, sizeof r->host_name, tok);
    if (csv_field(&cur, tok)) r->risk_level = atoi(tok);
}

/* getAllRecordsFromFile (buildEngine-serial.c:70-108) */
record *oracle_load_csv(const char *path, long long *n_out) {
    FILE *f = fopen(path, "r");
    if (!f) return NULL;
    char line[1024];
    record *rows = NULL;
    long long n = 0, cap = 0;
    int first = 1;
    while (fgets(line, sizeof line, f)) {
        if (first) {
            first = 0;
            continue;
        }
        if (n == cap) {
            cap = cap ? cap * 2 : 1024;
            rows = realloc(rows, sizeof(record) * (size_t)cap);
        }
        record_from_line(line, &rows[n++]);
    }
    fclose(f);
    if (!rows) rows = malloc(sizeof(record));
    *n_out = n;
    return rows;
}

void oracle_free(void *p) { free(p); }
