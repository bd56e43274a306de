/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin entry points around the UNMODIFIED reference (compiled in place from
 * /root/reference by oracle/Makefile into oracle/_ref/) so that tests and
 * bench.py's cpu_baseline leg can drive it:
 *
 *   - as a shared library (libqpe_ref.so) from Python/ctypes: parse a SQL text
 *     with the reference's own tokenizer (tokenizer/src/tokenizer.c:8,191),
 *     convert it with the reference's convert_conditions (connectEngine.c:65)
 *     and run the reference's executeQuerySelectSerial / DeleteSerial /
 *     linearSearchRecords on it;
 *   - as an executable (qpe_ref_dump): QPESeq.c's main loop (QPESeq.c:16-97)
 *     with the query file, the row limit and the number of indexes taken from
 *     argv instead of being hard-coded (SURVEY.md App. D "full-dump harness").
 *
 * This file contains no reference code: it only calls the reference's public
 * functions through the reference's own headers.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "executeEngine-serial.h"
#include "connectEngine.h"
#include "sql.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* engine with the first `num_indexes` of the reference's optimalIndexes[] (connectEngine.c:48-62) */
struct engineS *ref_open(const char *csv, int num_indexes) {
    if (num_indexes > numOptimalIndexes) num_indexes = numOptimalIndexes;
    return initializeEngineSerial(num_indexes, optimalIndexes,
                                  (const int *)optimalIndexTypes, csv, TABLE_NAME);
}

/* engine with caller-chosen indexes */
struct engineS *ref_open_idx(const char *csv, int n, const char **attrs, const int *types) {
    return initializeEngineSerial(n, attrs, types, csv, TABLE_NAME);
}

void ref_close(struct engineS *e) { destroyEngineSerial(e); }

int ref_num_records(struct engineS *e) { return e->num_records; }

/* SELECT through tokenizer -> parser -> convert_conditions -> executeQuerySelectSerial.
 * Returns the reference's own resultSetS (free with ref_free_result). NULL if not a SELECT. */
struct resultSetS *ref_select(struct engineS *e, const char *sql) {
    Token tokens[MAX_TOKENS];
    if (tokenize(sql, tokens, MAX_TOKENS) <= 0) return NULL;
    ParsedSQL parsed = parse_tokens(tokens);
    if (parsed.command != CMD_SELECT) return NULL;
    const char *items[10];
    int n = 0;
    if (!parsed.select_all)
        for (n = 0; n < parsed.num_columns; n++) items[n] = parsed.columns[n];
    struct whereClauseS *wc = convert_conditions(&parsed);
    struct resultSetS *r = executeQuerySelectSerial(e, items, n, parsed.table, wc);
    free_where_clause_list(wc);
    return r;
}

/* DELETE through the same front end; returns rows affected or -1 */
int ref_delete(struct engineS *e, const char *sql) {
    Token tokens[MAX_TOKENS];
    if (tokenize(sql, tokens, MAX_TOKENS) <= 0) return -1;
    ParsedSQL parsed = parse_tokens(tokens);
    if (parsed.command != CMD_DELETE) return -1;
    struct whereClauseS *wc = convert_conditions(&parsed);
    struct resultSetS *r = executeQueryDeleteSerial(e, parsed.table, wc);
    int n = r ? r->numRecords : -1;
    if (r) freeResultSet(r);
    free_where_clause_list(wc);
    return n;
}

/* any statement through run_test_query (prints to stdout exactly as QPESeq does) */
void ref_run(struct engineS *e, const char *sql, int max_rows) { run_test_query(e, sql, max_rows); fflush(stdout); }

void ref_free_result(struct resultSetS *r) { freeResultSet(r); }

/* accessors so ctypes does not need the struct layout */
int ref_result_rows(struct resultSetS *r) { return r->numRecords; }
int ref_result_cols(struct resultSetS *r) { return r->numColumns; }
const char *ref_result_colname(struct resultSetS *r, int j) { return r->columnNames[j]; }
const char *ref_result_cell(struct resultSetS *r, int i, int j) { return r->data[i][j]; }

/* number of matches of the WHERE of `sql` on the full-scan path, and the positions
 * (index into all_records) of the matches, in result order. out may be NULL. */
int ref_scan_positions(struct engineS *e, const char *sql, int *out, int cap) {
    Token tokens[MAX_TOKENS];
    if (tokenize(sql, tokens, MAX_TOKENS) <= 0) return -1;
    ParsedSQL parsed = parse_tokens(tokens);
    struct whereClauseS *wc = convert_conditions(&parsed);
    int m = 0, k = 0;
    record **hits = linearSearchRecords(e->all_records, e->num_records, wc, &m);
    if (out) {
        /* hits is an ordered subsequence of all_records: recover positions by a merge walk */
        int p = 0;
        for (k = 0; k < m && k < cap; k++) {
            while (e->all_records[p] != hits[k]) p++;
            out[k] = p++;
        }
    }
    free(hits);
    free_where_clause_list(wc);
    return m;
}

/* wall seconds of `reps` runs of the reference's hot loop A (linearSearchRecords,
 * engine/serial/executeEngine-serial.c:854-878) on the whole table; *matches = last count */
double ref_time_scan(struct engineS *e, const char *sql, int reps, int *matches) {
    Token tokens[MAX_TOKENS];
    if (tokenize(sql, tokens, MAX_TOKENS) <= 0) return -1.0;
    ParsedSQL parsed = parse_tokens(tokens);
    struct whereClauseS *wc = convert_conditions(&parsed);
    int m = 0;
    double t0 = now_s();
    for (int r = 0; r < reps; r++) {
        record **hits = linearSearchRecords(e->all_records, e->num_records, wc, &m);
        free(hits);
    }
    double t1 = now_s();
    if (matches) *matches = m;
    free_where_clause_list(wc);
    return t1 - t0;
}

/* wall seconds of `reps` full executeQuerySelectSerial calls (match + projection) */
double ref_time_select(struct engineS *e, const char *sql, int reps, int *matches) {
    double t0 = now_s();
    int m = -1;
    for (int r = 0; r < reps; r++) {
        struct resultSetS *res = ref_select(e, sql);
        if (res) { m = res->numRecords; freeResultSet(res); }
    }
    double t1 = now_s();
    if (matches) *matches = m;
    return t1 - t0;
}

/* wall seconds of `n` findRange calls (engine/bplus.c:282-314) on the engine's index number `index_no`, with the
 * caller-allocated result arrays sized to the table as executeQuerySelectSerial sizes them (:438-439);
 * lo / hi are unsigned 64-bit (index type 0) or int (index type 1) keys; *found = rows found over all calls */
double ref_time_probe(struct engineS *e, int index_no, const unsigned long long *lo, const unsigned long long *hi,
                      int n, long long *found) {
    if (index_no < 0 || index_no >= e->num_indexes) return -1.0;
    node *root = e->bplus_tree_roots[index_no];
    const FieldType type = e->attribute_types[index_no];
    KEY_T *keys = malloc(sizeof(KEY_T) * (size_t)(e->num_records > 0 ? e->num_records : 1));
    ROW_PTR *rows = malloc(sizeof(ROW_PTR) * (size_t)(e->num_records > 0 ? e->num_records : 1));
    long long total = 0;
    double t0 = now_s();
    for (int i = 0; i < n; i++) {
        KEY_T a, b;
        a.type = b.type = (type == FIELD_UINT64) ? KEY_UINT64 : KEY_INT;
        if (type == FIELD_UINT64) { a.v.u64 = lo[i]; b.v.u64 = hi[i]; }
        else { a.v.i32 = (int)lo[i]; b.v.i32 = (int)hi[i]; }
        total += findRange(root, a, b, false, keys, rows);
    }
    double t1 = now_s();
    free(keys);
    free(rows);
    if (found) *found = total;
    return t1 - t0;
}

/* the whereClauseS list the reference's tokenizer + parser + convert_conditions build for `sql`,
 * rendered as text (same rendering as qpe_sql_where_to_text in the product's front end);
 * also reports the parsed command. Caller frees. */
static void render_where(const struct whereClauseS *w, char *out, size_t cap) {
    for (; w; w = w->next) {
        size_t n = strlen(out);
        if (w->sub) {
            snprintf(out + n, cap - n, "( ");
            render_where(w->sub, out, cap);
            n = strlen(out);
            snprintf(out + n, cap - n, " )");
        } else {
            snprintf(out + n, cap - n, "%s %s %s", w->attribute ? w->attribute : "<null>",
                     w->operator ? w->operator : "<null>", w->value ? w->value : "<null>");
        }
        if (w->next) {
            n = strlen(out);
            snprintf(out + n, cap - n, " %s ", w->logical_op ? w->logical_op : "<none>");
        }
    }
}

char *ref_where_text(const char *sql, int *command_out) {
    Token tokens[MAX_TOKENS];
    memset(tokens, 0, sizeof tokens);
    char *out = calloc(1, 8192);
    if (command_out) *command_out = -1;
    if (tokenize(sql, tokens, MAX_TOKENS) <= 0) return out;
    ParsedSQL parsed = parse_tokens(tokens);
    if (command_out) *command_out = (int)parsed.command;
    if (parsed.command != CMD_SELECT && parsed.command != CMD_DELETE) return out;
    struct whereClauseS *wc = convert_conditions(&parsed);
    render_where(wc, out, 8192);
    free_where_clause_list(wc);
    return out;
}

void ref_free(void *p) { free(p); }

#ifdef REF_HARNESS_MAIN
/* qpe_ref_dump <csv> <query-file> [max_rows=0] [num_indexes=5]
 * Same statement splitting as QPESeq.c:74-82 (strtok on ';', leading-space trim). */
int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <csv> <query-file> [max_rows] [num_indexes]\n", argv[0]);
        return 2;
    }
    int max_rows = argc > 3 ? atoi(argv[3]) : 0;
    int nidx = argc > 4 ? atoi(argv[4]) : numOptimalIndexes;
    double t0 = now_s();
    struct engineS *e = ref_open(argv[1], nidx);
    double t1 = now_s();
    FILE *fp = fopen(argv[2], "r");
    if (!fp) { perror("query file"); return 1; }
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    char *buf = malloc((size_t)sz + 1);
    if (fread(buf, 1, (size_t)sz, fp) != (size_t)sz) { perror("read"); return 1; }
    buf[sz] = 0;
    fclose(fp);
    for (char *q = strtok(buf, ";"); q; q = strtok(NULL, ";")) {
        q = trim(q);
        if (*q) run_test_query(e, q, max_rows);
    }
    double t2 = now_s();
    free(buf);
    ref_close(e);
    fprintf(stderr, "ref_dump: init %.4f s, queries %.4f s (wall)\n", t1 - t0, t2 - t1);
    return 0;
}
#endif
