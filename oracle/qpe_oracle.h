/*
 * qpe_oracle.h -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's SELECT/WHERE path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load liboracle.so.  The product (libqpegpu.so) never links, loads or calls anything here.
 *
 * Pinned against: the known-answer tests of the reference's own test programs
 * (tests/executeEngine-serial-test.c:39,70,113; tests/duplicate-test.c:29,40,54;
 * tests/bplus-serial-test.c:40,43; tests/delete-test.c:28-105), against the compiled UNMODIFIED
 * reference (oracle/_ref/libqpe_ref.so) on every probe query, and against the golden vectors
 * under tests/golden/ that were produced by that compiled reference (tools/make_golden.py).
 */
#ifndef QPE_ORACLE_H
#define QPE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/qpe_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_NUM_COLS 12

/* Columnar table, same physical layout as the device table: column c is an array of n cells of
 * width[c] bytes -- u64 (8), int32 (4), bool (1), or NUL-padded text (each cell holds a
 * terminator).  Column order = struct order of `record` (include/logType.h:11-24). */
typedef struct oracle_table {
    long long n;
    const void *col[ORACLE_NUM_COLS];
    unsigned width[ORACLE_NUM_COLS];
} oracle_table;

int oracle_col_by_name(const char *name); /* -1 if unknown */

/* evaluateWhereClause (engine/serial/executeEngine-serial.c:292-316) on row `row`. */
int oracle_eval_row(const oracle_table *t, long long row, const struct whereClauseS *wc);

/* linearSearchRecords (:854-878): ordered filter over the whole table. Returns the match count;
 * ids (may be NULL) receives the matching positions in table order. */
long long oracle_scan(const oracle_table *t, const struct whereClauseS *wc, uint32_t *ids);

/* Same, over rows [first, first + n) only (used to time a bounded sample). */
long long oracle_scan_range(const oracle_table *t, long long first, long long n, const struct whereClauseS *wc,
                            uint32_t *ids);

/* Leaf-chain order of the reference's B+ tree on column `col` (u64 / int columns): positions
 * sorted by (key ascending, position descending) -- engine/bplus.c:317-358, :471-490, :723-740.
 * perm must hold n entries.  Returns 0, or -1 for a column that cannot be indexed. */
int oracle_index_order(const oracle_table *t, int col, uint32_t *perm);

/* The same order produced by literally replaying the reference's insertion rule (each new entry
 * goes before all existing entries with an equal or greater key): O(n^2) worst case, for small n.
 * Exists to validate oracle_index_order's closed form. */
int oracle_index_order_by_insertion(const oracle_table *t, int col, uint32_t *perm);

/* findRange (engine/bplus.c:282-314) on a leaf chain given as `perm`: inclusive [lo, hi].
 * Keys are passed as KEY_T (type tag KEY_UINT64 / KEY_INT). *first / *count delimit the slice. */
int oracle_find_range(const oracle_table *t, int col, const uint32_t *perm, KEY_T lo, KEY_T hi, long long *first,
                      long long *count);

/* Match phase of executeQuerySelectSerial (:328-476) with the candidate rule of :358-459:
 * indexes are (attribute name, type code 0=u64 1=int 2=string 3=bool) in build order.
 * *ids_out receives a malloc'ed array (free with oracle_free) of positions in result order.
 * Returns the number of results, or -1 on error. *used_index = 1 when the index path ran. */
long long oracle_select(const oracle_table *t, int num_indexes, const char *const *index_attrs,
                        const int *index_types, const struct whereClauseS *wc, uint32_t **ids_out,
                        int *used_index);

/* get_attribute_string_value (:216-248): text of cell (row, attribute) into buf (cap bytes). */
void oracle_cell_text(const oracle_table *t, long long row, const char *attribute, char *buf, size_t cap);

/* Array-of-structs -> columnar (malloc'ed columns at the given widths; text widths must be
 * >= longest value + 1).  Pass widths == NULL for the struct's own field sizes rounded to 16. */
oracle_table *oracle_table_from_records(const record *rows, long long n, const unsigned *widths);
void oracle_table_free(oracle_table *t);

/* CSV loader of engine/serial/buildEngine-serial.c:70-221 (fgets(1024) chunks, quoted fields,
 * strtoull/atoi/strncpy field rules).  Returns malloc'ed rows, *n_out rows; NULL if unreadable. */
record *oracle_load_csv(const char *path, long long *n_out);

void oracle_free(void *p);

#ifdef __cplusplus
}
#endif

#endif /* QPE_ORACLE_H */
