/*
 * qpe_abi.h -- binary interface shared with the reference's C engines.
 *
 * The GPU engine is called by the reference's bridge (connectEngine.c:125-245) and drivers
 * (QPESeq.c:16-97) with the reference's own structs, so the types below are laid out
 * byte-for-byte like the reference's:
 *
 *   record          include/logType.h:11-24        (1040-byte row; offsets asserted below)
 *   FieldType       include/recordSchema.h:12-17
 *   KeyType, KEY_T  include/bplus.h:15-30
 *   engineS         include/executeEngine-serial.h:15-25
 *   resultSetS      include/executeEngine-serial.h:30-38
 *   whereClauseS    include/executeEngine-serial.h:48-56
 *
 * When this header is compiled inside the reference tree, include the reference's
 * "executeEngine-serial.h" FIRST: its include guard (EXECUTE_ENGINE_SERIAL_H) is detected and
 * the definitions below are skipped, so the two never collide.
 */
#ifndef QPE_ABI_H
#define QPE_ABI_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef EXECUTE_ENGINE_SERIAL_H /* not inside the reference tree: define the shared types */
#define QPE_ABI_OWNS_TYPES 1

/* One command-log row, array-of-structs form (what INSERT receives and the CSV describes). */
typedef struct record {
    unsigned long long command_id;
    char raw_command[512];
    char base_command[100];
    char shell_type[20];
    int exit_code;
    char timestamp[30];
    bool sudo_used;
    char working_directory[200];
    int user_id;
    char user_name[50];
    char host_name[100];
    int risk_level;
} record;

typedef enum { FIELD_UINT64, FIELD_INT, FIELD_STRING, FIELD_BOOL } FieldType;

typedef enum { KEY_INT, KEY_UINT64, KEY_BOOL, KEY_STRING } KeyType;

typedef struct {
    KeyType type;
    union {
        uint64_t u64;
        int i32;
        bool b;
        const char *str;
    } v;
} KEY_T;

typedef void *ROW_PTR;
typedef struct node node; /* B+ tree node of the CPU engines; opaque (always NULL) for the GPU engine */

/* Engine handle head.  The GPU engine allocates a larger object whose FIRST member is this
 * struct, so code that only reads the public fields keeps working. */
struct engineS {
    char *tableName;
    node **bplus_tree_roots; /* GPU engine: array of num_indexes NULLs (index lives in HBM) */
    int num_indexes;
    char **indexed_attributes;
    FieldType *attribute_types;
    record **all_records; /* GPU engine: NULL (rows live as columns in HBM) */
    int num_records;
    char *datafile;
    void *record_block; /* GPU engine: NULL */
};

struct resultSetS {
    int numRecords;
    int numColumns;
    char **columnNames;
    FieldType *columnTypes;
    char ***data; /* data[row][col], NUL-terminated text */
    double queryTime;
    bool success;
};

struct whereClauseS {
    const char *attribute; /* NULL for a parenthesised group (see sub) */
#ifdef __cplusplus
    const char *op_; /* the C name of this member is `operator` */
#else
    const char *operator;
#endif
    const char *value;
    int value_type;
    struct whereClauseS *next;
    const char *logical_op; /* joins this node to ->next: "AND" / "OR" / NULL */
    struct whereClauseS *sub;
};

#endif /* EXECUTE_ENGINE_SERIAL_H */

#ifdef __cplusplus
}
#endif

#if defined(QPE_ABI_OWNS_TYPES) && !defined(QPE_ABI_NO_LAYOUT_CHECK)
#ifdef __cplusplus
#define QPE_SA(c, m) static_assert(c, m)
#else
#define QPE_SA(c, m) _Static_assert(c, m)
#endif
QPE_SA(sizeof(record) == 1040, "record must be 1040 bytes (logType.h:11-24)");
QPE_SA(offsetof(record, raw_command) == 8, "raw_command");
QPE_SA(offsetof(record, base_command) == 520, "base_command");
QPE_SA(offsetof(record, shell_type) == 620, "shell_type");
QPE_SA(offsetof(record, exit_code) == 640, "exit_code");
QPE_SA(offsetof(record, timestamp) == 644, "timestamp");
QPE_SA(offsetof(record, sudo_used) == 674, "sudo_used");
QPE_SA(offsetof(record, working_directory) == 675, "working_directory");
QPE_SA(offsetof(record, user_id) == 876, "user_id");
QPE_SA(offsetof(record, user_name) == 880, "user_name");
QPE_SA(offsetof(record, host_name) == 930, "host_name");
QPE_SA(offsetof(record, risk_level) == 1032, "risk_level");
QPE_SA(sizeof(KEY_T) == 16, "KEY_T is a 16-byte tagged union (bplus.h:22-30)");
#undef QPE_SA
#endif

#endif /* QPE_ABI_H */
