/*
 * executeEngine-gpu.h -- B200 execute engine: the fourth engine beside serial / omp / mpi.
 *
 * Same shape as the reference's include/executeEngine-omp.h:6-53 and -mpi.h: the shared
 * structs come from the serial header (here: qpe_abi.h, which steps aside when the
 * reference's executeEngine-serial.h was included first) and every entry point carries the
 * mode suffix.  Each prototype names the reference function it replaces.
 *
 * All functions are C-ABI (`extern "C"`, plain pointers and sizes) and are exported by
 * libqpegpu.so.  There is no CPU fallback: when no CUDA device is usable the entry points
 * print a diagnostic to stderr and initializeEngineGPU returns NULL.
 */
#ifndef EXECUTE_ENGINE_GPU_H
#define EXECUTE_ENGINE_GPU_H

#include "qpe_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Replaces initializeEngineSerial (engine/serial/executeEngine-serial.c:727-771).
 * Loads the CSV with the serial loader's exact field rules, builds the columnar table in HBM
 * and one flattened index per indexed attribute.  attribute_types: 0=u64 1=int 2=string 3=bool
 * (mapAttributeType, engine/serial/buildEngine-serial.c:224-237).  The returned pointer is the
 * head of a larger private object; release it with destroyEngineGPU only. */
struct engineS *initializeEngineGPU(int num_indexes, const char *indexed_attributes[],
                                    const int attribute_types[], const char *datafile,
                                    const char *tableName);

/* Replaces destroyEngineSerial (executeEngine-serial.c:774-814). */
void destroyEngineGPU(struct engineS *engine);

/* Replaces executeQuerySelectSerial (executeEngine-serial.c:328-528): same candidate rule
 * (index path iff a top-level condition names a u64/int index, else full scan), same row
 * order, same cell text.  selectItems NULL / numSelectItems 0 means the 12 schema columns.
 * tableName is ignored, as in the reference.  whereClause and selectItems are borrowed.
 * The result is heap-owned by the caller and is released with freeResultSet(). */
struct resultSetS *executeQuerySelectGPU(struct engineS *engine, const char **selectItems,
                                         int numSelectItems, const char *tableName,
                                         struct whereClauseS *whereClause);

/* Replaces executeQueryDeleteSerial (executeEngine-serial.c:627-715): full-scan match mask,
 * stable compaction of every column, index maintenance, CSV rewritten in the serial format. */
struct resultSetS *executeQueryDeleteGPU(struct engineS *engine, const char *tableName,
                                         struct whereClauseS *whereClause);

/* Replaces executeQueryInsertSerial (executeEngine-serial.c:538-617): same validation, same
 * CSV line appended, row appended to the table, indexes updated.  *r is copied. */
bool executeQueryInsertGPU(struct engineS *engine, const char *tableName, const record *r);

/* Replaces addAttributeIndexSerial (executeEngine-serial.c:825-841).  Note the reference
 * returns FALSE on success (inverted test at :833-840); this keeps that return value. */
bool addAttributeIndexGPU(struct engineS *engine, const char *tableName,
                          const char *attributeName, int attributeType);

/* Un-suffixed helpers every reference engine defines (executeEngine-serial.c:844-851,
 * :881-908; executeEngine-omp.c:985,1012): position of the attribute among the indexes or -1,
 * and the result destructor the bridge calls (connectEngine.c:203,229). */
int isAttributeIndexed(struct engineS *engine, const char *attributeName);
void freeResultSet(struct resultSetS *result);

#ifdef __cplusplus
}
#endif

#endif /* EXECUTE_ENGINE_GPU_H */
