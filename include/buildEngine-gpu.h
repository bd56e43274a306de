/*
 * buildEngine-gpu.h -- build side of the B200 engine (ingest + index construction).
 *
 * Counterpart of the reference's include/buildEngine-omp.h:30-35 (mode-suffixed variants of
 * include/buildEngine-serial.h:29-95).  The CSV rules are those of
 * engine/serial/buildEngine-serial.c:70-221; see csrc/ingest.cpp for the restatement.
 */
#ifndef BUILD_ENGINE_GPU_H
#define BUILD_ENGINE_GPU_H

#include "qpe_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Replaces makeIndexSerial (buildEngine-serial.c:13-31): registers the attribute as index
 * number engine->num_indexes and builds its flattened index in HBM (u64 / int attributes; bool
 * and string attributes are registered but never probed, as in executeEngine-serial.c:425-429).
 * Returns true on success.  The caller must have reserved room in the engine's index arrays
 * (initializeEngineGPU / addAttributeIndexGPU do). */
bool makeIndexGPU(struct engineS *engine, const char *indexName, int attributeType);

/* Replaces getAllRecordsFromFile (buildEngine-serial.c:70-108): parses the CSV into malloc'ed
 * `record`s with the serial loader's exact rules.  Host-only helper (the engine itself ingests
 * straight into columns); the caller frees each row and the array. */
record **getAllRecordsFromFileGPU(const char *filepath, int *num_records);

/* Replaces getRecordFromLine (buildEngine-serial.c:159-221). Caller frees. */
record *getRecordFromLineGPU(char *line);

/* Replaces mapAttributeType (buildEngine-serial.c:224-237): 0,1,2,3 -> FieldType, else -1. */
FieldType mapAttributeTypeGPU(int attributeType);

#ifdef __cplusplus
}
#endif

#endif /* BUILD_ENGINE_GPU_H */
