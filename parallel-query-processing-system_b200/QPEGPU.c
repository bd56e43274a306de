/* QPEGPU.c -- the fourth driver, beside the reference's QPESeq.c / QPEOMP.c / QPEMPI.c.
 *
 * Same protocol as QPESeq.c:16-97: build the engine over a CSV with the five default indexes
 * (connectEngine.c:48-62), read a query file, split it on ';', trim leading blanks, run every
 * statement with the 20-row print limit (connectEngine.h:14) and print the timing banner.
 * Plain C over the C-ABI of libqpegpu.so only.
 *
 *   QPEGPU [csv] [query-file] [max_rows]      defaults: data-generation/commands_50k.csv,
 *                                             sample-queries.txt, 20
 */
#define _POSIX_C_SOURCE 200809L
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "executeEngine-gpu.h"
#include "qpe_gpu.h"

#define CYAN "\x1b[36m"
#define YELLOW "\x1b[33m"
#define BOLD "\x1b[1m"
#define RESET "\x1b[0m"

static const char *kIndexes[] = {"command_id", "user_id", "risk_level", "exit_code", "sudo_used"};
static const int kIndexTypes[] = {0, 1, 1, 1, 3};

static double wall(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv) {
    const char *csv = argc > 1 ? argv[1] : "data-generation/commands_50k.csv";
    const char *qfile = argc > 2 ? argv[2] : "sample-queries.txt";
    const int max_rows = argc > 3 ? atoi(argv[3]) : 20;

    const double t0 = wall();
    struct engineS *engine = initializeEngineGPU(5, kIndexes, kIndexTypes, csv, "commands");
    if (!engine) {
        fprintf(stderr, "QPEGPU: engine initialisation failed: %s\n", qpe_gpu_last_error());
        return EXIT_FAILURE;
    }
    const double t_init = wall() - t0;

    FILE *fp = fopen(qfile, "r");
    if (!fp) {
        perror("Failed to open query file");
        destroyEngineGPU(engine);
        return EXIT_FAILURE;
    }
    fseek(fp, 0, SEEK_END);
    const long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    char *text = malloc((size_t)size + 1);
    if (!text || fread(text, 1, (size_t)size, fp) != (size_t)size) {
        perror("Failed to read query file");
        fclose(fp);
        destroyEngineGPU(engine);
        return EXIT_FAILURE;
    }
    text[size] = '\0';
    fclose(fp);
    const double t_load = wall() - t0;

    for (char *q = strtok(text, ";"); q; q = strtok(NULL, ";")) {
        while (*q && isspace((unsigned char)*q)) q++;
        if (*q) qpe_sql_run(engine, q, max_rows, NULL);
    }
    free(text);
    destroyEngineGPU(engine);

    const double t_total = wall() - t0;
    printf(CYAN "======= B200 GPU Execution Summary =======" RESET "\n");
    printf(CYAN "Engine Initialization Time: " RESET YELLOW "%.4f seconds\n" RESET, t_init);
    printf(CYAN "Query Loading Time: " RESET YELLOW "%.4f seconds\n" RESET, t_load - t_init);
    printf(CYAN "Query Execution Time: " RESET YELLOW "%.4f seconds\n" RESET, t_total - t_load);
    printf(BOLD CYAN "Total Execution Time: " RESET BOLD YELLOW "%.4f seconds" RESET "\n", t_total);
    printf(CYAN "========================================" RESET "\n");
    return EXIT_SUCCESS;
}
