/* orc_bench.c -- multi-threaded driver that times the oracle over a set of streams.
 * TEST INFRASTRUCTURE ONLY: used by bench.py's cpu_baseline / --impl reference legs. */
#define _GNU_SOURCE
#include "sse_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    const uint8_t *arena; const uint64_t *off; const uint32_t *len; const uint8_t *mode;
    size_t n_streams; int n_threads, tid; uint64_t out_bytes, frames, chunks_ok;
} job;

static void *worker(void *p) {
    job *j = (job *)p;
    orc_result *r = orc_result_new();
    for (size_t s = (size_t)j->tid; s < j->n_streams; s += (size_t)j->n_threads) {
        orc_result_clear(r);
        if (j->mode[s] & 1) orc_reframe_stream(r, j->arena + j->off[s], j->len[s], 0);
        else orc_passthrough(r, j->arena + j->off[s], j->len[s], (j->mode[s] & 2) != 0);
        j->out_bytes += r->out_len;
        for (size_t i = 0; i < r->n_lines; i++) if (r->lines[i].out_len) j->frames++;
        for (size_t i = 0; i < r->n_chunks; i++) j->chunks_ok += r->chunks[i].json_ok;
    }
    orc_result_free(r);
    return 0;
}

/* Runs every stream through the oracle with n_threads threads; returns seconds. mode bit0: R, bit1: parse.
 * The oracle keeps its builder/decoder scratch in _Thread_local globals, so threads are independent. */
double orc_bench_run(const uint8_t *arena, const uint64_t *off, const uint32_t *len, const uint8_t *mode,
                     size_t n_streams, int n_threads, uint64_t *out_bytes, uint64_t *frames, uint64_t *chunks_ok) {
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof *th);
    job *jobs = (job *)calloc((size_t)n_threads, sizeof *jobs);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (job){ arena, off, len, mode, n_streams, n_threads, t, 0, 0, 0 };
        pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    *out_bytes = *frames = *chunks_ok = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], 0);
        *out_bytes += jobs[t].out_bytes; *frames += jobs[t].frames; *chunks_ok += jobs[t].chunks_ok;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
