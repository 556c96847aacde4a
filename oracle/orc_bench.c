/* orc_bench.c -- multi-threaded driver that times the oracle over a set of streams.
 * TEST INFRASTRUCTURE ONLY: used by bench.py's cpu_baseline / --impl reference legs.
 * A pool of n_threads threads is started once per call; the clock runs from the moment all of them stand at the start
 * barrier until the last one has finished `passes` passes over the streams (streams are handed out in blocks of 16 from
 * a shared counter, so cores that are slower or shared do not hold the others up). */
#define _GNU_SOURCE
#include "sse_oracle.h"
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    const uint8_t *arena; const uint64_t *off; const uint32_t *len; const uint8_t *mode;
    size_t n_streams; int passes;
    atomic_size_t *next;            /* one work counter per pass */
    pthread_barrier_t *start;
    uint64_t out_bytes, frames, chunks_ok;
} job;

static void *worker(void *p) {
    job *j = (job *)p;
    orc_result *r = orc_result_new();
    pthread_barrier_wait(j->start);
    for (int ps = 0; ps < j->passes; ps++) {
        for (;;) {
            size_t s0 = atomic_fetch_add(&j->next[ps], 16);
            if (s0 >= j->n_streams) break;
            size_t s1 = s0 + 16 < j->n_streams ? s0 + 16 : j->n_streams;
            for (size_t s = s0; s < s1; s++) {
                orc_result_clear(r);
                if (j->mode[s] & 1) orc_reframe_stream(r, j->arena + j->off[s], j->len[s], 0);
                else orc_passthrough(r, j->arena + j->off[s], j->len[s], (j->mode[s] & 2) != 0);
                j->out_bytes += r->out_len;
                for (size_t i = 0; i < r->n_lines; i++) if (r->lines[i].out_len) j->frames++;
                for (size_t i = 0; i < r->n_chunks; i++) j->chunks_ok += r->chunks[i].json_ok;
            }
        }
    }
    orc_result_free(r);
    return 0;
}

/* Runs every stream through the oracle `passes` times with n_threads threads; returns the seconds of all passes together.
 * mode bit0: R, bit1: parse. Totals over all passes in out_bytes / frames / chunks_ok.
 * The oracle keeps its builder/decoder scratch in _Thread_local globals, so threads are independent. */
double orc_bench_run(const uint8_t *arena, const uint64_t *off, const uint32_t *len, const uint8_t *mode,
                     size_t n_streams, int n_threads, int passes, uint64_t *out_bytes, uint64_t *frames, uint64_t *chunks_ok) {
    if (n_threads < 1) n_threads = 1;
    if (passes < 1) passes = 1;
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof *th);
    job *jobs = (job *)calloc((size_t)n_threads, sizeof *jobs);
    atomic_size_t *next = (atomic_size_t *)calloc((size_t)passes, sizeof *next);
    pthread_barrier_t start;
    pthread_barrier_init(&start, 0, (unsigned)n_threads + 1u);
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (job){ arena, off, len, mode, n_streams, passes, next, &start, 0, 0, 0 };
        pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    struct timespec t0, t1;
    pthread_barrier_wait(&start);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    *out_bytes = *frames = *chunks_ok = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], 0);
        *out_bytes += jobs[t].out_bytes; *frames += jobs[t].frames; *chunks_ok += jobs[t].chunks_ok;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    pthread_barrier_destroy(&start);
    free(th); free(jobs); free(next);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
