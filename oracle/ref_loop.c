/*
 * ref_loop.c -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A tight C driver that dlopen()s a library exporting the reference's seven
 * symbols (C/ChiSquare.h:31-49) -- normally oracle/_ref/libChiSq.so, the
 * unmodified reference compiled in place -- and calls calculateLnLiklihood
 * once per walker with the box-prior test of mda.py:1574-1576 in front of it.
 * It exists for two things only:
 *   * the "pure C, no Python" secondary CPU-baseline line of bench.py
 *     (SURVEY.md section 8d), and
 *   * fast generation of reference outputs for larger golden / parity cases.
 * Ex bins are the unit of parallelism, exactly as in the reference
 * (multiprocessing.Pool over bins, mda.py:186-188): thread t owns bins
 * t, t+T, t+2T, ... and builds its own structs (a handle is not re-entrant,
 * C/ChiSquare.c:183).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef void *(*make_fn)(int, int);
typedef void (*free_fn)(void *);
typedef void (*setdata_fn)(void *, double *);
typedef void (*setdist_fn)(void *, int, double *);
typedef double (*calc_fn)(void *, double *);

typedef struct {
    make_fn make; free_fn release; setdata_fn set_data; setdist_fn set_dist;
    calc_fn lnl; calc_fn chi;
} ref_api;

typedef struct {
    const ref_api *api;
    int tid, n_threads;
    int B, L, n_pad, W, repeats, mode;          /* mode 0: lnpost, 1: chi */
    const int *n_pts;
    const double *consts, *lo, *hi, *params;
    double *out;
    pthread_barrier_t *start, *stop;
} job;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *worker(void *arg)
{
    job *jb = (job *)arg;
    const ref_api *api = jb->api;
    const int L = jb->L, W = jb->W, n_pad = jb->n_pad;
    const size_t bin_stride = (size_t)(L + 1) * (size_t)n_pad;
    int n_mine = 0;
    for (int b = jb->tid; b < jb->B; b += jb->n_threads) ++n_mine;
    void **handles = (void **)calloc((size_t)(n_mine > 0 ? n_mine : 1), sizeof(void *));
    double *row = (double *)malloc(sizeof(double) * (size_t)(n_pad > 0 ? n_pad : 1));
    double *a = (double *)malloc(sizeof(double) * (size_t)L);
    int k = 0;
    for (int b = jb->tid; b < jb->B; b += jb->n_threads, ++k) {
        const int n = jb->n_pts[b];
        const double *blk = jb->consts + (size_t)b * bin_stride;
        handles[k] = api->make(n, L);
        memcpy(row, blk, sizeof(double) * (size_t)n);
        api->set_data(handles[k], row);
        for (int l = 0; l < L; ++l) {
            memcpy(row, blk + (size_t)(l + 1) * n_pad, sizeof(double) * (size_t)n);
            api->set_dist(handles[k], l, row);
        }
    }
    pthread_barrier_wait(jb->start);
    for (int rep = 0; rep < jb->repeats; ++rep) {
        k = 0;
        for (int b = jb->tid; b < jb->B; b += jb->n_threads, ++k) {
            const double *p = jb->params + (size_t)b * W * L;
            double *o = jb->out + (size_t)b * W;
            for (int w = 0; w < W; ++w) {
                memcpy(a, p + (size_t)w * L, sizeof(double) * (size_t)L);
                if (jb->mode == 1) { o[w] = api->chi(handles[k], a); continue; }
                int outside = 0;
                for (int l = 0; l < L; ++l)
                    if (a[l] < jb->lo[l] || a[l] > jb->hi[l]) { outside = 1; break; }
                o[w] = outside ? -INFINITY : api->lnl(handles[k], a);
            }
        }
    }
    pthread_barrier_wait(jb->stop);
    for (int i = 0; i < n_mine; ++i) api->release(handles[i]);
    free(handles); free(row); free(a);
    return NULL;
}

/*
 * Evaluate lnpost (mode 0) or chi^2 (mode 1) for params [B][W][L] against
 * consts [B][1+L][n_pad] (row 0 data, rows 1..L dists, zero padded) with the
 * library at lib_path, `repeats` times, on n_threads threads.  Returns the
 * wall-clock seconds of the evaluation region only (struct set-up excluded),
 * or a negative value on failure.
 */
double ref_loop_run(const char *lib_path, int mode, int B, int L, int n_pad,
                    const int *n_pts, const double *consts,
                    const double *lo, const double *hi,
                    const double *params, int W, double *out,
                    int n_threads, int repeats)
{
    void *h = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "ref_loop: dlopen(%s): %s\n", lib_path, dlerror()); return -1.0; }
    ref_api api;
    api.make = (make_fn)dlsym(h, "makeMdaStruct");
    api.release = (free_fn)dlsym(h, "freeMdaStruct");
    api.set_data = (setdata_fn)dlsym(h, "setMdaData");
    api.set_dist = (setdist_fn)dlsym(h, "setMdaDist");
    api.lnl = (calc_fn)dlsym(h, "calculateLnLiklihood");
    api.chi = (calc_fn)dlsym(h, "calculateChi");
    if (!api.make || !api.release || !api.set_data || !api.set_dist || !api.lnl || !api.chi) {
        fprintf(stderr, "ref_loop: %s lacks a reference symbol\n", lib_path);
        dlclose(h);
        return -2.0;
    }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > B) n_threads = B > 0 ? B : 1;
    pthread_barrier_t start, stop;
    pthread_barrier_init(&start, NULL, (unsigned)n_threads + 1u);
    pthread_barrier_init(&stop, NULL, (unsigned)n_threads + 1u);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    job *jobs = (job *)malloc(sizeof(job) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) {
        job jb = { &api, t, n_threads, B, L, n_pad, W, repeats, mode,
                   n_pts, consts, lo, hi, params, out, &start, &stop };
        jobs[t] = jb;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    pthread_barrier_wait(&start);
    const double t0 = now_s();
    pthread_barrier_wait(&stop);
    const double t1 = now_s();
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&start);
    pthread_barrier_destroy(&stop);
    free(th); free(jobs);
    dlclose(h);
    return t1 - t0;
}
