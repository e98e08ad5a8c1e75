/*
 * C / pthreads restatement of the reference's self-consistent update and gradient — TEST INFRASTRUCTURE.
 *
 * Same arithmetic as pymbar/mbar_solvers.py:231-242 (Eq. C3) and :284-292 (Eq. C6): two max-shifted
 * log-sum-exp sweeps over u_kn [K, N] (row-major), first over states per sample (weights N_k, rows with
 * N_k = 0 excluded before the max, as scipy.special.logsumexp(b=...) does), then over samples per state.
 * Purpose: (i) a multi-threaded CPU baseline beside the single-threaded numpy port (the reference itself is
 * single-threaded numpy), (ii) a faster checker than numpy for mid-size cases.  Never linked into the
 * product.  Build: `make -C oracle` -> oracle/_ref/libmbar_oracle.so.  (This image has no libgomp, hence
 * plain pthreads.)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

typedef struct {
    const double *u, *N_k, *f;
    double *L, *out;
    int64_t K, N, lo, hi;
} job_t;

static int n_threads(void) {
    const char* e = getenv("MBAR_ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}
int mbar_oracle_threads(void) { return n_threads(); }

static void run(void* (*fn)(void*), job_t proto, int64_t total) {
    int T = n_threads();
    if ((int64_t)T > total) T = (int)(total > 0 ? total : 1);
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < T; ++t) {
        jobs[t] = proto;
        jobs[t].lo = total * t / T;
        jobs[t].hi = total * (t + 1) / T;
        if (t > 0) pthread_create(&th[t], NULL, fn, &jobs[t]);
    }
    fn(&jobs[0]);
    for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
}

/* L_n = log sum_k N_k exp(f_k - u_kn)  (mbar_solvers.py:238) for n in [lo, hi) */
static void* denom_job(void* p) {
    const job_t* j = (const job_t*)p;
    for (int64_t n = j->lo; n < j->hi; ++n) {
        double m = -INFINITY;
        for (int64_t k = 0; k < j->K; ++k)
            if (j->N_k[k] > 0) {
                const double a = j->f[k] - j->u[k * j->N + n];
                if (a > m) m = a;
            }
        double s = 0.0;
        for (int64_t k = 0; k < j->K; ++k)
            if (j->N_k[k] > 0) s += j->N_k[k] * exp(j->f[k] - j->u[k * j->N + n] - m);
        j->L[n] = m + log(s);
    }
    return NULL;
}

/* f'_k = -logsumexp_n(-L_n - u_kn)  (mbar_solvers.py:240) for k in [lo, hi) */
static void* numer_job(void* p) {
    const job_t* j = (const job_t*)p;
    for (int64_t k = j->lo; k < j->hi; ++k) {
        const double* uk = j->u + k * j->N;
        double m = -INFINITY;
        for (int64_t n = 0; n < j->N; ++n) {
            const double a = -j->L[n] - uk[n];
            if (a > m) m = a;
        }
        double s = 0.0;
        for (int64_t n = 0; n < j->N; ++n) s += exp(-j->L[n] - uk[n] - m);
        j->out[k] = -(m + log(s));
    }
    return NULL;
}

void mbar_oracle_self_consistent_update(const double* u, int64_t K, int64_t N, const double* N_k, const double* f,
                                        double* f_out) {
    double* L = (double*)malloc((size_t)N * sizeof(double));
    job_t proto = {u, N_k, f, L, f_out, K, N, 0, 0};
    run(denom_job, proto, N);
    run(numer_job, proto, K);
    free(L);
}

/* g_k = -N_k (1 - exp(f_k + logsumexp_n(-L_n - u_kn)))  (mbar_solvers.py:290-292) */
void mbar_oracle_gradient(const double* u, int64_t K, int64_t N, const double* N_k, const double* f, double* g) {
    double* fo = (double*)malloc((size_t)K * sizeof(double));
    mbar_oracle_self_consistent_update(u, K, N, N_k, f, fo);
    for (int64_t k = 0; k < K; ++k) g[k] = -N_k[k] * (1.0 - exp(f[k] - fo[k]));
    free(fo);
}
