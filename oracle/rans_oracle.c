/*
 * oracle/rans_oracle.c -- CPU restatement of the reference's entropy coder.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the
 * reported CPU baseline.
 *
 * What it restates (all citations into /root/reference):
 *   rans/rans.pyx:13-22   geometry constants
 *   rans/rans.pyx:25-26   logistic()         (generated C: rans/rans.cpp:1283-1329)
 *   rans/rans.pyx:31-35   CDF()              (generated C: rans/rans.cpp:1395-1484)
 *   rans/rans.pyx:37-67   encode()           (generated C: rans/rans.cpp:1668-1845)
 *   rans/rans.pyx:69-110  decode()           (generated C: rans/rans.cpp:2005-2362)
 * The float/double promotions follow the *generated C++*, not the .pyx source text
 * (SURVEY.md Appendix A.2).  The transcendental is the host libm expf(), exactly as
 * the reference links it (expf@GLIBC_2.27); flic_oracle_expf_restated() is the
 * published glibc >= 2.27 algorithm (sysdeps/ieee754/flt-32/e_expf.c) written out so
 * the device function can be compared with both.
 *
 * Pinning: this restatement is checked bit-for-bit against the reference's own
 * rans.pyx, re-cythonised from /root/reference by oracle/build_ref.py into
 * oracle/_ref/ (tests/test_oracle_pinning.py), and against the committed golden
 * vectors under tests/golden/ that were produced by that build.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -pthread (see oracle/Makefile).
 * -ffp-contract=off matters: the reference is built for baseline x86-64 (no FMA).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FLIC_OK 0
#define FLIC_ERR_ZERO_SCALE 1  /* reference: ZeroDivisionError "float division" (rans.cpp:1435) */
#define FLIC_ERR_ZERO_FREQ 2   /* reference: ZeroDivisionError "integer division" (rans.cpp:1826) */
#define FLIC_ERR_NEGATIVE 3    /* reference: OverflowError, negative -> unsigned (rans.cpp:2194) */
#define FLIC_ERR_UNDERRUN 4    /* reference: unchecked operator[] (rans.cpp:2109); UB there */

/* rans/rans.pyx:13-22 */
static const uint64_t RANS_L = 0x100000000ull;
static const uint64_t RANS_MASK = 0xffffffffull;
static const uint64_t RANS_M = 0x1000000ull;

/* ---- glibc expf, restated ------------------------------------ */
static const uint64_t EXP2F_T[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

static inline uint64_t as_u64(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double as_f64(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint32_t as_u32(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* use_fma: evaluate the polynomial the way glibc's FMA ifunc variant contracts it. */
static float expf_restated_impl(float x, int use_fma)
{
    const double N = 32.0;
    const double InvLn2N = 0x1.71547652b82fep+0 * N;
    const double SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / N / N / N;
    const double C1 = 0x1.ebfce50fac4f3p-3 / N / N;
    const double C2 = 0x1.62e42ff0c52d6p-1 / N;
    uint32_t abstop = (as_u32(x) >> 20) & 0x7ff;
    double xd = (double)x;
    if (abstop >= (as_u32(88.0f) >> 20)) {
        if (as_u32(x) == as_u32(-INFINITY)) return 0.0f;
        if (abstop >= (as_u32(INFINITY) >> 20)) return x + x;
        if (x > 0x1.62e42ep6f) return INFINITY;
        if (x < -0x1.9fe368p6f) return 0.0f;
    }
    double z = InvLn2N * xd;
    volatile double kdv = z + SHIFT; /* "needs to be double" */
    double kd = kdv;
    uint64_t ki = as_u64(kd);
    kd -= SHIFT;
    double r = z - kd;
    uint64_t t = EXP2F_T[ki % 32];
    t += ki << 47;
    double s = as_f64(t);
    double y;
    if (use_fma) {
        double zz = fma(C0, r, C1);
        double r2 = r * r;
        y = fma(C2, r, 1.0);
        y = fma(zz, r2, y);
    } else {
        double zz = C0 * r + C1;
        double r2 = r * r;
        y = C2 * r + 1.0;
        y = zz * r2 + y;
    }
    y = y * s;
    return (float)y;
}

float flic_oracle_expf_restated(float x) { return expf_restated_impl(x, 0); }
float flic_oracle_expf_restated_fma(float x) { return expf_restated_impl(x, 1); }
float flic_oracle_expf_host(float x) { return expf(x); }

/* Count inputs with |x| <= bound where the restatement differs from host expf.
 * Sweeps float bit patterns [lo_bits, hi_bits) (callers split by sign). */
uint64_t flic_oracle_expf_sweep(uint32_t lo_bits, uint32_t hi_bits, int use_fma,
                                uint32_t *first_bad, uint32_t max_bad)
{
    uint64_t bad = 0;
    for (uint64_t b = lo_bits; b < hi_bits; ++b) {
        uint32_t u = (uint32_t)b;
        float x;
        memcpy(&x, &u, 4);
        float a = expf(x), c = expf_restated_impl(x, use_fma);
        if (as_u32(a) != as_u32(c)) {
            if (bad < max_bad) first_bad[bad] = u;
            ++bad;
        }
    }
    return bad;
}

/* got[i] is some implementation's expf(float with bit pattern lo_bits + i); count disagreements
 * with the host libm and with the restatement. */
void flic_oracle_expf_compare(uint32_t lo_bits, int64_t n, const float *got, uint64_t *bad_vs_host,
                              uint64_t *bad_vs_restated, uint32_t *first_bad, uint32_t max_bad)
{
    uint64_t bh = 0, br = 0;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t u = lo_bits + (uint32_t)i;
        float x;
        memcpy(&x, &u, 4);
        uint32_t g = as_u32(got[i]);
        if (g != as_u32(expf(x))) {
            if (bh < max_bad) first_bad[bh] = u;
            ++bh;
        }
        if (g != as_u32(expf_restated_impl(x, 1))) ++br;
    }
    *bad_vs_host = bh;
    *bad_vs_restated = br;
}

/* ---- logistic / CDF (rans/rans.pyx:25-35 with rans.cpp promotions) ------------ */

/* rans.cpp:1301-1306: argument float32, expf float32, sum and divide in double. */
static inline double logistic_f(float x) { return 1.0 / (1.0 + (double)expf(-x)); }

/* rans.cpp:1418-1449.  Returns part1+part2; *err set when scale == 0. */
static inline int cdf_ref(float x, float mean, float scale, float lower, int *err)
{
    int part2 = (int)round((double)(x - lower) * 256.0) + 1;
    double t4 = ((double)x + (0.5 / 256.0)) - (double)mean;
    if (scale == 0.0f) { *err = FLIC_ERR_ZERO_SCALE; return 0; }
    float arg = (float)(t4 / (double)scale);
    double prod = logistic_f(arg) * 16775168.0; /* PyFloat * PyLong(M - 2048) */
    int part1 = (int)roundf((float)prod);
    return part1 + part2;
}

/* got[i] is some implementation's part1 for the float argument with bit pattern lo_bits + i:
 * (int) roundf((float)(logistic(arg) * 16775168.0)), rans/rans.pyx:25-26,34 with the promotions of
 * rans.cpp:1301-1306,1441-1449.  Counts disagreements; NaN arguments are skipped. */
void flic_oracle_part1_compare(uint32_t lo_bits, int64_t n, const int32_t *got, uint64_t *bad,
                               uint32_t *first_bad, uint32_t max_bad)
{
    uint64_t b = 0;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t u = lo_bits + (uint32_t)i;
        float arg;
        memcpy(&arg, &u, 4);
        if (arg != arg) continue;
        double prod = logistic_f(arg) * 16775168.0;
        int want = (int)roundf((float)prod);
        if (got[i] != want) {
            if (b < max_bad) first_bad[b] = u;
            ++b;
        }
    }
    *bad = b;
}

int flic_oracle_cdf(float x, float mean, float scale, float lower)
{
    int err = 0;
    return cdf_ref(x, mean, scale, lower, &err);
}

/* rans.pyx:51 / rans.cpp:1683: C double round(), half away from zero. */
static inline int lower_i_ref(float mean) { return (int)round((double)mean * 256.0 - 1024.0); }

int flic_oracle_lower(float mean) { return lower_i_ref(mean); }

/* encode pass 1 (rans.pyx:49-56): per-symbol (lower_i, start, freq).
 * start/freq are the values pushed into vector<unsigned long long> (int -> u64). */
int flic_oracle_tables(const float *x, const float *mean, const float *scale, int64_t n,
                       int32_t *lower_out, uint64_t *start_out, uint64_t *freq_out)
{
    for (int64_t i = 0; i < n; ++i) {
        int err = 0;
        int li = lower_i_ref(mean[i]);
        float lower = (float)((double)li / 256.0);
        int start = cdf_ref((float)((double)x[i] - (1.0 / 256.0)), mean[i], scale[i], lower, &err);
        int end = cdf_ref(x[i], mean[i], scale[i], lower, &err);
        if (err) return err;
        if (lower_out) lower_out[i] = li;
        start_out[i] = (uint64_t)(long long)start;
        freq_out[i] = (uint64_t)(long long)(end - start);
    }
    return FLIC_OK;
}

/* encode (rans.pyx:37-67).  buf must hold n words (<=1 word per symbol). */
int flic_oracle_encode(uint64_t state, int64_t n, const float *x, const float *mean,
                       const float *scale, uint32_t *buf, int64_t *n_words, uint64_t *state_out)
{
    int64_t nw = 0;
    for (int64_t i = 0; i < n; ++i) {
        int err = 0;
        int li = lower_i_ref(mean[i]);
        float lower = (float)((double)li / 256.0);
        int start_i = cdf_ref((float)((double)x[i] - (1.0 / 256.0)), mean[i], scale[i], lower, &err);
        int end_i = cdf_ref(x[i], mean[i], scale[i], lower, &err);
        if (err) return err;
        uint64_t cdf = (uint64_t)(long long)start_i;
        uint64_t freq = (uint64_t)(long long)(end_i - start_i);
        if (state >= (freq << 40)) {
            buf[nw++] = (uint32_t)(state & RANS_MASK);
            state >>= 32;
        }
        if (freq == 0) return FLIC_ERR_ZERO_FREQ;
        state = ((state / freq) << 24) + (state % freq) + cdf;
    }
    *n_words = nw;
    *state_out = state;
    return FLIC_OK;
}

/* decode (rans.pyx:69-110).  buf / mean / scale are in the order the reference's
 * caller passes them: REVERSED (trainer.py:317).  msg comes out reversed too. */
int flic_oracle_decode(uint64_t state, const uint32_t *buf, int64_t n_buf, int64_t n,
                       const float *mean, const float *scale, float *msg, uint64_t *state_out,
                       int64_t *words_used)
{
    int64_t pos = 0;
    for (int64_t i = 0; i < n; ++i) {
        int err = 0;
        if (state < RANS_L) {
            if (pos >= n_buf) return FLIC_ERR_UNDERRUN;
            state = (state << 32) | buf[pos];
            pos += 1;
        }
        uint64_t mod = state & 0xffffff;
        int lower = lower_i_ref(mean[i]);
        int upper = lower + 2047;
        float lower_f = (float)((double)lower / 256.0);
        while (lower <= upper) {
            int s = (lower + upper) >> 1;
            int c = cdf_ref((float)((double)s / 256.0), mean[i], scale[i], lower_f, &err);
            if (err) return err;
            if (c < 0) return FLIC_ERR_NEGATIVE;
            if ((uint64_t)c > mod) upper = s - 1; else lower = s + 1;
        }
        int s = lower;
        msg[i] = (float)((double)s / 256.0);
        int c0 = cdf_ref((float)((double)(s - 1) / 256.0), mean[i], scale[i], lower_f, &err);
        int c1 = cdf_ref((float)((double)s / 256.0), mean[i], scale[i], lower_f, &err);
        if (err) return err;
        if (c0 < 0 || c1 - c0 < 0) return FLIC_ERR_NEGATIVE;
        uint64_t cdf_s = (uint64_t)c0, freq_s = (uint64_t)(c1 - c0);
        state = (state >> 24) * freq_s + (state & 0xffffff) - cdf_s;
    }
    *state_out = state;
    if (words_used) *words_used = pos;
    return FLIC_OK;
}

/* ---- many independent streams (the partition the GPU path codes) --------------
 * Stream s owns symbols [offsets[s], offsets[s+1]) and starts at state 1<<32
 * (trainer.py:310).  Words of stream s are written to words + offsets[s]
 * (worst case one word per symbol) in emission order; counts[s] says how many.
 * Threads split the stream range; used for the all-cores CPU baseline. */
typedef struct {
    const float *x, *mean, *scale;
    const int64_t *offsets;
    int64_t s0, s1;
    uint32_t *words;
    int64_t *counts;
    uint64_t *states;
    int32_t *status;
    float *out; /* decode */
    int decode;
} stream_job;

static void *stream_worker(void *p)
{
    stream_job *j = (stream_job *)p;
    for (int64_t s = j->s0; s < j->s1; ++s) {
        int64_t a = j->offsets[s], b = j->offsets[s + 1], n = b - a;
        if (!j->decode) {
            int64_t nw = 0;
            uint64_t st = RANS_L;
            int rc = flic_oracle_encode(RANS_L, n, j->x + a, j->mean + a, j->scale + a,
                                        j->words + a, &nw, &st);
            j->counts[s] = nw;
            j->states[s] = st;
            j->status[s] = rc;
        } else {
            /* Reverse the stream's inputs the way trainer.py:317 does. */
            int64_t nw = j->counts[s];
            float *rm = (float *)malloc(sizeof(float) * (size_t)(n ? n : 1));
            float *rs = (float *)malloc(sizeof(float) * (size_t)(n ? n : 1));
            float *ro = (float *)malloc(sizeof(float) * (size_t)(n ? n : 1));
            uint32_t *rb = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nw ? nw : 1));
            for (int64_t i = 0; i < n; ++i) { rm[i] = j->mean[a + n - 1 - i]; rs[i] = j->scale[a + n - 1 - i]; }
            for (int64_t i = 0; i < nw; ++i) rb[i] = j->words[a + nw - 1 - i];
            uint64_t st = 0;
            int rc = flic_oracle_decode(j->states[s], rb, nw, n, rm, rs, ro, &st, NULL);
            for (int64_t i = 0; i < n; ++i) j->out[a + i] = ro[n - 1 - i];
            j->states[s] = st;
            j->status[s] = rc;
            free(rm); free(rs); free(ro); free(rb);
        }
    }
    return NULL;
}

static int run_streams(stream_job base, int64_t n_streams, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    stream_job *jobs = (stream_job *)malloc(sizeof(stream_job) * (size_t)n_threads);
    int64_t per = (n_streams + n_threads - 1) / n_threads;
    int started = 0;
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = base;
        jobs[t].s0 = (int64_t)t * per;
        jobs[t].s1 = jobs[t].s0 + per < n_streams ? jobs[t].s0 + per : n_streams;
        if (jobs[t].s0 >= jobs[t].s1) break;
        if (n_threads == 1) stream_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, stream_worker, &jobs[t]);
        ++started;
    }
    if (n_threads > 1) for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return FLIC_OK;
}

int flic_oracle_encode_streams(const float *x, const float *mean, const float *scale,
                               const int64_t *offsets, int64_t n_streams, uint32_t *words,
                               int64_t *counts, uint64_t *states, int32_t *status, int n_threads)
{
    stream_job j;
    memset(&j, 0, sizeof j);
    j.x = x; j.mean = mean; j.scale = scale; j.offsets = offsets;
    j.words = words; j.counts = counts; j.states = states; j.status = status; j.decode = 0;
    return run_streams(j, n_streams, n_threads);
}

/* states[] holds the final encoder states on entry and the decoder's end states
 * (1<<32 for an intact stream) on return. */
int flic_oracle_decode_streams(const float *mean, const float *scale, const int64_t *offsets,
                               int64_t n_streams, uint32_t *words, int64_t *counts,
                               uint64_t *states, int32_t *status, float *out, int n_threads)
{
    stream_job j;
    memset(&j, 0, sizeof j);
    j.mean = mean; j.scale = scale; j.offsets = offsets;
    j.words = words; j.counts = counts; j.states = states; j.status = status; j.out = out;
    j.decode = 1;
    return run_streams(j, n_streams, n_threads);
}
