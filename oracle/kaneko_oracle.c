/* TEST INFRASTRUCTURE ONLY -- see kaneko_oracle.h.  Plain C11, single thread.
 *
 * The arithmetic, iteration order, stale-state behaviour and counters of the
 * reference are kept exactly (SURVEY.md 8c lists the quirks); only names,
 * memory management and structure are ours.  Build with -fwrapv.
 */
#include "kaneko_oracle.h"
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LOG0 ULONG_MAX /* stands for the reference's log[0] = LONG_MAX sentinel (main.cpp:68) */

typedef unsigned long gf_t; /* the reference stores field elements in unsigned long */

/* 2x2 polynomial matrix of the Sugiyama recursion (headers/Decoder.h:10-19). */
typedef struct {
    gf_t *ff, *fs, *sf, *ss;
    int nff, nfs, nsf, nss;
} polymat;

struct ko_code {
    int m, n, k, t, gsize;
    gf_t *alog, *log;
    uint8_t *g;
    long J; /* <0: uncapped */
    /* RNG: std::default_random_engine == minstd_rand0, state x (bchCoder.cpp:20) */
    uint64_t rng;
    /* Kaneko state (headers/KanekoKernelProcessor.h:16-31) */
    double sd;
    double *alpha, *skey;
    int *sidx;
    uint8_t *yH, *x, *err;
    long mm, mm0;
    uint64_t n_cmp, n_sum, n_dec;
    /* Decoder state (headers/Decoder.h:21-46) */
    int l;
    gf_t *synd;
    long synd_size;
    uint8_t *old_word;
    gf_t *lambda;
    int lambda_size;
    gf_t *locators;
    gf_t *s, *p, *q, *tmp, *f, *res1;
    int ns, np, nq, ntmp;
    polymat a, b;
};

/* ------------------------------------------------------------------ a1 */
/* src/main.cpp:59-78 */
static void build_field(ko_code *c) {
    static const gf_t prim[16] = {3, 7, 11, 19, 37, 67, 137, 285, 529, 1033, 2053, 4179, 8219, 17475, 32771, 69643};
    int n = c->n, m = c->m;
    c->alog = calloc((size_t)n + 1, sizeof(gf_t));
    c->log = calloc((size_t)n + 2, sizeof(gf_t));
    c->alog[0] = 1;
    c->log[0] = LOG0;
    c->log[1] = 0;
    for (int i = 1; i < n; ++i) {
        gf_t v = c->alog[i - 1] << 1;
        if ((v >> m) == 1) v ^= prim[m - 1];
        c->alog[i] = v;
        c->log[v] = (gf_t)i;
    }
}

/* ------------------------------------------------------------------ a2 */
/* GF(2)[x] helpers on byte-per-coefficient arrays, index = power of x
 * (src/bchCoder.cpp:104-226).  Sizes are "highest index + 1". */
static int p2_trim(const uint8_t *a, int na) {
    while (na > 1 && !a[na - 1]) --na;
    return na;
}
static int p2_mul(const uint8_t *a, int na, const uint8_t *b, int nb, uint8_t *r) {
    memset(r, 0, (size_t)(na + nb - 1));
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) r[i + j] ^= a[i] & b[j]; /* bchCoder.cpp:124-128 */
    return na + nb - 1;
}
/* quotient and remainder of a / b, b monic in GF(2) (bchCoder.cpp:134-184) */
static void p2_divmod(const uint8_t *a, int na, const uint8_t *b, int nb, uint8_t *quo, int *nq, uint8_t *rem, int *nr) {
    uint8_t *w = malloc((size_t)na);
    memcpy(w, a, (size_t)na);
    int s = na;
    int qn = na - nb + 1;
    if (qn < 1) qn = 1;
    memset(quo, 0, (size_t)qn);
    while (s >= nb) {
        for (int i = 0; i < nb; ++i) w[s - nb + i] ^= b[i];
        quo[s - nb] = 1;
        while (s > 0 && !w[s - 1]) --s;
    }
    *nq = na - nb + 1;
    *nr = s ? s : 1;
    for (int i = 0; i < *nr; ++i) rem[i] = w[i] ? 1 : 0;
    free(w);
}
/* minimal polynomial of alpha^i: product over the conjugates (bchCoder.cpp:25-91) */
static int min_poly(const ko_code *c, int i, uint8_t *out) {
    int n = c->n;
    gf_t poly[40] = {1}, nxt[40];
    int deg = 0;
    int e = i % n;
    do {
        gf_t root = c->alog[e];
        memset(nxt, 0, sizeof nxt);
        for (int d = 0; d <= deg; ++d) { /* poly * (x + root) */
            nxt[d + 1] ^= poly[d];
            if (poly[d]) nxt[d] ^= c->alog[(c->log[poly[d]] + c->log[root]) % (gf_t)n];
        }
        memcpy(poly, nxt, sizeof nxt);
        ++deg;
        e = (2 * e) % n;
    } while (e != i % n);
    for (int d = 0; d <= deg; ++d) out[d] = (poly[d] % 2) ? 1 : 0; /* bchCoder.cpp:82-84 */
    return deg + 1;
}
/* g(x) = lcm of the minimal polynomials of alpha^1..alpha^(2t-1) (main.cpp:80-95;
 * lcm = a*b/gcd(a,b), bchCoder.cpp:186-226) */
static void build_generator(ko_code *c) {
    int n = c->n;
    uint8_t *g = calloc((size_t)2 * n + 64, 1), *mp = calloc(64, 1);
    uint8_t *prod = calloc((size_t)2 * n + 64, 1), *a = calloc((size_t)2 * n + 64, 1), *b = calloc((size_t)2 * n + 64, 1);
    uint8_t *quo = calloc((size_t)2 * n + 64, 1), *rem = calloc((size_t)2 * n + 64, 1);
    int ng = min_poly(c, 1, g);
    for (int i = 2; i < 2 * c->t; ++i) {
        int nmp = min_poly(c, i, mp);
        /* gcd by Euclid */
        int na = ng, nb = nmp, nq, nr;
        memcpy(a, g, (size_t)ng);
        memcpy(b, mp, (size_t)nmp);
        while (!(nb == 1 && b[0] == 0)) {
            if (na < nb) { /* a mod b = a */
                uint8_t *tp = a; a = b; b = tp;
                int ti = na; na = nb; nb = ti;
                continue;
            }
            p2_divmod(a, na, b, nb, quo, &nq, rem, &nr);
            memcpy(a, b, (size_t)nb); na = nb;
            memcpy(b, rem, (size_t)nr); nb = nr;
        }
        na = p2_trim(a, na);
        int np = p2_mul(g, ng, mp, nmp, prod);
        p2_divmod(prod, np, a, na, quo, &nq, rem, &nr);
        memcpy(g, quo, (size_t)nq);
        ng = nq;
    }
    c->g = malloc((size_t)ng);
    memcpy(c->g, g, (size_t)ng);
    c->gsize = ng;
    c->k = n - ng + 1;
    free(g); free(mp); free(prod); free(a); free(b); free(quo); free(rem);
}

static void polymat_alloc(polymat *M, int len) {
    M->ff = calloc((size_t)len, sizeof(gf_t)); M->fs = calloc((size_t)len, sizeof(gf_t));
    M->sf = calloc((size_t)len, sizeof(gf_t)); M->ss = calloc((size_t)len, sizeof(gf_t));
    M->nff = M->nfs = M->nsf = M->nss = 1;
}
static void polymat_free(polymat *M) { free(M->ff); free(M->fs); free(M->sf); free(M->ss); }

ko_code *ko_create(int m, int t) {
    if (m < 2 || m > 16 || t <= 0 || t >= (1 << (m - 1))) return NULL; /* main.cpp:55 */
    ko_code *c = calloc(1, sizeof *c);
    c->m = m; c->t = t; c->n = (1 << m) - 1;
    build_field(c);
    build_generator(c);
    int n = c->n;
    c->J = -1;
    c->rng = 1; /* default-constructed minstd_rand0 */
    /* KanekoKernelProcessor ctor with snr = 0.5 (KanekoKernelProcessor.cpp:17-26, main.cpp:176) */
    {
        long k = c->k, nn = c->n;
        double signalToNoiseRatio = 0.5;
        c->sd = sqrt(1 / (pow(10, signalToNoiseRatio / 10) * 2 * k / nn));
    }
    c->alpha = calloc((size_t)n + 1, sizeof(double));
    c->skey = calloc((size_t)n + 1, sizeof(double));
    c->sidx = calloc((size_t)n + 1, sizeof(int));
    c->yH = calloc((size_t)n, 1); c->x = calloc((size_t)n, 1); c->err = calloc((size_t)n, 1);
    /* Decoder ctor (Decoder.cpp:12-38) */
    c->l = 2 * t;
    int len = 2 * c->l + 8;
    c->synd = calloc((size_t)len, sizeof(gf_t));
    c->old_word = calloc((size_t)n, 1);
    c->lambda = calloc((size_t)len, sizeof(gf_t));
    c->locators = calloc((size_t)len, sizeof(gf_t));
    c->s = calloc((size_t)len, sizeof(gf_t)); c->p = calloc((size_t)len, sizeof(gf_t));
    c->q = calloc((size_t)len, sizeof(gf_t)); c->tmp = calloc((size_t)len, sizeof(gf_t));
    c->f = calloc((size_t)len, sizeof(gf_t)); c->res1 = calloc((size_t)len, sizeof(gf_t));
    polymat_alloc(&c->a, len);
    polymat_alloc(&c->b, len);
    return c;
}

void ko_destroy(ko_code *c) {
    if (!c) return;
    free(c->alog); free(c->log); free(c->g); free(c->alpha); free(c->skey); free(c->sidx);
    free(c->yH); free(c->x); free(c->err); free(c->synd); free(c->old_word); free(c->lambda);
    free(c->locators); free(c->s); free(c->p); free(c->q); free(c->tmp); free(c->f); free(c->res1);
    polymat_free(&c->a); polymat_free(&c->b);
    free(c);
}

void ko_info(const ko_code *c, int *n, int *k, int *t, int *gsize, uint8_t *g_out) {
    if (n) *n = c->n;
    if (k) *k = c->k;
    if (t) *t = c->t;
    if (gsize) *gsize = c->gsize;
    if (g_out) memcpy(g_out, c->g, (size_t)c->gsize);
}

void ko_tables(const ko_code *c, uint64_t *antilog_out, uint64_t *log_out) {
    for (int i = 0; i < c->n; ++i) antilog_out[i] = c->alog[i];
    for (int i = 0; i <= c->n; ++i) log_out[i] = (i == 0) ? (uint64_t)LONG_MAX : c->log[i];
}

void ko_set_J(ko_code *c, long J) { c->J = J; }

/* ------------------------------------------------------------------ a3-a5 */
/* minstd_rand0: x <- 16807 x mod (2^31 - 1); range [1, 2^31-2] */
static inline uint64_t rng_next(ko_code *c) {
    c->rng = (c->rng * 16807ULL) % 2147483647ULL;
    return c->rng;
}
void ko_seed(ko_code *c, uint64_t seed) {
    /* linear_congruential_engine::seed: if (s mod m) == 0 -> 1 */
    uint64_t s = seed % 2147483647ULL;
    c->rng = s ? s : 1;
}
/* libstdc++ uniform_int_distribution<unsigned short>(0,1) on minstd_rand0: the
 * generic "downscaling" branch (bits/uniform_int_dist.h). */
static inline unsigned rng_bit(ko_code *c) {
    const uint64_t urngrange = 2147483646ULL - 1ULL;
    const uint64_t uerange = 2;
    const uint64_t scaling = urngrange / uerange;
    const uint64_t past = uerange * scaling;
    uint64_t r;
    do r = rng_next(c) - 1ULL; while (r >= past);
    return (unsigned)(r / scaling);
}
/* libstdc++ generate_canonical<double,53>(minstd_rand0): two draws. */
static inline double rng_canonical(ko_code *c) {
    const long double r = 2147483646.0L;
    double sum = 0, tmp = 1;
    for (int k = 2; k != 0; --k) {
        sum += (double)(rng_next(c) - 1ULL) * tmp;
        tmp = (double)((long double)tmp * r);
    }
    double ret = sum / tmp;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}
/* One draw of a FRESH std::normal_distribution<double>(0, sd): Marsaglia polar; the
 * object is re-created per addNoise call (bchCoder.cpp:244) so the saved second
 * variate alternates within a frame and is dropped at the frame end. */
typedef struct { int have; double saved; } normal_state;
static inline double rng_normal(ko_code *c, normal_state *st, double sd) {
    double ret;
    if (st->have) {
        st->have = 0;
        ret = st->saved;
    } else {
        double x, y, r2;
        do {
            x = 2.0 * rng_canonical(c) - 1.0;
            y = 2.0 * rng_canonical(c) - 1.0;
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        double mult = sqrt(-2 * log(r2) / r2);
        st->saved = x * mult;
        st->have = 1;
        ret = y * mult;
    }
    return ret * sd + 0.0;
}

/* src/bchCoder.cpp:120-132 with first = info, second = g */
static void encode_one(const ko_code *c, const uint8_t *info, uint8_t *cw) {
    memset(cw, 0, (size_t)c->n);
    for (int i = 0; i < c->k; ++i)
        for (int j = 0; j < c->gsize; ++j) cw[i + j] ^= info[i] & c->g[j];
}
void ko_encode(const ko_code *c, const uint8_t *info, long B, uint8_t *cw) {
    for (long f = 0; f < B; ++f) encode_one(c, info + f * c->k, cw + f * c->n);
}
/* src/dataForPlot.cpp:45-48 */
void ko_gen_frames(ko_code *c, double ebn0_db, long B, uint8_t *info, uint8_t *cw, double *y) {
    long k = c->k, n = c->n;
    for (long f = 0; f < B; ++f) {
        double sd = sqrt(1 / (pow(10, ebn0_db / 10) * 2 * k / n));
        for (long i = 0; i < k; ++i) info[f * k + i] = (uint8_t)rng_bit(c);
        encode_one(c, info + f * k, cw + f * n);
        normal_state st = {0, 0.0};
        for (long i = 0; i < n; ++i) y[f * n + i] = (cw[f * n + i] ? 1 : -1) + rng_normal(c, &st, sd);
    }
}

/* ------------------------------------------------------------------ a8-a12 */
/* src/Decoder.cpp:184-207 */
static void dec_syndromes(ko_code *c, const uint8_t *word) {
    long n = c->n, start = n - 1;
    while (!word[start] && start > 0) start--;
    c->synd_size = 0;
    for (gf_t i = 1; i <= (gf_t)c->l; ++i) {
        gf_t v = word[start];
        for (long j = start - 1; j >= 0; --j) {
            if (v) {
                gf_t s = i + c->log[v];
                v = (s < (gf_t)n) ? c->alog[s] : c->alog[s - n];
            }
            v ^= (gf_t)word[j];
        }
        c->synd[i - 1] = v;
        if (v) c->synd_size = (long)i;
    }
    memcpy(c->old_word, word, (size_t)n);
}
/* src/Decoder.cpp:210-230 */
static void dec_alter_syndromes(ko_code *c, const uint8_t *word) {
    int n = c->n;
    for (int i = 0; i < n; ++i) {
        if (c->old_word[i] != word[i]) {
            gf_t residual = c->log[c->old_word[i] ^ word[i]]; /* = log[1] = 0 */
            for (int j = 1; j <= c->l; ++j) {
                if (residual != LOG0) {
                    gf_t s = residual + (gf_t)((j * i) % n);
                    c->synd[j - 1] ^= (s < (gf_t)n) ? c->alog[s] : c->alog[s - n];
                }
            }
        }
    }
    c->synd_size = c->l;
    for (int i = c->l - 1; i >= 0 && !c->synd[i]; --i) c->synd_size = i;
    memcpy(c->old_word, word, (size_t)n);
}
/* src/Decoder.cpp:71-92 -- note *nres is left untouched when the operands have
 * equal length and cancel completely. */
static void dec_poly_add(const gf_t *a, int na, const gf_t *b, int nb, gf_t *r, int *nres) {
    int mn = na < nb ? na : nb;
    for (int i = 0; i < mn; ++i) {
        r[i] = a[i] ^ b[i];
        if (r[i]) *nres = i + 1;
    }
    if (na > mn) {
        for (int i = mn; i < na; ++i) r[i] = a[i];
        *nres = na;
    } else if (nb > mn) {
        for (int i = mn; i < nb; ++i) r[i] = b[i];
        *nres = nb;
    }
}
/* src/Decoder.cpp:94-110 */
static void dec_poly_mul(const ko_code *c, const gf_t *a, int na, const gf_t *b, int nb, gf_t *r, int *nres) {
    gf_t n = (gf_t)c->n;
    for (int i = 0; i < na + nb - 1; ++i) r[i] = 0;
    for (int i = 0; i < na; ++i) {
        gf_t la = c->log[a[i]];
        for (int j = 0; j < nb; ++j) {
            if (la != LOG0 && b[j]) {
                gf_t s = la + c->log[b[j]];
                r[i + j] ^= (s < n) ? c->alog[s] : c->alog[s - n];
            }
        }
    }
    *nres = na + nb - 1;
}
/* src/Decoder.cpp:112-161: (a, b) <- (b, a mod b), quotient to q */
static void dec_poly_divstep(ko_code *c, gf_t *a, int *na, gf_t *b, int *nb, gf_t *q, int *nq) {
    gf_t n = (gf_t)c->n;
    gf_t *f = c->f, *res1 = c->res1;
    for (int i = 0; i < *na; ++i) f[i] = a[i];
    int s = *na, s2 = *nb;
    gf_t coeff = LOG0;
    for (int i = 0; i < s - s2 + 1; ++i) res1[i] = 0;
    while (s >= s2) {
        if (f[s - 1]) {
            gf_t d = n + c->log[f[s - 1]] - c->log[b[s2 - 1]];
            coeff = (d < n) ? d : d - n;
        }
        for (int i = 0; i < *nb; ++i) {
            if (coeff != LOG0 && b[i]) {
                gf_t e = coeff + c->log[b[i]];
                f[s - s2 + i] ^= (e < n) ? c->alog[e] : c->alog[e - n];
            }
        }
        res1[s - s2] = c->alog[coeff];
        while (s > 0 && f[s - 1] == 0) --s;
    }
    *nq = *na - *nb + 1;
    for (int i = 0; i < *nq; ++i) q[i] = res1[i];
    *na = *nb;
    for (int i = 0; i < *nb; ++i) a[i] = b[i];
    *nb = s ? s : 1;
    for (int i = 0; i < *nb; ++i) b[i] = f[i];
}
/* src/Decoder.cpp:164-180 (Horner) */
static gf_t dec_eval(const ko_code *c, const gf_t *poly, int size, gf_t elem) {
    gf_t n = (gf_t)c->n;
    gf_t v = poly[size - 1];
    if (size < 2) return v;
    gf_t le = c->log[elem];
    for (int i = size - 2; i >= 0; --i) {
        if (v && le != LOG0) {
            gf_t s = le + c->log[v];
            v = (s < n) ? c->alog[s] : c->alog[s - n];
        }
        v ^= poly[i];
    }
    return v;
}
/* src/Decoder.cpp:233-277 (Sugiyama) */
static int dec_euclid(ko_code *c) {
    int l = c->l, t = c->t;
    c->ns = l + 1;
    for (int i = 0; i < l; ++i) c->s[i] = 0;
    c->s[l] = 1;
    memcpy(c->p, c->synd, (size_t)c->synd_size * sizeof(gf_t));
    c->np = (int)c->synd_size;
    polymat *a = &c->a, *b = &c->b;
    for (int i = 0; i <= l; ++i) a->ff[i] = a->fs[i] = a->sf[i] = a->ss[i] = 0;
    a->ff[0] = 1; a->ss[0] = 1;
    a->nff = a->nfs = a->nsf = a->nss = 1;
    while (c->np > t) {
        dec_poly_divstep(c, c->s, &c->ns, c->p, &c->np, c->q, &c->nq);
        dec_poly_add(a->sf, a->nsf, NULL, 0, b->ff, &b->nff);
        dec_poly_add(a->ss, a->nss, NULL, 0, b->fs, &b->nfs);
        dec_poly_mul(c, c->q, c->nq, a->sf, a->nsf, c->tmp, &c->ntmp);
        dec_poly_add(a->ff, a->nff, c->tmp, c->ntmp, b->sf, &b->nsf);
        dec_poly_mul(c, c->q, c->nq, a->ss, a->nss, c->tmp, &c->ntmp);
        dec_poly_add(a->fs, a->nfs, c->tmp, c->ntmp, b->ss, &b->nss);
        polymat sw = *a; *a = *b; *b = sw; /* std::swap(tempA, a) */
    }
    if (!a->ss[0]) return 0;
    c->lambda_size = a->nss;
    memcpy(c->lambda, a->ss, (size_t)a->nss * sizeof(gf_t));
    return 1;
}
/* src/Decoder.cpp:279-296 (Chien with the `count == size - 2` acceptance test) */
static int dec_roots(ko_code *c) {
    int size = c->lambda_size, count = 0;
    gf_t k = 0, n = (gf_t)c->n;
    for (int i = 0; i < size; ++i) {
        while (k < n && dec_eval(c, c->lambda, size, c->alog[k]) != 0) ++k;
        if (k < n) {
            c->locators[i] = (n - k) % n;
            count = i;
            ++k;
        }
    }
    return count == size - 2;
}
/* src/Decoder.cpp:298-321 */
static int dec_decode(ko_code *c, const uint8_t *word, uint8_t *answer) {
    if (!dec_euclid(c)) return 0;
    if (!dec_roots(c)) return 0;
    int n = c->n;
    for (int i = 0; i < n; ++i) answer[i] = 0;
    for (int i = 0; i < c->lambda_size - 1; ++i) answer[c->locators[i]] = 1;
    for (int i = 0; i < n; ++i) answer[i] ^= word[i];
    return 1;
}

void ko_bdd(ko_code *c, const uint8_t *words, long B, uint8_t *answers, uint8_t *ok, uint64_t *synd,
            uint64_t *lambda, int *lambda_size) {
    int n = c->n, t = c->t;
    for (long f = 0; f < B; ++f) {
        dec_syndromes(c, words + f * n);
        if (synd)
            for (int i = 0; i < 2 * t; ++i) synd[f * 2 * t + i] = c->synd[i];
        ok[f] = (uint8_t)dec_decode(c, words + f * n, answers + f * n);
        if (lambda_size) lambda_size[f] = c->lambda_size;
        if (lambda)
            for (int i = 0; i <= t; ++i) lambda[f * (t + 1) + i] = c->lambda[i];
    }
}

/* ------------------------------------------------------------------ std::sort */
/* libstdc++ introsort (bits/stl_algo.h: __sort, __introsort_loop, threshold 16,
 * __move_median_to_first, __unguarded_partition, __final_insertion_sort) on
 * pair<double,int> with a comparator that looks at .first only
 * (KanekoKernelProcessor.cpp:148,343).  Restated so that tie order is identical. */
typedef struct { double k; int i; } kv;
#define KV_LT(a, b) ((a).k < (b).k)
static void kv_swap(kv *a, kv *b) { kv t = *a; *a = *b; *b = t; }
static void kv_unguarded_linear_insert(kv *last) {
    kv val = *last;
    kv *next = last - 1;
    while (KV_LT(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void kv_insertion_sort(kv *first, kv *last) {
    if (first == last) return;
    for (kv *i = first + 1; i != last; ++i) {
        if (KV_LT(*i, *first)) {
            kv val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(kv));
            *first = val;
        } else
            kv_unguarded_linear_insert(i);
    }
}
/* heap helpers for the depth-limit fallback (__partial_sort == heap select + sort_heap) */
static void kv_push_heap(kv *first, long hole, long top, kv val) {
    long parent = (hole - 1) / 2;
    while (hole > top && KV_LT(first[parent], val)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = val;
}
static void kv_adjust_heap(kv *first, long hole, long len, kv val) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (KV_LT(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    kv_push_heap(first, hole, top, val);
}
static void kv_heapsort(kv *first, kv *last) {
    long len = last - first;
    if (len < 2) return;
    for (long parent = (len - 2) / 2;; --parent) { /* make_heap */
        kv v = first[parent];
        kv_adjust_heap(first, parent, len, v);
        if (parent == 0) break;
    }
    while (last - first > 1) { /* sort_heap */
        --last;
        kv v = *last;
        *last = *first;
        kv_adjust_heap(first, 0, last - first, v);
    }
}
static void kv_move_median_to_first(kv *result, kv *a, kv *b, kv *c) {
    if (KV_LT(*a, *b)) {
        if (KV_LT(*b, *c)) kv_swap(result, b);
        else if (KV_LT(*a, *c)) kv_swap(result, c);
        else kv_swap(result, a);
    } else if (KV_LT(*a, *c)) kv_swap(result, a);
    else if (KV_LT(*b, *c)) kv_swap(result, c);
    else kv_swap(result, b);
}
static kv *kv_unguarded_partition(kv *first, kv *last, kv *pivot) {
    for (;;) {
        while (KV_LT(*first, *pivot)) ++first;
        --last;
        while (KV_LT(*pivot, *last)) --last;
        if (!(first < last)) return first;
        kv_swap(first, last);
        ++first;
    }
}
static void kv_introsort_loop(kv *first, kv *last, long depth) {
    while (last - first > 16) {
        if (depth == 0) { kv_heapsort(first, last); return; }
        --depth;
        kv *mid = first + (last - first) / 2;
        kv_move_median_to_first(first, first + 1, mid, last - 1);
        kv *cut = kv_unguarded_partition(first + 1, last, first);
        kv_introsort_loop(cut, last, depth);
        last = cut;
    }
}
static void kv_std_sort(kv *first, kv *last) {
    if (first == last) return;
    long n = last - first, lg = 0;
    while ((1L << (lg + 1)) <= n) ++lg;
    kv_introsort_loop(first, last, 2 * lg);
    if (last - first > 16) {
        kv_insertion_sort(first, first + 16);
        for (kv *i = first + 16; i != last; ++i) kv_unguarded_linear_insert(i);
    } else
        kv_insertion_sort(first, last);
}
void ko_std_sort_pairs(double *key, int *idx, int n) {
    kv *a = malloc((size_t)n * sizeof(kv));
    for (int i = 0; i < n; ++i) { a[i].k = key[i]; a[i].i = idx[i]; }
    kv_std_sort(a, a + n);
    for (int i = 0; i < n; ++i) { key[i] = a[i].k; idx[i] = a[i].i; }
    free(a);
}

/* ------------------------------------------------------------------ a6, a7, a13 */
/* KanekoKernelProcessor.cpp:336-343 (identical in all decode flavours) */
static void kan_load(ko_code *c, const double *word) {
    int n = c->n;
    for (int i = 0; i < n; ++i) {
        c->alpha[i] = 2 * word[i] / pow(c->sd, 2);
        c->skey[i] = fabs(c->alpha[i]);
        c->sidx[i] = i;
        c->yH[i] = (c->alpha[i] <= 0.0) ? 0 : 1;
        c->alpha[i] = fabs(c->alpha[i]);
    }
    ko_std_sort_pairs(c->skey, c->sidx, n);
    c->skey[n] = 0.0; /* the reference reads one past the end in calcT(n-t); value unused */
}
/* KanekoKernelProcessor.cpp:36-51 */
static void kan_pattern(ko_code *c, long i) {
    long pos = 0;
    memset(c->err, 0, (size_t)c->n);
    while (i > 0) {
        if (i & 1) c->err[c->sidx[pos]] = 1;
        ++pos;
        i >>= 1;
    }
}
/* KanekoKernelProcessor.cpp:89-97 */
static long kan_m(const ko_code *c) {
    long cnt = 0;
    for (int i = 0; i < c->n; ++i) cnt += (c->yH[i] != c->x[i]);
    return cnt;
}
/* KanekoKernelProcessor.cpp:69-77 */
static double kan_l(const ko_code *c) {
    double l = 0;
    for (int i = 0; i < c->n; ++i)
        if (c->yH[i] != c->x[i]) l += c->alpha[i];
    return l;
}
/* KanekoKernelProcessor.cpp:54-67 */
static double kan_rhs(const ko_code *c) {
    long border = (2 * c->t + 1) - (c->mm + c->mm0) / 2;
    double l = 0;
    long i = 0, j = 0;
    while (i < border && j < c->n) {
        if (c->yH[c->sidx[j]] == c->x[c->sidx[j]]) { l += c->skey[j]; ++i; }
        ++j;
    }
    return l;
}
/* KanekoKernelProcessor.cpp:110-126 */
static double kan_T(const ko_code *c, long j) {
    long border = c->t - (c->mm + c->mm0) / 2;
    long i = 0, k = 0;
    double l = 0;
    while (i < border) {
        if (c->yH[c->sidx[k]] == c->x[c->sidx[k]]) { l += c->skey[k]; ++i; }
        ++k;
    }
    for (i = 0; i <= c->t; ++i) l += c->skey[j + i];
    return l;
}

/* KanekoKernelProcessor.cpp:335-407.  `1 << T` on an int with a runtime count is
 * x86 SHL (count masked to 5 bits); with -fwrapv the "- 1" wraps. */
static inline long pattern_bound(long T) { return (long)(int32_t)(((uint32_t)1 << ((uint32_t)T & 31u)) - 1u); }

static void kan_decode3(ko_code *c, const double *word, uint8_t *res) {
    int n = c->n, t = c->t;
    kan_load(c, word);
    long j = 0, i = 0, T = n;
    double l = 0, l0 = DBL_MAX;
    int success, first_ok = 1;
    dec_syndromes(c, c->yH);
    while (i < pattern_bound(T)) {
        kan_pattern(c, i);
        for (int k = 0; k < n; ++k) c->err[k] ^= c->yH[k];
        dec_alter_syndromes(c, c->err);
        ++c->n_dec;
        success = dec_decode(c, c->err, c->x);
        if (!i && !success) first_ok = 0;
        if (success) {
            c->mm = kan_m(c);
            if (!i || !first_ok) c->mm0 = c->mm;
            l = kan_l(c);
            if (l < l0) {
                memcpy(res, c->x, (size_t)n);
                l0 = l;
                if (l < kan_rhs(c)) return;
                while (j <= n - 1 - t && l >= kan_T(c, j)) { /* operands commute: kan_T is pure */
                    ++j;
                    ++c->n_cmp;
                    ++c->n_sum;
                }
                T = (c->J >= 0 && j > c->J) ? c->J : j;
                j = 0;
                ++c->n_cmp;
            }
        }
        ++i;
        c->n_cmp += (uint64_t)n + 6;
        c->n_sum += (uint64_t)n + 1;
    }
}

/* KanekoKernelProcessor.cpp:212-276 (uint64 bound, unconditional m/l, 2n+1 sort cost) */
static void kan_decode2(ko_code *c, const double *word, uint8_t *res) {
    int n = c->n;
    kan_load(c, word);
    c->n_sum += 2 * (uint64_t)n + 1;
    c->n_cmp += 2 * (uint64_t)n + 1;
    uint64_t j = 0, T = LONG_MAX;
    long i = 0;
    double l = 0, l0 = DBL_MAX;
    int success, first_ok = 1;
    dec_syndromes(c, c->yH);
    /* `uint64_t(1) << T` with a runtime count is x86 SHL r64 (count masked to 6 bits) */
    while ((uint64_t)i < ((T == (uint64_t)LONG_MAX) ? (uint64_t)LONG_MAX : ((uint64_t)1 << (T & 63u)))) {
        kan_pattern(c, i);
        for (int k = 0; k < n; ++k) c->err[k] ^= c->yH[k];
        dec_alter_syndromes(c, c->err);
        success = dec_decode(c, c->err, c->x);
        ++c->n_dec;
        c->mm = kan_m(c);
        if (!i && success) c->mm0 = c->mm;
        else if (!i) first_ok = 0;
        if (!first_ok) c->mm0 = c->mm;
        l = kan_l(c);
        if (success && l < l0) {
            memcpy(res, c->x, (size_t)n);
            l0 = l;
            if (l < kan_rhs(c)) return;
            while (l >= kan_T(c, (long)j)) {
                ++j;
                ++c->n_cmp;
                ++c->n_sum;
            }
            T = j;
            j = 0;
            ++c->n_cmp;
        }
        ++i;
        c->n_cmp += (uint64_t)n + 6;
        c->n_sum += (uint64_t)n + 1;
    }
}

void ko_kaneko_decode(ko_code *c, const double *y, long B, uint8_t *decided, uint32_t *trials, uint64_t *cmp,
                      uint64_t *sum) {
    for (long f = 0; f < B; ++f) {
        c->n_dec = c->n_cmp = c->n_sum = 0;
        kan_decode3(c, y + f * c->n, decided + f * c->n);
        if (trials) trials[f] = (uint32_t)c->n_dec;
        if (cmp) cmp[f] = c->n_cmp;
        if (sum) sum[f] = c->n_sum;
    }
}
void ko_kaneko_decode2(ko_code *c, const double *y, long B, uint8_t *decided, uint32_t *trials, uint64_t *cmp,
                       uint64_t *sum) {
    for (long f = 0; f < B; ++f) {
        c->n_dec = c->n_cmp = c->n_sum = 0;
        kan_decode2(c, y + f * c->n, decided + f * c->n);
        if (trials) trials[f] = (uint32_t)c->n_dec;
        if (cmp) cmp[f] = c->n_cmp;
        if (sum) sum[f] = c->n_sum;
    }
}

/* ------------------------------------------------------------------ a14 */
/* src/dataForPlot.cpp:16-115 without the file I/O: same stop rule, same counters,
 * same never-reset bit-error counter (the BER* column). */
int ko_fun(ko_code *c, long p, long e, double max_snr, double *rows, uint64_t *raw) {
    int n = c->n, k = c->k;
    int count = 0, countErr = 0, countE = 0;
    uint8_t *info = malloc((size_t)k), *res = malloc((size_t)n), *decoded = calloc((size_t)n, 1);
    double *err = malloc((size_t)n * sizeof(double));
    unsigned long words = 0;
    int pt = 0;
    c->n_dec = c->n_cmp = c->n_sum = 0;
    for (double stnr = 0.0; stnr <= max_snr; stnr += 0.5) {
        while (count < p && countErr < e) {
            ko_gen_frames(c, stnr, 1, info, res, err);
            kan_decode3(c, err, decoded);
            if (memcmp(res, decoded, (size_t)n) != 0) ++countErr;
            for (int i = 0; i < n; ++i)
                if (res[i] != decoded[i]) countE++;
            ++count;
            ++words;
        }
        if (rows) {
            rows[pt * 6 + 0] = stnr;
            rows[pt * 6 + 1] = ((double)countErr) / count;
            rows[pt * 6 + 2] = ((double)countE) / count / n;
            rows[pt * 6 + 3] = ((double)c->n_dec) / words;
            rows[pt * 6 + 4] = ((double)c->n_cmp) / words;
            rows[pt * 6 + 5] = ((double)c->n_sum) / words;
        }
        if (raw) {
            raw[pt * 6 + 0] = (uint64_t)count; raw[pt * 6 + 1] = (uint64_t)countErr; raw[pt * 6 + 2] = (uint64_t)countE;
            raw[pt * 6 + 3] = c->n_dec; raw[pt * 6 + 4] = c->n_cmp; raw[pt * 6 + 5] = c->n_sum;
        }
        c->n_dec = c->n_cmp = c->n_sum = 0;
        words = 0;
        count = 0; countErr = 0;
        ++pt;
    }
    free(info); free(res); free(decoded); free(err);
    return pt;
}

/* ------------------------------------------------------------------ kernel matrix */
/* src/bchCoder.cpp:317-345 */
void ko_make_matrix(const ko_code *c, uint8_t *out) {
    int m = c->m, len = c->n;
    int amount = ((1 << m) - 2) / 2;
    uint8_t *g = calloc((size_t)2 * len + 64, 1), *poly = calloc(64, 1);
    uint8_t *quo = calloc((size_t)2 * len + 64, 1), *rem = calloc((size_t)2 * len + 64, 1), *row = calloc((size_t)2 * len + 64, 1);
    int ng = 1;
    g[0] = 1;
    memset(out, 0, (size_t)len * len);
    out[0] = 1;
    for (int i = 2; i <= amount; ++i) {
        int np = min_poly(c, i, poly);
        if (ng >= np) {
            int nq, nr;
            p2_divmod(g, ng, poly, np, quo, &nq, rem, &nr);
            if (nr == 1 && rem[0] == 0) continue;
        }
        int nn = np + ng - 1;
        p2_mul(poly, np, g, ng, row);
        memcpy(out + (size_t)(nn - 1) * len, row, (size_t)nn);
        int shift = 1;
        for (int j = ng; j < nn - 1; ++j) {
            for (int q = 0; q < ng; ++q) out[(size_t)j * len + shift + q] = g[q];
            ++shift;
        }
        memcpy(g, row, (size_t)nn);
        ng = nn;
    }
    free(g); free(poly); free(quo); free(rem); free(row);
}

/* ------------------------------------------------------------------ extended codes / exact rules
 * NOT IN THE REFERENCE (src/main.cpp:60 builds n = 2^m - 1 only; `parity unpinned`).  CPU statement of the definition in
 * DESIGN.md that libpkb200's pk_kaneko_create_ext implements, for the GPU parity tests:
 *   ext = 1 : code (n+1, k, 2t+2), position n = overall parity of the BCH codeword.  alpha, yH and the reliability order
 *             run over all n+1 positions (std::sort replay as in kan_load); test pattern bit b flips the b-th least
 *             reliable BCH position (the parity position is skipped); the algebraic decoder sees the BCH part; the decided
 *             parity position is the parity of the decoded BCH word; m, l, calcRightSide (d = 2t+2) and calcT use all n+1
 *             positions; T starts at n (so the initial loop bound is the BCH code's).
 *   rules   : 0 = the 3-argument flavour's bookkeeping (kan_decode3), 2 = exact rules: loop bound 1 << T (64-bit),
 *             m0 = m on every success, calcRightSide border d - m, calcT without its border sum.
 * lbest (optional) receives l of the decision (DBL_MAX if none). */
void ko_ext_gen_frames(ko_code *c, int ext, double ebn0_db, long B, uint8_t *info, uint8_t *cw, double *y) {
    long k = c->k, n = c->n, ne = n + ext;
    for (long f = 0; f < B; ++f) {
        double sd = sqrt(1 / (pow(10, ebn0_db / 10) * 2 * k / ne));
        for (long i = 0; i < k; ++i) info[f * k + i] = (uint8_t)rng_bit(c);
        encode_one(c, info + f * k, cw + f * ne);
        if (ext) {
            uint8_t p = 0;
            for (long i = 0; i < n; ++i) p ^= cw[f * ne + i];
            cw[f * ne + n] = p;
        }
        normal_state st = {0, 0.0};
        for (long i = 0; i < ne; ++i) y[f * ne + i] = (cw[f * ne + i] ? 1 : -1) + rng_normal(c, &st, sd);
    }
}

void ko_ext_kaneko_decode(ko_code *c, int ext, int rules, double llr_snr_db, const double *y, long B, uint8_t *decided,
                          uint32_t *trials, uint64_t *cmp, uint64_t *sum, double *lbest) {
    const int n = c->n, ne = n + ext, t = c->t, d = 2 * t + 1 + ext;
    double *alpha = malloc((size_t)(ne + 1) * sizeof(double)), *skey = malloc((size_t)(ne + 2) * sizeof(double));
    int *sidx = malloc((size_t)(ne + 1) * sizeof(int)), *pidx = malloc((size_t)(ne + 1) * sizeof(int));
    uint8_t *yH = malloc((size_t)ne + 1), *x = malloc((size_t)ne + 1), *word = malloc((size_t)ne + 1);
    const double sd0 = sqrt(1 / (pow(10, llr_snr_db / 10) * 2 * (long)c->k / (long)ne));
    for (long f = 0; f < B; ++f) {
        const double *w = y + f * ne;
        uint8_t *res = decided + f * ne;
        uint64_t n_dec = 0, n_cmp = 0, n_sum = 0;
        for (int i = 0; i < ne; ++i) {
            const double a = 2 * w[i] / pow(sd0, 2);
            yH[i] = (a <= 0.0) ? 0 : 1;
            alpha[i] = fabs(a);
            skey[i] = alpha[i];
            sidx[i] = i;
        }
        ko_std_sort_pairs(skey, sidx, ne);
        skey[ne] = 0.0;
        int np = 0;   /* reliability order of the BCH positions alone: the pattern positions */
        for (int r = 0; r < ne; ++r)
            if (sidx[r] < n) pidx[np++] = sidx[r];
        long mm = 0, mm0 = 0, j = 0, T = n;
        uint64_t i = 0;
        double l0 = DBL_MAX;
        int first_ok = 1, have = 0;
        dec_syndromes(c, yH);
        for (;;) {
            uint64_t bound;
            if (rules == 2) bound = (T >= 63) ? UINT64_MAX : ((uint64_t)1 << T);
            else bound = (uint64_t)(int64_t)pattern_bound(T);   /* (1 << T) - 1 on an int, wrapped */
            if (rules != 2 && pattern_bound(T) < 0) bound = 0;
            if (!(i < bound)) break;
            memcpy(word, yH, (size_t)n);
            for (int b = 0; b < 63 && b < np; ++b)
                if ((i >> b) & 1) word[pidx[b]] ^= 1;
            dec_alter_syndromes(c, word);
            ++n_dec;
            const int success = dec_decode(c, word, x);
            if (!i && !success) first_ok = 0;
            if (success) {
                if (ext) {
                    uint8_t p = 0;
                    for (int q = 0; q < n; ++q) p ^= x[q];
                    x[n] = p;
                }
                mm = 0;
                double l = 0;
                for (int q = 0; q < ne; ++q)
                    if (yH[q] != x[q]) { ++mm; l += alpha[q]; }
                if (!i || !first_ok || rules == 2) mm0 = mm;
                if (l < l0) {
                    memcpy(res, x, (size_t)ne);
                    have = 1;
                    l0 = l;
                    {   /* calcRightSide */
                        const long border = d - (mm + mm0) / 2;
                        double rs = 0;
                        long cnt = 0, q = 0;
                        while (cnt < border && q < ne) {
                            if (yH[sidx[q]] == x[sidx[q]]) { rs += skey[q]; ++cnt; }
                            ++q;
                        }
                        if (l < rs) { if (trials) trials[f] = (uint32_t)n_dec; goto done; }
                    }
                    for (;;) {   /* while (j <= n-1-t && l >= calcT(j)) */
                        if (!(j <= ne - 1 - t)) break;
                        const long border = (rules == 2) ? 0 : t - (mm + mm0) / 2;
                        double tj = 0;
                        long cnt = 0, q = 0;
                        while (cnt < border && q < ne) {
                            if (yH[sidx[q]] == x[sidx[q]]) { tj += skey[q]; ++cnt; }
                            ++q;
                        }
                        for (int q2 = 0; q2 <= t; ++q2) tj += skey[j + q2];
                        if (!(l >= tj)) break;
                        ++j; ++n_cmp; ++n_sum;
                    }
                    T = (rules != 2 && c->J >= 0 && j > c->J) ? c->J : j;
                    j = 0;
                    ++n_cmp;
                }
            }
            ++i;
            n_cmp += (uint64_t)ne + 6;
            n_sum += (uint64_t)ne + 1;
        }
        if (trials) trials[f] = (uint32_t)n_dec;
    done:
        if (cmp) cmp[f] = n_cmp;
        if (sum) sum[f] = n_sum;
        if (lbest) lbest[f] = have ? l0 : DBL_MAX;
    }
    free(alpha); free(skey); free(sidx); free(pidx); free(yH); free(x); free(word);
}
