// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI harness around the UNMODIFIED reference objects (compiled in place from
// /root/reference/src/*.cpp by oracle/Makefile into oracle/_ref/).  It lets the
// Python tests and bench.py's `--impl reference` arm drive the reference's own
// KanekoKernelProcessor / Decoder / bchCoder code frame by frame and dump
// golden vectors.  Nothing here restates reference algorithms except the field
// table / generator setup that lives inside the reference's main()
// (src/main.cpp:59-95), which cannot be called as a function.
//
// Build variants (see oracle/Makefile):
//   libkaneko_ref.so      HEAD semantics (T = j, src/KanekoKernelProcessor.cpp:393)
//   libkaneko_ref_cap.so  capped semantics (T = min(j, J), line 392) from a
//                         temporary sed-patched copy; J is made a settable static.
#include <cstdint>
#include <climits>
#include <cstring>
#include <random>
#include <fstream>
#include <string>

#define private public            // harness needs Decoder internals + J (test-only)
#include "KanekoKernelProcessor.h"
#undef private
#include "bchCoder.h"
#include "dataForPlot.h"

extern std::default_random_engine generator;   // src/bchCoder.cpp:20

#ifdef REF_CAP_VARIANT
long KanekoKernelProcessor::J = 15;            // patched copy: `static long J;`
#endif

namespace {
const unsigned long kPrimPoly[16] = {3, 7, 11, 19, 37, 67, 137, 285, 529, 1033,
                                     2053, 4179, 8219, 17475, 32771, 69643};
struct RefCode {
    int m, n, k, t, gSize;
    unsigned long *antilog, *log;
    unsigned char *g;
    KanekoKernelProcessor *kan;
    Decoder *dec;
};
}  // namespace

extern "C" {

// Mirrors src/main.cpp:59-95 (table + generator construction) by CALLING the
// reference's findMinimalPolynomial / lcm.
void *ref_create(int m, int t) {
    RefCode *c = new RefCode;
    c->m = m; c->t = t;
    int n = (1 << m) - 1;
    c->n = n;
    unsigned long prim = kPrimPoly[m - 1];
    c->antilog = new unsigned long[n];
    c->log = new unsigned long[n + 1];
    c->antilog[0] = 1; c->antilog[1] = 2;
    c->log[0] = LONG_MAX; c->log[1] = 0; c->log[2] = 1;
    for (unsigned long i = 2; i < (unsigned long)n; ++i) {
        c->antilog[i] = c->antilog[i - 1] << 1;
        if (c->antilog[i] >> m == 1) c->antilog[i] ^= prim;
        c->log[c->antilog[i]] = i;
    }
    int gSize, mpSize;
    unsigned char *g = new unsigned char[n];
    findMinimalPolynomial(1, m, c->antilog, &gSize, g);
    unsigned char *mp = new unsigned char[m + 1];
    for (unsigned long i = 2; i < 2UL * t; ++i) {
        findMinimalPolynomial(i, m, c->antilog, &mpSize, mp);
        unsigned char *tmp = lcm(g, gSize, mp, mpSize, &gSize);
        delete[] g;
        g = tmp;
    }
    delete[] mp;
    c->g = g; c->gSize = gSize; c->k = n - gSize + 1;
    c->kan = new KanekoKernelProcessor(m, n, t, c->k, c->antilog, c->log, 0.5);  // main.cpp:176
    c->dec = new Decoder(m, n, t, c->k, c->antilog, c->log);
    return c;
}

void ref_info(void *h, int *n, int *k, int *t, int *gSize, unsigned char *g_out) {
    RefCode *c = (RefCode *)h;
    *n = c->n; *k = c->k; *t = c->t; *gSize = c->gSize;
    if (g_out) memcpy(g_out, c->g, c->gSize);
}

void ref_tables(void *h, unsigned long *antilog_out, unsigned long *log_out) {
    RefCode *c = (RefCode *)h;
    memcpy(antilog_out, c->antilog, sizeof(unsigned long) * c->n);
    memcpy(log_out, c->log, sizeof(unsigned long) * (c->n + 1));
}

void ref_seed(unsigned long s) { generator.seed(s); }

int ref_set_J(long J) {
#ifdef REF_CAP_VARIANT
    KanekoKernelProcessor::J = J;
    return 0;
#else
    (void)J;
    return -1;   // HEAD build has no cap
#endif
}

// One frame exactly as src/dataForPlot.cpp:45-48 generates it.
void ref_gen_frames(void *h, double ebn0_db, long B, unsigned char *info, unsigned char *cw, double *y) {
    RefCode *c = (RefCode *)h;
    for (long f = 0; f < B; ++f) {
        double sd = sqrt(1 / (pow(10, ebn0_db / 10) * 2 * c->k / c->n));
        generateRandomPoly(info + f * c->k, c->k);
        multiplyPolynomials(info + f * c->k, c->k, c->g, c->gSize, cw + f * c->n);
        addNoise(sd, cw + f * c->n, y + f * c->n, c->n);
    }
}

void ref_encode(void *h, const unsigned char *info, long B, unsigned char *cw) {
    RefCode *c = (RefCode *)h;
    for (long f = 0; f < B; ++f)
        multiplyPolynomials(info + f * c->k, c->k, c->g, c->gSize, cw + f * c->n);
}

// 3-argument decode (the one fun() uses, src/KanekoKernelProcessor.cpp:335).
// `decided` rows are only written on improvement, like the reference.
void ref_kaneko_decode(void *h, const double *y, const unsigned char *answer, long B,
                       unsigned char *decided, uint32_t *trials, uint64_t *cmp, uint64_t *sum) {
    RefCode *c = (RefCode *)h;
    for (long f = 0; f < B; ++f) {
        c->kan->setDecodingCount(); c->kan->setComparisonCount(); c->kan->setSummCount();
        c->kan->decode(answer + f * c->n, y + f * c->n, decided + f * c->n);
        if (trials) trials[f] = (uint32_t)c->kan->getDecodingCount();
        if (cmp) cmp[f] = c->kan->getComparisonCount();
        if (sum) sum[f] = c->kan->getSummCount();
    }
}

// 2-argument decode (file mode, src/KanekoKernelProcessor.cpp:212).
void ref_kaneko_decode2(void *h, const double *y, long B, unsigned char *decided,
                        uint32_t *trials, uint64_t *cmp, uint64_t *sum) {
    RefCode *c = (RefCode *)h;
    for (long f = 0; f < B; ++f) {
        c->kan->setDecodingCount(); c->kan->setComparisonCount(); c->kan->setSummCount();
        c->kan->decode(y + f * c->n, decided + f * c->n);
        if (trials) trials[f] = (uint32_t)c->kan->getDecodingCount();
        if (cmp) cmp[f] = c->kan->getComparisonCount();
        if (sum) sum[f] = c->kan->getSummCount();
    }
}

// Algebraic decoder alone: findSyndromPoly + decode (src/Decoder.cpp:184,298).
// Also returns the 2t syndromes and the locator polynomial the reference built.
void ref_bdd(void *h, const unsigned char *words, long B, unsigned char *answers, unsigned char *ok,
             unsigned long *synd /*[B][2t] or null*/, unsigned long *lambda /*[B][t+1] or null*/,
             int *lambda_size /*[B] or null*/) {
    RefCode *c = (RefCode *)h;
    for (long f = 0; f < B; ++f) {
        c->dec->findSyndromPoly(words + f * c->n);
        if (synd) memcpy(synd + f * 2 * c->t, c->dec->syndromPoly, sizeof(unsigned long) * 2 * c->t);
        bool r = c->dec->decode(words + f * c->n, answers + f * c->n);
        ok[f] = r ? 1 : 0;
        if (lambda_size) lambda_size[f] = c->dec->sizeLambda;
        if (lambda) memcpy(lambda + f * (c->t + 1), c->dec->lambda, sizeof(unsigned long) * (c->t + 1));
    }
}

// The reference's own Monte-Carlo sweep (src/dataForPlot.cpp:16): writes <file>.csv.
void ref_fun(void *h, const char *file, long p, long e, double maxSTNR) {
    RefCode *c = (RefCode *)h;
    c->kan->setDecodingCount(); c->kan->setComparisonCount(); c->kan->setSummCount();
    fun(std::string(file), *c->kan, c->g, (unsigned long)c->gSize, p, e, maxSTNR);
}

// n x n nested-BCH kernel matrix (src/bchCoder.cpp:317), row-major bytes.
void ref_make_matrix(void *h, unsigned char *out) {
    RefCode *c = (RefCode *)h;
    unsigned char **mtx = new unsigned char *[c->n];
    for (int i = 0; i < c->n; ++i) mtx[i] = new unsigned char[c->n];
    makeMatrix(c->m, c->antilog, mtx);
    for (int i = 0; i < c->n; ++i) { memcpy(out + (long)i * c->n, mtx[i], c->n); delete[] mtx[i]; }
    delete[] mtx;
}

}  // extern "C"
