#include "kaneko_processor.hpp"

#include <cmath>
#include <cstring>

KanekoKernelProcessor::KanekoKernelProcessor(long pw, long n, long t, long k, unsigned long *antilogarithms,
                                             unsigned long *logarithms, double signalToNoiseRatio, long J)
    : decoder(pw, n, t, k, antilogarithms, logarithms), n_(n), t_(t), pw_(pw), J_(J), snr_(signalToNoiseRatio), kan_(nullptr), alpha_(new double[n]()),
      yH_(new unsigned char[n]()) {
    sd = sqrt(1 / (pow(10, signalToNoiseRatio / 10) * 2 * k / n));   // KanekoKernelProcessor.cpp:20
    if (pk_kaneko_create(decoder.handle(), signalToNoiseRatio, J, 0, &kan_) != PK_OK) throw pk_last_error();
}

void KanekoKernelProcessor::useGpus(int ngpus) {
    if (ngpus < 1) throw "Invalid values of arguments\n";
    if (ckan_) { pk_comm_kaneko_destroy(ckan_); ckan_ = nullptr; }
    if (comm_) { pk_comm_destroy(comm_); comm_ = nullptr; }
    ngpus_ = 1;
    if (ngpus == 1) return;
    if (pk_comm_create(ngpus, nullptr, &comm_) != PK_OK) throw pk_last_error();
    if (pk_comm_kaneko_create(comm_, (int)pw_, (int)t_, snr_, J_, 0, &ckan_) != PK_OK) throw pk_last_error();
    ngpus_ = ngpus;
}

KanekoKernelProcessor::~KanekoKernelProcessor() {
    if (ckan_) pk_comm_kaneko_destroy(ckan_);
    if (comm_) pk_comm_destroy(comm_);
    pk_kaneko_destroy(kan_);
    delete[] alpha_;
    delete[] yH_;
}

void KanekoKernelProcessor::account(const pk_point_result &r) {
    decodingCount += r.trials;
    comparisonCount += r.cmp;
    summCount += r.sum;
}

void KanekoKernelProcessor::set(const double *word) const {
    for (long i = 0; i < n_; ++i) {
        const double a = 2 * word[i] / pow(sd, 2);
        yH_[i] = (a <= 0.0) ? 0 : 1;
        alpha_[i] = fabs(a);
    }
}

double KanekoKernelProcessor::calcL(const unsigned char *word) const {
    double l = 0;
    for (long i = 0; i < n_; ++i)
        if (yH_[i] != word[i]) l += alpha_[i];
    return l;
}

void KanekoKernelProcessor::decode(const unsigned char *answer, const double *word, unsigned char *res) {
    (void)answer;   // only feeds an unused local in the reference (:345)
    set(word);
    decodeBatch(word, 1, res);
}

void KanekoKernelProcessor::decode(const double *word, unsigned char *res) {
    // file-mode flavour (:212-276): same kernels, different loop bound and counter bookkeeping
    set(word);
    if (pk_kaneko_set_variant(kan_, 1) != PK_OK) throw pk_last_error();
    try {
        decodeBatch(word, 1, res);
    } catch (...) {
        pk_kaneko_set_variant(kan_, 0);
        throw;
    }
    pk_kaneko_set_variant(kan_, 0);
}
void KanekoKernelProcessor::decode(unsigned char *) {
    throw "KanekoKernelProcessor::decode(res): the DEBUG flavour reads an uninitialised flag in the reference and is not provided\n";
}

void KanekoKernelProcessor::decodeBatch(const double *words, long B, unsigned char *res, uint32_t *trials) {
    // the reference leaves `res` untouched when no trial succeeds; keep a copy to restore such rows
    pk_point_result tot;
    unsigned char *keep = new unsigned char[(size_t)B * n_];
    std::memcpy(keep, res, (size_t)B * n_);
    pk_frame_rec *recs = new pk_frame_rec[B];
    const int rc = pk_kaneko_decode_batch(kan_, words, B, res, trials, recs, &tot);
    if (rc == PK_OK && (tot.flags_or & PK_FLAG_NO_DECISION))
        for (long f = 0; f < B; ++f)
            if (recs[f].flags & PK_FLAG_NO_DECISION) std::memcpy(res + f * n_, keep + f * n_, (size_t)n_);
    delete[] keep;
    delete[] recs;
    if (rc != PK_OK) throw pk_last_error();
    account(tot);
}

pk_point_result KanekoKernelProcessor::runPoint(double ebn0_db, int snr_index, uint64_t seed, long p, long e) {
    pk_point_result r;
    const int rc = ckan_ ? pk_comm_run_point(ckan_, ebn0_db, snr_index, seed, p, e, &r)
                         : pk_kaneko_run_point(kan_, ebn0_db, snr_index, seed, p, e, &r);
    if (rc != PK_OK) throw pk_last_error();
    account(r);
    return r;
}
