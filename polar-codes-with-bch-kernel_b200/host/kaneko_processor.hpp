// kaneko_processor.hpp -- class KanekoKernelProcessor with the reference's public interface
// (headers/KanekoKernelProcessor.h:47-68) on top of the C ABI.  Besides the per-word
// decode() of the reference it exposes the batched entry points the Monte-Carlo driver uses.
#pragma once
#include <cstdint>

#include "bch_decoder.hpp"
#include "pk_capi.h"

class KanekoKernelProcessor {
public:
    // J = cap on the test-pattern exponent (reference: compile-time constant 15 that HEAD ignores,
    // KanekoKernelProcessor.h:35, .cpp:392-393); J < 0 keeps HEAD behaviour.
    KanekoKernelProcessor(long pw, long n, long t, long k, unsigned long *antilogarithms, unsigned long *logarithms,
                          double signalToNoiseRatio, long J = -1);
    virtual ~KanekoKernelProcessor();
    KanekoKernelProcessor(const KanekoKernelProcessor &) = delete;
    KanekoKernelProcessor &operator=(const KanekoKernelProcessor &) = delete;

    // decode(answer, word, res): the flavour fun() uses (KanekoKernelProcessor.cpp:335-407).
    void decode(const unsigned char *answer, const double *word, unsigned char *res);
    // decode(word, res): the file-mode flavour (:212-276), also on the device (pk_kaneko_set_variant).
    // decode(res): the DEBUG flavour (:161-210) reads an uninitialised flag in the reference and is not
    // provided -- it throws instead of silently answering with other semantics.
    void decode(const double *word, unsigned char *res);
    void decode(unsigned char *res);
    void set(const double *word) const;
    double calcL(const unsigned char *word) const;   // :79-87, against the last word given to set()/decode()

    long getN() const { return decoder.getN(); }
    long getT() const { return decoder.getT(); }
    long getK() const { return decoder.getK(); }
    unsigned long getComparisonCount() const { return comparisonCount; }
    unsigned long getSummCount() const { return summCount; }
    unsigned long getDecodingCount() const { return decodingCount; }
    void setDecodingCount(unsigned long c = 0) { decodingCount = c; }
    void setComparisonCount(unsigned long c = 0) { comparisonCount = c; }
    void setSummCount(unsigned long c = 0) { summCount = c; }

    // ---- batched extensions (what the sm_100a path is for)
    // replay: B frames of channel output -> decisions; counters accumulate like B decode() calls
    void decodeBatch(const double *words, long B, unsigned char *res, uint32_t *trials = nullptr);
    // generation mode: one SNR point with fun()'s stop rule, frames drawn on the device
    pk_point_result runPoint(double ebn0_db, int snr_index, uint64_t seed, long p, long e);
    // shard the frames of runPoint over the first `ngpus` devices of the box (pk_comm_create / pk_comm_run_point:
    // one NCCL all-reduce of the counters per SNR point); the results do not depend on ngpus
    void useGpus(int ngpus);
    int gpus() const { return ngpus_; }
    pk_kaneko *handle() const { return kan_; }

private:
    Decoder decoder;
    long n_, t_, pw_, J_;
    double sd, snr_;
    pk_kaneko *kan_;
    int ngpus_{1};
    pk_comm *comm_{nullptr};
    pk_comm_kaneko *ckan_{nullptr};
    mutable double *alpha_;
    mutable unsigned char *yH_;
    unsigned long comparisonCount{0}, summCount{0}, decodingCount{0};
    void account(const pk_point_result &r);
};
