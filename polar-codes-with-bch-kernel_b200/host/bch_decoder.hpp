// bch_decoder.hpp -- class Decoder with the reference's public interface
// (headers/Decoder.h:67-78), backed by the sm_100a algebraic decoder through the C ABI.
#pragma once
#include <cstdint>

struct pk_code;

class Decoder {
public:
    // same argument list as the reference; the two tables are BORROWED (never freed), and
    // are only used to fill the public syndrome fields -- the device owns its own copies.
    Decoder(long pw, long n, long t, long k, unsigned long *antilogarithms, unsigned long *logarithms);
    ~Decoder();
    Decoder(const Decoder &) = delete;
    Decoder &operator=(const Decoder &) = delete;

    void findSyndromPoly(const unsigned char *word);    // Decoder.cpp:184
    void alterSyndromPoly(const unsigned char *word);   // Decoder.cpp:210
    // t-error bounded-distance decode of `word` on the GPU; `answer` untouched on failure
    // (Decoder.cpp:298-321).  A zero syndrome reports failure, exactly like the reference.
    bool decode(const unsigned char *word, unsigned char *answer);
    long getN() const { return n_; }
    long getT() const { return t_; }
    long getK() const { return k_; }
    pk_code *handle() const { return code_; }

    unsigned long *syndromPoly;   // S_1 .. S_2t of the last word given to find/alterSyndromPoly
    long syndromPolySize;         // index (1-based) of the highest non-zero syndrome, 0 if none

private:
    long power_, n_, t_, k_;
    unsigned long *antilog_, *log_;
    pk_code *code_;
    unsigned char *last_word_;
};
