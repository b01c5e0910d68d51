#include "bch_decoder.hpp"

#include <cstring>

#include "pk_capi.h"

Decoder::Decoder(long pw, long n, long t, long k, unsigned long *antilogarithms, unsigned long *logarithms)
    : syndromPoly(new unsigned long[2 * t]()), syndromPolySize(0), power_(pw), n_(n), t_(t), k_(k),
      antilog_(antilogarithms), log_(logarithms), code_(nullptr), last_word_(new unsigned char[n]()) {
    if (pk_code_create((int)pw, (int)t, 0, &code_) != PK_OK) throw pk_last_error();
    int nn = 0, kk = 0;
    pk_code_info(code_, &nn, &kk, nullptr, nullptr, nullptr);
    if (nn != n || kk != k) throw "Invalid values of arguments\n";
}

Decoder::~Decoder() {
    pk_code_destroy(code_);
    delete[] syndromPoly;
    delete[] last_word_;
}

// S_j = word(alpha^j) = XOR over set positions i of alpha^{(i*j) mod n}
void Decoder::findSyndromPoly(const unsigned char *word) {
    syndromPolySize = 0;
    for (long j = 1; j <= 2 * t_; ++j) {
        unsigned long s = 0;
        for (long i = 0; i < n_; ++i)
            if (word[i]) s ^= antilog_[(i * j) % n_];
        syndromPoly[j - 1] = s;
        if (s) syndromPolySize = j;
    }
    std::memcpy(last_word_, word, (size_t)n_);
}

void Decoder::alterSyndromPoly(const unsigned char *word) {
    for (long i = 0; i < n_; ++i)
        if (last_word_[i] != word[i])
            for (long j = 1; j <= 2 * t_; ++j) syndromPoly[j - 1] ^= antilog_[(i * j) % n_];
    syndromPolySize = 0;
    for (long j = 2 * t_; j >= 1; --j)
        if (syndromPoly[j - 1]) { syndromPolySize = j; break; }
    std::memcpy(last_word_, word, (size_t)n_);
}

bool Decoder::decode(const unsigned char *word, unsigned char *answer) {
    uint8_t ok = 0;
    if (pk_bch_decode_batch(code_, word, 1, answer, &ok) != PK_OK) throw pk_last_error();
    return ok != 0;
}
