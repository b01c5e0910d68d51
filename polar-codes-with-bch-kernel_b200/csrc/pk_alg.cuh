// pk_alg.cuh -- bounded-distance algebraic BCH decoding from packed syndromes.
//
// Replaces Decoder::euclid + Decoder::locatorsAndRoots (reference src/Decoder.cpp:233-296).
// The reference runs Sugiyama's Euclid on (x^2t, S(x)) until deg r < t, rejects
// Lambda(0) == 0 (:270-273), then Chien-searches and accepts iff the number of distinct
// roots equals deg Lambda >= 1 (:279-296, the `count == size - 2` test; a zero syndrome
// gives Lambda = 1 and is REJECTED).  By the uniqueness of minimal key-equation solutions
// that is: Berlekamp-Massey's final LFSR length L <= t, with the same Lambda up to a
// scalar.  Note deg Lambda < L is accepted by the reference (no deg check), so we must not
// add one.  We therefore run an inversion-free binary BM (even-step discrepancies vanish
// because S_2j = S_j^2 for syndromes of a binary word) in lock-step registers, and a
// Horner Chien search whose multiplier row is warp-uniform (bank-conflict free).
//
// Exhaustive (n-k <= 15) and randomised differential tests against the compiled
// reference are in tests/test_gpu_parity.py (test_bch_decode_matches_oracle, test_bch_decode_exhaustive_cosets) and tests/test_capi_host.py.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PK_HD __host__ __device__ __forceinline__
#else
#define PK_HD inline
#endif

template <int M, int T>
struct PkCfg {
    static constexpr int N = (1 << M) - 1;       // code length
    static constexpr int NW = (N + 31) / 32;     // 32-bit words per position mask
    static constexpr int PER = 32 / M;           // syndromes per packed word (no straddling)
    static constexpr int NSYN = 2 * T;
    static constexpr int NSW = (NSYN + PER - 1) / PER;  // packed syndrome words
    static constexpr int NA = NSW + NW;          // "augmented column": syndromes | one-hot position
};

// Sw   : NSW packed words holding S_1..S_2t (S_j = r(alpha^j)), M bits each.
// mul  : full GF(2^M) product table, mul[(a << M) | b].
// xoff : xoff[p] = alpha^{-p} << M  (row of `mul` that multiplies by the Chien point of position p).
// A    : out, NW words: bit p set  <=>  Lambda(alpha^{-p}) == 0.
// returns the reference's Decoder::decode() verdict.
template <int M, int T>
PK_HD bool pk_alg_decode(const uint32_t *Sw, const uint8_t *mul, const uint16_t *xoff, uint32_t *A) {
    typedef PkCfg<M, T> C;
    uint32_t S[2 * T + 1];
#pragma unroll
    for (int j = 1; j <= 2 * T; ++j) S[j] = (Sw[(j - 1) / C::PER] >> (((j - 1) % C::PER) * M)) & C::N;

    // ---- inversion-free Berlekamp-Massey:  Lambda <- gamma*Lambda + delta*x^s*B
    uint32_t Lam[T + 1], Bp[T + 1];
#pragma unroll
    for (int k = 0; k <= T; ++k) { Lam[k] = 0; Bp[k] = 0; }
    Lam[0] = 1;
    Bp[0] = 1;
    int L = 0;
    uint32_t gamma = 1;
#pragma unroll
    for (int r = 1; r <= 2 * T; ++r) {
#pragma unroll
        for (int k = T; k >= 1; --k) Bp[k] = Bp[k - 1];   // B <- x*B (pure renaming once unrolled)
        Bp[0] = 0;
        if (r & 1) {
            const int dl = (r - 1 < T) ? r - 1 : T;        // deg Lambda <= min(T, r-1) here
            const int du = (r < T) ? r : T;                // deg of shifted B and of new Lambda <= min(T, r)
            uint32_t delta = 0;
#pragma unroll
            for (int k = 0; k <= dl; ++k) delta ^= mul[(Lam[k] << M) | S[r - k]];
            const uint8_t *rg = mul + (gamma << M);
            const uint8_t *rd = mul + (delta << M);
            const bool upd = (delta != 0) && (2 * L <= r - 1);
#pragma unroll
            for (int k = 0; k <= du; ++k) {
                uint32_t nl = (uint32_t)rg[Lam[k]] ^ (uint32_t)rd[Bp[k]];
                Bp[k] = upd ? Lam[k] : Bp[k];
                Lam[k] = nl;
            }
            L = upd ? r - L : L;
            gamma = upd ? delta : gamma;
        }
    }
    int d = 0;
#pragma unroll
    for (int k = 1; k <= T; ++k) d = Lam[k] ? k : d;

    // ---- Chien search, Horner form; position p <-> root alpha^{(n-p) % n} (Decoder.cpp:287)
    int nroots = 0;
#pragma unroll
    for (int w = 0; w < C::NW; ++w) A[w] = 0;
    if constexpr (C::NW <= 2) {
        // n <= 63: everything unrolled, A stays in registers
#pragma unroll
        for (int p = 0; p < C::N; ++p) {
            const uint8_t *xr = mul + xoff[p];
            uint32_t v = Lam[T];
#pragma unroll
            for (int j = T - 1; j >= 0; --j) v = (uint32_t)xr[v] ^ Lam[j];
            const bool z = (v == 0);
            nroots += z ? 1 : 0;
            A[p >> 5] |= z ? (1u << (p & 31)) : 0u;
        }
    } else {
        // n = 127 / 255: runtime loop over words (code size), 32 positions unrolled inside
#pragma unroll 1
        for (int w = 0; w < C::NW; ++w) {
            uint32_t aw = 0;
            const int pend = (w == C::NW - 1) ? (C::N - 32 * (C::NW - 1)) : 32;
#pragma unroll 8
            for (int b = 0; b < pend; ++b) {
                const uint8_t *xr = mul + xoff[w * 32 + b];
                uint32_t v = Lam[T];
#pragma unroll
                for (int j = T - 1; j >= 0; --j) v = (uint32_t)xr[v] ^ Lam[j];
                const bool z = (v == 0);
                nroots += z ? 1 : 0;
                aw |= z ? (1u << b) : 0u;
            }
#pragma unroll
            for (int ww = 0; ww < C::NW; ++ww)
                if (ww == w) A[ww] = aw;
        }
    }
    return (L <= T) && (d >= 1) && (Lam[0] != 0) && (nroots == d);
}
