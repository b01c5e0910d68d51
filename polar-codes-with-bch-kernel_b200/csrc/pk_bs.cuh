// pk_bs.cuh -- bit-sliced bounded-distance BCH decoding: 32 test patterns per 32-bit word.
//
// Same verdict as pk_alg_decode<M,T> (i.e. as the reference's Decoder::decode,
// src/Decoder.cpp:233-321) for each of the 32 trials packed in the bit lanes of a word:
// every GF(2^M) element is M bit-planes, every operation is a handful of LOP3s and there
// are no table lookups at all -- the long searches of Kaneko decoding (thousands of trials per
// frame, >99 % of them failures for the t >= 4 codes) run at ALU speed instead of shared-memory
// latency.  Multiplication by the constants alpha^e (Chien stepping) is an XOR network fixed
// at compile time.  Host+device so the CPU tests can compare it with pk_alg_decode.
#pragma once
#include <stdint.h>

#include <utility>

#include "pk_alg.cuh"

template <int M>
struct PkGF {
    static constexpr int N = (1 << M) - 1;
    // primitive polynomials of the reference (src/main.cpp:14-15)
    static constexpr uint32_t PRIM = (M == 3) ? 11u : (M == 4) ? 19u : (M == 5) ? 37u : (M == 6) ? 67u : (M == 7) ? 137u : 285u;
    static constexpr uint32_t PLOW = PRIM ^ (1u << M);
    static constexpr uint32_t apow(int e) {
        uint32_t v = 1;
        e %= N;
        for (int i = 0; i < e; ++i) {
            v <<= 1;
            if (v >> M) v ^= PRIM;
        }
        return v;
    }
};
// bit B of alpha^(E+I): coefficient of input plane I in output plane B of "multiply by alpha^E"
template <int M, int E, int I, int B>
struct PkCBit {
    static constexpr bool v = ((PkGF<M>::apow(E + I) >> B) & 1u) != 0;
};

template <int M, int E, int B, int... I>
PK_HD uint32_t pk_bs_cmul_row(const uint32_t *v, std::integer_sequence<int, I...>) {
    return (0u ^ ... ^ (PkCBit<M, E, I, B>::v ? v[I] : 0u));
}
template <int M, int E, int... B>
PK_HD void pk_bs_cmul_all(uint32_t *v, std::integer_sequence<int, B...>) {
    const uint32_t o[M] = {pk_bs_cmul_row<M, E, B>(v, std::make_integer_sequence<int, M>{})...};
#pragma unroll
    for (int b = 0; b < M; ++b) v[b] = o[b];
}
// v <- v * alpha^E (E compile-time)
template <int M, int E>
PK_HD void pk_bs_cmul(uint32_t *v) {
    pk_bs_cmul_all<M, E>(v, std::make_integer_sequence<int, M>{});
}

// acc (2M-1 planes, unreduced) ^= a * b
template <int M>
PK_HD void pk_bs_mac(uint32_t *acc, const uint32_t *a, const uint32_t *b) {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) acc[i + j] ^= a[i] & b[j];
}
// fold planes M..2M-2 back (x^M = PLOW)
template <int M>
PK_HD void pk_bs_reduce(uint32_t *acc) {
#pragma unroll
    for (int k = 2 * M - 2; k >= M; --k)
#pragma unroll
        for (int e = 0; e < M; ++e)
            if ((PkGF<M>::PLOW >> e) & 1u) acc[k - M + e] ^= acc[k];
}

// Chien stepping constants: term_j <- term_j * alpha^{-j}, one position at a time.
template <int M, int T, int J>
struct PkBsChienStep {
    PK_HD static void run(uint32_t (*term)[M]) {
        pk_bs_cmul<M, (PkGF<M>::N - (J % PkGF<M>::N)) % PkGF<M>::N>(term[J]);
        PkBsChienStep<M, T, J + 1>::run(term);
    }
};
template <int M, int T>
struct PkBsChienStep<M, T, T + 1> {
    PK_HD static void run(uint32_t (*)[M]) {}
};

// SFn: functor  void operator()(int j /*1..2T*/, uint32_t *planes /*[M]*/)  giving syndrome S_j.
// Z  : out, Z[p * zstride] = word whose bit q is set iff Lambda_q(alpha^{-p}) == 0 (position p is
//      located for trial q); written for every p in [0, N).
// returns the word of verdicts (bit q = Decoder::decode() of trial q).
// LOOP = true: the t Berlekamp-Massey steps run as ONE loop body (fits the 32 KB instruction cache; products
// known to be zero are skipped by warp-uniform branches); LOOP = false: fully unrolled (more constant folding).
template <int M, int T, bool LOOP, class SFn>
PK_HD uint32_t pk_bs_decode(SFn getS, uint32_t *Z, int zstride) {
    constexpr int N = PkGF<M>::N;
    constexpr int LB = (2 * T < 2) ? 1 : (2 * T < 4) ? 2 : (2 * T < 8) ? 3 : (2 * T < 16) ? 4 : 5;   // bits of L <= 2T
    uint32_t Lam[T + 1][M], Bp[T + 1][M], gamma[M], Lb[LB];
#pragma unroll
    for (int k = 0; k <= T; ++k)
#pragma unroll
        for (int b = 0; b < M; ++b) { Lam[k][b] = 0; Bp[k][b] = 0; }
    Lam[0][0] = ~0u;
    Bp[0][0] = ~0u;
#pragma unroll
    for (int b = 0; b < M; ++b) gamma[b] = 0;
    gamma[0] = ~0u;
#pragma unroll
    for (int b = 0; b < LB; ++b) Lb[b] = 0;

    if constexpr (LOOP) {
#pragma unroll 1
        for (int it = 0; it < T; ++it) {
            const int r = 2 * it + 1;
            // B <- x*B for step r
#pragma unroll
            for (int k = T; k >= 1; --k)
#pragma unroll
                for (int b = 0; b < M; ++b) Bp[k][b] = Bp[k - 1][b];
#pragma unroll
            for (int b = 0; b < M; ++b) Bp[0][b] = 0;
            const int dl = (r - 1 < T) ? r - 1 : T;   // deg Lambda <= dl, deg x^s B and new Lambda <= du
            const int du = (r < T) ? r : T;
            uint32_t acc[2 * M - 1], delta[M];
#pragma unroll
            for (int b = 0; b < 2 * M - 1; ++b) acc[b] = 0;
#pragma unroll
            for (int k = 0; k <= T; ++k) {
                if (k <= dl) {   // warp-uniform: skips the products known to be zero
                    uint32_t sj[M];
                    getS(r - k, sj);
                    pk_bs_mac<M>(acc, Lam[k], sj);
                }
            }
            pk_bs_reduce<M>(acc);
            uint32_t nz = 0;
#pragma unroll
            for (int b = 0; b < M; ++b) { delta[b] = acc[b]; nz |= acc[b]; }
            // 2L <= r-1  <=>  L <= it
            uint32_t lt = 0, eq = ~0u;
#pragma unroll
            for (int b = LB - 1; b >= 0; --b) {
                const uint32_t cb = 0u - (uint32_t)((it >> b) & 1);
                lt |= eq & ~Lb[b] & cb;
                eq &= ~(Lb[b] ^ cb);
            }
            const uint32_t upd = nz & (lt | eq);
#pragma unroll
            for (int k = 0; k <= T; ++k) {
                if (k <= du) {
#pragma unroll
                    for (int b = 0; b < 2 * M - 1; ++b) acc[b] = 0;
                    pk_bs_mac<M>(acc, gamma, Lam[k]);
                    pk_bs_mac<M>(acc, delta, Bp[k]);
                    pk_bs_reduce<M>(acc);
#pragma unroll
                    for (int b = 0; b < M; ++b) {
                        Bp[k][b] = (Bp[k][b] & ~upd) | (Lam[k][b] & upd);
                        Lam[k][b] = acc[b];
                    }
                }
            }
            uint32_t carry = ~0u;   // L <- upd ? r - L : L
#pragma unroll
            for (int b = 0; b < LB; ++b) {
                const uint32_t rb = 0u - (uint32_t)((r >> b) & 1);
                const uint32_t nl = ~Lb[b];
                const uint32_t sum = rb ^ nl ^ carry;
                carry = (rb & nl) | (carry & (rb ^ nl));
                Lb[b] = (Lb[b] & ~upd) | (sum & upd);
            }
#pragma unroll
            for (int b = 0; b < M; ++b) gamma[b] = (gamma[b] & ~upd) | (delta[b] & upd);
            // B <- x*B for the even step r+1 (its discrepancy is zero for a binary code)
#pragma unroll
            for (int k = T; k >= 1; --k)
#pragma unroll
                for (int b = 0; b < M; ++b) Bp[k][b] = Bp[k - 1][b];
#pragma unroll
            for (int b = 0; b < M; ++b) Bp[0][b] = 0;
        }
    } else {
#pragma unroll
    for (int r = 1; r <= 2 * T; ++r) {
#pragma unroll
        for (int k = T; k >= 1; --k)
#pragma unroll
            for (int b = 0; b < M; ++b) Bp[k][b] = Bp[k - 1][b];
#pragma unroll
        for (int b = 0; b < M; ++b) Bp[0][b] = 0;
        if (r & 1) {
            const int dl = (r - 1 < T) ? r - 1 : T;
            const int du = (r < T) ? r : T;
            uint32_t acc[2 * M - 1], delta[M];
#pragma unroll
            for (int b = 0; b < 2 * M - 1; ++b) acc[b] = 0;
#pragma unroll
            for (int k = 0; k <= dl; ++k) {
                uint32_t s[M];
                getS(r - k, s);
                pk_bs_mac<M>(acc, Lam[k], s);
            }
            pk_bs_reduce<M>(acc);
            uint32_t nz = 0;
#pragma unroll
            for (int b = 0; b < M; ++b) { delta[b] = acc[b]; nz |= acc[b]; }
            // 2L <= r-1  <=>  L <= (r-1)/2
            const int c = (r - 1) / 2;
            uint32_t lt = 0, eq = ~0u;
#pragma unroll
            for (int b = LB - 1; b >= 0; --b) {
                if ((c >> b) & 1) { lt |= eq & ~Lb[b]; eq &= Lb[b]; }
                else eq &= ~Lb[b];
            }
            const uint32_t upd = nz & (lt | eq);
#pragma unroll
            for (int k = 0; k <= du; ++k) {
#pragma unroll
                for (int b = 0; b < 2 * M - 1; ++b) acc[b] = 0;
                pk_bs_mac<M>(acc, gamma, Lam[k]);
                pk_bs_mac<M>(acc, delta, Bp[k]);
                pk_bs_reduce<M>(acc);
#pragma unroll
                for (int b = 0; b < M; ++b) {
                    Bp[k][b] = (Bp[k][b] & ~upd) | (Lam[k][b] & upd);
                    Lam[k][b] = acc[b];
                }
            }
            // L <- upd ? r - L : L   (r + ~L + 1, ripple over LB bits)
            uint32_t carry = ~0u;
#pragma unroll
            for (int b = 0; b < LB; ++b) {
                const uint32_t rb = ((r >> b) & 1) ? ~0u : 0u;
                const uint32_t nl = ~Lb[b];
                const uint32_t sum = rb ^ nl ^ carry;
                carry = (rb & nl) | (carry & (rb ^ nl));
                Lb[b] = (Lb[b] & ~upd) | (sum & upd);
            }
#pragma unroll
            for (int b = 0; b < M; ++b) gamma[b] = (gamma[b] & ~upd) | (delta[b] & upd);
        }
    }
    }
    // L <= T
    uint32_t okL;
    {
        uint32_t lt = 0, eq = ~0u;
#pragma unroll
        for (int b = LB - 1; b >= 0; --b) {
            if ((T >> b) & 1) { lt |= eq & ~Lb[b]; eq &= Lb[b]; }
            else eq &= ~Lb[b];
        }
        okL = lt | eq;
    }
    // d = deg Lambda, as LB bit-planes
    uint32_t dB[LB], dge1 = 0, lam0 = 0;
#pragma unroll
    for (int b = 0; b < LB; ++b) dB[b] = 0;
#pragma unroll
    for (int b = 0; b < M; ++b) lam0 |= Lam[0][b];
#pragma unroll
    for (int k = 1; k <= T; ++k) {
        uint32_t nzk = 0;
#pragma unroll
        for (int b = 0; b < M; ++b) nzk |= Lam[k][b];
        dge1 |= nzk;
#pragma unroll
        for (int b = 0; b < LB; ++b) dB[b] = (dB[b] & ~nzk) | (((k >> b) & 1) ? nzk : 0u);
    }
    // Chien: term_j(p) = Lambda_j * alpha^{-jp}; root count as LB bit-planes
    uint32_t cnt[LB];
#pragma unroll
    for (int b = 0; b < LB; ++b) cnt[b] = 0;
#pragma unroll 1
    for (int p = 0; p < N; ++p) {
        uint32_t nzv = 0;
#pragma unroll
        for (int b = 0; b < M; ++b) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k <= T; ++k) v ^= Lam[k][b];
            nzv |= v;
        }
        const uint32_t z = ~nzv;
        Z[p * zstride] = z;
        uint32_t carry = z;
#pragma unroll
        for (int b = 0; b < LB; ++b) {
            const uint32_t t2 = cnt[b] & carry;
            cnt[b] ^= carry;
            carry = t2;
        }
        PkBsChienStep<M, T, 1>::run(Lam);
    }
    uint32_t same = ~0u;
#pragma unroll
    for (int b = 0; b < LB; ++b) same &= ~(cnt[b] ^ dB[b]);
    return okL & dge1 & lam0 & same;
}

// ------------------------------------------------------------------ large codes: BM state in (shared) memory
// Same algorithm as pk_bs_decode<.., LOOP = true>, for codes whose Lambda/B planes do not fit the register file
// ((t+1) m > 56, e.g. (127,64,21): 77 planes each, (255,139,31): 128 planes each).  Lambda and B live in a
// lane-private column of `st` (element (k, b) of Lambda at st[(k*M + b) * sstride], of B behind it), the loops over
// coefficients are real loops (small code), only the Chien terms are pulled into registers.
// Z as in pk_bs_decode (may point to global scratch).
template <int M, int T, class SFn>
PK_HD uint32_t pk_bs_decode_mem(SFn getS, uint32_t *st, int sstride, uint32_t *Z, int zstride) {
    constexpr int N = PkGF<M>::N;
    constexpr int LB = (2 * T < 2) ? 1 : (2 * T < 4) ? 2 : (2 * T < 8) ? 3 : (2 * T < 16) ? 4 : 5;
    uint32_t *Lam = st, *Bp = st + (size_t)(T + 1) * M * sstride;
    uint32_t gamma[M], Lb[LB];
    for (int e = 0; e < 2 * (T + 1) * M; ++e) st[(size_t)e * sstride] = 0;
    Lam[0] = ~0u;
    Bp[0] = ~0u;
#pragma unroll
    for (int b = 0; b < M; ++b) gamma[b] = 0;
    gamma[0] = ~0u;
#pragma unroll
    for (int b = 0; b < LB; ++b) Lb[b] = 0;

#pragma unroll 1
    for (int it = 0; it < T; ++it) {
        const int r = 2 * it + 1;
        const int dl = (r - 1 < T) ? r - 1 : T, du = (r < T) ? r : T;
        // B <- x*B (only coefficients 0..du can be non-zero afterwards)
#pragma unroll 1
        for (int k = du; k >= 1; --k)
#pragma unroll
            for (int b = 0; b < M; ++b) Bp[(size_t)(k * M + b) * sstride] = Bp[(size_t)((k - 1) * M + b) * sstride];
#pragma unroll
        for (int b = 0; b < M; ++b) Bp[(size_t)b * sstride] = 0;
        uint32_t acc[2 * M - 1], delta[M];
#pragma unroll
        for (int b = 0; b < 2 * M - 1; ++b) acc[b] = 0;
#pragma unroll 1
        for (int k = 0; k <= dl; ++k) {
            uint32_t sj[M], lk[M];
            getS(r - k, sj);
#pragma unroll
            for (int b = 0; b < M; ++b) lk[b] = Lam[(size_t)(k * M + b) * sstride];
            pk_bs_mac<M>(acc, lk, sj);
        }
        pk_bs_reduce<M>(acc);
        uint32_t nz = 0;
#pragma unroll
        for (int b = 0; b < M; ++b) { delta[b] = acc[b]; nz |= acc[b]; }
        uint32_t lt = 0, eq = ~0u;
#pragma unroll
        for (int b = LB - 1; b >= 0; --b) {
            const uint32_t cb = 0u - (uint32_t)((it >> b) & 1);
            lt |= eq & ~Lb[b] & cb;
            eq &= ~(Lb[b] ^ cb);
        }
        const uint32_t upd = nz & (lt | eq);
#pragma unroll 1
        for (int k = 0; k <= du; ++k) {
            uint32_t lk[M], bk[M];
#pragma unroll
            for (int b = 0; b < M; ++b) {
                lk[b] = Lam[(size_t)(k * M + b) * sstride];
                bk[b] = Bp[(size_t)(k * M + b) * sstride];
            }
#pragma unroll
            for (int b = 0; b < 2 * M - 1; ++b) acc[b] = 0;
            pk_bs_mac<M>(acc, gamma, lk);
            pk_bs_mac<M>(acc, delta, bk);
            pk_bs_reduce<M>(acc);
#pragma unroll
            for (int b = 0; b < M; ++b) {
                Bp[(size_t)(k * M + b) * sstride] = (bk[b] & ~upd) | (lk[b] & upd);
                Lam[(size_t)(k * M + b) * sstride] = acc[b];
            }
        }
        uint32_t carry = ~0u;
#pragma unroll
        for (int b = 0; b < LB; ++b) {
            const uint32_t rb = 0u - (uint32_t)((r >> b) & 1);
            const uint32_t nl = ~Lb[b];
            const uint32_t sum = rb ^ nl ^ carry;
            carry = (rb & nl) | (carry & (rb ^ nl));
            Lb[b] = (Lb[b] & ~upd) | (sum & upd);
        }
#pragma unroll
        for (int b = 0; b < M; ++b) gamma[b] = (gamma[b] & ~upd) | (delta[b] & upd);
        // B <- x*B for the even step r+1
        const int du2 = (r + 1 < T) ? r + 1 : T;
#pragma unroll 1
        for (int k = du2; k >= 1; --k)
#pragma unroll
            for (int b = 0; b < M; ++b) Bp[(size_t)(k * M + b) * sstride] = Bp[(size_t)((k - 1) * M + b) * sstride];
#pragma unroll
        for (int b = 0; b < M; ++b) Bp[(size_t)b * sstride] = 0;
    }
    uint32_t okL;
    {
        uint32_t lt = 0, eq = ~0u;
#pragma unroll
        for (int b = LB - 1; b >= 0; --b) {
            if ((T >> b) & 1) { lt |= eq & ~Lb[b]; eq &= Lb[b]; }
            else eq &= ~Lb[b];
        }
        okL = lt | eq;
    }
    // Chien terms into registers; degree and Lambda(0) on the way
    uint32_t term[T + 1][M], dB[LB], dge1 = 0, lam0 = 0;
#pragma unroll
    for (int b = 0; b < LB; ++b) dB[b] = 0;
#pragma unroll
    for (int k = 0; k <= T; ++k) {
        uint32_t nzk = 0;
#pragma unroll
        for (int b = 0; b < M; ++b) { term[k][b] = Lam[(size_t)(k * M + b) * sstride]; nzk |= term[k][b]; }
        if (k == 0) lam0 = nzk;
        else {
            dge1 |= nzk;
#pragma unroll
            for (int b = 0; b < LB; ++b) dB[b] = (dB[b] & ~nzk) | (((k >> b) & 1) ? nzk : 0u);
        }
    }
    uint32_t cnt[LB];
#pragma unroll
    for (int b = 0; b < LB; ++b) cnt[b] = 0;
#pragma unroll 1
    for (int p = 0; p < N; ++p) {
        uint32_t nzv = 0;
#pragma unroll
        for (int b = 0; b < M; ++b) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k <= T; ++k) v ^= term[k][b];
            nzv |= v;
        }
        const uint32_t z = ~nzv;
        Z[(size_t)p * zstride] = z;
        uint32_t carry = z;
#pragma unroll
        for (int b = 0; b < LB; ++b) {
            const uint32_t t2 = cnt[b] & carry;
            cnt[b] ^= carry;
            carry = t2;
        }
        PkBsChienStep<M, T, 1>::run(term);
    }
    uint32_t same = ~0u;
#pragma unroll
    for (int b = 0; b < LB; ++b) same &= ~(cnt[b] ^ dB[b]);
    return okL & dge1 & lam0 & same;
}
