// pk_kernels.cuh -- sm_100a kernels of the Kaneko/BCH Monte-Carlo hot path.
//
// Mapping (reference file:line -> here):
//   KanekoKernelProcessor::decode(answer,word,res)  src/KanekoKernelProcessor.cpp:335-407 -> KanekoWarp::decode
//   calcError / alterSyndromPoly                    :36-51, src/Decoder.cpp:210-230       -> XOR of "augmented columns"
//   Decoder::decode (euclid + Chien)                src/Decoder.cpp:233-321               -> pk_alg_decode / coset table
//   calcM / calcL / calcRightSide / calcT           :54-126                               -> popc, ordered fp64 sums
//   generateRandomPoly / multiplyPolynomials / addNoise  src/bchCoder.cpp:120-132,236-250 -> k_generate front end
//   fun() per-frame bookkeeping                     src/dataForPlot.cpp:66-73             -> k_generate back end
//
// Execution model: persistent CTAs, ONE WARP PER FRAME, frames pulled from an atomic queue
// (per-frame work spans 1 .. 2^31 trials).  Inside a frame the 32 lanes evaluate 32
// consecutive test patterns per step speculatively; improvements are then committed IN
// PATTERN ORDER (ballot + shuffles), so the sequential semantics of the reference loop
// -- l0, m0, the shrinking bound (1 << T) - 1 and all three operation counters -- are
// reproduced exactly, and only trials the sequential loop would have run are counted.
// Pattern i's word differs from yH by an XOR of per-position columns, so its syndromes
// are S(yH) ^ (uniform part for i >> 5) ^ (per-lane part for i & 31): calcError and
// alterSyndromPoly collapse into a handful of XORs.
#pragma once
#include <cfloat>
#include <cuda_runtime.h>

#include "pk_alg.cuh"
#include "pk_kernels.h"

#define PK_FULL 0xFFFFFFFFu
#ifndef PK_WARPS
#define PK_WARPS 8
#endif

// totals slots (pk_point_result layout)
enum { PK_T_FRAMES = 0, PK_T_FERR, PK_T_BERR, PK_T_TRIALS, PK_T_CMP, PK_T_SUM, PK_T_MAXTR, PK_T_FLAGS };

__host__ __device__ constexpr size_t pk_align16(size_t x) { return (x + 15) & ~(size_t)15; }

// (1 << T) - 1 as the reference's x86 build evaluates it: 32-bit SHL masks the count to
// 5 bits, -fwrapv wraps the subtraction (KanekoKernelProcessor.cpp:354,361).
__device__ __forceinline__ uint32_t pk_pattern_bound(int T) { return (1u << (T & 31)) - 1u; }

// ------------------------------------------------------------------ Philox4x32-10
struct PkPhilox {
    uint32_t c[4];
};
__device__ __forceinline__ PkPhilox pk_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    PkPhilox o;
    o.c[0] = c0; o.c[1] = c1; o.c[2] = c2; o.c[3] = c3;
    return o;
}

// ------------------------------------------------------------------ shared memory plan
template <int M, int T, bool LUT>
struct PkSmem {
    typedef PkCfg<M, T> C;
    static constexpr int NP = C::NW * 32;             // padded length
    static constexpr int SW = LUT ? 1 : C::NSW;       // syndrome words per column
    static constexpr size_t MUL_OFF = 0;
    static constexpr size_t MUL_SZ = LUT ? 0 : ((size_t)1 << (2 * M));
    static constexpr size_t XOFF_OFF = MUL_OFF + MUL_SZ;
    static constexpr size_t XOFF_SZ = LUT ? 0 : pk_align16((size_t)C::N * 2);
    static constexpr size_t COL_OFF = XOFF_OFF + XOFF_SZ;
    static constexpr size_t COL_SZ = pk_align16((size_t)C::N * SW * 4);
    static constexpr size_t LUT_OFF = COL_OFF + COL_SZ;
    __host__ __device__ static constexpr size_t lut_sz(int nk) { return LUT ? ((size_t)2 << nk) : 0; }
    // per warp
    static constexpr size_t W_ALPHA = 0;                              // double[NP]   |alpha| by position
    static constexpr size_t W_SKEY = W_ALPHA + (size_t)NP * 8;        // double[NP+2] |alpha| ascending
    static constexpr size_t W_SIDX = W_SKEY + (size_t)(NP + 2) * 8;   // uint8 [NP]   position of rank r
    static constexpr size_t W_U = W_SIDX + (size_t)NP;                // uint32[NW+1] info bits (generation)
    static constexpr size_t W_SZ = pk_align16(W_U + (size_t)(C::NW + 1) * 4);
    __host__ __device__ static constexpr size_t total(int nk, int warps) {
        return LUT_OFF + pk_align16(lut_sz(nk)) + (size_t)warps * W_SZ;
    }
};

template <int NW>
__device__ __forceinline__ uint32_t pk_getbit(const uint32_t (&F)[NW], int p) {
    uint32_t w = F[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) w = ((p >> 5) == i) ? F[i] : w;
    return (w >> (p & 31)) & 1u;
}

// ------------------------------------------------------------------ one frame, one warp
template <int M, int T, bool LUT>
struct KanekoWarp {
    typedef PkCfg<M, T> C;
    typedef PkSmem<M, T, LUT> SM;
    static constexpr int N = C::N, NW = C::NW, SW = SM::SW, NA = SW + NW;

    struct Result {
        uint32_t YH[NW];    // hard decisions of the received word
        uint32_t F[NW];     // best flip set: decided = YH ^ F
        uint32_t trials, extra_cmp, extra_sum, flags;
    };

    // yv[w]: channel output of position lane + 32 w.  All lanes return the same Result.
    __device__ static void decode(const uint8_t *s_mul, const uint16_t *s_xoff, const uint32_t *s_col,
                                  const uint16_t *s_lut, double *w_alpha, double *w_skey, uint8_t *w_sidx,
                                  const double (&yv)[NW], const PkKanekoParams &kp, Result &R) {
        const int lane = threadIdx.x & 31;
        uint32_t flags = 0;

        // ---- alpha_i = 2 y_i / sd0^2, yH_i = (alpha_i > 0), reliabilities (KanekoKernelProcessor.cpp:336-342)
        unsigned long long keyb[NW];
        bool hard[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int p = lane + 32 * w;
            const double a = (2.0 * yv[w]) / kp.llr_den;
            const bool valid = p < N;
            hard[w] = valid && !(a <= 0.0);
            const double key = fabs(a);
            keyb[w] = (unsigned long long)__double_as_longlong(key);
            if (valid) w_alpha[p] = key;
            R.YH[w] = __ballot_sync(PK_FULL, hard[w]);
        }
        __syncwarp();

        // ---- std::sort by |alpha| ascending (:343) as a rank sort; ties broken by position
        // (== std::sort for n <= 16 where libstdc++ runs a plain insertion sort; flagged otherwise).
        {
            int rank[NW];
            bool tie = false;
#pragma unroll
            for (int w = 0; w < NW; ++w) rank[w] = 0;
            const unsigned long long *ak = reinterpret_cast<const unsigned long long *>(w_alpha);
#pragma unroll 4
            for (int q = 0; q < N; ++q) {
                const unsigned long long kq = ak[q];
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const int p = lane + 32 * w;
                    const bool eq = (kq == keyb[w]);
                    rank[w] += (kq < keyb[w] || (eq && q < p)) ? 1 : 0;
                    tie |= eq && (q != p) && (p < N);
                }
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int p = lane + 32 * w;
                if (p < N) {
                    w_skey[rank[w]] = __longlong_as_double((long long)keyb[w]);
                    w_sidx[rank[w]] = (uint8_t)p;
                }
            }
            if (lane == 0) w_skey[N] = 0.0;  // the reference reads one past the end in calcT(n-t); value unused
            if (__any_sync(PK_FULL, tie)) flags |= PK_FLAG_SORT_TIE;
        }
        __syncwarp();

        // ---- syndrome of yH (Decoder::findSyndromPoly, Decoder.cpp:184-207): XOR of columns
        uint32_t S0[SW];
        {
            uint32_t acc[SW];
#pragma unroll
            for (int s = 0; s < SW; ++s) acc[s] = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int p = lane + 32 * w;
                if (hard[w]) {
#pragma unroll
                    for (int s = 0; s < SW; ++s) acc[s] ^= s_col[p * SW + s];
                }
            }
#pragma unroll
            for (int s = 0; s < SW; ++s) S0[s] = __reduce_xor_sync(PK_FULL, acc[s]);
        }

        // ---- lane b keeps the augmented column of the b-th least reliable position
        // (pattern bit b flips that position, calcError :36-51).  Bits >= 31 are never set
        // because the bound never exceeds 2^31 - 1.
        uint32_t aug[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) aug[a] = 0;
        if (lane < 31 && lane < N) {
            const int p = w_sidx[lane];
#pragma unroll
            for (int s = 0; s < SW; ++s) aug[s] = s_col[p * SW + s];
#pragma unroll
            for (int w = 0; w < NW; ++w) aug[SW + w] = ((p >> 5) == w) ? (1u << (p & 31)) : 0u;
        }
        uint32_t Vl[NA], Vb[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) { Vl[a] = 0; Vb[a] = 0; }
#pragma unroll
        for (int b = 0; b < 5; ++b) {
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const uint32_t c = __shfl_sync(PK_FULL, aug[a], b);
                Vl[a] ^= ((lane >> b) & 1) ? c : 0u;
            }
        }

        // ---- the test-pattern loop (:354-405)
        uint32_t bound = pk_pattern_bound(N);   // long T = n
        double l0 = DBL_MAX;
        bool first_ok = true, have = false, early = false;
        int m0 = 0;
        uint32_t bestF[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) bestF[w] = 0;
        uint32_t trials = 0, tsteps = 0, nimpr = 0;

        uint32_t step = 0;
        for (uint32_t base = 0;; base += 32, ++step) {
            if (base >= kp.max_trials) { trials = base; flags |= PK_FLAG_TRUNCATED; break; }
            if (step) {   // uniform part of the pattern: bits 5.. of i
                uint32_t diff = step ^ (step - 1);
                while (diff) {
                    const int b = __ffs(diff) - 1;
                    diff &= diff - 1;
#pragma unroll
                    for (int a = 0; a < NA; ++a) Vb[a] ^= __shfl_sync(PK_FULL, aug[a], 5 + b);
                }
            }
            const uint32_t i = base + lane;
            bool active = i < bound;

            uint32_t Sx[SW], A[NW], F[NW];
#pragma unroll
            for (int s = 0; s < SW; ++s) Sx[s] = S0[s] ^ Vb[s] ^ Vl[s];
            bool succ;
            if constexpr (LUT) {
                const uint32_t e = s_lut[Sx[0]];
                succ = !(e & 0x8000u);
#pragma unroll
                for (int w = 0; w < NW; ++w) A[w] = 0;
#pragma unroll
                for (int j = 0; j < T; ++j) {
                    const uint32_t p = (e >> (j * M)) & (uint32_t)N;   // N = "none"
                    if constexpr (NW == 1) {
                        A[0] |= 1u << p;                               // p == N sets the unused top bit
                    } else {
#pragma unroll
                        for (int w = 0; w < NW; ++w) A[w] |= ((p >> 5) == (uint32_t)w) ? (1u << (p & 31)) : 0u;
                    }
                }
                A[NW - 1] &= ~(1u << (N & 31));                        // drop the "none" marker bit
            } else {
                succ = pk_alg_decode<M, T>(Sx, s_mul, s_xoff, A);
            }
            succ = succ && active;

            // m = d_H(yH, x) (calcM :89-97), l = sum_{yH != x} alpha in index order (calcL :69-77)
            int m = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                F[w] = Vb[SW + w] ^ Vl[SW + w] ^ A[w];
                m += __popc(F[w]);
            }
            double l = 0.0;
            if (succ) {
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    uint32_t f = F[w];
                    while (f) {
                        const int b = __ffs(f) - 1;
                        f &= f - 1;
                        l += w_alpha[32 * w + b];
                    }
                }
            }
            if (base == 0) first_ok = __shfl_sync(PK_FULL, succ ? 1 : 0, 0) != 0;   // :371

            // ---- in-order commit of improvements (:372-399)
            bool impr_here = false;
            uint32_t last_is = 0;
            uint32_t cand = __ballot_sync(PK_FULL, succ && (l < l0));
            while (cand) {
                const int src = __ffs(cand) - 1;
                const double ls = __shfl_sync(PK_FULL, l, src);
                const int ms = __shfl_sync(PK_FULL, m, src);
                uint32_t Fs[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) Fs[w] = __shfl_sync(PK_FULL, F[w], src);
                const uint32_t is = base + src;
                if (is == 0 || !first_ok) m0 = ms;   // :374
                l0 = ls;
                have = true;
                impr_here = true;
                last_is = is;
#pragma unroll
                for (int w = 0; w < NW; ++w) bestF[w] = Fs[w];

                // calcRightSide (:54-67)
                {
                    const int border = (2 * T + 1) - (ms + m0) / 2;
                    double rs = 0.0;
                    int cnt = 0, j = 0;
                    while (cnt < border && j < N) {
                        const int p = w_sidx[j];
                        if (!pk_getbit<NW>(Fs, p)) { rs += w_skey[j]; ++cnt; }
                        ++j;
                    }
                    if (ls < rs) { early = true; trials = is + 1; break; }   // :380-382
                }
                // while (l >= calcT(j) && j <= n-1-t) ++j   (:384-390, calcT :110-126)
                int jn;
                {
                    const int border = T - (ms + m0) / 2;
                    double bs = 0.0;
                    int cnt = 0, k = 0;
                    while (cnt < border && k < N) {
                        const int p = w_sidx[k];
                        if (!pk_getbit<NW>(Fs, p)) { bs += w_skey[k]; ++cnt; }
                        ++k;
                    }
                    const int jmax = N - 1 - T;
                    jn = jmax + 1;
                    for (int c0 = 0; c0 <= jmax; c0 += 32) {
                        const int jj = c0 + lane;
                        const bool in = jj <= jmax;
                        double tj = bs;
                        if (in) {
#pragma unroll
                            for (int q = 0; q <= T; ++q) tj += w_skey[jj + q];
                        }
                        const uint32_t stop = __ballot_sync(PK_FULL, in && !(ls >= tj));
                        if (stop) { jn = c0 + __ffs(stop) - 1; break; }
                    }
                }
                tsteps += (uint32_t)jn;
                ++nimpr;
                const int Tn = (kp.J >= 0 && jn > kp.J) ? kp.J : jn;   // :392-393
                bound = pk_pattern_bound(Tn);
                active = i < bound;
                const uint32_t later = (src == 31) ? 0u : (PK_FULL << (src + 1));
                cand = __ballot_sync(PK_FULL, succ && active && (l < l0)) & later;
            }
            if (early) break;
            if (bound <= base + 32) {   // the sequential loop ends inside this step
                trials = bound;
                if (impr_here && last_is + 1 > trials) trials = last_is + 1;
                break;
            }
        }

        if (early) flags |= PK_FLAG_EARLY_RETURN;
        if (!have) flags |= PK_FLAG_NO_DECISION;
#pragma unroll
        for (int w = 0; w < NW; ++w) R.F[w] = bestF[w];
        R.trials = trials;
        R.extra_cmp = tsteps + nimpr;   // :386-397
        R.extra_sum = tsteps;
        R.flags = flags;
        __syncwarp();
    }
};

// ------------------------------------------------------------------ table staging
template <int M, int T, bool LUT>
__device__ __forceinline__ void pk_stage_tables(unsigned char *smem, const PkDevTables &tb) {
    typedef PkSmem<M, T, LUT> SM;
    typedef PkCfg<M, T> C;
    const int tid = threadIdx.x, nth = blockDim.x;
    if constexpr (!LUT) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tb.mul);
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem + SM::MUL_OFF);
        for (int i = tid; i < (int)(SM::MUL_SZ / 4); i += nth) dst[i] = src[i];
        uint16_t *xo = reinterpret_cast<uint16_t *>(smem + SM::XOFF_OFF);
        for (int i = tid; i < C::N; i += nth) xo[i] = tb.xoff[i];
        uint32_t *col = reinterpret_cast<uint32_t *>(smem + SM::COL_OFF);
        for (int i = tid; i < C::N * SM::SW; i += nth) col[i] = tb.hcol[i];
    } else {
        uint32_t *col = reinterpret_cast<uint32_t *>(smem + SM::COL_OFF);
        for (int i = tid; i < C::N; i += nth) col[i] = tb.rcol[i];
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tb.lut);
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem + SM::LUT_OFF);
        for (int i = tid; i < (int)(SM::lut_sz(tb.nk) / 4); i += nth) dst[i] = src[i];
    }
    __syncthreads();
}

struct PkWarpTotals {
    unsigned long long frames, ferr, berr, trials, cmp, sum, maxtr, flags;
    __device__ void clear() { frames = ferr = berr = trials = cmp = sum = maxtr = flags = 0; }
    __device__ void add(uint32_t n, uint32_t tr, uint32_t ecmp, uint32_t esum, uint32_t fl, uint32_t be) {
        const unsigned long long run = (unsigned long long)tr - ((fl & PK_FLAG_EARLY_RETURN) ? 1ull : 0ull);
        frames += 1;
        ferr += (fl & PK_FLAG_FRAME_ERROR) ? 1 : 0;
        berr += be;
        trials += tr;
        cmp += run * (n + 6) + ecmp;   // KanekoKernelProcessor.cpp:401-404
        sum += run * (n + 1) + esum;
        maxtr = tr > maxtr ? tr : maxtr;
        flags |= fl;
    }
    __device__ void flush(unsigned long long *tot) const {
        if (!tot || !frames) return;
        atomicAdd(tot + PK_T_FRAMES, frames);
        if (ferr) atomicAdd(tot + PK_T_FERR, ferr);
        if (berr) atomicAdd(tot + PK_T_BERR, berr);
        atomicAdd(tot + PK_T_TRIALS, trials);
        atomicAdd(tot + PK_T_CMP, cmp);
        atomicAdd(tot + PK_T_SUM, sum);
        atomicMax(tot + PK_T_MAXTR, maxtr);
        if (flags) atomicOr(tot + PK_T_FLAGS, flags);
    }
};

// ------------------------------------------------------------------ replay kernel
template <int M, int T, bool LUT>
__global__ void __launch_bounds__(PK_WARPS * 32)
k_replay(PkDevTables tb, PkKanekoParams kp, const double *__restrict__ y, long B, uint8_t *__restrict__ decided,
         uint32_t *__restrict__ trials, pk_frame_rec *__restrict__ recs, unsigned long long *totals,
         unsigned long long *queue) {
    typedef PkSmem<M, T, LUT> SM;
    typedef KanekoWarp<M, T, LUT> KW;
    constexpr int N = KW::N, NW = KW::NW;
    extern __shared__ __align__(16) unsigned char smem[];
    pk_stage_tables<M, T, LUT>(smem, tb);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wb = smem + SM::LUT_OFF + pk_align16(SM::lut_sz(tb.nk)) + (size_t)warp * SM::W_SZ;
    double *w_alpha = reinterpret_cast<double *>(wb + SM::W_ALPHA);
    double *w_skey = reinterpret_cast<double *>(wb + SM::W_SKEY);
    uint8_t *w_sidx = wb + SM::W_SIDX;
    const uint8_t *s_mul = smem + SM::MUL_OFF;
    const uint16_t *s_xoff = reinterpret_cast<const uint16_t *>(smem + SM::XOFF_OFF);
    const uint32_t *s_col = reinterpret_cast<const uint32_t *>(smem + SM::COL_OFF);
    const uint16_t *s_lut = reinterpret_cast<const uint16_t *>(smem + SM::LUT_OFF);

    PkWarpTotals tot;
    tot.clear();
    const int grab = kp.frames_per_grab;
    for (;;) {
        unsigned long long f0 = 0;
        if (lane == 0) f0 = atomicAdd(queue, (unsigned long long)grab);
        f0 = __shfl_sync(PK_FULL, f0, 0);
        if ((long)f0 >= B) break;
        const long f1 = ((long)f0 + grab < B) ? (long)f0 + grab : B;
        for (long f = (long)f0; f < f1; ++f) {
            double yv[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int p = lane + 32 * w;
                yv[w] = (p < N) ? __ldg(y + f * N + p) : 0.0;
            }
            typename KW::Result R;
            KW::decode(s_mul, s_xoff, s_col, s_lut, w_alpha, w_skey, w_sidx, yv, kp, R);
            if (!(R.flags & PK_FLAG_NO_DECISION)) {
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const int p = lane + 32 * w;
                    if (p < N) decided[f * N + p] = (uint8_t)(((R.YH[w] ^ R.F[w]) >> lane) & 1u);
                }
            }
            if (lane == 0) {
                if (trials) trials[f] = R.trials;
                if (recs) {
                    pk_frame_rec r;
                    r.trials = R.trials; r.extra_cmp = R.extra_cmp; r.extra_sum = R.extra_sum;
                    r.bit_errors = 0; r.flags = (uint8_t)R.flags; r.reserved = 0;
                    recs[f] = r;
                }
                tot.add(N, R.trials, R.extra_cmp, R.extra_sum, R.flags, 0);
            }
        }
    }
    if (lane == 0) tot.flush(totals);
}

// ------------------------------------------------------------------ generation kernel
// Frame f of SNR point s draws from Philox4x32-10 with key = seed and counter
// (f_lo, f_hi, block, s): blocks 0.. hold the info bits (128 per block), blocks
// 0x100+q the Box-Muller pair for positions 2q, 2q+1.
template <int M, int T, bool LUT>
__global__ void __launch_bounds__(PK_WARPS * 32)
k_generate(PkDevTables tb, PkKanekoParams kp, PkGenParams gp, long B, pk_frame_rec *__restrict__ recs,
           unsigned long long *totals, unsigned long long *queue, uint8_t *__restrict__ d_info,
           uint8_t *__restrict__ d_cw, double *__restrict__ d_y, int dump_only) {
    typedef PkSmem<M, T, LUT> SM;
    typedef KanekoWarp<M, T, LUT> KW;
    constexpr int N = KW::N, NW = KW::NW;
    extern __shared__ __align__(16) unsigned char smem[];
    pk_stage_tables<M, T, LUT>(smem, tb);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wb = smem + SM::LUT_OFF + pk_align16(SM::lut_sz(tb.nk)) + (size_t)warp * SM::W_SZ;
    double *w_alpha = reinterpret_cast<double *>(wb + SM::W_ALPHA);
    double *w_skey = reinterpret_cast<double *>(wb + SM::W_SKEY);
    uint8_t *w_sidx = wb + SM::W_SIDX;
    uint32_t *w_u = reinterpret_cast<uint32_t *>(wb + SM::W_U);
    const uint8_t *s_mul = smem + SM::MUL_OFF;
    const uint16_t *s_xoff = reinterpret_cast<const uint16_t *>(smem + SM::XOFF_OFF);
    const uint32_t *s_col = reinterpret_cast<const uint32_t *>(smem + SM::COL_OFF);
    const uint16_t *s_lut = reinterpret_cast<const uint16_t *>(smem + SM::LUT_OFF);
    const int K = tb.k, NK = tb.nk;
    const uint32_t k0 = (uint32_t)gp.seed, k1 = (uint32_t)(gp.seed >> 32);
    uint32_t gm[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) gm[w] = tb.gmask[w];

    PkWarpTotals tot;
    tot.clear();
    const int grab = kp.frames_per_grab;
    for (;;) {
        unsigned long long f0 = 0;
        if (lane == 0) f0 = atomicAdd(queue, (unsigned long long)grab);
        f0 = __shfl_sync(PK_FULL, f0, 0);
        if ((long)f0 >= B) break;
        const long f1 = ((long)f0 + grab < B) ? (long)f0 + grab : B;
        for (long f = (long)f0; f < f1; ++f) {
            const unsigned long long gf = gp.first_frame + (unsigned long long)f;
            const uint32_t c0 = (uint32_t)gf, c1 = (uint32_t)(gf >> 32);
            // ---- k information bits (generateRandomPoly, bchCoder.cpp:236-240)
            uint32_t uw = 0;
            if (lane * 32 < K) {
                PkPhilox r = pk_philox(c0, c1, (uint32_t)(lane >> 2), gp.snr_index, k0, k1);
                uw = r.c[lane & 3];
                const int rem = K - lane * 32;
                if (rem < 32) uw &= (1u << rem) - 1u;
            }
            if (lane <= NW) w_u[lane] = uw;   // lane < 32 always >= NW+1 entries
            __syncwarp();
            // ---- c(x) = u(x) g(x) over GF(2) (multiplyPolynomials, bchCoder.cpp:120-132):
            // XOR of (u << d) over the set coefficients g_d, d split across lanes.
            uint32_t cwp[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) cwp[w] = 0;
            for (int d = lane; d <= NK; d += 32) {
                if ((gm[d >> 5] >> (d & 31)) & 1u) {
                    const int ws = d >> 5, bs = d & 31;
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const int lo = w - ws;
                        const uint32_t a = (lo >= 0) ? w_u[lo] : 0u;
                        const uint32_t b = (lo >= 1) ? w_u[lo - 1] : 0u;
                        cwp[w] ^= __funnelshift_l(b, a, bs);
                    }
                }
            }
            uint32_t CW[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) CW[w] = __reduce_xor_sync(PK_FULL, cwp[w]);
            // ---- BPSK + AWGN (addNoise, bchCoder.cpp:243-250): y = (c ? +1 : -1) + N(0, sigma^2)
            for (int q = lane; 2 * q < N; q += 32) {
                PkPhilox r = pk_philox(c0, c1, 0x100u + (uint32_t)q, gp.snr_index, k0, k1);
                const unsigned long long a = ((unsigned long long)r.c[0] << 32) | r.c[1];
                const unsigned long long b = ((unsigned long long)r.c[2] << 32) | r.c[3];
                const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);   // (0,1]
                const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);           // [0,1)
                const double rad = sqrt(-2.0 * log(u1));
                double sn, cs;
                sincospi(2.0 * u2, &sn, &cs);
                const int p0 = 2 * q, p1 = 2 * q + 1;
                const double b0 = ((CW[p0 >> 5] >> (p0 & 31)) & 1u) ? 1.0 : -1.0;
                w_skey[p0] = b0 + gp.sigma * (rad * cs);
                if (p1 < N) {
                    const double b1 = pk_getbit<NW>(CW, p1) ? 1.0 : -1.0;
                    w_skey[p1] = b1 + gp.sigma * (rad * sn);
                }
            }
            __syncwarp();
            double yv[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int p = lane + 32 * w;
                yv[w] = (p < N) ? w_skey[p] : 0.0;
            }
            __syncwarp();
            if (d_info) {
                for (int i = lane; i < K; i += 32) d_info[f * K + i] = (uint8_t)((w_u[i >> 5] >> (i & 31)) & 1u);
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int p = lane + 32 * w;
                if (p < N) {
                    if (d_cw) d_cw[f * N + p] = (uint8_t)((CW[w] >> lane) & 1u);
                    if (d_y) d_y[f * N + p] = yv[w];
                }
            }
            __syncwarp();
            if (dump_only) continue;

            typename KW::Result R;
            KW::decode(s_mul, s_xoff, s_col, s_lut, w_alpha, w_skey, w_sidx, yv, kp, R);
            // ---- compare with the transmitted word (dataForPlot.cpp:66-73)
            uint32_t be = 0;
            if (!(R.flags & PK_FLAG_NO_DECISION)) {
#pragma unroll
                for (int w = 0; w < NW; ++w) be += __popc(R.YH[w] ^ R.F[w] ^ CW[w]);
            } else {
#pragma unroll
                for (int w = 0; w < NW; ++w) be += __popc(CW[w]);   // undecided buffer counted as all-zero
            }
            uint32_t fl = R.flags | (be ? PK_FLAG_FRAME_ERROR : 0);
            if (lane == 0) {
                if (recs) {
                    pk_frame_rec r;
                    r.trials = R.trials; r.extra_cmp = R.extra_cmp; r.extra_sum = R.extra_sum;
                    r.bit_errors = (uint16_t)be; r.flags = (uint8_t)fl; r.reserved = 0;
                    recs[f] = r;
                }
                tot.add(N, R.trials, R.extra_cmp, R.extra_sum, fl, be);
            }
        }
    }
    if (lane == 0) tot.flush(totals);
}

// ------------------------------------------------------------------ algebraic decoder alone
// One thread per word: findSyndromPoly + decode (Decoder.cpp:184-207,298-321).
template <int M, int T>
__global__ void __launch_bounds__(128)
k_bdd(PkDevTables tb, const uint8_t *__restrict__ words, long B, uint8_t *__restrict__ answers,
      uint8_t *__restrict__ ok) {
    typedef PkSmem<M, T, false> SM;
    typedef PkCfg<M, T> C;
    extern __shared__ __align__(16) unsigned char smem[];
    pk_stage_tables<M, T, false>(smem, tb);
    const uint8_t *s_mul = smem + SM::MUL_OFF;
    const uint16_t *s_xoff = reinterpret_cast<const uint16_t *>(smem + SM::XOFF_OFF);
    const uint32_t *s_col = reinterpret_cast<const uint32_t *>(smem + SM::COL_OFF);
    for (long f = (long)blockIdx.x * blockDim.x + threadIdx.x; f < B; f += (long)gridDim.x * blockDim.x) {
        uint32_t Sw[C::NSW], A[C::NW];
#pragma unroll
        for (int s = 0; s < C::NSW; ++s) Sw[s] = 0;
        for (int p = 0; p < C::N; ++p) {
            if (words[f * C::N + p]) {
#pragma unroll
                for (int s = 0; s < C::NSW; ++s) Sw[s] ^= s_col[p * C::NSW + s];
            }
        }
        const bool good = pk_alg_decode<M, T>(Sw, s_mul, s_xoff, A);
        ok[f] = good ? 1 : 0;
        if (good) {
            for (int p = 0; p < C::N; ++p)
                answers[f * C::N + p] = (uint8_t)((words[f * C::N + p] ? 1u : 0u) ^ pk_getbit<C::NW>(A, p));
        }
    }
}

// ------------------------------------------------------------------ encoder alone
// One warp per frame: c = u*g, byte per bit in and out (bchCoder.cpp:120-132).
template <int M>
__global__ void __launch_bounds__(256)
k_encode(PkDevTables tb, const uint8_t *__restrict__ info, long B, uint8_t *__restrict__ cw) {
    constexpr int N = (1 << M) - 1, NW = (N + 31) / 32;
    __shared__ uint32_t s_u[8][NW + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = tb.k, NK = tb.nk;
    uint32_t gm[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) gm[w] = tb.gmask[w];
    for (long f = (long)blockIdx.x * 8 + warp; f < B; f += (long)gridDim.x * 8) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int i = lane + 32 * w;
            const uint32_t bits = __ballot_sync(PK_FULL, i < K && info[f * K + i] != 0);
            if (lane == 0) s_u[warp][w] = bits;
        }
        if (lane == 0) s_u[warp][NW] = 0;
        __syncwarp();
        uint32_t cwp[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) cwp[w] = 0;
        for (int d = lane; d <= NK; d += 32) {
            if ((gm[d >> 5] >> (d & 31)) & 1u) {
                const int ws = d >> 5, bs = d & 31;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const int lo = w - ws;
                    const uint32_t a = (lo >= 0) ? s_u[warp][lo] : 0u;
                    const uint32_t b = (lo >= 1) ? s_u[warp][lo - 1] : 0u;
                    cwp[w] ^= __funnelshift_l(b, a, bs);
                }
            }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t c = __reduce_xor_sync(PK_FULL, cwp[w]);
            const int p = lane + 32 * w;
            if (p < N) cw[f * N + p] = (uint8_t)((c >> lane) & 1u);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ launch wrappers
template <int M, int T>
struct PkLaunch {
    typedef PkCfg<M, T> C;

    static bool host_alg(const uint32_t *Sw, const uint8_t *mul, const uint16_t *xoff, uint32_t *A) {
        return pk_alg_decode<M, T>(Sw, mul, xoff, A);
    }

    template <bool LUT>
    static cudaError_t geom_k(int nk, int sm_count, PkLaunchGeom *out) {
        const size_t smem = PkSmem<M, T, LUT>::total(nk, PK_WARPS);
        cudaError_t e = cudaFuncSetAttribute(k_replay<M, T, LUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_generate<M, T, LUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm_a = 0, per_sm_b = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, k_replay<M, T, LUT>, PK_WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, k_generate<M, T, LUT>, PK_WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm_a < 1 || per_sm_b < 1) return cudaErrorLaunchOutOfResources;
        // persistent grids: every resident slot of every SM; out[0] = replay, out[1] = generate
        out[0].grid = sm_count * per_sm_a;
        out[1].grid = sm_count * per_sm_b;
        out[0].block = out[1].block = PK_WARPS * 32;
        out[0].smem = out[1].smem = smem;
        return cudaSuccess;
    }
    // coset-table kernels exist only where the u16 entry (t positions of m bits + flag) fits
    static constexpr bool LUT_OK = (T * M <= 15);
    static cudaError_t geom_kaneko(bool lut, int nk, int sm_count, PkLaunchGeom *out) {
        if constexpr (LUT_OK) {
            if (lut) return geom_k<true>(nk, sm_count, out);
        } else {
            if (lut) return cudaErrorInvalidValue;
        }
        return geom_k<false>(nk, sm_count, out);
    }
    static cudaError_t geom_bdd(int sm_count, PkLaunchGeom *out) {
        const size_t smem = PkSmem<M, T, false>::total(0, 0);
        cudaError_t e = cudaFuncSetAttribute(k_bdd<M, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bdd<M, T>, 128, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        out->grid = sm_count * per_sm;
        out->block = 128;
        out->smem = smem;
        return cudaSuccess;
    }

    static cudaError_t replay(bool lut, const PkLaunchGeom &g, const PkDevTables &tb, const PkKanekoParams &kp,
                              const double *d_y, long B, uint8_t *d_decided, uint32_t *d_trials, pk_frame_rec *d_recs,
                              unsigned long long *d_totals, unsigned int *d_queue, cudaStream_t st) {
        unsigned long long *q = reinterpret_cast<unsigned long long *>(d_queue);
        cudaError_t e = cudaMemsetAsync(q, 0, sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
        if constexpr (LUT_OK) {
            if (lut) {
                k_replay<M, T, true><<<g.grid, g.block, g.smem, st>>>(tb, kp, d_y, B, d_decided, d_trials, d_recs, d_totals, q);
                ++g_pk_launches;
                return cudaGetLastError();
            }
        }
        k_replay<M, T, false><<<g.grid, g.block, g.smem, st>>>(tb, kp, d_y, B, d_decided, d_trials, d_recs, d_totals, q);
        ++g_pk_launches;
        return cudaGetLastError();
    }
    static cudaError_t generate(bool lut, const PkLaunchGeom &g, const PkDevTables &tb, const PkKanekoParams &kp,
                                const PkGenParams &gp, long B, pk_frame_rec *d_recs, unsigned long long *d_totals,
                                unsigned int *d_queue, uint8_t *d_info, uint8_t *d_cw, double *d_y, int dump_only,
                                cudaStream_t st) {
        unsigned long long *q = reinterpret_cast<unsigned long long *>(d_queue);
        cudaError_t e = cudaMemsetAsync(q, 0, sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
        if constexpr (LUT_OK) {
            if (lut) {
                k_generate<M, T, true><<<g.grid, g.block, g.smem, st>>>(tb, kp, gp, B, d_recs, d_totals, q, d_info, d_cw, d_y, dump_only);
                ++g_pk_launches;
                return cudaGetLastError();
            }
        }
        k_generate<M, T, false><<<g.grid, g.block, g.smem, st>>>(tb, kp, gp, B, d_recs, d_totals, q, d_info, d_cw, d_y, dump_only);
        ++g_pk_launches;
        return cudaGetLastError();
    }
    static cudaError_t bdd(const PkLaunchGeom &g, const PkDevTables &tb, const uint8_t *d_words, long B,
                           uint8_t *d_answers, uint8_t *d_ok, cudaStream_t st) {
        long need = (B + g.block - 1) / g.block;
        int grid = (int)(need < g.grid ? (need < 1 ? 1 : need) : g.grid);
        k_bdd<M, T><<<grid, g.block, g.smem, st>>>(tb, d_words, B, d_answers, d_ok);
        ++g_pk_launches;
        return cudaGetLastError();
    }
    static cudaError_t encode(const PkDevTables &tb, const uint8_t *d_info, long B, uint8_t *d_cw, cudaStream_t st) {
        long need = (B + 7) / 8;
        int grid = (int)(need < 148 * 8 ? (need < 1 ? 1 : need) : 148 * 8);
        k_encode<M><<<grid, 256, 0, st>>>(tb, d_info, B, d_cw);
        ++g_pk_launches;
        return cudaGetLastError();
    }

    static constexpr PkKernelSet make() {
        return PkKernelSet{M, T, &host_alg, &geom_kaneko, &geom_bdd, &replay, &generate, &bdd, &encode};
    }
};
