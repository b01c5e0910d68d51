// pk_kernels.cuh -- sm_100a kernels of the Kaneko/BCH Monte-Carlo hot path.
//
// Mapping (reference file:line -> here):
//   KanekoKernelProcessor::decode(answer,word,res)  src/KanekoKernelProcessor.cpp:335-407 -> KanekoWarp
//   calcError / alterSyndromPoly                    :36-51, src/Decoder.cpp:210-230       -> XOR of "augmented columns"
//   Decoder::decode (euclid + Chien)                src/Decoder.cpp:233-321               -> pk_alg_decode / pk_bs_decode / coset table
//   calcM / calcL / calcRightSide / calcT           :54-126                               -> popc, ordered fp64 sums (KanekoWarp::commit)
//   generateRandomPoly / multiplyPolynomials / addNoise  src/bchCoder.cpp:120-132,236-250 -> pk_load_frame<.., GEN = true>
//   fun() per-frame bookkeeping                     src/dataForPlot.cpp:66-73             -> pk_emit
//
// Execution model.  Persistent CTAs, ONE WARP PER FRAME, frames pulled from an atomic queue
// (per-frame work spans 1 .. 2^31 trials).
//   Phase A (k_phase_a): "narrow" search -- the 32 lanes evaluate 32 consecutive test patterns
//     per step (BM+Chien in registers or a coset-table lookup).  Almost every frame at medium
//     and high SNR finishes here.  A frame still running after kp.limit_a trials is parked in a
//     long list with its search state.
//   Phase B (k_phase_b): "wide" search of the parked frames, 1024 test patterns per warp step:
//     lane l, bit q  <->  pattern base + 32 l + q.  For t*m > 15 codes the algebraic decoder
//     runs BIT-SLICED (pk_bs.cuh: no table lookups, ~9 LOP3 per trial); coset-table codes do
//     32 lookups per lane.  Candidates are rare and are committed one by one.
// In both phases improvements are committed IN PATTERN ORDER, so the sequential semantics of
// the reference loop -- l0, m0, the shrinking bound (1 << T) - 1 and all three operation
// counters -- are reproduced exactly, and only trials the sequential loop would have run are
// counted.  Pattern i's word differs from yH by an XOR of per-position columns, so its
// syndromes are S(yH) ^ (XOR of the columns of i's set bits): calcError and alterSyndromPoly
// collapse into a handful of XORs.
#pragma once
#include <cfloat>
#include <cstdlib>
#include <cuda_runtime.h>

#include "pk_alg.cuh"
#include "pk_bs.cuh"
#include "pk_kernels.h"
#include "pk_philox.cuh"
#include "pk_stdsort.cuh"

#define PK_FULL 0xFFFFFFFFu
#define PK_WARPS_A 8        // warps per CTA, phase A
// PK_WARPS_B, PK_WARPS_B_CT, PK_WARPS_B_LUT (warps per phase-B CTA): pk_kernels.h
#ifndef PK_BS_LOOP
#define PK_BS_LOOP true     // bit-sliced BM as one loop body (instruction-cache friendly)
#endif

// totals slots (pk_point_result layout)
enum { PK_T_FRAMES = 0, PK_T_FERR, PK_T_BERR, PK_T_TRIALS, PK_T_CMP, PK_T_SUM, PK_T_MAXTR, PK_T_FLAGS };

__host__ __device__ constexpr size_t pk_align16(size_t x) { return (x + 15) & ~(size_t)15; }

// (1 << T) - 1 as the reference's x86 build evaluates it: 32-bit SHL masks the count to
// 5 bits, -fwrapv wraps the subtraction (KanekoKernelProcessor.cpp:354,361).
__device__ __forceinline__ uint32_t pk_pattern_bound(int T) { return (1u << (T & 31)) - 1u; }
// The 2-argument flavour (file mode, KanekoKernelProcessor.cpp:229,236) loops while i < (uint64_t(1) << T): a 64-bit
// SHL (count masked to 6 bits), no "- 1", and T starts at LONG_MAX (= no bound).  Pattern indices here are 31-bit,
// anything larger is "unbounded" (a search that long ends with PK_FLAG_TRUNCATED at max_trials).
__device__ __forceinline__ uint32_t pk_pattern_bound2(int T) { return ((T & 63) >= 31) ? 0x7FFFFFFFu : (1u << (T & 63)); }

// ------------------------------------------------------------------ compile-time switches per (M,T)
template <int M, int T>
struct PkTraits {
    static constexpr bool LUT_OK = (T * M <= 15);                 // u16 coset-table entry fits
    static constexpr bool BS_OK = ((T + 1) * M <= 56) && !LUT_OK; // bit-sliced decoder fits the register file
    static constexpr bool BSM_OK = !LUT_OK && !BS_OK;             // bit-sliced decoder with its BM state in shared memory
    static constexpr int MINB = BSM_OK ? 2 : 3;                   // phase-B CTAs per SM the register budget is set for
    // cyclic-class table (PkClassTable): key of n-k-m+1 <= 28 bits and key + t*m <= 64 bits, 2^(2m)-byte rank tables
    static constexpr bool CT_OK = !LUT_OK && ((M == 5 && (T == 5 || T == 7)) || (M == 6 && T >= 3 && T <= 6));
    static constexpr int MINB_CT = 16 / PK_WARPS_B_CT;           // 16 warps of 128 registers per SM
};

// Is S_j (odd j >= 3) an independent syndrome, i.e. is the cyclotomic coset of j new among 1, 3, .., j-2?
__host__ __device__ constexpr int pk_ct_rep(int M, int j) {
    const int n = (1 << M) - 1;
    int e = j % n, rep = e;
    for (int i = 0; i < M; ++i) { e = (2 * e) % n; rep = e < rep ? e : rep; }
    return rep;
}
__host__ __device__ constexpr bool pk_ct_used(int M, int j) {
    for (int i = 1; i < j; i += 2)
        if (pk_ct_rep(M, i) == pk_ct_rep(M, j)) return false;
    return true;
}
// number of key fields before S_j's
__host__ __device__ constexpr int pk_ct_index(int M, int j) {
    int c = 0;
    for (int i = 3; i < j; i += 2) c += pk_ct_used(M, i) ? 1 : 0;
    return c;
}

// ------------------------------------------------------------------ shared memory plan
template <int M, int T, bool LUT, bool CT = false>
struct PkSmem {
    static_assert(!(LUT && CT), "coset-table and class-table modes are exclusive");
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
    // per warp (common)
    static constexpr size_t W_ALPHA = 0;                              // double[NP]   |alpha| by position
    static constexpr size_t W_SKEY = W_ALPHA + (size_t)NP * 8;        // double[NP+2] |alpha| ascending
    static constexpr size_t W_PREF = W_SKEY + (size_t)(NP + 2) * 8;   // double[NP+2] prefix sums of skey
    static constexpr size_t W_SIDX = W_PREF + (size_t)(NP + 2) * 8;   // uint8 [NP]   position of rank r
    static constexpr size_t W_U = W_SIDX + (size_t)NP;                // uint32[NW+1] info bits (generation)
    static constexpr size_t W_SZ_A = pk_align16(W_U + (size_t)(C::NW + 1) * 4);
    // per warp, phase B extras
    static constexpr size_t W_CM = W_SZ_A;                                          // uint32[2T*M] planes / [32] deltas / [32][SW] (CT)
    static constexpr size_t W_CM_SZ = pk_align16((size_t)(LUT ? 32 : CT ? 32 * SW : 2 * T * M) * 4);
    static constexpr size_t W_PB = W_CM + W_CM_SZ;                                  // uint32[32][NW] one-hot XORs of the in-word patterns
    static constexpr size_t W_PB_SZ = pk_align16((size_t)32 * C::NW * 4);
    static constexpr size_t W_WL = W_PB + W_PB_SZ;                                  // double[32] in-word pattern reliabilities (LUT, CT)
    static constexpr size_t W_WL_SZ = (LUT || CT) ? 256 : 0;
    static constexpr size_t W_Z = W_WL + W_WL_SZ;                                   // uint32[N][33] root words (bit-sliced)
    static constexpr bool BSM = !LUT && !CT && ((T + 1) * M > 56);                  // BM state in shared memory, Z in global scratch
    static constexpr size_t W_Z_SZ = (LUT || CT || BSM) ? 0 : pk_align16((size_t)C::N * 33 * 4);
    static constexpr size_t W_ST = W_Z + W_Z_SZ;                                    // uint32[2][(T+1)*M][32] Lambda / B planes (BSM)
    static constexpr size_t W_ST_SZ = BSM ? (size_t)2 * (T + 1) * M * 32 * 4 : 0;
    static constexpr size_t W_SZ_B = W_ST + W_ST_SZ;
    static constexpr int ZS = BSM ? 32 : 33;                                        // stride of the root-word buffer
    __host__ __device__ static constexpr size_t tables(int nk) { return LUT_OFF + pk_align16(lut_sz(nk)); }
    __host__ __device__ static constexpr size_t tables_a(int nk) { return CT ? COL_SZ + pk_align16((size_t)1 << M) : tables(nk); }
    __host__ __device__ static constexpr size_t total_a(int nk) { return tables_a(nk) + (size_t)PK_WARPS_A * W_SZ_A; }
    // phase B of the bit-sliced codes needs neither the GF product table nor the Chien offsets: columns only
    static constexpr size_t B_COL_OFF = LUT ? COL_OFF : 0;
    // class-table mode: log S_1 behind the columns; the rank tables (one 2^m x 2^m byte table per independent S_j)
    // are STATIC shared memory (pk_ct_norm_smem) so that their address folds into the LDS immediate
    static constexpr int CT_NJ = CT ? pk_ct_index(M, 2 * T + 1) : 0;
    static constexpr size_t CT_NORM_SZ = (size_t)CT_NJ << (2 * M);
    static constexpr size_t CT_LOG_OFF = COL_SZ;
    static constexpr size_t CT_LOG_SZ = CT ? pk_align16((size_t)1 << M) : 0;
    __host__ __device__ static constexpr size_t tables_b(int nk) { return LUT ? tables(nk) : COL_SZ + CT_LOG_SZ; }
    static constexpr int WB = LUT ? PK_WARPS_B_LUT : CT ? PK_WARPS_B_CT : PK_WARPS_B;   // warps per phase-B CTA
    __host__ __device__ static constexpr size_t total_b(int nk) { return tables_b(nk) + (size_t)WB * W_SZ_B; }
};

template <int NW>
__device__ __forceinline__ uint32_t pk_getbit(const uint32_t (&F)[NW], int p) {
    uint32_t w = F[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) w = ((p >> 5) == i) ? F[i] : w;
    return (w >> (p & 31)) & 1u;
}

template <int M, int NJ>
__device__ __forceinline__ uint8_t *pk_ct_norm_smem() {
    __shared__ __align__(16) uint8_t tab[(size_t)(NJ > 0 ? NJ : 1) << (2 * M)];
    return tab;
}

// ------------------------------------------------------------------ one frame, one warp
template <int M, int T, bool LUT, bool CT = false>
struct KanekoWarp {
    typedef PkCfg<M, T> C;
    typedef PkSmem<M, T, LUT, CT> SM;
    static constexpr int N = C::N, NW = C::NW, SW = SM::SW, NA = SW + NW;

    struct Tables {      // CTA-shared tables
        const uint8_t *mul;
        const uint16_t *xoff;
        const uint32_t *col;
        const uint16_t *lut;
        // class-table mode: rank tables and log S_1 in shared memory, bitmap and position table in global memory
        const uint8_t *ctlog;
        const uint32_t *ctbits;
        unsigned long long cttex;
        const unsigned long long *cthash;
        uint32_t cthshift, cthmask;
        uint32_t ctmult[8];
    };
    struct WarpMem {     // per-warp shared scratch
        double *alpha, *skey, *pref;
        uint8_t *sidx;
        uint32_t *cm, *pb, *z, *st;
        double *wl;
    };
    struct Frame {       // per-frame registers
        uint32_t YH[NW];   // hard decisions (uniform)
        uint32_t S0[SW];   // syndromes / coset index of yH (uniform)
        uint32_t aug[NA];  // lane b: column of the b-th least reliable position | its one-hot mask
        uint32_t flags;
        // extended code (kp.ext): position N carries the overall parity of the BCH codeword
        uint32_t py;       // parity of ALL hard decisions (uniform)
        int rp;            // rank of position N in the reliability order (patterns skip it); huge when not extended
        uint32_t topmask;  // mask of the BCH positions inside word NW-1 (all ones when not extended)
    };
    struct Search {      // the sequential state of the reference loop (uniform)
        double l0;
        int m0;
        bool first_ok, have, early;
        uint32_t bound, trials, tsteps, nimpr, flags;
        uint32_t step_last;   // wide search: 1 + index of the last improvement of the current step, 0 = none
        uint32_t bestF[NW];
    };

    __device__ static WarpMem warp_mem(unsigned char *wb) {
        WarpMem w;
        w.alpha = reinterpret_cast<double *>(wb + SM::W_ALPHA);
        w.skey = reinterpret_cast<double *>(wb + SM::W_SKEY);
        w.pref = reinterpret_cast<double *>(wb + SM::W_PREF);
        w.sidx = wb + SM::W_SIDX;
        w.cm = reinterpret_cast<uint32_t *>(wb + SM::W_CM);
        w.pb = reinterpret_cast<uint32_t *>(wb + SM::W_PB);
        w.z = reinterpret_cast<uint32_t *>(wb + SM::W_Z);
        w.wl = reinterpret_cast<double *>(wb + SM::W_WL);
        w.st = reinterpret_cast<uint32_t *>(wb + SM::W_ST);
        return w;
    }

    // ---- per-frame set-up: alpha, yH, reliability order, syndrome of yH, lane columns
    // (KanekoKernelProcessor.cpp:336-343,359).  yv[w]: channel output of position lane + 32 w.
    __device__ static void setup(const Tables &tb, const WarpMem &wm, const double (&yv)[NW], const PkKanekoParams &kp,
                                 Frame &f) {
        const int lane = threadIdx.x & 31;
        const int NE = N + kp.ext;   // positions of the (possibly extended) code; NE <= 32 NW always
        f.flags = 0;
        unsigned long long keyb[NW];
        bool hard[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int p = lane + 32 * w;
            const double a = (2.0 * yv[w]) / kp.llr_den;      // alpha_i = 2 y_i / sd^2, IEEE divide
            const bool valid = p < NE;
            hard[w] = valid && !(a <= 0.0);                   // yH = (alpha <= 0) ? 0 : 1
            const double key = fabs(a);
            keyb[w] = (unsigned long long)__double_as_longlong(key);
            if (valid) wm.alpha[p] = key;
            f.YH[w] = __ballot_sync(PK_FULL, hard[w]);
        }
        __syncwarp();
        // std::sort by |alpha| ascending (:343) as a rank sort on the f64 bit patterns; ties broken by position
        // (== std::sort for n <= 16, where libstdc++ runs a plain, stable insertion sort).  For n > 16 a frame
        // with equal keys is re-sorted by a literal replay of libstdc++'s introsort (pk_stdsort.cuh).
        {
            // fast pass: rank = number of strictly smaller keys (one fp64 compare per pair: the keys are non-negative
            // doubles, so the fp order is the order of the bit patterns).  Equal keys collide on a rank -- the loser
            // finds somebody else's index in its slot -- and send the frame to the exact pass below.
            int rank[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) rank[w] = 0;
            double keyd[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) keyd[w] = __longlong_as_double((long long)keyb[w]);
#pragma unroll 8
            for (int q = 0; q < NE; ++q) {
                const double kq = wm.alpha[q];
#pragma unroll
                for (int w = 0; w < NW; ++w) rank[w] += (kq < keyd[w]) ? 1 : 0;
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int p = lane + 32 * w;
                if (p < NE) {
                    wm.skey[rank[w]] = keyd[w];
                    wm.sidx[rank[w]] = (uint8_t)p;
                }
            }
            __syncwarp();
            bool clash = false;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int p = lane + 32 * w;
                if (p < NE) clash |= (wm.sidx[rank[w]] != (uint8_t)p);
            }
            if (__any_sync(PK_FULL, clash)) {
                // exact pass: integer compares of the bit patterns, ties broken by position
                __syncwarp();
                bool tie = false;
#pragma unroll
                for (int w = 0; w < NW; ++w) rank[w] = 0;
                const unsigned long long *ak = reinterpret_cast<const unsigned long long *>(wm.alpha);
#pragma unroll 4
                for (int q = 0; q < NE; ++q) {
                    const unsigned long long kq = ak[q];
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const int p = lane + 32 * w;
                        const bool eq = (kq == keyb[w]);
                        rank[w] += (kq < keyb[w] || (eq && q < p)) ? 1 : 0;
                        tie |= eq && (q != p) && (p < NE);
                    }
                }
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const int p = lane + 32 * w;
                    if (p < NE) {
                        wm.skey[rank[w]] = keyd[w];
                        wm.sidx[rank[w]] = (uint8_t)p;
                    }
                }
                if (__any_sync(PK_FULL, tie)) {
                    f.flags |= PK_FLAG_SORT_TIE;
                    if (NE > 16) {
                        // equal keys: std::sort's order is an artefact of libstdc++'s introsort -- replay it on one lane
                        __syncwarp();
                        if (lane == 0) {
                            for (int i = 0; i < NE; ++i) { wm.skey[i] = wm.alpha[i]; wm.sidx[i] = (uint8_t)i; }
                            pk_stdsort::sort(wm.skey, wm.sidx, NE);
                        }
                    }
                }
            }
            if (lane == 0) wm.skey[NE] = 0.0;  // the reference reads one past the end in calcT(n-t); value unused
        }
        __syncwarp();
        // prefix sums of the sorted reliabilities: pref[m] <= l of ANY flip set of weight m (used only to
        // skip exact l computations that cannot beat l0; never decides anything by itself)
        // (a warp scan: the summation order differs from a sequential sum by a few ulp, far inside the 1e-9 margin)
        {
            double carry = 0.0;
            if (lane == 0) wm.pref[0] = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int r = lane + 32 * w;
                double v = (r < NE) ? wm.skey[r] : 0.0;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const double t = __shfl_up_sync(PK_FULL, v, off);
                    if (lane >= off) v += t;
                }
                v += carry;
                if (r < NE) wm.pref[r + 1] = v;
                carry = __shfl_sync(PK_FULL, v, 31);
            }
        }
        // syndrome of yH (Decoder::findSyndromPoly, Decoder.cpp:184-207): XOR of columns
        {
            uint32_t acc[SW];
#pragma unroll
            for (int s = 0; s < SW; ++s) acc[s] = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int p = lane + 32 * w;
                if (hard[w] && p < N) {   // (the parity position of an extended code is not part of the BCH word)
#pragma unroll
                    for (int s = 0; s < SW; ++s) acc[s] ^= tb.col[p * SW + s];
                }
            }
#pragma unroll
            for (int s = 0; s < SW; ++s) f.S0[s] = __reduce_xor_sync(PK_FULL, acc[s]);
        }
        // extended code: parity of all hard decisions, and where the parity position sits in the reliability order --
        // test patterns flip BCH positions only, so pattern bit b is the b-th least reliable position NOT counting it
        {
            uint32_t par = 0;
            f.rp = 0x7FFFFFFF;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                par ^= f.YH[w];
                const int r = lane + 32 * w;
                const uint32_t hit = __ballot_sync(PK_FULL, kp.ext && r < NE && wm.sidx[r] == (uint8_t)N);
                if (hit) f.rp = 32 * w + __ffs(hit) - 1;
            }
            f.py = __popc(par) & 1u;
            f.topmask = ~((uint32_t)kp.ext << (N & 31));
        }
        // lane b keeps the augmented column of the b-th least reliable position (pattern bit b flips that
        // position, calcError :36-51).  Bits >= 31 never occur: the bound never exceeds 2^31 - 1.
#pragma unroll
        for (int a = 0; a < NA; ++a) f.aug[a] = 0;
        if (lane < 31 && lane < N) {
            const int p = wm.sidx[lane + (lane >= f.rp ? 1 : 0)];
#pragma unroll
            for (int s = 0; s < SW; ++s) f.aug[s] = tb.col[p * SW + s];
#pragma unroll
            for (int w = 0; w < NW; ++w) f.aug[SW + w] = ((p >> 5) == w) ? (1u << (p & 31)) : 0u;
        }
        __syncwarp();
    }

    __device__ static void search_init(Search &s, const Frame &f, int variant) {
        s.l0 = DBL_MAX;
        s.m0 = 0;
        s.first_ok = true; s.have = false; s.early = false;
        s.bound = variant ? 0x7FFFFFFFu : pk_pattern_bound(N);      // long T = n (:354) / uint64_t T = LONG_MAX (:229)
        s.trials = 0; s.tsteps = 0; s.nimpr = 0; s.step_last = 0;
        s.flags = f.flags;
#pragma unroll
        for (int w = 0; w < NW; ++w) s.bestF[w] = 0;
    }

    // Could a flip set of weight m still beat l0?  pref[m] is a lower bound of its l; the margin makes the
    // skip exact under fp64 rounding of both sums (2 m ulp << 1e-9).
    __device__ static __forceinline__ bool may_improve(const WarpMem &wm, int m, double l0) {
        return wm.pref[m] * (1.0 - 1e-9) < l0;
    }
    // l = sum over the set bits of F of alpha, ascending position (calcL :69-77)
    __device__ static __forceinline__ double calc_l(const WarpMem &wm, const uint32_t (&F)[NW]) {
        double l = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint32_t fw = F[w];
            while (fw) {
                const int b = __ffs(fw) - 1;
                fw &= fw - 1;
                l += wm.alpha[32 * w + b];
            }
        }
        return l;
    }

    // ---- an improvement found at trial `is` (l < l0): lines :374-398.  Uniform across the warp.
    // Returns true when the search ends through `if (l < calcRightSide()) return;`.
    __device__ static bool commit(Search &s, const WarpMem &wm, const PkKanekoParams &kp, double ls, int ms,
                                  const uint32_t (&Fs)[NW], uint32_t is) {
        const int lane = threadIdx.x & 31;
        const int NE = N + kp.ext;
        if (is == 0 || !s.first_ok || kp.variant == 2) s.m0 = ms;   // :374 (exact rules: m0 = m, the bound of the candidate alone)
        s.l0 = ls;
        s.have = true;
#pragma unroll
        for (int w = 0; w < NW; ++w) s.bestF[w] = Fs[w];
        {   // calcRightSide (:54-67)
            const int border = (2 * T + 1 + kp.ext) - (ms + s.m0) / 2;   // d = 2t+1 (+1: overall parity)
            double rs = 0.0;
            int cnt = 0, j = 0;
            while (cnt < border && j < NE) {
                const int p = wm.sidx[j];
                if (!pk_getbit<NW>(Fs, p)) { rs += wm.skey[j]; ++cnt; }
                ++j;
            }
            if (ls < rs) { s.early = true; s.trials = is + 1; return true; }   // :380-382
        }
        // while (l >= calcT(j) && j <= n-1-t) ++j   (:384-390, calcT :110-126); lanes try 32 j at a time
        int jn;
        {
            const int border = (kp.variant == 2) ? 0 : T - (ms + s.m0) / 2;
            double bs = 0.0;
            int cnt = 0, k = 0;
            while (cnt < border && k < NE) {
                const int p = wm.sidx[k];
                if (!pk_getbit<NW>(Fs, p)) { bs += wm.skey[k]; ++cnt; }
                ++k;
            }
            // 3-argument flavour: j <= n-1-t is part of the loop condition (:384).  The 2-argument flavour (:257) has no
            // such test and would read past the reliability array (undefined behaviour) -- flagged, same stop.
            const int jmax = NE - 1 - T;
            jn = jmax + 1;
            for (int c0 = 0; c0 <= jmax; c0 += 32) {
                const int jj = c0 + lane;
                const bool in = jj <= jmax;
                double tj = bs;
                if (in) {
#pragma unroll
                    for (int q = 0; q <= T; ++q) tj += wm.skey[jj + q];
                }
                const uint32_t stop = __ballot_sync(PK_FULL, in && !(ls >= tj));
                if (stop) { jn = c0 + __ffs(stop) - 1; break; }
            }
        }
        s.tsteps += (uint32_t)jn;
        ++s.nimpr;
        if (kp.variant) {
            if (kp.variant == 1 && jn == NE - T) s.flags |= PK_FLAG_REF_UNDEFINED;
            s.bound = pk_pattern_bound2(jn);                   // T = j (:264); exact rules: every subset of the j least reliable
        } else {
            const int Tn = (kp.J >= 0 && jn > kp.J) ? kp.J : jn;   // :392-393
            s.bound = pk_pattern_bound(Tn);
        }
        return false;
    }

    // extended code: the decided word's parity position is the parity of its BCH part, so its flip bit is
    // parity(yH) ^ parity(F over the BCH positions)
    __device__ static __forceinline__ void ext_bit(const Frame &f, const PkKanekoParams &kp, uint32_t (&F)[NW]) {
        if (kp.ext) {
            uint32_t x = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) x ^= F[w];
            F[NW - 1] |= ((f.py ^ (uint32_t)__popc(x)) & 1u) << (N & 31);
        }
    }
    // reliability of the position pattern bit b flips (the b-th least reliable BCH position)
    __device__ static __forceinline__ double pat_key(const WarpMem &wm, const Frame &f, int b) { return wm.skey[b + (b >= f.rp ? 1 : 0)]; }

    // coset-table entry -> located positions; returns the verdict
    __device__ static __forceinline__ bool lut_positions(uint32_t e, uint32_t (&A)[NW]) {
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
        return !(e & 0x8000u);
    }

    // ---- class-table mode (PkClassTable): key of a packed syndrome, position-table probe, positions
    // w: XOR of pair-packed columns (PkClassTable::col): field i = S_1 | S_{j_i} << M at bit (i % 2) * 2M of word i / 2
    template <int I>
    __device__ static __forceinline__ void ct_key_term(const Tables &tb, const uint32_t (&w)[SW], uint32_t &key) {
        if constexpr (I < SM::CT_NJ) {
            const uint8_t *tab = pk_ct_norm_smem<M, SM::CT_NJ>();
            const uint32_t idx = (I % 2) ? (w[I / 2] >> (2 * M)) : (w[I / 2] & ((1u << (2 * M)) - 1u));
            key += (uint32_t)tab[(I << (2 * M)) + idx] * tb.ctmult[I];
            ct_key_term<I + 1>(tb, w, key);
        }
    }
    __device__ static __forceinline__ uint32_t ct_key(const Tables &tb, const uint32_t (&w)[SW]) {
        uint32_t key = 0;
        ct_key_term<0>(tb, w, key);
        return key;
    }
    __device__ static __forceinline__ unsigned long long ct_find(const Tables &tb, uint32_t key) {
        constexpr int TM = T * M;
        constexpr unsigned long long PM = (TM >= 64) ? ~0ull : ((1ull << (TM & 63)) - 1ull);
        uint32_t h = (key * 0x9E3779B1u) >> tb.cthshift;
        for (;;) {
            const unsigned long long e = __ldg(tb.cthash + h);
            if (e == ~0ull || ((uint32_t)(e >> TM) == key && (e & PM) != PM)) return e;
            h = (h + 1) & tb.cthmask;
        }
    }
    // position j of a class entry, shifted back by s = log S_1; N = none
    __device__ static __forceinline__ uint32_t ct_pos(unsigned long long e, int j, uint32_t s) {
        uint32_t p = (uint32_t)(e >> (j * M)) & (uint32_t)N;
        if (p != (uint32_t)N) {
            p += s;
            p = (p >= (uint32_t)N) ? p - (uint32_t)N : p;
        }
        return p;
    }
    __device__ static __forceinline__ void ct_positions(unsigned long long e, uint32_t s, uint32_t (&A)[NW]) {
#pragma unroll
        for (int w = 0; w < NW; ++w) A[w] = 0;
#pragma unroll
        for (int j = 0; j < T; ++j) {
            const uint32_t p = ct_pos(e, j, s);
            if (p != (uint32_t)N) {
#pragma unroll
                for (int w = 0; w < NW; ++w) A[w] |= ((p >> 5) == (uint32_t)w) ? (1u << (p & 31)) : 0u;
            }
        }
    }

    // ---- narrow search: 32 patterns per step starting at trial base0 (multiple of 32).  Returns true when
    // the frame is finished, false when trial `limit` was reached and the frame must be parked (s holds the
    // state, *next = the next trial to run).
    __device__ static bool narrow(const Tables &tb, const WarpMem &wm, const PkKanekoParams &kp, const Frame &f,
                                  Search &s, uint32_t base0, uint32_t limit, uint32_t *next) {
        const int lane = threadIdx.x & 31;
        uint32_t Vl[NA], Vb[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) { Vl[a] = 0; Vb[a] = 0; }
#pragma unroll
        for (int b = 0; b < 5; ++b) {
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const uint32_t c = __shfl_sync(PK_FULL, f.aug[a], b);
                Vl[a] ^= ((lane >> b) & 1) ? c : 0u;
            }
        }
        {
            uint32_t hb = base0 >> 5;
            while (hb) {
                const int b = __ffs(hb) - 1;
                hb &= hb - 1;
#pragma unroll
                for (int a = 0; a < NA; ++a) Vb[a] ^= __shfl_sync(PK_FULL, f.aug[a], 5 + b);
            }
        }
        for (uint32_t base = base0;; base += 32) {
            if (base >= kp.max_trials) { s.trials = base; s.flags |= PK_FLAG_TRUNCATED; return true; }
            if (base >= limit) { *next = base; return false; }
            const uint32_t i = base + lane;
            uint32_t Sx[SW], A[NW], F[NW];
#pragma unroll
            for (int q = 0; q < SW; ++q) Sx[q] = f.S0[q] ^ Vb[q] ^ Vl[q];
            bool succ;
            if constexpr (LUT) {
                succ = lut_positions(tb.lut[Sx[0]], A);
            } else if constexpr (CT) {
                const uint32_t key = ct_key(tb, Sx);
                succ = ((__ldg(tb.ctbits + (key >> 5)) >> (key & 31)) & 1u) != 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) A[w] = 0;
                if (succ) ct_positions(ct_find(tb, key), tb.ctlog[Sx[0] & (uint32_t)N], A);
            } else {
                succ = pk_alg_decode<M, T>(Sx, tb.mul, tb.xoff, A);
            }
            // `succ` is the decoder's verdict alone: whether trial i is inside the loop bound is tested against the LIVE
            // bound at every ballot below, because an improvement may also RAISE the bound (a larger m shrinks calcT's
            // border sum), which admits patterns of this step that the bound at step start excluded.
            // m = d_H(yH, x) (calcM :89-97), l = sum_{yH != x} alpha in index order (calcL :69-77)
            int m = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) F[w] = Vb[SW + w] ^ Vl[SW + w] ^ A[w];
            ext_bit(f, kp, F);
#pragma unroll
            for (int w = 0; w < NW; ++w) m += __popc(F[w]);
            double l = DBL_MAX;
            if (succ && may_improve(wm, m, s.l0)) l = calc_l(wm, F);
            if (base == 0) s.first_ok = __shfl_sync(PK_FULL, succ ? 1 : 0, 0) != 0;   // :371

            // in-order commit of improvements (:372-399)
            bool impr_here = false;
            uint32_t last_is = 0;
            uint32_t cand = __ballot_sync(PK_FULL, succ && (i < s.bound) && (l < s.l0));
            while (cand) {
                const int src = __ffs(cand) - 1;
                const double ls = __shfl_sync(PK_FULL, l, src);
                const int ms = __shfl_sync(PK_FULL, m, src);
                uint32_t Fs[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) Fs[w] = __shfl_sync(PK_FULL, F[w], src);
                impr_here = true;
                last_is = base + src;
                if (commit(s, wm, kp, ls, ms, Fs, base + src)) return true;
                const uint32_t later = (src == 31) ? 0u : (PK_FULL << (src + 1));
                cand = __ballot_sync(PK_FULL, succ && (i < s.bound) && (l < s.l0)) & later;
            }
            if (s.bound <= base + 32) {   // the sequential loop ends inside this step
                s.trials = s.bound;
                if (impr_here && last_is + 1 > s.trials) s.trials = last_is + 1;
                return true;
            }
            {   // uniform part of the next pattern block: bits 5.. of i
                uint32_t diff = ((base + 32u) ^ base) >> 5;
                while (diff) {
                    const int b = __ffs(diff) - 1;
                    diff &= diff - 1;
#pragma unroll
                    for (int a = 0; a < NA; ++a) Vb[a] ^= __shfl_sync(PK_FULL, f.aug[a], 5 + b);
                }
            }
        }
    }

    // ---- wide search (phase B): 1024 patterns per step, pattern = base + 32*lane + bit.
    // Resumes a parked search at trial `start` (> 0, any multiple of 32): the first step covers the block of
    // 1024 patterns containing `start` with the already-run patterns masked out.
    // G = warps of the CTA cooperating on this frame (warp wi takes block gbase + 1024 wi of every group step);
    // for G > 1 the search state travels through `shared` in warp order (baton passing), so improvements are
    // still committed in pattern order.  All G warps return with identical s.
    //
    // stop_at (a group-step boundary of this search, or 0xFFFFFFFF): the search pauses there and PK_W_STOPPED is returned
    // with s ready to be resumed at `start = stop_at`.
    // dry (helpers of a mega frame, G > 1): nothing is committed; s is a SNAPSHOT of (l0, have, bestF) with the bound wide
    // open, and the answer is PK_W_DIRTY as soon as any pattern of [start, stop_at) survives the candidate filters
    // against it, else PK_W_STOPPED = "no pattern of this range can improve on that state or on any later one"
    // (l0 only decreases, and every former best codeword stays a non-improvement).
    enum { PK_W_FINISHED = 0, PK_W_STOPPED = 1, PK_W_DIRTY = 2 };
    template <int G>
    __device__ static int wide(const Tables &tb, const WarpMem &wm, const PkKanekoParams &kp, const Frame &f,
                               Search &s, uint32_t start, Search *shared, uint32_t *votes, int wi,
                               uint32_t stop_at = 0xFFFFFFFFu, bool dry = false) {
        const int lane = threadIdx.x & 31;
        const uint32_t base0 = (start & ~1023u) + 1024u * (uint32_t)wi;
        // pattern bits 0..4 = bit index in the word: their column contributions are per-frame constants
        if constexpr (LUT || CT) {
            // cm[q][.] = coset-index / packed-syndrome delta of in-word pattern q; pb[q][w] = its one-hot XOR
            uint32_t c[SW], ph[NW];
#pragma unroll
            for (int a = 0; a < SW; ++a) c[a] = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) ph[w] = 0;
#pragma unroll
            for (int b = 0; b < 5; ++b) {
#pragma unroll
                for (int a = 0; a < SW; ++a) {
                    const uint32_t cc = __shfl_sync(PK_FULL, f.aug[a], b);
                    c[a] ^= ((lane >> b) & 1) ? cc : 0u;
                }
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const uint32_t pp = __shfl_sync(PK_FULL, f.aug[SW + w], b);
                    ph[w] ^= ((lane >> b) & 1) ? pp : 0u;
                }
            }
#pragma unroll
            for (int a = 0; a < SW; ++a) wm.cm[lane * SW + a] = c[a];
#pragma unroll
            for (int w = 0; w < NW; ++w) wm.pb[lane * NW + w] = ph[w];
            // wl[q] = sum of the reliabilities flipped by in-word pattern q (for the approximate-l filter)
            double ws = 0.0;
#pragma unroll
            for (int b = 0; b < 5; ++b)
                if (b < N && ((lane >> b) & 1)) ws += pat_key(wm, f, b);
            wm.wl[lane] = ws;
        } else {
            // cm[j*M + b] = plane (over the 32 in-word patterns) of bit b of syndrome S_{j+1}
            for (int si = lane; si < 2 * T * M; si += 32) {
                const int j = si / M, b = si % M;
                const int word = j / C::PER, sh = (j % C::PER) * M + b;
                uint32_t plane = 0;
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    if (q < N) {
                        const uint32_t pat = (q == 0) ? 0xAAAAAAAAu : (q == 1) ? 0xCCCCCCCCu : (q == 2) ? 0xF0F0F0F0u : (q == 3) ? 0xFF00FF00u : 0xFFFF0000u;
                        const uint32_t col = tb.col[(int)wm.sidx[q + (q >= f.rp ? 1 : 0)] * SW + word];
                        plane ^= ((col >> sh) & 1u) ? pat : 0u;
                    }
                }
                wm.cm[si] = plane;
            }
            // pb[q][w] = positions flipped by in-word pattern q (known-codeword test of the candidates)
            uint32_t ph[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) ph[w] = 0;
#pragma unroll
            for (int b = 0; b < 5; ++b) {
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const uint32_t pp = __shfl_sync(PK_FULL, f.aug[SW + w], b);
                    ph[w] ^= ((lane >> b) & 1) ? pp : 0u;
                }
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) wm.pb[lane * NW + w] = ph[w];
        }
        // pattern bits 5..9 = lane index
        uint32_t Ul[NA], Ub[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) { Ul[a] = 0; Ub[a] = 0; }
#pragma unroll
        for (int b = 0; b < 5; ++b) {
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const uint32_t c = __shfl_sync(PK_FULL, f.aug[a], 5 + b);
                Ul[a] ^= ((lane >> b) & 1) ? c : 0u;
            }
        }
        double lsum = 0.0;   // reliabilities flipped by pattern bits 5..9 (lane part)
        if constexpr (LUT || CT) {
#pragma unroll
            for (int b = 0; b < 5; ++b)
                if (5 + b < N && ((lane >> b) & 1)) lsum += pat_key(wm, f, 5 + b);
        }
        // pattern bits 10.. = base
        {
            uint32_t hb = base0 >> 10;
            while (hb) {
                const int b = __ffs(hb) - 1;
                hb &= hb - 1;
#pragma unroll
                for (int a = 0; a < NA; ++a) Ub[a] ^= __shfl_sync(PK_FULL, f.aug[a], 10 + b);
            }
        }
        __syncwarp();
        // Flip sets of the last codewords found worse than l0 (table modes: by this lane's look-ups; bit-sliced mode: by
        // the warp's exact evaluations, identical in all lanes).  Any later pattern within distance t of one of them
        // decodes to that codeword again (never an improvement: l0 only decreases), so it needs no probe / evaluation.
        // All-ones = empty (farther than t from every pattern).
        constexpr int KR = 2;
        uint32_t rej[KR][NW];
#pragma unroll
        for (int k = 0; k < KR; ++k)
#pragma unroll
            for (int w = 0; w < NW; ++w) rej[k][w] = PK_FULL;

        // lo: patterns of the current step below it are not to be run -- `start` in the first step (phase A ran them), and
        // on a RE-RUN of a step the bound the step was first masked with (see the end of the loop body)
        uint32_t lo = start;
        bool rerun = false;
        for (uint32_t base = base0;;) {
            const uint32_t gbase = base - 1024u * (uint32_t)wi;   // first pattern of this group step
            if (gbase >= stop_at && !rerun) return PK_W_STOPPED;
            if (gbase >= kp.max_trials) {
                if (dry) return PK_W_DIRTY;   // the master decides about truncation
                s.trials = gbase; s.flags |= PK_FLAG_TRUNCATED; return PK_W_FINISHED;
            }
            if (!rerun) s.step_last = 0;
            const uint32_t bound0 = s.bound;   // the bound this pass masks its patterns with (identical in all G warps)
            uint32_t u[SW];
#pragma unroll
            for (int q = 0; q < SW; ++q) u[q] = f.S0[q] ^ Ul[q] ^ Ub[q];
            const uint32_t lane_first = base + 32u * lane;
            uint32_t vmask = (s.bound <= lane_first) ? 0u : ((s.bound - lane_first >= 32u) ? PK_FULL : ((1u << (s.bound - lane_first)) - 1u));
            if (lo > lane_first) vmask &= (lo - lane_first >= 32u) ? 0u : (PK_FULL << (lo - lane_first));
            uint32_t cand = 0;   // bit q: trial lane_first + q succeeded (and, LUT / CT: its metric may still beat l0)
            double bsum = 0.0;   // reliabilities flipped by pattern bits 10.. (base part)
            if constexpr (LUT || CT) {
                uint32_t hb = base >> 10;
                while (hb) {
                    const int b = __ffs(hb) - 1;
                    hb &= hb - 1;
                    bsum += pat_key(wm, f, 10 + b);
                }
            }
            // Per-lane filter of decodable trials (bits of `ok`) against the search state `st`, with an APPROXIMATE l
            // (pattern part + located positions, any summation order).  It only discards trials whose l exceeds l0 by
            // far more than the rounding slack, so the exact in-order test below sees every possible improvement
            // (a stale l0 is conservative: l0 only decreases).  Run again on the not yet visited candidates
            // whenever an improvement lowers l0 inside a step.
            auto refine = [&](uint32_t ok, const Search &st) -> uint32_t {
                uint32_t out = 0;
                if constexpr (LUT) {
                    while (ok) {
                        const int q = __ffs(ok) - 1;
                        ok &= ok - 1;
                        uint32_t pat[NW];   // positions flipped by this pattern
#pragma unroll
                        for (int w2 = 0; w2 < NW; ++w2) pat[w2] = Ul[SW + w2] ^ Ub[SW + w2] ^ wm.pb[q * NW + w2];
                        {   // within distance t of a known codeword: decodes to it again (see class-table mode below)
                            int dist = st.have ? 0 : 99;
#pragma unroll
                            for (int w2 = 0; w2 < NW; ++w2) dist += __popc((pat[w2] ^ st.bestF[w2]) & (w2 == NW - 1 ? f.topmask : PK_FULL));
                            bool known = dist <= T;
#pragma unroll
                            for (int k = 0; k < KR; ++k) {
                                int dk = 0;
#pragma unroll
                                for (int w2 = 0; w2 < NW; ++w2) dk += __popc(pat[w2] ^ rej[k][w2]);
                                known = known || (dk <= T);
                            }
                            if (known) continue;
                        }
                        const uint32_t e = tb.lut[u[0] ^ wm.cm[q]];
                        const double lp = wm.wl[q] + lsum + bsum;
                        double la = lp;
                        uint32_t fl[NW];
#pragma unroll
                        for (int w2 = 0; w2 < NW; ++w2) fl[w2] = pat[w2];
#pragma unroll
                        for (int j = 0; j < T; ++j) {
                            const uint32_t p = (e >> (j * M)) & (uint32_t)N;
                            if (p != (uint32_t)N) {
                                uint32_t pw = pat[0];
#pragma unroll
                                for (int w = 1; w < NW; ++w) pw = ((p >> 5) == (uint32_t)w) ? pat[w] : pw;
                                const double a = wm.alpha[p];
                                la += ((pw >> (p & 31)) & 1u) ? -a : a;
#pragma unroll
                                for (int w2 = 0; w2 < NW; ++w2) fl[w2] ^= ((p >> 5) == (uint32_t)w2) ? (1u << (p & 31)) : 0u;
                            }
                        }
                        if (!((la - st.l0) > 1e-9 * (lp + st.l0))) {
                            out |= 1u << q;
                        } else {
#pragma unroll
                            for (int k = KR - 1; k > 0; --k)
#pragma unroll
                                for (int w2 = 0; w2 < NW; ++w2) rej[k][w2] = rej[k - 1][w2];
#pragma unroll
                            for (int w2 = 0; w2 < NW; ++w2) rej[0][w2] = fl[w2];
                        }
                    }
                } else if constexpr (CT) {
                    while (ok) {
                        const int q = __ffs(ok) - 1;
                        ok &= ok - 1;
                        uint32_t pat[NW];   // positions flipped by this pattern
#pragma unroll
                        for (int w2 = 0; w2 < NW; ++w2) pat[w2] = Ul[SW + w2] ^ Ub[SW + w2] ^ wm.pb[q * NW + w2];
                        {
                            // A pattern whose flip set is within distance t of a known codeword's decodes to that
                            // codeword again (bounded-distance decoding is unique).  The best one: l == l0, never an
                            // improvement -- most decodable patterns of a low-SNR frame are of this kind.
                            int dist = st.have ? 0 : 99;
#pragma unroll
                            for (int w2 = 0; w2 < NW; ++w2) dist += __popc((pat[w2] ^ st.bestF[w2]) & (w2 == NW - 1 ? f.topmask : PK_FULL));
                            bool known = dist <= T;
#pragma unroll
                            for (int k = 0; k < KR; ++k) {
                                int dk = 0;
#pragma unroll
                                for (int w2 = 0; w2 < NW; ++w2) dk += __popc(pat[w2] ^ rej[k][w2]);
                                known = known || (dk <= T);
                            }
                            if (known) continue;
                        }
                        uint32_t w[SW];
#pragma unroll
                        for (int a = 0; a < SW; ++a) w[a] = u[a] ^ wm.cm[q * SW + a];
                        const unsigned long long e = ct_find(tb, ct_key(tb, w));
                        const uint32_t sh = tb.ctlog[w[0] & (uint32_t)N];
                        const double lp = wm.wl[q] + lsum + bsum;
                        double la = lp;
                        uint32_t fl[NW];    // flip set of the decoded codeword = pattern ^ located positions
#pragma unroll
                        for (int w2 = 0; w2 < NW; ++w2) fl[w2] = pat[w2];
#pragma unroll
                        for (int j = 0; j < T; ++j) {
                            const uint32_t p = ct_pos(e, j, sh);
                            if (p != (uint32_t)N) {
                                uint32_t pw = pat[0];
#pragma unroll
                                for (int w2 = 1; w2 < NW; ++w2) pw = ((p >> 5) == (uint32_t)w2) ? pat[w2] : pw;
                                const double a = wm.alpha[p];
                                la += ((pw >> (p & 31)) & 1u) ? -a : a;
#pragma unroll
                                for (int w2 = 0; w2 < NW; ++w2) fl[w2] ^= ((p >> 5) == (uint32_t)w2) ? (1u << (p & 31)) : 0u;
                            }
                        }
                        if (!((la - st.l0) > 1e-9 * (lp + st.l0))) {
                            out |= 1u << q;
                        } else {
#pragma unroll
                            for (int k = KR - 1; k > 0; --k)
#pragma unroll
                                for (int w2 = 0; w2 < NW; ++w2) rej[k][w2] = rej[k - 1][w2];
#pragma unroll
                            for (int w2 = 0; w2 < NW; ++w2) rej[0][w2] = fl[w2];
                        }
                    }
                } else {
                    // bit-sliced mode: drop the decodable trials that fall back on a known codeword, the rest is
                    // evaluated exactly
                    while (ok) {
                        const int q = __ffs(ok) - 1;
                        ok &= ok - 1;
                        int dist = st.have ? 0 : 99, d0 = 0, d1 = 0;
#pragma unroll
                        for (int w2 = 0; w2 < NW; ++w2) {
                            const uint32_t pat = Ul[SW + w2] ^ Ub[SW + w2] ^ wm.pb[q * NW + w2];
                            dist += __popc((pat ^ st.bestF[w2]) & (w2 == NW - 1 ? f.topmask : PK_FULL));
                            d0 += __popc(pat ^ rej[0][w2]);
                            d1 += __popc(pat ^ rej[1][w2]);
                        }
                        if (!(dist <= T || d0 <= T || d1 <= T)) out |= 1u << q;
                    }
                }
                return out;
            };
            const uint32_t nimpr0 = s.nimpr;
            if (base >= s.bound || base + 1024u <= lo) {
                // nothing to run in this block (G > 1: blocks past the bound; first step / re-run: blocks below `lo`)
            } else if constexpr (LUT) {
                uint32_t ok = 0;
#pragma unroll 8
                for (int q = 0; q < 32; ++q) {
                    const uint32_t e = tb.lut[u[0] ^ wm.cm[q]];
                    ok |= (e & 0x8000u) ? 0u : (1u << q);
                }
                cand = refine(ok & vmask, s);
            } else if constexpr (CT) {
                // (Skipping the probe of patterns within distance t of a known codeword was tried in round 2: +41 % instructions
                // for the per-lane distance masks, no measurable drop in L2 sectors, 6.1 -> 8.9 ms per launch -- profiles/r2_notes.md.)
                // one bitmap probe per pattern: does ANY error pattern of weight <= t have this syndrome class?
                uint32_t ok = 0;
#pragma unroll 8
                for (int q = 0; q < 32; ++q) {
                    uint32_t w[SW];
#pragma unroll
                    for (int a = 0; a < SW; ++a) w[a] = u[a] ^ wm.cm[q * SW + a];
                    const uint32_t key = ct_key(tb, w);
                    // through the TEX path: the gathers then share the L1TEX data stage with nothing but themselves (the
                    // rank-table lookups go through the LSU path); +3.5 % over __ldg, both limited by the one L2 request
                    // per SM per clock of the L1TEX -> XBAR port
                    const uint32_t word = tex1Dfetch<unsigned int>((cudaTextureObject_t)tb.cttex, (int)(key >> 5));
                    ok |= ((word >> (key & 31)) & 1u) << q;
                }
                // the decodable ones: positions from the class entry, approximate-l filter as in coset-table mode
                cand = refine(ok & vmask, s);
            } else {
                auto getS = [&](int j, uint32_t *o) {   // j may be a run-time value (looped BM)
                    const int wi = (j - 1) / C::PER, sh = ((j - 1) % C::PER) * M;
                    uint32_t uw = u[0];
#pragma unroll
                    for (int q = 1; q < SW; ++q) uw = (wi == q) ? u[q] : uw;
                    uw >>= sh;
#pragma unroll
                    for (int b = 0; b < M; ++b) o[b] = wm.cm[(j - 1) * M + b] ^ (0u - ((uw >> b) & 1u));
                };
                if constexpr (SM::BSM)
                    cand = pk_bs_decode_mem<M, T>(getS, wm.st + lane, 32, wm.z + lane, SM::ZS) & vmask;
                else
                    cand = pk_bs_decode<M, T, PK_BS_LOOP>(getS, wm.z + lane, SM::ZS) & vmask;
                cand = refine(cand, s);
            }
            __syncwarp();

            // exact evaluation of candidate (lane src, bit q) against state st: true iff its metric beats st.l0
            auto eval = [&](int src, int q, uint32_t usrc, uint32_t is, const Search &st, uint32_t (&F)[NW], int &m,
                            double &l) -> bool {
                uint32_t A[NW], P[NW];
                const bool sel = (lane < 31) && ((is >> lane) & 1u);
#pragma unroll
                for (int w = 0; w < NW; ++w) P[w] = __reduce_xor_sync(PK_FULL, sel ? f.aug[SW + w] : 0u);
                if constexpr (!LUT && !CT) {
                    // falls back on a known codeword (the best one or one rejected since the step began)?
                    int dist = st.have ? 0 : 99, d0 = 0, d1 = 0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        dist += __popc((P[w] ^ st.bestF[w]) & (w == NW - 1 ? f.topmask : PK_FULL));
                        d0 += __popc(P[w] ^ rej[0][w]);
                        d1 += __popc(P[w] ^ rej[1][w]);
                    }
                    if (dist <= T || d0 <= T || d1 <= T) return false;
                }
                if constexpr (LUT) {
                    lut_positions(tb.lut[usrc ^ wm.cm[q]], A);
                } else if constexpr (CT) {
                    uint32_t w[SW];
#pragma unroll
                    for (int a = 0; a < SW; ++a) w[a] = __shfl_sync(PK_FULL, u[a], src) ^ wm.cm[q * SW + a];
                    ct_positions(ct_find(tb, ct_key(tb, w)), tb.ctlog[w[0] & (uint32_t)N], A);
                } else {
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const int p = lane + 32 * w;
                        const uint32_t zb = (p < N) ? ((wm.z[p * SM::ZS + src] >> q) & 1u) : 0u;
                        A[w] = __ballot_sync(PK_FULL, zb);
                    }
                }
                m = 0;
                bool better = false;
#pragma unroll
                for (int w = 0; w < NW; ++w) F[w] = P[w] ^ A[w];
                uint32_t Fb[NW];   // flip set over the BCH positions (what the known-codeword tests compare)
#pragma unroll
                for (int w = 0; w < NW; ++w) Fb[w] = F[w];
                ext_bit(f, kp, F);
#pragma unroll
                for (int w = 0; w < NW; ++w) m += __popc(F[w]);
                bool same = st.have;   // same codeword as the current best: l == l0 exactly, no improvement
#pragma unroll
                for (int w = 0; w < NW; ++w) same = same && (F[w] == st.bestF[w]);
                if (same) return false;
                if (may_improve(wm, m, st.l0)) {
                    l = calc_l(wm, F);
                    better = l < st.l0;
                }
                if constexpr (!LUT && !CT) {
                    if (!better) {   // remember the rejected codeword (uniform across the warp)
#pragma unroll
                        for (int w = 0; w < NW; ++w) { rej[1][w] = rej[0][w]; rej[0][w] = Fb[w]; }
                    }
                }
                return better;
            };
            if (G > 1 && !(LUT || CT)) {
                // cooperative search: every warp first thins its own candidates IN PARALLEL against the state
                // at the start of the step (conservative: l0 only decreases), so that the ordered hand-over
                // below only sees the rare real improvements (table modes: refine() has done that already)
                uint32_t keep = 0;
                uint32_t lanes_with = __ballot_sync(PK_FULL, cand != 0);
                while (lanes_with) {
                    const int src = __ffs(lanes_with) - 1;
                    lanes_with &= lanes_with - 1;
                    uint32_t word = __shfl_sync(PK_FULL, cand, src), word2 = 0;
                    const uint32_t usrc = __shfl_sync(PK_FULL, u[0], src);
                    while (word) {
                        const int q = __ffs(word) - 1;
                        word &= word - 1;
                        uint32_t F[NW];
                        int m;
                        double l;
                        if (eval(src, q, usrc, base + 32u * src + q, s, F, m, l)) word2 |= 1u << q;
                    }
                    if (lane == src) keep = word2;
                }
                cand = keep;
            }
            // ---- candidates in pattern order (for G > 1: warp 0's block first, then warp 1's, ...)
            uint32_t turns = 1u;   // bit g: warp g has candidates and takes a turn in the ordered hand-over
            if (G > 1) {
                // most steps leave no candidate in any warp: one barrier settles that and the state is unchanged
                volatile uint32_t *fl = votes + (((gbase >> 10) / G) & 1u) * G;   // double-buffered by step parity
                const uint32_t mine = __ballot_sync(PK_FULL, cand != 0);
                if (lane == 0) fl[wi] = mine ? 1u : 0u;
                __syncthreads();
                turns = 0;
#pragma unroll
                for (int g = 0; g < G; ++g) turns |= (fl[g] ? 1u : 0u) << g;
            }
            if (dry) {
                if (G == 1) turns = __ballot_sync(PK_FULL, cand != 0) ? 1u : 0u;
                if (turns) return PK_W_DIRTY;
            }
            if (turns) {
                // only the warps that have candidates take a turn (one barrier each): late in a long search that is
                // one warp in sixteen
                bool first = true;
#pragma unroll 1
                for (uint32_t mk = turns; mk; mk &= mk - 1) {
                    const int turn = __ffs(mk) - 1;
                    if (turn == wi) {
                        if (G > 1 && !first) s = *shared;
                        bool stop = s.early;
                        if (s.nimpr != nimpr0) cand = refine(cand, s);   // earlier warps of this step lowered l0
                        while (!stop) {
                            const uint32_t lanes_with = __ballot_sync(PK_FULL, cand != 0);
                            if (!lanes_with) break;
                            const int src = __ffs(lanes_with) - 1;
                            uint32_t word = __shfl_sync(PK_FULL, cand, src);
                            const uint32_t usrc = __shfl_sync(PK_FULL, u[0], src);
                            bool improved = false;
                            while (word) {
                                const int q = __ffs(word) - 1;
                                word &= word - 1;
                                if (lane == src) cand &= ~(1u << q);   // visited
                                const uint32_t is = base + 32u * src + q;
                                if (is >= s.bound) { stop = true; break; }
                                uint32_t F[NW];
                                int m;
                                double l;
                                if (eval(src, q, usrc, is, s, F, m, l)) {
                                    s.step_last = is + 1;
                                    if (commit(s, wm, kp, l, m, F, is)) { stop = true; break; }
                                    improved = true;
                                    break;
                                }
                            }
                            // l0 went down: filter what is left of the step again, all lanes in parallel, instead of
                            // evaluating every stale candidate exactly one by one
                            if (improved && !stop) cand = refine(cand, s);
                        }
                        if (G > 1 && lane == 0) *shared = s;
                    }
                    first = false;
                    if (G > 1) __syncthreads();
                }
                if (G > 1) {
                    s = *shared;
                    __syncthreads();   // everyone has re-read the state before a later step overwrites it
                }
            }
            if (s.early) return PK_W_FINISHED;
            // An improvement may RAISE the bound (T = j can grow when a larger m shrinks calcT's border sum).  If the bound
            // this pass was masked with ended inside the step, the patterns [bound0, min(new bound, end of step)) have not
            // been looked at yet although the sequential loop runs them: re-run the step for exactly those.
            if (s.bound > bound0 && bound0 < gbase + 1024u * G) {
                lo = bound0;
                rerun = true;
                __syncwarp();
                continue;
            }
            if (s.bound <= gbase + 1024u * G) {
                s.trials = s.bound;
                if (s.step_last > s.trials) s.trials = s.step_last;
                return PK_W_FINISHED;
            }
            lo = 0;
            rerun = false;
            {   // next base: bits 10.. change
                uint32_t diff = ((base + 1024u * G) ^ base) >> 10;
                while (diff) {
                    const int b = __ffs(diff) - 1;
                    diff &= diff - 1;
#pragma unroll
                    for (int a = 0; a < NA; ++a) Ub[a] ^= __shfl_sync(PK_FULL, f.aug[a], 10 + b);
                }
            }
            base += 1024u * G;
            __syncwarp();
        }
    }

    __device__ static void search_finish(Search &s) {
        if (s.early) s.flags |= PK_FLAG_EARLY_RETURN;
        if (!s.have) s.flags |= PK_FLAG_NO_DECISION;
    }
    __device__ static void park(const Search &s, uint32_t frame, uint32_t base, PkLongRec *r) {
        r->l0 = s.l0;
        r->frame = frame; r->base = base; r->bound = s.bound; r->m0 = (uint32_t)s.m0;
        r->tsteps = s.tsteps; r->nimpr = s.nimpr;
        r->sflags = (s.first_ok ? 1u : 0u) | (s.have ? 2u : 0u) | (s.flags << 8);
        r->pad = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) r->bestF[w] = s.bestF[w];
    }
    __device__ static void unpark(Search &s, const PkLongRec *r) {
        s.l0 = r->l0;
        s.m0 = (int)r->m0;
        s.bound = r->bound;
        s.tsteps = r->tsteps; s.nimpr = r->nimpr;
        s.first_ok = (r->sflags & 1u) != 0;
        s.have = (r->sflags & 2u) != 0;
        s.early = false;
        s.flags = r->sflags >> 8;
        s.trials = 0; s.step_last = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s.bestF[w] = r->bestF[w];
    }
};

// ------------------------------------------------------------------ table staging
template <int M, int T, bool LUT, bool CT = false>
__device__ __forceinline__ void pk_stage_tables(unsigned char *smem, const PkDevTables &tb, bool need_mul) {
    typedef PkSmem<M, T, LUT, CT> SM;
    typedef PkCfg<M, T> C;
    const int tid = threadIdx.x, nth = blockDim.x;
    if constexpr (!LUT) {
        if (need_mul) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(tb.mul);
            uint32_t *dst = reinterpret_cast<uint32_t *>(smem + SM::MUL_OFF);
            for (int i = tid; i < (int)(SM::MUL_SZ / 4); i += nth) dst[i] = src[i];
            uint16_t *xo = reinterpret_cast<uint16_t *>(smem + SM::XOFF_OFF);
            for (int i = tid; i < C::N; i += nth) xo[i] = tb.xoff[i];
        }
        uint32_t *col = reinterpret_cast<uint32_t *>(smem + (need_mul ? SM::COL_OFF : SM::B_COL_OFF));
        if constexpr (CT) {
            // class-table search: pair-packed columns instead of the S_1..S_2t columns (the algebraic decoder is not used)
            for (int i = tid; i < C::N * SM::SW; i += nth) col[i] = tb.ct_col[i];
            const uint32_t *src = reinterpret_cast<const uint32_t *>(tb.ct_norm);
            uint32_t *dst = reinterpret_cast<uint32_t *>(pk_ct_norm_smem<M, SM::CT_NJ>());
            for (int i = tid; i < (int)(SM::CT_NORM_SZ / 4); i += nth) dst[i] = src[i];
            for (int i = tid; i < (1 << M); i += nth) smem[SM::CT_LOG_OFF + i] = tb.ct_log[i];
        } else {
            for (int i = tid; i < C::N * SM::SW; i += nth) col[i] = tb.hcol[i];
        }
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

// Frame f of SNR point s draws from Philox4x32-10 with key = seed and counter (f_lo, f_hi, block, s):
// blocks 0.. hold the info bits (128 per block), blocks 0x100+q the Box-Muller pair of positions 2q, 2q+1.
template <int M, int NW, bool GEN>
__device__ __forceinline__ void pk_load_frame(const PkIo &io, const PkDevTables &tb, long f, double *stage, uint32_t *w_u,
                                              bool dump, double (&yv)[NW], uint32_t (&CW)[NW], int ext) {
    constexpr int N = (1 << M) - 1;
    const int NE = N + ext;   // frame length: the extended code appends the overall parity at position N
    const int lane = threadIdx.x & 31;
    if constexpr (!GEN) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int p = lane + 32 * w;
            yv[w] = (p < NE) ? __ldcs(io.y + f * NE + p) : 0.0;   // streamed once (evict-first): L2 belongs to the 64 MB of class tables
            CW[w] = 0;
        }
    } else {
        const int K = tb.k, NK = tb.nk;
        const uint32_t k0 = (uint32_t)io.gp.seed, k1 = (uint32_t)(io.gp.seed >> 32);
        const unsigned long long gf = io.gp.first_frame + (unsigned long long)f;
        const uint32_t c0 = (uint32_t)gf, c1 = (uint32_t)(gf >> 32);
        // k information bits (generateRandomPoly, bchCoder.cpp:236-240)
        uint32_t uw = 0;
        if (lane * 32 < K) {
            PkPhilox r = pk_philox(c0, c1, (uint32_t)(lane >> 2), io.gp.snr_index, k0, k1);
            uw = r.c[lane & 3];
            const int rem = K - lane * 32;
            if (rem < 32) uw &= (1u << rem) - 1u;
        }
        if (lane <= NW) w_u[lane] = uw;
        __syncwarp();
        // c(x) = u(x) g(x) over GF(2) (multiplyPolynomials, bchCoder.cpp:120-132): XOR of (u << d) over the
        // set coefficients g_d, d split across lanes, then one warp XOR-reduce per word.
        uint32_t cwp[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) cwp[w] = 0;
        for (int d = lane; d <= NK; d += 32) {
            if ((tb.gmask[d >> 5] >> (d & 31)) & 1u) {
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
#pragma unroll
        for (int w = 0; w < NW; ++w) CW[w] = __reduce_xor_sync(PK_FULL, cwp[w]);
        if (ext) {   // overall parity of the BCH codeword
            uint32_t x = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) x ^= CW[w];
            CW[NW - 1] |= ((uint32_t)__popc(x) & 1u) << (N & 31);
        }
        // BPSK + AWGN (addNoise, bchCoder.cpp:243-250): y = (c ? +1 : -1) + N(0, sigma^2), f64 Box-Muller
        for (int q = lane; 2 * q < NE; q += 32) {
            PkPhilox r = pk_philox(c0, c1, 0x100u + (uint32_t)q, io.gp.snr_index, k0, k1);
            const unsigned long long a = ((unsigned long long)r.c[0] << 32) | r.c[1];
            const unsigned long long b = ((unsigned long long)r.c[2] << 32) | r.c[3];
            const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);   // (0,1]
            const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);           // [0,1)
            const double rad = sqrt(-2.0 * log(u1));
            double sn, cs;
            sincospi(2.0 * u2, &sn, &cs);
            const int p0 = 2 * q, p1 = 2 * q + 1;
            stage[p0] = (pk_getbit<NW>(CW, p0) ? 1.0 : -1.0) + io.gp.sigma * (rad * cs);
            if (p1 < NE) stage[p1] = (pk_getbit<NW>(CW, p1) ? 1.0 : -1.0) + io.gp.sigma * (rad * sn);
        }
        __syncwarp();
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int p = lane + 32 * w;
            yv[w] = (p < NE) ? stage[p] : 0.0;
        }
        __syncwarp();
        if (dump) {
            if (io.d_info) {
                for (int i = lane; i < K; i += 32) io.d_info[f * K + i] = (uint8_t)((w_u[i >> 5] >> (i & 31)) & 1u);
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int p = lane + 32 * w;
                if (p < NE) {
                    if (io.d_cw) io.d_cw[f * NE + p] = (uint8_t)((CW[w] >> lane) & 1u);
                    if (io.d_y) io.d_y[f * NE + p] = yv[w];
                }
            }
        }
        __syncwarp();
    }
}

// finished frame -> outputs (replay: decisions + trial count; generation: compare with the sent word,
// dataForPlot.cpp:66-73) and the warp's running totals
template <int M, int NW, bool GEN>
__device__ __forceinline__ void pk_emit(const PkIo &io, long f, const uint32_t (&YH)[NW], const uint32_t (&bestF)[NW],
                                        const uint32_t (&CW)[NW], uint32_t trials, uint32_t ecmp, uint32_t esum,
                                        uint32_t flags, PkWarpTotals &tot, int ext, const double *alpha) {
    const int N = (1 << M) - 1 + ext;   // length of the (possibly extended) code
    const int lane = threadIdx.x & 31;
    uint32_t be = 0;
    if constexpr (!GEN) {
        if (!(flags & PK_FLAG_NO_DECISION)) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int p = lane + 32 * w;
                if (p < N) __stcs(io.decided + f * N + p, (uint8_t)(((YH[w] ^ bestF[w]) >> lane) & 1u));
            }
        }
        if (lane == 0 && io.trials) __stcs(io.trials + f, trials);
    } else {
        if (!(flags & PK_FLAG_NO_DECISION)) {
#pragma unroll
            for (int w = 0; w < NW; ++w) be += __popc(YH[w] ^ bestF[w] ^ CW[w]);
        } else {
#pragma unroll
            for (int w = 0; w < NW; ++w) be += __popc(CW[w]);   // undecided buffer counted as all-zero
        }
        flags |= be ? PK_FLAG_FRAME_ERROR : 0;
        if (be && !(flags & PK_FLAG_NO_DECISION)) {
            // the reference's DEBUG build logs the frames whose transmitted word is MORE likely than the decision
            // (calcL(res) < calcL(decoded), dataForPlot.cpp:55-64: a decoder that is not maximum-likelihood there)
            double l_tx = 0.0, l_dec = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                uint32_t a = YH[w] ^ CW[w], b = bestF[w];
                while (a) { const int i = __ffs(a) - 1; a &= a - 1; l_tx += alpha[32 * w + i]; }
                while (b) { const int i = __ffs(b) - 1; b &= b - 1; l_dec += alpha[32 * w + i]; }
            }
            if (l_tx < l_dec) flags |= PK_FLAG_NON_ML;
        }
    }
    if (lane == 0) {
        if (io.recs) {
            pk_frame_rec r;
            r.trials = trials; r.extra_cmp = ecmp; r.extra_sum = esum;
            r.bit_errors = (uint16_t)be; r.flags = (uint8_t)flags; r.reserved = 0;
            io.recs[f] = r;
        }
        tot.add(N, trials, ecmp, esum, flags, be);
    }
}

// ------------------------------------------------------------------ phase A
template <int M, int T, bool LUT, bool GEN, bool CT = false>
__global__ void __launch_bounds__(PK_WARPS_A * 32)
k_phase_a(PkDevTables tb, PkKanekoParams kp, PkIo io, long B, PkPhaseCtl *ctl, PkLongRec *longs, long long_cap) {
    typedef PkSmem<M, T, LUT, CT> SM;
    typedef KanekoWarp<M, T, LUT, CT> KW;
    constexpr int NW = KW::NW;
    extern __shared__ __align__(16) unsigned char smem[];
    pk_stage_tables<M, T, LUT, CT>(smem, tb, !CT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wb = smem + SM::tables_a(tb.nk) + (size_t)warp * SM::W_SZ_A;
    typename KW::WarpMem wm = KW::warp_mem(wb);
    uint32_t *w_u = reinterpret_cast<uint32_t *>(wb + SM::W_U);
    typename KW::Tables tabs;
    tabs.mul = smem + SM::MUL_OFF;
    tabs.xoff = reinterpret_cast<const uint16_t *>(smem + SM::XOFF_OFF);
    tabs.col = reinterpret_cast<const uint32_t *>(smem + (CT ? SM::B_COL_OFF : SM::COL_OFF));
    tabs.lut = reinterpret_cast<const uint16_t *>(smem + SM::LUT_OFF);
    if constexpr (CT) {   // class-table mode: the narrow search probes the class table as well
        tabs.ctlog = smem + SM::CT_LOG_OFF;
        tabs.ctbits = tb.ct_bits;
        tabs.cttex = tb.ct_tex;
        tabs.cthash = tb.ct_hash;
        tabs.cthshift = tb.ct_hshift;
        tabs.cthmask = tb.ct_hmask;
#pragma unroll
        for (int i = 0; i < 8; ++i) tabs.ctmult[i] = tb.ct_mult[i];
    }
    // frames are parked only when a wide kernel exists for this code (long_cap > 0 says so)
    const uint32_t limit = (long_cap > 0) ? kp.limit_a : 0xFFFFFFFFu;

    PkWarpTotals tot;
    tot.clear();
    const int grab = kp.frames_per_grab;
    for (;;) {
        unsigned long long f0 = 0;
        if (lane == 0) f0 = atomicAdd(&ctl->queue_a, (unsigned long long)grab);
        f0 = __shfl_sync(PK_FULL, f0, 0);
        if ((long)f0 >= B) break;
        const long f1 = ((long)f0 + grab < B) ? (long)f0 + grab : B;
        for (long f = (long)f0; f < f1; ++f) {
            double yv[NW];
            uint32_t CW[NW];
            pk_load_frame<M, NW, GEN>(io, tb, f, wm.skey, w_u, true, yv, CW, kp.ext);
            if (GEN && io.dump_only) continue;
            typename KW::Frame fr;
            typename KW::Search s;
            KW::setup(tabs, wm, yv, kp, fr);
            KW::search_init(s, fr, kp.variant);
            uint32_t next = 0;
            bool done = KW::narrow(tabs, wm, kp, fr, s, 0u, limit, &next);
            if (!done) {
                // park: frames with a long pattern range left go to the "big" end of the list (one CTA each in
                // phase B), the others to the "small" end (one warp each)
                long slot = -1;
                if (lane == 0) {
                    if ((long)atomicAdd(&ctl->n_total, 1ull) < long_cap) {
                        const uint32_t span = (s.bound > next) ? s.bound - next : 0u;
                        // "huge" frames (uncapped searches: up to 2^31 patterns) get their own short list behind the
                        // main one and are always searched by a whole CTA, however many big frames there are
                        if (s.have && span >= kp.huge_span) {   // (no decision yet: the bound is still the initial one, not a measure of the work left)
                            const unsigned long long h = atomicAdd(&ctl->n_huge, 1ull);
                            if (h < (unsigned long long)PK_HUGE_CAP) slot = long_cap + (long)h;
                        }
                        if (slot < 0)
                            slot = (span >= kp.big_span) ? long_cap - 1 - (long)atomicAdd(&ctl->n_big, 1ull)
                                                         : (long)atomicAdd(&ctl->n_long, 1ull);
                        KW::park(s, (uint32_t)f, next, longs + slot);
                    }
                }
                slot = __shfl_sync(PK_FULL, slot, 0);
                if (slot >= 0) continue;
                // list full: finish here (phase B ignores slots >= long_cap)
                KW::narrow(tabs, wm, kp, fr, s, next, 0xFFFFFFFFu, &next);
            }
            KW::search_finish(s);
            pk_emit<M, NW, GEN>(io, f, fr.YH, s.bestF, CW, s.trials, s.tsteps + s.nimpr + kp.extra_ops, s.tsteps + kp.extra_ops, s.flags, tot, kp.ext, wm.alpha);
        }
    }
    if (lane == 0) tot.flush(io.totals);
}

// ------------------------------------------------------------------ phase B
__device__ __forceinline__ unsigned int pk_ld_vol(const unsigned int *p) { return *reinterpret_cast<const volatile unsigned int *>(p); }
__device__ __forceinline__ unsigned long long pk_ld_vol(const unsigned long long *p) { return *reinterpret_cast<const volatile unsigned long long *>(p); }
__device__ __forceinline__ void pk_st_vol(unsigned int *p, unsigned int v) { *reinterpret_cast<volatile unsigned int *>(p) = v; }

// Work order inside one launch (every CTA walks the same stages; all queues are global atomics):
//   1. parked HUGE frames and the cooperative share of the big ones: one CTA per frame (`master`), warp w takes block w
//      of every 1024 * WB-pattern step, commits in pattern order through shared memory;
//   2. the other parked frames: one warp per frame.  A one-warp search still running after kp.solo_patterns is handed
//      to a whole CTA through the LATE list;
//   3. end game: CTAs without work take late frames as masters, and otherwise HELP the masters that are still searching.
// MEGA frames.  A master whose search has kp.mega_span patterns or more ahead of it (an uncapped search after an early
// decision: up to 2^28 patterns; a large code before its first decodable pattern: up to 2^31) registers the frame in a
// mega slot.  Idle CTAs then scan chunks of kp.mega_chunk patterns AHEAD of the master in `dry` mode against a snapshot
// of (l0, best codeword) and set a bit per chunk that holds no possible improvement; the master skips those chunks and
// searches every other chunk itself, so commits stay in pattern order and the master never waits for anybody.  A clean
// verdict cannot go stale: l0 only decreases and a former best codeword never becomes an improvement again; the
// bound, m0 and the counters are the master's business alone.  (Before this, one CTA ground through such a frame at
// 16K patterns per 11 us while the other 147 SMs idled: the tail of every uncapped low-SNR launch.)
template <int M, int T, bool LUT, bool GEN, bool CT = false>
__global__ void __launch_bounds__(PkSmem<M, T, LUT, CT>::WB * 32, CT ? PkTraits<M, T>::MINB_CT : LUT ? 1 : PkTraits<M, T>::MINB)   // (coset-table mode: 16 warps, one CTA per SM next to its 64 KB table -- 128 registers, no spills)
k_phase_b(PkDevTables tb, PkKanekoParams kp, PkIo io, PkPhaseCtl *ctl, PkLongRec *longs, long long_cap) {
    typedef PkSmem<M, T, LUT, CT> SM;
    typedef KanekoWarp<M, T, LUT, CT> KW;
    constexpr int NW = KW::NW;
    constexpr int GW = SM::WB;
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned long long n_long = ctl->n_long, n_big = ctl->n_big;
    const unsigned long long n_huge = ctl->n_huge < (unsigned long long)PK_HUGE_CAP ? ctl->n_huge : (unsigned long long)PK_HUGE_CAP;
    if (n_long + n_big + n_huge == 0) return;
    pk_stage_tables<M, T, LUT, CT>(smem, tb, false);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wb = smem + SM::tables_b(tb.nk) + (size_t)warp * SM::W_SZ_B;
    typename KW::WarpMem wm = KW::warp_mem(wb);
    uint32_t *w_u = reinterpret_cast<uint32_t *>(wb + SM::W_U);
    typename KW::Tables tabs;
    tabs.mul = nullptr;    // the wide search is bit-sliced or table-driven by the coset table only
    tabs.xoff = nullptr;
    tabs.col = reinterpret_cast<const uint32_t *>(smem + SM::B_COL_OFF);
    tabs.lut = reinterpret_cast<const uint16_t *>(smem + SM::LUT_OFF);
    if constexpr (CT) {
        tabs.ctlog = smem + SM::CT_LOG_OFF;
        tabs.ctbits = tb.ct_bits;
        tabs.cttex = tb.ct_tex;
        tabs.cthash = tb.ct_hash;
        tabs.cthshift = tb.ct_hshift;
        tabs.cthmask = tb.ct_hmask;
#pragma unroll
        for (int i = 0; i < 8; ++i) tabs.ctmult[i] = tb.ct_mult[i];
    }
    if constexpr (SM::BSM)   // root words of the bit-sliced Chien search go to a per-warp slice of global scratch
        wm.z = io.zscratch + ((size_t)blockIdx.x * SM::WB + warp) * (size_t)KW::N * 32;
    __shared__ typename KW::Search s_shared;
    __shared__ unsigned long long s_idx;
    __shared__ uint32_t s_votes[2 * SM::WB];
    __shared__ uint32_t s_cmd[4];
    __shared__ double s_snap_l0;
    __shared__ uint32_t s_snap[NW + 1];
    PkLongRec *late = longs + long_cap + PK_HUGE_CAP;
    const long late_cap = long_cap;   // every parked frame of the launch may be handed over
    const uint32_t CH = (kp.mega_chunk % (1024u * GW)) ? 0u : kp.mega_chunk;   // 0: no mega frames (chunks must be whole cooperative steps)

    PkWarpTotals tot;
    tot.clear();
    // frame tables of the frame this CTA has set up last (every warp keeps its own copy)
    typename KW::Frame fr;
    uint32_t CW[NW];
    long cur_frame = -1;
    auto load = [&](long f) {
        double yv[NW];
        pk_load_frame<M, NW, GEN>(io, tb, f, wm.skey, w_u, false, yv, CW, kp.ext);   // dumps were written in phase A
        KW::setup(tabs, wm, yv, kp, fr);
        cur_frame = f;
    };

    // Stage 1 / 3 share ONE call site of the cooperative search (the CTA-level loop below is a small state machine:
    // several inlined copies of wide<GW> would triple the kernel's code and stack).
    // master state (CTA-uniform)
    bool m_active = false, m_late = false, m_may_open = false;
    long m_f = 0;
    PkMegaSlot *m_ms = nullptr;
    uint32_t *m_bits = nullptr;
    uint32_t m_limit = 0, m_cbase = 0, m_gen = 0, m_own = 0, m_start = 0, m_g0 = 0;
    typename KW::Search s;   // the master's sequential state, or a helper's snapshot
    // parked frames of this launch
    const bool all_coop = (n_long + n_big + n_huge) <= (unsigned long long)gridDim.x;
    // Cooperation costs ~10 % throughput (hand-over barriers), so when there are far more big frames than CTAs only one
    // round of them is searched cooperatively and the rest go warp-per-frame; with few big frames (the tail regime of
    // medium / high SNR launches) all of them are, and with fewer parked frames than CTAs the small ones as well.
    const unsigned long long n_coop = all_coop ? n_big + n_long : (n_big > 8ull * gridDim.x) ? (unsigned long long)gridDim.x : n_big;
    const unsigned long long n_solo = all_coop ? 0ull : n_long + (n_big - n_coop);
    int stage = 1;

    auto begin_master = [&](const PkLongRec *rec, bool known_long, bool from_late) {
        m_f = (long)rec->frame;
        load(m_f);
        KW::unpark(s, rec);
        m_g0 = rec->base & ~1023u;   // the search's step grid (and chunk grid) starts here
        m_start = rec->base;
        m_ms = nullptr; m_bits = nullptr; m_limit = 0; m_cbase = 0; m_gen = 0;
        m_may_open = CH != 0;
        m_own = known_long ? 0u : kp.mega_after;
        m_late = from_late;
        m_active = true;
    };
    // (re)open a window of PK_MEGA_WINDOW chunks at the master's position: fields + tagged bitmap words, then the state word
    auto open_window = [&]() {
        if (threadIdx.x == 0) pk_st_vol(&m_ms->state, 0u);
        __threadfence();
        __syncthreads();
        m_gen = (m_gen + 1u) ? (m_gen + 1u) : 1u;
        m_cbase = (m_start - m_g0) / CH;
        m_limit = m_cbase + PK_MEGA_WINDOW;
        for (uint32_t w = threadIdx.x; w < PK_MEGA_WORDS; w += blockDim.x) m_bits[w] = m_gen << 16;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            m_ms->gen = m_gen;
            m_ms->frame = (unsigned int)m_f; m_ms->g0 = m_g0; m_ms->cbase = m_cbase; m_ms->limit = m_limit;
            m_ms->pos = m_cbase; m_ms->next = m_cbase + 1; m_ms->bound = s.bound;
            m_ms->l0 = s.l0; m_ms->have = s.have ? 1u : 0u;
#pragma unroll
            for (int w = 0; w < NW; ++w) m_ms->bestF[w] = s.bestF[w];
            m_ms->seq = 0;
            __threadfence();
            pk_st_vol(&m_ms->state, m_gen);
            const uint32_t slot = (uint32_t)(m_ms - io.mega);
            atomicOr(&ctl->mega_mask[slot >> 6], 1ull << (slot & 63u));
        }
    };
    auto end_master = [&]() {
        if (m_ms && threadIdx.x == 0) {
            const uint32_t slot = (uint32_t)(m_ms - io.mega);
            atomicAnd(&ctl->mega_mask[slot >> 6], ~(1ull << (slot & 63u)));
            pk_st_vol(&m_ms->state, 0u);
            __threadfence();
            atomicExch(&m_ms->owner, 0u);   // the slot is free again (its generation counter lives on in m_ms->gen)
        }
        if (warp == 0) {
            KW::search_finish(s);
            pk_emit<M, NW, GEN>(io, m_f, fr.YH, s.bestF, CW, s.trials, s.tsteps + s.nimpr + kp.extra_ops, s.tsteps + kp.extra_ops, s.flags, tot, kp.ext, wm.alpha);
        }
        if (m_late) {
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(&ctl->masters, ~0ull);
        }
        m_active = false;
    };

    for (;;) {
        uint32_t h_slot = 0, h_chunk = 0, h_gen = 0;
        bool h_job = false;
        // ---------------- A: something to do
        if (!m_active && stage == 1) {
            // huge frames (own list behind the main one) and the cooperative share of the big ones: one CTA per frame
            __syncthreads();
            if (threadIdx.x == 0) s_idx = atomicAdd(&ctl->queue_big, 1ull);
            __syncthreads();
            const unsigned long long si = s_idx;
            if (si < n_huge + n_coop) {
                const unsigned long long idx = si - (si < n_huge ? 0ull : n_huge);
                const PkLongRec *rec = (si < n_huge) ? longs + (long_cap + (long)idx)
                                       : (idx < n_big) ? longs + (long_cap - 1 - (long)idx) : longs + (idx - n_big);
                begin_master(rec, si < n_huge, false);
            } else {
                stage = 2;
            }
        }
        if (!m_active && stage == 2) {
            // the other parked frames (big leftovers first), one warp each
            for (;;) {
                unsigned long long idx = 0;
                if (lane == 0) idx = atomicAdd(&ctl->queue_b, 1ull);
                idx = __shfl_sync(PK_FULL, idx, 0);
                if (idx >= n_solo) break;
                const PkLongRec *rec = (idx < n_big - n_coop) ? longs + (long_cap - 1 - (long)(n_coop + idx)) : longs + (idx - (n_big - n_coop));
                const long f = (long)rec->frame;
                double yv[NW];
                pk_load_frame<M, NW, GEN>(io, tb, f, wm.skey, w_u, false, yv, CW, kp.ext);
                KW::setup(tabs, wm, yv, kp, fr);
                KW::unpark(s, rec);
                uint32_t start = rec->base;
                const unsigned long long st = (unsigned long long)(rec->base & ~1023u) + kp.solo_patterns;
                uint32_t stop = (kp.solo_patterns && st < 0x80000000ull) ? (uint32_t)st : 0xFFFFFFFFu;
                bool handed = false;
                for (;;) {
                    const int r = KW::template wide<1>(tabs, wm, kp, fr, s, start, nullptr, nullptr, 0, stop, false);
                    if (r != KW::PK_W_STOPPED) break;
                    // still running: hand the search to a whole CTA (and through it to the idle part of the grid)
                    long slot = -1;
                    if (lane == 0) {
                        const unsigned long long q = atomicAdd(&ctl->n_late, 1ull);
                        if (q < (unsigned long long)late_cap) {
                            slot = (long)q;
                            KW::park(s, (uint32_t)f, stop, late + slot);
                            __threadfence();
                            pk_st_vol(&late[slot].pad, kp.epoch);   // the record is complete
                        }
                    }
                    slot = __shfl_sync(PK_FULL, slot, 0);
                    if (slot >= 0) { handed = true; break; }
                    start = stop;            // late list full: finish here
                    stop = 0xFFFFFFFFu;
                }
                if (handed) continue;
                KW::search_finish(s);
                pk_emit<M, NW, GEN>(io, f, fr.YH, s.bestF, CW, s.trials, s.tsteps + s.nimpr + kp.extra_ops, s.tsteps + kp.extra_ops, s.flags, tot, kp.ext, wm.alpha);
            }
            cur_frame = -1;   // the warps of this CTA hold different frames now
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(&ctl->ctas_past_solo, 1ull);
            stage = 3;
        }
        if (!m_active && stage == 3) {
            // end game: late frames as a master, else help a master, else wait for the masters or leave
            if (threadIdx.x == 0) {
                uint32_t cmd = 0, a0 = 0, a1 = 0, a2 = 0;   // 0 = look again, 1 = master of late frame a0, 2 = help slot a0 (generation a2) with chunk a1, 3 = leave
                for (;;) {
                    const unsigned long long nl = pk_ld_vol(&ctl->n_late), ql = pk_ld_vol(&ctl->queue_late);
                    if (ql >= (nl < (unsigned long long)late_cap ? nl : (unsigned long long)late_cap)) break;
                    atomicAdd(&ctl->masters, 1ull);
                    if (atomicCAS(&ctl->queue_late, ql, ql + 1ull) == ql) { cmd = 1; a0 = (uint32_t)ql; break; }
                    atomicAdd(&ctl->masters, ~0ull);
                }
                if (!cmd && CH) {
                    // slots with an open window (a bit mask: scanning all the slot records costs a helper ~100 us)
                    unsigned long long act[PK_MEGA_SLOTS / 64];
#pragma unroll
                    for (int w = 0; w < PK_MEGA_SLOTS / 64; ++w) act[w] = pk_ld_vol(&ctl->mega_mask[w]);
                    for (uint32_t i = 0; i < (uint32_t)PK_MEGA_SLOTS && !cmd; ++i) {
                        if (!((act[i >> 6] >> (i & 63u)) & 1ull)) {
                            if (!(act[i >> 6] >> (i & 63u))) i |= 63u;   // nothing left in this word
                            continue;
                        }
                        PkMegaSlot *ms = io.mega + i;
                        const uint32_t st = pk_ld_vol(&ms->state);
                        if (!st) continue;
                        __threadfence();
                        const uint32_t lim = pk_ld_vol(&ms->limit), g0 = pk_ld_vol(&ms->g0);
                        const uint32_t nx = pk_ld_vol(&ms->next);
                        if (nx >= lim || (unsigned long long)g0 + (unsigned long long)nx * CH >= (unsigned long long)pk_ld_vol(&ms->bound)) continue;
                        const uint32_t c = atomicAdd(&ms->next, 1u);
                        __threadfence();
                        if (pk_ld_vol(&ms->state) != st) continue;   // the window moved on meanwhile
                        if (c >= lim || (unsigned long long)g0 + (unsigned long long)c * CH >= (unsigned long long)pk_ld_vol(&ms->bound) || c <= pk_ld_vol(&ms->pos)) continue;
                        cmd = 2; a0 = i; a1 = c; a2 = st;
                    }
                }
                if (!cmd) {
                    // leave when nobody is a master any more and nobody can become one: all CTAs are past the one-warp
                    // loop (n_late is final), every late frame has been taken and is finished
                    bool leave = pk_ld_vol(&ctl->ctas_past_solo) >= (unsigned long long)gridDim.x;
                    __threadfence();
                    if (leave) {
                        const unsigned long long nl = pk_ld_vol(&ctl->n_late);
                        leave = pk_ld_vol(&ctl->queue_late) >= (nl < (unsigned long long)late_cap ? nl : (unsigned long long)late_cap);
                        __threadfence();
                        leave = leave && pk_ld_vol(&ctl->masters) == 0ull;
                    }
                    if (leave) cmd = 3;
                    else __nanosleep(400);
                }
                s_cmd[0] = cmd; s_cmd[1] = a0; s_cmd[2] = a1; s_cmd[3] = a2;
            }
            __syncthreads();
            const uint32_t cmd = s_cmd[0], a0 = s_cmd[1], a1 = s_cmd[2], a2 = s_cmd[3];
            __syncthreads();
            if (cmd == 3) break;
            if (cmd == 0) continue;
            if (cmd == 1) {
                PkLongRec *rec = late + a0;
                if (threadIdx.x == 0)
                    while (pk_ld_vol(&rec->pad) != kp.epoch) __nanosleep(100);   // the one-warp search is still writing it
                __threadfence();
                __syncthreads();
                begin_master(rec, true, true);
            } else {
                h_job = true; h_slot = a0; h_chunk = a1; h_gen = a2;
            }
        }
        // ---------------- B: the next stretch of patterns to run
        uint32_t start, stop, h_cbase = 0;
        bool dry = false;
        uint32_t nimpr0 = 0;
        if (m_active) {
            if (!m_ms) {
                // plain cooperative search up to the point where it is opened to helpers (if ever)
                stop = 0xFFFFFFFFu;
                if (m_may_open) {
                    const unsigned long long c = ((unsigned long long)(m_start - m_g0) + m_own + CH - 1) / CH;
                    const unsigned long long st = (unsigned long long)m_g0 + c * CH;
                    if (st < 0x80000000ull) stop = (uint32_t)st;
                }
            } else {
                const uint32_t c = (m_start - m_g0) / CH;
                if (c >= m_limit) { open_window(); continue; }   // end of the window: the next one starts here
                if (threadIdx.x == 0) {
                    // chunks marked clean by helpers, from c on, as far as the window and the bound reach
                    pk_st_vol(&m_ms->pos, c);
                    atomicMax(&m_ms->next, c + 1);
                    const unsigned long long nch = ((unsigned long long)(s.bound - m_g0) + CH - 1) / CH;
                    const uint32_t cend = nch < (unsigned long long)m_limit ? (uint32_t)nch : m_limit;
                    uint32_t cc = c;
                    while (cc < cend) {
                        const uint32_t rel = cc - m_cbase, k = rel & 15u;
                        const uint32_t rem = (~(pk_ld_vol(m_bits + (rel >> 4)) >> k)) & 0xFFFFu;   // first zero flag ends the run (the k vacated top flags read as "not clean")
                        uint32_t run = rem ? (uint32_t)(__ffs(rem) - 1) : 16u;
                        const bool word_end = run >= 16u - k;
                        if (run > cend - cc) run = cend - cc;
                        cc += run;
                        if (!word_end) break;
                    }
                    s_cmd[0] = cc - c;
                }
                __syncthreads();
                const uint32_t skip = s_cmd[0];
                __syncthreads();
                if (skip) {
                    // no pattern of these chunks improves on the state: the sequential loop just runs through them
                    const unsigned long long st = (unsigned long long)m_start + (unsigned long long)skip * CH;
                    if (st >= (unsigned long long)s.bound) { s.trials = s.bound; end_master(); continue; }
                    m_start = (uint32_t)st;
                    if (m_start >= kp.max_trials) { s.trials = m_start; s.flags |= PK_FLAG_TRUNCATED; end_master(); }
                    continue;
                }
                const unsigned long long st = (unsigned long long)m_start + CH;
                stop = st < 0x80000000ull ? (uint32_t)st : 0xFFFFFFFFu;
            }
            start = m_start;
            nimpr0 = s.nimpr;
        } else if (h_job) {
            // a helper's share: chunk h_chunk of mega slot h_slot against a snapshot of the master's state, nothing committed
            PkMegaSlot *ms = io.mega + h_slot;
            if (threadIdx.x == 0) {
                // frame, grid origin and the snapshot, all of generation h_gen (else the job is void)
                uint32_t ok = 1;
                for (;;) {
                    const unsigned int q0 = pk_ld_vol(&ms->seq);
                    if (pk_ld_vol(&ms->state) != h_gen) { ok = 0; break; }
                    if (q0 & 1u) continue;
                    __threadfence();
                    s_snap_l0 = *reinterpret_cast<volatile double *>(&ms->l0);
                    s_snap[NW] = pk_ld_vol(&ms->have);
#pragma unroll
                    for (int w = 0; w < NW; ++w) s_snap[w] = pk_ld_vol(&ms->bestF[w]);
                    s_cmd[1] = pk_ld_vol(&ms->frame);
                    s_cmd[2] = pk_ld_vol(&ms->g0);
                    s_cmd[3] = pk_ld_vol(&ms->cbase);
                    __threadfence();
                    if (pk_ld_vol(&ms->seq) == q0) { ok = pk_ld_vol(&ms->state) == h_gen; break; }
                }
                s_cmd[0] = ok;
            }
            __syncthreads();
            const uint32_t h_ok = s_cmd[0], h_frame = s_cmd[1], h_g0 = s_cmd[2];
            h_cbase = s_cmd[3];
            __syncthreads();
            if (!h_ok) continue;
            if (cur_frame != (long)h_frame) load((long)h_frame);
            s.l0 = s_snap_l0;
            s.have = s_snap[NW] != 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s.bestF[w] = s_snap[w];
            s.m0 = 0; s.first_ok = true; s.early = false;
            s.bound = 0x7FFFFFFFu;
            s.trials = 0; s.tsteps = 0; s.nimpr = 0; s.step_last = 0; s.flags = 0;
            __syncthreads();
            start = h_g0 + h_chunk * CH;
            stop = start + CH;
            dry = true;
        } else {
            continue;
        }
        // ---------------- C: run it (the only call site of the cooperative search)
        const int r = KW::template wide<GW>(tabs, wm, kp, fr, s, start, &s_shared, s_votes, warp, stop, dry);
        // ---------------- D
        if (dry) {
            if (r == KW::PK_W_STOPPED && threadIdx.x == 0) {
                // mark the chunk clean -- only in a word that still carries the generation the scan was made for
                const uint32_t rel = h_chunk - h_cbase;
                unsigned int *word = io.mega_bits + (size_t)h_slot * PK_MEGA_WORDS + (rel >> 4);
                unsigned int old = pk_ld_vol(word);
                while ((old >> 16) == (h_gen & 0xFFFFu)) {
                    const unsigned int seen = atomicCAS(word, old, old | (1u << (rel & 15u)));
                    if (seen == old) break;
                    old = seen;
                }
            }
            continue;
        }
        if (r == KW::PK_W_FINISHED) { end_master(); continue; }
        m_start = stop;
        if (!m_ms) {
            // still running after its own share: open the search to the idle CTAs of the grid if it is long
            if (s.bound > m_start && s.bound - m_start >= kp.mega_span) {
                if (threadIdx.x == 0) {
                    uint32_t got = 0xFFFFFFFFu;
                    const uint32_t h0 = (blockIdx.x * 7u) % (uint32_t)PK_MEGA_SLOTS;   // spread the probes
                    for (uint32_t q = 0; q < (uint32_t)PK_MEGA_SLOTS; ++q) {
                        const uint32_t i = (h0 + q) % (uint32_t)PK_MEGA_SLOTS;
                        if (pk_ld_vol(&io.mega[i].owner) == 0u && atomicCAS(&io.mega[i].owner, 0u, 1u) == 0u) { got = i; break; }
                    }
                    s_cmd[0] = got;
                    if (got != 0xFFFFFFFFu) s_cmd[1] = pk_ld_vol(&io.mega[got].gen);
                }
                __syncthreads();
                const uint32_t slot = s_cmd[0], gen0 = s_cmd[1];
                __syncthreads();
                if (slot != 0xFFFFFFFFu) {
                    m_ms = io.mega + slot;
                    m_bits = io.mega_bits + (size_t)slot * PK_MEGA_WORDS;
                    m_gen = gen0;        // generations of a slot go on where its last owner stopped
                    m_may_open = false;
                    open_window();
                } else {
                    m_own = 8u * CH;     // every slot is taken: go on alone, ask again later
                }
            } else {
                m_may_open = false;
            }
        } else if (s.nimpr != nimpr0 && threadIdx.x == 0) {
            // publish the new state for the helpers (sequence lock)
            const unsigned int q = m_ms->seq;
            pk_st_vol(&m_ms->seq, q + 1);
            __threadfence();
            m_ms->l0 = s.l0; m_ms->have = s.have ? 1u : 0u; m_ms->bound = s.bound;
#pragma unroll
            for (int w = 0; w < NW; ++w) m_ms->bestF[w] = s.bestF[w];
            __threadfence();
            pk_st_vol(&m_ms->seq, q + 2);
        }
    }
    if (lane == 0) tot.flush(io.totals);
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
    pk_stage_tables<M, T, false>(smem, tb, true);
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
    typedef PkTraits<M, T> TR;

    static bool host_alg(const uint32_t *Sw, const uint8_t *mul, const uint16_t *xoff, uint32_t *A) {
        return pk_alg_decode<M, T>(Sw, mul, xoff, A);
    }

    // One shared-memory / L1 split for every Kaneko kernel: kernels of different handles (codes) run concurrently on the
    // same SMs, and a CTA that becomes resident next to CTAs of a kernel with a larger carveout inherits that carveout
    // for its whole (persistent) life.  The class-table search needs >= 64 KB of L1 to keep its bitmap gathers in flight
    // (1.7 ms vs 3.1 ms per launch with 32 KB), so every kernel asks for the 164 KB configuration and sizes its grid for
    // it; only kernels whose single CTA does not fit take the maximum.
    // big_smem: the bit-sliced wide search of the large codes keeps its Berlekamp-Massey state in shared memory (20-32 KB per
    // warp) and gathers nothing from global memory: it takes the whole 227 KB so that a second CTA fits next to the first
    // (BCH(127,64,21): 4 -> 8 warps per SM; with one warp per scheduler the ALU pipe idles on every dependent LOP3).
    template <class K>
    static cudaError_t fit(K kernel, int threads, size_t dyn, int sm_count, PkLaunchGeom *g, bool big_smem = false) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, kernel);
        if (e != cudaSuccess) return e;
        const size_t budget = (size_t)(big_smem ? 227 : 164) << 10, per_cta = dyn + fa.sharedSizeBytes + 1024;
        const bool fits = per_cta <= budget;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (fits && !big_smem) ? 70 : 100);
        if (e != cudaSuccess) return e;
        int per = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, threads, dyn);
        if (e != cudaSuccess) return e;
        if (fits && per > (int)(budget / per_cta)) per = (int)(budget / per_cta);
        if (per < 1) return cudaErrorLaunchOutOfResources;
        g->grid = sm_count * per;      // persistent: every resident slot of every SM
        g->block = threads;
        g->smem = dyn;
        return cudaSuccess;
    }
    template <bool LUT, bool GEN, bool CT>
    static cudaError_t geom_one(int nk, int sm_count, PkLaunchGeom *ga, PkLaunchGeom *gb) {
        cudaError_t e = fit(k_phase_a<M, T, LUT, GEN, CT>, PK_WARPS_A * 32, PkSmem<M, T, LUT, CT>::total_a(nk), sm_count, ga);
        if (e != cudaSuccess) return e;
        return fit(k_phase_b<M, T, LUT, GEN, CT>, PkSmem<M, T, LUT, CT>::WB * 32, PkSmem<M, T, LUT, CT>::total_b(nk), sm_count, gb,
                   PkSmem<M, T, LUT, CT>::BSM);
    }
    // out[0..1] = replay phase A / B, out[2..3] = generation phase A / B
    static cudaError_t geom_kaneko(int mode, int nk, int sm_count, PkLaunchGeom *out) {
        cudaError_t e;
        if (mode == PK_MODE_LUT) {
            if constexpr (TR::LUT_OK) {
                e = geom_one<true, false, false>(nk, sm_count, out + 0, out + 1);
                if (e != cudaSuccess) return e;
                return geom_one<true, true, false>(nk, sm_count, out + 2, out + 3);
            }
            return cudaErrorInvalidValue;
        }
        if (mode == PK_MODE_CLASS) {
            if constexpr (TR::CT_OK) {
                e = geom_one<false, false, true>(nk, sm_count, out + 0, out + 1);
                if (e != cudaSuccess) return e;
                return geom_one<false, true, true>(nk, sm_count, out + 2, out + 3);
            }
            return cudaErrorInvalidValue;
        }
        e = geom_one<false, false, false>(nk, sm_count, out + 0, out + 1);
        if (e != cudaSuccess) return e;
        return geom_one<false, true, false>(nk, sm_count, out + 2, out + 3);
    }
    static cudaError_t geom_bdd(int sm_count, PkLaunchGeom *out) {
        const size_t smem = PkSmem<M, T, false>::tables(0);
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

    template <bool LUT, bool GEN, bool CT>
    static cudaError_t run(const PkLaunchGeom *g, const PkDevTables &tb, const PkKanekoParams &kp, const PkIo &io, long B,
                           PkPhaseCtl *ctl, PkLongRec *longs, long long_cap, cudaStream_t st) {
        cudaError_t e = cudaMemsetAsync(ctl, 0, sizeof(PkPhaseBlock), st);   // ctl is the head of a PkPhaseBlock (control words + mega slots)
        if (e != cudaSuccess) return e;
        const bool wide_ok = g[1].grid > 0 && long_cap > 0 && !(GEN && io.dump_only);
        k_phase_a<M, T, LUT, GEN, CT><<<g[0].grid, g[0].block, g[0].smem, st>>>(tb, kp, io, B, ctl, longs, wide_ok ? long_cap : 0);
        ++g_pk_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (wide_ok) {
            k_phase_b<M, T, LUT, GEN, CT><<<g[1].grid, g[1].block, g[1].smem, st>>>(tb, kp, io, ctl, longs, long_cap);
            ++g_pk_launches;
            e = cudaGetLastError();
        }
        return e;
    }
    static cudaError_t kaneko(int mode, bool gen, const PkLaunchGeom *g4, const PkDevTables &tb, const PkKanekoParams &kp,
                              const PkIo &io, long B, PkPhaseCtl *ctl, PkLongRec *longs, long long_cap, cudaStream_t st) {
        const PkLaunchGeom *g = g4 + (gen ? 2 : 0);
        if (mode == PK_MODE_LUT) {
            if constexpr (TR::LUT_OK)
                return gen ? run<true, true, false>(g, tb, kp, io, B, ctl, longs, long_cap, st)
                           : run<true, false, false>(g, tb, kp, io, B, ctl, longs, long_cap, st);
            return cudaErrorInvalidValue;
        }
        if (mode == PK_MODE_CLASS) {
            if constexpr (TR::CT_OK)
                return gen ? run<false, true, true>(g, tb, kp, io, B, ctl, longs, long_cap, st)
                           : run<false, false, true>(g, tb, kp, io, B, ctl, longs, long_cap, st);
            return cudaErrorInvalidValue;
        }
        return gen ? run<false, true, false>(g, tb, kp, io, B, ctl, longs, long_cap, st)
                   : run<false, false, false>(g, tb, kp, io, B, ctl, longs, long_cap, st);
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
        return PkKernelSet{M, T, !TR::LUT_OK, TR::CT_OK, &host_alg, &geom_kaneko, &geom_bdd, &kaneko, &bdd, &encode};
    }
};
