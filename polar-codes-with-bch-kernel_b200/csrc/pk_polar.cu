// pk_polar.cu -- sm_100a kernels and C ABI for polar codes with binary matrix (e.g. extended-BCH) kernels.
//
// Mapping (reference file:line -> here):
//   CMixedKernelEncoder::Encode               out/external/MixedKernelEncoder.cpp:142-176  -> k_polar_encode
//   CMatrixBinaryKernel::Multiply / MatrixMultiply   Kernel.cpp:202-208, LinAlg.cpp:685-711 -> kernel_multiply()
//   CTrellisKernelProcessor::GetLLRs          TrellisKernelProcessor.cpp:234-295          -> get_llrs() / viterbi()
//   CListKernelEngine::IterativelyCalcS / IterativelyUpdateC   KernelListEngine.cpp:370-447,266-315 -> calc_s() / update_c()
//   CMixedKernelListDecoder::Decode / ContinuePathsFrozen / ContinuePathsUnfrozen
//                                             MixedKernelListDecoder.cpp:211-269,61-98,100-185 -> k_polar_lanes, k_polar_decode
//   CTVMemoryEngine path stack (Pop / Push, lazily initialised)   TVMemoryEngine.cpp:85-141, misc.h:212-226 -> PathStack
//
// Two decoders share this file's set-up and C ABI:
//   k_polar_lanes (pk_polar_lanes.cuh) -- list paths AND frames across the lanes of a warp, in-place Viterbi on the
//     fixed-bit-position numbering of the kernel trellises, bits packed across slots.  Runs for every list size <= 32
//     whenever the trellises fit the shared memory of an SM (every code of the tests).
//   k_polar_decode (below) -- one CTA per frame group, ONE WARP PER LIST PATH, lanes across trellis states in gather
//     form; the general fall-back (any L <= 32) and the decoder the lanes kernel is tested against.
// In both, every value is a single fp32 add or min of the reference's operands in the reference's order, so kernel
// LLRs and path metrics are bit-identical to the reference's floats (tests/test_gpu_polar.py).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/pk_capi.h"
#include "pk_philox.cuh"
#include "pk_polar.h"

extern unsigned long long g_pk_launches;
int pk_set_error(int code, const std::string &msg);   // pk_capi.cu

#define PKP_FULL 0xFFFFFFFFu
#define PKP_UPPER 100000.0f   // MTYPE_UPPER_BOUND (SeqConfigOrig.h:173): LLR of a shortened symbol

struct PkPolarKernelDev {
    const uint8_t *mat;    // [l][l]
    const uint8_t *ab;     // [l][l+1]
    const uint32_t *pred;
    const uint32_t *off;   // [l*l]
    int size, max_ab, npred;
};
struct PkPolarDev {
    int N, K, N0, layers, nw;
    int ksize[PK_POLAR_MAX_LAYERS];
    int outer[PK_POLAR_MAX_LAYERS + 1];
    PkPolarKernelDev kern[PK_POLAR_MAX_LAYERS];
    const uint8_t *frozen;     // [N0] 1 = frozen symbol
    const uint32_t *cmask;     // [N0][nw]
    const uint16_t *info_pos;  // [K]
    const uint8_t *symtype;    // [N0] or null
    const uint16_t *compact;   // [N0] index of symbol i among the transmitted ones (shortened / punctured codes), or null
    int all_static;            // every freezing constraint is "symbol = 0"
    int max_ab;                // over layers
};

// floats of Viterbi scratch per warp: 2 * 2^max_ab for lanes-across-states, 32 slices of 33 for lane-per-element
__host__ __device__ inline int polar_met_floats(int max_ab) { return (2 << max_ab) + 2 > 1120 ? (2 << max_ab) + 2 : 1120; }

// ------------------------------------------------------------------ shared building blocks
// dest[c*stride + i] (^)= XOR_r K[r][c] * src[r*stride + i]   (MatrixMultiply, LinAlg.cpp:685-711), one warp
__device__ __forceinline__ void kernel_multiply(const PkPolarKernelDev &k, int stride, const uint8_t *src, uint8_t *dest) {
    const int lane = threadIdx.x & 31, l = k.size;
    for (int e = lane; e < l * stride; e += 32) {
        const int c = e / stride, i = e - c * stride;
        uint8_t v = 0;
        for (int r = 0; r < l; ++r) v ^= k.mat[r * l + c] & src[r * stride + i];
        dest[e] = v;
    }
    __syncwarp();
}

// min-sum Viterbi of one kernel phase for stride element i (TrellisKernelProcessor.cpp:260-293), one warp,
// lanes across the next states.  sec[j] = offset of section j's predecessor table | (state bits after j) << 24.
// A predecessor entry holds two (state | branch bit << 15) halves; a missing branch points to the DUMMY state
// (index 2^max_ab) whose metric is +inf, which keeps the loop branch-free: x + 0.0f and min(x, inf) are exact.
// m0/m1: 2^max_ab + 1 floats each of warp-private shared memory.
__device__ __forceinline__ float viterbi(const uint32_t *__restrict__ pred, const uint32_t *__restrict__ sec, int l, int stride,
                                         int dummy, const float *chan, const uint8_t *offs, float *m0, float *m1) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) { m0[0] = 0.0f; m0[dummy] = HUGE_VALF; m1[dummy] = HUGE_VALF; }
    __syncwarp();
    for (int j = 0; j < l; ++j) {
        float y = chan[j * stride];
        if (offs[j * stride]) y = -y;
        const uint32_t hd = y < 0.0f;
        const float ay = fabsf(y);
        const uint32_t sj = sec[j];
        const uint32_t *tab = pred + (sj & 0xFFFFFFu);
        const int ns = 1 << (sj >> 24);
        for (int s1 = lane; s1 < ns; s1 += 32) {
            const uint32_t e = tab[s1];
            const float va = m0[e & 0x7FFFu] + ((((e >> 15) & 1u) ^ hd) ? ay : 0.0f);
            const float vb = m0[(e >> 16) & 0x7FFFu] + (((e >> 31) ^ hd) ? ay : 0.0f);
            m1[s1] = vb < va ? vb : va;
        }
        __syncwarp();
        float *t = m0; m0 = m1; m1 = t;
    }
    const float r = m0[1] - m0[0];   // :292
    __syncwarp();
    return r;
}

// Same recursion with ONE LANE PER STRIDE ELEMENT (lanes 0..stride-1 each walk their own trellis sequentially):
// used at the outer layers when the phase's trellis is small (<= 16 states), where lanes-across-states would
// leave the warp idle.  met: lane-private slice of 2 * 17 floats (index 16 = DUMMY).  Same float operations.
__device__ __forceinline__ float viterbi_lane(const uint32_t *__restrict__ pred, const uint32_t *__restrict__ sec, int l, int stride,
                                              const float *chan, const uint8_t *offs, float *met) {
    float *m0 = met, *m1 = met + 17;
    m0[0] = 0.0f;
    m0[16] = HUGE_VALF;
    m1[16] = HUGE_VALF;
    for (int j = 0; j < l; ++j) {
        float y = chan[j * stride];
        if (offs[j * stride]) y = -y;
        const uint32_t hd = y < 0.0f;
        const float ay = fabsf(y);
        const uint32_t sj = sec[j];
        const uint32_t *tab = pred + (sj & 0xFFFFFFu);
        const int ns = 1 << (sj >> 24);
        for (int s1 = 0; s1 < ns; ++s1) {
            const uint32_t e = tab[s1];
            const uint32_t ia = min(e & 0x7FFFu, 16u), ib = min((e >> 16) & 0x7FFFu, 16u);
            const float va = m0[ia] + ((((e >> 15) & 1u) ^ hd) ? ay : 0.0f);
            const float vb = m0[ib] + (((e >> 31) ^ hd) ? ay : 0.0f);
            m1[s1] = vb < va ? vb : va;
        }
        float *t = m0; m0 = m1; m1 = t;
    }
    return m0[1] - m0[0];
}

// GetLLRs (TrellisKernelProcessor.cpp:234-295): offset update for the newly known input `phase-1`, then one
// Viterbi per stride element.  known: [l][stride] decided kernel inputs; offs: [l][stride] state.
__device__ __forceinline__ void get_llrs(const PkPolarKernelDev &k, int stride, int phase, const uint8_t *known,
                                         const float *chan, float *out, uint8_t *offs, float *met) {
    const int lane = threadIdx.x & 31, l = k.size;
    if (phase == 0) {
        for (int e = lane; e < l * stride; e += 32) offs[e] = 0;
    } else {
        const uint8_t *row = k.mat + (phase - 1) * l;
        for (int e = lane; e < l * stride; e += 32) {
            const int c = e / stride, i = e - c * stride;
            if (row[c]) offs[e] ^= known[(phase - 1) * stride + i];
        }
    }
    __syncwarp();
    const uint32_t *pred = k.pred, *sec = k.off + phase * l;   // off[] holds the packed section words
    const int mab = k.ab[phase * (l + 1)];                     // ab[p][0] carries the phase's state complexity
    if (stride > 1 && mab <= 4) {
        for (int i0 = 0; i0 < stride; i0 += 32) {
            const int i = i0 + lane;
            if (i < stride) out[i] = viterbi_lane(pred, sec, l, stride, chan + i, offs + i, met + lane * 35);   // 35: conflict-free slices
        }
    } else {
        const int dummy = 1 << k.max_ab;
        float *m0 = met, *m1 = met + dummy + 1;
        for (int i = 0; i < stride; ++i) {
            const float v = viterbi(pred, sec, l, stride, dummy, chan + i, offs + i, m0, m1);
            if (lane == 0) out[i] = v;
        }
    }
    __syncwarp();
}

// per-path shared-memory layout
struct PathLayout {
    int s_off[PK_POLAR_MAX_LAYERS + 1];   // S arrays of layers 1..m (floats, offset in floats)
    int c_off[PK_POLAR_MAX_LAYERS + 1];   // C arrays of layers 0..m (bytes)
    int o_off[PK_POLAR_MAX_LAYERS];       // kernel-processor offsets of layers 0..m-1 (bytes)
    int u_off;                            // decided symbols, N0 bits as uint32 (bytes offset, 4-aligned)
    int floats, bytes;                    // totals per path
};
__host__ __device__ inline PathLayout path_layout(const PkPolarDev &d) {
    PathLayout p;
    int f = 0, b = 0;
    p.s_off[0] = 0;
    for (int j = 1; j <= d.layers; ++j) { p.s_off[j] = f; f += d.outer[j]; }
    for (int j = 0; j <= d.layers; ++j) { p.c_off[j] = b; b += d.outer[j] * (j > 0 ? d.ksize[j - 1] : 1); }
    for (int j = 0; j < d.layers; ++j) { p.o_off[j] = b; b += d.outer[j]; }
    b = (b + 3) & ~3;
    p.u_off = b;
    b += d.nw * 4;
    p.floats = f;
    p.bytes = (b + 15) & ~15;
    return p;
}

// ------------------------------------------------------------------ encoder (a15)
// One warp: a[] holds the K information symbols at their positions (zero elsewhere) on entry and the unshortened
// codeword on exit (the returned pointer is a or b).  Frozen symbols are evaluated in order
// (MixedKernelEncoder.cpp:148-159), then the layers m-1 .. 0 multiply every block by its kernel (:161-173).
__device__ __forceinline__ uint8_t *polar_encode_warp(const PkPolarDev &d, uint8_t *a, uint8_t *b) {
    const int lane = threadIdx.x & 31;
    if (!d.all_static && lane == 0) {
        for (int i = 0; i < d.N0; ++i) {
            if (!d.frozen[i]) continue;
            uint8_t v = 0;
            for (int w = 0; w < d.nw; ++w) {
                uint32_t m = d.cmask[i * d.nw + w];
                while (m) { const int t = __ffs(m) - 1; m &= m - 1; v ^= a[32 * w + t]; }
            }
            a[i] = v;
        }
    }
    __syncwarp();
    int stride = 1;
    for (int L = d.layers - 1; L >= 0; --L) {
        const int l = d.ksize[L], bs = l * stride;
        for (int e = lane; e < d.N0; e += 32) {
            const int blk = e / bs, r0 = e - blk * bs, c = r0 / stride, i = r0 - c * stride;
            uint8_t v = 0;
            for (int r = 0; r < l; ++r) v ^= d.kern[L].mat[r * l + c] & a[blk * bs + r * stride + i];
            b[e] = v;
        }
        __syncwarp();
        uint8_t *t = a; a = b; b = t;
        stride = bs;
    }
    return a;
}

__global__ void __launch_bounds__(128)
k_polar_encode(PkPolarDev d, const uint8_t *__restrict__ info, long B, uint8_t *__restrict__ cw) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint8_t *a0 = smem + (size_t)warp * 2 * d.N0, *b0 = a0 + d.N0;
    for (long f = (long)blockIdx.x * nwarps + warp; f < B; f += (long)gridDim.x * nwarps) {
        for (int i = lane; i < d.N0; i += 32) a0[i] = 0;
        __syncwarp();
        for (int q = lane; q < d.K; q += 32) a0[d.info_pos[q]] = info[f * d.K + q] ? 1 : 0;
        __syncwarp();
        const uint8_t *a = polar_encode_warp(d, a0, b0);
        // Shorten (:115-139): drop shortened / punctured symbols
        if (!d.symtype) {
            for (int i = lane; i < d.N0; i += 32) cw[f * d.N + i] = a[i];
        } else if (lane == 0) {
            int o = 0;
            for (int i = 0; i < d.N0; ++i)
                if (d.symtype[i] == 0) cw[f * d.N + o++] = a[i];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ generation mode (the simulator loop on the device)
// CSimulator::Iterate (out/external/Simulator.cpp:139-335, not buildable: GSL / Windows) restated for the device:
// random information symbols -> Encode -> BPSK (bit 0 -> +1, Modem.h:64) + AWGN -> LLR = 2 y / sigma^2 (> 0 <=> bit 0,
// :78) as fp32.  Frame f of SNR point s draws from Philox4x32-10 with key = seed, counter (f_lo, f_hi, block, s): blocks
// 0.. hold the information bits (128 per block), blocks 0x1000+q the Box-Muller pair of transmitted symbols 2q, 2q+1.
__global__ void __launch_bounds__(128)
k_polar_generate(PkPolarDev d, double sigma, uint64_t seed, uint64_t first_frame, uint32_t snr_index, long B,
                 uint8_t *__restrict__ info_out, uint8_t *__restrict__ cw_out, float *__restrict__ llr_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint8_t *a0 = smem + (size_t)warp * 3 * d.N0, *b0 = a0 + d.N0, *tx = b0 + d.N0;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (long f = (long)blockIdx.x * nwarps + warp; f < B; f += (long)gridDim.x * nwarps) {
        const unsigned long long gf = first_frame + (unsigned long long)f;
        const uint32_t c0 = (uint32_t)gf, c1 = (uint32_t)(gf >> 32);
        for (int i = lane; i < d.N0; i += 32) a0[i] = 0;
        __syncwarp();
        for (int q0 = 0; q0 < d.K; q0 += 32) {
            const int q = q0 + lane;
            const PkPhilox r = pk_philox(c0, c1, (uint32_t)(q0 >> 7), snr_index, k0, k1);   // same word in all lanes of a group of 32
            const uint32_t word = r.c[(q0 >> 5) & 3];
            if (q < d.K) {
                const uint8_t bit = (uint8_t)((word >> lane) & 1u);
                a0[d.info_pos[q]] = bit;
                if (info_out) info_out[f * d.K + q] = bit;
            }
        }
        __syncwarp();
        const uint8_t *a = polar_encode_warp(d, a0, b0);
        // Shorten: the transmitted symbols
        if (!d.symtype) {
            for (int i = lane; i < d.N0; i += 32) tx[i] = a[i];
        } else if (lane == 0) {
            int o = 0;
            for (int i = 0; i < d.N0; ++i)
                if (d.symtype[i] == 0) tx[o++] = a[i];
        }
        __syncwarp();
        for (int q = lane; 2 * q < d.N; q += 32) {
            const PkPhilox r = pk_philox(c0, c1, 0x1000u + (uint32_t)q, snr_index, k0, k1);
            const unsigned long long ra = ((unsigned long long)r.c[0] << 32) | r.c[1];
            const unsigned long long rb = ((unsigned long long)r.c[2] << 32) | r.c[3];
            const double u1 = ((double)(ra >> 11) + 1.0) * (1.0 / 9007199254740992.0);   // (0,1]
            const double u2 = (double)(rb >> 11) * (1.0 / 9007199254740992.0);           // [0,1)
            const double rad = sqrt(-2.0 * log(u1));
            double sn, cs;
            sincospi(2.0 * u2, &sn, &cs);
            const int p0 = 2 * q, p1 = 2 * q + 1;
            const double y0 = (tx[p0] ? -1.0 : 1.0) + sigma * (rad * cs);
            llr_out[f * d.N + p0] = (float)(2.0 * y0 / (sigma * sigma));
            if (cw_out) cw_out[f * d.N + p0] = tx[p0];
            if (p1 < d.N) {
                const double y1 = (tx[p1] ? -1.0 : 1.0) + sigma * (rad * sn);
                llr_out[f * d.N + p1] = (float)(2.0 * y1 / (sigma * sigma));
                if (cw_out) cw_out[f * d.N + p1] = tx[p1];
            }
        }
        __syncwarp();
    }
}

// decided information vector of the best path against the transmitted one: frame / bit error counters
// (pk_point_result layout: [0] frames, [1] frame errors, [2] information-bit errors, [7] flags)
__global__ void __launch_bounds__(256)
k_polar_compare(int K, int L, const uint8_t *__restrict__ info, const uint8_t *__restrict__ inf_out, const int *__restrict__ count,
                long B, unsigned long long *__restrict__ totals) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    unsigned long long fr = 0, fe = 0, be = 0;
    for (long f = warp; f < B; f += nwarps) {
        int e = 0;
        for (int q = lane; q < K; q += 32) e += (info[f * K + q] != inf_out[(f * L) * K + q]) ? 1 : 0;
        e = __reduce_add_sync(PKP_FULL, e);
        if (count[f] < 1) e = K;   // no path survived (cannot happen with list size >= 1; counted as a total loss)
        fr += 1;
        fe += e ? 1 : 0;
        be += (unsigned long long)e;
    }
    if (lane == 0 && fr) {
        atomicAdd(totals + 0, fr);
        if (fe) atomicAdd(totals + 1, fe);
        if (be) atomicAdd(totals + 2, be);
    }
}

// ------------------------------------------------------------------ kernel LLRs of independent kernel blocks (a16/a17)
// One warp per block: phases 0..l-1 in order with the given (genie) kernel inputs, stride 1.
__global__ void __launch_bounds__(128)
k_polar_kernel_llr(PkPolarKernelDev k, const float *__restrict__ chan, const uint8_t *__restrict__ u, long B,
                   float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, l = k.size;
    const size_t per = (size_t)polar_met_floats(k.max_ab) * 4 + 4 * l + 2 * l + 16;
    unsigned char *wb = smem + (size_t)warp * ((per + 15) & ~(size_t)15);
    float *met = reinterpret_cast<float *>(wb);
    float *ch = met + polar_met_floats(k.max_ab);
    uint8_t *known = reinterpret_cast<uint8_t *>(ch + l), *offs = known + l;
    for (long f = (long)blockIdx.x * nwarps + warp; f < B; f += (long)gridDim.x * nwarps) {
        for (int i = lane; i < l; i += 32) { ch[i] = chan[f * l + i]; known[i] = u[f * l + i] ? 1 : 0; }
        __syncwarp();
        for (int ph = 0; ph < l; ++ph) {
            float v;
            get_llrs(k, 1, ph, known, ch, &v, offs, met);   // out[0] written by lane 0 into v (register of lane 0)
            v = __shfl_sync(PKP_FULL, v, 0);
            if (lane == 0) out[f * l + ph] = v;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ SC / SC-list decoder (a18, a19)
struct ListCtl {            // CTA-shared list state
    float R[32];            // path metrics m_pR
    float llr[32];          // LLR of the current phase per path
    uint8_t active[32];
    uint8_t cont[32];       // continuation mask: 1 = extend with 0, 2 = extend with 1, 3 = both (clone)
    uint8_t bit[32];        // decided symbol of this phase per path
    int8_t clone_src[32];   // for a path created in this phase: the path it was cloned from, else -1
    uint32_t stack[34];     // inactive path indices (lazily initialised stack, misc.h:212-226)
    float cand_m[64];
    uint32_t cand_s[64];
    int n_cand;
};
__device__ __forceinline__ uint32_t stack_pop(uint32_t *st) {
    if (st[st[0]] == 0xFFFFFFFFu) {
        const uint32_t s = --st[0];
        if (s > 0) st[st[0]] = 0xFFFFFFFFu;
        return s;
    }
    return st[st[0]--];
}
__device__ __forceinline__ void stack_push(uint32_t x, uint32_t *st) { st[++st[0]] = x; }


// copies the tables of every distinct kernel into shared memory and returns descriptors pointing there
__device__ __forceinline__ void stage_polar_tables(const PkPolarDev &d, PkPolarKernelDev *sk, unsigned char *&sm) {
    for (int j = 0; j < d.layers; ++j) {
        int same = -1;
        for (int i = 0; i < j; ++i)
            if (d.kern[i].pred == d.kern[j].pred) same = i;
        if (same >= 0) { sk[j] = sk[same]; continue; }
        const PkPolarKernelDev &g = d.kern[j];
        const int l = g.size;
        uint32_t *pred = reinterpret_cast<uint32_t *>(sm);
        uint32_t *off = pred + g.npred;
        uint8_t *ab = reinterpret_cast<uint8_t *>(off + l * l);
        uint8_t *mat = ab + l * (l + 1);
        for (int i = threadIdx.x; i < g.npred; i += blockDim.x) pred[i] = g.pred[i];
        for (int i = threadIdx.x; i < l * l; i += blockDim.x) { off[i] = g.off[i]; mat[i] = g.mat[i]; }
        for (int i = threadIdx.x; i < l * (l + 1); i += blockDim.x) ab[i] = g.ab[i];
        sk[j] = g;
        sk[j].pred = pred; sk[j].off = off; sk[j].ab = ab; sk[j].mat = mat;
        sm += ((size_t)g.npred * 4 + (size_t)l * l * 4 + (size_t)l * (l + 1) + (size_t)l * l + 15) & ~(size_t)15;
    }
    __syncthreads();
}
__host__ inline size_t polar_table_bytes(const PkPolarDev &d) {
    size_t t = 0;
    for (int j = 0; j < d.layers; ++j) {
        bool dup = false;
        for (int i = 0; i < j; ++i) dup = dup || d.kern[i].pred == d.kern[j].pred;
        if (dup) continue;
        const int l = d.kern[j].size;
        t += ((size_t)d.kern[j].npred * 4 + (size_t)l * l * 4 + (size_t)l * (l + 1) + (size_t)l * l + 15) & ~(size_t)15;
    }
    return t;
}

// CTA = FPC frames x L list paths, one warp per (frame, path).  All frames of a CTA walk the N0 phases in lock
// step (the frozen pattern is a property of the code), so CTA-wide barriers are uniform.
__global__ void __launch_bounds__(1024)
k_polar_decode(PkPolarDev d, int L, int FPC, const float *__restrict__ llr_in, long B, int *__restrict__ count,
               uint8_t *__restrict__ inf_out, uint8_t *__restrict__ cw_out, float *__restrict__ metric_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ PkPolarKernelDev sk[PK_POLAR_MAX_LAYERS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fi = warp / L, path = warp - fi * L;           // frame slot in the CTA, list path
    const PathLayout pl = path_layout(d);
    unsigned char *sm = smem;
    stage_polar_tables(d, sk, sm);
    // shared layout after the tables: per frame [ListCtl][chan N0 floats][L paths], then per warp Viterbi scratch
    const size_t path_sz = (((size_t)pl.floats * 4 + 15) & ~(size_t)15) + pl.bytes;
    const size_t frame_sz = ((sizeof(ListCtl) + 15) & ~(size_t)15) + (size_t)d.N0 * 4 + (size_t)L * path_sz;
    unsigned char *fb = sm + (size_t)fi * frame_sz;
    ListCtl *lc = reinterpret_cast<ListCtl *>(fb);
    float *chan = reinterpret_cast<float *>(fb + ((sizeof(ListCtl) + 15) & ~(size_t)15));
    unsigned char *paths = reinterpret_cast<unsigned char *>(chan + d.N0);
    float *met = reinterpret_cast<float *>(sm + (size_t)FPC * frame_sz) + (size_t)warp * polar_met_floats(d.max_ab);
    auto pS = [&](int p) { return reinterpret_cast<float *>(paths + (size_t)p * path_sz); };
    auto pB = [&](int p) { return paths + (size_t)p * path_sz + (((size_t)pl.floats * 4 + 15) & ~(size_t)15); };
    const int last = d.layers - 1, lsz = d.ksize[last];

    for (long f0 = (long)blockIdx.x * FPC; f0 < B; f0 += (long)gridDim.x * FPC) {
        const long f = f0 + fi;
        const bool live = f < B;       // a frame slot past the end idles but keeps the barriers
        __syncthreads();
        if (live) {
            // LoadLLRs (MixedKernelEncoder.cpp:179-203)
            if (!d.symtype) {
                for (int i = path * 32 + lane; i < d.N0; i += 32 * L) chan[i] = llr_in[f * d.N + i];
            } else if (path == 0 && lane == 0) {
                int o = 0;
                for (int i = 0; i < d.N0; ++i)
                    chan[i] = d.symtype[i] == 0 ? llr_in[f * d.N + o++] : (d.symtype[i] == 1 ? PKP_UPPER : 0.0f);
            }
            if (path == 0 && lane == 0) {
                // Cleanup + AssignInitialPath (TVMemoryEngine.cpp:58-94)
                lc->stack[0] = (uint32_t)L;
                lc->stack[L] = 0xFFFFFFFFu;
                for (int p = 0; p < 32; ++p) { lc->active[p] = 0; lc->clone_src[p] = -1; }
                const uint32_t pid = stack_pop(lc->stack);
                lc->active[pid] = 1;
                lc->R[pid] = 0.0f;
            }
        }
        __syncthreads();
        if (live && lc->active[path]) {
            uint32_t *uh = reinterpret_cast<uint32_t *>(pB(path) + pl.u_off);
            for (int w = lane; w < d.nw; w += 32) uh[w] = 0;
        }
        __syncthreads();

        for (int phi = 0; phi < d.N0; ++phi) {
            // ---- every active path: LLR of symbol phi (IterativelyCalcS, KernelListEngine.cpp:370-447)
            if (live && lc->active[path]) {
                float *S = pS(path);
                uint8_t *Bp = pB(path);
                int pv = phi, mm = last;
                while (mm > 0 && (pv % d.ksize[mm]) == 0) { pv /= d.ksize[mm]; --mm; }
                const float *src = (mm == 0) ? chan : S + pl.s_off[mm];
                for (int j = mm; j <= last; ++j) {
                    float *dest = S + pl.s_off[j + 1];
                    get_llrs(sk[j], d.outer[j + 1], pv % d.ksize[j], Bp + pl.c_off[j + 1], src, dest, Bp + pl.o_off[j], met);
                    pv = 0;
                    src = dest;
                }
                if (lane == 0) lc->llr[path] = S[pl.s_off[d.layers]];
            }
            __syncthreads();
            const bool frozen = d.frozen[phi] != 0;
            if (frozen) {
                // ---- ContinuePathsFrozen (MixedKernelListDecoder.cpp:61-98)
                if (live && lc->active[path]) {
                    const uint32_t *uh = reinterpret_cast<const uint32_t *>(pB(path) + pl.u_off);
                    uint32_t par = 0;
                    if (!d.all_static)
                        for (int w = lane; w < d.nw; w += 32) par ^= __popc(uh[w] & d.cmask[phi * d.nw + w]) & 1u;
                    par = __reduce_xor_sync(PKP_FULL, par);
                    if (lane == 0) {
                        const float v = lc->llr[path];
                        if ((par != 0) ^ (v < 0.0f)) lc->R[path] -= fabsf(v);
                        lc->bit[path] = (uint8_t)par;
                    }
                }
            } else if (live && path == 0) {
                // ---- ContinuePathsUnfrozen (:100-185), list bookkeeping by the first warp of the frame
                int J = 0;
                if (lane == 0) {
                    for (int p = 0; p < L; ++p) {
                        lc->cont[p] = 0;
                        lc->clone_src[p] = -1;
                        if (!lc->active[p]) continue;
                        const float v = lc->llr[p];
                        const uint32_t D = v < 0.0f;
                        lc->cand_m[J] = lc->R[p];
                        lc->cand_s[J] = 2u * p + D;
                        lc->cand_m[J + 1] = lc->R[p] - fabsf(v);
                        lc->cand_s[J + 1] = 2u * p + (D ^ 1u);
                        J += 2;
                    }
                    lc->n_cand = J;
                }
                __syncwarp();
                J = lc->n_cand;
                // descending order of (metric, second) pairs (CPairComparator = std::greater<pair>): rank by counting
                const int keep = J < L ? J : L;
                for (int c = lane; c < J; c += 32) {
                    const float mc = lc->cand_m[c];
                    const uint32_t sc = lc->cand_s[c];
                    int rank = 0;
                    for (int k = 0; k < J; ++k) {
                        const float mk = lc->cand_m[k];
                        rank += (mk > mc || (mk == mc && lc->cand_s[k] > sc)) ? 1 : 0;
                    }
                    if (rank < keep) atomicOr(reinterpret_cast<unsigned int *>(lc->cont) + (sc >> 3), (1u << (sc & 1u)) << (8 * ((sc >> 1) & 3)));
                }
                __syncwarp();
                if (lane == 0) {
                    for (int p = 0; p < L; ++p)
                        if (lc->active[p] && !lc->cont[p]) { stack_push((uint32_t)p, lc->stack); lc->active[p] = 0; }   // KillPath
                    for (int p = 0; p < L; ++p) {
                        switch (lc->cont[p]) {
                        case 1: lc->bit[p] = 0; break;
                        case 2: lc->bit[p] = 1; break;
                        case 3: {
                            const float v = lc->llr[p];
                            const uint8_t C = v < 0.0f ? 1 : 0;
                            lc->bit[p] = C;
                            const uint32_t p1 = stack_pop(lc->stack);   // ClonePath
                            lc->bit[p1] = C ^ 1;
                            lc->active[p1] = 1;
                            lc->R[p1] = lc->R[p] - fabsf(v);
                            lc->clone_src[p1] = (int8_t)p;
                            break;
                        }
                        default: break;
                        }
                    }
                }
            }
            __syncthreads();
            // ---- clones copy their parent's arrays (eager version of the reference's copy-on-write)
            if (live && !frozen && lc->active[path] && lc->clone_src[path] >= 0) {
                const uint32_t *s = reinterpret_cast<const uint32_t *>(paths + (size_t)lc->clone_src[path] * path_sz);
                uint32_t *t = reinterpret_cast<uint32_t *>(paths + (size_t)path * path_sz);
                for (int i = lane; i < (int)(path_sz / 4); i += 32) t[i] = s[i];
            }
            __syncthreads();
            // ---- write the decided symbol and propagate completed kernel blocks (IterativelyUpdateC, :266-315)
            if (live && lc->active[path]) {
                uint8_t *Bp = pB(path);
                const uint8_t C = lc->bit[path];
                if (lane == 0) {
                    Bp[pl.c_off[d.layers] + (phi % lsz)] = C;
                    uint32_t *uh = reinterpret_cast<uint32_t *>(Bp + pl.u_off);
                    if (C) uh[phi >> 5] |= 1u << (phi & 31);
                }
                __syncwarp();
                int lambda = d.layers, stride = 1, pv = phi;
                while (lambda > 0 && ((pv + 1) % d.ksize[lambda - 1]) == 0) {
                    const int psi = pv / d.ksize[lambda - 1];
                    const int next = stride * d.ksize[lambda - 1];
                    const int phi0 = (lambda > 1) ? (psi % d.ksize[lambda - 2]) * next : 0;
                    kernel_multiply(sk[lambda - 1], stride, Bp + pl.c_off[lambda], Bp + pl.c_off[lambda - 1] + phi0);
                    stride = next;
                    pv = psi;
                    --lambda;
                }
            }
            __syncthreads();
        }

        // ---- final ordering (MixedKernelListDecoder.cpp:253-266): active paths by (R, index) descending
        if (live && lc->active[path]) {
            const float r = lc->R[path];
            int rank = 0;
            for (int p = 0; p < L; ++p)
                if (lc->active[p] && (lc->R[p] > r || (lc->R[p] == r && p > path))) ++rank;
            const uint8_t *Bp = pB(path);
            const uint32_t *uh = reinterpret_cast<const uint32_t *>(Bp + pl.u_off);
            if (inf_out)
                for (int q = lane; q < d.K; q += 32) {
                    const int pos = d.info_pos[q];
                    inf_out[(f * L + rank) * d.K + q] = (uint8_t)((uh[pos >> 5] >> (pos & 31)) & 1u);
                }
            if (cw_out) {
                const uint8_t *c0 = Bp + pl.c_off[0];
                if (!d.symtype) {
                    for (int i = lane; i < d.N0; i += 32) cw_out[(f * L + rank) * d.N + i] = c0[i];
                } else if (lane == 0) {
                    int o = 0;
                    for (int i = 0; i < d.N0; ++i)
                        if (d.symtype[i] == 0) cw_out[(f * L + rank) * d.N + o++] = c0[i];
                }
            }
            if (metric_out && lane == 0) metric_out[f * L + rank] = r;
        }
        if (live && path == 0 && lane == 0 && count) {
            int J = 0;
            for (int p = 0; p < L; ++p) J += lc->active[p] ? 1 : 0;
            count[f] = J;
        }
    }
}


#include "pk_polar_lanes.cuh"

// ------------------------------------------------------------------ host handle + C ABI
struct pk_polar {
    pk_polar_code code;
    PkPolarDev dev{};
    int L = 1, device = 0;
    std::vector<void *> allocs;
    cudaStream_t stream = nullptr;
    size_t smem_decode = 0;
    int fpc = 1;   // frames per CTA
    // decoder with paths across lanes (pk_polar_lanes.cuh); ln_g = 0: not applicable, k_polar_decode runs
    PkLanesDev lanes{};
    int ln_g = 0, ln_ls = 0, ln_warps = 0, ln_grid = 0;   // lanes per slot, slots per frame (list size rounded up to a power of two), warps per CTA, CTAs
    size_t ln_smem = 0;
    float *ln_chan = nullptr;                    // transposed channel LLRs, one block per warp of the grid
    // generation-mode workspaces (one chunk of frames)
    long gen_cap = 0;
    uint8_t *g_info = nullptr, *g_inf = nullptr;
    float *g_llr = nullptr;
    int *g_cnt = nullptr;
    unsigned long long *g_tot = nullptr, *h_tot = nullptr;
};

namespace {
template <class Tp>
cudaError_t up(pk_polar *h, const std::vector<Tp> &v, const Tp **d) {
    *d = nullptr;
    if (v.empty()) return cudaSuccess;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, v.size() * sizeof(Tp));
    if (e != cudaSuccess) return e;
    h->allocs.push_back(p);
    *d = static_cast<const Tp *>(p);
    return cudaMemcpy(p, v.data(), v.size() * sizeof(Tp), cudaMemcpyHostToDevice);
}

// ---- lanes decoder set-up: tables in the kernel's format, shared-memory budget, scratch
#define PK_LANES_CASES(X) X(1, 1) X(1, 2) X(1, 4) X(2, 1) X(2, 2) X(2, 4) X(4, 1) X(4, 2) X(4, 4) X(8, 1) X(8, 2) X(8, 4) X(16, 1) X(16, 2) X(32, 1)

cudaError_t lanes_launch(pk_polar *h, const float *d_llr, long B, int *d_count, uint8_t *d_inf, uint8_t *d_cw, float *d_metric, cudaStream_t st) {
#define X(LL, GG)                                                                                                                       \
    if (h->ln_ls == LL && h->ln_g == GG) {                                                                                                \
        k_polar_lanes<LL, GG><<<h->ln_grid, 32 * h->ln_warps, h->ln_smem, st>>>(h->dev, h->lanes, h->L, d_llr, B, d_count, d_inf, d_cw, d_metric, h->ln_chan); \
        return cudaGetLastError();                                                                                                        \
    }
    PK_LANES_CASES(X)
#undef X
    return cudaErrorInvalidConfiguration;
}
cudaError_t lanes_attr(pk_polar *h) {
#define X(LL, GG) \
    if (h->ln_ls == LL && h->ln_g == GG) return cudaFuncSetAttribute(k_polar_lanes<LL, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ln_smem);
    PK_LANES_CASES(X)
#undef X
    return cudaErrorInvalidConfiguration;
}

// Leaves h->ln_g = 0 (k_polar_decode runs instead) when the code does not fit the scheme: a kernel trellis too wide for
// the 16-bit row offsets / the shared memory of an SM.
cudaError_t lanes_setup(pk_polar *h) {
    const pk_polar_code &c = h->code;
    const PkPolarDev &d = h->dev;
    int L = 1;                      // slots per frame: the list size rounded up to a power of two (the spare slots stay empty)
    while (L < h->L) L <<= 1;
    const char *off = getenv("PK_POLAR_LANES");
    if (off && atoi(off) == 0) return cudaSuccess;
    const char *env = getenv("PK_POLAR_LANES_G");
    int G = env ? atoi(env) : 2;
    if (G != 1 && G != 2 && G != 4) G = 2;
    while (L * G > 32) G >>= 1;
    PkLanesDev &ld = h->lanes;
    ld.nk = (int)c.kernels.size();
    ld.ns_rows = 1;
    for (const PkKernelTrellis &k : c.kernels) {
        if (!k.ip_ok || k.ip_bits > 14) return cudaSuccess;
        ld.ns_rows = std::max(ld.ns_rows, 1 << k.ip_bits);
    }
    // wide trellises: more lanes per slot make the metric rows shorter (16-bit row offsets, shared memory per warp)
    while ((size_t)ld.ns_rows * (32 / G) * 4 > 65534 && G < 4 && L * G * 2 <= 32) G *= 2;
    const int nslot = 32 / G;
    if ((size_t)ld.ns_rows * nslot * 4 > 65534) return cudaSuccess;
    for (int j = 0; j < c.layers; ++j) ld.kidx[j] = c.kid[j];
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < ld.nk && e == cudaSuccess; ++i) {
        const PkKernelTrellis &k = c.kernels[i];
        const int l = k.size;
        // entries: state index -> byte offset of its metric row | PAD << 1 | label bit; 16 bits each.  The G lanes of a slot
        // take the entries of a group of four interleaved (lane g: entries g, g + G, ..), so that the lanes of a slot read
        // neighbouring rows, which lie in different banks: the table stores every group in lane order.
        std::vector<uint16_t> ent;
        std::vector<uint32_t> sec(k.ip_sec);
        for (int p = 0; p < l; ++p)
            for (int j = 0; j <= l; ++j) {
                uint32_t *w = &sec[((size_t)p * (l + 1) + j) * 2];
                const uint32_t first = w[0] & 0xFFFFFFu, ngroups = w[0] >> 24, q = w[1] & 0xFFu, type = w[1] >> 8;
                const uint32_t qoff = (((1u << q) * (uint32_t)nslot * 4u) & 0xFFFFu);
                if (j == l || ngroups == 0) { w[0] = 0; w[1] = qoff; continue; }
                uint32_t n = 0;
                for (uint32_t x = 0; x < 4 * ngroups; ++x) n += k.ip_x[first + x] != 0xFFFFu;
                const bool tiny = n < 8;   // n is a power of two
                w[0] = (uint32_t)(ent.size() * 2) | ((tiny ? 0u : n / 8) << 24);
                w[1] = qoff | ((type | (tiny ? 4u : 0u) | 8u) << 16);
                for (uint32_t gr = 0; gr < ngroups; ++gr)
                    for (int gg = 0; gg < G; ++gg)
                        for (int u = 0; u < 4 / G; ++u) {
                            const uint32_t v = k.ip_x[first + 4 * gr + gg + G * u];
                            ent.push_back(v == 0xFFFFu ? (uint16_t)2u : (uint16_t)(((v & 0x7FFFu) * (uint32_t)nslot * 4u) | (v >> 15)));
                        }
            }
        ent.resize((ent.size() + 1) & ~(size_t)1, 2u);
        ent.resize(ent.size() + 16, 2u);
        std::vector<uint32_t> tab(ent.size() / 2);
        for (size_t x = 0; x < tab.size(); ++x) tab[x] = (uint32_t)ent[2 * x] | ((uint32_t)ent[2 * x + 1] << 16);
        // bounds check of everything the kernel will dereference from these tables (it has no checks of its own): every
        // batch lies inside the table, every row offset and partner offset inside the metric buffer
        {
            const uint32_t met_bytes = (uint32_t)ld.ns_rows * (uint32_t)nslot * 4u;
            for (int p = 0; p < l; ++p)
                for (int j = 0; j <= l; ++j) {
                    const uint32_t *w = &sec[((size_t)p * (l + 1) + j) * 2];
                    const uint32_t off = w[0] & 0xFFFFFFu, nb = w[0] >> 24, qoff = w[1] & 0xFFFFu, ty = w[1] >> 16;
                    if (j == l) { if (qoff >= met_bytes) return cudaErrorInvalidValue; continue; }
                    if (!(ty & 8u)) continue;
                    const uint32_t nent = (ty & 4u) ? 4u : 8u * nb;
                    if ((off & 7u) || off / 2 + nent > ent.size() - 16) return cudaErrorInvalidValue;
                    for (uint32_t x = 0; x < nent; ++x) {
                        const uint32_t en = ent[off / 2 + x];
                        if (en & 2u) { if (!(ty & 4u)) return cudaErrorInvalidValue; continue; }   // padding only in one-group sections
                        const uint32_t ro = en & 0xFFE0u;
                        if ((ro % ((uint32_t)nslot * 4u)) || ro + ((ty & 3u) ? qoff : 0u) >= met_bytes) return cudaErrorInvalidValue;
                    }
                }
        }
        std::vector<unsigned long long> masks((size_t)2 * l, 0);
        for (int r = 0; r < l; ++r)
            for (int cc = 0; cc < l; ++cc)
                if (k.matrix[(size_t)r * l + cc]) { masks[cc] |= 1ull << r; masks[l + r] |= 1ull << cc; }
        ld.k[i].ntab = (int)tab.size();
        ld.k[i].size = l;
        e = up(h, tab, &ld.k[i].tab);
        if (e == cudaSuccess) e = up(h, sec, &ld.k[i].sec);
        if (e == cudaSuccess) e = up(h, masks, &ld.k[i].masks);
    }
    if (e != cudaSuccess) return e;
    const PathLayout pl = path_layout(d);
    const LanesLayout ly = lanes_layout(d, ld, pl, L, nslot);
    int smax = 0, sms = 0;
    cudaDeviceGetAttribute(&smax, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const long fit = ((long)smax - (long)ly.tables) / (long)ly.per_warp;
    const char *wenv = getenv("PK_POLAR_LANES_WARPS");
    const int wmax = G == 1 ? 8 : 16;   // __launch_bounds__ of k_polar_lanes
    const int warps = (int)std::min<long>(wenv ? std::min(wmax, atoi(wenv)) : wmax, fit);
    if (warps < 1) return cudaSuccess;
    h->ln_warps = warps;
    h->ln_grid = sms;
    h->ln_smem = (size_t)ly.tables + (size_t)warps * ly.per_warp;
    e = cudaMalloc(&h->ln_chan, (size_t)h->ln_grid * warps * d.N0 * (nslot / L) * sizeof(float));
    if (e != cudaSuccess) return e;
    h->ln_g = G;
    h->ln_ls = L;
    return lanes_attr(h);
}
}  // namespace

extern "C" {

// CMixedKernelListDecoder(std::istream& Spec, unsigned ListSize) (MixedKernelListDecoder.cpp:10): spec_text in the
// reference's specification format, L = list size (1 = plain successive cancellation), 1 <= L <= 32.
int pk_polar_create(const char *spec_text, int L, int device, pk_polar **out) {
    if (!out || !spec_text) return pk_set_error(PK_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (L < 1 || L > 32) return pk_set_error(PK_ERR_ARG, "list size must be in [1,32]");
    pk_polar *h = new (std::nothrow) pk_polar;
    if (!h) return pk_set_error(PK_ERR_ALLOC, "out of memory");
    std::string err = pk_polar_parse(h->code, spec_text);
    if (!err.empty()) { delete h; return pk_set_error(PK_ERR_ARG, err); }
    h->L = L;
    h->device = device;
    if (device < 0) { *out = h; return PK_OK; }   // host-only handle (introspection / CPU tests)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { delete h; return pk_set_error(PK_ERR_CUDA, "no CUDA device: libpkb200 has no CPU path"); }
    if (device >= ndev) { delete h; return pk_set_error(PK_ERR_ARG, "bad device ordinal"); }
    cudaError_t e = cudaSetDevice(device);
    const pk_polar_code &c = h->code;
    PkPolarDev &d = h->dev;
    d.N = c.N; d.K = c.K; d.N0 = c.N0; d.layers = c.layers; d.nw = (c.N0 + 31) / 32;
    d.max_ab = 0;
    std::vector<PkPolarKernelDev> kd(c.kernels.size());
    for (size_t i = 0; i < c.kernels.size() && e == cudaSuccess; ++i) {
        const PkKernelTrellis &k = c.kernels[i];
        kd[i].size = k.size; kd[i].max_ab = k.max_ab; kd[i].npred = (int)k.pred.size();
        // device copies: section word = table offset | (state bits after the section) << 24; the (always zero)
        // entry ab[p][0] carries the phase's maximal state complexity instead
        std::vector<uint32_t> sec(k.off);
        std::vector<uint8_t> abd(k.ab);
        for (int p = 0; p < k.size; ++p) {
            uint8_t mx = 0;
            for (int j = 0; j < k.size; ++j) {
                const uint8_t a = k.ab[(size_t)p * (k.size + 1) + j + 1];
                sec[(size_t)p * k.size + j] |= (uint32_t)a << 24;
                mx = std::max(mx, a);
            }
            abd[(size_t)p * (k.size + 1)] = mx;
        }
        std::vector<uint32_t> predd(k.pred);   // missing branch (0xFFFF) -> DUMMY state 2^max_ab, branch bit 0
        const uint32_t dummy = 1u << k.max_ab;
        for (auto &w : predd) {
            if ((w & 0xFFFFu) == 0xFFFFu) w = (w & 0xFFFF0000u) | dummy;
            if ((w >> 16) == 0xFFFFu) w = (w & 0x0000FFFFu) | (dummy << 16);
        }
        e = up(h, k.matrix, &kd[i].mat);
        if (e == cudaSuccess) e = up(h, abd, &kd[i].ab);
        if (e == cudaSuccess) e = up(h, predd, &kd[i].pred);
        if (e == cudaSuccess) e = up(h, sec, &kd[i].off);
    }
    for (int j = 0; j < c.layers; ++j) {
        d.ksize[j] = c.ksize[j];
        d.kern[j] = kd[c.kid[j]];
        d.max_ab = std::max(d.max_ab, c.kernels[c.kid[j]].max_ab);
    }
    for (int j = 0; j <= c.layers; ++j) d.outer[j] = c.outer[j];
    std::vector<uint8_t> fz(c.N0);
    bool all_static = true;
    for (int i = 0; i < c.N0; ++i) fz[i] = c.decision[i] >= 0;
    for (uint32_t m : c.cmask) all_static = all_static && (m == 0);
    d.all_static = all_static ? 1 : 0;
    std::vector<uint16_t> ip(c.info_pos.begin(), c.info_pos.end());
    if (e == cudaSuccess) e = up(h, fz, &d.frozen);
    if (e == cudaSuccess) e = up(h, c.cmask, &d.cmask);
    if (e == cudaSuccess) e = up(h, ip, &d.info_pos);
    if (e == cudaSuccess) e = up(h, c.symtype, &d.symtype);
    {
        std::vector<uint16_t> compact;
        if (!c.symtype.empty()) {
            compact.assign(c.N0, 0);
            int o = 0;
            for (int i = 0; i < c.N0; ++i) { compact[i] = (uint16_t)o; if (c.symtype[i] == 0) ++o; }
        }
        if (e == cudaSuccess) e = up(h, compact, &d.compact);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        const PathLayout pl = path_layout(d);
        const size_t path_sz = (((size_t)pl.floats * 4 + 15) & ~(size_t)15) + pl.bytes;
        const size_t frame_sz = ((sizeof(ListCtl) + 15) & ~(size_t)15) + (size_t)d.N0 * 4 + (size_t)L * path_sz;
        int smem_max = 0;
        e = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        h->fpc = std::max(1, 8 / L);   // at least 8 warps per CTA share the staged trellis tables ..
        auto need = [&](int fpc) { return polar_table_bytes(d) + (size_t)fpc * frame_sz + (size_t)fpc * L * polar_met_floats(d.max_ab) * 4; };
        while (h->fpc > 1 && e == cudaSuccess && need(h->fpc) > (size_t)smem_max) h->fpc /= 2;   // .. fewer where the trellises are wide
        h->smem_decode = need(h->fpc);
        if (e == cudaSuccess && h->smem_decode > (size_t)smem_max) {
            for (void *p : h->allocs) cudaFree(p);
            cudaStreamDestroy(h->stream);
            const std::string msg = "polar decoder needs " + std::to_string(h->smem_decode) + " bytes of shared memory per CTA (list size x code length x trellis states), the device offers " + std::to_string(smem_max);
            delete h;
            return pk_set_error(PK_ERR_UNSUPPORTED, msg);
        }
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_polar_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_decode);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_polar_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2 * d.N0);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_polar_generate, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 3 * d.N0);
        if (e == cudaSuccess) e = lanes_setup(h);
    }
    if (e != cudaSuccess) {
        std::string msg = std::string("pk_polar_create: ") + cudaGetErrorString(e);
        for (void *p : h->allocs) cudaFree(p);
        delete h;
        return pk_set_error(PK_ERR_CUDA, msg);
    }
    *out = h;
    return PK_OK;
}

void pk_polar_destroy(pk_polar *h) {
    if (!h) return;
    if (h->device >= 0) {
        cudaSetDevice(h->device);
        if (h->stream) cudaStreamDestroy(h->stream);
        for (void *p : h->allocs) cudaFree(p);
        cudaFree(h->g_info); cudaFree(h->g_inf); cudaFree(h->g_llr); cudaFree(h->g_cnt); cudaFree(h->g_tot);
        cudaFree(h->ln_chan);
        if (h->h_tot) cudaFreeHost(h->h_tot);
    }
    delete h;
}

int pk_polar_info(const pk_polar *h, int *N, int *K, int *N0, int *layers, int *L) {
    if (!h) return pk_set_error(PK_ERR_ARG, "NULL handle");
    if (N) *N = h->code.N;
    if (K) *K = h->code.K;
    if (N0) *N0 = h->code.N0;
    if (layers) *layers = h->code.layers;
    if (L) *L = h->L;
    return PK_OK;
}

// number of trellis state bits after every section, per phase, of the kernel of `layer`: out[l][l+1]
// (m_ppNumOfActiveBits, TrellisKernelProcessor.cpp:105,154)
int pk_polar_trellis_profile(const pk_polar *h, int layer, int *size, uint8_t *out) {
    if (!h || layer < 0 || layer >= h->code.layers) return pk_set_error(PK_ERR_ARG, "bad layer");
    const PkKernelTrellis &k = h->code.kernels[h->code.kid[layer]];
    if (size) *size = k.size;
    if (out) std::memcpy(out, k.ab.data(), k.ab.size());
    return PK_OK;
}

// Host self-check of the trellis tables of the kernel of `layer`: the in-place numbering k_polar_lanes runs on against the
// gather form k_polar_decode runs on, `ntests` random cost vectors, all phases.  *state_bits = index bits of the in-place
// numbering.  PK_OK, or PK_ERR_UNSUPPORTED with the first difference in pk_last_error().
int pk_polar_trellis_selfcheck(const pk_polar *h, int layer, uint64_t seed, int ntests, int *state_bits) {
    if (!h || layer < 0 || layer >= h->code.layers || ntests < 1) return pk_set_error(PK_ERR_ARG, "bad arguments");
    const PkKernelTrellis &k = h->code.kernels[h->code.kid[layer]];
    if (state_bits) *state_bits = k.ip_ok ? k.ip_bits : -1;
    const std::string err = pk_polar_check_inplace(k, seed, ntests);
    if (!err.empty()) return pk_set_error(PK_ERR_UNSUPPORTED, err);
    return PK_OK;
}

// (2^m) x (2^m) extended-BCH polarisation kernel (root bchCoder.cpp:356-389), row-major bytes
int pk_make_ebch_kernel(int m, uint8_t *out) {
    if (m < 3 || m > 6 || !out) return pk_set_error(PK_ERR_ARG, "m must be in [3,6]");
    std::vector<uint8_t> k;
    pk_polar_ebch_kernel(m, k);
    std::memcpy(out, k.data(), k.size());
    return PK_OK;
}

#define PKP_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) { cudaFree(d0); cudaFree(d1); cudaFree(d2); cudaFree(d3); cudaFree(d4); \
            return pk_set_error(PK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); }     \
    } while (0)

// CBinaryEncoder::Encode (Codec.h:52-70; MixedKernelEncoder.cpp:142) for B frames, bits as 0/1 bytes
int pk_polar_encode_batch(pk_polar *h, const uint8_t *info, long B, uint8_t *cw) {
    if (!h || B < 0 || (B && (!info || !cw))) return pk_set_error(PK_ERR_ARG, "bad arguments");
    if (h->device < 0) return pk_set_error(PK_ERR_CUDA, "host-only handle: libpkb200 has no CPU compute path");
    if (!B) return PK_OK;
    void *d0 = nullptr, *d1 = nullptr, *d2 = nullptr, *d3 = nullptr, *d4 = nullptr;
    PKP_CUDA(cudaSetDevice(h->device));
    PKP_CUDA(cudaMalloc(&d0, (size_t)B * h->code.K));
    PKP_CUDA(cudaMalloc(&d1, (size_t)B * h->code.N));
    PKP_CUDA(cudaMemcpyAsync(d0, info, (size_t)B * h->code.K, cudaMemcpyHostToDevice, h->stream));
    const int grid = (int)std::min<long>((B + 3) / 4, 148L * 8);
    k_polar_encode<<<grid, 128, 4 * 2 * h->dev.N0, h->stream>>>(h->dev, (const uint8_t *)d0, B, (uint8_t *)d1);
    ++g_pk_launches;
    PKP_CUDA(cudaGetLastError());
    PKP_CUDA(cudaMemcpyAsync(cw, d1, (size_t)B * h->code.N, cudaMemcpyDeviceToHost, h->stream));
    PKP_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d0); cudaFree(d1);
    return PK_OK;
}

// CKernProcLLR::GetLLRs (KernProc.h:40-60; TrellisKernelProcessor.cpp:234) for B independent kernel blocks of the
// kernel of `layer`: chan [B][l] LLRs of the kernel outputs, u [B][l] the (known) kernel inputs; out [B][l]:
// out[b][p] = LLR of input p given inputs 0..p-1.  Stride 1.
int pk_polar_kernel_llrs(pk_polar *h, int layer, const float *chan, const uint8_t *u, long B, float *out) {
    if (!h || layer < 0 || layer >= h->code.layers || B < 0 || (B && (!chan || !u || !out))) return pk_set_error(PK_ERR_ARG, "bad arguments");
    if (h->device < 0) return pk_set_error(PK_ERR_CUDA, "host-only handle: libpkb200 has no CPU compute path");
    if (!B) return PK_OK;
    const PkPolarKernelDev &k = h->dev.kern[layer];
    const int l = k.size;
    void *d0 = nullptr, *d1 = nullptr, *d2 = nullptr, *d3 = nullptr, *d4 = nullptr;
    PKP_CUDA(cudaSetDevice(h->device));
    PKP_CUDA(cudaMalloc(&d0, (size_t)B * l * 4));
    PKP_CUDA(cudaMalloc(&d1, (size_t)B * l));
    PKP_CUDA(cudaMalloc(&d2, (size_t)B * l * 4));
    PKP_CUDA(cudaMemcpyAsync(d0, chan, (size_t)B * l * 4, cudaMemcpyHostToDevice, h->stream));
    PKP_CUDA(cudaMemcpyAsync(d1, u, (size_t)B * l, cudaMemcpyHostToDevice, h->stream));
    const size_t per = (((size_t)polar_met_floats(k.max_ab) * 4 + 4 * l + 2 * l + 16) + 15) & ~(size_t)15;
    PKP_CUDA(cudaFuncSetAttribute(k_polar_kernel_llr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * per)));
    const int grid = (int)std::min<long>((B + 3) / 4, 148L * 8);
    k_polar_kernel_llr<<<grid, 128, 4 * per, h->stream>>>(k, (const float *)d0, (const uint8_t *)d1, B, (float *)d2);
    ++g_pk_launches;
    PKP_CUDA(cudaGetLastError());
    PKP_CUDA(cudaMemcpyAsync(out, d2, (size_t)B * l * 4, cudaMemcpyDeviceToHost, h->stream));
    PKP_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d0); cudaFree(d1); cudaFree(d2);
    return PK_OK;
}

// CBinarySoftDecoder::Decode (Codec.h:100-118; MixedKernelListDecoder.cpp:211) for B frames, device buffers,
// asynchronous on `stream` (NULL = the handle's stream).  llr [B][N] (positive = bit 0); count [B] list entries
// per frame; inf [B][L][K], cw [B][L][N] (may be NULL), metric [B][L] (may be NULL), best path first.
int pk_polar_decode_batch_dev(pk_polar *h, const float *d_llr, long B, int *d_count, uint8_t *d_inf, uint8_t *d_cw,
                              float *d_metric, void *stream) {
    if (!h || B < 0 || (B && !d_llr)) return pk_set_error(PK_ERR_ARG, "bad arguments");
    if (h->device < 0) return pk_set_error(PK_ERR_CUDA, "host-only handle: libpkb200 has no CPU compute path");
    if (!B) return PK_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return pk_set_error(PK_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    if (h->ln_g) {
        // paths across lanes (one launch in flight per handle: the channel scratch belongs to the handle)
        const cudaError_t e = lanes_launch(h, d_llr, B, d_count, d_inf, d_cw, d_metric, st);
        ++g_pk_launches;
        if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
        return PK_OK;
    }
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_polar_decode, 32 * h->L * h->fpc, h->smem_decode);
    if (per_sm < 1) return pk_set_error(PK_ERR_CUDA, "polar decode kernel does not fit (list size x code length too large for shared memory)");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, h->device);
    const int grid = (int)std::min<long>((B + h->fpc - 1) / h->fpc, (long)prop.multiProcessorCount * per_sm);
    k_polar_decode<<<grid, 32 * h->L * h->fpc, h->smem_decode, st>>>(h->dev, h->L, h->fpc, d_llr, B, d_count, d_inf, d_cw, d_metric);
    ++g_pk_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
    return PK_OK;
}

// same with host buffers (H2D / D2H inside the call)
int pk_polar_decode_batch(pk_polar *h, const float *llr, long B, int *count, uint8_t *inf, uint8_t *cw, float *metric) {
    if (!h || B < 0 || (B && (!llr || !inf))) return pk_set_error(PK_ERR_ARG, "bad arguments");
    if (h->device < 0) return pk_set_error(PK_ERR_CUDA, "host-only handle: libpkb200 has no CPU compute path");
    if (!B) return PK_OK;
    const int N = h->code.N, K = h->code.K, L = h->L;
    void *d0 = nullptr, *d1 = nullptr, *d2 = nullptr, *d3 = nullptr, *d4 = nullptr;
    PKP_CUDA(cudaSetDevice(h->device));
    PKP_CUDA(cudaMalloc(&d0, (size_t)B * N * 4));
    PKP_CUDA(cudaMalloc(&d1, (size_t)B * 4));
    PKP_CUDA(cudaMalloc(&d2, (size_t)B * L * K));
    PKP_CUDA(cudaMalloc(&d3, (size_t)B * L * N));
    PKP_CUDA(cudaMalloc(&d4, (size_t)B * L * 4));
    PKP_CUDA(cudaMemcpyAsync(d0, llr, (size_t)B * N * 4, cudaMemcpyHostToDevice, h->stream));
    PKP_CUDA(cudaMemsetAsync(d2, 0, (size_t)B * L * K, h->stream));
    PKP_CUDA(cudaMemsetAsync(d3, 0, (size_t)B * L * N, h->stream));
    PKP_CUDA(cudaMemsetAsync(d4, 0, (size_t)B * L * 4, h->stream));
    int rc = pk_polar_decode_batch_dev(h, (const float *)d0, B, (int *)d1, (uint8_t *)d2, (uint8_t *)d3, (float *)d4, h->stream);
    if (rc) { cudaFree(d0); cudaFree(d1); cudaFree(d2); cudaFree(d3); cudaFree(d4); return rc; }
    if (count) PKP_CUDA(cudaMemcpyAsync(count, d1, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
    PKP_CUDA(cudaMemcpyAsync(inf, d2, (size_t)B * L * K, cudaMemcpyDeviceToHost, h->stream));
    if (cw) PKP_CUDA(cudaMemcpyAsync(cw, d3, (size_t)B * L * N, cudaMemcpyDeviceToHost, h->stream));
    if (metric) PKP_CUDA(cudaMemcpyAsync(metric, d4, (size_t)B * L * 4, cudaMemcpyDeviceToHost, h->stream));
    PKP_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d0); cudaFree(d1); cudaFree(d2); cudaFree(d3); cudaFree(d4);
    return PK_OK;
}

// ------------------------------------------------------------------ generation mode
static int polar_gen_ws(pk_polar *h) {
    if (h->gen_cap) return PK_OK;
    const long cap = std::max(1024L, (1L << 21) / h->L / std::max(1, h->code.K / 128));
    cudaError_t e = cudaMalloc(&h->g_info, (size_t)cap * h->code.K);
    if (e == cudaSuccess) e = cudaMalloc(&h->g_inf, (size_t)cap * h->L * h->code.K);
    if (e == cudaSuccess) e = cudaMalloc(&h->g_llr, (size_t)cap * h->code.N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->g_cnt, (size_t)cap * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->g_tot, 8 * 8);
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_tot, 8 * 8);
    if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, std::string("pk_polar generation workspaces: ") + cudaGetErrorString(e));
    h->gen_cap = cap;
    return PK_OK;
}
static double polar_sigma(const pk_polar *h, double ebn0_db) {
    return sqrt(1.0 / (2.0 * ((double)h->code.K / (double)h->code.N) * pow(10.0, ebn0_db / 10.0)));   // Simulator.cpp:104 (SetEbN0)
}

// The frames generation mode draws, written out (device buffers, any of d_info [B][K] / d_cw [B][N] may be NULL; d_llr [B][N])
int pk_polar_generate_frames_dev(pk_polar *h, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame, long nframes,
                                 uint8_t *d_info, uint8_t *d_cw, float *d_llr, void *stream) {
    if (!h || nframes < 0 || (nframes && !d_llr)) return pk_set_error(PK_ERR_ARG, "bad arguments");
    if (h->device < 0) return pk_set_error(PK_ERR_CUDA, "host-only handle: libpkb200 has no CPU compute path");
    if (!nframes) return PK_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return pk_set_error(PK_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const int grid = (int)std::min<long>((nframes + 3) / 4, 148L * 8);
    k_polar_generate<<<grid, 128, 4 * 3 * h->dev.N0, st>>>(h->dev, polar_sigma(h, ebn0_db), seed, first_frame, (uint32_t)snr_index, nframes, d_info, d_cw, d_llr);
    ++g_pk_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
    return PK_OK;
}

// Monte-Carlo body for frames [first_frame, first_frame + nframes) of SNR point snr_index, fully on the device:
// generate -> SC / SC-list decode -> compare the best path's information vector with the transmitted one.  d_totals
// (8 x u64, pk_point_result layout: frames, frame_errors, bit_errors = information-bit errors) is ACCUMULATED into.
// Asynchronous on `stream` (NULL = the handle's).  Results do not depend on how the frame range is split.
int pk_polar_run_frames_dev(pk_polar *h, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame, long nframes,
                            uint64_t *d_totals, void *stream) {
    if (!h || nframes < 0 || !d_totals) return pk_set_error(PK_ERR_ARG, "bad arguments");
    if (h->device < 0) return pk_set_error(PK_ERR_CUDA, "host-only handle: libpkb200 has no CPU compute path");
    if (!nframes) return PK_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return pk_set_error(PK_ERR_CUDA, "cudaSetDevice failed");
    int rc = polar_gen_ws(h);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    for (long off = 0; off < nframes; off += h->gen_cap) {
        const long nb = std::min(h->gen_cap, nframes - off);
        rc = pk_polar_generate_frames_dev(h, ebn0_db, snr_index, seed, first_frame + (uint64_t)off, nb, h->g_info, nullptr, h->g_llr, st);
        if (rc) return rc;
        rc = pk_polar_decode_batch_dev(h, h->g_llr, nb, h->g_cnt, h->g_inf, nullptr, nullptr, st);
        if (rc) return rc;
        const int grid = (int)std::min<long>((nb + 7) / 8, 148L * 8);
        k_polar_compare<<<grid, 256, 0, st>>>(h->code.K, h->L, h->g_info, h->g_inf, h->g_cnt, nb, (unsigned long long *)d_totals);
        ++g_pk_launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
    }
    return PK_OK;
}

// Same, synchronous, host result (adds into *totals).
int pk_polar_run_frames(pk_polar *h, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame, long nframes,
                        pk_point_result *totals) {
    if (!h || nframes < 0 || !totals) return pk_set_error(PK_ERR_ARG, "bad arguments");
    if (h->device < 0) return pk_set_error(PK_ERR_CUDA, "host-only handle: libpkb200 has no CPU compute path");
    if (!nframes) return PK_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return pk_set_error(PK_ERR_CUDA, "cudaSetDevice failed");
    int rc = polar_gen_ws(h);
    if (rc) return rc;
    cudaError_t e = cudaMemsetAsync(h->g_tot, 0, 64, h->stream);
    if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
    rc = pk_polar_run_frames_dev(h, ebn0_db, snr_index, seed, first_frame, nframes, (uint64_t *)h->g_tot, h->stream);
    if (rc) return rc;
    e = cudaMemcpyAsync(h->h_tot, h->g_tot, 64, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
    uint64_t *t = reinterpret_cast<uint64_t *>(totals);
    for (int i = 0; i < 6; ++i) t[i] += h->h_tot[i];
    return PK_OK;
}

// host copies of the generated frames (tests): info [B][K], cw [B][N] (any may be NULL), llr [B][N]
int pk_polar_generate_frames(pk_polar *h, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame, long nframes,
                             uint8_t *info, uint8_t *cw, float *llr) {
    if (!h || nframes < 0 || (nframes && !llr)) return pk_set_error(PK_ERR_ARG, "bad arguments");
    if (h->device < 0) return pk_set_error(PK_ERR_CUDA, "host-only handle: libpkb200 has no CPU compute path");
    if (!nframes) return PK_OK;
    const int N = h->code.N, K = h->code.K;
    void *d0 = nullptr, *d1 = nullptr, *d2 = nullptr, *d3 = nullptr, *d4 = nullptr;
    PKP_CUDA(cudaSetDevice(h->device));
    PKP_CUDA(cudaMalloc(&d0, (size_t)nframes * K));
    PKP_CUDA(cudaMalloc(&d1, (size_t)nframes * N));
    PKP_CUDA(cudaMalloc(&d2, (size_t)nframes * N * 4));
    int rc = pk_polar_generate_frames_dev(h, ebn0_db, snr_index, seed, first_frame, nframes, (uint8_t *)d0, (uint8_t *)d1, (float *)d2, h->stream);
    if (rc) { cudaFree(d0); cudaFree(d1); cudaFree(d2); return rc; }
    if (info) PKP_CUDA(cudaMemcpyAsync(info, d0, (size_t)nframes * K, cudaMemcpyDeviceToHost, h->stream));
    if (cw) PKP_CUDA(cudaMemcpyAsync(cw, d1, (size_t)nframes * N, cudaMemcpyDeviceToHost, h->stream));
    PKP_CUDA(cudaMemcpyAsync(llr, d2, (size_t)nframes * N * 4, cudaMemcpyDeviceToHost, h->stream));
    PKP_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d0); cudaFree(d1); cudaFree(d2);
    return PK_OK;
}

}  // extern "C"
