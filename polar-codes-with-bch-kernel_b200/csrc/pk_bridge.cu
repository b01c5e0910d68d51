// pk_bridge.cu -- Kaneko decoding as a polarisation-kernel processor (what the repository's name
// promises and the reference never wrote: KanekoKernelProcessor does not derive from CKernProcLLR).
//
// CKernProcLLR::GetLLRs(Stride, phase, known, chanLLR, out, ..) (headers/external/KernProc.h:40-60) for an
// extended-BCH kernel K (l = 2^m, makeMatrix of the root bchCoder.cpp:356-389):
//     out = M[1] - M[0],   M[v] = min over codewords c of the coset { sum_{r<phase} u_r K_r + v K_phase + span(K_{phase+1..l-1}) }
//                                 of  sum_{j : c_j != hard decision_j} |chanLLR_j|
// which is exactly pStateMetric0[1] - pStateMetric0[0] of the trellis processor (out/external/TrellisKernelProcessor.cpp:292),
// whose Viterbi search is exponential in the trellis state count and refuses kernels of size >= 64 (:71-72).
// Here each M[v] is a maximum-likelihood decoding of an extended BCH code: the row space R of rows phase+1..l-1 is
// split as  R = ext(C_t) + span(e extra rows)  with the largest narrow-sense BCH code C_t whose extension lies in R;
// for each of the 2^e combinations of the extra rows (and both v) the received word is offset and decoded by the
// Kaneko search of libpkb200 with the EXACT stopping rules (pk_kaneko_create_ext, rules = 2: the decision is the ML
// codeword).  Tails too small for a BCH code are enumerated (dimension <= 12); the even-weight code has a closed form.
// Metrics are fp32 sums in column order, like the Viterbi recursion's accumulation along the sections.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/pk_capi.h"
#include "pk_kernels.h"
#include "pk_polar.h"

int pk_set_error(int code, const std::string &msg);   // pk_capi.cu
extern unsigned long long g_pk_launches;

namespace {

enum { KP_ENUM = 0, KP_EVEN = 1, KP_BCH = 2 };
constexpr int KP_MAX_EXTRA = 12;

struct PhasePlan {
    int mode = KP_ENUM, t = 0, nextra = 0;
    uint64_t extra[KP_MAX_EXTRA] = {};   // rows enumerated on top of the decoded code (ENUM: a basis of the whole tail)
};

// offset of candidate (v, combo) of block b:  sum_{r<phase} u_r K_r  +  v K_phase  +  sum_{i in combo} extra_i
__device__ __forceinline__ uint64_t kp_offset(const uint64_t *rows, const uint8_t *u, long ustride, int phase, int v, uint32_t combo,
                                              const uint64_t *extra, int nextra) {
    uint64_t o = 0;
    for (int r = 0; r < phase; ++r)
        if (u[(long)r * ustride]) o ^= rows[r];
    if (v) o ^= rows[phase];
    for (int i = 0; i < nextra; ++i)
        if ((combo >> i) & 1u) o ^= extra[i];
    return o;
}
// sum over the columns j (ascending) where `diff` is set of |chan_j|, fp32
__device__ __forceinline__ float kp_metric(const float *chan, long cstride, int l, uint64_t diff) {
    float m = 0.0f;
    for (int j = 0; j < l; ++j)
        if ((diff >> j) & 1ull) m += fabsf(chan[(long)j * cstride]);
    return m;
}
__device__ __forceinline__ uint64_t kp_hard(const float *chan, long cstride, int l) {
    uint64_t h = 0;
    for (int j = 0; j < l; ++j)
        if (chan[(long)j * cstride] < 0.0f) h |= 1ull << j;
    return h;
}

// ENUM / EVEN phases: one thread per element, everything in registers
__global__ void k_bridge_small(const uint64_t *rows, PhasePlan pl, int l, int phase, const float *chan, long cstride, long celem,
                               const uint8_t *u, long ustride, long uelem, long B, float *out) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *ch = chan + b * celem;
    const uint8_t *ub = u + b * uelem;
    const uint64_t hard = kp_hard(ch, cstride, l);
    float M[2];
    for (int v = 0; v < 2; ++v) {
        const uint64_t o = kp_offset(rows, ub, ustride, phase, v, 0, pl.extra, 0);
        float best = HUGE_VALF;
        if (pl.mode == KP_EVEN) {
            // tail = all even-weight words: flip nothing, or the least reliable position when the parity is odd
            const uint64_t d = hard ^ o;
            if (__popcll(d) & 1) {
                float mn = HUGE_VALF;
                for (int j = 0; j < l; ++j) mn = fminf(mn, fabsf(ch[(long)j * cstride]));
                best = mn;
            } else {
                best = 0.0f;
            }
        } else {
            for (uint32_t c = 0; c < (1u << pl.nextra); ++c) {
                uint64_t w = o;
                for (int i = 0; i < pl.nextra; ++i)
                    if ((c >> i) & 1u) w ^= pl.extra[i];
                best = fminf(best, kp_metric(ch, cstride, l, w ^ hard));
            }
        }
        M[v] = best;
    }
    out[b] = M[1] - M[0];
}

// BCH phases, step 1: the offset received words, in the Kaneko frame layout (position p < l-1 = column p+1, the
// overall-parity position l-1 = column 0), BPSK convention of the Kaneko path: y > 0 <=> bit 1
__global__ void k_bridge_prep(const uint64_t *rows, PhasePlan pl, int l, int phase, const float *chan, long cstride, long celem,
                              const uint8_t *u, long ustride, long uelem, long B, double *y) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per = 2L << pl.nextra;
    if (idx >= B * per) return;
    const long b = idx / per;
    const int q = (int)(idx - b * per), v = q >> pl.nextra;
    const uint32_t combo = (uint32_t)q & ((1u << pl.nextra) - 1u);
    const float *ch = chan + b * celem;
    const uint64_t o = kp_offset(rows, u + b * uelem, ustride, phase, v, combo, pl.extra, pl.nextra);
    double *yy = y + idx * l;
    for (int j = 0; j < l; ++j) {
        const float c = ch[(long)j * cstride];
        const uint64_t bit = ((c < 0.0f) ? 1ull : 0ull) ^ ((o >> j) & 1ull);
        const int p = (j == 0) ? l - 1 : j - 1;
        yy[p] = bit ? (double)fabsf(c) : -(double)fabsf(c);
    }
}
// step 3: metric of every decided codeword, minimum per v, difference
__global__ void k_bridge_metric(const uint64_t *rows, PhasePlan pl, int l, int phase, const float *chan, long cstride, long celem,
                                const uint8_t *u, long ustride, long uelem, long B, const uint8_t *decided, const pk_frame_rec *recs,
                                float *out, unsigned long long *bad) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *ch = chan + b * celem;
    const uint64_t hard = kp_hard(ch, cstride, l);
    const long per = 2L << pl.nextra;
    float M[2] = {HUGE_VALF, HUGE_VALF};
    unsigned long long nb = 0;
    for (int q = 0; q < (int)per; ++q) {
        const int v = q >> pl.nextra;
        const uint32_t combo = (uint32_t)q & ((1u << pl.nextra) - 1u);
        const long f = b * per + q;
        if (recs[f].flags & (PK_FLAG_TRUNCATED | PK_FLAG_NO_DECISION)) { ++nb; if (recs[f].flags & PK_FLAG_NO_DECISION) continue; }
        const uint8_t *d = decided + f * l;
        uint64_t cw = 0;   // decided codeword of the extended BCH code, back in column order
        for (int j = 0; j < l; ++j) {
            const int p = (j == 0) ? l - 1 : j - 1;
            if (d[p]) cw |= 1ull << j;
        }
        const uint64_t o = kp_offset(rows, u + b * uelem, ustride, phase, v, combo, pl.extra, pl.nextra);
        M[v] = fminf(M[v], kp_metric(ch, cstride, l, cw ^ o ^ hard));
    }
    out[b] = M[1] - M[0];
    if (nb) atomicAdd(bad, nb);
}

// ---- GF(2) helpers on 64-bit row masks
int gf2_rank_insert(std::vector<uint64_t> &basis, uint64_t v) {   // returns 1 if v enlarged the span
    for (uint64_t b : basis) v = std::min(v, v ^ b);
    if (!v) return 0;
    basis.push_back(v);
    std::sort(basis.begin(), basis.end(), std::greater<uint64_t>());
    // keep it reduced: re-reduce every vector by the others (small sizes)
    for (size_t i = 0; i < basis.size(); ++i)
        for (size_t j = 0; j < basis.size(); ++j)
            if (i != j && (basis[j] ^ basis[i]) < basis[j]) basis[j] ^= basis[i];
    std::sort(basis.begin(), basis.end(), std::greater<uint64_t>());
    return 1;
}
bool gf2_in_span(const std::vector<uint64_t> &basis, uint64_t v) {
    for (uint64_t b : basis) v = std::min(v, v ^ b);
    return v == 0;
}

}  // namespace

struct pk_kproc {
    int m = 0, l = 0, device = 0;
    long max_trials = 0;
    std::vector<uint64_t> rows;
    std::vector<PhasePlan> plan;
    std::map<int, pk_code *> codes;
    std::map<int, pk_kaneko *> decs;
    uint64_t *d_rows = nullptr;
    cudaStream_t stream = nullptr;
    unsigned long long *d_bad = nullptr;
};

#define PKB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) return pk_set_error(PK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

extern "C" {

int pk_kproc_create(int m, int device, long max_trials, int enum_dim, pk_kproc **out) {
    if (!out) return pk_set_error(PK_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (m < 3 || m > 6) return pk_set_error(PK_ERR_ARG, "pk_kproc_create: kernel size 2^m with m in [3,6]");
    if (enum_dim < 0 || enum_dim > KP_MAX_EXTRA) enum_dim = KP_MAX_EXTRA;   // row tails up to this dimension are enumerated
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return pk_set_error(PK_ERR_CUDA, "no CUDA device: libpkb200 has no CPU path");
    if (device < 0 || device >= ndev) return pk_set_error(PK_ERR_ARG, "bad device ordinal");
    pk_kproc *h = new (std::nothrow) pk_kproc;
    if (!h) return pk_set_error(PK_ERR_ALLOC, "out of memory");
    h->m = m; h->l = 1 << m; h->device = device;
    h->max_trials = max_trials > 0 ? max_trials : (1L << 22);
    const int l = h->l, n = l - 1;
    std::vector<uint8_t> K;
    pk_polar_ebch_kernel(m, K);
    h->rows.assign(l, 0);
    for (int r = 0; r < l; ++r)
        for (int c = 0; c < l; ++c)
            if (K[(size_t)r * l + c]) h->rows[r] |= 1ull << c;
    // extended BCH codes ext(C_t) in column order: column p+1 = coefficient of x^p, column 0 = overall parity
    // (several designed t give the same code; the largest one with sm_100a kernels decodes it best)
    struct Cand { int t, k; std::vector<uint64_t> gen; };
    std::vector<Cand> cands;
    for (int t = 1; t < (1 << (m - 1)); ++t) {
        if (!pk_find_kernels(m, t)) continue;
        pk_code *c = nullptr;
        if (pk_code_create_host(m, t, &c) != PK_OK) continue;
        int nn = 0, kk = 0, gs = 0;
        pk_code_info(c, &nn, &kk, nullptr, &gs, nullptr);
        std::vector<uint8_t> g(gs);
        pk_code_info(c, nullptr, nullptr, nullptr, nullptr, g.data());
        pk_code_destroy(c);
        Cand cd;
        cd.t = t; cd.k = kk;
        for (int s = 0; s < kk; ++s) {
            uint64_t w = 0;
            int wt = 0;
            for (int i = 0; i < gs; ++i)
                if (g[i]) { w |= 1ull << (i + s + 1); ++wt; }
            if (wt & 1) w |= 1ull;
            cd.gen.push_back(w);
        }
        if (!cands.empty() && cands.back().k == kk) cands.back() = cd;   // same code, larger designed t
        else cands.push_back(cd);
    }
    h->plan.assign(l, PhasePlan());
    std::string err;
    for (int ph = 0; ph < l && err.empty(); ++ph) {
        PhasePlan &pl = h->plan[ph];
        std::vector<uint64_t> tail;   // reduced basis of span(rows ph+1 .. l-1)
        for (int r = ph + 1; r < l; ++r) gf2_rank_insert(tail, h->rows[r]);
        const int dim = (int)tail.size();
        // the even-weight code (dimension l-1)?
        bool even = dim == l - 1;
        for (int j = 1; j < l && even; ++j) even = gf2_in_span(tail, 1ull | (1ull << j));
        std::vector<uint64_t> have;   // span already covered by the decoded code
        if (even) {
            pl.mode = KP_EVEN;
            continue;
        }
        pl.mode = KP_ENUM;
        if (dim > enum_dim) {
            // largest extended BCH code inside the tail for which sm_100a kernels exist
            for (const Cand &cd : cands) {
                bool inside = true;
                for (uint64_t w : cd.gen) inside = inside && gf2_in_span(tail, w);
                if (!inside) continue;
                pl.mode = KP_BCH;
                pl.t = cd.t;
                for (uint64_t w : cd.gen) gf2_rank_insert(have, w);
                break;
            }
            if (pl.mode != KP_BCH && dim > KP_MAX_EXTRA) { err = "phase " + std::to_string(ph) + ": no extended BCH code inside the row tail"; break; }
        }
        // rows of the kernel that complete the covered span to the whole tail
        for (int r = ph + 1; r < l; ++r) {
            if (gf2_rank_insert(have, h->rows[r])) {
                if (pl.nextra >= KP_MAX_EXTRA) { err = "phase " + std::to_string(ph) + ": more than 12 rows left outside the decoded code"; break; }
                pl.extra[pl.nextra++] = h->rows[r];
            }
        }
    }
    if (!err.empty()) { delete h; return pk_set_error(PK_ERR_UNSUPPORTED, "pk_kproc_create: " + err); }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_rows, (size_t)l * 8);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_rows, h->rows.data(), (size_t)l * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_bad, 8);
    if (e != cudaSuccess) { pk_kproc_destroy(h); return pk_set_error(PK_ERR_CUDA, std::string("pk_kproc_create: ") + cudaGetErrorString(e)); }
    for (int ph = 0; ph < l; ++ph) {
        const int t = h->plan[ph].t;
        if (h->plan[ph].mode != KP_BCH || h->decs.count(t)) continue;
        pk_code *c = nullptr;
        pk_kaneko *d = nullptr;
        int rc = pk_code_create(m, t, device, &c);
        if (rc == PK_OK) { h->codes[t] = c; rc = pk_kaneko_create_ext(c, 0.5, -1, h->max_trials, 1, 2, &d); }
        if (rc != PK_OK) { pk_kproc_destroy(h); return rc; }
        h->decs[t] = d;
    }
    (void)n;
    *out = h;
    return PK_OK;
}

void pk_kproc_destroy(pk_kproc *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto &kv : h->decs) pk_kaneko_destroy(kv.second);
    for (auto &kv : h->codes) pk_code_destroy(kv.second);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaFree(h->d_rows);
    cudaFree(h->d_bad);
    delete h;
}

// size = l; per phase: mode (0 enumeration, 1 even-weight closed form, 2 Kaneko search), t of the decoded BCH code, rows enumerated
int pk_kproc_info(const pk_kproc *h, int *size, int *mode, int *t, int *nextra) {
    if (!h) return pk_set_error(PK_ERR_ARG, "NULL handle");
    if (size) *size = h->l;
    for (int p = 0; p < h->l; ++p) {
        if (mode) mode[p] = h->plan[p].mode;
        if (t) t[p] = h->plan[p].t;
        if (nextra) nextra[p] = h->plan[p].nextra;
    }
    return PK_OK;
}

// device-side core: element b reads chan[b*celem + j*cstride], u[b*uelem + r*ustride]; out[b]
static int kproc_phase(pk_kproc *h, int phase, const float *d_chan, long cstride, long celem, const uint8_t *d_u, long ustride, long uelem,
                       long B, float *d_out, long *truncated) {
    const PhasePlan &pl = h->plan[phase];
    const int l = h->l;
    const int thr = 128;
    if (pl.mode != KP_BCH) {
        k_bridge_small<<<(unsigned)((B + thr - 1) / thr), thr, 0, h->stream>>>(h->d_rows, pl, l, phase, d_chan, cstride, celem, d_u, ustride, uelem, B, d_out);
        ++g_pk_launches;
        PKB_CUDA(cudaGetLastError());
        return PK_OK;
    }
    const long per = 2L << pl.nextra, F = B * per;
    double *d_y = nullptr;
    uint8_t *d_dec = nullptr;
    pk_frame_rec *d_rec = nullptr;
    cudaError_t e = cudaMalloc(&d_y, (size_t)F * l * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_dec, (size_t)F * l);
    if (e == cudaSuccess) e = cudaMalloc(&d_rec, (size_t)F * sizeof(pk_frame_rec));
    int rc = PK_OK;
    if (e != cudaSuccess) rc = pk_set_error(PK_ERR_CUDA, std::string("pk_kproc: ") + cudaGetErrorString(e));
    if (rc == PK_OK) {
        cudaMemsetAsync(h->d_bad, 0, 8, h->stream);
        cudaMemsetAsync(d_dec, 0, (size_t)F * l, h->stream);
        k_bridge_prep<<<(unsigned)((F + thr - 1) / thr), thr, 0, h->stream>>>(h->d_rows, pl, l, phase, d_chan, cstride, celem, d_u, ustride, uelem, B, d_y);
        ++g_pk_launches;
        rc = pk_kaneko_decode_batch_dev(h->decs[pl.t], d_y, F, d_dec, nullptr, d_rec, nullptr, h->stream);
    }
    if (rc == PK_OK) {
        k_bridge_metric<<<(unsigned)((B + thr - 1) / thr), thr, 0, h->stream>>>(h->d_rows, pl, l, phase, d_chan, cstride, celem, d_u, ustride, uelem, B, d_dec, d_rec, d_out, h->d_bad);
        ++g_pk_launches;
        unsigned long long bad = 0;
        e = cudaMemcpyAsync(&bad, h->d_bad, 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = pk_set_error(PK_ERR_CUDA, std::string("pk_kproc: ") + cudaGetErrorString(e));
        if (truncated) *truncated += (long)bad;
    } else {
        cudaStreamSynchronize(h->stream);
    }
    cudaFree(d_y); cudaFree(d_dec); cudaFree(d_rec);
    return rc;
}

// CKernProcLLR::GetLLRs(Stride, phase, pKnownInputSymbols [l][Stride], pChannelLLRs [l][Stride], pLLRs [Stride]) (KernProc.h:40-60):
// host buffers; the known inputs of rows >= phase are not read.  *truncated (may be NULL) counts the searches that were
// stopped by the handle's trial budget (their minimum is then over the codewords found so far).
int pk_kproc_get_llrs(pk_kproc *h, int stride, int phase, const uint8_t *known, const float *chan, float *out, long *truncated) {
    if (!h || stride < 1 || phase < 0 || phase >= h->l || !known || !chan || !out) return pk_set_error(PK_ERR_ARG, "bad arguments");
    const int l = h->l;
    PKB_CUDA(cudaSetDevice(h->device));
    float *d_chan = nullptr, *d_out = nullptr;
    uint8_t *d_u = nullptr;
    PKB_CUDA(cudaMalloc(&d_chan, (size_t)l * stride * 4));
    cudaError_t e = cudaMalloc(&d_u, (size_t)l * stride);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)stride * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_chan, chan, (size_t)l * stride * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_u, known, (size_t)l * stride, cudaMemcpyHostToDevice, h->stream);
    int rc = e == cudaSuccess ? PK_OK : pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
    if (truncated) *truncated = 0;
    if (rc == PK_OK) rc = kproc_phase(h, phase, d_chan, stride, 1, d_u, stride, 1, stride, d_out, truncated);
    if (rc == PK_OK) {
        e = cudaMemcpyAsync(out, d_out, (size_t)stride * 4, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(d_chan); cudaFree(d_u); cudaFree(d_out);
    return rc;
}

// All phases of B independent kernel blocks (the layout of pk_polar_kernel_llrs): chan [B][l], u [B][l] -> out [B][l],
// out[b][p] = LLR of input p given inputs u[b][0..p).
int pk_kproc_kernel_llrs(pk_kproc *h, const float *chan, const uint8_t *u, long B, float *out, long *truncated) {
    if (!h || B < 0 || (B && (!chan || !u || !out))) return pk_set_error(PK_ERR_ARG, "bad arguments");
    if (truncated) *truncated = 0;
    if (!B) return PK_OK;
    const int l = h->l;
    PKB_CUDA(cudaSetDevice(h->device));
    float *d_chan = nullptr, *d_out = nullptr, *d_all = nullptr;
    uint8_t *d_u = nullptr;
    PKB_CUDA(cudaMalloc(&d_chan, (size_t)B * l * 4));
    cudaError_t e = cudaMalloc(&d_u, (size_t)B * l);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)B * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_all, (size_t)B * l * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_chan, chan, (size_t)B * l * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_u, u, (size_t)B * l, cudaMemcpyHostToDevice, h->stream);
    int rc = e == cudaSuccess ? PK_OK : pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
    for (int ph = 0; ph < l && rc == PK_OK; ++ph) {
        rc = kproc_phase(h, ph, d_chan, 1, l, d_u, 1, l, B, d_out, truncated);
        if (rc == PK_OK) {
            e = cudaMemcpy2DAsync(d_all + ph, (size_t)l * 4, d_out, 4, 4, (size_t)B, cudaMemcpyDeviceToDevice, h->stream);
            if (e != cudaSuccess) rc = pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
        }
    }
    if (rc == PK_OK) {
        e = cudaMemcpyAsync(out, d_all, (size_t)B * l * 4, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = pk_set_error(PK_ERR_CUDA, cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(h->stream);
    }
    cudaFree(d_chan); cudaFree(d_u); cudaFree(d_out); cudaFree(d_all);
    return rc;
}

}  // extern "C"
