// pk_capi.cu -- the C ABI (include/pk_capi.h) over the sm_100a kernels.
// No CPU fallback lives here: every compute entry point launches CUDA kernels or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/pk_capi.h"
#include "pk_code.h"
#include "pk_kernels.h"

unsigned long long g_pk_launches = 0;

namespace {
thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define PK_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(PK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
    } while (0)

template <class Tp>
int upload(pk_code *c, const std::vector<Tp> &h, const Tp **dptr) {
    *dptr = nullptr;
    if (h.empty()) return PK_OK;
    void *d = nullptr;
    PK_CUDA(cudaMalloc(&d, h.size() * sizeof(Tp)));
    c->dev_allocs.push_back(d);
    PK_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(Tp), cudaMemcpyHostToDevice));
    *dptr = static_cast<const Tp *>(d);
    return PK_OK;
}
}  // namespace

// shared with pk_polar.cu
int pk_set_error(int code, const std::string &msg) { return fail(code, msg); }

struct pk_kaneko {
    pk_code *code = nullptr;
    PkKanekoParams kp{};
    int mode = PK_MODE_ALG;    // fixed at creation from the code's table switches (pk_code_set_lut)
    PkLaunchGeom geom4[4]{};   // replay phase A/B, generation phase A/B
    cudaStream_t stream[2] = {nullptr, nullptr};
    // one control block + parked-frame list per concurrent launch slot:
    // slots 0/1 = the handle's two pipeline streams, slot 2 = caller-supplied streams
    PkPhaseBlock *d_ctl = nullptr;            // [3]: control words + mega slots, zeroed before every launch pair
    PkLongRec *d_longs[3] = {nullptr, nullptr, nullptr};   // parked frames: [long_cap] main, [PK_HUGE_CAP] huge, [long_cap] late
    uint32_t *d_mega_bits[3] = {nullptr, nullptr, nullptr};   // clean-chunk bitmaps of the mega slots
    uint32_t epoch = 0;                       // launch counter (PkKanekoParams::epoch)
    uint32_t *d_zscr[3] = {nullptr, nullptr, nullptr};   // root-word scratch of the large-code phase B
    long long_cap = 1L << 18;                 // frames per launch pair (and capacity of a list)
    unsigned long long *d_totals = nullptr;   // [8]
    unsigned long long *h_totals = nullptr;   // pinned [8]
    // host-batch pipeline, one set per stream
    long chunk = 0;
    int next_slot = 0;          // stream / buffer set of the next chunk (alternates across calls as well)
    bool pending = false;       // asynchronous batches enqueued since the last pk_kaneko_wait
    cudaEvent_t ev_zero = nullptr;
    double *d_y[2] = {nullptr, nullptr};
    uint8_t *d_dec[2] = {nullptr, nullptr};
    uint32_t *d_tr[2] = {nullptr, nullptr};
    pk_frame_rec *d_rec[2] = {nullptr, nullptr};
    // generation-mode scratch (run_point)
    pk_frame_rec *d_grec = nullptr;
    pk_frame_rec *h_grec = nullptr;
    long grec_cap = 0;
};

// frame length: the BCH length, plus the overall-parity position of an extended code (pk_kaneko_create_ext)
static inline int frame_len(const pk_kaneko *d) { return d->code->n + d->kp.ext; }

extern "C" {

const char *pk_last_error(void) { return g_err.c_str(); }

int pk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

uint64_t pk_launch_count(void) { return g_pk_launches; }
void pk_launch_count_reset(void) { g_pk_launches = 0; }

// ------------------------------------------------------------------ code
int pk_code_create(int m, int t, int device, pk_code **out) {
    if (!out) return fail(PK_ERR_ARG, "out is NULL");
    *out = nullptr;
    pk_code *c = new (std::nothrow) pk_code;
    if (!c) return fail(PK_ERR_ALLOC, "out of memory");
    std::string msg = pk_code_build_host(*c, m, t);
    if (!msg.empty()) {
        delete c;
        return fail(PK_ERR_ARG, msg);
    }
    c->device = device;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        delete c;
        return fail(PK_ERR_CUDA, "no CUDA device: libpkb200 has no CPU path");
    }
    if (device < 0 || device >= ndev) {
        delete c;
        return fail(PK_ERR_ARG, "bad device ordinal");
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        delete c;
        return fail(PK_ERR_CUDA, cudaGetErrorString(e));
    }
    int rc;
    if ((rc = upload(c, c->mul, &c->dev.mul)) || (rc = upload(c, c->xoff, &c->dev.xoff)) ||
        (rc = upload(c, c->hcol, &c->dev.hcol)) || (rc = upload(c, c->rcol, &c->dev.rcol)) ||
        (rc = upload(c, c->lut, &c->dev.lut)) || (rc = upload(c, c->gmask, &c->dev.gmask))) {
        pk_code_destroy(c);
        return rc;
    }
    if (c->ct) {
        const PkClassTable &ct = *c->ct;
        if ((rc = upload(c, ct.norm, &c->dev.ct_norm)) || (rc = upload(c, ct.col, &c->dev.ct_col)) || (rc = upload(c, ct.logt, &c->dev.ct_log)) ||
            (rc = upload(c, ct.bits, &c->dev.ct_bits)) ||
            (rc = upload(c, ct.hash, reinterpret_cast<const uint64_t **>(&c->dev.ct_hash)))) {
            pk_code_destroy(c);
            return rc;
        }
        {
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeLinear;
            rd.res.linear.devPtr = const_cast<uint32_t *>(c->dev.ct_bits);
            rd.res.linear.desc = cudaCreateChannelDesc<unsigned int>();
            rd.res.linear.sizeInBytes = ct.bits.size() * sizeof(uint32_t);
            cudaTextureDesc td{};
            td.readMode = cudaReadModeElementType;
            cudaTextureObject_t tex = 0;
            cudaError_t te = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
            if (te != cudaSuccess) {
                pk_code_destroy(c);
                return fail(PK_ERR_CUDA, std::string("cudaCreateTextureObject: ") + cudaGetErrorString(te));
            }
            c->dev.ct_tex = (unsigned long long)tex;
        }
        c->dev.ct_hshift = 32u - (uint32_t)ct.hbits;
        c->dev.ct_hmask = (1u << ct.hbits) - 1u;
        for (size_t i = 0; i < 8; ++i) c->dev.ct_mult[i] = i < ct.mult.size() ? ct.mult[i] : 0u;
    }
    c->dev.k = c->k;
    c->dev.nk = c->nk;
    *out = c;
    return PK_OK;
}

// Host tables only (device = -1): lets CPU-only tooling and tests inspect g(x), the field
// tables, the kernel matrix and the coset table.  Every compute call on such a handle fails.
int pk_code_create_host(int m, int t, pk_code **out) {
    if (!out) return fail(PK_ERR_ARG, "out is NULL");
    *out = nullptr;
    pk_code *c = new (std::nothrow) pk_code;
    if (!c) return fail(PK_ERR_ALLOC, "out of memory");
    std::string msg = pk_code_build_host(*c, m, t);
    if (!msg.empty()) {
        delete c;
        return fail(PK_ERR_ARG, msg);
    }
    c->device = -1;
    *out = c;
    return PK_OK;
}

void pk_code_destroy(pk_code *c) {
    if (!c) return;
    if (c->device >= 0) {
        cudaSetDevice(c->device);
        if (c->dev.ct_tex) cudaDestroyTextureObject((cudaTextureObject_t)c->dev.ct_tex);
        for (void *p : c->dev_allocs) cudaFree(p);
    }
    delete c;
}

int pk_code_coset_table(const pk_code *c, uint16_t *out, long *nentries) {
    if (!c) return fail(PK_ERR_ARG, "code is NULL");
    if (nentries) *nentries = (long)c->lut.size();
    if (out && !c->lut.empty()) std::memcpy(out, c->lut.data(), c->lut.size() * sizeof(uint16_t));
    return PK_OK;
}

int pk_code_info(const pk_code *c, int *n, int *k, int *d, int *gsize, uint8_t *g_out) {
    if (!c) return fail(PK_ERR_ARG, "code is NULL");
    if (n) *n = c->n;
    if (k) *k = c->k;
    if (d) *d = 2 * c->t + 1;   // main.cpp:94
    if (gsize) *gsize = c->gsize;
    if (g_out) std::memcpy(g_out, c->g.data(), c->g.size());
    return PK_OK;
}

int pk_code_tables(const pk_code *c, uint64_t *antilog_out, uint64_t *log_out) {
    if (!c) return fail(PK_ERR_ARG, "code is NULL");
    if (antilog_out)
        for (int i = 0; i < c->n; ++i) antilog_out[i] = c->alog[i];
    if (log_out) {
        log_out[0] = (uint64_t)LONG_MAX;   // main.cpp:68
        for (int i = 1; i <= c->n; ++i) log_out[i] = c->log[i];
    }
    return PK_OK;
}

// 0: algebraic decoding only; 1: coset table (n-k <= 16); 2: cyclic-class table (wide search of the mid-size codes)
int pk_code_uses_lut(const pk_code *c) { return !c ? 0 : c->use_lut ? 1 : c->use_ct ? 2 : 0; }

int pk_code_set_lut(pk_code *c, int enable) {
    if (!c) return fail(PK_ERR_ARG, "code is NULL");
    if (enable && c->lut.empty() && !c->ct)
        return fail(PK_ERR_UNSUPPORTED, "no lookup table for this code (coset table: n-k <= 16, t*m <= 15; class table: see pk_code.h)");
    c->use_lut = enable != 0 && !c->lut.empty();
    c->use_ct = enable != 0 && !c->use_lut && (bool)c->ct;
    return PK_OK;
}

// Differential self-check of the class table on the host: `ntrials` random error patterns of weight 0 .. t+3 through
// PkClassTable::lookup and through the algebraic decoder (pk_alg_decode); *mismatches = patterns where verdict or
// positions differ.  info[0..3] = key bits, log2 position-table slots, position-table entries, bitmap bytes.
int pk_code_class_table_check(const pk_code *c, uint64_t seed, long ntrials, long *mismatches, long *info) {
    if (!c) return fail(PK_ERR_ARG, "code is NULL");
    if (!c->ct) return fail(PK_ERR_UNSUPPORTED, "no class table for this code");
    const PkClassTable &ct = *c->ct;
    if (info) { info[0] = ct.kb; info[1] = ct.hbits; info[2] = (long)ct.entries; info[3] = (long)(ct.bits.size() * 4); }
    const int n = c->n, nw = (n + 31) / 32, per = 32 / c->m, nsw = (2 * c->t + per - 1) / per;
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    long bad = 0;
    std::vector<uint32_t> sw(nsw), A1(nw), A2(nw);
    for (long it = 0; it < ntrials; ++it) {
        const int wgt = (int)(rnd() % (uint64_t)(c->t + 4));
        std::fill(sw.begin(), sw.end(), 0u);
        for (int i = 0; i < wgt; ++i) {   // repeated positions cancel: the weight is "at most wgt"
            const int p = (int)(rnd() % (uint64_t)n);
            for (int w = 0; w < nsw; ++w) sw[w] ^= c->hcol[(size_t)p * nsw + w];
        }
        std::fill(A1.begin(), A1.end(), 0u);
        const bool ok1 = c->ks->host_alg_decode(sw.data(), c->mul.data(), c->xoff.data(), A1.data());
        const bool ok2 = ct.lookup(c->m, c->t, sw.data(), A2.data());
        if (ok1 != ok2 || (ok1 && A1 != A2)) ++bad;
    }
    if (mismatches) *mismatches = bad;
    return PK_OK;
}

int pk_make_kernel_matrix(const pk_code *c, uint8_t *out) {
    if (!c || !out) return fail(PK_ERR_ARG, "NULL argument");
    pk_code_kernel_matrix(*c, out);
    return PK_OK;
}

static int need_kernels(const pk_code *c) {
    if (!c) return fail(PK_ERR_ARG, "code is NULL");
    if (c->device < 0) return fail(PK_ERR_CUDA, "host-only code handle: libpkb200 has no CPU compute path");
    if (!c->ks)
        return fail(PK_ERR_UNSUPPORTED, "no sm_100a kernel instantiated for (m=" + std::to_string(c->m) +
                                            ", t=" + std::to_string(c->t) + "); see csrc/pk_inst_m*.cu");
    return PK_OK;
}

// ------------------------------------------------------------------ encoder / algebraic decoder
int pk_encode_batch(pk_code *c, const uint8_t *info, long B, uint8_t *cw) {
    int rc = need_kernels(c);
    if (rc) return rc;
    if (B < 0 || (B && (!info || !cw))) return fail(PK_ERR_ARG, "bad buffers");
    if (!B) return PK_OK;
    PK_CUDA(cudaSetDevice(c->device));
    uint8_t *d_in = nullptr, *d_out = nullptr;
    PK_CUDA(cudaMalloc(&d_in, (size_t)B * c->k));
    cudaError_t e = cudaMalloc(&d_out, (size_t)B * c->n);
    if (e != cudaSuccess) { cudaFree(d_in); return fail(PK_ERR_CUDA, cudaGetErrorString(e)); }
    e = cudaMemcpy(d_in, info, (size_t)B * c->k, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = c->ks->launch_encode(c->dev, d_in, B, d_out, 0);
    if (e == cudaSuccess) e = cudaMemcpy(cw, d_out, (size_t)B * c->n, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(PK_ERR_CUDA, cudaGetErrorString(e));
    return PK_OK;
}

int pk_bch_decode_batch(pk_code *c, const uint8_t *words, long B, uint8_t *answers, uint8_t *ok) {
    int rc = need_kernels(c);
    if (rc) return rc;
    if (B < 0 || (B && (!words || !answers || !ok))) return fail(PK_ERR_ARG, "bad buffers");
    if (!B) return PK_OK;
    PK_CUDA(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    PK_CUDA(cudaGetDeviceProperties(&prop, c->device));
    PkLaunchGeom g;
    PK_CUDA(c->ks->geom_bdd(prop.multiProcessorCount, &g));
    uint8_t *d_w = nullptr, *d_a = nullptr, *d_ok = nullptr;
    PK_CUDA(cudaMalloc(&d_w, (size_t)B * c->n));
    cudaError_t e = cudaMalloc(&d_a, (size_t)B * c->n);
    if (e == cudaSuccess) e = cudaMalloc(&d_ok, (size_t)B);
    if (e == cudaSuccess) e = cudaMemcpy(d_w, words, (size_t)B * c->n, cudaMemcpyHostToDevice);
    // rows of failed words keep the caller's bytes, like Decoder::decode (Decoder.cpp:300-307)
    if (e == cudaSuccess) e = cudaMemcpy(d_a, answers, (size_t)B * c->n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = c->ks->launch_bdd(g, c->dev, d_w, B, d_a, d_ok, 0);
    if (e == cudaSuccess) e = cudaMemcpy(answers, d_a, (size_t)B * c->n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ok, d_ok, (size_t)B, cudaMemcpyDeviceToHost);
    cudaFree(d_w);
    cudaFree(d_a);
    cudaFree(d_ok);
    if (e != cudaSuccess) return fail(PK_ERR_CUDA, cudaGetErrorString(e));
    return PK_OK;
}

// ------------------------------------------------------------------ Kaneko handle
int pk_kaneko_create(pk_code *c, double llr_snr_db, long J, long max_trials, pk_kaneko **out) {
    return pk_kaneko_create_ext(c, llr_snr_db, J, max_trials, 0, 0, out);
}

int pk_kaneko_create_ext(pk_code *c, double llr_snr_db, long J, long max_trials, int extended, int rules, pk_kaneko **out) {
    if (!out) return fail(PK_ERR_ARG, "out is NULL");
    if ((extended != 0 && extended != 1) || rules < 0 || rules > 2) return fail(PK_ERR_ARG, "pk_kaneko_create_ext: extended in {0,1}, rules in {0,1,2}");
    *out = nullptr;
    int rc = need_kernels(c);
    if (rc) return rc;
    if (llr_snr_db < 0) return fail(PK_ERR_ARG, "Invalid values of arguments");   // main.cpp:55
    PK_CUDA(cudaSetDevice(c->device));
    pk_kaneko *d = new (std::nothrow) pk_kaneko;
    if (!d) return fail(PK_ERR_ALLOC, "out of memory");
    d->code = c;
    {   // sd = sqrt(1 / (pow(10, snr/10) * 2 * k / n)); alpha = 2*y / pow(sd, 2)
        // (KanekoKernelProcessor.cpp:20,337) -- same expression, same operand types.
        long k = c->k, n = c->n + extended;
        double sd = sqrt(1 / (pow(10, llr_snr_db / 10) * 2 * k / n));
        d->kp.llr_den = pow(sd, 2);
    }
    d->kp.ext = extended;
    d->kp.J = (J < 0 || J >= c->n) ? -1 : (int)J;
    d->kp.max_trials = (max_trials <= 0 || max_trials > 0x7FFFFFFFL) ? 0x7FFFFFFFu : (uint32_t)max_trials;
    d->kp.frames_per_grab = 2;
    // a narrow BM+Chien step (32 trials) costs about as much latency as a bit-sliced wide step (1024 trials),
    // so those frames move to phase B almost at once; coset-table steps are cheap and stay longer
    d->kp.limit_a = c->use_lut ? 256u : c->use_ct ? 128u : 64u;
    d->mode = c->use_lut ? PK_MODE_LUT : c->use_ct ? PK_MODE_CLASS : PK_MODE_ALG;
    d->kp.big_span = 8192u;
    d->kp.huge_span = 65536u;
    {   // long searches shared by the grid: chunk = a few cooperative steps (1024 patterns per warp per step)
        const uint32_t step = 1024u * (uint32_t)(d->mode == PK_MODE_LUT ? PK_WARPS_B_LUT : d->mode == PK_MODE_CLASS ? PK_WARPS_B_CT : PK_WARPS_B);
        d->kp.mega_chunk = d->mode == PK_MODE_LUT ? 2 * step : d->mode == PK_MODE_CLASS ? 4 * step : step;
        d->kp.mega_span = 8 * d->kp.mega_chunk;
        d->kp.mega_after = d->mode == PK_MODE_LUT ? 4 * step : 8 * step;
        d->kp.solo_patterns = d->mode == PK_MODE_ALG ? 16384u : 65536u;
        d->kp.epoch = 0;
    }
    d->kp.variant = rules;
    d->kp.extra_ops = rules == 1 ? (uint32_t)(2 * (c->n + extended) + 1) : 0u;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, c->device);
    if (e == cudaSuccess) e = c->ks->geom_kaneko(d->mode, c->nk, prop.multiProcessorCount, d->geom4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->stream[0], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->stream[1], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&d->d_ctl, 3 * sizeof(PkPhaseBlock));
    if (e == cudaSuccess) e = cudaMalloc(&d->d_totals, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(&d->h_totals, 8 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        pk_kaneko_destroy(d);
        return fail(PK_ERR_CUDA, std::string("pk_kaneko_create: ") + cudaGetErrorString(e));
    }
    *out = d;
    return PK_OK;
}

void pk_kaneko_destroy(pk_kaneko *d) {
    if (!d) return;
    cudaSetDevice(d->code->device);
    for (int s = 0; s < 2; ++s) {
        if (d->stream[s]) { cudaStreamSynchronize(d->stream[s]); cudaStreamDestroy(d->stream[s]); }
        cudaFree(d->d_y[s]); cudaFree(d->d_dec[s]); cudaFree(d->d_tr[s]); cudaFree(d->d_rec[s]);
    }
    cudaFree(d->d_ctl);
    for (int i = 0; i < 3; ++i) { cudaFree(d->d_longs[i]); cudaFree(d->d_zscr[i]); cudaFree(d->d_mega_bits[i]); }
    cudaFree(d->d_totals);
    cudaFree(d->d_grec);
    if (d->h_totals) cudaFreeHost(d->h_totals);
    if (d->h_grec) cudaFreeHost(d->h_grec);
    if (d->ev_zero) cudaEventDestroy(d->ev_zero);
    delete d;
}

int pk_kaneko_set_frames_per_grab(pk_kaneko *d, int g) {
    if (!d || g < 1) return fail(PK_ERR_ARG, "bad argument");
    d->kp.frames_per_grab = g;
    return PK_OK;
}

// 0: KanekoKernelProcessor::decode(answer, word, res) -- what fun() uses (default);
// 1: KanekoKernelProcessor::decode(word, res), the file-mode flavour of main.cpp:158
int pk_kaneko_set_variant(pk_kaneko *d, int two_argument) {
    if (!d) return fail(PK_ERR_ARG, "NULL handle");
    if (d->kp.variant == 2) return fail(PK_ERR_ARG, "handle runs the exact rules (pk_kaneko_create_ext)");
    d->kp.variant = two_argument ? 1 : 0;
    d->kp.extra_ops = two_argument ? (uint32_t)(2 * frame_len(d) + 1) : 0u;
    return PK_OK;
}

int pk_kaneko_set_phase_a_limit(pk_kaneko *d, long trials) {
    if (!d || trials < 32) return fail(PK_ERR_ARG, "bad argument");
    d->kp.limit_a = (uint32_t)std::min(trials, 1L << 30) & ~31u;
    return PK_OK;
}

int pk_kaneko_launch_geometry(const pk_kaneko *d, int *grid, int *block, long *smem) {
    if (!d) return fail(PK_ERR_ARG, "NULL");
    if (grid) *grid = d->geom4[0].grid;
    if (block) *block = d->geom4[0].block;
    if (smem) *smem = (long)d->geom4[0].smem;
    return PK_OK;
}

// ------------------------------------------------------------------ launch helpers
// A batch is cut into launch pairs (phase A + phase B) of at most long_cap frames, so that every
// frame that outlives phase A finds a slot in the parked-frame list.
static int launch_pairs(pk_kaneko *d, int slot, bool gen, PkIo io, long B, cudaStream_t st) {
    const pk_code *c = d->code;
    const bool wide = d->geom4[gen ? 3 : 1].grid > 0;
    if (wide && !d->d_longs[slot]) {
        const size_t bytes = (size_t)(2 * d->long_cap + PK_HUGE_CAP) * sizeof(PkLongRec);   // parked, huge, late
        PK_CUDA(cudaMalloc(&d->d_longs[slot], bytes));
        PK_CUDA(cudaMemsetAsync(d->d_longs[slot], 0, bytes, st));   // late-list records carry the launch epoch (never 0)
        PK_CUDA(cudaMalloc(&d->d_mega_bits[slot], (size_t)PK_MEGA_SLOTS * PK_MEGA_WORDS * sizeof(uint32_t)));
    }
    if (wide && d->mode == PK_MODE_ALG && (c->t + 1) * c->m > 56 && !d->d_zscr[slot]) {
        const int gmax = std::max(d->geom4[1].grid, d->geom4[3].grid);
        PK_CUDA(cudaMalloc(&d->d_zscr[slot], (size_t)gmax * 4 * c->n * 32 * sizeof(uint32_t)));
    }
    io.zscratch = d->d_zscr[slot];
    io.mega = d->d_ctl[slot].mega;
    io.mega_bits = d->d_mega_bits[slot];
    for (long off = 0; off < B; off += d->long_cap) {
        const long nb = std::min(d->long_cap, B - off);
        PkIo part = io;
        PkKanekoParams kp = d->kp;
        if (++d->epoch == 0) d->epoch = 1;
        kp.epoch = d->epoch;
        if (gen) {
            part.gp.first_frame = io.gp.first_frame + (uint64_t)off;
            if (io.d_info) part.d_info = io.d_info + off * c->k;
            if (io.d_cw) part.d_cw = io.d_cw + off * frame_len(d);
            if (io.d_y) part.d_y = io.d_y + off * frame_len(d);
        } else {
            part.y = io.y + off * frame_len(d);
            part.decided = io.decided + off * frame_len(d);
            if (io.trials) part.trials = io.trials + off;
        }
        if (io.recs) part.recs = io.recs + off;
        PK_CUDA(c->ks->launch_kaneko(d->mode, gen, d->geom4, c->dev, kp, part, nb, &d->d_ctl[slot].ctl,
                                     d->d_longs[slot], wide ? d->long_cap : 0, st));
    }
    return PK_OK;
}
static int launch_replay(pk_kaneko *d, int slot, const double *d_y, long B, uint8_t *d_dec, uint32_t *d_tr,
                         pk_frame_rec *d_recs, unsigned long long *d_totals, cudaStream_t st) {
    PkIo io{};
    io.y = d_y; io.decided = d_dec; io.trials = d_tr; io.recs = d_recs; io.totals = d_totals;
    return launch_pairs(d, slot, false, io, B, st);
}
static int launch_gen(pk_kaneko *d, int slot, const PkGenParams &gp, long B, pk_frame_rec *d_recs,
                      unsigned long long *d_totals, uint8_t *d_info, uint8_t *d_cw, double *d_y, int dump_only,
                      cudaStream_t st) {
    PkIo io{};
    io.gp = gp; io.d_info = d_info; io.d_cw = d_cw; io.d_y = d_y; io.dump_only = dump_only;
    io.recs = d_recs; io.totals = d_totals;
    return launch_pairs(d, slot, true, io, B, st);
}

// ------------------------------------------------------------------ replay mode
int pk_kaneko_decode_batch_dev(pk_kaneko *d, const double *d_y, long B, uint8_t *d_decided, uint32_t *d_trials,
                               pk_frame_rec *d_recs, uint64_t *d_totals, void *stream) {
    if (!d || B < 0 || (B && (!d_y || !d_decided))) return fail(PK_ERR_ARG, "bad arguments");
    if (!B) return PK_OK;
    PK_CUDA(cudaSetDevice(d->code->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream[0];
    return launch_replay(d, stream ? 2 : 0, d_y, B, d_decided, d_trials, d_recs, (unsigned long long *)d_totals, st);
}

static int ensure_pipeline(pk_kaneko *d) {
    if (d->chunk) return PK_OK;
    const int n = frame_len(d);
    long chunk = (32L << 20) / (8L * n);
    chunk = std::max(1024L, chunk / 1024 * 1024);
    for (int s = 0; s < 2; ++s) {
        PK_CUDA(cudaMalloc(&d->d_y[s], (size_t)chunk * n * sizeof(double)));
        PK_CUDA(cudaMalloc(&d->d_dec[s], (size_t)chunk * n));
        PK_CUDA(cudaMalloc(&d->d_tr[s], (size_t)chunk * sizeof(uint32_t)));
        PK_CUDA(cudaMalloc(&d->d_rec[s], (size_t)chunk * sizeof(pk_frame_rec)));
    }
    d->chunk = chunk;
    return PK_OK;
}

// Chunked, double-buffered: H2D(y) -> kernels -> D2H(results) on alternating streams; nothing here waits for the device.
static int enqueue_batch(pk_kaneko *d, const double *y, long B, uint8_t *decided, uint32_t *trials, pk_frame_rec *recs) {
    int rc = ensure_pipeline(d);
    if (rc) return rc;
    const int n = frame_len(d);
    if (!d->pending) {
        // first batch since the last wait: zero the device totals ahead of both streams
        if (!d->ev_zero) PK_CUDA(cudaEventCreateWithFlags(&d->ev_zero, cudaEventDisableTiming));
        PK_CUDA(cudaMemsetAsync(d->d_totals, 0, 8 * sizeof(unsigned long long), d->stream[0]));
        PK_CUDA(cudaEventRecord(d->ev_zero, d->stream[0]));
        PK_CUDA(cudaStreamWaitEvent(d->stream[1], d->ev_zero, 0));
        d->pending = true;
    }
    for (long off = 0; off < B; off += d->chunk) {
        const int s = d->next_slot;
        d->next_slot ^= 1;
        const long nb = std::min(d->chunk, B - off);
        cudaStream_t st = d->stream[s];
        PK_CUDA(cudaMemcpyAsync(d->d_y[s], y + off * n, (size_t)nb * n * sizeof(double), cudaMemcpyHostToDevice, st));
        // undecided rows (PK_FLAG_NO_DECISION) come back zero-filled
        PK_CUDA(cudaMemsetAsync(d->d_dec[s], 0, (size_t)nb * n, st));
        rc = launch_replay(d, s, d->d_y[s], nb, d->d_dec[s], trials ? d->d_tr[s] : nullptr, recs ? d->d_rec[s] : nullptr,
                           d->d_totals, st);
        if (rc) return rc;
        PK_CUDA(cudaMemcpyAsync(decided + off * n, d->d_dec[s], (size_t)nb * n, cudaMemcpyDeviceToHost, st));
        if (trials)
            PK_CUDA(cudaMemcpyAsync(trials + off, d->d_tr[s], (size_t)nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (recs)
            PK_CUDA(cudaMemcpyAsync(recs + off, d->d_rec[s], (size_t)nb * sizeof(pk_frame_rec), cudaMemcpyDeviceToHost, st));
    }
    return PK_OK;
}

// An enqueue failed part-way (e.g. a workspace allocation): kernels and D2H copies already queued still write into
// the caller's buffers, so nothing may return before both streams are idle.  The error message survives.
static void drain_after_error(pk_kaneko *d) {
    const std::string keep = g_err;
    cudaStreamSynchronize(d->stream[0]);
    cudaStreamSynchronize(d->stream[1]);
    d->pending = false;
    g_err = keep;
}
// The generation-mode entry points zero / read the handle's totals on stream[0]: not while asynchronous batches
// (pk_kaneko_decode_batch_async) are outstanding.
static int reject_pending(const pk_kaneko *d) {
    if (d->pending) return fail(PK_ERR_ARG, "asynchronous batches are pending on this handle: call pk_kaneko_wait first");
    return PK_OK;
}

int pk_kaneko_wait(pk_kaneko *d, pk_point_result *totals) {
    if (!d) return fail(PK_ERR_ARG, "NULL handle");
    if (totals) std::memset(totals, 0, sizeof(*totals));
    PK_CUDA(cudaSetDevice(d->code->device));
    PK_CUDA(cudaStreamSynchronize(d->stream[0]));
    PK_CUDA(cudaStreamSynchronize(d->stream[1]));
    if (d->pending && totals) {
        PK_CUDA(cudaMemcpy(d->h_totals, d->d_totals, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        std::memcpy(totals, d->h_totals, sizeof(*totals));
    }
    d->pending = false;
    return PK_OK;
}

int pk_kaneko_decode_batch_async(pk_kaneko *d, const double *y, long B, uint8_t *decided, uint32_t *trials,
                                 pk_frame_rec *recs) {
    if (!d || B < 0 || (B && (!y || !decided))) return fail(PK_ERR_ARG, "bad arguments");
    if (!B) return PK_OK;
    PK_CUDA(cudaSetDevice(d->code->device));
    int rc = enqueue_batch(d, y, B, decided, trials, recs);
    if (rc) drain_after_error(d);
    return rc;
}

int pk_kaneko_decode_batch(pk_kaneko *d, const double *y, long B, uint8_t *decided, uint32_t *trials,
                           pk_frame_rec *recs, pk_point_result *totals) {
    if (!d || B < 0 || (B && (!y || !decided))) return fail(PK_ERR_ARG, "bad arguments");
    if (totals) std::memset(totals, 0, sizeof(*totals));
    if (d->pending) return fail(PK_ERR_ARG, "asynchronous batches are pending on this handle: call pk_kaneko_wait first");
    if (!B) return PK_OK;
    PK_CUDA(cudaSetDevice(d->code->device));
    int rc = enqueue_batch(d, y, B, decided, trials, recs);
    if (rc) { drain_after_error(d); return rc; }
    return pk_kaneko_wait(d, totals);
}

// ------------------------------------------------------------------ generation mode
static double channel_sigma(const pk_kaneko *d, double ebn0_db) {
    long k = d->code->k, n = frame_len(d);
    return sqrt(1 / (pow(10, ebn0_db / 10) * 2 * k / n));   // dataForPlot.cpp:45
}

int pk_kaneko_run_frames_dev(pk_kaneko *d, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame,
                             long nframes, pk_frame_rec *d_recs, uint64_t *d_totals, void *stream) {
    if (!d || nframes < 0) return fail(PK_ERR_ARG, "bad arguments");
    if (!nframes) return PK_OK;
    PK_CUDA(cudaSetDevice(d->code->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream[0];
    PkGenParams gp;
    gp.sigma = channel_sigma(d, ebn0_db);
    gp.seed = seed;
    gp.first_frame = first_frame;
    gp.snr_index = (uint32_t)snr_index;
    return launch_gen(d, stream ? 2 : 0, gp, nframes, d_recs, (unsigned long long *)d_totals, nullptr, nullptr, nullptr, 0, st);
}

static int ensure_grec(pk_kaneko *d, long cap) {
    if (cap <= d->grec_cap) return PK_OK;
    cudaFree(d->d_grec);
    if (d->h_grec) cudaFreeHost(d->h_grec);
    d->d_grec = nullptr; d->h_grec = nullptr; d->grec_cap = 0;
    PK_CUDA(cudaMalloc(&d->d_grec, (size_t)cap * sizeof(pk_frame_rec)));
    PK_CUDA(cudaMallocHost(&d->h_grec, (size_t)cap * sizeof(pk_frame_rec)));
    d->grec_cap = cap;
    return PK_OK;
}

int pk_kaneko_run_frames(pk_kaneko *d, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame,
                         long nframes, pk_frame_rec *recs, pk_point_result *totals) {
    if (!d || nframes < 0 || !totals) return fail(PK_ERR_ARG, "bad arguments");
    if (int rp = reject_pending(d)) return rp;
    if (!nframes) return PK_OK;
    PK_CUDA(cudaSetDevice(d->code->device));
    cudaStream_t st = d->stream[0];
    const long step = recs ? (1L << 20) : nframes;
    if (recs) {
        int rc = ensure_grec(d, std::min(step, nframes));
        if (rc) return rc;
    }
    PK_CUDA(cudaMemsetAsync(d->d_totals, 0, 8 * sizeof(unsigned long long), st));
    for (long off = 0; off < nframes; off += step) {
        const long nb = std::min(step, nframes - off);
        int rc = pk_kaneko_run_frames_dev(d, ebn0_db, snr_index, seed, first_frame + (uint64_t)off, nb,
                                          recs ? d->d_grec : nullptr, (uint64_t *)d->d_totals, st);
        if (rc) return rc;
        if (recs) {
            PK_CUDA(cudaMemcpyAsync(d->h_grec, d->d_grec, (size_t)nb * sizeof(pk_frame_rec), cudaMemcpyDeviceToHost, st));
            PK_CUDA(cudaStreamSynchronize(st));
            std::memcpy(recs + off, d->h_grec, (size_t)nb * sizeof(pk_frame_rec));
        }
    }
    PK_CUDA(cudaMemcpyAsync(d->h_totals, d->d_totals, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    PK_CUDA(cudaStreamSynchronize(st));
    uint64_t *t = reinterpret_cast<uint64_t *>(totals);
    for (int i = 0; i < 6; ++i) t[i] += d->h_totals[i];
    t[6] = std::max<uint64_t>(t[6], d->h_totals[6]);
    t[7] |= d->h_totals[7];
    return PK_OK;
}

int pk_generate_frames_dev(pk_kaneko *d, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame,
                           long nframes, uint8_t *d_info, uint8_t *d_cw, double *d_y, void *stream) {
    if (!d || nframes < 0) return fail(PK_ERR_ARG, "bad arguments");
    if (!nframes) return PK_OK;
    const pk_code *c = d->code;
    PK_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream[0];
    PkGenParams gp;
    gp.sigma = channel_sigma(d, ebn0_db);
    gp.seed = seed;
    gp.first_frame = first_frame;
    gp.snr_index = (uint32_t)snr_index;
    return launch_gen(d, stream ? 2 : 0, gp, nframes, nullptr, nullptr, d_info, d_cw, d_y, 1, st);
}

int pk_generate_frames(pk_kaneko *d, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame, long nframes,
                       uint8_t *info, uint8_t *cw, double *y) {
    if (!d || nframes < 0) return fail(PK_ERR_ARG, "bad arguments");
    if (int rp = reject_pending(d)) return rp;
    if (!nframes) return PK_OK;
    const pk_code *c = d->code;
    PK_CUDA(cudaSetDevice(c->device));
    uint8_t *d_info = nullptr, *d_cw = nullptr;
    double *d_y = nullptr;
    cudaError_t e = cudaSuccess;
    if (info) e = cudaMalloc(&d_info, (size_t)nframes * c->k);
    const size_t fl = (size_t)frame_len(d);
    if (e == cudaSuccess && cw) e = cudaMalloc(&d_cw, (size_t)nframes * fl);
    if (e == cudaSuccess && y) e = cudaMalloc(&d_y, (size_t)nframes * fl * sizeof(double));
    if (e == cudaSuccess) {
        PkGenParams gp;
        gp.sigma = channel_sigma(d, ebn0_db);
        gp.seed = seed;
        gp.first_frame = first_frame;
        gp.snr_index = (uint32_t)snr_index;
        if (launch_gen(d, 0, gp, nframes, nullptr, nullptr, d_info, d_cw, d_y, 1, d->stream[0]) != PK_OK) e = cudaErrorUnknown;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream[0]);
    if (e == cudaSuccess && info) e = cudaMemcpy(info, d_info, (size_t)nframes * c->k, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && cw) e = cudaMemcpy(cw, d_cw, (size_t)nframes * fl, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && y) e = cudaMemcpy(y, d_y, (size_t)nframes * fl * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_info);
    cudaFree(d_cw);
    cudaFree(d_y);
    if (e != cudaSuccess) return fail(PK_ERR_CUDA, cudaGetErrorString(e));
    return PK_OK;
}

// One SNR point with fun()'s stop rule (dataForPlot.cpp:43): frames are decoded in growing
// chunks; the per-frame records are then scanned IN FRAME ORDER and the scan stops at the
// frame where `count < p && countErr < e` first fails, so the totals equal a sequential run.
int pk_kaneko_run_point(pk_kaneko *d, double ebn0_db, int snr_index, uint64_t seed, long p, long e,
                        pk_point_result *out) {
    if (!d || !out || p <= 0) return fail(PK_ERR_ARG, "Invalid values of arguments");
    if (int rp = reject_pending(d)) return rp;
    std::memset(out, 0, sizeof(*out));
    if (e <= 0) return pk_kaneko_run_frames(d, ebn0_db, snr_index, seed, 0, p, nullptr, out);
    const int n = frame_len(d);
    long done = 0, chunk = 4096;
    std::vector<pk_frame_rec> recs;
    while (done < p && (long)out->frame_errors < e) {
        const long nb = std::min(chunk, p - done);
        recs.resize((size_t)nb);
        pk_point_result scratch;
        std::memset(&scratch, 0, sizeof scratch);
        int rc = pk_kaneko_run_frames(d, ebn0_db, snr_index, seed, (uint64_t)done, nb, recs.data(), &scratch);
        if (rc) return rc;
        for (long f = 0; f < nb && (long)out->frame_errors < e; ++f) {
            const pk_frame_rec &r = recs[(size_t)f];
            const uint64_t run = (uint64_t)r.trials - ((r.flags & PK_FLAG_EARLY_RETURN) ? 1 : 0);
            out->frames += 1;
            out->frame_errors += (r.flags & PK_FLAG_FRAME_ERROR) ? 1 : 0;
            out->bit_errors += r.bit_errors;
            out->trials += r.trials;
            out->cmp += run * (uint64_t)(n + 6) + r.extra_cmp;
            out->sum += run * (uint64_t)(n + 1) + r.extra_sum;
            out->max_trials_seen = std::max<uint64_t>(out->max_trials_seen, r.trials);
            out->flags_or |= r.flags;
        }
        done += nb;
        chunk = std::min(chunk * 2, 1L << 20);
    }
    return PK_OK;
}

}  // extern "C"
