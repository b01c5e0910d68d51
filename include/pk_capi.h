/*
 * pk_capi.h -- C ABI of libpkb200.so, the B200-native (sm_100a) replacement for the
 * Monte-Carlo hot path of lizmoscow/polar-codes-with-bch-kernel.
 *
 * Every entry point names the reference interface it replaces (file:line under the
 * reference tree).  Plain pointers and sizes only; no exceptions cross the boundary:
 * every function returns PK_OK (0) or a negative pk_status and leaves a message in
 * pk_last_error().  There is NO CPU fallback: without a CUDA device every compute
 * call fails with PK_ERR_CUDA.
 *
 * Conventions (identical to the reference):
 *   - bits are one byte each (0/1), polynomial index == array index == power of x
 *     (headers/bchCoder.h:10-48);
 *   - a frame is n bytes (codeword / hard decisions) or n doubles (channel output y);
 *   - batches are row-major [B][n], caller-owned.
 *   Pointers named d_* must be device pointers on the handle's device, all others host.
 */
#ifndef PK_CAPI_H
#define PK_CAPI_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PK_OK = 0,
    PK_ERR_ARG = -1,         /* the reference throws "Invalid values of arguments" (src/main.cpp:55-57) */
    PK_ERR_UNSUPPORTED = -2, /* no sm_100a kernel instantiated for this (m,t) */
    PK_ERR_CUDA = -3,        /* CUDA runtime error / no device */
    PK_ERR_ALLOC = -4
} pk_status;

/* per-frame flags in pk_frame_rec.flags */
#define PK_FLAG_EARLY_RETURN 0x01 /* left through `if (l < calcRightSide()) return;` (KanekoKernelProcessor.cpp:380) */
#define PK_FLAG_NO_DECISION 0x02  /* no trial ever succeeded: the reference leaves `res` untouched */
#define PK_FLAG_SORT_TIE 0x04     /* informational: equal |alpha| keys met; their order was resolved by replaying libstdc++'s std::sort */
#define PK_FLAG_FRAME_ERROR 0x08  /* generation mode: decided != transmitted (dataForPlot.cpp:66) */
#define PK_FLAG_TRUNCATED 0x10    /* stopped by the max_trials safety cap (never set with the default cap) */
#define PK_FLAG_NON_ML 0x40       /* generation mode: the transmitted word is more likely than the decision -- what the
                                      reference's DEBUG build writes to out/errWords.txt (dataForPlot.cpp:55-64) */
#define PK_FLAG_REF_UNDEFINED 0x20 /* 2-argument flavour only: the reference's unbounded `while (l >= calcT(j))` (:257) ran
                                      past the reliability array here, i.e. its own result is undefined */

/* One record per decoded frame (16 bytes). The reference's three operation counters
 * (KanekoKernelProcessor.cpp:367-369,386-404) are reconstructed as
 *   decodingCount   = trials
 *   comparisonCount = (trials - early) * (n + 6) + extra_cmp
 *   summCount       = (trials - early) * (n + 1) + extra_sum        early = flags & 1 */
typedef struct {
    uint32_t trials;
    uint32_t extra_cmp;
    uint32_t extra_sum;
    uint16_t bit_errors; /* generation mode only (dataForPlot.cpp:69-73) */
    uint8_t flags;
    uint8_t reserved;
} pk_frame_rec;

/* Totals of one SNR point / one batch: what fun() accumulates (dataForPlot.cpp:20,66-87). */
typedef struct {
    uint64_t frames;
    uint64_t frame_errors;
    uint64_t bit_errors;
    uint64_t trials; /* decodingCount   */
    uint64_t cmp;    /* comparisonCount */
    uint64_t sum;    /* summCount       */
    uint64_t max_trials_seen;
    uint64_t flags_or;
} pk_point_result;

typedef struct pk_code pk_code;     /* field tables + g(x) + device tables  */
typedef struct pk_kaneko pk_kaneko; /* decoder instance: stream + workspaces */

const char *pk_last_error(void);
int pk_device_count(void);

/* ---- code construction: replaces src/main.cpp:59-95 (tables, g(x) = lcm of minimal
 * polynomials via findMinimalPolynomial/lcm, src/bchCoder.cpp:25,217).  m in [3,8],
 * 0 < t < 2^(m-1).  `device` is the CUDA ordinal the tables are uploaded to. */
int pk_code_create(int m, int t, int device, pk_code **out);
/* Host tables only, no CUDA (introspection / CPU tests); compute calls on it return PK_ERR_CUDA. */
int pk_code_create_host(int m, int t, pk_code **out);
void pk_code_destroy(pk_code *code);
/* (n,k,d) banner values of main.cpp:93-97 and g(x) (gsize = n-k+1 bytes; g_out may be NULL) */
int pk_code_info(const pk_code *code, int *n, int *k, int *d, int *gsize, uint8_t *g_out);
/* antilogarithms[n], logarithms[n+1] exactly as main.cpp:63-78 builds them (log[0] = LONG_MAX) */
int pk_code_tables(const pk_code *code, uint64_t *antilog_out, uint64_t *log_out);
/* Which lookup table replaces the algebraic decoder (Decoder::decode, src/Decoder.cpp:298-321) in the searches:
 * 1 = shared-memory coset table (n-k <= 16, t*m <= 15), 2 = cyclic-class table (bitmap + position table over the
 * syndrome classes under cyclic shifts; (31,11,11), (31,6,15), (63,45,7) .. (63,30,13)), 0 = none.
 * pk_code_set_lut(code, 0) forces the Berlekamp-Massey + Chien kernels instead (call before pk_kaneko_create);
 * pk_code_set_lut(code, 1) goes back to the code's table. */
int pk_code_uses_lut(const pk_code *code);
int pk_code_set_lut(pk_code *code, int enable);
/* Host-side differential self-check of the cyclic-class table against the algebraic decoder on `ntrials` random
 * error patterns of weight 0 .. t+3; info[4] = key bits, log2 slots, entries, bitmap bytes (may be NULL). */
int pk_code_class_table_check(const pk_code *code, uint64_t seed, long ntrials, long *mismatches, long *info);
/* The coset table itself: entry r = up to t error positions (m bits each, n = none) of the
 * algebraic decoder's answer for syndrome r(x) = word mod g, 0xFFFF = decoding failure. */
int pk_code_coset_table(const pk_code *code, uint16_t *out /*[2^(n-k)] or NULL*/, long *nentries);

/* ---- encoder: multiplyPolynomials(info,k,g,gSize,res) (src/bchCoder.cpp:120-132) */
int pk_encode_batch(pk_code *code, const uint8_t *info /*[B][k]*/, long B, uint8_t *cw /*[B][n]*/);

/* ---- algebraic decoder: Decoder::findSyndromPoly + Decoder::decode
 * (src/Decoder.cpp:184-207,298-321).  answers[f] is written only when ok[f] == 1. */
int pk_bch_decode_batch(pk_code *code, const uint8_t *words /*[B][n]*/, long B,
                        uint8_t *answers /*[B][n]*/, uint8_t *ok /*[B]*/);

/* ---- Kaneko decoder: KanekoKernelProcessor(pw,n,t,k,antilog,log,snr) ctor
 * (src/KanekoKernelProcessor.cpp:17-26); llr_snr_db is the ctor's signalToNoiseRatio
 * (main.cpp:176 passes 0.5).  J < 0: HEAD semantics T = j (line 393); J >= 0: capped
 * T = min(j,J) (line 392, the *_e*.csv runs).  max_trials <= 0: the reference bound. */
int pk_kaneko_create(pk_code *code, double llr_snr_db, long J, long max_trials, pk_kaneko **out);
/* Two things the reference does not have (it only builds n = 2^m - 1, src/main.cpp:60, and only its own stopping rules):
 *   extended = 1: the EXTENDED code (n+1, k, 2t+2) -- position n carries the overall parity of the BCH codeword; frames are
 *     n+1 long everywhere (y, decisions, generation mode), Eb/N0 uses the rate k/(n+1).  Test patterns flip the least
 *     reliable BCH positions, the algebraic decoder works on the BCH part, the decided parity position follows from it, and
 *     m, l, calcRightSide (with d = 2t+2) and calcT run over all n+1 positions.  eBCH(128,64,22) = (m,t) = (7,10), extended.
 *     DESIGN.md states the definition; it is pinned against exhaustive ML on the small extended codes, not against the reference.
 *   rules: 0 = decode(answer, word, res) (:335-407), 1 = decode(word, res) (:212-276), 2 = EXACT rules: loop bound 1 << T,
 *     optimality test l < sum of the d - m least reliable agreeing positions, T = first j with l < sum_{i=j..j+t} alpha_(i):
 *     both are sufficient conditions, so the decision is the ML codeword (what the kernel-LLR bridge needs). */
int pk_kaneko_create_ext(pk_code *code, double llr_snr_db, long J, long max_trials, int extended, int rules, pk_kaneko **out);
void pk_kaneko_destroy(pk_kaneko *dec);
/* Which of the reference's decode flavours the handle runs: 0 (default) decode(answer, word, res), the one fun()
 * uses (KanekoKernelProcessor.cpp:335-407); 1 decode(word, res), the file-mode flavour of main.cpp:158 (:212-276:
 * bound 1 << T without "- 1", T starts unbounded, no cap J, 2n+1 added to both synthetic counters). */
int pk_kaneko_set_variant(pk_kaneko *dec, int two_argument);
/* tuning / introspection */
int pk_kaneko_set_frames_per_grab(pk_kaneko *dec, int frames);
/* trials (multiple of 32) a frame may spend in the narrow phase-A search before it is handed to the
 * wide phase-B search (1024 patterns per warp step); tuning only, results do not depend on it */
int pk_kaneko_set_phase_a_limit(pk_kaneko *dec, long trials);
int pk_kaneko_launch_geometry(const pk_kaneko *dec, int *grid, int *block, long *smem_bytes);

/* Replay mode, host buffers: decode(answer, word, res) for B frames
 * (src/KanekoKernelProcessor.cpp:335-407).  H2D of y and D2H of the results happen
 * inside the call (chunked, double-buffered).  decided rows of frames flagged
 * PK_FLAG_NO_DECISION come back zero-filled (the reference leaves the caller's stale bytes;
 * it takes 2^15-1 .. 2^31-1 consecutive failed trials to get there).  recs / totals may be NULL. */
int pk_kaneko_decode_batch(pk_kaneko *dec, const double *y /*[B][n]*/, long B, uint8_t *decided /*[B][n]*/,
                           uint32_t *trials /*[B] or NULL*/, pk_frame_rec *recs /*[B] or NULL*/,
                           pk_point_result *totals /*or NULL*/);

/* The same without waiting: the batch is enqueued on the handle's two streams and the call returns.  Several batches
 * may be enqueued back to back (their copies overlap the kernels of the batches before them); pk_kaneko_wait blocks
 * until all of them are finished and returns the totals accumulated over them.  The host buffers must stay valid
 * until then, and must be page-locked (cudaHostAlloc / cudaHostRegister) for the copies to overlap. */
int pk_kaneko_decode_batch_async(pk_kaneko *dec, const double *y /*[B][n]*/, long B, uint8_t *decided /*[B][n]*/,
                                 uint32_t *trials /*[B] or NULL*/, pk_frame_rec *recs /*[B] or NULL*/);
int pk_kaneko_wait(pk_kaneko *dec, pk_point_result *totals /*or NULL*/);

/* Replay mode, device-resident buffers, asynchronous on `stream` (a cudaStream_t, NULL =
 * the handle's own NON-BLOCKING stream: work the caller queued elsewhere -- e.g. the fill that
 * zeroes d_totals -- is not ordered before it, synchronise first or pass your stream).
 * d_totals (8 x u64, pk_point_result layout) is ACCUMULATED into, the caller zeroes it.
 * ANY non-NULL `stream` -- including one of the handle's own -- uses the one control block / parked-frame list
 * reserved for caller streams: at most one launch in flight per handle across all caller streams. */
int pk_kaneko_decode_batch_dev(pk_kaneko *dec, const double *d_y, long B, uint8_t *d_decided,
                               uint32_t *d_trials /*or NULL*/, pk_frame_rec *d_recs /*or NULL*/,
                               uint64_t *d_totals /*or NULL*/, void *stream);

/* Generation mode: the body of fun()'s while loop (src/dataForPlot.cpp:43-74) for frames
 * [first_frame, first_frame + nframes) of SNR point `snr_index` at ebn0_db, fully on the
 * device: Philox4x32-10 info bits + AWGN (counter = (frame, snr_index, draw), key = seed)
 * -> encode -> Kaneko decode -> compare.  Results do not depend on how the frame range
 * is split over calls or GPUs.  d_recs (nframes records) may be NULL.  Asynchronous on
 * `stream`; d_totals is accumulated into. */
int pk_kaneko_run_frames_dev(pk_kaneko *dec, double ebn0_db, int snr_index, uint64_t seed,
                             uint64_t first_frame, long nframes, pk_frame_rec *d_recs /*or NULL*/,
                             uint64_t *d_totals, void *stream);

/* Same, synchronous, host-side result (adds into *totals). recs (host, nframes) may be NULL. */
int pk_kaneko_run_frames(pk_kaneko *dec, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame,
                         long nframes, pk_frame_rec *recs /*or NULL*/, pk_point_result *totals);

/* The frames generation mode would draw, written out (host buffers; any may be NULL):
 * generateRandomPoly + multiplyPolynomials + addNoise (dataForPlot.cpp:45-48) on the Philox stream. */
int pk_generate_frames(pk_kaneko *dec, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame,
                       long nframes, uint8_t *info /*[B][k]*/, uint8_t *cw /*[B][n]*/, double *y /*[B][n]*/);

/* Same, into device buffers (any may be NULL), asynchronous on `stream`. */
int pk_generate_frames_dev(pk_kaneko *dec, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame,
                           long nframes, uint8_t *d_info, uint8_t *d_cw, double *d_y, void *stream);

/* One whole SNR point on ONE device with fun()'s stop rule `count < p && countErr < e` (dataForPlot.cpp:43): frames
 * 0, 1, 2, .. of the Philox stream are decoded in growing chunks and the per-frame records are scanned in frame order,
 * so the totals equal a sequential run over the same frames (e <= 0: exactly p frames, no records needed).  Sharding a
 * point over several GPUs is pk_comm_run_point (below) or, per process, pk_kaneko_run_frames on a frame sub-range. */
int pk_kaneko_run_point(pk_kaneko *dec, double ebn0_db, int snr_index, uint64_t seed, long p, long e,
                        pk_point_result *out);

/* ======================================================================================================
 * Multi-GPU: the frames of an SNR point are independent, so they are sharded over the GPUs of the box
 * by global frame index and the per-point counters are combined by ONE NCCL all-reduce over NVLink.  The reference
 * has no counterpart (src/dataForPlot.cpp:41-95 is one sequential loop); results do not depend on the GPU count.
 * NCCL is bound at run time (dlopen libnccl.so.2); a communicator of one rank needs none.
 * ====================================================================================================== */
typedef struct pk_comm pk_comm;
typedef struct pk_comm_kaneko pk_comm_kaneko;

/* One process driving ndev devices (devices = NULL: ordinals 0 .. ndev-1): ncclCommInitAll, one stream per device. */
int pk_comm_create(int ndev, const int *devices, pk_comm **out);
/* One process per GPU (torchrun, mpirun): rank 0 draws an id (pk_comm_unique_id), the caller hands it to the other
 * processes, every process calls pk_comm_create_rank(world, rank, id, device).  world == 1 ignores id. */
int pk_comm_unique_id(uint8_t *id128 /*[128]*/);
int pk_comm_create_rank(int world, int rank, const uint8_t *id128, int device, pk_comm **out);
void pk_comm_destroy(pk_comm *comm);
int pk_comm_size(const pk_comm *comm);            /* ranks of the whole job */
int pk_comm_rank(const pk_comm *comm);            /* global rank of this process' first device */
int pk_comm_local_devices(const pk_comm *comm);   /* devices this process drives */
void *pk_comm_stream(const pk_comm *comm, int local_device);   /* cudaStream_t the collectives of that device run on */
/* In-place reduction over all ranks of the pk_point_result each local device holds at d_results[local] (device
 * memory): sum of frames / frame_errors / bit_errors / trials / cmp / sum, maximum of max_trials_seen, OR of flags_or.
 * Enqueued on the communicator's streams (asynchronous); pk_comm_sync waits for them. */
int pk_allreduce_point(pk_comm *comm, pk_point_result *const *d_results /*[local devices]*/);
int pk_comm_sync(pk_comm *comm);
/* A Kaneko decoder on every local device of the communicator (pk_code_create + pk_kaneko_create per device). */
int pk_comm_kaneko_create(pk_comm *comm, int m, int t, double llr_snr_db, long J, long max_trials, pk_comm_kaneko **out);
void pk_comm_kaneko_destroy(pk_comm_kaneko *dec);
pk_kaneko *pk_comm_kaneko_local(pk_comm_kaneko *dec, int local_device);
/* fun()'s SNR point (dataForPlot.cpp:43-95) sharded over the communicator.  e <= 0: exactly p frames, contiguous share
 * per rank, ONE all-reduce.  e > 0: the stop rule `count < p && countErr < e` in global frame order (rounds of
 * world x chunk frames; per round the error counts of the chunks are exchanged, the chunk holding the e-th error is cut
 * right after it): *out equals pk_kaneko_run_point on one device, on every rank. */
int pk_comm_run_point(pk_comm_kaneko *dec, double ebn0_db, int snr_index, uint64_t seed, long p, long e,
                      pk_point_result *out);

/* n x n nested-BCH polarisation kernel, makeMatrix (src/bchCoder.cpp:317-345); row-major bytes. */
int pk_make_kernel_matrix(const pk_code *code, uint8_t *out /*[n][n]*/);

/* ======================================================================================================
 * Polar codes with binary matrix kernels (the reference's vendored library, headers/external + out/external):
 * mixed-kernel encoder, trellis kernel processor and SC / SC-list decoder.  Bits are 0/1 bytes, LLRs are
 * fp32 with LLR > 0 <=> bit 0 (Modem.h:64,78; MixedKernelListDecoder.cpp:80).
 * ====================================================================================================== */
typedef struct pk_polar pk_polar;

/* CMixedKernelListDecoder(std::istream& Spec, unsigned ListSize) (MixedKernelListDecoder.cpp:10,
 * CMixedKernelEncoder::CMixedKernelEncoder MixedKernelEncoder.cpp:7-97): spec_text is the reference's code
 * specification ("N K dmin layers nShort nPunct", kernel names, shortened / punctured indices, N0-K freezing
 * constraints "w i_1 .. i_w").  Kernels are matrix kernels loaded from a file ("-path" or "<path",
 * Kernel.cpp:93-107,255-262).  list_size 1 = plain SC; <= 32.  device < 0: host tables only. */
int pk_polar_create(const char *spec_text, int list_size, int device, pk_polar **out);
void pk_polar_destroy(pk_polar *p);
int pk_polar_info(const pk_polar *p, int *N, int *K, int *N0, int *layers, int *list_size);
/* m_ppNumOfActiveBits of the kernel of `layer` (TrellisKernelProcessor.cpp:105,154): out[l][l+1] */
int pk_polar_trellis_profile(const pk_polar *p, int layer, int *size, uint8_t *out);
/* Host check of the two trellis table forms of a layer's kernel against each other (the in-place numbering of the lanes
 * decoder vs the gather form of the warp-per-path decoder) on random integer costs; works on host-only handles. */
int pk_polar_trellis_selfcheck(const pk_polar *p, int layer, uint64_t seed, int ntests, int *state_bits);
/* (2^m) x (2^m) extended-BCH polarisation kernel, makeMatrix of the root bchCoder.cpp:356-389; m in [3,6] */
int pk_make_ebch_kernel(int m, uint8_t *out /*[2^m][2^m]*/);
/* CBinaryEncoder::Encode (Codec.h:52; MixedKernelEncoder.cpp:142-176) */
int pk_polar_encode_batch(pk_polar *p, const uint8_t *info /*[B][K]*/, long B, uint8_t *cw /*[B][N]*/);
/* CKernProcLLR::GetLLRs (KernProc.h:40; TrellisKernelProcessor.cpp:234-295), stride 1, for B independent
 * kernel blocks: out[b][ph] = LLR of kernel input ph given inputs u[b][0..ph) and output LLRs chan[b][:] */
int pk_polar_kernel_llrs(pk_polar *p, int layer, const float *chan /*[B][l]*/, const uint8_t *u /*[B][l]*/, long B,
                         float *out /*[B][l]*/);
/* CBinarySoftDecoder::Decode (Codec.h:100-118; MixedKernelListDecoder.cpp:211-269): count[b] list entries,
 * best first; inf [B][L][K]; cw [B][L][N] and metric [B][L] may be NULL */
int pk_polar_decode_batch(pk_polar *p, const float *llr /*[B][N]*/, long B, int *count, uint8_t *inf, uint8_t *cw,
                          float *metric);
/* The same with device buffers, asynchronous on `stream` (NULL = the handle's own stream).  A handle owns one set of
 * decoder scratch (transposed channel LLRs): launches of ONE handle must be ordered on one stream; use one handle per
 * stream for concurrent decoding. */
int pk_polar_decode_batch_dev(pk_polar *p, const float *d_llr, long B, int *d_count, uint8_t *d_inf, uint8_t *d_cw,
                              float *d_metric, void *stream);

/* Generation mode of the polar path -- the body of the reference's simulator loop (out/external/Simulator.cpp:139-335,
 * not buildable: GSL / Windows) on the device: Philox4x32-10 information bits (counter = (frame, snr_index, draw), key =
 * seed) -> Encode -> BPSK (bit 0 -> +1, Modem.h:64) + AWGN -> LLR = 2y/sigma^2 as fp32 (Modem.h:78) -> Decode -> compare
 * the best path's information vector with the transmitted one.  sigma^2 = 1 / (2 (K/N) 10^(EbN0/10)) (Simulator.cpp:104).
 * d_totals (8 x u64, pk_point_result layout: frames, frame_errors, bit_errors = information-bit errors) is accumulated
 * into; results do not depend on how the frame range is split over calls or GPUs. */
int pk_polar_run_frames_dev(pk_polar *p, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame, long nframes,
                            uint64_t *d_totals, void *stream);
int pk_polar_run_frames(pk_polar *p, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame, long nframes,
                        pk_point_result *totals);
/* the frames themselves: info [B][K], cw [B][N] (either may be NULL), llr [B][N]; device / host buffers */
int pk_polar_generate_frames_dev(pk_polar *p, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame, long nframes,
                                 uint8_t *d_info, uint8_t *d_cw, float *d_llr, void *stream);
int pk_polar_generate_frames(pk_polar *p, double ebn0_db, int snr_index, uint64_t seed, uint64_t first_frame, long nframes,
                             uint8_t *info, uint8_t *cw, float *llr);

/* ======================================================================================================
 * Kaneko decoding as a polarisation-kernel processor.  CKernProcLLR::GetLLRs
 * (headers/external/KernProc.h:40-60) for the (2^m) x (2^m) extended-BCH kernel of the root bchCoder.cpp:356-389:
 * LLR of kernel input `phase` = M[1] - M[0], M[v] = metric of the best codeword of the coset with u_phase = v -- the value
 * the trellis processor computes as pStateMetric0[1] - pStateMetric0[0] (out/external/TrellisKernelProcessor.cpp:292),
 * here by maximum-likelihood Kaneko decodings of extended BCH codes (exact rules of pk_kaneko_create_ext), which also
 * serves the 64 x 64 kernel the trellis processor rejects (:71-72).  m in [3,6]; row tails of dimension <= enum_dim
 * (<= 12, negative = 12) are enumerated instead of decoded; max_trials <= 0: 2^22 per search.
 * ====================================================================================================== */
typedef struct pk_kproc pk_kproc;
int pk_kproc_create(int m, int device, long max_trials, int enum_dim, pk_kproc **out);
void pk_kproc_destroy(pk_kproc *h);
/* size = 2^m; per phase: mode (0 enumeration, 1 even-weight closed form, 2 Kaneko search), t of the BCH code decoded,
 * number of kernel rows enumerated on top of it (2 * 2^rows searches per LLR).  Arrays of `size` ints, any may be NULL. */
int pk_kproc_info(const pk_kproc *h, int *size, int *mode, int *t, int *nextra);
/* GetLLRs(Stride, phase, pKnownInputSymbols [l][Stride], pChannelLLRs [l][Stride], pLLRs [Stride]); host buffers.
 * *truncated (may be NULL): searches stopped by the trial budget (their minimum is over the codewords found so far). */
int pk_kproc_get_llrs(pk_kproc *h, int stride, int phase, const uint8_t *known, const float *chan, float *out, long *truncated);
/* all phases of B independent kernel blocks, layout of pk_polar_kernel_llrs: chan [B][l], u [B][l] -> out [B][l] */
int pk_kproc_kernel_llrs(pk_kproc *h, const float *chan, const uint8_t *u, long B, float *out, long *truncated);

/* ---- kernel construction tooling of the reference's newer bchCoder.cpp
 * Work of the trellis kernel processor (TrellisKernelProcessor.cpp:234-295) on a binary size x size kernel: branch
 * evaluations of one GetLLRs pass over all phases (11 712 for the 16 x 16 extended-BCH kernel) and the largest number of
 * state bits.  Stands in for the Sum/Cmp counters of the `SectionedTrellisKernelProcessor` the reference's own scoring
 * calls but does not ship (root bchCoder.cpp:13,497-511). */
int pk_kernel_trellis_cost(int size, const uint8_t *matrix, uint64_t *branches, int *max_state_bits);
/* swapColumns (root bchCoder.cpp:478-496): columns 0..2 stay, column i >= 3 <- column j+1 with fieldElements[j] == i */
int pk_kernel_swap_columns(int power, const uint64_t *field_elements /*[2^power]*/, const uint8_t *matrix, uint8_t *out);
/* candidate `trial` of the random search: B = L U (randomInvertibleMatrix :766-784) drawn from Philox(seed, trial),
 * column j <- column B j (:604-622); basis_out [power] = newBasis */
int pk_kernel_permute_columns(int power, const uint8_t *matrix, uint64_t seed, uint64_t trial, uint8_t *out, uint32_t *basis_out);
/* randomSwapColumns (:541-699) on the GPU, one thread per candidate: ntrials <= 2^26 (the reference: 2*10^7) */
int pk_kernel_random_search(int power, const uint8_t *matrix, long ntrials, uint64_t seed, int device, int max_state_bits,
                            uint8_t *best_matrix, uint32_t *best_basis, uint64_t *best_cost, uint64_t *best_trial, uint64_t *input_cost,
                            uint64_t *all_costs /*[ntrials] or NULL*/);

/* Introspection for bench.py: kernels launched by this library since load / reset. */
uint64_t pk_launch_count(void);
void pk_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* PK_CAPI_H */
