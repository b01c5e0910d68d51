// pk_kernels.h -- launch interface between the C ABI (pk_capi.cu) and the per-(m,t)
// template instantiations of the sm_100a kernels (pk_kernels.cuh, pk_inst_m*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pk_capi.h"
#include "pk_code.h"

#define PK_HUGE_CAP 4096   // records behind the main parked-frame list for the huge frames
#define PK_MEGA_SLOTS 256  // long searches open to helpers at any one time
struct PkMegaSlot;

// warps per phase-B CTA
#define PK_WARPS_B 4        // bit-sliced decoder (needs ~170 registers)
#ifndef PK_WARPS_B_CT
#define PK_WARPS_B_CT 4     // class-table mode (8: no faster in bulk, slower tails at mid SNR)
#endif
#ifndef PK_WARPS_B_LUT
#define PK_WARPS_B_LUT 16   // coset-table mode: a whole CTA searches one long frame, 16K patterns per step
                            // (the uncapped searches of these codes end in a few monster frames: latency matters)
#endif

// Per-decoder run-time parameters (constant for the life of a pk_kaneko handle).
struct PkKanekoParams {
    double llr_den;       // pow(sd0, 2): alpha = 2*y / llr_den  (KanekoKernelProcessor.cpp:337)
    int J;                // cap on T, < 0 = none (KanekoKernelProcessor.cpp:392-393)
    uint32_t max_trials;  // safety cap, 0x7FFFFFFF = the reference bound
    int frames_per_grab;  // frames a warp takes from the queue per atomic
    uint32_t limit_a;     // trials (multiple of 32) a frame may spend in the narrow phase A before it is parked
    uint32_t big_span;    // parked frames with at least this many patterns left are searched by a whole CTA
    uint32_t huge_span;   // ... and with at least this many go to the separate "huge" list (always cooperative)
    int variant;          // 0: decode(answer, word, res) (:335-407); 1: decode(word, res), the file-mode flavour (:212-276);
                          // 2: exact rules (ours, for the kernel-LLR bridge): bound 1 << T, calcRightSide over d - m positions, calcT without the border sum
    int ext;              // 1: extended code -- position n carries the overall parity of the BCH codeword (ours: src/main.cpp:60 only builds n = 2^m - 1)
    uint32_t extra_ops;   // per-frame constant added to both synthetic counters (2n+1 sort cost of the file-mode flavour, :221-224)
    // long searches shared by the whole grid (see "mega frames" in pk_kernels.cuh)
    uint32_t mega_chunk;    // patterns per helper chunk (multiple of 1024 * warps per phase-B CTA)
    uint32_t mega_span;     // a cooperative search with at least this many patterns left is opened to helpers ...
    uint32_t mega_after;    // ... once it has run this many patterns on its own CTA
    uint32_t solo_patterns; // a one-warp search still running after this many patterns is handed to a whole CTA (late list)
    uint32_t epoch;         // launch counter (marks the late-list records of THIS launch as written)
};

// Generation-mode parameters of one launch.
struct PkGenParams {
    double sigma;         // channel noise sd at this SNR point (dataForPlot.cpp:45)
    uint64_t seed;
    uint64_t first_frame;
    uint32_t snr_index;
};

// Where frames come from / go to (one launch pair).
struct PkIo {
    // replay mode
    const double *y;
    uint8_t *decided;
    uint32_t *trials;
    // generation mode
    PkGenParams gp;
    uint8_t *d_info, *d_cw;
    double *d_y;
    int dump_only;
    // both
    uint32_t *zscratch;   // phase B of large codes: root-word scratch, grid_b * 4 warps * n * 32 words
    PkMegaSlot *mega;     // [PK_MEGA_SLOTS] per launch slot
    uint32_t *mega_bits;  // [PK_MEGA_SLOTS][PK_MEGA_WORDS] clean-chunk bitmaps
    pk_frame_rec *recs;
    unsigned long long *totals;
};

// A frame parked by phase A for the wide search of phase B.
struct PkLongRec {
    double l0;
    uint32_t frame, base, bound, m0;
    uint32_t tsteps, nimpr, sflags, pad;   // sflags: bit0 first_ok, bit1 have, bits 8.. = frame flags; pad: late list: == launch epoch once written
    uint32_t bestF[8];
};
// Device-side control block of one phase A / phase B launch pair (zeroed by the launcher).
struct PkPhaseCtl {
    unsigned long long queue_a;     // next frame of phase A
    unsigned long long queue_b;     // next parked "small" frame of phase B (one warp each)
    unsigned long long n_long;      // parked small frames: longs[0 .. n_long)
    unsigned long long n_total;     // parked frames of both kinds (capacity check)
    unsigned long long queue_big;   // next parked "big" frame (one CTA each)
    unsigned long long n_big;       // parked big frames: longs[cap-1 .. cap-n_big] (filled from the top)
    unsigned long long n_huge;      // parked huge frames: longs[cap .. cap+n_huge) (may count past PK_HUGE_CAP: the overflow went to the big list)
    // phase B end game
    unsigned long long n_late;          // one-warp searches handed over to a whole CTA: longs[cap + PK_HUGE_CAP ..) (cap of them at most)
    unsigned long long queue_late;      // next of them to take
    unsigned long long mega_mask[PK_MEGA_SLOTS / 64];   // slots with an open window
    unsigned long long ctas_past_solo;  // CTAs that have left the one-warp loop: once all have, n_late is final
    unsigned long long masters;         // CTAs that are (or are about to become) the master of a late frame
};

// A long search opened to the idle CTAs of the grid ("mega frame").  The master CTA keeps the sequential state and
// commits in pattern order; helpers scan chunks AHEAD of it against a snapshot of (l0, best codeword) and mark the ones
// that hold no possible improvement in a bitmap, which the master then skips.  A slot covers a WINDOW of
// PK_MEGA_WINDOW chunks from the master's position; the master re-registers (same slot, next generation) when it
// reaches the end of the window and gives the slot back when the frame is finished.
struct PkMegaSlot {
    unsigned int owner;      // 0 = free, 1 = taken by a master (atomicCAS)
    unsigned int state;      // 0 = fields invalid, else the generation (never 0) the fields below belong to
    unsigned int gen;        // last generation handed out (owner only)
    unsigned int seq;        // sequence lock over (l0, have, bestF): odd while the master writes
    unsigned int pos;        // chunk the master is working on (helpers take later ones)
    unsigned int next;       // next chunk to hand to a helper
    unsigned int bound;      // pattern bound as of the last commit (chunks beyond it are not handed out)
    unsigned int frame;      // frame index in the batch
    unsigned int g0;         // first pattern of chunk 0 (the search's step grid starts there)
    unsigned int cbase;      // first chunk of the window
    unsigned int limit;      // one past the last chunk of the window
    unsigned int have;
    double l0;
    unsigned int bestF[8];
};
// bitmap word of a window: clean flags of 16 chunks in the low half, generation tag in the high half -- a helper that was
// scanning for an earlier generation of the slot cannot mark anything (its compare-and-swap fails on the tag)
#define PK_MEGA_WINDOW 2048                      // chunks per window
#define PK_MEGA_WORDS (PK_MEGA_WINDOW / 16)      // bitmap words per slot

// what the launcher zeroes before every launch pair: the control words and, right behind them, the mega slots
struct PkPhaseBlock {
    PkPhaseCtl ctl;
    PkMegaSlot mega[PK_MEGA_SLOTS];
};


enum { PK_MODE_ALG = 0, PK_MODE_LUT = 1, PK_MODE_CLASS = 2 };


struct PkLaunchGeom {
    int grid, block;
    size_t smem;
};

struct PkKernelSet {
    int m, t;
    bool has_bitsliced;   // phase B runs the bit-sliced decoder (pk_bs.cuh) for this code
    bool has_class;       // phase B can run on a cyclic-class table (PkClassTable) instead
    // host instance of the algebraic decoder (coset-table construction)
    bool (*host_alg_decode)(const uint32_t *, const uint8_t *, const uint16_t *, uint32_t *);
    // geometry (queries the device once; sets the dynamic shared-memory attributes):
    // out[0..1] = replay phase A / B, out[2..3] = generation phase A / B (grid 0 = no such kernel)
    // mode: PK_MODE_ALG (BM+Chien / bit-sliced), PK_MODE_LUT (coset table) or PK_MODE_CLASS (cyclic-class table)
    cudaError_t (*geom_kaneko)(int mode, int nk, int sm_count, PkLaunchGeom *out);
    cudaError_t (*geom_bdd)(int sm_count, PkLaunchGeom *out);
    // Kaneko decode of B frames: phase A (+ phase B when a wide kernel exists and long_cap > 0)
    cudaError_t (*launch_kaneko)(int mode, bool gen, const PkLaunchGeom *g4, const PkDevTables &tb,
                                 const PkKanekoParams &kp, const PkIo &io, long B, PkPhaseCtl *ctl, PkLongRec *longs,
                                 long long_cap, cudaStream_t st);
    // algebraic decoder alone, one thread per word
    cudaError_t (*launch_bdd)(const PkLaunchGeom &g, const PkDevTables &tb, const uint8_t *d_words, long B,
                              uint8_t *d_answers, uint8_t *d_ok, cudaStream_t st);
    // encoder, one warp per frame
    cudaError_t (*launch_encode)(const PkDevTables &tb, const uint8_t *d_info, long B, uint8_t *d_cw,
                                 cudaStream_t st);
};

// null when no kernel is instantiated for (m,t)
const PkKernelSet *pk_find_kernels(int m, int t);

// bumped by every launch wrapper (pk_launch_count())
extern unsigned long long g_pk_launches;
