// pk_kernels.h -- launch interface between the C ABI (pk_capi.cu) and the per-(m,t)
// template instantiations of the sm_100a kernels (pk_kernels.cuh, pk_inst_m*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pk_capi.h"
#include "pk_code.h"

// Per-decoder run-time parameters (constant for the life of a pk_kaneko handle).
struct PkKanekoParams {
    double llr_den;       // pow(sd0, 2): alpha = 2*y / llr_den  (KanekoKernelProcessor.cpp:337)
    int J;                // cap on T, < 0 = none (KanekoKernelProcessor.cpp:392-393)
    uint32_t max_trials;  // safety cap, 0x7FFFFFFF = the reference bound
    int frames_per_grab;  // frames a warp takes from the queue per atomic
};

// Generation-mode parameters of one launch.
struct PkGenParams {
    double sigma;         // channel noise sd at this SNR point (dataForPlot.cpp:45)
    uint64_t seed;
    uint64_t first_frame;
    uint32_t snr_index;
};

struct PkLaunchGeom {
    int grid, block;
    size_t smem;
};

struct PkKernelSet {
    int m, t;
    // host instance of the algebraic decoder (coset-table construction)
    bool (*host_alg_decode)(const uint32_t *, const uint8_t *, const uint16_t *, uint32_t *);
    // geometry (queries the device once; sets the dynamic shared-memory attribute)
    cudaError_t (*geom_kaneko)(bool lut, int nk, int sm_count, PkLaunchGeom *out /*[2]: replay, generate*/);
    cudaError_t (*geom_bdd)(int sm_count, PkLaunchGeom *out);
    // Kaneko decode of B frames from y (replay mode).  queue: device u32 zeroed by the callee.
    cudaError_t (*launch_replay)(bool lut, const PkLaunchGeom &g, const PkDevTables &tb, const PkKanekoParams &kp,
                                 const double *d_y, long B, uint8_t *d_decided, uint32_t *d_trials,
                                 pk_frame_rec *d_recs, unsigned long long *d_totals, unsigned int *d_queue,
                                 cudaStream_t st);
    // generate + encode + AWGN + decode + compare (generation mode)
    cudaError_t (*launch_generate)(bool lut, const PkLaunchGeom &g, const PkDevTables &tb, const PkKanekoParams &kp,
                                   const PkGenParams &gp, long B, pk_frame_rec *d_recs,
                                   unsigned long long *d_totals, unsigned int *d_queue,
                                   uint8_t *d_info, uint8_t *d_cw, double *d_y, int dump_only, cudaStream_t st);
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
