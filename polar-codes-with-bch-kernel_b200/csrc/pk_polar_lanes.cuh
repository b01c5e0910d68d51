// pk_polar_lanes.cuh -- SC / SC-list decoder with LIST PATHS (AND FRAMES) ACROSS LANES.  Included by pk_polar.cu.
//
// Every path of every frame walks the same phases, the same kernel trellises and the same sections: only the numbers
// differ.  k_polar_lanes therefore gives every (frame, path) pair one of NSLOT = 32 / G slots of a warp -- FPW = NSLOT / L
// frames of L paths each -- and runs the Viterbi recursion of CTrellisKernelProcessor::GetLLRs
// (TrellisKernelProcessor.cpp:260-293) with control flow that is uniform over the warp.  (k_polar_decode, the general
// fall-back, spends one warp per path with lanes across trellis states: the 2..16-state sections at both ends of every
// pass leave it mostly idle, and it ran 1.5 warp instructions per branch evaluation where this kernel runs 0.21.)
//
// The recursion runs IN PLACE on one metric buffer M[state][slot] in shared memory: the trellises are renumbered on the
// host (pk_polar.h: every generator row keeps one bit position of the state index for its whole span) so that a section
// only ever combines the pairs (x, x + q) and writes the results back over them -- one load and one store per state,
// no second buffer, no ordering between the pairs of a section.  A section's pairs come from a staged table (16-bit
// entries, the same for all slots: broadcast loads) in batches of eight, the G lanes of a slot taking interleaved
// entries; every access of a warp to M is one wavefront.
//
// Bits are packed ACROSS SLOTS: the partial-sum arrays C, the kernel-processor offsets and the decided symbols are one
// 32-bit word per element whose bit s belongs to slot s, so IterativelyUpdateC (KernelListEngine.cpp:266-315) and the
// offset update (TrellisKernelProcessor.cpp:245-259) are word-wide XORs, one element per lane.
//
// List management (MixedKernelListDecoder.cpp:100-185) runs in registers: both candidates of a path are ranked against
// the 2 L candidates of its frame with shuffles, the path-index stack (TVMemoryEngine.cpp:85-141, misc.h:206-226) is
// replayed with ranks inside the kill / clone masks.  Cloning copies the innermost arrays at once (two words per kernel
// row) and defers everything else: `cmap` names the slot whose column still holds a path's copy of the outer arrays; the
// columns are brought home when an innermost block completes, which is the only time outer arrays are written.
//
// Same fp32 operations in the same order per state as the reference (one add per branch, one min per state -- except
// that adding the zero cost of a branch that agrees with the hard decision is skipped, x + 0.0f being x; the list
// metrics as in :83, :121, :170), so LLRs, metrics and lists are bit-identical.
#pragma once

struct PkLanesKernel {
    const uint32_t *tab;                // in-place entries, two per word: byte offset of the metric row of state x | PAD << 1 | t (pk_polar.h)
    const uint32_t *sec;                // [l][l+1][2]: {byte offset of the entries | batches of 8 << 24, byte offset of row 1 << q | (type | TINY << 2 | 8) << 16}
    const unsigned long long *masks;    // [2 l]: column masks (bit r: K[r][c]), then row masks (bit c: K[r][c])
    int ntab, size;
};
struct PkLanesDev {
    PkLanesKernel k[PK_POLAR_MAX_LAYERS];   // distinct kernels
    int kidx[PK_POLAR_MAX_LAYERS];          // layer -> distinct kernel
    int nk, ns_rows;                        // metric rows (states of the widest section)
};
struct LanesLayout {
    int tab[PK_POLAR_MAX_LAYERS], sec[PK_POLAR_MAX_LAYERS], masks[PK_POLAR_MAX_LAYERS];   // byte offsets of the staged tables
    int tables;                                                                           // bytes of all tables
    int met, S, W, U, st, par, per_warp;                                                  // byte offsets inside a warp's region
    int nwords;
};
__host__ __device__ inline LanesLayout lanes_layout(const PkPolarDev &d, const PkLanesDev &ld, const PathLayout &pl, int L, int nslot) {
    LanesLayout y;
    int o = 0;
    for (int k = 0; k < ld.nk; ++k) {
        const int l = ld.k[k].size;
        y.tab[k] = o; o += ld.k[k].ntab * 4;
        o = (o + 7) & ~7;
        y.sec[k] = o; o += l * (l + 1) * 8;
        o = (o + 7) & ~7;
        y.masks[k] = o; o += 2 * l * 8;
        o = (o + 15) & ~15;
    }
    y.tables = o;
    int w = 0;
    y.met = w; w += ld.ns_rows * nslot * 4;
    y.S = w; w += (pl.floats + 1) * nslot * 4;
    y.nwords = pl.u_off;
    y.W = w; w += y.nwords * 4;
    y.U = w; w += d.N0 * 4;
    y.st = w; w += (nslot / L) * (L + 1) * 4;
    y.par = w; w += 32 * 4;
    y.per_warp = (w + 15) & ~15;
    return y;
}

// One section of the in-place recursion: the lane's share of the state pairs (x, x + q) of the section.  The number of
// pairs is a power of two: sections of 8 and more run in batches of 8 pairs per slot (2 CH per lane) whose loads are all
// issued before the first store (the pairs of a section are disjoint, so the order inside a section is free); sections of
// 1, 2 or 4 pairs are one group of four entries padded with PAD entries whose stores are suppressed (TINY).
// `ay` = |LLR| of the section's symbol, `hd` = its hard decision: a branch labelled like the hard decision costs nothing,
// the other one `ay` (x + 0.0f = x: the reference's add of a zero cost is skipped, same value).  sw = t ^ hd.
//   TYPE 0: M[x]  = M[x] + c(t)                                  (no generator row starts or ends)
//   TYPE 1: M[x], M[x+q] = M[x] + c(t), M[x] + c(!t)             (a row starts)
//   TYPE 2: M[x]  = min(M[x] + c(t), M[x+q] + c(!t))             (a row ends)
//   TYPE 3: M[x]  = min(M[x] + c(t), M[x+q] + c(!t)), M[x+q] = min(M[x] + c(!t), M[x+q] + c(t))   (both: butterfly)
// Entry (16 bits): byte offset of the metric row of x (a multiple of 32) | PAD << 1 | t.
// (Two batches per step, software-pipelined batches and repacked entry words were measured and dropped: profiles/r2_notes.md 3.)
template <int G, int TYPE, bool TINY>
__device__ __forceinline__ void lanes_section(const unsigned char *__restrict__ tb, int nbatch, unsigned char *mcol, uint32_t qoff, float ay, uint32_t hd) {
    constexpr int CH = 4 / G, NP = TINY ? CH : 2 * CH;
    for (int it = 0; it < (TINY ? 1 : nbatch); ++it, tb += 16) {
        uint32_t e[NP];
#pragma unroll
        for (int k = 0; k < NP / CH; ++k) {
            if constexpr (CH == 4) {
                const uint2 q = *reinterpret_cast<const uint2 *>(tb + 8 * k);
                e[4 * k] = q.x; e[4 * k + 1] = q.x >> 16; e[4 * k + 2] = q.y; e[4 * k + 3] = q.y >> 16;
            } else if constexpr (CH == 2) {
                const uint32_t q = *reinterpret_cast<const uint32_t *>(tb + 8 * k);
                e[2 * k] = q; e[2 * k + 1] = q >> 16;
            } else {
                e[k] = *reinterpret_cast<const uint16_t *>(tb + 8 * k);
            }
        }
        float a[NP], b[NP];
        float *px[NP];
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            px[u] = reinterpret_cast<float *>(mcol + (e[u] & 0xFFE0u));
            a[u] = *px[u];
            if (TYPE >= 2) b[u] = *reinterpret_cast<const float *>(reinterpret_cast<const unsigned char *>(px[u]) + qoff);
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            float *pq = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(px[u]) + qoff);
            const bool sw = ((e[u] ^ hd) & 1u) != 0, ok = !TINY || !(e[u] & 2u);
            if (TYPE == 0) {
                const float P = a[u] + ay;
                if (ok) *px[u] = sw ? P : a[u];
            } else if (TYPE == 1) {
                const float P = a[u] + ay;
                if (ok) { *px[u] = sw ? P : a[u]; *pq = sw ? a[u] : P; }
            } else if (TYPE == 2) {
                const float r0 = fminf(a[u], b[u] + ay), r1 = fminf(a[u] + ay, b[u]);
                if (ok) *px[u] = sw ? r1 : r0;
            } else {
                const float r0 = fminf(a[u], b[u] + ay), r1 = fminf(a[u] + ay, b[u]);
                if (ok) { *px[u] = sw ? r1 : r0; *pq = sw ? r0 : r1; }
            }
        }
    }
}

// One Viterbi pass (TrellisKernelProcessor.cpp:260-293) for stride element i of every slot of the warp, on the in-place
// numbering of the trellis (pk_polar.h); returns M[tag = 1] - M[tag = 0] (:292) in all lanes of the slot.
//   tab/sec: staged tables (shared); src: the slot's value of section 0 (then + step per section); offw: offset words.
template <int G>
__device__ __forceinline__ float lanes_viterbi(const unsigned char *__restrict__ tab, const uint2 *__restrict__ sec, int l,
                                               const float *src, int src_step, const uint32_t *offw,
                                               int off_step, int slot, int g, unsigned char *mcol) {
    constexpr int CH = 4 / G;
    if (g == 0) *reinterpret_cast<float *>(mcol) = 0.0f;
    if (G > 1) __syncwarp();
    float y = src[0];   // (plain loads: the transposed channel LLRs were written by this warp)
    uint32_t ow = offw[0];
    uint2 sj = sec[0];
    for (int j = 0; j < l; ++j) {
        if ((ow >> slot) & 1u) y = -y;
        const float ay = fabsf(y);
        const uint32_t hd = y < 0.0f ? 1u : 0u;
        const unsigned char *tb = tab + (sj.x & 0xFFFFFFu) + g * CH * 2;
        const int nb = (int)(sj.x >> 24);
        const uint32_t qoff = sj.y & 0xFFFFu, type = sj.y >> 16;   // type | TINY << 2 | something to do << 3
        // next section's operands while this one runs
        if (j + 1 < l) { src += src_step; offw += off_step; }
        y = *src;
        ow = *offw;
        sj = sec[j + 1];   // (sec[l] holds the position of the tagged row)
        switch (type) {
        case 8 + 0: lanes_section<G, 0, false>(tb, nb, mcol, qoff, ay, hd); break;
        case 8 + 1: lanes_section<G, 1, false>(tb, nb, mcol, qoff, ay, hd); break;
        case 8 + 2: lanes_section<G, 2, false>(tb, nb, mcol, qoff, ay, hd); break;
        case 8 + 3: lanes_section<G, 3, false>(tb, nb, mcol, qoff, ay, hd); break;
        case 12 + 0: lanes_section<G, 0, true>(tb, nb, mcol, qoff, ay, hd); break;
        case 12 + 1: lanes_section<G, 1, true>(tb, nb, mcol, qoff, ay, hd); break;
        case 12 + 2: lanes_section<G, 2, true>(tb, nb, mcol, qoff, ay, hd); break;
        case 12 + 3: lanes_section<G, 3, true>(tb, nb, mcol, qoff, ay, hd); break;
        default: break;
        }
        if (G > 1) __syncwarp();
    }
    const float r = *reinterpret_cast<const float *>(mcol + (sj.y & 0xFFFFu)) - *reinterpret_cast<const float *>(mcol);
    if (G > 1) __syncwarp();
    return r;
}

// L = slots per frame (a power of two), Lr = the list size (<= L: the slots Lr .. L-1 of a frame are never used)
template <int L, int G>
__global__ void __launch_bounds__(G == 1 ? 256 : 512, 1)
k_polar_lanes(PkPolarDev d, PkLanesDev ld, int Lr, const float *__restrict__ llr_in, long B, int *__restrict__ count,
              uint8_t *__restrict__ inf_out, uint8_t *__restrict__ cw_out, float *__restrict__ metric_out, float *__restrict__ scr_chan) {
    constexpr int NSLOT = 32 / G, FPW = NSLOT / L;
    static_assert(L * G <= 32 && (L & (L - 1)) == 0 && (G == 1 || G == 2 || G == 4), "slots");
    constexpr uint32_t SLOTS = NSLOT == 32 ? 0xFFFFFFFFu : ((1u << NSLOT) - 1u);   // the g = 0 lanes: lane == slot
    constexpr uint32_t GM = L == 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
    extern __shared__ __align__(16) unsigned char smem[];
    // (the warp index through a shuffle: the compiler then knows it is uniform over the warp and keeps what derives from
    // it -- the warp's shared-memory region -- in uniform registers)
    const int lane = threadIdx.x & 31, warp = __shfl_sync(PKP_FULL, (int)(threadIdx.x >> 5), 0), nwarps = blockDim.x >> 5;
    const int g = lane / NSLOT, slot = lane % NSLOT, fslot = slot / L, p = slot % L, gs = fslot * L;
    const PathLayout pl = path_layout(d);
    const LanesLayout ly = lanes_layout(d, ld, pl, L, NSLOT);
    // ---- stage the tables of every distinct kernel
    for (int k = 0; k < ld.nk; ++k) {
        const int l = ld.k[k].size;
        uint32_t *t = reinterpret_cast<uint32_t *>(smem + ly.tab[k]);
        for (int i = threadIdx.x; i < ld.k[k].ntab; i += blockDim.x) t[i] = ld.k[k].tab[i];
        uint32_t *s = reinterpret_cast<uint32_t *>(smem + ly.sec[k]);
        for (int i = threadIdx.x; i < l * (l + 1) * 2; i += blockDim.x) s[i] = ld.k[k].sec[i];
        unsigned long long *m = reinterpret_cast<unsigned long long *>(smem + ly.masks[k]);
        for (int i = threadIdx.x; i < 2 * l; i += blockDim.x) m[i] = ld.k[k].masks[i];
    }
    __syncthreads();
    unsigned char *wb = smem + ly.tables + (size_t)warp * ly.per_warp;
    float *met = reinterpret_cast<float *>(wb + ly.met);
    float *S = reinterpret_cast<float *>(wb + ly.S);
    uint32_t *W = reinterpret_cast<uint32_t *>(wb + ly.W);
    uint32_t *U = reinterpret_cast<uint32_t *>(wb + ly.U);
    uint32_t *st = reinterpret_cast<uint32_t *>(wb + ly.st) + fslot * (L + 1);
    uint32_t *par = reinterpret_cast<uint32_t *>(wb + ly.par);
    const long gw = (long)blockIdx.x * nwarps + warp, tw = (long)gridDim.x * nwarps;
    float *chanT = scr_chan + (size_t)gw * d.N0 * FPW;   // [N0][FPW]
    const int last = d.layers - 1, lsz = d.ksize[last];
    const int cm_off = pl.c_off[d.layers], ol_off = pl.o_off[last];   // innermost C array and offsets: copied eagerly

    // (the trip count depends on the CTA only: control flow that is uniform over the whole CTA, not just the warp, measured
    // 9 % faster at L = 32; a warp past the end runs one idle group)
    for (long grp0 = (long)blockIdx.x * nwarps; grp0 * FPW < B; grp0 += tw) {
        const long grp = grp0 + warp;
        const long fr = grp * FPW + fslot;
        const bool live = fr < B;
        // ---- LoadLLRs (MixedKernelEncoder.cpp:179-203), transposed: chanT[i][frame slot]
        for (int idx = lane; idx < d.N0 * FPW; idx += 32) {
            const int i = idx / FPW, fs = idx - i * FPW;
            const long f2 = grp * FPW + fs;
            float v = 0.0f;
            if (f2 < B) {
                if (!d.symtype) v = llr_in[f2 * d.N + i];
                else if (d.symtype[i] == 1) v = PKP_UPPER;
                else if (d.symtype[i] == 2) v = 0.0f;
                else v = llr_in[f2 * d.N + d.compact[i]];
            }
            chanT[idx] = v;
        }
        // ---- Cleanup + AssignInitialPath (TVMemoryEngine.cpp:58-94): the first Pop of the lazily initialised stack is L - 1
        bool act = live && p == Lr - 1;
        float R = 0.0f;
        int cnt = Lr - 1;       // free path indices on the stack (positions 0 .. cnt-1)
        int cmap = slot;        // column holding this path's copy of the outer arrays
        if (L > 1) {
            if (g == 0) st[p] = (uint32_t)p;
        }
        __syncwarp();

        for (int phi = 0; phi < d.N0; ++phi) {
            // ---- LLR of symbol phi (IterativelyCalcS, KernelListEngine.cpp:370-447)
            int pv = phi, mm = last;
            while (mm > 0 && (pv % d.ksize[mm]) == 0) { pv /= d.ksize[mm]; --mm; }
            float v = 0.0f;
            for (int j = mm; j <= last; ++j) {
                const int kx = ld.kidx[j], l = d.ksize[j];
                const int stride = d.outer[j + 1], phase = pv % l;
                uint32_t *offs = W + pl.o_off[j];
                // offset update for the newly known input phase-1 (TrellisKernelProcessor.cpp:245-259)
                if (phase == 0) {
                    for (int e = lane; e < l * stride; e += 32) offs[e] = 0;
                } else {
                    const unsigned long long row = reinterpret_cast<const unsigned long long *>(smem + ly.masks[kx])[l + phase - 1];
                    const uint32_t *known = W + pl.c_off[j + 1] + (phase - 1) * stride;
                    for (int e = lane; e < l * stride; e += 32) {
                        const int c = e / stride, i = e - c * stride;
                        if ((row >> c) & 1ull) offs[e] ^= known[i];
                    }
                }
                __syncwarp();
                const unsigned char *tab = smem + ly.tab[kx];
                const uint2 *sec = reinterpret_cast<const uint2 *>(smem + ly.sec[kx]) + phase * (l + 1);
                float *dest = S + (size_t)pl.s_off[j + 1] * NSLOT;
                // one call site (the recursion is inlined once): layer 0 reads the transposed channel LLRs, the others the
                // S array of the layer above through the path's column
                const float *src = j == 0 ? chanT + fslot : S + (size_t)pl.s_off[j] * NSLOT + cmap;
                const int unit = j == 0 ? FPW : NSLOT;
                for (int i = 0; i < stride; ++i) {
                    v = lanes_viterbi<G>(tab, sec, l, src + (size_t)i * unit, stride * unit, offs + i, stride, slot, g, reinterpret_cast<unsigned char *>(met + slot));
                    if (g == 0) dest[(size_t)i * NSLOT + slot] = v;
                }
                __syncwarp();
                pv = 0;
            }
            // ---- decision
            const bool frozen = d.frozen[phi] != 0;
            uint32_t bit = 0;
            int from = slot;   // slot this path was cloned from in this phase
            if (frozen) {
                // ContinuePathsFrozen (MixedKernelListDecoder.cpp:61-98)
                if (!d.all_static) {
                    const int blk0 = phi - phi % lsz;
                    for (int w = 0; w < d.nw; ++w) {
                        uint32_t m = d.cmask[phi * d.nw + w];
                        while (m) {
                            const int t = 32 * w + __ffs(m) - 1;
                            m &= m - 1;
                            bit ^= t >= blk0 ? (W[cm_off + t - blk0] >> slot) : (U[t] >> cmap);
                        }
                    }
                    bit &= 1u;
                }
                if (act && ((bit != 0) ^ (v < 0.0f))) R -= fabsf(v);
            } else if (L == 1) {
                // one path: a zero LLR ties the two candidates and the larger index, the flipped decision 1, wins
                bit = (v < 0.0f || v == 0.0f) ? 1u : 0u;
            } else {
                // ContinuePathsUnfrozen (:100-185)
                const uint32_t amask = __ballot_sync(PKP_FULL, act);
                const uint32_t D = v < 0.0f ? 1u : 0u;
                const float ma = R, mb = R - fabsf(v);
                const uint32_t sa = 2u * p + D, sb = 2u * p + (D ^ 1u);
                int ra = 0, rb = 0;
                for (int k = 0; k < L; ++k) {
                    const float Rk = __shfl_sync(PKP_FULL, R, gs + k), vk = __shfl_sync(PKP_FULL, v, gs + k);
                    if (!((amask >> (gs + k)) & 1u)) continue;
                    const uint32_t Dk = vk < 0.0f ? 1u : 0u;
                    const float m1 = Rk, m2 = Rk - fabsf(vk);
                    const uint32_t s1 = 2u * k + Dk, s2 = 2u * k + (Dk ^ 1u);
                    ra += (m1 > ma || (m1 == ma && s1 > sa)) ? 1 : 0;
                    ra += (m2 > ma || (m2 == ma && s2 > sa)) ? 1 : 0;
                    rb += (m1 > mb || (m1 == mb && s1 > sb)) ? 1 : 0;
                    rb += (m2 > mb || (m2 == mb && s2 > sb)) ? 1 : 0;
                }
                const int J = 2 * __popc((amask >> gs) & GM), keep = J < Lr ? J : Lr;
                const bool ka = act && ra < keep, kb = act && rb < keep;
                const bool kill = act && !ka && !kb, clone = ka && kb;
                bit = ka ? D : (D ^ 1u);   // the one continuation, or the better one of two (C = LLR < 0, :165)
                const uint32_t km = (__ballot_sync(PKP_FULL, kill) >> gs) & GM, cm = (__ballot_sync(PKP_FULL, clone) >> gs) & GM;
                const int nk = __popc(km), nc = __popc(cm);
                const uint32_t below = (1u << p) - 1u;
                if (g == 0) par[slot] = 0xFFu;
                // KillPath pushes in ascending path order, then ClonePath pops (:137-176)
                if (kill && g == 0) st[cnt + __popc(km & below)] = (uint32_t)p;
                __syncwarp();
                if (clone && g == 0) par[gs + st[cnt + nk - 1 - __popc(cm & below)]] = (uint32_t)slot;
                cnt += nk - nc;
                __syncwarp();
                const uint32_t pr = par[slot];
                if (kill) act = false;
                const float Rb = __shfl_sync(PKP_FULL, mb, pr & 31u);
                const uint32_t bp = __shfl_sync(PKP_FULL, bit, pr & 31u);
                const int cp = __shfl_sync(PKP_FULL, cmap, pr & 31u);
                if (pr != 0xFFu) {   // a new path: the other continuation of path pr
                    from = (int)pr;
                    act = true;
                    R = Rb;
                    bit = bp ^ 1u;
                    cmap = cp;
                }
                // clones take their parent's innermost arrays now
                const uint32_t newm = __ballot_sync(PKP_FULL, from != slot) & SLOTS;
                if (newm) {
                    for (int i0 = 0; i0 < 2 * lsz; i0 += 32) {   // uniform trip count: every lane takes part in the shuffles
                        const int idx = i0 + lane;
                        const bool mine = idx < 2 * lsz;
                        uint32_t *wp = idx < lsz ? W + cm_off + idx : W + ol_off + idx - lsz;
                        uint32_t w = mine ? *wp : 0u;
                        for (uint32_t m = newm; m; m &= m - 1) {
                            const int dd = __ffs(m) - 1;
                            const int ss = __shfl_sync(PKP_FULL, from, dd);
                            w = (w & ~(1u << dd)) | (((w >> ss) & 1u) << dd);
                        }
                        if (mine) *wp = w;
                    }
                    __syncwarp();
                }
            }
            if (!act) bit = 0;
            const uint32_t bw = __ballot_sync(PKP_FULL, bit != 0) & SLOTS;
            if (lane == 0) W[cm_off + phi % lsz] = bw;
            __syncwarp();
            // ---- an innermost block is complete: bring the outer arrays home, file the block's decided symbols
            if ((phi + 1) % lsz == 0) {
                if (L > 1) {
                    const uint32_t nonid = __ballot_sync(PKP_FULL, cmap != slot) & SLOTS;
                    if (nonid) {
                        const int nrows = pl.floats - 1;   // S arrays of layers 1 .. last
                        for (int r0 = 0; r0 < nrows; r0 += 8) {
                            float t[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) t[u] = r0 + u < nrows ? S[(size_t)(r0 + u) * NSLOT + cmap] : 0.0f;
                            __syncwarp();
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if (r0 + u < nrows && g == 0) S[(size_t)(r0 + u) * NSLOT + slot] = t[u];
                            __syncwarp();
                        }
                        const int ufirst = phi + 1 - lsz;   // decided symbols filed so far
                        const int total = ly.nwords + ufirst;
                        for (int i0 = 0; i0 < total; i0 += 32) {
                            const int idx = i0 + lane;
                            const bool mine = idx < total && !(idx >= cm_off && idx < cm_off + lsz) && !(idx >= ol_off && idx < ol_off + lsz) &&
                                              idx >= pl.c_off[1];
                            uint32_t *wp = idx < ly.nwords ? W + idx : U + (idx - ly.nwords);
                            uint32_t w = mine ? *wp : 0u, nw2 = w & ~nonid;
                            for (uint32_t m = nonid; m; m &= m - 1) {
                                const int dd = __ffs(m) - 1;
                                const int ss = __shfl_sync(PKP_FULL, cmap, dd);
                                nw2 |= ((w >> ss) & 1u) << dd;
                            }
                            if (mine) *wp = nw2;
                        }
                        cmap = slot;
                        __syncwarp();
                    }
                }
                for (int t = lane; t < lsz; t += 32) U[phi + 1 - lsz + t] = W[cm_off + t];
                __syncwarp();
            }
            // ---- propagate completed kernel blocks (IterativelyUpdateC, KernelListEngine.cpp:266-315)
            int lambda = d.layers, stride = 1;
            pv = phi;
            while (lambda > 0 && ((pv + 1) % d.ksize[lambda - 1]) == 0) {
                const int psi = pv / d.ksize[lambda - 1];
                const int next = stride * d.ksize[lambda - 1];
                const int phi0 = (lambda > 1) ? (psi % d.ksize[lambda - 2]) * next : 0;
                if (lambda > 1 || cw_out) {   // the outermost product is the codeword: only needed when it is asked for
                    const int l = d.ksize[lambda - 1];
                    const unsigned long long *cmk = reinterpret_cast<const unsigned long long *>(smem + ly.masks[ld.kidx[lambda - 1]]);
                    const uint32_t *csrc = W + pl.c_off[lambda];
                    uint32_t *cdst = W + pl.c_off[lambda - 1] + phi0;
                    for (int e = lane; e < l * stride; e += 32) {
                        const int c = e / stride, i = e - c * stride;
                        uint32_t x = 0;
                        for (unsigned long long m = cmk[c]; m; m &= m - 1) x ^= csrc[(__ffsll((long long)m) - 1) * stride + i];
                        cdst[e] = x;
                    }
                    __syncwarp();
                }
                stride = next;
                pv = psi;
                --lambda;
            }
        }

        // ---- final ordering (MixedKernelListDecoder.cpp:253-266): active paths by (R, index) descending
        const uint32_t amask = __ballot_sync(PKP_FULL, act) & SLOTS;
        int rank = 0;
        if (L > 1) {
            for (int k = 0; k < L; ++k) {
                const float Rk = __shfl_sync(PKP_FULL, R, gs + k);
                if (((amask >> (gs + k)) & 1u) && (Rk > R || (Rk == R && k > p))) ++rank;
            }
        }
        for (uint32_t m = amask; m; m &= m - 1) {
            const int s = __ffs(m) - 1;
            const int rk = __shfl_sync(PKP_FULL, rank, s);
            const long row = (grp * FPW + s / L) * Lr + rk;
            if (inf_out)
                for (int q = lane; q < d.K; q += 32) inf_out[row * d.K + q] = (uint8_t)((U[d.info_pos[q]] >> s) & 1u);
            if (cw_out) {
                const uint32_t *c0 = W + pl.c_off[0];
                for (int i = lane; i < d.N0; i += 32) {
                    if (!d.symtype) cw_out[row * d.N + i] = (uint8_t)((c0[i] >> s) & 1u);
                    else if (d.symtype[i] == 0) cw_out[row * d.N + d.compact[i]] = (uint8_t)((c0[i] >> s) & 1u);
                }
            }
        }
        if (act && g == 0 && metric_out) metric_out[fr * Lr + rank] = R;
        if (live && p == 0 && g == 0 && count) count[fr] = __popc((amask >> gs) & GM);
        __syncwarp();
    }
}
