// pk_polar.h -- host-side description of a mixed-kernel polar code with binary matrix kernels
// (the reference's CMixedKernelEncoder / CListKernelEngine set-up, out/external/MixedKernelEncoder.cpp:7-97,
// KernelListEngine.cpp:6-170, and the per-phase kernel trellises of TrellisKernelProcessor.cpp:69-179).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#define PK_POLAR_MAX_LAYERS 4

// One binary l x l kernel with, for every phase, the minimal trellis of the coset code spanned by rows
// phase..l-1 (row `phase` tagged by an extra column), stored in GATHER form for the Viterbi kernel.
struct PkKernelTrellis {
    int size = 0;                              // l
    std::vector<uint8_t> matrix;               // [l][l] row-major 0/1 (Kernel.cpp:93-107 file format)
    // per phase p, per section j in [0,l): number of state bits AFTER section j
    std::vector<uint8_t> ab;                   // [l][l+1]: ab[p][0] = 0, ab[p][j+1]
    int max_ab = 0;
    // predecessor table: entry for (phase p, section j, next state s1) at pred[off[p*l + j] + s1]:
    //   two 16-bit halves, each = previous state | (branch bit << 15), 0xFFFF = no such branch
    std::vector<uint32_t> pred;
    std::vector<uint32_t> off;                 // [l*l]
    // The same trellises for IN-PLACE evaluation (k_polar_lanes): every generator row keeps ONE bit position of the state
    // index for its whole span (a row that starts where another ends inherits its position), so a section only ever
    // combines the state pairs (x, x | 1 << q) and writes its results back over them.
    //   ip_x: per section the states x (bit q clear) | t << 15, t = label of the branch leaving x with the starting
    //         row's coefficient 0; padded with 0xFFFF to a multiple of 4.
    //   ip_sec[p][j], j < l: {entry offset | groups of 4 << 24, q | type << 8}; type 0: no row starts or ends (or one does
    //         both: nothing to do, 0 groups), 1: a row starts at q, 2: a row ends at q, 3: both (butterfly).
    //   ip_sec[p][l]: {0, position of the tagged row}: the LLR is M[1 << q] - M[0].
    std::vector<uint16_t> ip_x;
    std::vector<uint32_t> ip_sec;              // [l][l+1][2]
    int ip_bits = 0;                           // state index bits used by the in-place numbering
    bool ip_ok = true;                         // false: a section has more groups than the 8-bit count holds (k_polar_lanes is not used)
};

struct pk_polar_code {
    int N = 0, K = 0, N0 = 0, layers = 0, min_dist = 0;   // N0 = unshortened length (product of kernel sizes)
    std::vector<int> ksize;                     // kernel size per layer (layer 0 = outermost)
    std::vector<int> kid;                       // index into `kernels` per layer
    std::vector<PkKernelTrellis> kernels;       // distinct kernels
    std::vector<int> outer;                     // outer[j] = N0 / prod_{i<j} ksize[i], outer[layers] = 1
    std::vector<uint8_t> symtype;               // [N0] 0 normal, 1 shortened, 2 punctured (empty = none)
    std::vector<int> decision;                  // [N0] constraint index of a frozen symbol, -1 = information symbol
    std::vector<std::vector<int>> constraints;  // terms of every freezing constraint (last = the frozen symbol)
    std::vector<uint32_t> cmask;                // [N0][N0/32] mask of the earlier symbols a frozen symbol depends on
    std::vector<int> info_pos;                  // positions of the K information symbols, ascending
};

// Parses the reference's code specification format (MixedKernelEncoder.cpp:10-93): kernels named "-file" or
// "<file" are loaded from `file` (size, then size*size 0/1 integers).  Returns "" or an error message.
std::string pk_polar_parse(pk_polar_code &c, const std::string &spec_text);
// Builds the per-phase trellises of one kernel.
std::string pk_polar_build_trellis(PkKernelTrellis &k);
// Gather-form and in-place tables of a kernel against each other on random integer costs (host, tests).
std::string pk_polar_check_inplace(const PkKernelTrellis &k, unsigned long long seed, int ntests);
// (2^m) x (2^m) extended BCH kernel of the reference's newer makeMatrix (root bchCoder.cpp:356-389).
void pk_polar_ebch_kernel(int m, std::vector<uint8_t> &out);
