// pk_code.h -- host-side description of one primitive narrow-sense BCH code and the
// lookup tables the sm_100a kernels stage into shared memory.
#pragma once
#include <stdint.h>
#include <memory>
#include <string>
#include <vector>

struct PkKernelSet;  // pk_kernels.h

// Device-resident tables, passed to kernels by value.
struct PkDevTables {
    const uint8_t *mul;    // [2^(2m)]  GF(2^m) product table
    const uint16_t *xoff;  // [n]       alpha^{-p} << m  (Chien row offsets)
    const uint32_t *hcol;  // [n][NSW]  packed syndromes S_1..S_2t of x^p
    const uint32_t *rcol;  // [n]       x^p mod g(x)  (coset index bits), LUT mode
    const uint16_t *lut;   // [2^(n-k)] coset -> up to t error positions, LUT mode
    const uint32_t *gmask; // [NW]      g(x) as a bit mask
    int k;
    int nk;                // n - k
    // cyclic-class table (PkClassTable), class-table mode
    const uint32_t *ct_col;             // [n][NSW] columns in pair packing: field i = alpha^p | alpha^{j_i p} << m
    const uint8_t *ct_norm;             // [fields][2^m (S_j)][2^m (S_1)] rank of S_j * alpha^{-j log S_1}
    const uint8_t *ct_log;              // [2^m] log S_1 (0 for S_1 = 0)
    const uint32_t *ct_bits;            // [2^(kb+1) / 32] decodable-class bitmap
    const unsigned long long *ct_hash;  // [2^hbits] key << (t m) | t packed positions
    uint32_t ct_hshift, ct_hmask;       // slot = (key * 0x9E3779B1) >> hshift, linear probing under hmask
    unsigned long long ct_tex;          // cudaTextureObject_t over ct_bits (u32 texels): the bitmap gathers of the wide search go through the TEX path
    uint32_t ct_mult[8];                // 1 << (field offset in the key) per independent syndrome
};

// Cyclic-class table of the weight <= t error patterns (class-table mode of the wide search).
//
// A binary BCH syndrome is (S_1, S_3, .., S_{2t-1}); a cyclic shift of the error pattern by r multiplies S_j by
// alpha^{j r}.  Shifting by -log S_1 normalises S_1 to 1, so a pattern class is identified by the remaining
// S_j' = S_j alpha^{-j log S_1} alone: kb = n - k - m bits (each S_j' stored as its rank inside the subfield it lives
// in).  Syndromes with S_1 = 0 cannot be normalised and are keyed by their raw (S_3, .., S_{2t-1}) under a flag bit.
// `bits` says for every key whether a pattern of weight <= t has that syndrome class -- one 32-byte sector read
// per test pattern decides "the algebraic decoder would fail" for the ~99 % of patterns where it does; `hash` gives
// the (normalised) error positions of the rest.  Built once per (m, t) per process from an enumeration of all
// patterns of weight <= t that contain position 0.
struct PkClassTable {
    int kb = 0, hbits = 0;
    size_t entries = 0;
    std::vector<int> js;               // the odd j > 1 whose S_j is independent: one key field each (pk_ct_used)
    std::vector<uint8_t> norm, logt;
    std::vector<uint32_t> mult, bits, col;
    std::vector<uint64_t> hash;
    // verdict and error positions (bit mask, nw words) for packed syndromes S_1..S_2t (hcol layout); host mirror of the device lookup
    bool lookup(int m, int t, const uint32_t *packed, uint32_t *A) const;
};

struct pk_code {
    int m = 0, n = 0, k = 0, t = 0, gsize = 0, nk = 0;
    int device = 0;
    bool use_lut = false;              // coset table (n-k <= 16) in use
    bool use_ct = false;               // cyclic-class table in use
    std::shared_ptr<const PkClassTable> ct;
    std::vector<uint8_t> g;            // g(x), index = power of x   (main.cpp:80-95)
    std::vector<uint32_t> alog, log;   // antilog[i] = alpha^i, log[v]; log[0] = 0xFFFFFFFF sentinel
    std::vector<uint8_t> mul;
    std::vector<uint16_t> xoff;
    std::vector<uint32_t> hcol, rcol, gmask;
    std::vector<uint16_t> lut;
    PkDevTables dev{};                 // device copies
    std::vector<void *> dev_allocs;
    const PkKernelSet *ks = nullptr;   // sm_100a kernels instantiated for (m,t), or null
};

// Fills everything host-side (no CUDA calls).  Returns "" or an error message.
std::string pk_code_build_host(pk_code &c, int m, int t);
// n x n nested-BCH polarisation kernel (reference src/bchCoder.cpp:317-345).
void pk_code_kernel_matrix(const pk_code &c, uint8_t *out);
