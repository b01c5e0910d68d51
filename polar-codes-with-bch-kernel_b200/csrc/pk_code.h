// pk_code.h -- host-side description of one primitive narrow-sense BCH code and the
// lookup tables the sm_100a kernels stage into shared memory.
#pragma once
#include <stdint.h>
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
};

struct pk_code {
    int m = 0, n = 0, k = 0, t = 0, gsize = 0, nk = 0;
    int device = 0;
    bool use_lut = false;
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
