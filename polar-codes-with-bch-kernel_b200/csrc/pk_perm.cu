// pk_perm.cu -- kernel construction tooling of the reference's newer bchCoder.cpp:
//   swapColumns        (root bchCoder.cpp:478-528)  columns re-ordered by field-element value
//   randomSwapColumns  (:541-699)                   2*10^7 random GL(m,2) column permutations, keep the cheapest
//   randomInvertibleMatrix (:766-784)               B = L U, unit lower times unit upper triangular
// The reference scores a candidate with the operation counters of a `SectionedTrellisKernelProcessor` that is not in
// its tree (bchCoder.cpp:13 does not compile).  Here the score is the work of the trellis kernel processor that IS there
// (out/external/TrellisKernelProcessor.cpp:234-295): the branch evaluations of one pass over all phases,
//     cost = sum_phase sum_{section j} 2 * 2^{s_j(phase)},
// s_j = state bits before section j of the minimal trellis of rows phase..l-1 with row `phase` tagged (SURVEY 8a table:
// 11 712 for the 16 x 16 kernel in natural column order).  s_j needs no trellis: with p_j = rank of the first j columns
// and f_j = rank of the columns j.. (tag included), s_j = p_j + f_j - rows; both ranks come from two echelon bases that
// grow by one row per phase, so a candidate costs O(l^2) word operations -- one GPU thread per candidate.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pk_capi.h"
#include "pk_philox.cuh"

int pk_set_error(int code, const std::string &msg);   // pk_capi.cu
extern unsigned long long g_pk_launches;

namespace {

// cost and maximal state bits of a kernel given as row masks (bit c = column c), l <= 64
__host__ __device__ inline unsigned long long trellis_cost(const unsigned long long *rows, int l, int *max_bits) {
    unsigned long long lead[64], trail[64];   // echelon bases indexed by leading (lowest) / trailing (highest) column
    for (int i = 0; i < l; ++i) { lead[i] = 0; trail[i] = 0; }
    unsigned long long leadmask = 0, trailmask = 0, cost = 0;
    int mx = 0;
    for (int ph = l - 1; ph >= 0; --ph) {
        // rows ph..l-1: leading columns (row ph joins the left basis now, the right basis after this phase)
        unsigned long long v = rows[ph];
        while (v) {
#ifdef __CUDA_ARCH__
            const int c = __ffsll((long long)v) - 1;
#else
            const int c = __builtin_ctzll(v);
#endif
            if (lead[c]) v ^= lead[c];
            else { lead[c] = v; leadmask |= 1ull << c; break; }
        }
        const int nrows = l - ph;
        for (int j = 0; j < l; ++j) {
            const unsigned long long below = (j == 0) ? 0ull : (~0ull >> (64 - j));
#ifdef __CUDA_ARCH__
            const int p = __popcll(leadmask & below), f = __popcll(trailmask & ~below) + 1;
#else
            const int p = __builtin_popcountll(leadmask & below), f = __builtin_popcountll(trailmask & ~below) + 1;
#endif
            const int s = p + f - nrows;   // (the tagged row is independent of the others on every column suffix)
            cost += 2ull << s;
            mx = s > mx ? s : mx;
        }
        v = rows[ph];
        while (v) {
#ifdef __CUDA_ARCH__
            const int c = 63 - __clzll((long long)v);
#else
            const int c = 63 - __builtin_clzll(v);
#endif
            if (trail[c]) v ^= trail[c];
            else { trail[c] = v; trailmask |= 1ull << c; break; }
        }
    }
    if (max_bits) *max_bits = mx;
    return cost;
}

// candidate `trial`: B = L U from Philox (randomInvertibleMatrix), column j of the new kernel = column B j of the old one
__host__ __device__ inline void candidate_basis(int power, unsigned long long seed, unsigned long long trial, unsigned int *basis) {
    unsigned int bits[4];
#ifdef __CUDA_ARCH__
    const PkPhilox r = pk_philox((uint32_t)trial, (uint32_t)(trial >> 32), 0x5EA7C4u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    bits[0] = r.c[0]; bits[1] = r.c[1]; bits[2] = r.c[2]; bits[3] = r.c[3];
#else
    // host mirror of pk_philox (same rounds)
    uint32_t c0 = (uint32_t)trial, c1 = (uint32_t)(trial >> 32), c2 = 0x5EA7C4u, c3 = 0u, k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    bits[0] = c0; bits[1] = c1; bits[2] = c2; bits[3] = c3;
#endif
    // row i of L (unit lower triangular) and column j of U (unit upper triangular) as bit masks over the inner index
    unsigned int Lrow[8], Ucol[8];
    for (int i = 0; i < 8; ++i) { Lrow[i] = 1u << i; Ucol[i] = 1u << i; }
    int q = 0;
    for (int i = 0; i < power; ++i) {
        for (int j = 0; j < i; ++j) { Lrow[i] |= ((bits[q >> 5] >> (q & 31)) & 1u) << j; ++q; }
        for (int j = i + 1; j < power; ++j) { Ucol[j] |= ((bits[q >> 5] >> (q & 31)) & 1u) << i; ++q; }
    }
    for (int j = 0; j < power; ++j) {
        unsigned int col = 0;   // newBasis[j] = sum_k e_k b[k][j], b = L U
        for (int k = 0; k < power; ++k) {
            unsigned int x = Lrow[k] & Ucol[j];
            x ^= x >> 4; x ^= x >> 2; x ^= x >> 1;
            col |= (x & 1u) << k;
        }
        basis[j] = col;
    }
}
__host__ __device__ inline void permute_rows(int power, const unsigned long long *rows, const unsigned int *basis, unsigned long long *out) {
    const int n = 1 << power;
    unsigned char src[64];   // column j of the new kernel is column src[j] = B j of the old one
    for (int j = 0; j < n; ++j) {
        unsigned int s = 0;
        for (int k = 0; k < power; ++k)
            if ((j >> k) & 1) s ^= basis[k];
        src[j] = (unsigned char)(s & (unsigned int)(n - 1));
    }
    for (int r = 0; r < n; ++r) {
        const unsigned long long w = rows[r];
        unsigned long long o = 0;
        for (int j = 0; j < n; ++j) o |= ((w >> src[j]) & 1ull) << j;
        out[r] = o;
    }
}

__global__ void __launch_bounds__(128)
k_perm_search(int power, const unsigned long long *rows, unsigned long long seed, unsigned long long first, long ntrials, int max_bits_allowed,
              unsigned long long *best, unsigned long long *costs_out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntrials) return;
    const int n = 1 << power;
    unsigned long long cand[64];
    unsigned int basis[8];
    candidate_basis(power, seed, first + (unsigned long long)i, basis);
    permute_rows(power, rows, basis, cand);
    int mb = 0;
    const unsigned long long cost = trellis_cost(cand, n, &mb);
    if (costs_out) costs_out[i] = cost;
    if (mb > max_bits_allowed) return;
    atomicMin(best, (cost << 26) | ((first + (unsigned long long)i) & 0x3FFFFFFull));
}

void rows_from_matrix(int n, const uint8_t *matrix, std::vector<unsigned long long> &rows) {
    rows.assign(n, 0);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c)
            if (matrix[(size_t)r * n + c]) rows[r] |= 1ull << c;
}

}  // namespace

extern "C" {

int pk_kernel_trellis_cost(int size, const uint8_t *matrix, uint64_t *branches, int *max_state_bits) {
    if (size < 2 || size > 64 || !matrix) return pk_set_error(PK_ERR_ARG, "kernel size must be in [2,64]");
    std::vector<unsigned long long> rows;
    rows_from_matrix(size, matrix, rows);
    int mb = 0;
    const unsigned long long c = trellis_cost(rows.data(), size, &mb);
    if (branches) *branches = c;
    if (max_state_bits) *max_state_bits = mb;
    return PK_OK;
}

// swapColumns (root bchCoder.cpp:478-528) without its scoring tail: columns 0..2 stay, column i (3 <= i < n) becomes
// column j+1 of the input where fieldElements[j] == i (j searched from 2 upwards, :486-495).
int pk_kernel_swap_columns(int power, const uint64_t *field_elements, const uint8_t *matrix, uint8_t *out) {
    if (power < 2 || power > 6 || !field_elements || !matrix || !out) return pk_set_error(PK_ERR_ARG, "bad arguments");
    const int n = 1 << power;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= 2 && j < n; ++j) out[i * n + j] = matrix[i * n + j];
    for (int i = 3; i < n; ++i) {
        int j = 2;
        for (; j < n; ++j)
            if (field_elements[j] == (uint64_t)i) break;
        if (j + 1 >= n) return pk_set_error(PK_ERR_ARG, "field element table does not name column " + std::to_string(i));
        for (int k = 0; k < n; ++k) out[k * n + i] = matrix[k * n + j + 1];
    }
    return PK_OK;
}

// one candidate of the search, on the host (what the device evaluated for `trial`)
int pk_kernel_permute_columns(int power, const uint8_t *matrix, uint64_t seed, uint64_t trial, uint8_t *out, uint32_t *basis_out) {
    if (power < 2 || power > 6 || !matrix || !out) return pk_set_error(PK_ERR_ARG, "bad arguments");
    const int n = 1 << power;
    std::vector<unsigned long long> rows, perm(n);
    rows_from_matrix(n, matrix, rows);
    unsigned int basis[8];
    candidate_basis(power, seed, trial, basis);
    permute_rows(power, rows.data(), basis, perm.data());
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) out[(size_t)r * n + c] = (uint8_t)((perm[r] >> c) & 1ull);
    if (basis_out)
        for (int k = 0; k < power; ++k) basis_out[k] = basis[k];
    return PK_OK;
}

// randomSwapColumns (:541-699) on the GPU: `ntrials` (<= 2^26, the reference runs 2*10^7) random linear column
// permutations j -> B j, B = L U in GL(m,2); returns the cheapest kernel, its basis (newBasis of :604-609), its cost, the
// trial that produced it and the cost of the input.  Candidates whose trellis would exceed max_state_bits (<= 22, the
// default) are discarded.
int pk_kernel_random_search(int power, const uint8_t *matrix, long ntrials, uint64_t seed, int device, int max_state_bits,
                            uint8_t *best_matrix, uint32_t *best_basis, uint64_t *best_cost, uint64_t *best_trial, uint64_t *input_cost,
                            uint64_t *all_costs /*[ntrials] or NULL: the cost of every candidate (tests)*/) {
    if (power < 2 || power > 6 || !matrix || ntrials < 1 || ntrials > (1L << 26)) return pk_set_error(PK_ERR_ARG, "bad arguments (power in [2,6], 1 <= ntrials <= 2^26)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return pk_set_error(PK_ERR_CUDA, "no CUDA device: libpkb200 has no CPU path");
    if (device < 0 || device >= ndev) return pk_set_error(PK_ERR_ARG, "bad device ordinal");
    const int n = 1 << power;
    std::vector<unsigned long long> rows;
    rows_from_matrix(n, matrix, rows);
    if (input_cost) *input_cost = trellis_cost(rows.data(), n, nullptr);
    unsigned long long *d_rows = nullptr, *d_best = nullptr, *d_costs = nullptr, h_best = ~0ull;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&d_rows, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_best, 8);
    if (e == cudaSuccess) e = cudaMemcpy(d_rows, rows.data(), (size_t)n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_best, &h_best, 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && all_costs) e = cudaMalloc(&d_costs, (size_t)ntrials * 8);
    if (e == cudaSuccess) {
        k_perm_search<<<(unsigned)((ntrials + 127) / 128), 128>>>(power, d_rows, seed, 0ull, ntrials, (max_state_bits > 0 && max_state_bits <= 22) ? max_state_bits : 22, d_best, d_costs);
        ++g_pk_launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(&h_best, d_best, 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && all_costs) e = cudaMemcpy(all_costs, d_costs, (size_t)ntrials * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_rows);
    cudaFree(d_best);
    cudaFree(d_costs);
    if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, std::string("pk_kernel_random_search: ") + cudaGetErrorString(e));
    if (h_best == ~0ull) return pk_set_error(PK_ERR_UNSUPPORTED, "no candidate within the state-bit limit");
    const uint64_t trial = h_best & 0x3FFFFFFull;
    if (best_cost) *best_cost = h_best >> 26;
    if (best_trial) *best_trial = trial;
    if (best_matrix) return pk_kernel_permute_columns(power, matrix, seed, trial, best_matrix, best_basis);
    return PK_OK;
}

}  // extern "C"
