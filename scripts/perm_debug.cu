// debug harness: device vs host evaluation of permutation-search candidates (not part of the product)
#include <cstdio>
#include <string>
int pk_set_error(int code, const std::string &msg) { fprintf(stderr, "err %d %s\n", code, msg.c_str()); return code; }
unsigned long long g_pk_launches = 0;
#include "../polar-codes-with-bch-kernel_b200/csrc/pk_perm.cu"

struct Dbg { unsigned int basis[8]; unsigned long long cand[64]; unsigned long long cost; int mb; unsigned long long base[4]; };
__global__ void k_dbg(int power, const unsigned long long *rows, unsigned long long seed, unsigned long long trial, Dbg *o) {
    const int n = 1 << power;
    unsigned long long base[64], cand[64];
    for (int r = 0; r < n; ++r) base[r] = rows[r];
    unsigned int basis[8];
    candidate_basis(power, seed, trial, basis);
    permute_rows(power, base, basis, cand);
    int mb = 0;
    o->cost = trellis_cost(cand, n, &mb);
    o->mb = mb;
    for (int k = 0; k < power; ++k) o->basis[k] = basis[k];
    for (int r = 0; r < n; ++r) o->cand[r] = cand[r];
    for (int r = 0; r < 4; ++r) o->base[r] = base[r];
}
__global__ void k_cost_only(const unsigned long long *cand, int n, unsigned long long *o) { o[0] = trellis_cost(cand, n, nullptr); }
int main() {
    const int power = 4, n = 16;
    unsigned long long rows[64] = {0x1ull,0x3ull,0x5ull,0x9ull,0x11ull,0x27ull,0x4dull,0x99ull,0x131ull,0x3a3ull,0x745ull,0xa6full,0x14ddull,0x29b9ull,0x5371ull,0xffffull};
    unsigned long long *d_rows, *d_best, *d_costs; Dbg *d_o, h;
    cudaMalloc(&d_rows, 8 * 64); cudaMalloc(&d_o, sizeof(Dbg)); cudaMalloc(&d_best, 8); cudaMalloc(&d_costs, 8 * 512);
    cudaMemcpy(d_rows, rows, 8 * 64, cudaMemcpyHostToDevice);
    k_perm_search<<<4, 128>>>(power, d_rows, 5ull, 0ull, 500, 22, d_best, d_costs);
    unsigned long long hc[512];
    cudaMemcpy(hc, d_costs, 8 * 500, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (unsigned long long trial = 0; trial < 500; ++trial) {
        unsigned int basis[8]; unsigned long long cand[64]; int mb = 0;
        candidate_basis(power, 5ull, trial, basis);
        permute_rows(power, rows, basis, cand);
        unsigned long long cost = trellis_cost(cand, n, &mb);
        if (cost != hc[trial]) {
            if (bad++ < 3) {
                k_dbg<<<1, 1>>>(power, d_rows, 5ull, trial, d_o);
                cudaMemcpy(&h, d_o, sizeof(Dbg), cudaMemcpyDeviceToHost);
                int diff = 0; for (int r = 0; r < n; ++r) diff += cand[r] != h.cand[r];
                unsigned long long *d_c, co;
                cudaMalloc(&d_c, 8 * 64); cudaMemcpy(d_c, cand, 8 * 64, cudaMemcpyHostToDevice);
                k_cost_only<<<1, 1>>>(d_c, n, d_best); cudaMemcpy(&co, d_best, 8, cudaMemcpyDeviceToHost);
                for (int r = 0; r < n; ++r) if (cand[r] != h.cand[r]) printf("   row %d: in %llx host %llx dev %llx\n", r, rows[r], cand[r], h.cand[r]);
                printf("   base on device after: %llx %llx %llx %llx\n", h.base[0], h.base[1], h.base[2], h.base[3]);
                printf("trial %llu: host %llu search kernel %llu single-thread kernel %llu cost-only kernel on host cand %llu; cand rows differing %d; basis host %x %x %x %x dev %x %x %x %x\n",
                       trial, cost, hc[trial], h.cost, co, diff, basis[0], basis[1], basis[2], basis[3], h.basis[0], h.basis[1], h.basis[2], h.basis[3]);
            }
        }
    }
    printf("mismatching candidates: %d of 500 (%s)\n", bad, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
