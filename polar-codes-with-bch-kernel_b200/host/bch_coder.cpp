// bch_coder.cpp -- see bch_coder.hpp.  Own implementation (std::vector based); only the
// signatures and observable behaviour follow the reference (src/bchCoder.cpp).
#include "bch_coder.hpp"

#include <cstring>
#include <iostream>
#include <random>
#include <vector>

#include "pk_capi.h"

namespace {
typedef std::vector<unsigned char> P2;

std::default_random_engine g_engine;                                // unseeded, like src/bchCoder.cpp:20
std::uniform_int_distribution<unsigned short> g_bit(0, 1);

P2 make(const unsigned char *a, int n) { return P2(a, a + n); }
void strip(P2 &a) {
    while (a.size() > 1 && !a.back()) a.pop_back();
}
P2 product(const P2 &a, const P2 &b) {
    P2 r(a.size() + b.size() - 1, 0);
    for (size_t i = 0; i < a.size(); ++i)
        for (size_t j = 0; j < b.size(); ++j) r[i + j] ^= a[i] & b[j];
    return r;
}
void long_division(P2 a, const P2 &b, P2 &quo, P2 &rem) {
    const size_t nq = a.size() >= b.size() ? a.size() - b.size() + 1 : 1;
    quo.assign(nq, 0);
    size_t s = a.size();
    while (s >= b.size()) {
        const size_t sh = s - b.size();
        for (size_t i = 0; i < b.size(); ++i) a[sh + i] ^= b[i];
        quo[sh] = 1;
        while (s > 0 && !a[s - 1]) --s;
    }
    rem.assign(a.begin(), a.begin() + (s ? s : 1));
}
unsigned char *release(const P2 &a) {
    unsigned char *r = new unsigned char[a.size()];
    std::memcpy(r, a.data(), a.size());
    return r;
}
template <class S, class V>
void print_row(S &out, const V *v, int size, bool as_bit) {
    for (int i = 0; i < size; ++i) {
        if (as_bit) out << (v[i] ? 1 : 0) << ' ';
        else out << v[i] << ' ';
    }
    out << std::endl;
}
}  // namespace

void findMinimalPolynomial(int i, int power, const unsigned long *fieldElements, int *size, unsigned char *res) {
    const int n = (1 << power) - 1;
    std::vector<int> lg(n + 1, -1);
    for (int e = 0; e < n; ++e) lg[fieldElements[e]] = e;
    std::vector<unsigned long> p(1, 1);
    int e = i % n;
    do {   // multiply by (x + alpha^e) for every conjugate
        std::vector<unsigned long> nx(p.size() + 1, 0);
        for (size_t d = 0; d < p.size(); ++d) {
            nx[d + 1] ^= p[d];
            if (p[d]) nx[d] ^= fieldElements[(lg[p[d]] + e) % n];
        }
        p.swap(nx);
        e = (2 * e) % n;
    } while (e != i % n);
    *size = (int)p.size();
    for (size_t d = 0; d < p.size(); ++d) res[d] = (p[d] % 2) ? 1 : 0;
}

bool comparePoly(const unsigned char *poly1, int size1, const unsigned char *poly2, int size2) {
    return size1 == size2 && std::memcmp(poly1, poly2, (size_t)size1) == 0;
}

unsigned char *multiplyPolynomials(const unsigned char *first, int size1, const unsigned char *second, int size2,
                                   int *sizeRes) {
    P2 r = product(make(first, size1), make(second, size2));
    if (sizeRes) *sizeRes = (int)r.size();
    return release(r);
}

void multiplyPolynomials(const unsigned char *first, int size1, const unsigned char *second, int size2,
                         unsigned char *res, int *sizeRes) {
    P2 r = product(make(first, size1), make(second, size2));
    std::memcpy(res, r.data(), r.size());
    if (sizeRes) *sizeRes = (int)r.size();
}

unsigned char *dividePolynomial(const unsigned char *first, int size1, const unsigned char *second, int size2,
                                int *size, bool needRemainder) {
    P2 q, r;
    long_division(make(first, size1), make(second, size2), q, r);
    const P2 &o = needRemainder ? r : q;
    *size = (int)o.size();
    return release(o);
}

unsigned char *lcm(const unsigned char *first, int size1, const unsigned char *second, int size2, int *sizeRes) {
    P2 a = make(first, size1), b = make(second, size2);
    while (!(b.size() == 1 && !b[0])) {   // Euclid
        P2 q, r;
        long_division(a, b, q, r);
        a = b;
        b = r;
    }
    strip(a);
    P2 q, r;
    long_division(product(make(first, size1), make(second, size2)), a, q, r);
    *sizeRes = (int)q.size();
    return release(q);
}

unsigned char *generateRandomPoly(long k) {
    unsigned char *r = new unsigned char[k];
    generateRandomPoly(r, k);
    return r;
}
void generateRandomPoly(unsigned char *res, long k) {
    for (long i = 0; i < k; ++i) res[i] = (unsigned char)g_bit(g_engine);
}
void addNoise(double standartDeviation, const unsigned char *codeword, double *wordWithNoise, unsigned long n) {
    std::normal_distribution<double> gauss(0.0, standartDeviation);
    for (unsigned long i = 0; i < n; ++i) wordWithNoise[i] = (codeword[i] ? 1 : -1) + gauss(g_engine);
}

void printVec(const unsigned char *poly, int size) { print_row(std::cout, poly, size, true); }
void printVec(const unsigned long *poly, int size) { print_row(std::cout, poly, size, false); }
void printVec(const double *poly, int size) { print_row(std::cout, poly, size, false); }
void printVec(std::ofstream &out, const unsigned char *poly, int size) { print_row(out, poly, size, true); }
void printVec(std::ofstream &out, const double *poly, int size) { print_row(out, poly, size, false); }

void printMatrix(unsigned char **const matrix, int sizeI, int sizeJ) {
    if (sizeJ == -1) sizeJ = sizeI;
    for (int i = 0; i < sizeI; ++i) print_row(std::cout, matrix[i], sizeJ, true);
    std::cout << std::endl;
}
void printMatrix(std::ofstream &out, unsigned char **const matrix, int sizeI, int sizeJ) {
    if (sizeJ == -1) sizeJ = sizeI;
    for (int i = 0; i < sizeI; ++i) print_row(out, matrix[i], sizeJ, true);
    out << std::endl;
}

void makeMatrix(int power, const unsigned long *fieldElements, unsigned char **matrix) {
    (void)fieldElements;   // the library rebuilds the same field from `power`
    pk_code *code = nullptr;
    if (pk_code_create_host(power, 1, &code) != PK_OK) throw "Invalid values of argument(s)\n";
    const int n = (1 << power) - 1;
    std::vector<unsigned char> flat((size_t)n * n);
    pk_make_kernel_matrix(code, flat.data());
    for (int i = 0; i < n; ++i) std::memcpy(matrix[i], &flat[(size_t)i * n], (size_t)n);
    pk_code_destroy(code);
}
