// pk_code.cpp -- host-side code construction: GF(2^m) tables, generator polynomial,
// packed syndrome columns, Chien row offsets and (for n-k <= 16) the coset table.
// Replaces the set-up part of the reference's main() (src/main.cpp:59-95) and the
// GF(2)[x] helpers it calls (src/bchCoder.cpp:25-226).  Runs once per code; the per-frame
// work is all in the CUDA kernels.
#include "pk_code.h"
#include "pk_kernels.h"

#include <algorithm>

namespace {

typedef std::vector<uint8_t> Poly2;  // GF(2)[x], coefficient i = power i, no trailing zeros (except "0" = {0})

// primitive polynomials, same table as src/main.cpp:14-15 (index m-1)
const uint32_t kPrimitive[16] = {3, 7, 11, 19, 37, 67, 137, 285, 529, 1033, 2053, 4179, 8219, 17475, 32771, 69643};

void trim(Poly2 &a) {
    while (a.size() > 1 && !a.back()) a.pop_back();
}
Poly2 mul2(const Poly2 &a, const Poly2 &b) {
    Poly2 r(a.size() + b.size() - 1, 0);
    for (size_t i = 0; i < a.size(); ++i)
        if (a[i])
            for (size_t j = 0; j < b.size(); ++j) r[i + j] ^= b[j];
    trim(r);
    return r;
}
// a = q*b + r
void divmod2(Poly2 a, const Poly2 &b, Poly2 &q, Poly2 &r) {
    q.assign(a.size() >= b.size() ? a.size() - b.size() + 1 : 1, 0);
    trim(a);
    while (a.size() >= b.size() && !(a.size() == 1 && !a[0])) {
        size_t sh = a.size() - b.size();
        q[sh] = 1;
        for (size_t i = 0; i < b.size(); ++i) a[sh + i] ^= b[i];
        trim(a);
    }
    r = a;
}
bool is_zero(const Poly2 &a) { return a.size() == 1 && !a[0]; }
Poly2 gcd2(Poly2 a, Poly2 b) {
    while (!is_zero(b)) {
        Poly2 q, r;
        divmod2(a, b, q, r);
        a = b;
        b = r;
    }
    return a;
}

}  // namespace

// minimal polynomial of alpha^i over GF(2): prod over the cyclotomic coset of i
static Poly2 minimal_poly(const pk_code &c, int i) {
    const int n = c.n;
    std::vector<uint32_t> p(1, 1);  // GF(2^m)[x]
    int e = i % n;
    do {
        uint32_t root = c.alog[e];
        std::vector<uint32_t> nx(p.size() + 1, 0);
        for (size_t d = 0; d < p.size(); ++d) {
            nx[d + 1] ^= p[d];
            if (p[d]) nx[d] ^= c.alog[(c.log[p[d]] + c.log[root]) % n];
        }
        p.swap(nx);
        e = (2 * e) % n;
    } while (e != i % n);
    Poly2 out(p.size());
    for (size_t d = 0; d < p.size(); ++d) out[d] = (uint8_t)(p[d] & 1);
    return out;
}

std::string pk_code_build_host(pk_code &c, int m, int t) {
    // same validity rule as src/main.cpp:55 (`1 << power - 1` parses as 1 << (power-1))
    if (m < 3 || m > 8) return "Invalid values of arguments (m must be in [3,8])";
    if (t <= 0 || t >= (1 << (m - 1))) return "Invalid values of arguments";
    c.m = m;
    c.t = t;
    c.n = (1 << m) - 1;
    const int n = c.n;

    // ---- field tables (main.cpp:59-78)
    c.alog.assign(n, 0);
    c.log.assign(n + 1, 0);
    c.log[0] = 0xFFFFFFFFu;
    uint32_t v = 1;
    for (int i = 0; i < n; ++i) {
        c.alog[i] = v;
        c.log[v] = (uint32_t)i;
        v <<= 1;
        if (v >> m) v ^= kPrimitive[m - 1];
    }

    // ---- generator g(x) = lcm{ M_i(x) : 1 <= i <= 2t-1 } (main.cpp:80-95)
    Poly2 g = minimal_poly(c, 1);
    for (int i = 2; i < 2 * t; ++i) {
        Poly2 mp = minimal_poly(c, i);
        Poly2 q, r, prod = mul2(g, mp);
        divmod2(prod, gcd2(g, mp), q, r);
        trim(q);
        g = q;
    }
    c.g = g;
    c.gsize = (int)g.size();
    c.k = n - c.gsize + 1;
    c.nk = n - c.k;
    if (c.k <= 0) return "Invalid values of arguments (k <= 0)";

    c.ks = pk_find_kernels(m, t);

    // ---- GF product table and Chien row offsets
    const int q = 1 << m;
    c.mul.assign((size_t)q * q, 0);
    for (int a = 1; a < q; ++a)
        for (int b = 1; b < q; ++b) c.mul[(size_t)a * q + b] = (uint8_t)c.alog[(c.log[a] + c.log[b]) % n];
    c.xoff.assign(n, 0);
    for (int p = 0; p < n; ++p) c.xoff[p] = (uint16_t)(c.alog[(n - p) % n] << m);

    // ---- packed syndrome columns: position p contributes S_j += alpha^{j p}, j = 1..2t
    const int per = 32 / m;
    const int nsw = (2 * t + per - 1) / per;
    c.hcol.assign((size_t)n * nsw, 0);
    for (int p = 0; p < n; ++p)
        for (int j = 1; j <= 2 * t; ++j) {
            uint32_t s = c.alog[((long)j * p) % n];
            c.hcol[(size_t)p * nsw + (j - 1) / per] |= s << (((j - 1) % per) * m);
        }

    // ---- g as a bit mask; x^p mod g (coset index of a single error at p)
    const int nw = (n + 31) / 32;
    c.gmask.assign(nw, 0);
    for (int i = 0; i < c.gsize; ++i)
        if (g[i]) c.gmask[i >> 5] |= 1u << (i & 31);
    c.rcol.assign(n, 0);
    c.use_lut = false;
    if (c.nk <= 16) {
        Poly2 xp(1, 1);
        for (int p = 0; p < n; ++p) {
            Poly2 qq, r;
            divmod2(xp, g, qq, r);
            uint32_t bits = 0;
            for (size_t i = 0; i < r.size(); ++i) bits |= (uint32_t)r[i] << i;
            c.rcol[p] = bits;
            xp.insert(xp.begin(), 0);  // * x
        }
    }

    // ---- coset table: the verdict and error positions of the algebraic decoder for every
    // syndrome, computed with the same pk_alg_decode<M,T> the kernels run.
    if (c.ks && c.nk <= 16 && t * m <= 15) {
        c.use_lut = true;
        const uint32_t ncos = 1u << c.nk;
        c.lut.assign(ncos, 0xFFFF);
        std::vector<uint32_t> sw(nsw), A(nw);
        for (uint32_t r = 0; r < ncos; ++r) {
            std::fill(sw.begin(), sw.end(), 0u);
            for (int p = 0; p < c.nk; ++p)   // the word r(x) itself lies in coset r
                if ((r >> p) & 1)
                    for (int w = 0; w < nsw; ++w) sw[w] ^= c.hcol[(size_t)p * nsw + w];
            bool ok = c.ks->host_alg_decode(sw.data(), c.mul.data(), c.xoff.data(), A.data());
            if (!ok) continue;
            uint32_t e = 0;
            int cnt = 0;
            for (int p = 0; p < n; ++p)
                if ((A[p >> 5] >> (p & 31)) & 1) e |= (uint32_t)p << (m * cnt++);
            for (; cnt < t; ++cnt) e |= (uint32_t)n << (m * cnt);  // "no position" = n (all ones)
            c.lut[r] = (uint16_t)e;
        }
    }
    return "";
}

// makeMatrix (src/bchCoder.cpp:317-345): rows are nested BCH generator polynomials and
// their shifts; row 0 = 1.
void pk_code_kernel_matrix(const pk_code &c, uint8_t *out) {
    const int len = c.n;
    const int amount = ((1 << c.m) - 2) / 2;
    std::fill(out, out + (size_t)len * len, 0);
    out[0] = 1;
    Poly2 g(1, 1);
    for (int i = 2; i <= amount; ++i) {
        Poly2 mp = minimal_poly(c, i);
        if (g.size() >= mp.size()) {
            Poly2 q, r;
            divmod2(g, mp, q, r);
            if (is_zero(r)) continue;
        }
        Poly2 ng = mul2(mp, g);
        const int go = (int)g.size(), gn = (int)ng.size();
        for (int j = 0; j < gn; ++j) out[(size_t)(gn - 1) * len + j] = ng[j];
        for (int row = go, sh = 1; row < gn - 1; ++row, ++sh)
            for (int q = 0; q < go; ++q) out[(size_t)row * len + sh + q] = g[q];
        g = ng;
    }
}
