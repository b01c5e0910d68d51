// pk_code.cpp -- host-side code construction: GF(2^m) tables, generator polynomial,
// packed syndrome columns, Chien row offsets and (for n-k <= 16) the coset table.
// Replaces the set-up part of the reference's main() (src/main.cpp:59-95) and the
// GF(2)[x] helpers it calls (src/bchCoder.cpp:25-226).  Runs once per code; the per-frame
// work is all in the CUDA kernels.
#include "pk_code.h"
#include "pk_kernels.h"

#include <algorithm>
#include <map>
#include <mutex>

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

// ------------------------------------------------------------------ cyclic-class table (see PkClassTable)
namespace {

struct CtBuilder {
    const pk_code &c;
    PkClassTable &ct;
    int m, n, t, Q, per, nsw;
    std::vector<int> js;                       // 3, 5, .., 2t-1
    std::vector<std::vector<uint8_t>> rank;    // per js: value -> rank in its subfield (0xFF: not in the subfield)
    std::vector<int> pos;
    std::vector<uint32_t> S;                   // S_1, S_3, .., S_{2t-1} of the current pattern
    uint64_t posmask;

    CtBuilder(const pk_code &c_, PkClassTable &ct_) : c(c_), ct(ct_), m(c_.m), n(c_.n), t(c_.t), Q(1 << c_.m) {}

    uint32_t gmul(uint32_t a, int e) const {   // a * alpha^e
        return a ? c.alog[(c.log[a] + (uint32_t)(((e % n) + n) % n)) % n] : 0u;
    }
    int coset_rep(int j, int *size) const {
        int e = j % n, rep = e, sz = 0;
        do { rep = std::min(rep, e); e = (2 * e) % n; ++sz; } while (e != j % n);
        if (size) *size = sz;
        return rep;
    }
    bool prepare() {
        if (t < 2) return false;
        std::vector<int> reps(1, coset_rep(1, nullptr));
        std::vector<int> deg;
        for (int j = 3; j < 2 * t; j += 2) {
            int d;
            const int rep = coset_rep(j, &d);
            if (std::find(reps.begin(), reps.end(), rep) != reps.end()) continue;   // S_j is a power of an earlier syndrome
            reps.push_back(rep);
            js.push_back(j);
            deg.push_back(d);
        }
        int kb = 0;
        for (int d : deg) kb += d;
        if (kb + 1 > 28 || t * m + kb + 1 > 64 || (int)js.size() > 8) return false;
        ct.kb = kb;
        ct.js = js;
        // field offsets: S_3 on top (its table carries the S_1 = 0 flag one bit above its rank)
        ct.mult.assign(js.size(), 0);
        int sh = 0;
        for (int i = (int)js.size() - 1; i >= 0; --i) { ct.mult[i] = 1u << sh; sh += deg[i]; }
        rank.resize(js.size());
        for (size_t i = 0; i < js.size(); ++i) {
            const int sub = (1 << deg[i]) - 1, step = n / sub;
            std::vector<uint32_t> el(1, 0u);
            for (int e = 0; e < sub; ++e) el.push_back(c.alog[(e * step) % n]);
            std::sort(el.begin(), el.end());
            rank[i].assign(Q, 0xFF);
            for (size_t r = 0; r < el.size(); ++r) rank[i][el[r]] = (uint8_t)r;
        }
        ct.logt.assign(Q, 0);
        for (int v = 1; v < Q; ++v) ct.logt[v] = (uint8_t)c.log[v];
        ct.norm.assign(js.size() * (size_t)Q * Q, 0);
        for (size_t i = 0; i < js.size(); ++i)
            for (int s1 = 0; s1 < Q; ++s1)
                for (int v = 0; v < Q; ++v) {
                    const uint32_t w = s1 ? gmul((uint32_t)v, -(int)((js[i] * (long)c.log[s1]) % n)) : (uint32_t)v;
                    uint8_t r = rank[i][w];
                    if (r == 0xFF) r = 0;   // not a syndrome value of this code
                    if (!s1 && i == 0) r |= (uint8_t)(1u << deg[0]);
                    ct.norm[(i * Q + v) * Q + s1] = r;
                }
        // per-position columns in "pair" packing: field i of a column = alpha^p | alpha^{js[i] p} << m, two fields per
        // 32-bit word -- the XOR of columns then carries (S_1, S_js[i]) side by side, i.e. the rank-table index itself
        {
            const int per = 32 / m, nsw = (2 * t + per - 1) / per;
            ct.col.assign((size_t)n * nsw, 0u);
            if ((int)(js.size() + 1) / 2 > nsw) return false;
            for (int p = 0; p < n; ++p)
                for (size_t i = 0; i < js.size(); ++i) {
                    const uint32_t pair = c.alog[p % n] | (c.alog[(js[i] * (long)p) % n] << m);
                    ct.col[(size_t)p * nsw + i / 2] |= pair << ((i % 2) * 2 * m);
                }
        }
        // capacity: about V(n,t)/n classes with S_1 != 0 and as many keys again with S_1 = 0
        double vol = 0, binom = 1;
        for (int w = 0; w <= t; ++w) { vol += binom; binom = binom * (n - w) / (w + 1); }
        ct.hbits = 10;
        while ((double)(1ull << ct.hbits) < 3.5 * vol / n) ++ct.hbits;
        ct.hash.assign((size_t)1 << ct.hbits, ~0ull);
        ct.bits.assign(((size_t)1 << (kb + 1)) / 32 + 1, 0u);
        posmask = (t * m == 64) ? ~0ull : ((1ull << (t * m)) - 1);
        return true;
    }
    uint32_t key_of(const std::vector<uint32_t> &Sv) const {
        uint32_t key = 0;
        for (size_t i = 0; i < js.size(); ++i) key += (uint32_t)ct.norm[(i * Q + Sv[i + 1]) * Q + Sv[0]] * ct.mult[i];
        return key;
    }
    void insert(uint32_t key, uint64_t packed) {
        ct.bits[key >> 5] |= 1u << (key & 31);
        const uint32_t mask = (1u << ct.hbits) - 1;
        uint32_t h = (uint32_t)(key * 0x9E3779B1u) >> (32 - ct.hbits);
        for (;;) {
            uint64_t &e = ct.hash[h];
            if (e == ~0ull) { e = ((uint64_t)key << (t * m)) | packed; ++ct.entries; return; }
            if ((uint32_t)(e >> (t * m)) == key && (e & posmask) != posmask) return;   // same class seen before
            h = (h + 1) & mask;
        }
    }
    std::vector<uint32_t> Sv;
    void emit(int w) {
        Sv = S;
        auto pack = [&](int shift) {
            uint64_t e = 0;
            for (int i = 0; i < t; ++i) {
                const uint64_t p = (i < w) ? (uint64_t)((pos[i] + shift) % n) : (uint64_t)n;
                e |= p << (i * m);
            }
            return e;
        };
        if (S[0]) {
            insert(key_of(Sv), pack(n - (int)c.log[S[0]]));
        } else {
            for (int r = 0; r < n; ++r) {
                for (size_t i = 0; i < js.size(); ++i) Sv[i + 1] = gmul(S[i + 1], (int)((js[i] * (long)r) % n));
                insert(key_of(Sv), pack(r));
            }
        }
    }
    void rec(int w) {
        emit(w);
        if (w == t) return;
        for (int p = pos[w - 1] + 1; p < n; ++p) {
            pos[w] = p;
            S[0] ^= c.alog[p % n];
            for (size_t i = 0; i < js.size(); ++i) S[i + 1] ^= c.alog[(js[i] * (long)p) % n];
            rec(w + 1);
            S[0] ^= c.alog[p % n];
            for (size_t i = 0; i < js.size(); ++i) S[i + 1] ^= c.alog[(js[i] * (long)p) % n];
        }
    }
    bool build() {
        if (!prepare()) return false;
        pos.assign(t, 0);
        S.assign(js.size() + 1, 1u);   // the pattern {0}: every S_j = alpha^0
        rec(1);
        return true;
    }
};

std::mutex g_ct_mutex;
std::map<std::pair<int, int>, std::shared_ptr<const PkClassTable>> g_ct_cache;

std::shared_ptr<const PkClassTable> class_table_for(const pk_code &c) {
    std::lock_guard<std::mutex> lock(g_ct_mutex);
    auto it = g_ct_cache.find({c.m, c.t});
    if (it != g_ct_cache.end()) return it->second;
    auto ct = std::make_shared<PkClassTable>();
    CtBuilder b(c, *ct);
    std::shared_ptr<const PkClassTable> out;
    if (b.build()) out = ct;
    g_ct_cache[{c.m, c.t}] = out;
    return out;
}

}  // namespace

bool PkClassTable::lookup(int m, int t, const uint32_t *packed, uint32_t *A) const {
    const int n = (1 << m) - 1, Q = 1 << m, per = 32 / m, nw = (n + 31) / 32;
    auto syn = [&](int j) { return (packed[(j - 1) / per] >> (((j - 1) % per) * m)) & (uint32_t)n; };
    const uint32_t s1 = syn(1);
    uint32_t key = 0;
    for (size_t i = 0; i < mult.size(); ++i) key += (uint32_t)norm[(i * Q + syn(js[i])) * Q + s1] * mult[i];
    for (int w = 0; w < nw; ++w) A[w] = 0;
    if (!((bits[key >> 5] >> (key & 31)) & 1u)) return false;
    const uint64_t posmask = (t * m == 64) ? ~0ull : ((1ull << (t * m)) - 1);
    const uint32_t mask = (1u << hbits) - 1;
    uint32_t h = (uint32_t)(key * 0x9E3779B1u) >> (32 - hbits);
    for (;;) {
        const uint64_t e = hash[h];
        if (e == ~0ull) return false;   // cannot happen: the bitmap is exact
        if ((uint32_t)(e >> (t * m)) == key && (e & posmask) != posmask) {
            const int s = logt[s1];
            for (int i = 0; i < t; ++i) {
                int p = (int)((e >> (i * m)) & (uint64_t)n);
                if (p == n) continue;
                p = (p + s) % n;
                A[p >> 5] |= 1u << (p & 31);
            }
            return true;
        }
        h = (h + 1) & mask;
    }
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

    // ---- cyclic-class table for the codes whose wide search can run on it
    c.use_ct = false;
    if (c.ks && c.ks->has_class && !c.use_lut) {
        c.ct = class_table_for(c);
        c.use_ct = (bool)c.ct;
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
