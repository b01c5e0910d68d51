// pk_polar_host.cpp -- host-side set-up of mixed-kernel polar codes: specification parser, kernel files and the
// per-phase kernel trellises.  Runs once per code; decoding is all in pk_polar.cu.
#include "pk_polar.h"

#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>

#include "pk_code.h"

namespace {

typedef unsigned long long Row;   // kernel sizes < 64 (TrellisKernelProcessor.cpp:71-72): l+1 columns fit a word

int first_col(Row r) { return __builtin_ctzll(r); }
int last_col(Row r) { return 63 - __builtin_clzll(r); }

// Trellis-oriented (minimum-span) form: all span starts distinct and all span ends distinct.
void minimum_span_form(std::vector<Row> &g) {
    for (bool changed = true; changed;) {
        changed = false;
        for (size_t i = 0; i < g.size() && !changed; ++i)
            for (size_t k = i + 1; k < g.size() && !changed; ++k) {
                if (first_col(g[i]) == first_col(g[k])) {
                    // shorten the row that ends later (ties: either) -- its start moves right
                    if (last_col(g[i]) >= last_col(g[k])) g[i] ^= g[k]; else g[k] ^= g[i];
                    changed = true;
                } else if (last_col(g[i]) == last_col(g[k])) {
                    // shorten the row that starts earlier -- its end moves left
                    if (first_col(g[i]) <= first_col(g[k])) g[i] ^= g[k]; else g[k] ^= g[i];
                    changed = true;
                }
            }
    }
}

}  // namespace

std::string pk_polar_build_trellis(PkKernelTrellis &k) {
    const int l = k.size;
    if (l < 2 || l >= 64) return "Kernel is too big for the trellis processor";   // TrellisKernelProcessor.cpp:71
    k.ab.assign((size_t)l * (l + 1), 0);
    k.off.assign((size_t)l * l, 0);
    k.pred.clear();
    k.max_ab = 0;
    k.ip_x.clear();
    k.ip_sec.assign((size_t)l * (l + 1) * 2, 0);
    k.ip_bits = 0;
    k.ip_ok = true;
    for (int p = 0; p < l; ++p) {
        // extended generator: rows p..l-1, row p tagged in column l (TrellisKernelProcessor.cpp:88-98)
        std::vector<Row> g;
        for (int r = p; r < l; ++r) {
            Row row = 0;
            for (int c = 0; c < l; ++c)
                if (k.matrix[(size_t)r * l + c]) row |= 1ull << c;
            if (r == p) row |= 1ull << l;
            if (!row) return "Matrix is not full rank";
            g.push_back(row);
        }
        // full rank check + MSGM
        {
            std::vector<Row> t = g;
            size_t rank = 0;
            for (int c = 0; c <= l && rank < t.size(); ++c) {
                size_t piv = rank;
                while (piv < t.size() && !((t[piv] >> c) & 1)) ++piv;
                if (piv == t.size()) continue;
                std::swap(t[rank], t[piv]);
                for (size_t i = 0; i < t.size(); ++i)
                    if (i != rank && ((t[i] >> c) & 1)) t[i] ^= t[rank];
                ++rank;
            }
            if (rank != g.size()) return "Matrix is not full rank";
        }
        minimum_span_form(g);
        // walk the sections; state bit q <-> coefficient of active[q]
        std::vector<int> active;
        for (int j = 0; j < l; ++j) {
            int starting = -1, ending = -1;
            for (size_t i = 0; i < g.size(); ++i) {
                if (first_col(g[i]) == j) starting = (int)i;
                if (last_col(g[i]) == j) ending = (int)i;
            }
            std::vector<int> cur = active;            // rows that may be non-zero in column j
            if (starting >= 0) cur.push_back(starting);
            std::vector<int> next;
            for (int r : cur)
                if (r != ending) next.push_back(r);
            const int nb_cur = (int)cur.size(), nb_next = (int)next.size();
            // a predecessor half is (state | branch bit << 15) with 0xFFFF = "no branch": 15 state bits would let state
            // 0x7FFF with branch bit 1 collide with that marker, and the decoder's DUMMY state 1 << max_ab with bit 15
            if (nb_next > 14) return "Kernel trellis has too many states";
            k.off[(size_t)p * l + j] = (uint32_t)k.pred.size();
            k.pred.resize(k.pred.size() + ((size_t)1 << nb_next), 0xFFFFFFFFu);
            uint32_t *tab = &k.pred[k.off[(size_t)p * l + j]];
            // enumerate (previous state over `active`, coefficient of the starting row)
            for (uint32_t sc = 0; sc < (1u << nb_cur); ++sc) {
                uint32_t bit = 0, s1 = 0;
                int q1 = 0;
                for (int q = 0; q < nb_cur; ++q) {
                    const uint32_t coef = (sc >> q) & 1u;
                    bit ^= coef & (uint32_t)((g[cur[q]] >> j) & 1);
                    if (cur[q] != ending) s1 |= coef << q1++;
                }
                const uint32_t s0 = sc & ((1u << active.size()) - 1u);   // previous state (starting row excluded)
                const uint32_t half = s0 | (bit << 15);
                uint32_t &e = tab[s1];
                if ((e & 0xFFFFu) == 0xFFFFu) e = (e & 0xFFFF0000u) | half;
                else if ((e >> 16) == 0xFFFFu) e = (e & 0x0000FFFFu) | (half << 16);
                else return "internal error: more than two branches into a trellis state";
            }
            active = next;
            k.ab[(size_t)p * (l + 1) + j + 1] = (uint8_t)nb_next;
            k.max_ab = std::max(k.max_ab, nb_next);
        }
        if (active.size() != 1 || last_col(g[active[0]]) != l) return "internal error: tagged row is not the last active one";
        // ---- in-place numbering of the same trellis
        {
            std::vector<int> pos(g.size(), -1);   // bit position of every row while it is active
            std::vector<int> act;                 // active rows
            uint32_t used = 0;                    // positions in use
            for (int j = 0; j < l; ++j) {
                int starting = -1, ending = -1;
                for (size_t i = 0; i < g.size(); ++i) {
                    if (first_col(g[i]) == j) starting = (int)i;
                    if (last_col(g[i]) == j) ending = (int)i;
                }
                uint32_t *sw = &k.ip_sec[((size_t)p * (l + 1) + j) * 2];
                sw[0] = (uint32_t)k.ip_x.size();
                if (starting >= 0 && starting == ending) { sw[1] = 0; continue; }   // min(M + c0, M + c1) = M: nothing to do
                int type = 0, q = 0;
                if (starting >= 0 && ending >= 0) { type = 3; q = pos[ending]; }
                else if (ending >= 0) { type = 2; q = pos[ending]; }
                else if (starting >= 0) { type = 1; q = 0; while ((used >> q) & 1u) ++q; }
                if (q > 14) return "Kernel trellis has too many states";
                // states over the active rows except the ending one
                std::vector<int> free_rows;
                for (int r : act)
                    if (r != ending) free_rows.push_back(r);
                size_t n = 0;
                for (uint32_t sc = 0; sc < (1u << free_rows.size()); ++sc) {
                    uint32_t x = 0, t = 0;
                    for (size_t b = 0; b < free_rows.size(); ++b)
                        if ((sc >> b) & 1u) { x |= 1u << pos[free_rows[b]]; t ^= (uint32_t)((g[free_rows[b]] >> j) & 1); }
                    k.ip_x.push_back((uint16_t)(x | (t << 15)));
                    ++n;
                }
                while (n % 4) { k.ip_x.push_back(0xFFFFu); ++n; }
                if (n / 4 > 255) k.ip_ok = false;
                sw[0] |= (uint32_t)std::min<size_t>(n / 4, 255) << 24;
                sw[1] = (uint32_t)q | ((uint32_t)type << 8);
                if (ending >= 0) { used &= ~(1u << pos[ending]); act.erase(std::find(act.begin(), act.end(), ending)); pos[ending] = -1; }
                if (starting >= 0) { pos[starting] = q; used |= 1u << q; act.push_back(starting); k.ip_bits = std::max(k.ip_bits, q + 1); }
            }
            if (act.size() != 1) return "internal error: tagged row is not the last active one";
            k.ip_sec[((size_t)p * (l + 1) + l) * 2 + 1] = (uint32_t)pos[act[0]];
        }
    }
    return "";
}

// Host check of the two table forms against each other: the gather-form recursion (k_polar_decode, pinned to the
// reference) and the in-place recursion (k_polar_lanes) on integer costs -- one label of every section costs nothing,
// the other `ay`, as in the kernels -- for `ntests` random cost vectors and every phase.  Returns "" or what differs.
std::string pk_polar_check_inplace(const PkKernelTrellis &k, unsigned long long seed, int ntests) {
    const int l = k.size;
    if (!k.ip_ok) return "";
    const long long INF = 1ll << 60;
    unsigned long long s = seed * 6364136223846793005ull + 1442695040888963407ull;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(s >> 33); };
    for (int test = 0; test < ntests; ++test) {
        std::vector<long long> ay(l);
        std::vector<int> hd(l);
        for (int j = 0; j < l; ++j) { ay[j] = rnd() % 1000; hd[j] = rnd() & 1; }
        for (int p = 0; p < l; ++p) {
            // gather form
            std::vector<long long> m0(1, 0), m1;
            for (int j = 0; j < l; ++j) {
                const int ns = 1 << k.ab[(size_t)p * (l + 1) + j + 1];
                m1.assign(ns, INF);
                const uint32_t *tab = &k.pred[k.off[(size_t)p * l + j]];
                for (int s1 = 0; s1 < ns; ++s1)
                    for (int hf = 0; hf < 2; ++hf) {
                        const uint32_t v = (tab[s1] >> (16 * hf)) & 0xFFFFu;
                        if (v == 0xFFFFu) continue;
                        const long long c = (((v >> 15) & 1u) ^ (uint32_t)hd[j]) ? ay[j] : 0;
                        m1[s1] = std::min(m1[s1], m0[v & 0x7FFFu] + c);
                    }
                m0.swap(m1);
            }
            const long long want = m0[1] - m0[0];
            // in-place form
            std::vector<long long> M((size_t)1 << std::max(1, k.ip_bits), INF);
            M[0] = 0;
            for (int j = 0; j < l; ++j) {
                const uint32_t *sw = &k.ip_sec[((size_t)p * (l + 1) + j) * 2];
                const uint32_t first = sw[0] & 0xFFFFFFu, ngroups = sw[0] >> 24, q = sw[1] & 0xFFu, type = sw[1] >> 8;
                const size_t Q = (size_t)1 << q;
                for (uint32_t e = 0; e < 4 * ngroups; ++e) {
                    const uint32_t v = k.ip_x[first + e];
                    if (v == 0xFFFFu) continue;
                    const size_t x = v & 0x7FFFu;
                    const bool swp = ((v >> 15) ^ (uint32_t)hd[j]) & 1u;
                    const long long a = M[x], b = M[x | Q];
                    if (type == 0) M[x] = swp ? a + ay[j] : a;
                    else if (type == 1) { M[x] = swp ? a + ay[j] : a; M[x | Q] = swp ? a : a + ay[j]; }
                    else if (type == 2) M[x] = swp ? std::min(a + ay[j], b) : std::min(a, b + ay[j]);
                    else { const long long r0 = std::min(a, b + ay[j]), r1 = std::min(a + ay[j], b); M[x] = swp ? r1 : r0; M[x | Q] = swp ? r0 : r1; }
                }
            }
            const size_t tq = (size_t)1 << (k.ip_sec[((size_t)p * (l + 1) + l) * 2 + 1] & 0xFFu);
            const long long got = M[tq] - M[0];
            if (got != want) return "in-place trellis of phase " + std::to_string(p) + " gives " + std::to_string(got) + ", gather form " + std::to_string(want);
        }
    }
    return "";
}

void pk_polar_ebch_kernel(int m, std::vector<uint8_t> &out) {
    // [[1 0..0],[1^T | K]] with K the (2^m - 1) x (2^m - 1) nested-BCH matrix (root bchCoder.cpp:356-389 restated:
    // first column all ones, matrix[len+2] = 1, the polynomial rows shifted right by one column)
    pk_code c;
    pk_code_build_host(c, m, 1);
    const int n = (1 << m) - 1, l = n + 1;
    std::vector<uint8_t> k15((size_t)n * n);
    pk_code_kernel_matrix(c, k15.data());
    out.assign((size_t)l * l, 0);
    for (int i = 0; i < l; ++i) out[(size_t)i * l] = 1;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) out[(size_t)(i + 1) * l + j + 1] = k15[(size_t)i * n + j];
}

std::string pk_polar_parse(pk_polar_code &c, const std::string &spec_text) {
    std::istringstream in(spec_text);
    int n_short = 0, n_punct = 0;
    in >> c.N >> c.K >> c.min_dist >> c.layers >> n_short >> n_punct;
    if (!in) return "Error reading file header";
    if (c.K > c.N) return "Code dimension cannot exceed code length";
    if (c.layers < 1 || c.layers > PK_POLAR_MAX_LAYERS) return "Unsupported number of layers";
    std::map<std::string, int> seen;
    c.N0 = 1;
    c.ksize.clear(); c.kid.clear(); c.kernels.clear();
    for (int i = 0; i < c.layers; ++i) {
        std::string name;
        in >> name;
        if (!in) return "Error reading kernel names";
        auto it = seen.find(name);
        if (it == seen.end()) {
            if (name.empty() || (name[0] != '-' && name[0] != '<'))
                return "Unknown kernel " + name + " (only matrix kernels loaded from a file, \"-file\" or \"<file\", have sm_100a processors)";
            std::ifstream kf(name.substr(1));
            PkKernelTrellis kt;
            kf >> kt.size;
            if (!kf) return "Error reading kernel file " + name.substr(1);
            if (kt.size < 2 || kt.size >= 64) return "Kernel is too big for the trellis processor";
            kt.matrix.resize((size_t)kt.size * kt.size);
            for (auto &v : kt.matrix) {
                unsigned x;
                kf >> x;
                v = x ? 1 : 0;
            }
            if (!kf) return "Error parsing kernel file " + name.substr(1);
            std::string err = pk_polar_build_trellis(kt);
            if (!err.empty()) return err;
            seen[name] = (int)c.kernels.size();
            c.kernels.push_back(kt);
            it = seen.find(name);
        }
        c.kid.push_back(it->second);
        c.ksize.push_back(c.kernels[it->second].size);
        c.N0 *= c.ksize.back();
    }
    if (c.N + n_short + n_punct != c.N0) return "Code length mismatch";
    if (c.N0 > 4096) return "Unsupported code length";
    c.symtype.clear();
    if (n_short + n_punct) {
        c.symtype.assign(c.N0, 0);
        for (int i = 0; i < n_short; ++i) {
            int s;
            in >> s;
            if (!in) return "Error loading shortened symbols";
            if (s < 0 || s >= c.N0) return "Invalid shortened symbol";
            c.symtype[s] = 1;
        }
        for (int i = 0; i < n_punct; ++i) {
            int s;
            in >> s;
            if (!in) return "Error loading punctured symbols";
            if (s < 0 || s >= c.N0) return "Invalid punctured symbol";
            c.symtype[s] = 2;
        }
    }
    c.decision.assign(c.N0, -1);
    c.constraints.clear();
    const int nw = (c.N0 + 31) / 32;
    c.cmask.assign((size_t)c.N0 * nw, 0);
    for (int i = 0; i < c.N0 - c.K; ++i) {
        int w;
        in >> w;
        if (!in || w < 1) return "Error reading freezing constraint " + std::to_string(i);
        std::vector<int> terms(w);
        for (int j = 0; j < w; ++j) {
            in >> terms[j];
            if (!in) return "Error parsing freezing constraint " + std::to_string(i);
            if (terms[j] < 0 || terms[j] >= c.N0) return "Invalid term in freezing constraint " + std::to_string(i);
            if (j > 0 && terms[j] <= terms[j - 1]) return "Invalid freezing constraint " + std::to_string(i);
        }
        const int fz = terms.back();
        if (c.decision[fz] != -1) return "Duplicate freezing constraint on symbol " + std::to_string(fz);
        c.decision[fz] = i;
        for (int j = 0; j + 1 < w; ++j) c.cmask[(size_t)fz * nw + (terms[j] >> 5)] |= 1u << (terms[j] & 31);
        c.constraints.push_back(terms);
    }
    c.info_pos.clear();
    for (int i = 0; i < c.N0; ++i)
        if (c.decision[i] < 0) c.info_pos.push_back(i);
    if ((int)c.info_pos.size() != c.K) return "Number of freezing constraints does not match the code dimension";
    c.outer.assign(c.layers + 1, 0);
    c.outer[0] = c.N0;
    for (int j = 0; j < c.layers; ++j) c.outer[j + 1] = c.outer[j] / c.ksize[j];
    return "";
}
