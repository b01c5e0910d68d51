// TEST HARNESS (compiled by tests/test_gpu_host_classes.py): drives the C++ drop-in class Decoder
// (host/bch_decoder.hpp, the reference's headers/Decoder.h:67-78 interface) from Python.
#include <cstring>
#include <vector>

#include "bch_decoder.hpp"
#include "pk_capi.h"

extern "C" int hd_check(int m, int t, const unsigned char *words, long B, unsigned long *synd_find, long *size_find,
                        unsigned long *synd_alter, long *size_alter, unsigned char *ok, unsigned char *answers) {
    try {
        pk_code *c = nullptr;
        if (pk_code_create_host(m, t, &c) != PK_OK) return -1;
        int n = 0, k = 0;
        pk_code_info(c, &n, &k, nullptr, nullptr, nullptr);
        std::vector<uint64_t> al(n), lg(n + 1);
        pk_code_tables(c, al.data(), lg.data());
        pk_code_destroy(c);
        std::vector<unsigned long> antilog(al.begin(), al.end()), logt(lg.begin(), lg.end());
        Decoder a(m, n, t, k, antilog.data(), logt.data()), b(m, n, t, k, antilog.data(), logt.data());
        for (long f = 0; f < B; ++f) {
            const unsigned char *w = words + f * n;
            a.findSyndromPoly(w);                         // Decoder.cpp:184-207
            std::memcpy(synd_find + f * 2 * t, a.syndromPoly, sizeof(unsigned long) * 2 * t);
            size_find[f] = a.syndromPolySize;
            if (f == 0) b.findSyndromPoly(w); else b.alterSyndromPoly(w);   // Decoder.cpp:210-230, incremental
            std::memcpy(synd_alter + f * 2 * t, b.syndromPoly, sizeof(unsigned long) * 2 * t);
            size_alter[f] = b.syndromPolySize;
            ok[f] = a.decode(w, answers + f * n) ? 1 : 0;   // Decoder.cpp:298-321 on the device
        }
        return 0;
    } catch (const char *) {
        return -2;
    }
}
