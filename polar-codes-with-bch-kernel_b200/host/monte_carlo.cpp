#include "monte_carlo.hpp"

#include <chrono>
#include <clocale>
#include <fstream>
#include <iostream>

#include <vector>

#include "bch_coder.hpp"

uint64_t fun_seed = 1;
std::string fun_errwords;   // DEBUG build of the reference: "out/errWords.txt" (dataForPlot.cpp:36-39); empty = off

void fun(const std::string &file, KanekoKernelProcessor &decoder, const unsigned char *g, unsigned long gSize, long p,
         long e, double maxSTNR) {
    (void)g;
    (void)gSize;
    setlocale(LC_ALL, "Russian");
    std::ofstream fout(file + ".csv");
    const auto start = std::chrono::steady_clock::now();
    unsigned long long countE = 0;   // never reset between points: the reference's BER* column (dataForPlot.cpp:20,71,95)
    int idx = 0;
    std::ofstream ferr;
    if (!fun_errwords.empty()) ferr.open(fun_errwords);
    for (double stnr = 0.0; stnr <= maxSTNR; stnr += 0.5, ++idx) {
        decoder.setDecodingCount();
        decoder.setComparisonCount();
        decoder.setSummCount();
        const pk_point_result r = decoder.runPoint(stnr, idx, fun_seed, p, e);
        countE += r.bit_errors;
        if (ferr.is_open() && (r.flags_or & PK_FLAG_NON_ML)) {
            // dataForPlot.cpp:55-64 (DEBUG): frames whose transmitted word is more likely than the decision.  The kernels
            // flag them (PK_FLAG_NON_ML); the frames of this point are decoded once more with per-frame records and the
            // flagged ones are drawn again (Philox: frame index = counter) to be written out in the reference's format.
            const long n = decoder.getN(), k = decoder.getK();
            std::vector<pk_frame_rec> recs((size_t)r.frames);
            pk_point_result again{};
            if (pk_kaneko_run_frames(decoder.handle(), stnr, idx, fun_seed, 0, (long)r.frames, recs.data(), &again) != PK_OK) throw pk_last_error();
            std::vector<unsigned char> info((size_t)k), cw((size_t)n);
            std::vector<double> y((size_t)n);
            for (uint64_t f = 0; f < r.frames; ++f) {
                if (!(recs[(size_t)f].flags & PK_FLAG_NON_ML)) continue;
                if (pk_generate_frames(decoder.handle(), stnr, idx, fun_seed, f, 1, info.data(), cw.data(), y.data()) != PK_OK) throw pk_last_error();
                ferr << stnr << "\n";
                printVec(ferr, cw.data(), (int)n);
                printVec(ferr, y.data(), (int)n);
                ferr << "\n";
            }
        }
        const double words = (double)r.frames;
        fout << stnr << "," << ((double)r.frame_errors) / r.frames << "," << ((double)countE) / r.frames / decoder.getN()
             << "," << ((double)decoder.getDecodingCount()) / words << "," << ((double)decoder.getComparisonCount()) / words
             << "," << ((double)decoder.getSummCount()) / words << "\n";
        decoder.setDecodingCount();
        decoder.setComparisonCount();
        decoder.setSummCount();
        std::cout << stnr << "\n";
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    std::cout << "Общее время: " << secs << " секунд\n";
    fout.close();
}
