#include "monte_carlo.hpp"

#include <chrono>
#include <clocale>
#include <fstream>
#include <iostream>

uint64_t fun_seed = 1;

void fun(const std::string &file, KanekoKernelProcessor &decoder, const unsigned char *g, unsigned long gSize, long p,
         long e, double maxSTNR) {
    (void)g;
    (void)gSize;
    setlocale(LC_ALL, "Russian");
    std::ofstream fout(file + ".csv");
    const auto start = std::chrono::steady_clock::now();
    unsigned long long countE = 0;   // never reset between points: the reference's BER* column (dataForPlot.cpp:20,71,95)
    int idx = 0;
    for (double stnr = 0.0; stnr <= maxSTNR; stnr += 0.5, ++idx) {
        decoder.setDecodingCount();
        decoder.setComparisonCount();
        decoder.setSummCount();
        const pk_point_result r = decoder.runPoint(stnr, idx, fun_seed, p, e);
        countE += r.bit_errors;
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
