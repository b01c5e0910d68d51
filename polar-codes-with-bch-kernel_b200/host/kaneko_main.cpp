// kaneko_main.cpp -- command-line driver with the reference's argv conventions
// (src/main.cpp:17-190):
//   kaneko_b200 <m> <t> <snr>                 one random word at <snr> dB
//   kaneko_b200 <m> <t> <snr> <file>          word + samples from <file>
//   kaneko_b200 <m> <t> <file> <p> <e>        FER sweep 0..5 dB -> <file>.csv
// optional trailing flags: --J <cap>  --seed <s>  --gpus <n> (frames of every SNR point sharded over n GPUs of the box)
//                          --errwords <file> (the reference's DEBUG dump of non-ML frames, dataForPlot.cpp:55-64; one GPU)
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "bch_coder.hpp"
#include "kaneko_processor.hpp"
#include "monte_carlo.hpp"

int main(int argc, char *argv[]) {
    try {
        long J = -1;
        int gpus = 1;
        std::vector<char *> pos;
        for (int i = 0; i < argc; ++i) {
            if (!std::strcmp(argv[i], "--J") && i + 1 < argc) J = std::atol(argv[++i]);
            else if (!std::strcmp(argv[i], "--seed") && i + 1 < argc) fun_seed = std::strtoull(argv[++i], nullptr, 10);
            else if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
            else if (!std::strcmp(argv[i], "--errwords") && i + 1 < argc) fun_errwords = argv[++i];
            else pos.push_back(argv[i]);
        }
        const int n_args = (int)pos.size();
        long power, t, p = 1, e = 1;
        double snr = 0;
        std::string filename;
        if (n_args == 4) { power = std::atoi(pos[1]); t = std::atoi(pos[2]); snr = std::atof(pos[3]); }
        else if (n_args == 5) { power = std::atoi(pos[1]); t = std::atoi(pos[2]); snr = std::atof(pos[3]); filename = pos[4]; }
        else if (n_args == 6) { power = std::atoi(pos[1]); t = std::atoi(pos[2]); filename = pos[3]; p = std::atoi(pos[4]); e = std::atoi(pos[5]); }
        else throw "Invalid arguments\n";
        if (t <= 0 || power <= 1 || t >= (1 << (power - 1)) || p <= 0 || e <= 0 || snr < 0) throw "Invalid values of arguments\n";

        pk_code *code = nullptr;
        if (pk_code_create_host((int)power, (int)t, &code) != PK_OK) throw pk_last_error();
        int n = 0, k = 0, d = 0, gSize = 0;
        pk_code_info(code, &n, &k, &d, &gSize, nullptr);
        std::vector<unsigned char> g(gSize);
        pk_code_info(code, nullptr, nullptr, nullptr, nullptr, g.data());
        std::vector<uint64_t> al(n), lg(n + 1);
        pk_code_tables(code, al.data(), lg.data());
        pk_code_destroy(code);
        std::vector<unsigned long> antilogarithms(al.begin(), al.end()), logarithms(lg.begin(), lg.end());
        printVec(g.data(), gSize);
        std::cout << "(" << n << ", " << k << ", " << d << ")\n";

        KanekoKernelProcessor decoder(power, n, t, k, antilogarithms.data(), logarithms.data(), 0.5, J);
        if (n_args == 6) {
            if (gpus != 1) decoder.useGpus(gpus);
            fun(filename, decoder, g.data(), (unsigned long)gSize, p, e, 5.0);
            return 0;
        }
        std::vector<unsigned char> res(n), decoded(n);
        std::vector<double> err(n);
        if (n_args == 4) {
            std::vector<unsigned char> info(k);
            generateRandomPoly(info.data(), k);
            multiplyPolynomials(info.data(), k, g.data(), gSize, res.data());
            printVec(res.data(), n);
            addNoise(sqrt(1 / (pow(10, snr / 10) * 2 * k / n)), res.data(), err.data(), n);
            printVec(err.data(), n);
        } else {
            std::ifstream in(filename);
            if (!in.is_open()) throw "File does not exsist!";
            for (int i = 0; i < n; ++i) { char c; in >> c; res[i] = (c == '1') ? 1 : 0; }
            printVec(res.data(), n);
            for (int i = 0; i < n; ++i) in >> err[i];
            printVec(err.data(), n);
        }
        if (n_args == 4) decoder.decode(res.data(), err.data(), decoded.data());   // main.cpp:116
        else decoder.decode(err.data(), decoded.data());                           // main.cpp:158 (file mode)
        printVec(decoded.data(), n);
        std::cout << (comparePoly(res.data(), n, decoded.data(), n) ? "Ok\n" : "Errors were not corrected!\n");
    } catch (const char *err) {
        std::cerr << err;
    }
    return 0;
}
