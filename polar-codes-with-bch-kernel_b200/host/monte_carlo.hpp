// monte_carlo.hpp -- fun(): the reference's Monte-Carlo sweep (headers/dataForPlot.h:8,
// src/dataForPlot.cpp:16-115) with the whole per-frame loop on the GPU.
#pragma once
#include <cstdint>
#include <string>

#include "kaneko_processor.hpp"

// Same signature, CSV layout (`stnr,FER,BER*,trials/word,cmp/word,sum/word`, default ostream
// precision), stdout progress lines and final timing line as the reference.  `g`/`gSize` are
// accepted for drop-in compatibility (the device holds its own copy of g(x)).
// Frames come from the device Philox stream selected by fun_seed (default 1).
void fun(const std::string &file, KanekoKernelProcessor &decoder, const unsigned char *g, unsigned long gSize, long p,
         long e, double maxSTNR = 5.0);
extern uint64_t fun_seed;
// non-empty: write the frames the decoder did not decode maximum-likelihood there (the reference's DEBUG build, out/errWords.txt)
extern std::string fun_errwords;
