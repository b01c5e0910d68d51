// TEST INFRASTRUCTURE ONLY -- C-ABI harness around the reference's vendored polar library
// (/root/reference/{headers,out}/external, compiled in place through oracle/polar_shim/compat.h by
// oracle/Makefile into oracle/_ref/libpolar_ref.so).  Exposes the mixed-kernel encoder
// (MixedKernelEncoder.cpp:142), the SC-list decoder (MixedKernelListDecoder.cpp:211) and the trellis
// kernel processor (TrellisKernelProcessor.cpp:234) to the Python tests.  Bits cross this boundary as 0/1.
#define protected public
#define private public
#include "MixedKernelListDecoder.h"
#include "TrellisKernelProcessor.h"
#undef protected
#undef private
#include <sstream>

namespace {
struct PRef {
    CMixedKernelListDecoder *dec;
    unsigned L;
};
}  // namespace

extern "C" {

void *pref_create(const char *spec_text, unsigned L) {
    try {
        std::istringstream ss(spec_text);
        PRef *p = new PRef;
        p->dec = new CMixedKernelListDecoder(ss, L);
        p->L = L;
        return p;
    } catch (std::exception &e) {
        fprintf(stderr, "pref_create: %s\n", e.what());
        return nullptr;
    }
}
void pref_dims(void *h, int *N, int *K, int *N0, int *layers) {
    PRef *p = (PRef *)h;
    *N = p->dec->m_Length; *K = p->dec->m_Dimension; *N0 = p->dec->m_UnshortenedLength; *layers = p->dec->m_NumOfLayers;
}
void pref_encode(void *h, const unsigned char *info, long B, unsigned char *cw) {
    PRef *p = (PRef *)h;
    const int N = p->dec->m_Length, K = p->dec->m_Dimension;
    for (long f = 0; f < B; ++f) {
        p->dec->Encode(info + f * K, cw + f * N);
        for (int i = 0; i < N; ++i) cw[f * N + i] = cw[f * N + i] ? 1 : 0;
    }
}
// returns per frame: count of list entries, information vectors [L][K], codewords [L][N] (best first),
// path metrics [L] (m_pSortingBuffer after the final sort, MixedKernelListDecoder.cpp:257)
void pref_decode(void *h, const float *llr, long B, int *count, unsigned char *inf, unsigned char *cw, float *metric) {
    PRef *p = (PRef *)h;
    const int N = p->dec->m_Length, K = p->dec->m_Dimension, L = (int)p->L;
    for (long f = 0; f < B; ++f) {
        unsigned char *pi = inf + f * L * K, *pc = cw + f * L * N;
        int c = p->dec->Decode(llr + f * N, pi, pc);
        count[f] = c;
        for (int i = 0; i < c * K; ++i) pi[i] = pi[i] ? 1 : 0;
        for (int i = 0; i < c * N; ++i) pc[i] = pc[i] ? 1 : 0;
        for (int l = 0; l < c; ++l) metric[f * L + l] = p->dec->m_pSortingBuffer[l].first;
    }
}
unsigned long long pref_op_counts(int which) { return which ? CmpCount : SumCount; }

// Trellis kernel processor alone: for phase = 0..l-1 in order, LLRs of input symbol `phase` given the known
// inputs u[0..phase) (0/1) and the channel LLRs of the l*stride kernel outputs.
int pref_trellis_llrs(const char *kernel_file, unsigned stride, const float *chan, const unsigned char *u, float *out,
                      unsigned *active_bits /*[l][l+1] or null*/) {
    try {
        CMatrixBinaryKernel K(kernel_file);
        CTrellisKernelProcessor P(K);
        const unsigned l = K.Size();
        void *state = P.GetState(stride), *temp = P.GetTemp(stride);
        tBit *known = new tBit[l * stride];
        for (unsigned i = 0; i < l * stride; ++i) known[i] = u[i] ? BIT_1 : BIT_0;
        for (unsigned ph = 0; ph < l; ++ph) P.GetLLRs(stride, ph, known, chan, out + ph * stride, state, temp);
        if (active_bits)
            for (unsigned ph = 0; ph < l; ++ph)
                for (unsigned j = 0; j <= l; ++j) active_bits[ph * (l + 1) + j] = P.m_ppNumOfActiveBits[ph][j];
        delete[] known;
        P.FreeState(state);
        P.FreeTemp(temp);
        return (int)l;
    } catch (std::exception &e) {
        fprintf(stderr, "pref_trellis_llrs: %s\n", e.what());
        return -1;
    }
}
}  // extern "C"
