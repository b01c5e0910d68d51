// sm_100a kernel instantiations, GF(2^5): BCH(31,26,3) .. (31,16,7), (31,11,11), (31,6,15) -- t in [5, 7]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m5_1[] = {PkLaunch<5, 5>::make(), PkLaunch<5, 7>::make()};
extern const int pk_sets_m5_1_n = sizeof(pk_sets_m5_1) / sizeof(pk_sets_m5_1[0]);
