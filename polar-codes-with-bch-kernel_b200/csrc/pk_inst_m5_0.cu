// sm_100a kernel instantiations, GF(2^5): BCH(31,26,3) .. (31,16,7), (31,11,11), (31,6,15) -- t in [1, 2, 3]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m5_0[] = {PkLaunch<5, 1>::make(), PkLaunch<5, 2>::make(), PkLaunch<5, 3>::make()};
extern const int pk_sets_m5_0_n = sizeof(pk_sets_m5_0) / sizeof(pk_sets_m5_0[0]);
