// sm_100a kernel instantiations, GF(2^3): BCH(7,4,3) -- t in [1]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m3_0[] = {PkLaunch<3, 1>::make()};
extern const int pk_sets_m3_0_n = sizeof(pk_sets_m3_0) / sizeof(pk_sets_m3_0[0]);
