// sm_100a kernel instantiations, GF(2^7): BCH(127,120,3) .. (127,64,21) -- t in [1, 2, 3]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m7_0[] = {PkLaunch<7, 1>::make(), PkLaunch<7, 2>::make(), PkLaunch<7, 3>::make()};
extern const int pk_sets_m7_0_n = sizeof(pk_sets_m7_0) / sizeof(pk_sets_m7_0[0]);
