// sm_100a kernel instantiations, GF(2^8): BCH(255,247,3) .. (255,139,31) -- t in [1, 2, 3]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m8_0[] = {PkLaunch<8, 1>::make(), PkLaunch<8, 2>::make(), PkLaunch<8, 3>::make()};
extern const int pk_sets_m8_0_n = sizeof(pk_sets_m8_0) / sizeof(pk_sets_m8_0[0]);
