// sm_100a kernel instantiations, GF(2^4): BCH(15,11,3), (15,7,5), (15,5,7) -- t in [1, 2, 3]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m4_0[] = {PkLaunch<4, 1>::make(), PkLaunch<4, 2>::make(), PkLaunch<4, 3>::make()};
extern const int pk_sets_m4_0_n = sizeof(pk_sets_m4_0) / sizeof(pk_sets_m4_0[0]);
