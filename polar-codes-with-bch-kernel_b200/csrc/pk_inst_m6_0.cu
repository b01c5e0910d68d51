// sm_100a kernel instantiations, GF(2^6): BCH(63,57,3) .. (63,30,13), (63,16,23), ... -- t in [1, 2, 3]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m6_0[] = {PkLaunch<6, 1>::make(), PkLaunch<6, 2>::make(), PkLaunch<6, 3>::make()};
extern const int pk_sets_m6_0_n = sizeof(pk_sets_m6_0) / sizeof(pk_sets_m6_0[0]);
