// sm_100a kernel instantiations, GF(2^6): BCH(63,57,3) .. (63,30,13), (63,16,23), ... -- t in [4, 5, 6]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m6_1[] = {PkLaunch<6, 4>::make(), PkLaunch<6, 5>::make(), PkLaunch<6, 6>::make()};
extern const int pk_sets_m6_1_n = sizeof(pk_sets_m6_1) / sizeof(pk_sets_m6_1[0]);
