// sm_100a kernel instantiations, GF(2^6): BCH(63,57,3) .. (63,30,13), (63,16,23), ... -- t in [7, 10, 11]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m6_2[] = {PkLaunch<6, 7>::make(), PkLaunch<6, 10>::make(), PkLaunch<6, 11>::make()};
extern const int pk_sets_m6_2_n = sizeof(pk_sets_m6_2) / sizeof(pk_sets_m6_2[0]);
