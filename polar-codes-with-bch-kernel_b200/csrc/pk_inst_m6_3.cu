// sm_100a kernel instantiations, GF(2^6): BCH(63,57,3) .. (63,30,13), (63,16,23), ... -- t in [13, 15]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m6_3[] = {PkLaunch<6, 13>::make(), PkLaunch<6, 15>::make()};
extern const int pk_sets_m6_3_n = sizeof(pk_sets_m6_3) / sizeof(pk_sets_m6_3[0]);
