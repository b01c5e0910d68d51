// sm_100a kernel instantiations, GF(2^7): BCH(127,120,3) .. (127,64,21) -- t in [4, 5, 6]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m7_1[] = {PkLaunch<7, 4>::make(), PkLaunch<7, 5>::make(), PkLaunch<7, 6>::make()};
extern const int pk_sets_m7_1_n = sizeof(pk_sets_m7_1) / sizeof(pk_sets_m7_1[0]);
