// sm_100a kernel instantiations, GF(2^7): BCH(127,120,3) .. (127,64,21) -- t in [7, 9, 10]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m7_2[] = {PkLaunch<7, 7>::make(), PkLaunch<7, 9>::make(), PkLaunch<7, 10>::make()};
extern const int pk_sets_m7_2_n = sizeof(pk_sets_m7_2) / sizeof(pk_sets_m7_2[0]);
