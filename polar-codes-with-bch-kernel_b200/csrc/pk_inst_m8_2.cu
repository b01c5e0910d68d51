// sm_100a kernel instantiations, GF(2^8): BCH(255,247,3) .. (255,139,31) -- t in [7, 8, 9]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m8_2[] = {PkLaunch<8, 7>::make(), PkLaunch<8, 8>::make(), PkLaunch<8, 9>::make()};
extern const int pk_sets_m8_2_n = sizeof(pk_sets_m8_2) / sizeof(pk_sets_m8_2[0]);
