// sm_100a kernel instantiations, GF(2^8): BCH(255,247,3) .. (255,139,31) -- t in [4, 5, 6]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m8_1[] = {PkLaunch<8, 4>::make(), PkLaunch<8, 5>::make(), PkLaunch<8, 6>::make()};
extern const int pk_sets_m8_1_n = sizeof(pk_sets_m8_1) / sizeof(pk_sets_m8_1[0]);
