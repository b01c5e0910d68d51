// sm_100a kernel instantiations, GF(2^8): BCH(255,247,3) .. (255,139,31) -- t in [10, 15]
#include "pk_kernels.cuh"
extern const PkKernelSet pk_sets_m8_3[] = {PkLaunch<8, 10>::make(), PkLaunch<8, 15>::make()};
extern const int pk_sets_m8_3_n = sizeof(pk_sets_m8_3) / sizeof(pk_sets_m8_3[0]);
