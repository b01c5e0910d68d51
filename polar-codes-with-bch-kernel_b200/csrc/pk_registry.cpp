// pk_registry.cpp -- lookup of the per-(m,t) kernel sets instantiated in pk_inst_m*.cu
#include "pk_kernels.h"

extern const PkKernelSet pk_sets_m3_0[];
extern const int pk_sets_m3_0_n;
extern const PkKernelSet pk_sets_m4_0[];
extern const int pk_sets_m4_0_n;
extern const PkKernelSet pk_sets_m5_0[];
extern const int pk_sets_m5_0_n;
extern const PkKernelSet pk_sets_m5_1[];
extern const int pk_sets_m5_1_n;
extern const PkKernelSet pk_sets_m6_0[];
extern const int pk_sets_m6_0_n;
extern const PkKernelSet pk_sets_m6_1[];
extern const int pk_sets_m6_1_n;
extern const PkKernelSet pk_sets_m6_2[];
extern const int pk_sets_m6_2_n;
extern const PkKernelSet pk_sets_m6_3[];
extern const int pk_sets_m6_3_n;
extern const PkKernelSet pk_sets_m7_0[];
extern const int pk_sets_m7_0_n;
extern const PkKernelSet pk_sets_m7_1[];
extern const int pk_sets_m7_1_n;
extern const PkKernelSet pk_sets_m7_2[];
extern const int pk_sets_m7_2_n;
extern const PkKernelSet pk_sets_m8_0[];
extern const int pk_sets_m8_0_n;
extern const PkKernelSet pk_sets_m8_1[];
extern const int pk_sets_m8_1_n;
extern const PkKernelSet pk_sets_m8_2[];
extern const int pk_sets_m8_2_n;
extern const PkKernelSet pk_sets_m8_3[];
extern const int pk_sets_m8_3_n;

const PkKernelSet *pk_find_kernels(int m, int t) {
    struct Unit { const PkKernelSet *sets; int n; };
    const Unit units[] = {
        {pk_sets_m3_0, pk_sets_m3_0_n},
        {pk_sets_m4_0, pk_sets_m4_0_n},
        {pk_sets_m5_0, pk_sets_m5_0_n},
        {pk_sets_m5_1, pk_sets_m5_1_n},
        {pk_sets_m6_0, pk_sets_m6_0_n},
        {pk_sets_m6_1, pk_sets_m6_1_n},
        {pk_sets_m6_2, pk_sets_m6_2_n},
        {pk_sets_m6_3, pk_sets_m6_3_n},
        {pk_sets_m7_0, pk_sets_m7_0_n},
        {pk_sets_m7_1, pk_sets_m7_1_n},
        {pk_sets_m7_2, pk_sets_m7_2_n},
        {pk_sets_m8_0, pk_sets_m8_0_n},
        {pk_sets_m8_1, pk_sets_m8_1_n},
        {pk_sets_m8_2, pk_sets_m8_2_n},
        {pk_sets_m8_3, pk_sets_m8_3_n},
    };
    for (const Unit &u : units)
        for (int i = 0; i < u.n; ++i)
            if (u.sets[i].m == m && u.sets[i].t == t) return &u.sets[i];
    return nullptr;
}
