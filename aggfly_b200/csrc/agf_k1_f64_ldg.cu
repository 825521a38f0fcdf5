// K1 instantiations of this unit: double raster, direct-load variant
// (see agf_k1_inst.cuh; rows are tried in order, cheapest first).
#define AGF_T double
#define AGF_TMA 0
#define AGF_FN agf_k1_f64_ldg
#define AGF_LIST \
    K1CASE(1, 0, false, KIND_ALL, NB_GENERAL, 0)   \
    K1CASE(4, 0, false, KIND_ALL, NB_GENERAL, 0)   \
    K1CASE(16, 0, true, KIND_ALL, NB_GENERAL, 0)   \
    K1CASE(16, 0, false, KIND_ALL, NB_GENERAL, 0)  \
    K1CASE(32, 0, true, KIND_ALL, NB_GENERAL, 0)   \
    K1CASE(32, 0, false, KIND_ALL, NB_GENERAL, 0) \
    K1CASE(1, 4, false, KIND_ALL, NB_GENERAL, 0)   \
    K1CASE(1, 32, false, KIND_ALL, NB_GENERAL, 0)  \
    K1CASE(4, 4, false, KIND_ALL, NB_GENERAL, 0)   \
    K1CASE(4, 32, false, KIND_ALL, NB_GENERAL, 0)  \
    K1CASE(16, 16, true, KIND_ALL, NB_GENERAL, 0)
#include "agf_k1_inst.cuh"
