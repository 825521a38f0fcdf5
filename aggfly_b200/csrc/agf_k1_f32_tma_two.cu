// K1 instantiations of this unit: float raster, TMA/shared-memory ring variant, two-level
// programs (see agf_k1_inst.cuh; rows are tried in order, cheapest first).
#define AGF_T float
#define AGF_TMA 1
#define AGF_FN agf_k1_f32_tma_two
#define AGF_LIST \
    K1CASE(1, 1, false, KIND_SUM, SK_SUM)             \
    K1CASE(1, 4, false, KIND_SUM, SK_SUM)             \
    K1CASE(1, 16, false, KIND_SUM, SK_BINS)           \
    K1CASE(1, 16, false, KIND_SUM, SK_SUM | SK_BINS)  \
    K1CASE(1, 32, false, KIND_SUM, SK_SUM | SK_BINS)  \
    K1CASE(1, 1, false, KIND_DD, SK_SUM)              \
    K1CASE(4, 4, true, KIND_DD, SK_SUM)               \
    K1CASE(16, 16, true, KIND_BINS, SK_SUM)           \
    K1CASE(16, 16, true, KIND_DD | KIND_BINS, SK_SUM) \
    K1CASE(1, 4, false, KIND_ALL, SK_ALL)             \
    K1CASE(1, 32, false, KIND_ALL, SK_ALL)            \
    K1CASE(4, 4, false, KIND_ALL, SK_ALL)             \
    K1CASE(4, 32, false, KIND_ALL, SK_ALL)            \
    K1CASE(16, 16, true, KIND_ALL, SK_ALL)
#include "agf_k1_inst.cuh"
