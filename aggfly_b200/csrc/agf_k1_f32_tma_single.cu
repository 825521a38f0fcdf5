// K1 instantiations of this unit: float raster, TMA/shared-memory ring variant
// (see agf_k1_inst.cuh; rows are tried in order, cheapest first).
#define AGF_T float
#define AGF_TMA 1
#define AGF_FN agf_k1_f32_tma_single
#define AGF_LIST \
    K1CASE(1, 0, false, KIND_SUM, NB_GENERAL, 0)                 \
    K1CASE(1, 0, false, KIND_DD, NB_GENERAL, 0)                  \
    K1CASE(1, 0, false, KIND_DDR, NB_GENERAL, 0)                 \
    K1CASE(4, 0, true, KIND_DD | KIND_DDR, NB_GENERAL, 0)        \
    K1CASE(2, 0, true, KIND_MIX_SD, NB_GENERAL, 0)               \
    K1CASE(4, 0, true, KIND_MIX_SD, NB_GENERAL, 0)               \
    K1CASE(8, 0, true, KIND_SUM | KIND_BINS, 6, 0)               \
    K1CASE(16, 0, true, KIND_SUM | KIND_BINS, 14, 0)             \
    K1CASE(32, 0, true, KIND_SUM | KIND_BINS, 28, 0)             \
    K1CASE(1, 0, false, KIND_ALL, NB_GENERAL, 0)   \
    K1CASE(4, 0, false, KIND_ALL, NB_GENERAL, 0)   \
    K1CASE(16, 0, true, KIND_ALL, NB_GENERAL, 0)   \
    K1CASE(16, 0, false, KIND_ALL, NB_GENERAL, 0)  \
    K1CASE(32, 0, true, KIND_ALL, NB_GENERAL, 0)   \
    K1CASE(32, 0, false, KIND_ALL, NB_GENERAL, 0)
#include "agf_k1_inst.cuh"
