// K1 instantiations of this unit: float raster, TMA/shared-memory ring variant, two-level
// programs (see agf_k1_inst.cuh; rows are K1CASE(lanes, slots, diag, lane kinds, NB, 0) and are
// tried in order, cheapest first; NB >= 0: typed slots = NB bin counters + (slots - NB) power sums).
#define AGF_T float
#define AGF_TMA 1
#define AGF_FN agf_k1_f32_tma_two
#define AGF_LIST \
    K1CASE(1, 1, false, KIND_SUM, 0, 0)                  \
    K1CASE(1, 4, false, KIND_SUM, 0, 0)                  \
    K1CASE(1, 8, false, KIND_SUM, 8, 0)                  \
    K1CASE(1, 16, false, KIND_SUM, 16, 0)                \
    K1CASE(1, 20, false, KIND_SUM, 16, 0)                \
    K1CASE(1, 32, false, KIND_SUM, 24, 0)                \
    K1CASE(1, 1, false, KIND_DD, 0, 0)                   \
    K1CASE(4, 4, true, KIND_DD, 0, 0)                    \
    K1CASE(2, 4, false, KIND_MIX_SD, 0, 0)               \
    K1CASE(4, 4, false, KIND_MIX_SD, 0, 0)               \
    K1CASE(4, 20, false, KIND_MIX_SD, 16, 0)             \
    K1CASE(16, 16, true, KIND_BINS, 0, 0)                \
    K1CASE(16, 16, true, KIND_DD | KIND_BINS, 0, 0)      \
    K1CASE(1, 4, false, KIND_ALL, NB_GENERAL, 0)         \
    K1CASE(1, 32, false, KIND_ALL, NB_GENERAL, 0)        \
    K1CASE(4, 4, false, KIND_ALL, NB_GENERAL, 0)         \
    K1CASE(4, 32, false, KIND_ALL, NB_GENERAL, 0)        \
    K1CASE(16, 16, true, KIND_ALL, NB_GENERAL, 0)
#include "agf_k1_inst.cuh"
