// K1 instantiations of this unit: float raster, TMA ring, UNIFORM level-1 groups (agf_k1_tma_uni:
// every group has GL rows; GL = 24 hourly -> date, 8 / 4 three- / six-hourly -> date, 1 daily data by date).  Rows are
// K1CASE(lanes, slots, diag, lane kinds, NB, GL); tried in order, cheapest first.  (KIND_SUM | KIND_MINMAX: daily minimum /
// maximum / mean of hourly values -- tmin, tmax, tavg -- per date or averaged over months / years: 35 ms on the general
// ragged-group kernel for a global year.)
#define AGF_T float
#define AGF_TMA 1
#define AGF_FN agf_k1_f32_tma_uni
#define AGF_LIST \
    K1CASE(1, 0, false, KIND_SUM, NB_GENERAL, 24)       \
    K1CASE(1, 1, false, KIND_SUM, 0, 24)                \
    K1CASE(1, 4, false, KIND_SUM, 0, 24)                \
    K1CASE(1, 8, false, KIND_SUM, 8, 24)                \
    K1CASE(1, 16, false, KIND_SUM, 16, 24)              \
    K1CASE(1, 20, false, KIND_SUM, 16, 24)              \
    K1CASE(1, 32, false, KIND_SUM, 24, 24)              \
    K1CASE(1, 0, false, KIND_DD, NB_GENERAL, 24)        \
    K1CASE(1, 1, false, KIND_DD, 0, 24)                 \
    K1CASE(4, 4, true, KIND_DD, 0, 24)                  \
    K1CASE(2, 0, true, KIND_MIX_SD, NB_GENERAL, 24)     \
    K1CASE(2, 4, false, KIND_MIX_SD, 0, 24)             \
    K1CASE(2, 20, false, KIND_MIX_SD, 16, 24)           \
    K1CASE(4, 0, true, KIND_MIX_SD, NB_GENERAL, 24)     \
    K1CASE(4, 4, false, KIND_MIX_SD, 0, 24)             \
    K1CASE(4, 20, false, KIND_MIX_SD, 16, 24)           \
    K1CASE(4, 0, false, KIND_SUM | KIND_MINMAX, NB_GENERAL, 24) \
    K1CASE(4, 4, false, KIND_SUM | KIND_MINMAX, 0, 24)  \
    K1CASE(8, 0, true, KIND_SUM | KIND_BINS, 6, 24)     \
    K1CASE(16, 0, true, KIND_SUM | KIND_BINS, 14, 24)   \
    K1CASE(32, 0, true, KIND_SUM | KIND_BINS, 28, 24)   \
    K1CASE(16, 16, true, KIND_BINS, 0, 24)              \
    K1CASE(1, 0, false, KIND_SUM, NB_GENERAL, 8)        \
    K1CASE(1, 4, false, KIND_SUM, 0, 8)                 \
    K1CASE(1, 20, false, KIND_SUM, 16, 8)               \
    K1CASE(1, 1, false, KIND_DD, 0, 8)                  \
    K1CASE(2, 4, false, KIND_MIX_SD, 0, 8)              \
    K1CASE(1, 0, false, KIND_SUM, NB_GENERAL, 4)        \
    K1CASE(1, 4, false, KIND_SUM, 0, 4)                 \
    K1CASE(1, 20, false, KIND_SUM, 16, 4)               \
    K1CASE(1, 1, false, KIND_DD, 0, 4)                  \
    K1CASE(2, 4, false, KIND_MIX_SD, 0, 4)              \
    K1CASE(1, 1, false, KIND_SUM, 0, 1)                 \
    K1CASE(1, 4, false, KIND_SUM, 0, 1)                 \
    K1CASE(1, 20, false, KIND_SUM, 16, 1)               \
    K1CASE(1, 1, false, KIND_DD, 0, 1)                  \
    K1CASE(4, 4, true, KIND_DD, 0, 1)
#include "agf_k1_inst.cuh"
