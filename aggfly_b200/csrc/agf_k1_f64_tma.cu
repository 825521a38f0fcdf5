// K1 instantiations of this unit: double raster (what xarray decodes int16-packed ERA5 NetCDF into),
// TMA/shared-memory ring variant; the common single-kind / mixed programs first, then the generic forms
// (see agf_k1_inst.cuh; rows are tried in order, cheapest first).
#define AGF_T double
#define AGF_TMA 1
#define AGF_FN agf_k1_f64_tma
#define AGF_LIST \
    K1CASE(1, 0, false, KIND_SUM, NB_GENERAL, 0)   \
    K1CASE(1, 0, false, KIND_DD, NB_GENERAL, 0)    \
    K1CASE(1, 1, false, KIND_SUM, 0, 0)            \
    K1CASE(1, 4, false, KIND_SUM, 0, 0)            \
    K1CASE(1, 20, false, KIND_SUM, 16, 0)          \
    K1CASE(1, 1, false, KIND_DD, 0, 0)             \
    K1CASE(2, 4, false, KIND_MIX_SD, 0, 0)         \
    K1CASE(4, 4, false, KIND_MIX_SD, 0, 0)         \
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
