// K1 instantiations: float raster, single-level programs (see agf_k1_inst.cuh)
#define AGF_T float
#define AGF_FN agf_k1_f32_single
#define AGF_PART 0
#include "agf_k1_inst.cuh"
