// K1 instantiations: float raster, two-level programs (see agf_k1_inst.cuh)
#define AGF_T float
#define AGF_FN agf_k1_f32_two
#define AGF_PART 1
#include "agf_k1_inst.cuh"
