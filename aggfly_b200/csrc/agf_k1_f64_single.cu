// K1 instantiations: double raster, single-level programs (see agf_k1_inst.cuh)
#define AGF_T double
#define AGF_FN agf_k1_f64_single
#define AGF_PART 0
#include "agf_k1_inst.cuh"
