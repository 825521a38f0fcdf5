// K1 instantiations: double raster, two-level programs (see agf_k1_inst.cuh)
#define AGF_T double
#define AGF_FN agf_k1_f64_two
#define AGF_PART 1
#include "agf_k1_inst.cuh"
