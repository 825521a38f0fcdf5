// K1R instantiations: float raster, 24-row periods (hourly -> date), single-level programs; see agf_regional.cuh.
// Rows are RCASE(kernel lanes, diagonal, lane kinds, NB, lanes per slot), tried in order, cheapest first.
#define AGF_T float
#include "agf_k1_inst.cuh"
#include "agf_regional.cuh"

namespace {

constexpr int R_GL = 24;

template <int NL, bool DIAG, unsigned KINDS, int NB, int LPS>
struct RShape {
    static constexpr bool TL = typed_lanes<0, NB>();
    static constexpr int NBL = TL ? NB : 0;
    static constexpr int NA = TL ? NL - NB : NL;          // float64 lanes
    static constexpr int N_INT = NBL + 1;                  // bin lanes + the denominator's 0 / 1
    static constexpr int N_INT_UNITS = (N_INT + 1) / 2;
    static constexpr int MAX_DBL = LPS - N_INT_UNITS;      // float64 units that fit behind the integers
};

template <int NL, bool DIAG, unsigned KINDS, int NB, int LPS>
bool regional_fits(const agf_program *p) {
    using S = RShape<NL, DIAG, KINDS, NB, LPS>;
    if (!k1_fits(p, NL, 0, DIAG, KINDS, NB, R_GL)) return false;
    const agf_program_desc_t &d = p->desc;
    if (S::TL || DIAG) return S::NA <= S::MAX_DBL;
    return d.n_cols <= S::MAX_DBL;  // columns transformed from the lanes: one float64 unit per column
}

template <int NL, bool DIAG, unsigned KINDS, int NB, int LPS>
int launch_regional(const RegionalLaunch &a, int mode, RegionalChoice *choice) {
    using T = float;
    using S = RShape<NL, DIAG, KINDS, NB, LPS>;
    const agf_program *p = a.k.p;
    const agf_rplan *plan = a.plan;
    // ring shape by register budget, like launch_k1
    constexpr int state_regs = S::TL ? NB + 2 * (NL - NB) : 2 * NL;
    constexpr int MINB0 = state_regs <= 30 ? 3 : (state_regs <= 48 ? 2 : 1);
    constexpr int TT = R_GL;
    constexpr int TILE = TT * TMA_CW * (int)sizeof(T);
    // shared memory: ring + barriers + staged rows + merge lists; three CTAs per SM need <= 75 KB each
    constexpr int fixed3 = 2 * TILE + 128 + stage_bytes<LPS>() + 16;
    constexpr int MINB = (MINB0 == 3 && fixed3 + 2 * 4 * 64 <= 75 * 1024) ? 3 : (MINB0 >= 2 ? 2 : 1);
    constexpr int STAGES = MINB == 3 ? 2 : (MINB == 2 ? 3 : 6);
    const int smem = STAGES * TILE + 128 + stage_bytes<LPS>() + 16 + 2 * plan->max_slots * 4;
    auto kern = agf_k1_regional<T, NL, DIAG, KINDS, NB, LPS, R_GL, TT, STAGES, MINB>;

    const int D = a.D;
    const int64_t n_blocks = (a.g_end - a.g_begin + D - 1) / D;
    const int64_t ring = (a.ring > 0 && a.ring < n_blocks) ? a.ring : n_blocks;
    const int64_t n_units = n_blocks * plan->n_active;
    const int64_t partial_bytes = ring * plan->n_gslots * (int64_t)D * LPS * 16;
    const int64_t cnt_bytes = ((ring * plan->n_regions * 4 + 255) / 256) * 256;
    const int64_t done_bytes = ((n_blocks * 4 + 255) / 256) * 256;
    const int64_t ws_bytes = partial_bytes + cnt_bytes + done_bytes + 256;
    if (n_units > 0x7fffffffLL) return agf_fail(AGF_E_UNSUPPORTED, "too many units of work for one launch");
    if (smem > 200 * 1024) return agf_fail(AGF_E_UNSUPPORTED, "a tile touches %d regions: merge lists do not fit", plan->max_slots);

    static int ctas_per_sm = 0;  // per instantiation
    static int smem_set = 0;
    int sms = 148;
    if (mode == 0 || choice) {
        int dev = -1;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (smem > smem_set) {
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            smem_set = smem;
            ctas_per_sm = 0;
        }
        if (ctas_per_sm == 0) {
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, TMA_THREADS, smem));
            if (ctas_per_sm < 1) return agf_fail(AGF_E_UNSUPPORTED, "regional kernel does not fit on an SM");
        }
    }
    // the grid must be entirely resident (tiles wait for earlier day-blocks when the partial ring is reused)
    const int grid = (int)std::min<int64_t>((int64_t)ctas_per_sm * sms, n_units);
    if (choice) {
        choice->lanes = NL;
        choice->typed_bins = NB;
        choice->lps = LPS;
        choice->smem_bytes = smem;
        choice->ctas_per_sm = ctas_per_sm;
        choice->grid = grid;
        choice->workspace_bytes = ws_bytes;
        choice->n_units = n_units;
    }
    if (mode != 0) return 0;
    if (a.workspace_bytes < ws_bytes)
        return agf_fail(AGF_E_INVALID, "workspace of %lld bytes, %lld needed", (long long)a.workspace_bytes, (long long)ws_bytes);

    K1Params<T, NL, 0> kp;
    k1_fill_params<T, NL, 0, DIAG, KINDS, NB>(kp, a.k);
    RegionalP q;
    memset(&q, 0, sizeof(q));
    q.tile_ids = plan->d_tile_ids;
    q.tile_slot_ptr = plan->d_tile_slot_ptr;
    q.slot_region = plan->d_slot_region;
    q.slot_ent_ptr = plan->d_slot_ent_ptr;
    q.entries = (const RgEntry *)plan->d_entries;
    q.region_slot_ptr = plan->d_region_slot_ptr;
    q.region_slots = plan->d_region_slots;
    q.n_active = plan->n_active;
    q.tiles_x = plan->tiles_x;
    q.n_regions = plan->n_regions;
    q.n_gslots = plan->n_gslots;
    q.max_slots = plan->max_slots;
    q.g_begin = (int)a.g_begin;
    q.g_end = (int)a.g_end;
    q.D = D;
    q.n_blocks = (int)n_blocks;
    q.ring = (int)ring;
    unsigned char *ws = (unsigned char *)(((uintptr_t)a.d_workspace + 255) & ~(uintptr_t)255);
    q.partial = (double *)ws;
    q.cnt = (int *)(ws + partial_bytes);
    q.done = (int *)(ws + partial_bytes + cnt_bytes);
    q.panel = a.d_panel;
    q.den_out = a.d_den;
    q.G = a.G;
    q.n_cols = a.out_ncols;
    q.n_int = S::N_INT;
    q.n_int_units = S::N_INT_UNITS;
    q.den_unit = (S::N_INT - 1) >> 1;
    q.den_half = (S::N_INT - 1) & 1;
    for (int i = 0; i < 32; ++i) q.dst_int[i] = -1;
    for (int i = 0; i < 16; ++i) q.dst_dbl[i] = -1;
    if constexpr (S::TL) {
        for (int j = 0; j < S::NBL; ++j) q.dst_int[j] = kp.cols[j].dst;
        for (int l = 0; l < S::NA; ++l) q.dst_dbl[l] = kp.cols[S::NBL + l].dst;
    } else if constexpr (DIAG) {
        for (int l = 0; l < NL; ++l) q.dst_dbl[l] = l < kp.n_cols ? kp.cols[l].dst : -1;
    } else {
        for (int c = 0; c < kp.n_cols && c < 16; ++c) q.dst_dbl[c] = kp.cols[c].dst;
    }
    CU(cudaMemsetAsync(q.cnt, 0, (size_t)(cnt_bytes + done_bytes), a.k.stream));
    if (plan->n_empty_regions > 0) {
        agf_regional_fill_empty<<<plan->n_regions, 256, 0, a.k.stream>>>(plan->d_region_slot_ptr, plan->n_regions, q.g_begin,
                                                                           q.g_end, q.G, q.n_cols, q.panel, q.den_out);
        CU(cudaGetLastError());
    }
    if (n_units == 0) return 0;
    // rows of the raster view this launch can touch
    const int64_t row_end = p->b1[a.g_end];
    TensorMap tm;
    int rc = agf_make_tensor_map3(&tm, a.k.d_x, (int)sizeof(T), (uint64_t)plan->n_lon, (uint64_t)plan->n_lat,
                                  (uint64_t)(row_end - a.k.row0), (uint64_t)a.k.ld, RG_TW, RG_TH, TT);
    if (rc) return rc;
    kern<<<grid, TMA_THREADS, smem, a.k.stream>>>(kp, q, tm);
    CU(cudaGetLastError());
    return 0;
}

}  // namespace

int agf_k1_f32_regional(const RegionalLaunch &a, int mode, RegionalChoice *choice, int *rc) {
    const agf_program *p = a.k.p;
    if (p->desc.n_slots != 0 || p->uniform_gl != R_GL) return 1;
#define RCASE(NL, DG, KINDS, NB, LPS)                                  \
    if (regional_fits<NL, DG, KINDS, NB, LPS>(p)) {                    \
        *rc = launch_regional<NL, DG, KINDS, NB, LPS>(a, mode, choice); \
        return 0;                                                      \
    }
    RCASE(1, false, KIND_SUM, NB_GENERAL, 4)
    RCASE(1, false, KIND_DD, NB_GENERAL, 4)
    RCASE(2, true, KIND_MIX_SD, NB_GENERAL, 4)
    RCASE(4, true, KIND_MIX_SD, NB_GENERAL, 8)
    RCASE(4, true, KIND_DD, NB_GENERAL, 8)
    RCASE(8, true, KIND_SUM | KIND_BINS, 6, 8)
    RCASE(14, true, KIND_SUM | KIND_BINS, 13, 8)
    RCASE(16, true, KIND_SUM | KIND_BINS, 14, 16)
#undef RCASE
    return 1;
}
