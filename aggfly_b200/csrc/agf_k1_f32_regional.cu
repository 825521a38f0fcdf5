// K1R instantiations: float raster, 24-row periods (hourly -> date), single-level programs; see agf_regional.cuh.
// Rows are RCASE(kernel lanes, diagonal, lane kinds, NB, lanes per slot), tried in order, cheapest first.
#define AGF_T float
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>


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
    // Shared memory: ring + barriers + the staged rows (+ the zero row) + the segment sums + the tile's tables.  One ring
    // stage is enough for three CTAs per SM: a stage is handed back as soon as its 24 values per thread sit in registers,
    // and the scan is bound by instruction issue, not by bytes in flight (13.5 KB per SM cover DRAM latency at 2 TB/s).
    // (228 KB per SM, 1 KB reserved per CTA: three CTAs get 75 KB each, two 113 KB)
    constexpr int budget3 = 75 * 1024, budget2 = 113 * 1024, budget1 = 200 * 1024;
    constexpr int rest = 128 + stage_bytes<LPS>() + rg_part_bytes<LPS>() + rg_table_bytes<LPS>();
    constexpr int MINB = (MINB0 == 3 && TILE + rest + 512 * 16 <= budget3) ? 3 : ((MINB0 >= 2 && TILE + rest + 512 * 16 <= budget2) ? 2 : 1);
    constexpr int budget = MINB == 3 ? budget3 : (MINB == 2 ? budget2 : budget1);
    constexpr int STAGES = (MINB != 3 && 2 * TILE + rest + 1024 * 16 <= budget) ? 2 : 1;
    constexpr int fixed = STAGES * TILE + rest;
    constexpr int sm_entries = (budget - fixed) / 16 < 4096 ? (budget - fixed) / 16 : 4096;
    constexpr int smem = fixed + sm_entries * 16;
    static_assert(sm_entries >= 512, "no room for the tile tables");
    auto kern = agf_k1_regional<T, NL, DIAG, KINDS, NB, LPS, R_GL, TT, STAGES, MINB>;

    static int ctas_per_sm = 0;  // per instantiation
    int sms = 148;
    {
        int dev = -1;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (ctas_per_sm == 0) {
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, TMA_THREADS, smem));
            if (ctas_per_sm < 1) return agf_fail(AGF_E_UNSUPPORTED, "regional kernel does not fit on an SM");
        }
    }
    if (choice) {
        choice->lanes = NL;
        choice->typed_bins = NB;
        choice->lps = LPS;
        choice->smem_bytes = smem;
        choice->ctas_per_sm = ctas_per_sm;
    }
    if (mode != 0) return 0;
    const int64_t n_groups = a.g_end - a.g_begin;
    const int64_t ws_bytes = (int64_t)plan->n_partial_rows * a.G * LPS * 16;
    unsigned char *ws = (unsigned char *)(((uintptr_t)a.d_workspace + 255) & ~(uintptr_t)255);
    if (plan->n_partial_rows > 0 && a.workspace_bytes - (ws - (unsigned char *)a.d_workspace) < ws_bytes)
        return agf_fail(AGF_E_INVALID, "workspace of %lld bytes, %lld needed", (long long)a.workspace_bytes, (long long)ws_bytes + 256);

    K1Params<T, NL, 0> kp;
    k1_fill_params<T, NL, 0, DIAG, KINDS, NB>(kp, a.k);
    RegionalP q;
    memset(&q, 0, sizeof(q));
    q.tile_ids = plan->d_tile_ids;
    q.tile_slot_ptr = plan->d_tile_slot_ptr;
    q.slot_dst = plan->d_slot_dst;
    q.slot_ent_ptr = plan->d_slot_ent_ptr;
    q.entries = (const RgEntry *)plan->d_entries;
    const agf_rplan::SegTables *sg = agf_rplan_segments(plan, LPS);
    if (!sg) return AGF_E_INVALID;
    q.tile_seg_ptr = sg->d_tile_seg_ptr;
    q.tile_pent_ptr = sg->d_tile_pent_ptr;
    q.grp = (const int2 *)sg->d_grp;
    q.seg = (const int2 *)sg->d_seg;
    q.slot_q = (const int2 *)sg->d_slot_q;
    q.pent = (const RgEntry *)sg->d_pent;
    q.n_active = plan->n_active;
    q.tiles_x = plan->tiles_x;
    q.sm_entries = sm_entries;
    q.g_begin = (int)a.g_begin;
    q.n_groups = (int)n_groups;
    q.row_begin = (long long)p->b1[a.g_begin] - a.k.row0;
    q.partial = (double *)ws;
    q.panel = a.d_panel;
    q.den_out = a.d_den;
    q.G = a.G;
    q.n_cols = a.out_ncols;
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
    if constexpr (S::TL && S::NBL > 4 && S::NA >= 1) {
        // contiguous ascending bins -> counted through their edges (agf_regional.cuh).  Interior edges that ARE floats
        // (lo_j+1 == hi_j) need the equality screen: usable when their bit patterns end in >= 16 zero bits.
        bool ok = true;
        unsigned low = 0xffffffffu;   // bits that are zero in every representable interior edge
        bool any_rep = false;
        for (int j = 0; j + 1 < S::NBL && ok; ++j) {
            const float hi = kp.lanes[j].hi, lo_next = kp.lanes[j + 1].lo;
            if (!(kp.lanes[j].lo < lo_next) || !(hi == hi) || !(lo_next == lo_next)) ok = false;
            if (lo_next == hi) {
                unsigned b;
                memcpy(&b, &hi, 4);
                low &= ~b;
                any_rep = true;
            } else if (lo_next != nextafterf(hi, -INFINITY)) {
                ok = false;
            }
        }
        const float hi_last = kp.lanes[S::NBL - 1].hi;
        if (!(kp.lanes[S::NBL - 1].lo < hi_last) || std::isinf(hi_last)) ok = false;
        unsigned mask = 0xffffffffu;
        if (any_rep) {
            int tz = 0;
            while (tz < 23 && ((low >> tz) & 1u)) ++tz;   // trailing bits that are zero in every such edge
            mask = (1u << tz) - 1u;
            if (tz < 16) ok = false;
        }
        int want = 2;
        if (getenv("AGF_BINS_BY_EDGES")) want = atoi(getenv("AGF_BINS_BY_EDGES"));
        q.bins_fast = (ok && want > 0) ? 1 : 0;
        q.eq_mask = mask;
        for (int j = 0; j < S::NBL; ++j) q.edge_f[j] = kp.lanes[j].lo;
        q.edge_f[S::NBL] = nextafterf(hi_last, -INFINITY);
        // Packed form: every edge representable in the packed type (interior edges: hi_j == lo_j+1), none a NaN; the
        // screen then covers the first lower and the last upper threshold too.
        if (ok && want >= 2) {
            bool pk_ok = true;
            unsigned low_all = 0xffffffffu;
            float e[17];
            for (int j = 0; j < S::NBL; ++j) e[j] = kp.lanes[j].lo;
            e[S::NBL] = hi_last;
            for (int j = 0; j + 1 < S::NBL; ++j)
                if (kp.lanes[j + 1].lo != kp.lanes[j].hi) pk_ok = false;
            unsigned epk[17];
            for (int k = 0; k <= S::NBL && pk_ok; ++k) {
                unsigned b;
                memcpy(&b, &e[k], 4);
                if (e[k] != e[k]) pk_ok = false;
                low_all &= ~b;
                epk[k] = 0u;
                if ((b & 0xffffu) != 0u) { pk_ok = false; break; }
                unsigned h = b >> 16;
                if (e[k] > 0.0f) h -= 1u;   // v >= e  <=>  trunc(v) > the bfloat16 below e
                epk[k] = h | (h << 16);
                if (e[k] == 0.0f) epk[k] = 0xffffffffu;   // the edge 0.0 is counted by signs
            }
            if (pk_ok) {
                int tz = 0;
                while (tz < 23 && ((low_all >> tz) & 1u)) ++tz;
                if (tz < 13) pk_ok = false;   // the screen would send most periods to the slow path
                if (pk_ok) {
                    q.eq_mask = (1u << tz) - 1u;
                    q.zero_k = -1;
                    for (int k = 0; k <= S::NBL; ++k) {
                        if (e[k] == 0.0f) q.zero_k = k;
                        q.edge_f[k] = e[k];
                        q.edge_pk[k] = epk[k];
                    }
                    q.bins_fast = 2;
                }
            }
        }
    }
    if (plan->n_empty_regions > 0) {
        agf_regional_fill_empty<<<plan->n_regions, 256, 0, a.k.stream>>>(plan->d_region_slot_ptr, plan->n_regions, q.g_begin,
                                                                           q.n_groups, q.G, q.n_cols, q.panel, q.den_out);
        CU(cudaGetLastError());
    }
    if (plan->n_active == 0) return 0;
    // periods per CTA: enough CTAs for ~8 waves of the resident grid, at least 8 periods each (the ring's ramp-up
    // and the tile's tables are paid once per CTA)
    // (16 / 32 / 64 waves measured: flat up to 48, slower beyond, profiles/r2_k1r_steps.jsonl)
    const int64_t want_ctas = 8LL * ctas_per_sm * sms;
    int64_t stripes = std::max<int64_t>(1, std::min<int64_t>((want_ctas + plan->n_active - 1) / plan->n_active, n_groups / 8));
    stripes = std::min<int64_t>(stripes, 65535);
    q.groups_per_cta = (int)((n_groups + stripes - 1) / stripes);
    stripes = (n_groups + q.groups_per_cta - 1) / q.groups_per_cta;
    // rows of the raster view this launch can touch
    const int64_t row_end = p->b1[a.g_end];
    TensorMap tm;
    int rc = agf_make_tensor_map3(&tm, a.k.d_x, (int)sizeof(T), (uint64_t)plan->n_lon, (uint64_t)plan->n_lat,
                                  (uint64_t)(row_end - a.k.row0), (uint64_t)a.k.ld, RG_TW, RG_TH, TT);
    if (rc) return rc;
    dim3 grid((unsigned)plan->n_active, (unsigned)stripes);
    kern<<<grid, TMA_THREADS, smem, a.k.stream>>>(kp, q, tm);
    CU(cudaGetLastError());
    if (plan->n_multi > 0) {
        MergeP m;
        memset(&m, 0, sizeof(m));
        m.multi_regions = plan->d_multi_regions;
        m.region_slot_ptr = plan->d_region_slot_ptr;
        m.region_slots = plan->d_region_slots;
        m.slot_dst = plan->d_slot_dst;
        m.n_multi = plan->n_multi;
        m.g_begin = q.g_begin;
        m.n_groups = q.n_groups;
        m.partial = q.partial;
        m.panel = q.panel;
        m.den_out = q.den_out;
        m.G = q.G;
        m.n_cols = q.n_cols;
        m.n_int_units = q.n_int_units;
        m.den_unit = q.den_unit;
        m.den_half = q.den_half;
        memcpy(m.dst_int, q.dst_int, sizeof(m.dst_int));
        memcpy(m.dst_dbl, q.dst_dbl, sizeof(m.dst_dbl));
        dim3 mgrid((unsigned)plan->n_multi);
        agf_regional_merge<LPS><<<mgrid, 256, 0, a.k.stream>>>(m);
        CU(cudaGetLastError());
    }
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
    RCASE(4, false, KIND_MMS, NB_GENERAL, 4)
    RCASE(2, true, KIND_MIX_SD, NB_GENERAL, 4)
    RCASE(4, true, KIND_MIX_SD, NB_GENERAL, 8)
    RCASE(4, true, KIND_DD, NB_GENERAL, 8)
    RCASE(8, true, KIND_SUM | KIND_BINS, 6, 8)
    RCASE(14, true, KIND_SUM | KIND_BINS, 13, 8)
    RCASE(16, true, KIND_SUM | KIND_BINS, 14, 16)
#undef RCASE
    return 1;
}
