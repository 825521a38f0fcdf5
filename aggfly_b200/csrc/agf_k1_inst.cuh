// agf_k1_inst.cuh -- K1 launcher + first-fit instantiation table; included by the agf_k1_*.cu
// units with AGF_T (float|double), AGF_TMA (0|1), AGF_FN (entry name) and AGF_LIST (the
// K1CASE(NL, NS, DIAG, KINDS) rows of that unit, cheapest first) defined.
#include <cmath>
#include <cstring>

#include "agf_host.h"

using namespace agf;

static float f32_round_down(double t) {
    float f = (float)t;
    if ((double)f > t) f = nextafterf(f, -INFINITY);
    return f;
}
static float f32_round_up(double t) {
    float f = (float)t;
    if ((double)f < t) f = nextafterf(f, INFINITY);
    return f;
}
// v > t0 (fp64 compare, as the reference does) <=> v > lo for a float v when lo = largest float <= t0;
// v < t1 <=> v < hi with hi = smallest float >= t1.  Keeps the hot loop in fp32 compares, bit-exact.
static inline void set_thresholds(LaneP<float> &L, double t0, double t1) {
    L.lo = f32_round_down(t0);
    L.hi = f32_round_up(t1);
}
static inline void set_thresholds(LaneP<double> &L, double t0, double t1) {
    L.lo = t0;
    L.hi = t1;
}

// descriptor + launch arguments -> the kernel's parameter block (lane / slot / column tables in KERNEL order)
// Typed bin lanes of a float program -> BinEdgesP (agf_kernels.cuh): the edge form needs contiguous ascending bins whose
// thresholds are all bfloat16s (>= 16 trailing zero bits; +-inf qualify), so that hi_j == lo_j+1 exactly and a value can
// equal an edge only if its low mantissa bits are zero.  AGF_BINS_BY_EDGES=0 keeps the bin-by-bin count.
template <typename LANE>
static void k1_fill_bin_edges(agf::BinEdgesP &be, const LANE *lanes, int n_bins) {
    memset(&be, 0, sizeof(be));
    be.zero_k = -1;
    for (int k = 0; k <= AGF_MAX_LANES; ++k) be.edge_f[k] = INFINITY;
    if (n_bins < 1 || n_bins > AGF_MAX_LANES) return;
    if (getenv("AGF_BINS_BY_EDGES") && atoi(getenv("AGF_BINS_BY_EDGES")) == 0) return;
    float e[AGF_MAX_LANES + 1];
    for (int j = 0; j < n_bins; ++j) e[j] = (float)lanes[j].lo;
    e[n_bins] = (float)lanes[n_bins - 1].hi;
    for (int j = 0; j + 1 < n_bins; ++j)
        if ((float)lanes[j + 1].lo != (float)lanes[j].hi) return;   // a gap, an overlap or a threshold that is not a float
    unsigned low_all = 0xffffffffu;
    for (int k = 0; k <= n_bins; ++k) {
        unsigned b;
        memcpy(&b, &e[k], 4);
        if (e[k] != e[k] || (b & 0xffffu) != 0u) return;
        if (k > 0 && !(e[k - 1] < e[k])) return;
        low_all &= ~b;
    }
    int tz = 0;
    while (tz < 23 && ((low_all >> tz) & 1u)) ++tz;
    be.eq_mask = (1u << tz) - 1u;
    be.n_edges = n_bins + 1;
    for (int k = 0; k <= n_bins; ++k) {
        unsigned b;
        memcpy(&b, &e[k], 4);
        unsigned h = b >> 16;
        if (e[k] > 0.0f) h -= 1u;   // v >= e  <=>  trunc(v) > the bfloat16 below e
        if (e[k] == 0.0f) be.zero_k = k;
        be.edge_f[k] = e[k];
        be.edge_pk[k] = h | (h << 16);
    }
    be.fast = 1;
}

template <typename T, int NL, int NS, bool DIAG, unsigned KINDS, int NB>
static void k1_fill_params(K1Params<T, NL, NS> &kp, const K1Launch &a) {
    const agf_program *p = a.p;
    memset(&kp, 0, sizeof(kp));
    const agf_program_desc_t &d = p->desc;
    kp.x = (const T *)a.d_x;
    kp.ld = a.ld;
    kp.row0 = a.row0;
    kp.n_cells = (int)p->n_cells;
    kp.stripe0 = a.s0;
    kp.b1 = p->d_b1;
    kp.b2 = p->d_b2;
    kp.stripes = p->d_stripes;
    kp.partial = a.d_partial;
    kp.out = a.d_out;
    kp.valid = a.d_valid;
    kp.n_lanes = d.n_lanes;
    kp.n_slots = d.n_slots;
    kp.n_cols = d.n_cols;
    kp.out_ncols = a.ncols;
    kp.valid_and = a.vand;
    kp.in_f64 = d.in_dtype == AGF_F64;
    kp.out_f64 = d.out_dtype == AGF_F64;
    kp.need_nan = p->need_nan;
    kp.need_cnt = p->need_cnt;
    kp.has_sine = p->has_sine;
    kp.diag = DIAG;
    kp.n_pre = d.n_pre;
    for (int i = 0; i < d.n_pre; ++i) {
        kp.pre[i].op = d.pre[i].op;
        kp.pre[i].c = (T)d.pre[i].c;
    }
    {
        // does the chain fit (v + b0) * a + b1, stage by stage?  stage 0: add, 1: multiply, 2: add
        T b0 = (T)-0.0, a = (T)1.0, b1 = (T)-0.0;
        int stage = 0;
        bool ok = d.n_pre > 0;
        for (int i = 0; i < d.n_pre && ok; ++i) {
            const T c = (T)d.pre[i].c;
            switch (d.pre[i].op) {
                case AGF_PRE_ADD:
                case AGF_PRE_SUB: {
                    const T add = d.pre[i].op == AGF_PRE_ADD ? c : -c;
                    if (stage == 0) { b0 = add; stage = 1; }
                    else if (stage <= 2) { b1 = add; stage = 3; }
                    else ok = false;
                    break;
                }
                case AGF_PRE_MUL:
                    if (stage <= 1) { a = c; stage = 2; }
                    else ok = false;
                    break;
                case AGF_PRE_NEG:
                    if (stage <= 1) { a = (T)-1.0; stage = 2; }
                    else ok = false;
                    break;
                case AGF_PRE_RSUB:  // c - x == x * (-1) + c
                    if (stage <= 1) { a = (T)-1.0; b1 = c; stage = 3; }
                    else ok = false;
                    break;
                default:
                    ok = false;  // divisions keep the exact out-of-line path
            }
        }
        kp.pre_linear = ok ? 1 : 0;
        kp.pre_b0 = b0;
        kp.pre_a = a;
        kp.pre_b1 = b1;
    }
    // kernel lane of each program lane: identity, except for typed lanes (single-level, NB >= 0),
    // where the bin lanes go to [0, NB) and the mean / sum lanes to [NB, NL), both in program order
    int lane_of[AGF_MAX_LANES];
    constexpr bool TL = typed_lanes<NS, NB>();
    constexpr bool MIX = !TL && KINDS == KIND_MIX_SD;  // lane 0 = the mean / sum lane, 1.. = the dd lanes
    constexpr bool MMS = !TL && KINDS == KIND_MMS && NL >= 3;   // lane 0 = mean / sum, 1 = min, 2 = max
    if constexpr (MMS) {
        for (int l = 0; l < d.n_lanes; ++l)
            lane_of[l] = d.lanes[l].calc == AGF_CALC_MIN ? 1 : (d.lanes[l].calc == AGF_CALC_MAX ? 2 : 0);
        for (int l = 0; l < NL; ++l) {  // inert pads: accumulators nobody reads
            LaneP<T> &L = kp.lanes[l];
            L.calc = l == 1 ? AGF_CALC_MIN : (l == 2 ? AGF_CALC_MAX : AGF_CALC_SUM);
            L.lo = (T)INFINITY;
            L.hi = (T)-INFINITY;
            L.t0 = INFINITY;
            L.t1 = -INFINITY;
        }
        kp.n_lanes = NL;
    } else if constexpr (MIX) {
        int nd = 1;
        for (int l = 0; l < d.n_lanes; ++l)
            lane_of[l] = (kind_of_calc(d.lanes[l].calc) == KIND_SUM) ? 0 : nd++;
        for (int l = 0; l < NL; ++l) {  // inert pads: a sum nobody reads / a dd range no value falls in
            LaneP<T> &L = kp.lanes[l];
            L.calc = l == 0 ? AGF_CALC_SUM : AGF_CALC_DD;
            L.lo = (T)INFINITY;
            L.hi = (T)-INFINITY;
            L.t0 = INFINITY;
            L.t1 = -INFINITY;
            if (NS == 0) kp.cols[l].dst = -1;
        }
        kp.n_lanes = NL;
        if (NS == 0) kp.n_cols = NL;  // diagonal columns travel with their lanes; pads have dst = -1
    } else if constexpr (TL) {
        int nb = 0, ns = NB;
        for (int l = 0; l < d.n_lanes; ++l) lane_of[l] = (d.lanes[l].calc == AGF_CALC_BINS) ? nb++ : ns++;
        for (int l = 0; l < NL; ++l) {  // inert pads: a bin no value falls in / a sum nobody stores
            LaneP<T> &L = kp.lanes[l];
            L.calc = l < NB ? AGF_CALC_BINS : AGF_CALC_SUM;
            L.lo = (T)INFINITY;
            L.hi = (T)-INFINITY;
            L.t0 = INFINITY;
            L.t1 = -INFINITY;
            kp.cols[l].dst = -1;
        }
    } else {
        for (int l = 0; l < d.n_lanes; ++l) lane_of[l] = l;
    }
    for (int l = 0; l < d.n_lanes; ++l) {
        LaneP<T> &L = kp.lanes[lane_of[l]];
        L.calc = d.lanes[l].calc;
        L.flag = d.lanes[l].flag;
        L.t0 = d.lanes[l].t0;
        L.t1 = d.lanes[l].t1;
        L.base = (d.lanes[l].flag == 0) ? L.t0 : L.t1;  // nb_kernels.py:168
        L.base_f = (float)L.base;
        L.base_is_f32 = (sizeof(T) == 4 && (double)L.base_f == L.base) ? 1 : 0;
        set_thresholds(L, L.t0, L.t1);
    }
    kp.be.fast = 0;
    if constexpr (TL && sizeof(T) == 4) {
        int n_bins = 0;
        for (int l = 0; l < d.n_lanes; ++l) n_bins += d.lanes[l].calc == AGF_CALC_BINS;
        if (NB > 4) k1_fill_bin_edges(kp.be, kp.lanes, n_bins);   // kernel lanes [0, n_bins): the bins in program order
    }
    if (NS > 0) {
        auto fill = [&](SlotP &S, int j) {
            S.src = lane_of[d.slots[j].src];
            S.xform = d.slots[j].xform;
            S.xparam = d.slots[j].xparam;
            S.x_f64 = d.slots[j].x_f64;
            S.calc = d.slots[j].calc;
            S.flag = d.slots[j].flag;
            S.t0 = d.slots[j].t0;
            S.t1 = d.slots[j].t1;
            S.base = (S.flag == 0) ? S.t0 : S.t1;
            S.ip = (S.xform == AGF_XF_POWI) ? (int)S.xparam : 1;
            S.flo = f32_round_down(S.t0);
            S.fhi = f32_round_up(S.t1);
            S.dst = j;
        };
        if (NB >= 0) {
            // typed form: kernel slots [0, NB) = the program's bin slots, [NB, NS) = its power
            // sums, both in program order; the rest are inert pads (dst = -1: never written back)
            int nb = 0, ns = NB;
            for (int j = 0; j < d.n_slots; ++j) {
                const bool bins = slot_kind_of(d.slots[j].calc, d.slots[j].xform) == SK_BINS;
                fill(kp.slots[bins ? nb++ : ns++], j);
            }
            for (; nb < NB; ++nb) {
                SlotP &S = kp.slots[nb];
                S.calc = AGF_CALC_BINS;
                S.t0 = S.flo = INFINITY;
                S.t1 = S.fhi = -INFINITY;
                S.ip = 1;
                S.dst = -1;
            }
            for (; ns < NS; ++ns) {
                SlotP &S = kp.slots[ns];
                S.calc = AGF_CALC_SUM;
                S.ip = 1;
                S.x_f64 = 1;
                S.dst = -1;
            }
        } else {
            for (int j = 0; j < d.n_slots; ++j) fill(kp.slots[j], j);
            for (int j = d.n_slots; j < NS; ++j) kp.slots[j].dst = -1;  // unused kernel slots name no program slot
        }
    } else {
        for (int c = 0; c < d.n_cols; ++c) {
            // typed lanes are diagonal (column c reads lane c): the column moves with its lane
            ColP &C = kp.cols[(TL || MIX) ? lane_of[d.cols[c].src] : c];
            C.src = (TL || MIX || MMS) ? lane_of[d.cols[c].src] : d.cols[c].src;
            C.xform = d.cols[c].xform;
            C.xparam = d.cols[c].xparam;
            C.x_f64 = d.cols[c].x_f64;
            C.dst = d.cols[c].dst;
        }
    }
    if (NS > 0) {
        // direct output: one stripe covers the whole time axis, so every level-2 group is reduced inside one
        // thread -- the kernel writes X / V itself; columns then name KERNEL slots (typed kernels reorder them)
        kp.direct_out = p->direct_out;
        if (kp.direct_out) {
            for (int c = 0; c < d.n_cols; ++c) {
                int ks = -1;
                for (int j = 0; j < NS; ++j)
                    if (kp.slots[j].dst == d.cols[c].src) ks = j;
                ColP &C = kp.cols[c];
                C.src = ks;
                C.xform = d.cols[c].xform;
                C.xparam = d.cols[c].xparam;
                C.x_f64 = d.cols[c].x_f64;
                C.dst = d.cols[c].dst;
                if (ks < 0) kp.direct_out = 0;  // cannot happen for a validated descriptor
            }
        }
    }
}

template <typename T, int NL, int NS, bool DIAG, unsigned KINDS, int NB, int GL, bool TMA>
static int launch_k1(const K1Launch &a) {
    const agf_program *p = a.p;
    const agf_program_desc_t &d = p->desc;
    K1Params<T, NL, NS> kp;
    k1_fill_params<T, NL, NS, DIAG, KINDS, NB>(kp, a);
    constexpr bool TL = typed_lanes<NS, NB>();
    if constexpr (TMA) {
        // rows of the view the stripes of this launch can touch
        const int64_t row_end = p->b1[p->stripes[a.s1 - 1].g1_end];
        // Ring shape by register budget of the instantiation (measured with tools/k1_sweep on the
        // C3 program: 3 CTAs x 3 stages 6.87 TB/s, 2 x 4 6.01 TB/s, 4 x 2 6.68 TB/s; tiles of 8 or
        // 12 rows lose 20-40% to per-tile synchronisation): three 9-warp CTAs per SM need <= 72
        // registers per thread, two need <= 112.
        // (a double raster holds its value batches and thresholds in register pairs: +27)
        constexpr int state_regs = (TL ? NB + 2 * (NL - NB) : 2 * NL + (NB >= 0 ? NB + 2 * (NS - NB) : 2 * NS)) +
                                   (sizeof(T) == 8 ? 27 : 0);
        constexpr int MINB = state_regs <= 30 ? 3 : (state_regs <= 48 ? 2 : 1);
        // kernels that stage their output block (single-level, several columns) give one ring stage
        // back to the staging area so that MINB CTAs still fit in the SM's 227 KB
        constexpr int SW = stage_bytes_per_warp<NL, NS>();
        constexpr int STAGES = (MINB == 3 ? 3 : (MINB == 2 ? 4 : 8)) - (SW > 0 ? 1 : 0);
        constexpr int TT = tma_rows<T>();
        const int out_esz = d.out_dtype == AGF_F64 ? 8 : 4;
        kp.stage_out = (SW > 0 && d.n_cols == a.ncols && a.ncols > 1 && 32 * a.ncols * out_esz <= SW) ? 1 : 0;
        TensorMap tm;
        int rc = agf_make_tensor_map(&tm, a.d_x, (int)sizeof(T), (uint64_t)p->n_cells, (uint64_t)(row_end - a.row0),
                                     (uint64_t)a.ld, TT);
        if (rc) return rc;
        constexpr int smem = STAGES * TMA_TILE_BYTES_DEFAULT + 2 * STAGES * 8 + (TMA_CW / 32) * SW;
        void (*kern)(const K1Params<T, NL, NS>, const TensorMap);
        if constexpr (GL > 0)
            kern = agf_k1_tma_uni<T, NL, NS, DIAG, KINDS, NB, GL, TT, STAGES, MINB>;
        else
            kern = agf_k1_tma<T, NL, NS, DIAG, KINDS, NB, TT, STAGES, MINB>;
        static bool attr_set = false;  // per instantiation
        if (!attr_set) {
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr_set = true;
        }
        dim3 grid((unsigned)((p->n_cells + TMA_CW - 1) / TMA_CW), (unsigned)(a.s1 - a.s0));
        kern<<<grid, TMA_THREADS, smem, a.stream>>>(kp, tm);
    } else {
        dim3 grid((unsigned)((p->n_cells + K1_THREADS - 1) / K1_THREADS), (unsigned)(a.s1 - a.s0));
        agf_k1_ldg<T, NL, NS, DIAG, KINDS, NB><<<grid, K1_THREADS, 0, a.stream>>>(kp);
    }
    CU(cudaGetLastError());
    return 0;
}

static inline bool k1_fits(const agf_program *p, int NL, int NS, bool DG, unsigned KINDS, int NB, int GL) {
    const agf_program_desc_t &d = p->desc;
    if (GL > 0 && p->uniform_gl != GL) return false;  // uniform-group kernel: every level-1 group has GL rows
    if (d.n_lanes > NL) return false;
    if ((NS == 0) != (d.n_slots == 0) || d.n_slots > NS) return false;
    if (DG && !p->diag_ok) return false;
    if (NS > 0 && NL > 4 && !DG) return false;  // select-chain form only for <= 4 lanes
    if (KINDS == KIND_MMS && NL >= 3 && !(NS == 0 && NB >= 0)) {  // fixed layout: one mean / sum lane, one min, one max
        int nsum = 0, nmin = 0, nmax = 0;
        for (int l = 0; l < d.n_lanes; ++l) {
            nsum += kind_of_calc(d.lanes[l].calc) == KIND_SUM;
            nmin += d.lanes[l].calc == AGF_CALC_MIN;
            nmax += d.lanes[l].calc == AGF_CALC_MAX;
        }
        if ((p->kinds & ~KIND_MMS) || nsum > 1 || nmin > 1 || nmax > 1 || nsum + nmin + nmax != d.n_lanes) return false;
        if (DG) return false;  // columns / slots pick their lane through the select chain
    }
    if (KINDS == KIND_MIX_SD && !(NS == 0 && NB >= 0)) {  // mixed layout: at most one mean / sum lane + dd lanes
        int nsum = 0;
        for (int l = 0; l < d.n_lanes; ++l) nsum += kind_of_calc(d.lanes[l].calc) == KIND_SUM;
        if ((p->kinds & ~KIND_MIX_SD) || nsum > 1 || d.n_lanes - nsum > NL - 1) return false;
        if (NS == 0 && (!DG || !p->diag_ok)) return false;
        if (NS > 0 && DG) return false;  // slots pick their lane through the select chain
    }
    if (NS == 0 && NB >= 0) {  // typed lanes: bins -> [0, NB), mean / sum -> [NB, NL); diagonal columns
        if (!DG || !p->diag_ok || (p->kinds & ~(KIND_SUM | KIND_BINS))) return false;
        if (p->max_group_rows >= (1 << 24)) return false;  // float bin counters stay exact below 2^24
        int nbins = 0;
        for (int l = 0; l < d.n_lanes; ++l) nbins += d.lanes[l].calc == AGF_CALC_BINS;
        if (nbins > NB || d.n_lanes - nbins > NL - NB) return false;
    }
    if (NS > 0 && NB >= 0) {                     // typed slots: bins -> [0, NB), power sums -> [NB, NS)
        if (p->slot_kinds & SK_GEN) return false;
        if (p->n_bin_slots > NB || d.n_slots - p->n_bin_slots > NS - NB) return false;
        if (DG && p->n_bin_slots != 0 && p->n_bin_slots != d.n_slots) return false;  // keep slot j == lane j
    }
    return (p->kinds & ~KINDS) == 0;
}

#ifdef AGF_LIST
int AGF_FN(const K1Launch &a, int mode, K1Choice *choice, int *rc) {
    const agf_program *p = a.p;
#define K1CASE(NL, NS, DG, KINDS, NB, GL)                                             \
    if (k1_fits(p, NL, NS, DG, KINDS, NB, GL)) {                                      \
        if (choice) *choice = K1Choice{NL, NS, DG ? 1 : 0, KINDS, NB, GL};            \
        if (mode == 0) *rc = launch_k1<AGF_T, NL, NS, DG, KINDS, NB, GL, AGF_TMA>(a); \
        return 0;                                                                     \
    }
    AGF_LIST
#undef K1CASE
    return 1;
}
#endif  // AGF_LIST
