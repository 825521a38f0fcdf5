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

template <typename T, int NL, int NS, bool DIAG, unsigned KINDS, unsigned SK, bool TMA>
static int launch_k1(const K1Launch &a) {
    const agf_program *p = a.p;
    K1Params<T, NL, NS> kp;
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
    for (int l = 0; l < d.n_lanes; ++l) {
        LaneP<T> &L = kp.lanes[l];
        L.calc = d.lanes[l].calc;
        L.flag = d.lanes[l].flag;
        L.t0 = d.lanes[l].t0;
        L.t1 = d.lanes[l].t1;
        L.base = (d.lanes[l].flag == 0) ? L.t0 : L.t1;  // nb_kernels.py:168
        set_thresholds(L, L.t0, L.t1);
    }
    if (NS > 0) {
        for (int j = 0; j < d.n_slots; ++j) {
            SlotP &S = kp.slots[j];
            S.src = d.slots[j].src;
            S.xform = d.slots[j].xform;
            S.xparam = d.slots[j].xparam;
            S.x_f64 = d.slots[j].x_f64;
            S.calc = d.slots[j].calc;
            S.flag = d.slots[j].flag;
            S.t0 = d.slots[j].t0;
            S.t1 = d.slots[j].t1;
            S.base = (S.flag == 0) ? S.t0 : S.t1;
            S.ip = (S.xform == AGF_XF_POWI) ? (int)S.xparam : 1;
        }
        // pad the unused slots of a fast kind set with reducers that cannot fault or branch
        // (their accumulators are never written back: l2_write_rec stops at n_slots)
        for (int j = d.n_slots; j < NS; ++j) {
            SlotP &S = kp.slots[j];
            S.ip = 1;
            S.x_f64 = 1;
            S.calc = (SK & SK_BINS) ? AGF_CALC_BINS : AGF_CALC_SUM;
            S.t0 = INFINITY;
            S.t1 = -INFINITY;
        }
    } else {
        for (int c = 0; c < d.n_cols; ++c) {
            ColP &C = kp.cols[c];
            C.src = d.cols[c].src;
            C.xform = d.cols[c].xform;
            C.xparam = d.cols[c].xparam;
            C.x_f64 = d.cols[c].x_f64;
            C.dst = d.cols[c].dst;
        }
    }
    if constexpr (TMA) {
        // rows of the view the stripes of this launch can touch
        const int64_t row_end = p->b1[p->stripes[a.s1 - 1].g1_end];
        TensorMap tm;
        int rc = agf_make_tensor_map(&tm, a.d_x, (int)sizeof(T), (uint64_t)p->n_cells, (uint64_t)(row_end - a.row0),
                                     (uint64_t)a.ld, tma_rows<T>());
        if (rc) return rc;
        constexpr int smem = TMA_STAGES_DEFAULT * TMA_TILE_BYTES_DEFAULT + 2 * TMA_STAGES_DEFAULT * 8;
        static bool attr_set = false;  // per instantiation
        if (!attr_set) {
            CU(cudaFuncSetAttribute(agf_k1_tma<T, NL, NS, DIAG, KINDS, SK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr_set = true;
        }
        dim3 grid((unsigned)((p->n_cells + TMA_CW - 1) / TMA_CW), (unsigned)(a.s1 - a.s0));
        agf_k1_tma<T, NL, NS, DIAG, KINDS, SK><<<grid, TMA_THREADS, smem, a.stream>>>(kp, tm);
    } else {
        dim3 grid((unsigned)((p->n_cells + K1_THREADS - 1) / K1_THREADS), (unsigned)(a.s1 - a.s0));
        agf_k1_ldg<T, NL, NS, DIAG, KINDS, SK><<<grid, K1_THREADS, 0, a.stream>>>(kp);
    }
    CU(cudaGetLastError());
    return 0;
}

static inline bool k1_fits(const agf_program *p, int NL, int NS, bool DG, unsigned KINDS, unsigned SK) {
    const agf_program_desc_t &d = p->desc;
    if (d.n_lanes > NL) return false;
    if ((NS == 0) != (d.n_slots == 0) || d.n_slots > NS) return false;
    if (DG && !p->diag_ok) return false;
    if (NS > 0 && NL > 4 && !DG) return false;  // select-chain form only for <= 4 lanes
    if (NS > 0 && (p->slot_kinds & ~SK) != 0) return false;
    return (p->kinds & ~KINDS) == 0;
}

int AGF_FN(const K1Launch &a, int mode, K1Choice *choice, int *rc) {
    const agf_program *p = a.p;
#define K1CASE(NL, NS, DG, KINDS, SK)                                             \
    if (k1_fits(p, NL, NS, DG, KINDS, SK)) {                                      \
        if (choice) *choice = K1Choice{NL, NS, DG ? 1 : 0, KINDS, SK};            \
        if (mode == 0) *rc = launch_k1<AGF_T, NL, NS, DG, KINDS, SK, AGF_TMA>(a); \
        return 0;                                                                 \
    }
    AGF_LIST
#undef K1CASE
    return 1;
}
