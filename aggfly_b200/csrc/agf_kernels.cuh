// agf_kernels.cuh -- device code of libaggfly_b200 (sm_100a).
//
// K1  agf_k1_*      fused temporal kernel: one thread per grid cell walks the time axis of its
//                   stripe once; level-1 reducers ("lanes": mean/sum/min/max/nanmean/dd/bins/
//                   sine_dd per group of bounds1) live in registers, are flushed at every group
//                   end into the level-2 reducers ("slots": sum/mean/min/max/dd/bins per group of
//                   bounds2, after an optional power/spline transform), also in registers.
//                   Semantics restate aggfly/aggregate/nb_kernels.py:121-251 (fp64 accumulation
//                   in time order, result rounded to the raster dtype, NaN rules :15-25) and
//                   aggfly/dataset/dataset.py:442-481,527-543 (power / spline).
// K1f agf_finalize  merges the per-stripe partial records, mean division, dtype rounding,
//                   trailing transforms, writes X[G, cells, n_cols] and the shared validity mask.
// K2  agf_spmm      CSR weighted regional average, one warp per (region, period):
//                   aggfly/aggregate/spatial.py:114-133, 181-186.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "aggfly_b200.h"

namespace agf {

constexpr int K1_THREADS = 256;  // cells per CTA (one cell per thread)
constexpr int K1_U = 8;          // rows per register batch

// Lane-kind sets.  The kinds a program uses are a COMPILE-TIME parameter of the temporal kernel:
// with a single kind the per-element code is straight-line (e.g. LDS -> F2F -> DADD for
// mean/sum); a runtime switch per element costs a branch tree per value and caps the kernel at
// ~20% of the HBM roofline (measured, profiles/).
enum : unsigned {
    KIND_SUM = 1u,      // mean, sum (NaN-propagating add)
    KIND_NANMEAN = 2u,  // nanmean
    KIND_MINMAX = 4u,   // min, max
    KIND_DD = 8u,       // degree days
    KIND_BINS = 16u,    // bin counts
    KIND_SINE = 32u,    // sine_dd + its hidden sum/min/max helper lanes
    KIND_DDR = 64u,     // degree days with every term rounded to the raster dtype (AGF_CALC_DD_R)
    KIND_ALL = 127u
};
// KIND_SUM | KIND_DD as a compile-time set means the MIXED layout: the launcher puts the program's one
// mean / sum lane in kernel lane 0 and its degree-day lanes in lanes 1.. (inert pads elsewhere), so the
// per-value code is straight-line -- the reference's own example config (daily mean -> annual mean
// plus degree days -> annual sum, examples/era5_counties_area.yaml) would otherwise run the generic
// switch-per-value kernel.
constexpr unsigned KIND_MIX_SD = KIND_SUM | KIND_DD;
// KIND_SUM | KIND_MINMAX as a compile-time set means the FIXED layout of daily minimum / maximum / mean programs (tmin,
// tmax, tavg): kernel lane 0 = the program's one mean / sum lane, lane 1 = its min lane, lane 2 = its max lane (inert
// pads where the program has none), so the per-value code is one add and two compare-selects instead of a dispatch on
// the lane's calc per lane and value (tmin + tmax + tavg by date -> year means of a global hourly year: 35 ms before, profiles/README.md).
constexpr unsigned KIND_MMS = KIND_SUM | KIND_MINMAX;

__host__ __device__ constexpr unsigned kind_of_calc(int calc) {
    return (calc == AGF_CALC_MEAN || calc == AGF_CALC_SUM)  ? KIND_SUM
           : (calc == AGF_CALC_NANMEAN)                      ? KIND_NANMEAN
           : (calc == AGF_CALC_MIN || calc == AGF_CALC_MAX)  ? KIND_MINMAX
           : (calc == AGF_CALC_DD)                           ? KIND_DD
           : (calc == AGF_CALC_DD_R)                         ? KIND_DDR
           : (calc == AGF_CALC_BINS)                         ? KIND_BINS
                                                             : KIND_SINE;
}

// Slot-kind sets (level-2 reducers), compile-time like the lane kinds: the per-group flush runs
// once per level-1 group and thread, so a runtime switch per slot (with pow() inlined per
// slot) made the flush ~30 instructions per raster value -- 5x the scan itself (ncu r1a).
// The kernels take this as the template parameter NB: NB >= 0 means "typed" slots -- the launcher
// sorts the program's slots so that kernel slots [0, NB) are bin counters (int registers) and
// [NB, NS) are power sums (double registers), padding each group with inert reducers; NB == -1
// is the general form (any calc / transform per slot, runtime switch).
enum : unsigned {
    SK_SUM = 1u,   // sum / mean of x or x^p (p a small non-negative integer)
    SK_BINS = 2u,  // bin count of x
    SK_GEN = 4u,   // min / max / dd, general pow / spline transforms
    SK_ALL = 7u
};
constexpr int NB_GENERAL = -1;

__host__ __device__ constexpr unsigned slot_kind_of(int calc, int xform) {
    return ((calc == AGF_CALC_MEAN || calc == AGF_CALC_SUM) && (xform == AGF_XF_NONE || xform == AGF_XF_POWI))
               ? SK_SUM
           : (calc == AGF_CALC_BINS && xform == AGF_XF_NONE) ? SK_BINS
                                                             : SK_GEN;
}

// ------------------------------------------------------------------------------------------
// kernel parameter blocks (passed by value, live in the constant bank)
// ------------------------------------------------------------------------------------------
template <typename T>
struct LaneP {
    int calc;
    int flag;
    T lo, hi;  // v > t0  <=>  v > lo ;  v < t1  <=>  v < hi   (thresholds pre-rounded for float)
    double t0, t1, base;
    // float raster and a dd base that IS a float: float(|double(v) - base|) == |v - base_f| computed in
    // float (rounding a float difference to double first never changes the float result: 53 >= 2*24+2),
    // which replaces three float<->double conversions per value by one.  Conversions run on the XU
    // pipe at a quarter of the fp64 rate and were the bound of the daily dd kernel (ncu r1n: XU 79 %).
    float base_f;
    int base_is_f32;
};

// one degree-day term rounded to the raster dtype (AGF_CALC_DD_R)
template <typename T>
__device__ __forceinline__ double dd_term_rounded(const LaneP<T> &L, T v, double vd) {
    if constexpr (sizeof(T) == 4) {
        if (L.base_is_f32) return (double)fabsf(v - L.base_f);
        return (double)(float)fabs(vd - L.base);
    } else {
        return fabs(vd - L.base);
    }
}

// acc += (t0 < v < t1) ? |v - base| : 0 -- with the term ZEROED instead of the add predicated.  ``if (in) acc += term``
// compiles to an unconditional DADD, two FSELs on the halves of the double and two moves (ncu r3: half of the dd_r
// kernel's instructions per value); selecting on the FLOAT side costs one FSEL, and adding +0.0 to a sum of non-negative
// terms changes nothing (C5, daily dd by month: K1 2.76 -> 2.35 ms, 0.82 -> 0.96 x measured peak).
template <typename T>
__device__ __forceinline__ void dd_add_rounded(const LaneP<T> &L, double &acc, T v, double vd) {
    if constexpr (sizeof(T) == 4) {
        if (L.base_is_f32) {
            float t;  // (one asm block: the compiler would move the select behind the conversion, onto both halves)
            asm("{\n\t"
                ".reg .pred p;\n\t"
                "setp.gt.f32 p, %1, %2;\n\t"
                "setp.lt.and.f32 p, %1, %3, p;\n\t"
                "sub.f32 %0, %1, %4;\n\t"
                "abs.f32 %0, %0;\n\t"
                "selp.f32 %0, %0, 0f00000000, p;\n\t"
                "}"
                : "=f"(t)
                : "f"((float)v), "f"((float)L.lo), "f"((float)L.hi), "f"(L.base_f));
            acc += (double)t;
            return;
        }
    }
    if (v > L.lo && v < L.hi) acc += dd_term_rounded(L, v, vd);
}
// The plain dd lane keeps the predicated add: its term is a float64 (|double(v) - base|), and zeroing it through the input
// (v' = in ? v : base) costs one conversion per LANE and value instead of one per value -- six dd thresholds of a global
// year went from 29 to 39 ms that way (XU pipe).
template <typename T>
__device__ __forceinline__ void dd_add(const LaneP<T> &L, double &acc, T v, double vd) {
    if (v > L.lo && v < L.hi) acc += fabs(vd - L.base);
}

struct SlotP {
    int src;
    int xform;
    double xparam;
    int x_f64;
    int calc;
    int flag;
    int ip;   // integer exponent of a POWI transform (1 for no transform)
    int dst;  // slot index in the partial records (kernel slot order may differ); -1 = padding
    float flo, fhi;  // t0 / t1 rounded outward to float: exact strict compares of float values
    double t0, t1, base;
};

struct ColP {
    int src;
    int xform;
    double xparam;
    int x_f64;
    int dst;  // column in the destination X
};

struct Stripe {  // one time stripe, cut at level-1 group boundaries
    int g1_begin, g1_end;
    int g2_first;  // level-2 group of g1_begin
    int rec0;      // first partial record of this stripe
};

template <typename T>
struct PreP {  // one preprocess operation (AGF_PRE_*), constant already in the raster dtype
    int op;
    T c;
};

// Contiguous ascending bins counted through their EDGES, two values per instruction (l1_acc_group, agf_regional.cuh):
// set by the launcher (k1_fill_bin_edges) when every edge of a typed-lane program is a bfloat16.
struct BinEdgesP {
    int fast;           // 0: bins are counted one by one
    unsigned eq_mask;   // a value can equal an edge only if (bits & eq_mask) == 0
    int zero_k;         // index of the edge 0.0 (-1: none)
    int n_edges;        // real bins + 1
    float edge_f[AGF_MAX_LANES + 1];     // lo_0 .. lo_{n-1}, hi_{n-1}; +inf behind them
    unsigned edge_pk[AGF_MAX_LANES + 1]; // what the packed compare tests against, in both halves
};

template <typename T, int NL, int NS>
struct K1Params {
    const T *x;
    long long ld;
    long long row0;
    int n_cells;
    int stripe0;
    const int *b1;
    const int *b2;
    const Stripe *stripes;
    double *partial;  // [n_recs, n_slots, n_cells]      (NS > 0)
    void *out;        // X[G1, n_cells, n_cols]           (NS == 0)
    unsigned char *valid;
    int n_lanes, n_slots, n_cols;
    int out_ncols, valid_and;  // column count of the (possibly shared) X; AND into V or overwrite
    int in_f64, out_f64;
    int need_nan, need_cnt, has_sine, diag;
    int n_pre;
    int pre_linear;  // the chain has the form (v + b0) * a + b1 (identity fillers: -0.0, 1): inlined as 3 ops
    T pre_b0, pre_a, pre_b1;
    int stage_out;  // single-level: stage a warp's X[g, 32 cells, :] block in shared memory, store it coalesced
    int direct_out;  // two-level program launched as ONE stripe: write X / V straight from the slots (no partial
                     // records, no finalize launch -- the records of a single stripe have nothing to merge with)
    PreP<T> pre[AGF_MAX_PRE];
    LaneP<T> lanes[NL];
    SlotP slots[NS > 0 ? NS : 1];
    ColP cols[AGF_MAX_COLS];  // NS == 0: columns read lanes; NS > 0 (direct output only): src = KERNEL slot index
    BinEdgesP be;
};

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double agf_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double agf_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

template <typename T>
__device__ __forceinline__ double round_to(double v);
template <>
__device__ __forceinline__ double round_to<float>(double v) {
    return (double)(float)v;
}
template <>
__device__ __forceinline__ double round_to<double>(double v) {
    return v;
}

__device__ __forceinline__ double ipow(double x, int p) {
    double r = 1.0;
    double b = x;
    // square-and-multiply; for p <= 4 on f32-derived x every product is exact or a single
    // rounding, i.e. correctly rounded like libm pow
    while (p > 0) {
        if (p & 1) r *= b;
        p >>= 1;
        if (p) b *= b;
    }
    return r;
}

// transform of a value whose dtype is the raster dtype T (dataset.py:442-481, 527-543)
template <typename T>
__device__ __forceinline__ double apply_xform(double x, int xform, double xparam, int x_f64) {
    double r = x;
    if (xform == AGF_XF_NONE) return x;
    if (xform == AGF_XF_POWI) {
        r = ipow(x, (int)xparam);
    } else if (xform == AGF_XF_POW) {
        r = pow(x, xparam);
    } else {  // AGF_XF_SPLINE2: (x > 20) * (x - 20) in the value's own dtype
        if (sizeof(T) == 4 && !x_f64) {
            float xf = (float)x;
            float d = xf - 20.0f;
            r = (double)((xf > 20.0f) ? d : d * 0.0f);
        } else {
            double d = x - 20.0;
            r = (x > 20.0) ? d : d * 0.0;
        }
        return r;
    }
    return x_f64 ? r : round_to<T>(r);
}

// transform of a value that is already float64-typed (only used for trailing transforms of f64 slots)
__device__ __forceinline__ double apply_xform64(double x, int xform, double xparam) {
    if (xform == AGF_XF_NONE) return x;
    if (xform == AGF_XF_POWI) return ipow(x, (int)xparam);
    if (xform == AGF_XF_POW) return pow(x, xparam);
    double d = x - 20.0;
    return (x > 20.0) ? d : d * 0.0;
}

template <int N>
__device__ __forceinline__ double select_reg(const double (&v)[N], int idx) {
    double r = v[0];
#pragma unroll
    for (int i = 1; i < N; ++i) r = (idx == i) ? v[i] : r;
    return r;
}

// single-sine degree days from a group's mean / min / max (nb_kernels.py:211-251)
__device__ __forceinline__ double sine_dd_value(double tavg, double tmin, double tmax, double t0,
                                                double t1, int kind) {
    const double PI = 3.141592653589793;
    double val = 0.0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        double thr = j == 0 ? t0 : t1;
        double part;
        if (kind == 0) {
            if (thr <= tmin) {
                part = tavg - thr;
            } else if (thr < tmax && tmin < thr) {
                double rng = tmax - tmin;
                double a = acos((2.0 * thr - tmax - tmin) / rng);
                part = ((tavg - thr) * a + rng * sin(a) / 2.0) / PI;
            } else {
                part = 0.0;
            }
            val += (j == 0) ? part : -part;
        } else {
            if (thr >= tmax) {
                part = thr - tavg;
            } else if (thr < tmax && tmin < thr) {
                double alpha = (tmax - tmin) / 2.0;
                double r = (thr - tavg) / alpha;
                double at = atan(r / sqrt(1.0 - r * r));
                part = (1.0 / PI) * ((thr - tavg) * (at + PI / 2.0) + alpha * cos(at));
            } else {
                part = 0.0;
            }
            val += (j == 0) ? -part : part;
        }
    }
    return val;
}

// ------------------------------------------------------------------------------------------
// per-thread reducer state (everything statically indexed -> registers)
// ------------------------------------------------------------------------------------------
// Typed lanes (single-level programs, NS == 0 and NB >= 0): the launcher orders the lanes so that
// kernel lanes [0, NB) are bin counters (int registers, predicated adds) and [NB, NL) are mean / sum
// lanes (double registers) -- the level-1 mirror of the typed slots.
template <int NS, int NB>
__host__ __device__ constexpr bool typed_lanes() {
    return NS == 0 && NB >= 0;
}

template <typename T, int NL, int NS, int NB = NB_GENERAL>
struct CellState {
    static constexpr bool TL = typed_lanes<NS, NB>();
    static constexpr int ND = (NB >= 0) ? (NS - NB) : NS;  // double-typed level-2 accumulators
    static constexpr int NA = TL ? (NL - NB) : NL;         // double-typed level-1 accumulators
    double a[NA > 0 ? NA : 1];  // level-1 accumulators (typed lanes: the mean / sum lanes)
    double b[ND > 0 ? ND : 1];  // level-2 accumulators (typed slots: the power sums)
    int c[NB > 0 ? NB : 1];     // bin counters: level-2 (typed slots)
    // level-1 bin counters (typed lanes) are FLOAT registers: the predicated "+ 1.0f" then issues on the
    // FMA pipe while the two compares of every (value, bin) pair keep the ALU pipe busy -- the hourly-bins
    // scan is ALU-bound (ncu r1h) and an integer add would be a third ALU instruction.  Exact while a
    // group has fewer than 2^24 rows (checked by the launcher).
    float cf[TL && NB > 0 ? NB : 1];
    int nn;                     // non-NaN values in the current level-1 group
    bool nan;                   // NaN seen in the current level-1 group
};

template <unsigned KINDS, typename T, int NL, int NS, typename ST>
__device__ __forceinline__ void l1_init(const K1Params<T, NL, NS> &p, ST &s) {
    if constexpr (ST::TL) {
#pragma unroll
        for (int j = 0; j < NL - ST::NA; ++j) s.cf[j] = 0.0f;
#pragma unroll
        for (int l = 0; l < ST::NA; ++l) s.a[l] = 0.0;
        s.nn = 0;
        s.nan = false;
        return;
    }
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        if constexpr ((KINDS & (KIND_MINMAX | KIND_SINE)) == 0) {
            s.a[l] = 0.0;
        } else {
            int c = (l < p.n_lanes) ? p.lanes[l].calc : AGF_CALC_SUM;
            s.a[l] = (c == AGF_CALC_MIN || c == AGF_CALC_HIDDEN_MIN)   ? agf_inf()
                     : (c == AGF_CALC_MAX || c == AGF_CALC_HIDDEN_MAX) ? -agf_inf()
                                                                       : 0.0;
        }
    }
    s.nn = 0;
    s.nan = false;
}

template <int NB, typename T, int NL, int NS, typename ST>
__device__ __forceinline__ void l2_init(const K1Params<T, NL, NS> &p, ST &s) {
    if constexpr (NS > 0) {
        if constexpr (NB >= 0) {
#pragma unroll
            for (int j = 0; j < NB; ++j) s.c[j] = 0;
#pragma unroll
            for (int j = 0; j < NS - NB; ++j) s.b[j] = 0.0;
        } else {
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                int c = (j < p.n_slots) ? p.slots[j].calc : AGF_CALC_SUM;
                s.b[j] = (c == AGF_CALC_MIN) ? agf_inf() : (c == AGF_CALC_MAX) ? -agf_inf() : 0.0;
            }
        }
    }
}

// c += (v > lo && v < hi): two compares + one predicated add (the C form compiles to add / select /
// move chains).  lo / hi are the thresholds rounded outward to the raster dtype (set_thresholds).
__device__ __forceinline__ void count_in_range(int &c, float v, float lo, float hi) {
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.gt.f32 p, %1, %2;\n\t"
        "setp.lt.and.f32 p, %1, %3, p;\n\t"
        "@p add.s32 %0, %0, 1;\n\t"
        "}"
        : "+r"(c)
        : "f"(v), "f"(lo), "f"(hi));
}
__device__ __forceinline__ void count_in_range(int &c, double v, double lo, double hi) {
    if (v > lo && v < hi) c += 1;
}
// float counters (typed lanes)
__device__ __forceinline__ void count_in_range(float &c, float v, float lo, float hi) {
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.gt.f32 p, %1, %2;\n\t"
        "setp.lt.and.f32 p, %1, %3, p;\n\t"
        "@p add.f32 %0, %0, 0f3F800000;\n\t"
        "}"
        : "+f"(c)
        : "f"(v), "f"(lo), "f"(hi));
}
__device__ __forceinline__ void count_in_range(float &c, double v, double lo, double hi) {
    if (v > lo && v < hi) c += 1.0f;
}

// the program's preprocess chain on one raster value, in the raster dtype (one rounding per op).
// General form: out of line -- the IEEE division routine inlined at every load site (24 per tile in
// the uniform kernel) made the kernels 3x larger for a path most programs never take.
template <typename T>
__device__ __noinline__ T pre_apply_general(const PreP<T> *pre, int n_pre, T v) {
#pragma unroll 1
    for (int i = 0; i < n_pre; ++i) {
        {
            const T c = pre[i].c;
            switch (pre[i].op) {
                case AGF_PRE_ADD: v = v + c; break;
                case AGF_PRE_SUB: v = v - c; break;
                case AGF_PRE_RSUB: v = c - v; break;
                case AGF_PRE_MUL: v = v * c; break;
                case AGF_PRE_DIV: v = v / c; break;
                case AGF_PRE_RDIV: v = c / v; break;
                default: v = -v; break;
            }
        }
    }
    return v;
}
// Linear chains (x - 273.15, x * 1000, 1.8 * x + 32, c - x ...) are three inline operations with
// identity fillers: adding -0.0 and multiplying by 1 return their operand bit for bit, x - c == x + (-c)
// and c - x == x * (-1) + c in IEEE arithmetic, and the library is built without FMA contraction.
template <typename T, int NL, int NS>
__device__ __forceinline__ T pre_apply(const K1Params<T, NL, NS> &p, T v) {
    if (p.pre_linear) {
        v = v + p.pre_b0;
        v = v * p.pre_a;
        v = v + p.pre_b1;
        return v;
    }
    return pre_apply_general<T>(p.pre, p.n_pre, v);
}
// a register batch of raster values: one uniform branch per batch, nothing when there is no chain
template <int N, typename T, int NL, int NS>
__device__ __forceinline__ void pre_apply_batch(const K1Params<T, NL, NS> &p, T (&v)[N]) {
    if (p.n_pre != 0) {
        if (p.pre_linear) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                v[i] = v[i] + p.pre_b0;
                v[i] = v[i] * p.pre_a;
                v[i] = v[i] + p.pre_b1;
            }
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] = pre_apply_general<T>(p.pre, p.n_pre, v[i]);  // N call sites: v stays in registers
        }
    }
}

// one raster value into every level-1 lane (nb_kernels.py:134-141, 170-177, 193-196, 213-220)
template <unsigned KINDS, typename T, int NL, int NS, typename ST>
__device__ __forceinline__ void l1_acc(const K1Params<T, NL, NS> &p, ST &s, T v) {
    const double vd = (double)v;
    if constexpr (KINDS == KIND_MIX_SD && !ST::TL) {  // lane 0: mean / sum, lanes 1..: degree days
        s.nan |= (v != v);
        s.a[0] += vd;
#pragma unroll
        for (int l = 1; l < NL; ++l)
            dd_add(p.lanes[l], s.a[l], v, vd);
        return;
    }
    if constexpr (KINDS == KIND_MMS && !ST::TL && NL >= 3) {  // lane 0: mean / sum, lane 1: min, lane 2: max (launcher)
        s.nan |= (v != v);
        s.a[0] += vd;
        if (vd < s.a[1]) s.a[1] = vd;
        if (vd > s.a[2]) s.a[2] = vd;
        return;
    }
    if constexpr (ST::TL) {  // typed lanes: straight-line, no per-lane dispatch
        constexpr int NBL = NL - ST::NA;
#pragma unroll
        for (int j = 0; j < NBL; ++j) count_in_range(s.cf[j], v, p.lanes[j].lo, p.lanes[j].hi);
#pragma unroll
        for (int l = 0; l < ST::NA; ++l) s.a[l] += vd;  // a NaN value poisons the sum
        return;
    }
    bool isn = false;
    if constexpr ((KINDS & (KIND_NANMEAN | KIND_MINMAX | KIND_DD | KIND_DDR | KIND_SINE)) != 0) isn = (v != v);
    if constexpr ((KINDS & (KIND_MINMAX | KIND_DD | KIND_DDR | KIND_SINE)) != 0) s.nan |= isn;
    if constexpr ((KINDS & (KIND_NANMEAN | KIND_SINE)) != 0) s.nn += isn ? 0 : 1;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        if (NL == 1 || l < p.n_lanes) {
            const LaneP<T> &L = p.lanes[l];
            if constexpr (KINDS == KIND_SUM) {
                s.a[l] += vd;  // a NaN value poisons the sum == "any NaN -> NaN"
            } else if constexpr (KINDS == KIND_BINS) {
                if (v > L.lo && v < L.hi) s.a[l] += 1.0;
            } else if constexpr (KINDS == KIND_DD) {
                dd_add(L, s.a[l], v, vd);
            } else if constexpr (KINDS == KIND_DDR) {
                dd_add_rounded(L, s.a[l], v, vd);
            } else {
                switch (L.calc) {
                    case AGF_CALC_MEAN:
                    case AGF_CALC_SUM:
                        if constexpr ((KINDS & KIND_SUM) != 0) s.a[l] += vd;
                        break;
                    case AGF_CALC_NANMEAN:
                        if constexpr ((KINDS & KIND_NANMEAN) != 0)
                            if (!isn) s.a[l] += vd;
                        break;
                    case AGF_CALC_HIDDEN_SUM:
                        if constexpr ((KINDS & KIND_SINE) != 0)
                            if (!isn) s.a[l] += vd;
                        break;
                    case AGF_CALC_MIN:
                        if constexpr ((KINDS & KIND_MINMAX) != 0)
                            if (vd < s.a[l]) s.a[l] = vd;
                        break;
                    case AGF_CALC_HIDDEN_MIN:
                        if constexpr ((KINDS & KIND_SINE) != 0)
                            if (vd < s.a[l]) s.a[l] = vd;
                        break;
                    case AGF_CALC_MAX:
                        if constexpr ((KINDS & KIND_MINMAX) != 0)
                            if (vd > s.a[l]) s.a[l] = vd;
                        break;
                    case AGF_CALC_HIDDEN_MAX:
                        if constexpr ((KINDS & KIND_SINE) != 0)
                            if (vd > s.a[l]) s.a[l] = vd;
                        break;
                    case AGF_CALC_DD:
                        if constexpr ((KINDS & KIND_DD) != 0)
                            dd_add(L, s.a[l], v, vd);
                        break;
                    case AGF_CALC_DD_R:
                        if constexpr ((KINDS & KIND_DDR) != 0)
                            dd_add_rounded(L, s.a[l], v, vd);
                        break;
                    case AGF_CALC_BINS:
                        if constexpr ((KINDS & KIND_BINS) != 0)
                            if (v > L.lo && v < L.hi) s.a[l] += 1.0;
                        break;
                    default:
                        break;
                }
            }
        }
    }
}

// ---- counting the values of a period above an edge, two values per instruction ----
// The 24 floats of a period are TRUNCATED to bfloat16 once (one PRMT packs the high halves of two values) and every edge
// inside the warp's range is tested with 12 HSET2.BF16 + 12 HADD2.BF16 instead of 24 FSETP + 24 predicated FADD.  That is
// exact, not approximate, on the fast path (no value equals an edge, no NaN next to a number):
//   * e < 0:  v > e  <=>  trunc(v) > e.   A negative v is truncated towards zero, i.e. UP onto the bfloat16 grid the edge
//     lies on, so it cannot cross e from below (v < e => trunc(v) <= e would need trunc(v) == e only for v in (e - ulp, e):
//     truncation moves such a v to the grid point ABOVE it only if that point is e itself -- and then v > e was false
//     and trunc(v) > e is false); a non-negative v stays non-negative.
//   * e > 0:  v > e  <=>  v >= e (screen)  <=>  trunc(v) >= e  <=>  trunc(v) > pred(e), pred(e) the bfloat16 below e:
//     a positive v is truncated DOWN onto the grid; a negative v stays <= -0 < pred(e) or == pred(e) = +0 (false).
//   * e == 0: truncation takes tiny values to +-0, so the SIGNS are counted instead (no value is +-0 on the fast path):
//     one PRMT replicates the sign bits of four values into bytes, one IDP4A adds the four -1 / 0.
// An all-NaN cell (a NaN may truncate to an infinity) is zeroed by the caller; partly-NaN cells never get here.
__device__ __forceinline__ unsigned rg_pack_hi(float a, float b) {  // (bfloat16 trunc(a), bfloat16 trunc(b))
    unsigned d;
    asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(d) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
    return d;
}
__device__ __forceinline__ unsigned rg_bf2_gt(unsigned a, unsigned e) {  // per half: 1.0 if a > e else 0.0
    unsigned d;
    asm("set.gt.bf16x2.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(e));
    return d;
}
__device__ __forceinline__ unsigned rg_bf2_add(unsigned a, unsigned b) {
    unsigned d;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// g += (v > edge): one compare (ALU pipe) + one predicated add (FMA pipe) -- the float32 form, for edges that are not
// bfloat16s (bins_fast == 1).
__device__ __forceinline__ void count_above(float &g, float v, float edge) {
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.gt.f32 p, %1, %2;\n\t"
        "@p add.f32 %0, %0, 0f3F800000;\n\t"
        "}"
        : "+f"(g)
        : "f"(v), "f"(edge));
}
template <int NP>
__device__ __forceinline__ float rg_count_above_packed(const unsigned (&pk)[NP], unsigned epk) {
    static_assert(NP >= 4, "four chains");
    unsigned a[4];  // four chains of exact small integers (<= NP / 4 per half)
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = rg_bf2_gt(pk[i], epk);
#pragma unroll
    for (int i = 4; i < NP; ++i) a[i & 3] = rg_bf2_add(a[i & 3], rg_bf2_gt(pk[i], epk));
    const unsigned s = rg_bf2_add(rg_bf2_add(a[0], a[1]), rg_bf2_add(a[2], a[3]));
    return __uint_as_float(s << 16) + __uint_as_float(s & 0xffff0000u);
}
template <int NP>
__device__ __forceinline__ float rg_count_positive_packed(const unsigned (&pk)[NP]) {
    static_assert(NP % 2 == 0, "pairs of packed registers");
    int c = 2 * NP;
#pragma unroll
    for (int i = 0; i < NP; i += 2) {
        unsigned sg;
        asm("prmt.b32 %0, %1, %2, 0xfdb9;" : "=r"(sg) : "r"(pk[i]), "r"(pk[i + 1]));  // 0xff per negative value
        asm("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(c) : "r"(sg), "r"(0x01010101));
    }
    return __uint_as_float(0x4B000000u + (unsigned)c) - 8388608.0f;
}

// A whole level-1 group held in registers (uniform-group kernel).  Typed lanes with many bins: the
// bins a warp can possibly hit are bounded by the min / max of its 32 x N values, and neighbouring
// cells at the same hours are close in value, so each bin's N x (2 compares + add) block is guarded
// by one warp-uniform test -- the cost follows the bins that are populated, not the bins that
// exist (the scan is ALU-bound: ncu r1f, 13 bins x 24 h = 1008 ALU-pipe instructions per warp-day).
// Must be called by all 32 lanes of a warp.
template <unsigned KINDS, int N, typename T, int NL, int NS, typename ST>
__device__ __forceinline__ void l1_acc_group(const K1Params<T, NL, NS> &p, ST &s, const T (&v)[N]) {
    if constexpr (ST::TL && (NL - ST::NA) > 4 && N >= 8) {
        constexpr int NBL = NL - ST::NA;
        T mn = v[0], mx = v[0];  // fmin / fmax skip NaNs; an all-NaN group compares false everywhere
#pragma unroll
        for (int r = 1; r < N; ++r) {
            mn = fmin(mn, v[r]);
            mx = fmax(mx, v[r]);
        }
        // Edge form (float rasters, every edge a bfloat16): this lane can take it unless one of its values could equal
        // an edge (low mantissa bits all zero) or only SOME of them are NaN (a float sum of the batch is NaN then; an
        // overflow to infinities only costs the slow path)
        constexpr bool EDGES = sizeof(T) == 4 && N % 4 == 0;
        bool fast = false, all_nan = false;
        if constexpr (EDGES) {
            if (p.be.fast) {
                unsigned em = 0xffffffffu;
                float f4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    em = min(em, __float_as_uint((float)v[r]) & p.be.eq_mask);
                    f4[r & 3] += (float)v[r];
                }
                const float fs = (f4[0] + f4[1]) + (f4[2] + f4[3]);
                all_nan = mn != mn;
                fast = !__any_sync(0xffffffffu, !all_nan && (em == 0u || fs != fs));
            }
        }
        if constexpr (sizeof(T) == 4) {
            // warp reductions (sm_100a: CREDUX.MIN / MAX.F32; NaNs are skipped like fmin / fmax skip them) instead of a
            // butterfly of ten dependent shuffles
            float rn, rx;
            asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(rn) : "f"((float)mn));
            asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(rx) : "f"((float)mx));
            mn = (T)rn;
            mx = (T)rx;
        } else {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
                mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            }
        }
        if (fast) {
            if constexpr (EDGES) {
                // G(e) = #(v > e) for the edges inside the warp's range (below it G = the lane's valid values, above it
                // 0), bin j += G(lo_j) - G(lo_j+1): one packed compare + add per TWO values and edge instead of two compares
                // + an add per value and bin (see rg_count_above_packed for why the truncated compare is exact)
                unsigned pk[N / 2];
#pragma unroll
                for (int i = 0; i < N / 2; ++i) pk[i] = rg_pack_hi((float)v[2 * i], (float)v[2 * i + 1]);
                const float n_valid = all_nan ? 0.0f : (float)N;
                float gprev = 0.0f;
#pragma unroll
                for (int k = 0; k <= NBL; ++k) {
                    const float edge = p.be.edge_f[k];
                    float gk;
                    if (edge >= (float)mn && edge < (float)mx) {  // warp-uniform
                        gk = (k == p.be.zero_k) ? rg_count_positive_packed(pk) : rg_count_above_packed(pk, p.be.edge_pk[k]);
                        if (all_nan) gk = 0.0f;
                    } else {
                        gk = (edge < (float)mn) ? n_valid : 0.0f;
                    }
                    if (k > 0) s.cf[k - 1] += gprev - gk;
                    gprev = gk;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < NBL; ++j) {
                if (mx > p.lanes[j].lo && mn < p.lanes[j].hi) {  // warp-uniform
#pragma unroll
                    for (int r = 0; r < N; ++r) count_in_range(s.cf[j], v[r], p.lanes[j].lo, p.lanes[j].hi);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < N; ++r) {
            const double vd = (double)v[r];
#pragma unroll
            for (int l = 0; l < ST::NA; ++l) s.a[l] += vd;
        }
    } else {
#pragma unroll
        for (int r = 0; r < N; ++r) l1_acc<KINDS>(p, s, v[r]);
    }
}

// s / n for a group of n rows.  With a compile-time row count GLC and a float raster this is the
// IEEE quotient without the division subroutine (whose special-case path ncu showed running for
// every NaN / zero sum, ~75 instructions): rcp = RN(1/n) is exact to half an ulp, q0 = RN(s*rcp)
// is within 1.5 ulp of s/n, the FMA residual r = s - n*q0 is exact, and RN(q0 + r*rcp) is the
// correctly rounded quotient (Markstein's correction step: the perturbation |r/n|*2^-53 cannot
// cross a rounding boundary because s/n with n < 2^31 is never closer than 2^-32 ulp to one, and
// an odd n > 1 cannot produce an exact midpoint).  A sum of floats is 0, NaN, +-inf or
// >= 2^-149 in magnitude, so there is no underflow; +-inf and NaN pass through.
template <typename T, int GLC>
__device__ __forceinline__ double mean_of(double s, int n) {
    if constexpr (GLC == 1) {
        return s;
    } else if constexpr (GLC > 1 && sizeof(T) == 4) {
        constexpr double dn = (double)GLC;
        constexpr double rcp = 1.0 / dn;
        const double q0 = s * rcp;
        const double r = fma(-dn, q0, s);
        const double q = fma(r, rcp, q0);
        return (fabs(q0) <= 1.7976931348623157e308) ? q : q0;
    } else {
        return s / (double)n;
    }
}

// value of lane l for a finished group of n_grp rows, rounded to the raster dtype (:143-155, :260)
template <unsigned KINDS, int GLC, typename T, int NL, int NS, typename ST>
__device__ __forceinline__ double l1_value(const K1Params<T, NL, NS> &p, const ST &s, int l, int n_grp) {
    const LaneP<T> &L = p.lanes[l];
    double r;
    if constexpr (KINDS == KIND_SUM) {
        r = (L.calc == AGF_CALC_MEAN) ? mean_of<T, GLC>(s.a[l], n_grp) : s.a[l];
    } else if constexpr (KINDS == KIND_BINS) {
        r = s.a[l];
    } else if constexpr (KINDS == KIND_DD || KINDS == KIND_DDR) {
        r = s.nan ? agf_nan() : s.a[l];
    } else if constexpr (KINDS == KIND_MIX_SD) {
        r = (l == 0) ? ((L.calc == AGF_CALC_MEAN) ? mean_of<T, GLC>(s.a[0], n_grp) : s.a[0])
                     : (s.nan ? agf_nan() : s.a[l]);
    } else {
        switch (L.calc) {
            case AGF_CALC_MEAN:
                r = mean_of<T, GLC>(s.a[l], n_grp);
                break;
            case AGF_CALC_NANMEAN:
                r = (s.nn > 0) ? s.a[l] / (double)s.nn : agf_nan();
                break;
            case AGF_CALC_MIN:
            case AGF_CALC_MAX:
            case AGF_CALC_DD:
            case AGF_CALC_DD_R:
                r = s.nan ? agf_nan() : s.a[l];
                break;
            case AGF_CALC_SINE_DD:
                // lanes 0..2 are the hidden sum / min / max helpers (host guarantees it)
                if constexpr ((KINDS & KIND_SINE) != 0)
                    r = (s.nan || s.nn == 0)
                            ? agf_nan()
                            : sine_dd_value(s.a[0] / (double)s.nn, s.a[NL > 1 ? 1 : 0], s.a[NL > 2 ? 2 : 0],
                                            L.t0, L.t1, L.flag);
                else
                    r = agf_nan();
                break;
            default:  // SUM, BINS, hidden helpers
                r = s.a[l];
                break;
        }
    }
    if constexpr (GLC == 0)
        if (n_grp == 0) r = agf_nan();  // empty resample bin -> NaN for every reducer
    return round_to<T>(r);
}

// one level-1 group value into slot j (same arithmetic as the reference's second
// numba_resample pass over the per-group series)
__device__ __forceinline__ void l2_acc_one(const SlotP &S, double &b, double xt) {
    switch (S.calc) {
        case AGF_CALC_MEAN:
        case AGF_CALC_SUM:
            b += xt;
            break;
        case AGF_CALC_MIN:
            b = (xt < b || xt != xt) ? xt : b;  // NaN is sticky: any NaN -> NaN
            break;
        case AGF_CALC_MAX:
            b = (xt > b || xt != xt) ? xt : b;
            break;
        case AGF_CALC_DD:
            if (xt > S.t0 && xt < S.t1) b += fabs(xt - S.base);
            if (xt != xt) b = xt;
            break;
        case AGF_CALC_BINS:
            if (xt > S.t0 && xt < S.t1) b += 1.0;
            break;
        default:
            break;
    }
}

// merge of two partial records of the same slot (stripe order)
__device__ __forceinline__ double l2_merge(int calc, double acc, double part) {
    switch (calc) {
        case AGF_CALC_MIN:
            return (part < acc || part != part) ? part : acc;
        case AGF_CALC_MAX:
            return (part > acc || part != part) ? part : acc;
        default:
            return acc + part;
    }
}

template <typename T>
__device__ __forceinline__ void store_col(void *out, int out_f64, size_t idx, double v) {
    if (out_f64)
        reinterpret_cast<double *>(out)[idx] = v;
    else
        reinterpret_cast<float *>(out)[idx] = (float)v;
}

// typed slots: straight-line code, no transform switch
template <typename T>
__device__ __forceinline__ void l2_acc_sum(const SlotP &S, double &b, double x) {
    double r = x;
    if (S.ip != 1) {  // uniform
        r = (S.ip == 2) ? x * x : ipow(x, S.ip);
        if (!S.x_f64) r = round_to<T>(r);
    }
    b += r;  // a NaN group value poisons the sum == "any NaN -> NaN"
}
// x is a value of the raster dtype: for float the compare runs in fp32 against thresholds rounded
// outward (same truth value as the reference's fp64 compare, see set_thresholds in agf_k1_inst.cuh)
template <typename T>
__device__ __forceinline__ void l2_acc_bins(const SlotP &S, int &c, double x) {
    if constexpr (sizeof(T) == 4) {
        count_in_range(c, (float)x, S.flo, S.fhi);  // exact: x holds a float
    } else {
        count_in_range(c, x, S.t0, S.t1);
    }
}

// Where a single-level program's columns go.  X is cell-major (X[g, cell, :]), so the 32 cells of a warp
// own one contiguous block of 32 * out_ncols elements; written straight from registers that is
// out_ncols stores per thread with a stride of out_ncols elements between lanes (every store
// instruction touches 32 sectors: +8.5 ms on the C3b daily panel, ncu r1i).  With a per-warp staging
// area in shared memory the block is assembled there and copied out with fully coalesced stores.
// All 32 lanes of the warp must reach l1_flush when `stage` is set (cells past the grid are `!active`).
struct OutSink {
    void *stage;  // this warp's staging area (nullptr: direct stores)
    bool active;  // this thread's cell exists
};

// bytes of staging per consumer warp an instantiation reserves (0: never stages)
template <int NL, int NS>
__host__ __device__ constexpr int stage_bytes_per_warp() {
    return (NS == 0 && NL > 1) ? 32 * NL * 4 : 0;
}

template <typename T, int NL, int NS>
__device__ __forceinline__ void sink_put(const K1Params<T, NL, NS> &p, const OutSink &o, size_t base, int dst, double v) {
    if (o.stage != nullptr) {
        const int i = (threadIdx.x & 31) * p.out_ncols + dst;
        if (p.out_f64)
            reinterpret_cast<double *>(o.stage)[i] = v;
        else
            reinterpret_cast<float *>(o.stage)[i] = (float)v;
    } else if (o.active) {
        store_col<T>(p.out, p.out_f64, base + dst, v);
    }
}

// copy the staged block of this warp to X (32-bit words, 128 contiguous bytes per store instruction)
template <typename T, int NL, int NS>
__device__ __forceinline__ void sink_commit(const K1Params<T, NL, NS> &p, const OutSink &o, int g, int cell) {
    if (o.stage == nullptr) return;
    __syncwarp();
    const int lane = threadIdx.x & 31;
    const int cell0 = cell - lane;
    const int n_valid = min(32, p.n_cells - cell0);
    if (n_valid > 0) {
        const int wpe = p.out_f64 ? 2 : 1;  // 32-bit words per element
        const int words = n_valid * p.out_ncols * wpe;
        uint32_t *dst = reinterpret_cast<uint32_t *>(p.out) + ((size_t)g * p.n_cells + cell0) * p.out_ncols * wpe;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(o.stage);
        for (int i = lane; i < words; i += 32) dst[i] = src[i];
    }
    __syncwarp();
}

// end of level-1 group g (n_grp rows): emit columns (single-level) or feed the slots
template <typename T, int NL, int NS, bool DIAG, unsigned KINDS, int NB, int GLC = 0, typename ST>
__device__ __forceinline__ void l1_flush(const K1Params<T, NL, NS> &p, ST &s, int g, int n_grp, int cell,
                                         const OutSink &o) {
    if constexpr (ST::TL) {
        // typed lanes: column of kernel lane l is cols[l].dst (-1: inert pad, nothing stored)
        constexpr int NBL = NL - ST::NA;
        const bool empty = (GLC == 0) && n_grp == 0;  // empty resample bin -> NaN for every reducer
        bool ok = !empty;
        const size_t base = ((size_t)g * p.n_cells + cell) * p.out_ncols;  // X[g, cell, :]
#pragma unroll
        for (int j = 0; j < NBL; ++j) {
            const int dst = p.cols[j].dst;
            if (dst >= 0) sink_put(p, o, base, dst, empty ? agf_nan() : (double)s.cf[j]);
        }
#pragma unroll
        for (int l = 0; l < ST::NA; ++l) {
            const int dst = p.cols[NBL + l].dst;
            if (dst >= 0) {
                double r = (p.lanes[NBL + l].calc == AGF_CALC_MEAN) ? mean_of<T, GLC>(s.a[l], n_grp) : s.a[l];
                r = empty ? agf_nan() : round_to<T>(r);
                ok &= (r == r);
                sink_put(p, o, base, dst, r);
            }
        }
        if (o.active) {
            unsigned char *vp = p.valid + (size_t)g * p.n_cells + cell;
            *vp = (p.valid_and ? (*vp != 0) && ok : ok) ? 1 : 0;
        }
        sink_commit(p, o, g, cell);
        return;
    }
    double val[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) val[l] = (NL == 1 || l < p.n_lanes) ? l1_value<KINDS, GLC>(p, s, l, n_grp) : 0.0;

    if constexpr (NS == 0) {
        bool ok = true;
        const size_t base = ((size_t)g * p.n_cells + cell) * p.out_ncols;  // X[g, cell, :]
        if (DIAG) {  // column c == lane c, no transform
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                if (l < p.n_cols && p.cols[l].dst >= 0) {  // dst < 0: an inert pad lane of the mixed layout
                    ok &= (val[l] == val[l]);
                    sink_put(p, o, base, p.cols[l].dst, val[l]);
                }
            }
        } else {
            for (int c = 0; c < p.n_cols; ++c) {
                const ColP &C = p.cols[c];
                double x = apply_xform<T>(select_reg<NL>(val, C.src), C.xform, C.xparam, C.x_f64);
                ok &= (x == x);
                sink_put(p, o, base, C.dst, x);
            }
        }
        if (o.active) {
            unsigned char *vp = p.valid + (size_t)g * p.n_cells + cell;
            *vp = (p.valid_and ? (*vp != 0) && ok : ok) ? 1 : 0;
        }
        sink_commit(p, o, g, cell);
    } else if constexpr (NB >= 0) {
        // typed slots; unused ones are padded by the launcher with inert reducers whose registers
        // are never written back, so there is no per-slot bound check or branch
#pragma unroll
        for (int j = 0; j < NB; ++j)
            l2_acc_bins<T>(p.slots[j], s.c[j], DIAG ? val[j < NL ? j : 0] : select_reg<NL>(val, p.slots[j].src));
#pragma unroll
        for (int j = NB; j < NS; ++j)
            l2_acc_sum<T>(p.slots[j], s.b[j - NB], DIAG ? val[j < NL ? j : 0] : select_reg<NL>(val, p.slots[j].src));
    } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if (j < p.n_slots) {
                const SlotP &S = p.slots[j];
                double x = DIAG ? val[j < NL ? j : 0] : select_reg<NL>(val, S.src);
                double xt = apply_xform<T>(x, S.xform, S.xparam, S.x_f64);
                l2_acc_one(S, s.b[j], xt);
            }
        }
    }
}

// one partial record: slot values of a finished (stripe, level-2 group) intersection
template <int NB, typename T, int NL, int NS, typename ST>
__device__ __forceinline__ void l2_write_rec(const K1Params<T, NL, NS> &p, const ST &s, int rec, int cell, int g2) {
    if constexpr (NS > 0) {
        if (p.direct_out) {
            // the whole level-2 group g2 was reduced by this thread: do what agf_finalize would do
            const int n2 = p.b2[g2 + 1] - p.b2[g2];
            bool ok = true;
            for (int c = 0; c < p.n_cols; ++c) {
                const ColP &C = p.cols[c];
                const SlotP &S = p.slots[C.src];
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < NS; ++j) {  // slot registers are picked by a select chain, never indexed
                    double sj;
                    if constexpr (NB >= 0)
                        sj = (j < NB) ? (double)s.c[j < NB ? j : 0] : s.b[j >= NB ? j - NB : 0];
                    else
                        sj = s.b[j];
                    v = (j == C.src) ? sj : v;
                }
                if (S.calc == AGF_CALC_MEAN) v = v / (double)n2;
                const bool slot_f64 = S.x_f64 || p.in_f64;
                if (!slot_f64) v = (double)(float)v;  // stored in the dtype of the slot's input series
                if (C.xform != AGF_XF_NONE) v = slot_f64 ? apply_xform64(v, C.xform, C.xparam) : apply_xform<float>(v, C.xform, C.xparam, C.x_f64);
                ok &= (v == v);
                store_col<T>(p.out, p.out_f64, ((size_t)g2 * p.n_cells + cell) * p.out_ncols + C.dst, v);
            }
            unsigned char *vp = p.valid + (size_t)g2 * p.n_cells + cell;
            *vp = (p.valid_and ? (*vp != 0) && ok : ok) ? 1 : 0;
            return;
        }
        if constexpr (NB >= 0) {
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const int dst = p.slots[j].dst;
                if (dst >= 0)
                    p.partial[((size_t)rec * p.n_slots + dst) * p.n_cells + cell] =
                        (j < NB) ? (double)s.c[j < NB ? j : 0] : s.b[j >= NB ? j - NB : 0];
            }
        } else {
#pragma unroll
            for (int j = 0; j < NS; ++j)
                if (j < p.n_slots)
                    p.partial[((size_t)rec * p.n_slots + j) * p.n_cells + cell] = s.b[j];
        }
    }
}

// ------------------------------------------------------------------------------------------
// K1, direct-load variant: any shape / alignment.  Register double-buffered batches of K1_U
// rows that never cross a level-1 group boundary; next batch's loads are issued before the
// current batch is reduced.
// ------------------------------------------------------------------------------------------
template <typename T, int NL, int NS, bool DIAG, unsigned KINDS, int NB>
__global__ void __launch_bounds__(K1_THREADS)
    agf_k1_ldg(const __grid_constant__ K1Params<T, NL, NS> p) {
    const int cell = blockIdx.x * K1_THREADS + threadIdx.x;
    if (cell >= p.n_cells) return;
    const Stripe st = p.stripes[p.stripe0 + blockIdx.y];
    int g = st.g1_begin;
    const int g_end = st.g1_end;
    if (g >= g_end) return;
    int g2 = st.g2_first;
    int rec = st.rec0;
    int next_b2 = (NS > 0) ? p.b2[g2 + 1] : 0;

    CellState<T, NL, NS, NB> s;
    l1_init<KINDS>(p, s);
    l2_init<NB>(p, s);

    const T *xc = p.x + cell;
    int glo = p.b1[g];      // first row of the current group
    int nb = p.b1[g + 1];   // one past its last row
    int ck = glo;
    int clen = min(K1_U, nb - ck);
    T cur[K1_U], nxt[K1_U];
#pragma unroll
    for (int i = 0; i < K1_U; ++i)
        cur[i] = (i < clen) ? __ldg(xc + (size_t)(ck + i - p.row0) * p.ld) : T(0);

    while (g < g_end) {
        // descriptor of the next batch
        const int nk = ck + clen;
        const bool ends = (nk == nb);
        int ng = g, nnb = nb;
        if (ends) {
            ng = g + 1;
            nnb = (ng < g_end) ? p.b1[ng + 1] : nk;
        }
        const int nlen = (ng < g_end) ? min(K1_U, nnb - nk) : 0;
#pragma unroll
        for (int i = 0; i < K1_U; ++i)
            nxt[i] = (i < nlen) ? __ldg(xc + (size_t)(nk + i - p.row0) * p.ld) : T(0);

        // reduce the current batch in time order
        pre_apply_batch(p, cur);
#pragma unroll
        for (int i = 0; i < K1_U; ++i)
            if (i < clen) l1_acc<KINDS>(p, s, cur[i]);

        if (ends) {
            l1_flush<T, NL, NS, DIAG, KINDS, NB>(p, s, g, nb - glo, cell, OutSink{nullptr, true});
            l1_init<KINDS>(p, s);
            if (NS > 0) {
                // close every level-2 group that ends with level-1 group g
                if (g + 1 == next_b2 || g + 1 == g_end) {
                    l2_write_rec<NB>(p, s, rec, cell, g2);
                    l2_init<NB>(p, s);
                    ++rec;
                    if (g + 1 < g_end && g + 1 == next_b2) {
                        do {  // skip zero-width level-2 groups (they get no record -> NaN)
                            ++g2;
                            next_b2 = p.b2[g2 + 1];
                        } while (next_b2 == g + 1);
                    }
                }
            }
            glo = nb;
        }
#pragma unroll
        for (int i = 0; i < K1_U; ++i) cur[i] = nxt[i];
        ck = nk;
        clen = nlen;
        g = ng;
        nb = nnb;
    }
}


// ------------------------------------------------------------------------------------------
// K1, TMA variant (sm_100a): a producer warp streams [TT rows x 256 cells] tiles of the raster
// into a shared-memory ring with cp.async.bulk.tensor (one elected thread, mbarrier
// complete_tx), 8 consumer warps reduce straight out of shared memory (conflict-free: thread
// t reads column t of every row).  Bytes in flight are set by the ring (STAGES x 24 KB per CTA,
// two CTAs per SM), not by registers, which is what the HBM roofline needs.  Needs a 16-byte
// aligned raster with a row stride that is a multiple of 16 bytes; otherwise the library runs
// agf_k1_ldg.
// ------------------------------------------------------------------------------------------
constexpr int TMA_CW = 256;         // cells per tile row == consumer threads
constexpr int TMA_TILE_BYTES_DEFAULT = 24 * 1024;
constexpr int TMA_STAGES_DEFAULT = 4;
constexpr int TMA_THREADS = TMA_CW + 32;

template <typename T>
__host__ __device__ constexpr int tma_rows() {
    return TMA_TILE_BYTES_DEFAULT / (TMA_CW * (int)sizeof(T));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// The producer lane's wait for a free stage.  try_wait returns after a few cycles when the phase has not completed
// (ncu r2d: ~900 polls per wait, YIELD + TRYWAIT + BRA each -- a quarter of all warp instructions the daily-panel
// kernel issued came from its three producer warps per SM spinning, on an SM whose issue slots are the scarce
// resource).  A stage frees up once per tile, microseconds apart: sleep between polls.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    while (!done) {
        __nanosleep(256);
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// (An L2 evict-first cache hint on these loads was measured: C3 5.40 ms against 5.23 ms without it, the small
// CONUS grid 0.130 against 0.133 ms -- not kept.)
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

struct alignas(64) TensorMap {  // same layout as CUtensorMap (128 opaque bytes)
    unsigned long long opaque[16];
};

template <typename T, int NL, int NS, bool DIAG, unsigned KINDS, int NB, int TT = tma_rows<T>(),
          int TMA_STAGES = TMA_STAGES_DEFAULT, int MINB = 2>
__global__ void __launch_bounds__(TMA_THREADS, MINB)
    agf_k1_tma(const __grid_constant__ K1Params<T, NL, NS> p, const __grid_constant__ TensorMap tmap) {
    constexpr int TMA_TILE_BYTES = TT * TMA_CW * (int)sizeof(T);
    // Typed lanes with many bins (bins of the raw values by month / year): runs are scanned a whole tile (then eight rows)
    // at a time through l1_acc_group, which tests only the bins inside the warp's min / max of the batch instead of every
    // bin for every value (13 bins: 39 instructions per value otherwise; C3-size hourly bins by year 16.9 -> 10.8 ms with
    // batches of eight).
    constexpr bool GROUPED = typed_lanes<NS, NB>() && NB > 4;
    constexpr int UNROLL = GROUPED ? 24 : (NL <= 4 ? 8 : (NL <= 16 ? 2 : 1));
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *tiles = reinterpret_cast<T *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + TMA_STAGES * TMA_TILE_BYTES);
    uint64_t *empty = full + TMA_STAGES;

    const Stripe st = p.stripes[p.stripe0 + blockIdx.y];
    int g = st.g1_begin;
    const int g_end = st.g1_end;
    if (g >= g_end) return;  // uniform
    const int k_begin = p.b1[g];
    const int k_end = p.b1[g_end];
    const int n_tiles = (k_end - k_begin + TT - 1) / TT;
    const int cell0 = blockIdx.x * TMA_CW;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], TMA_CW / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x >= TMA_CW) {
        // ===== producer warp: one elected lane issues every tile load =====
        if (threadIdx.x == TMA_CW) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % TMA_STAGES;
                if (i >= TMA_STAGES) mbar_wait_backoff(&empty[s], ((i / TMA_STAGES) - 1) & 1);
                mbar_expect_tx(&full[s], TMA_TILE_BYTES);
                tma_load_2d(smem_raw + s * TMA_TILE_BYTES, &tmap, cell0, (int)(k_begin + i * TT - p.row0), &full[s]);
            }
        }
        return;
    }

    // ===== consumers: thread t owns cell cell0 + t =====
    const int cell = cell0 + threadIdx.x;
    const bool active = cell < p.n_cells;
    // per-warp output staging lives behind the ring and its barriers (see OutSink)
    void *stage = nullptr;
    if constexpr (stage_bytes_per_warp<NL, NS>() > 0)
        if (p.stage_out)
            stage = smem_raw + TMA_STAGES * TMA_TILE_BYTES + 2 * TMA_STAGES * 8 +
                    (threadIdx.x >> 5) * stage_bytes_per_warp<NL, NS>();
    int g2 = st.g2_first;
    int rec = st.rec0;
    int next_b2 = (NS > 0) ? p.b2[g2 + 1] : 0;
    CellState<T, NL, NS, NB> s;
    l1_init<KINDS>(p, s);
    l2_init<NB>(p, s);
    int k = k_begin;
    int glo = k_begin;
    int nb = p.b1[g + 1];
    // bound after next, fetched one group ahead so the group-end path never waits on the load
    int nb_next = (g + 1 < g_end) ? p.b1[g + 2] : 0x7fffffff;

    // level-1 group g is complete (k == nb): flush it, close level-2 groups that end with it, move on
    auto close_group = [&]() {
        if (NS == 0 || active) l1_flush<T, NL, NS, DIAG, KINDS, NB>(p, s, g, nb - glo, cell, OutSink{stage, active});
        l1_init<KINDS>(p, s);
        if constexpr (NS > 0) {
            if (g + 1 == next_b2 || g + 1 == g_end) {
                if (active) l2_write_rec<NB>(p, s, rec, cell, g2);
                l2_init<NB>(p, s);
                ++rec;
                if (g + 1 < g_end && g + 1 == next_b2) {
                    do {  // skip zero-width level-2 groups (no record -> NaN in finalize)
                        ++g2;
                        next_b2 = p.b2[g2 + 1];
                    } while (next_b2 == g + 1);
                }
            }
        }
        glo = nb;
        ++g;
        nb = nb_next;  // 0x7fffffff after the last group: k == nb is never true again
        nb_next = (g + 1 < g_end) ? p.b1[g + 2] : 0x7fffffff;
    };

    // (A register-tile variant of this loop -- pull the whole column, release the stage early, then walk
    // the rows with a group-end test per row -- was tried for the month-bounded daily kernel and was
    // 2x SLOWER: 24 inlined group-end sites doubled the instruction count, ncu r1p.)
    for (int i = 0; i < n_tiles; ++i) {
        const int stg = i % TMA_STAGES;
        mbar_wait(&full[stg], (i / TMA_STAGES) & 1);
        const T *col = tiles + (size_t)stg * (TMA_TILE_BYTES / sizeof(T)) + threadIdx.x;
        const int rows = min(TT, k_end - (k_begin + i * TT));
        int r = 0;
        while (r < rows || (k == nb && g < g_end)) {
            const int run = min(rows - r, nb - k);
            const T *cp = col + r * TMA_CW;
            int j = run;
#pragma unroll 1
            for (; j >= UNROLL; j -= UNROLL, cp += UNROLL * TMA_CW) {
                T v[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) v[u] = cp[u * TMA_CW];
                pre_apply_batch(p, v);
                if constexpr (GROUPED) {
                    l1_acc_group<KINDS>(p, s, v);  // run lengths are the same for every lane: the whole warp is here
                } else {
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) l1_acc<KINDS>(p, s, v[u]);
                }
            }
            // tail of the run: one optional batch of 4, of 2, of 1 (a per-value loop cost ~12 instructions
            // of overhead per value, and month-bounded runs cut by 24-row tiles are mostly tail)
            if constexpr (GROUPED && UNROLL > 8) {
#pragma unroll 1
                for (; j >= 8; j -= 8, cp += 8 * TMA_CW) {
                    T v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = cp[u * TMA_CW];
                    pre_apply_batch(p, v);
                    l1_acc_group<KINDS>(p, s, v);
                }
            }
            if constexpr (UNROLL >= 8) {
                if (j >= 4) {
                    T v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = cp[u * TMA_CW];
                    pre_apply_batch(p, v);
#pragma unroll
                    for (int u = 0; u < 4; ++u) l1_acc<KINDS>(p, s, v[u]);
                    j -= 4;
                    cp += 4 * TMA_CW;
                }
                if (j >= 2) {
                    T v[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) v[u] = cp[u * TMA_CW];
                    pre_apply_batch(p, v);
#pragma unroll
                    for (int u = 0; u < 2; ++u) l1_acc<KINDS>(p, s, v[u]);
                    j -= 2;
                    cp += 2 * TMA_CW;
                }
            }
#pragma unroll 1
            for (; j > 0; --j, cp += TMA_CW) l1_acc<KINDS>(p, s, p.n_pre ? pre_apply(p, *cp) : *cp);
            r += run;
            k += run;
            if (k == nb) close_group();
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stg]);
    }
}

// ------------------------------------------------------------------------------------------
// K1, TMA variant for uniform level-1 groups: every group of the launched stripes has exactly
// GL rows and GL divides the tile height TT (hourly -> date: GL = TT = 24; daily data grouped by
// date: GL = 1).  Same ring as agf_k1_tma; what goes away is the per-group bookkeeping of the
// general kernel (run lengths, bound loads, min/compare chains: ~110 of its ~440 instructions per
// warp-day in ncu r1c), the loop around the scan (fully unrolled: GL loads, GL converts, GL adds)
// and the division subroutine (mean_of with a compile-time count).
// ------------------------------------------------------------------------------------------
template <typename T, int NL, int NS, bool DIAG, unsigned KINDS, int NB, int GL, int TT, int TMA_STAGES, int MINB>
__global__ void __launch_bounds__(TMA_THREADS, MINB)
    agf_k1_tma_uni(const __grid_constant__ K1Params<T, NL, NS> p, const __grid_constant__ TensorMap tmap) {
    static_assert(TT % GL == 0, "group length must divide the tile height");
    constexpr int TMA_TILE_BYTES = TT * TMA_CW * (int)sizeof(T);
    constexpr int GPT = TT / GL;  // groups per tile
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *tiles = reinterpret_cast<T *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + TMA_STAGES * TMA_TILE_BYTES);
    uint64_t *empty = full + TMA_STAGES;

    const Stripe st = p.stripes[p.stripe0 + blockIdx.y];
    int g = st.g1_begin;
    const int g_end = st.g1_end;
    if (g >= g_end) return;  // uniform
    const int k_begin = p.b1[g];
    const int n_groups = g_end - g;
    const int n_tiles = (n_groups + GPT - 1) / GPT;
    const int cell0 = blockIdx.x * TMA_CW;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], TMA_CW / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x >= TMA_CW) {
        // ===== producer warp: one elected lane issues every tile load =====
        if (threadIdx.x == TMA_CW) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            int s = 0, ph = 0;
            for (int i = 0; i < n_tiles; ++i) {
                if (i >= TMA_STAGES) mbar_wait_backoff(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], TMA_TILE_BYTES);
                tma_load_2d(smem_raw + s * TMA_TILE_BYTES, &tmap, cell0, (int)(k_begin + i * TT - p.row0), &full[s]);
                if (++s == TMA_STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        return;
    }

    // ===== consumers: thread t owns cell cell0 + t =====
    const int cell = cell0 + threadIdx.x;
    const bool active = cell < p.n_cells;
    // per-warp output staging lives behind the ring and its barriers (see OutSink)
    void *stage = nullptr;
    if constexpr (stage_bytes_per_warp<NL, NS>() > 0)
        if (p.stage_out)
            stage = smem_raw + TMA_STAGES * TMA_TILE_BYTES + 2 * TMA_STAGES * 8 +
                    (threadIdx.x >> 5) * stage_bytes_per_warp<NL, NS>();
    int g2 = st.g2_first;
    int rec = st.rec0;
    int next_b2 = (NS > 0) ? p.b2[g2 + 1] : 0;
    CellState<T, NL, NS, NB> s;
    l1_init<KINDS>(p, s);
    l2_init<NB>(p, s);

    // end of level-1 group g: flush, close level-2 groups that end with it
    auto group_end = [&]() {
        if (NS == 0 || active) l1_flush<T, NL, NS, DIAG, KINDS, NB, GL>(p, s, g, GL, cell, OutSink{stage, active});
        l1_init<KINDS>(p, s);
        if constexpr (NS > 0) {
            if (g + 1 == next_b2 || g + 1 == g_end) {
                if (active) l2_write_rec<NB>(p, s, rec, cell, g2);
                l2_init<NB>(p, s);
                ++rec;
                if (g + 1 < g_end && g + 1 == next_b2) {
                    do {  // skip zero-width level-2 groups (no record -> NaN in finalize)
                        ++g2;
                        next_b2 = p.b2[g2 + 1];
                    } while (next_b2 == g + 1);
                }
            }
        }
        ++g;
    };

    int stg = 0, ph = 0;
#pragma unroll 1
    for (int i = 0; i < n_tiles; ++i) {
        mbar_wait(&full[stg], ph);
        const T *col = tiles + (size_t)stg * (TMA_TILE_BYTES / sizeof(T)) + threadIdx.x;
        if constexpr (GPT == 1) {
            // one group per tile: pull the column into registers, reduce it, hand the stage back, THEN
            // flush.  The stage must not be released right after the loads are ISSUED: the arrive does
            // not wait for the shared-memory reads to return (no register dependence), and once all 8
            // warps have arrived the producer's next TMA may overwrite rows whose reads are still in
            // flight.  That showed as a few cells per launch with a wrong daily value (non-repeatable
            // sums, exact bins) in one instantiation.  The reduction consumes every loaded register,
            // so by its end the reads have completed; only the group-end flush stays overlapped.
            if constexpr (CellState<T, NL, NS, NB>::TL || TT % 2 != 0) {
                T v[TT];  // typed lanes cull their bins on the min / max of the whole group
#pragma unroll
                for (int r = 0; r < TT; ++r) v[r] = col[r * TMA_CW];
                pre_apply_batch(p, v);
                l1_acc_group<KINDS>(p, s, v);
            } else {
                // two half-columns: 12 value registers live instead of 24 (the 20-slot kernel sits at the
                // 72-register cap of three CTAs per SM)
                constexpr int H = TT / 2;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    T v[H];
#pragma unroll
                    for (int r = 0; r < H; ++r) v[r] = col[(h * H + r) * TMA_CW];
                    pre_apply_batch(p, v);
                    l1_acc_group<KINDS>(p, s, v);
                }
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stg]);
            group_end();
        } else {
            const int ng = min(GPT, g_end - g);
#pragma unroll 1
            for (int gi = 0; gi < ng; ++gi) {
                T v[GL];
#pragma unroll
                for (int r = 0; r < GL; ++r) v[r] = col[(gi * GL + r) * TMA_CW];
                pre_apply_batch(p, v);
                l1_acc_group<KINDS>(p, s, v);
                group_end();
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stg]);
        }
        if (++stg == TMA_STAGES) {
            stg = 0;
            ph ^= 1;
        }
    }
}

}  // namespace agf
