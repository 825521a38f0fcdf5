// agf_api.cu -- the C-ABI of libaggfly_b200.so (see include/aggfly_b200.h): descriptor
// validation, stripe / partial-record planning, kernel dispatch.  No CPU compute path.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda.h>

#include "agf_host.h"
#include "agf_post.cuh"

using namespace agf;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

int agf_cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return (int)e;
}

int agf_fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn();

extern "C" int agf_version(void) { return AGF_ABI_VERSION; }
extern "C" const char *agf_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------
struct agf_csr {
    int32_t n_regions;
    int64_t n_cells, nnz;
    const int32_t *d_row_ptr;
    const int32_t *d_cell_idx;
    const double *d_w;
    int device;
};

// ------------------------------------------------------------------------------------------
// descriptor validation + planning (host only)
// ------------------------------------------------------------------------------------------
static bool is_l1_calc(int c) { return c >= AGF_CALC_MEAN && c <= AGF_CALC_DD_R; }
static bool is_l2_calc(int c) {
    return c == AGF_CALC_MEAN || c == AGF_CALC_SUM || c == AGF_CALC_MIN || c == AGF_CALC_MAX ||
           c == AGF_CALC_DD || c == AGF_CALC_BINS;
}
static bool is_xform(int x) { return x >= AGF_XF_NONE && x <= AGF_XF_SPLINE2; }

static int validate_desc(const agf_program_desc_t *d, int64_t n_cells) {
    if (!d) return fail(AGF_E_INVALID, "null descriptor");
    if (n_cells <= 0 || n_cells > 0x7fffffff) return fail(AGF_E_INVALID, "n_cells out of range");
    if (d->in_dtype != AGF_F32 && d->in_dtype != AGF_F64) return fail(AGF_E_INVALID, "bad in_dtype");
    if (d->out_dtype != AGF_F32 && d->out_dtype != AGF_F64) return fail(AGF_E_INVALID, "bad out_dtype");
    if (d->in_dtype == AGF_F64 && d->out_dtype != AGF_F64)
        return fail(AGF_E_INVALID, "float64 raster needs float64 columns");
    if (d->n_lanes < 1 || d->n_lanes > AGF_MAX_LANES) return fail(AGF_E_INVALID, "n_lanes out of range");
    if (d->n_slots < 0 || d->n_slots > AGF_MAX_SLOTS) return fail(AGF_E_INVALID, "n_slots out of range");
    if (d->n_cols < 1 || d->n_cols > AGF_MAX_COLS) return fail(AGF_E_INVALID, "n_cols out of range");
    if (d->n_time <= 0 || d->n_time > 0x7fffffff) return fail(AGF_E_INVALID, "n_time out of range");
    if (d->n_groups1 < 1 || !d->bounds1) return fail(AGF_E_INVALID, "level-1 bounds missing");
    for (int64_t g = 0; g < d->n_groups1; ++g)
        if (d->bounds1[g] > d->bounds1[g + 1]) return fail(AGF_E_INVALID, "bounds1 not monotonic");
    if (d->bounds1[0] < 0 || d->bounds1[d->n_groups1] > d->n_time)
        return fail(AGF_E_INVALID, "bounds1 outside the time axis");
    if (d->n_pre < 0 || d->n_pre > AGF_MAX_PRE) return fail(AGF_E_INVALID, "n_pre out of range");
    for (int i = 0; i < d->n_pre; ++i)
        if (d->pre[i].op < AGF_PRE_ADD || d->pre[i].op > AGF_PRE_NEG) return fail(AGF_E_INVALID, "pre %d: bad op", i);
    bool sine = false;
    for (int l = 0; l < d->n_lanes; ++l) {
        if (!is_l1_calc(d->lanes[l].calc)) return fail(AGF_E_INVALID, "lane %d: bad calc", l);
        sine |= d->lanes[l].calc == AGF_CALC_SINE_DD;
    }
    if (sine) {
        if (d->n_lanes < 4 || d->lanes[0].calc != AGF_CALC_HIDDEN_SUM ||
            d->lanes[1].calc != AGF_CALC_HIDDEN_MIN || d->lanes[2].calc != AGF_CALC_HIDDEN_MAX)
            return fail(AGF_E_INVALID, "sine_dd lanes need hidden sum/min/max helpers in lanes 0..2");
    }
    if (d->n_slots > 0) {
        if (d->n_groups2 < 1 || !d->bounds2) return fail(AGF_E_INVALID, "level-2 bounds missing");
        if (d->bounds2[0] != 0 || d->bounds2[d->n_groups2] != d->n_groups1)
            return fail(AGF_E_INVALID, "bounds2 must cover the level-1 group axis");
        for (int64_t g = 0; g < d->n_groups2; ++g)
            if (d->bounds2[g] > d->bounds2[g + 1]) return fail(AGF_E_INVALID, "bounds2 not monotonic");
        for (int j = 0; j < d->n_slots; ++j) {
            const agf_slot_t &s = d->slots[j];
            if (s.src < 0 || s.src >= d->n_lanes) return fail(AGF_E_INVALID, "slot %d: bad src", j);
            if (!is_l2_calc(s.calc)) return fail(AGF_E_UNSUPPORTED, "slot %d: calc not fusable", j);
            if (!is_xform(s.xform)) return fail(AGF_E_INVALID, "slot %d: bad xform", j);
        }
    }
    const int nsrc = d->n_slots > 0 ? d->n_slots : d->n_lanes;
    for (int c = 0; c < d->n_cols; ++c) {
        const agf_col_t &k = d->cols[c];
        if (k.src < 0 || k.src >= nsrc) return fail(AGF_E_INVALID, "col %d: bad src", c);
        if (!is_xform(k.xform)) return fail(AGF_E_INVALID, "col %d: bad xform", c);
        if (d->n_slots == 0 && d->lanes[k.src].calc >= AGF_CALC_HIDDEN_SUM && d->lanes[k.src].calc <= AGF_CALC_HIDDEN_MAX)
            return fail(AGF_E_INVALID, "col %d reads a hidden lane", c);
        if (k.dst < 0) return fail(AGF_E_INVALID, "col %d: negative dst", c);
        if (k.x_f64 && d->out_dtype != AGF_F64)
            return fail(AGF_E_INVALID, "col %d is float64 but out_dtype is float32", c);
    }
    return 0;
}

// lane kinds + 1:1 mapping flags of a descriptor
static void analyse_desc(const agf_program_desc_t *d, unsigned *kinds, unsigned *slot_kinds, int *n_bin_slots,
                         int *diag_ok) {
    unsigned k = 0, sk = 0;
    int nbins = 0;
    for (int l = 0; l < d->n_lanes; ++l) k |= kind_of_calc(d->lanes[l].calc);
    for (int j = 0; j < d->n_slots; ++j) {
        const unsigned kind = slot_kind_of(d->slots[j].calc, d->slots[j].xform);
        sk |= kind;
        nbins += kind == SK_BINS;
    }
    *kinds = k;
    *slot_kinds = sk;
    *n_bin_slots = nbins;
    bool dg;
    if (d->n_slots == 0) {
        dg = d->n_cols <= d->n_lanes;
        for (int c = 0; c < d->n_cols && dg; ++c) dg = d->cols[c].src == c && d->cols[c].xform == AGF_XF_NONE;
    } else {
        dg = d->n_slots <= d->n_lanes;
        for (int j = 0; j < d->n_slots && dg; ++j) dg = d->slots[j].src == j;
    }
    *diag_ok = dg ? 1 : 0;
}

// first-fit over the instantiation tables of the K1 units; returns 0 or AGF_E_UNSUPPORTED
static int k1_select(const K1Launch &a, int mode, K1Choice *choice, int *rc) {
    const bool f32 = a.p->desc.in_dtype == AGF_F32;
    int miss;
    if (a.use_tma) {
        if (f32) {
            if (a.p->uniform_gl > 0 && !agf_k1_f32_tma_uni(a, mode, choice, rc)) return 0;
            miss = a.p->desc.n_slots > 0 ? agf_k1_f32_tma_two(a, mode, choice, rc) : agf_k1_f32_tma_single(a, mode, choice, rc);
        } else {
            miss = agf_k1_f64_tma(a, mode, choice, rc);
        }
    } else {
        miss = f32 ? agf_k1_f32_ldg(a, mode, choice, rc) : agf_k1_f64_ldg(a, mode, choice, rc);
    }
    if (miss)
        return fail(AGF_E_UNSUPPORTED,
                    "program with %d lanes / %d slots has no fused instantiation (split it: <=4 lanes feeding "
                    "<=32 slots, or <=16 lanes feeding one slot each; <=32 lanes single-level)",
                    a.p->desc.n_lanes, a.p->desc.n_slots);
    return 0;
}

static int uniform_group_rows(const agf_program *p) {
    const size_t G = p->b1.size() - 1;
    const int gl = p->b1[1] - p->b1[0];
    for (size_t g = 1; g < G; ++g)
        if (p->b1[g + 1] - p->b1[g] != gl) return 0;
    return gl > 0 ? gl : 0;
}

static int choose_kernel(agf_program *p) {
    p->uniform_gl = p->b1.size() >= 2 ? uniform_group_rows(p) : 0;
    p->max_group_rows = 0;
    for (size_t g = 0; g + 1 < p->b1.size(); ++g) p->max_group_rows = std::max(p->max_group_rows, p->b1[g + 1] - p->b1[g]);
    analyse_desc(&p->desc, &p->kinds, &p->slot_kinds, &p->n_bin_slots, &p->diag_ok);
    K1Launch q{};
    q.p = p;
    q.use_tma = 1;
    K1Choice ch{};
    int rc = 0;
    int r = k1_select(q, 1, &ch, &rc);
    if (r) return r;
    p->kernel_lanes = ch.lanes;
    p->kernel_slots = ch.slots;
    p->kernel_diag = ch.diag;
    p->kernel_kinds = ch.kinds;
    return 0;
}

struct Plan {
    std::vector<Stripe> stripes;
    std::vector<int32_t> g2_rec_ptr, g2_rec_idx;
    int n_recs = 0;
};

static void make_plan(const agf_program_desc_t *d, int64_t n_cells, int target_stripes, int sm_count,
                      Plan &plan) {
    const int64_t G1 = d->n_groups1;
    const int32_t *b1 = d->bounds1;
    const int64_t rows = (int64_t)b1[G1] - b1[0];
    int64_t S = target_stripes;
    if (S <= 0) {
        const int64_t cell_blocks = (n_cells + K1_THREADS - 1) / K1_THREADS;
        const int64_t want_ctas = 16LL * (sm_count > 0 ? sm_count : 148);
        S = (want_ctas + cell_blocks - 1) / cell_blocks;
        S = std::min<int64_t>(S, std::max<int64_t>(1, rows / 96));  // keep stripes >= ~96 rows
        // target_stripes < 0: "the heuristic, but at least -target_stripes" (streamed rasters want
        // stripes no longer than a few host->device chunks so kernels start before the copy ends)
        S = std::max<int64_t>(S, -(int64_t)target_stripes);
    }
    S = std::max<int64_t>(1, std::min<int64_t>(S, std::min<int64_t>(G1, 65535)));

    // cut at the level-1 group boundary closest to equal row shares
    std::vector<int64_t> cuts;  // group indices
    cuts.push_back(0);
    for (int64_t s = 1; s < S; ++s) {
        const int64_t want_row = b1[0] + rows * s / S;
        int64_t g = std::lower_bound(b1, b1 + G1 + 1, (int32_t)want_row) - b1;
        g = std::min<int64_t>(std::max<int64_t>(g, cuts.back()), G1);
        if (g > cuts.back() && g < G1) cuts.push_back(g);
    }
    cuts.push_back(G1);

    plan.stripes.clear();
    plan.n_recs = 0;
    const bool two = d->n_slots > 0;
    const int64_t G2 = two ? d->n_groups2 : 0;
    std::vector<std::vector<int32_t>> per_g2(two ? G2 : 0);
    int64_t g2 = 0;
    for (size_t i = 0; i + 1 < cuts.size(); ++i) {
        Stripe st;
        st.g1_begin = (int)cuts[i];
        st.g1_end = (int)cuts[i + 1];
        st.g2_first = 0;
        st.rec0 = plan.n_recs;
        if (two) {
            const int32_t *b2 = d->bounds2;
            // level-2 group containing level-1 group g1_begin (skip zero-width groups)
            while (!(b2[g2] <= st.g1_begin && st.g1_begin < b2[g2 + 1])) ++g2;
            st.g2_first = (int)g2;
            // one record per non-empty intersection, in order (mirrors the kernel loop)
            int64_t h = g2;
            while (h < G2 && b2[h] < st.g1_end) {
                if (b2[h + 1] > b2[h] && b2[h + 1] > st.g1_begin) {
                    per_g2[h].push_back(plan.n_recs);
                    ++plan.n_recs;
                }
                ++h;
            }
        }
        plan.stripes.push_back(st);
    }
    plan.g2_rec_ptr.assign(1, 0);
    plan.g2_rec_idx.clear();
    for (int64_t h = 0; h < G2; ++h) {
        for (int32_t r : per_g2[h]) plan.g2_rec_idx.push_back(r);
        plan.g2_rec_ptr.push_back((int32_t)plan.g2_rec_idx.size());
    }
}

// Host-only planning entry (used by agf_program_create and by the CPU tests of the planner):
// fills stripes_out with (g1_begin, g1_end, g2_first, rec0) quadruples.
extern "C" int agf_program_plan(const agf_program_desc_t *desc, int64_t n_cells, int32_t target_stripes,
                                int32_t sm_count, int32_t *stripes_out, int32_t max_stripes,
                                int32_t *n_stripes, int32_t *n_recs, int32_t *kernel_lanes,
                                int32_t *kernel_slots, int32_t *kernel_diag) {
    int rc = validate_desc(desc, n_cells);
    if (rc) return rc;
    agf_program tmp;
    tmp.desc = *desc;
    tmp.b1.assign(desc->bounds1, desc->bounds1 + desc->n_groups1 + 1);
    rc = choose_kernel(&tmp);
    if (rc) return rc;
    const int kl = tmp.kernel_lanes, ks = tmp.kernel_slots, dg = tmp.kernel_diag;
    Plan plan;
    make_plan(desc, n_cells, target_stripes, sm_count, plan);
    if (n_stripes) *n_stripes = (int32_t)plan.stripes.size();
    if (n_recs) *n_recs = plan.n_recs;
    if (kernel_lanes) *kernel_lanes = kl;
    if (kernel_slots) *kernel_slots = ks;
    if (kernel_diag) *kernel_diag = dg;
    if (stripes_out) {
        if ((int64_t)plan.stripes.size() > max_stripes) return fail(AGF_E_INVALID, "stripes_out too small");
        for (size_t i = 0; i < plan.stripes.size(); ++i) {
            stripes_out[4 * i + 0] = plan.stripes[i].g1_begin;
            stripes_out[4 * i + 1] = plan.stripes[i].g1_end;
            stripes_out[4 * i + 2] = plan.stripes[i].g2_first;
            stripes_out[4 * i + 3] = plan.stripes[i].rec0;
        }
    }
    return 0;
}

// The small device tables of a handle come from the device's stream-ordered memory pool, which is told to
// keep what is freed: a yearly loop creates and destroys programs every call, and cudaFree -- a device-wide
// synchronisation that also unmaps -- took 0.2-1.5 s now and then next to a 36 GB raster and its pinned host
// copy (bench e2e: runner.close() accounted for every slow call).
static int table_pool_ready(int dev) {
    static bool done[64] = {false};
    if (dev < 0 || dev >= 64 || done[dev]) return 0;
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long keep = ~0ULL;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    done[dev] = true;
    return 0;
}

template <typename V>
static int upload(V **dst, const std::vector<V> &src) {
    *dst = nullptr;
    if (src.empty()) return 0;
    CU(cudaMallocAsync((void **)dst, src.size() * sizeof(V), (cudaStream_t)0));
    CU(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(V), cudaMemcpyHostToDevice, (cudaStream_t)0));
    return 0;
}

static void release_table(void *p) {
    if (p) cudaFreeAsync(p, (cudaStream_t)0);
}

extern "C" int agf_program_destroy(agf_program_t *p) {
    if (!p) return 0;
    // the caller has waited for the launches that used this handle (see the header)
    release_table(p->d_b1);
    release_table(p->d_b2);
    release_table(p->d_g2_rec_ptr);
    release_table(p->d_g2_rec_idx);
    release_table(p->d_stripes);
    delete p;
    return 0;
}

extern "C" int agf_program_create(agf_program_t **out, const agf_program_desc_t *desc, int64_t n_cells,
                                  int32_t target_stripes) {
    if (!out) return fail(AGF_E_INVALID, "null out");
    *out = nullptr;
    int rc = validate_desc(desc, n_cells);
    if (rc) return rc;
    int dev = -1, sms = 148;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    agf_program *p = new agf_program();
    p->desc = *desc;
    p->b1.assign(desc->bounds1, desc->bounds1 + desc->n_groups1 + 1);
    if (desc->n_slots > 0) p->b2.assign(desc->bounds2, desc->bounds2 + desc->n_groups2 + 1);
    p->desc.bounds1 = nullptr;
    p->desc.bounds2 = nullptr;
    p->n_cells = n_cells;
    p->device = dev;
    if ((rc = choose_kernel(p))) {
        delete p;
        return rc;
    }
    for (int l = 0; l < desc->n_lanes; ++l) {
        const int c = desc->lanes[l].calc;
        if (c == AGF_CALC_MIN || c == AGF_CALC_MAX || c == AGF_CALC_DD || c == AGF_CALC_DD_R || c == AGF_CALC_SINE_DD)
            p->need_nan = 1;
        if (c == AGF_CALC_NANMEAN || c == AGF_CALC_SINE_DD) p->need_cnt = 1;
        if (c == AGF_CALC_SINE_DD) p->has_sine = 1;
    }
    Plan plan;
    make_plan(desc, n_cells, target_stripes, sms, plan);
    p->stripes = plan.stripes;
    p->g2_rec_ptr = plan.g2_rec_ptr;
    p->g2_rec_idx = plan.g2_rec_idx;
    p->n_recs = plan.n_recs;

    if (desc->n_slots > 0 && p->stripes.size() == 1) {
        p->direct_out = 1;  // nothing to merge: the kernel finalizes (see K1Params::direct_out)
        for (size_t g = 0; g + 1 < p->b2.size(); ++g)
            if (p->b2[g + 1] == p->b2[g]) p->direct_out = 0;  // an empty outer period gets its NaN from agf_finalize
    }
    if ((rc = table_pool_ready(dev)) || (rc = upload(&p->d_b1, p->b1)) || (rc = upload(&p->d_b2, p->b2)) ||
        (rc = upload(&p->d_stripes, p->stripes)) || (rc = upload(&p->d_g2_rec_ptr, p->g2_rec_ptr)) ||
        (rc = upload(&p->d_g2_rec_idx, p->g2_rec_idx))) {
        agf_program_destroy(p);
        return rc;
    }
    // the tables must be resident before a launch on ANY stream reads them (the copies ran on stream 0)
    cudaError_t e_sync = cudaStreamSynchronize((cudaStream_t)0);
    if (e_sync != cudaSuccess) {
        agf_program_destroy(p);
        return agf_cuda_fail(e_sync, "cudaStreamSynchronize");
    }
    *out = p;
    return 0;
}

static int64_t out_groups(const agf_program *p) {
    return p->desc.n_slots > 0 ? p->desc.n_groups2 : p->desc.n_groups1;
}

extern "C" int agf_program_info(const agf_program_t *p, agf_program_info_t *info) {
    if (!p || !info) return fail(AGF_E_INVALID, "null argument");
    memset(info, 0, sizeof(*info));
    info->n_stripes = (int32_t)p->stripes.size();
    info->n_recs = p->n_recs;
    info->n_cols = p->desc.n_cols;
    info->out_dtype = p->desc.out_dtype;
    info->n_out_groups = out_groups(p);
    info->partial_bytes = p->direct_out ? 0 : (int64_t)p->n_recs * p->desc.n_slots * p->n_cells * 8;
    info->direct_out = p->direct_out;
    info->out_bytes = info->n_out_groups * p->desc.n_cols * p->n_cells * (p->desc.out_dtype == AGF_F64 ? 8 : 4);
    info->valid_bytes = info->n_out_groups * p->n_cells;
    info->kernel_lanes = p->kernel_lanes;
    info->kernel_slots = p->kernel_slots;
    info->kernel_mode = p->kernel_diag;
    info->uses_tma = encode_tiled_fn() != nullptr && !getenv("AGF_DISABLE_TMA");
    info->kernel_kinds = (int32_t)p->kernel_kinds;
    return 0;
}

extern "C" int agf_program_stripe_rows(const agf_program_t *p, int32_t s, int64_t *row_begin, int64_t *row_end) {
    if (!p || s < 0 || s >= (int32_t)p->stripes.size()) return fail(AGF_E_INVALID, "bad stripe index");
    if (row_begin) *row_begin = p->b1[p->stripes[s].g1_begin];
    if (row_end) *row_end = p->b1[p->stripes[s].g1_end];
    return 0;
}

// ------------------------------------------------------------------------------------------
// TMA tensor map (driver entry point fetched through the runtime: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

bool agf_tma_eligible(const void *base, int elem_size, uint64_t ld) {
    if (getenv("AGF_DISABLE_TMA")) return false;
    return ((uintptr_t)base % 16 == 0) && ((ld * (uint64_t)elem_size) % 16 == 0) && encode_tiled_fn() != nullptr;
}

int agf_make_tensor_map(agf::TensorMap *out, const void *base, int elem_size, uint64_t n_cells, uint64_t n_rows,
                        uint64_t ld, int box_rows) {
    static_assert(sizeof(agf::TensorMap) == sizeof(CUtensorMap), "tensor map size");
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(AGF_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
    cuuint64_t dims[2] = {n_cells, n_rows};
    cuuint64_t strides[1] = {ld * (cuuint64_t)elem_size};
    cuuint32_t box[2] = {(cuuint32_t)agf::TMA_CW, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn((CUtensorMap *)out, elem_size == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                    2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AGF_E_UNSUPPORTED, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

static int check_device(int device) {
    int dev = -1;
    CU(cudaGetDevice(&dev));
    if (dev != device) return fail(AGF_E_STATE, "handle belongs to device %d, current device is %d", device, dev);
    return 0;
}

extern "C" int agf_temporal_run(const agf_program_t *p, const void *d_x, int64_t ld, int64_t row0,
                                int32_t stripe_begin, int32_t stripe_end, double *d_partial, void *d_out,
                                uint8_t *d_valid, int32_t out_ncols, int32_t valid_and, uintptr_t stream) {
    if (!p || !d_x) return fail(AGF_E_INVALID, "null argument");
    int rc = check_device(p->device);
    if (rc) return rc;
    if (stripe_begin < 0 || stripe_end > (int32_t)p->stripes.size() || stripe_begin >= stripe_end)
        return fail(AGF_E_INVALID, "bad stripe range [%d, %d)", stripe_begin, stripe_end);
    if (ld < p->n_cells) return fail(AGF_E_INVALID, "ld < n_cells");
    if (row0 < 0 || row0 > p->b1[p->stripes[stripe_begin].g1_begin])
        return fail(AGF_E_INVALID, "row0 is past the first row of the stripe range");
    if (p->desc.n_slots > 0 && !p->direct_out && !d_partial) return fail(AGF_E_INVALID, "two-level program needs d_partial");
    if ((p->desc.n_slots == 0 || p->direct_out) && (!d_out || !d_valid))
        return fail(AGF_E_INVALID, "this program writes X / V from agf_temporal_run: d_out and d_valid are required");
    for (int c = 0; c < p->desc.n_cols; ++c)
        if (p->desc.cols[c].dst >= out_ncols) return fail(AGF_E_INVALID, "col %d: dst outside X (out_ncols=%d)", c, out_ncols);
    const int esz = p->desc.in_dtype == AGF_F64 ? 8 : 4;
    K1Launch a{p, d_x, ld, row0, stripe_begin, stripe_end, d_partial, d_out, d_valid, out_ncols, valid_and,
               (cudaStream_t)stream, agf_tma_eligible(d_x, esz, (uint64_t)ld) ? 1 : 0};
    int krc = 0;
    int sel = k1_select(a, 0, nullptr, &krc);
    return sel ? sel : krc;
}

extern "C" int agf_temporal_finalize(const agf_program_t *p, const double *d_partial, void *d_out,
                                     uint8_t *d_valid, int32_t out_ncols, int32_t valid_and, uintptr_t stream) {
    if (!p) return fail(AGF_E_INVALID, "null program");
    if (p->desc.n_slots == 0 || p->direct_out) return 0;
    if (!d_partial || !d_out || !d_valid) return fail(AGF_E_INVALID, "null buffer");
    int rc = check_device(p->device);
    if (rc) return rc;
    const agf_program_desc_t &d = p->desc;
    FinParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.partial = d_partial;
    fp.out = d_out;
    fp.valid = d_valid;
    fp.b2 = p->d_b2;
    fp.g2_rec_ptr = p->d_g2_rec_ptr;
    fp.g2_rec_idx = p->d_g2_rec_idx;
    fp.n_cells = (int)p->n_cells;
    fp.n_slots = d.n_slots;
    fp.n_cols = d.n_cols;
    fp.in_f64 = d.in_dtype == AGF_F64;
    fp.out_f64 = d.out_dtype == AGF_F64;
    for (int c = 0; c < d.n_cols; ++c)
        if (d.cols[c].dst >= out_ncols) return fail(AGF_E_INVALID, "col %d: dst outside X (out_ncols=%d)", c, out_ncols);
    fp.out_ncols = out_ncols;
    fp.valid_and = valid_and;
    for (int j = 0; j < d.n_slots; ++j) {
        SlotP &S = fp.slots[j];
        S.src = d.slots[j].src;
        S.xform = d.slots[j].xform;
        S.xparam = d.slots[j].xparam;
        S.x_f64 = d.slots[j].x_f64;
        S.calc = d.slots[j].calc;
        S.flag = d.slots[j].flag;
        S.t0 = d.slots[j].t0;
        S.t1 = d.slots[j].t1;
        S.ip = (S.xform == AGF_XF_POWI) ? (int)S.xparam : 1;
    }
    for (int c = 0; c < d.n_cols; ++c) {
        ColP &C = fp.cols[c];
        C.src = d.cols[c].src;
        C.xform = d.cols[c].xform;
        C.xparam = d.cols[c].xparam;
        C.x_f64 = d.cols[c].x_f64;
        C.dst = d.cols[c].dst;
    }
    if (d.n_groups2 > 65535) return fail(AGF_E_UNSUPPORTED, "more than 65535 level-2 groups");
    dim3 grid((unsigned)((p->n_cells + 255) / 256), (unsigned)d.n_groups2);
    agf_finalize<<<grid, 256, 0, (cudaStream_t)stream>>>(fp);
    CU(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// CSR + K2
// ------------------------------------------------------------------------------------------
extern "C" int agf_csr_create(agf_csr_t **out, int32_t n_regions, int64_t n_cells, int64_t nnz,
                              const int32_t *d_row_ptr, const int32_t *d_cell_idx, const double *d_w) {
    if (!out) return fail(AGF_E_INVALID, "null out");
    *out = nullptr;
    if (n_regions <= 0 || n_cells <= 0 || nnz < 0 || nnz > 0x7fffffff) return fail(AGF_E_INVALID, "bad CSR sizes");
    if (!d_row_ptr || (nnz > 0 && (!d_cell_idx || !d_w))) return fail(AGF_E_INVALID, "null CSR array");
    int dev = -1;
    CU(cudaGetDevice(&dev));
    agf_csr *c = new agf_csr();
    c->n_regions = n_regions;
    c->n_cells = n_cells;
    c->nnz = nnz;
    c->d_row_ptr = d_row_ptr;
    c->d_cell_idx = d_cell_idx;
    c->d_w = d_w;
    c->device = dev;
    *out = c;
    return 0;
}

extern "C" int agf_csr_destroy(agf_csr_t *c) {
    delete c;
    return 0;
}

extern "C" int agf_spmm_run(const agf_csr_t *c, const void *d_x, int32_t x_dtype, const uint8_t *d_valid,
                            int64_t n_groups, int32_t n_cols, double *d_panel, double *d_den, uintptr_t stream) {
    if (!c || !d_x || !d_valid || !d_panel) return fail(AGF_E_INVALID, "null argument");
    if (n_groups <= 0 || n_cols <= 0) return fail(AGF_E_INVALID, "bad sizes");
    int rc = check_device(c->device);
    if (rc) return rc;
    // lanes per (period, region) pair: the smallest of 8 / 16 / 32 that covers the mean row length
    const double mean_row = (double)c->nnz / (double)c->n_regions;
    const long long pairs = (long long)c->n_regions * n_groups;
    int sms = 148;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    // enough pairs for >= 2048 threads on every SM: one thread per pair (no cross-lane reduction)
    int gs = pairs >= 2048LL * sms ? 1 : (mean_row <= 8.0 ? 8 : (mean_row <= 16.0 ? 16 : 32));
    if (const char *force = getenv("AGF_SPMM_GS")) {  // test hook: pin the variant (1, 8, 16 or 32)
        const int f = atoi(force);
        if (f == 1 || f == 8 || f == 16 || f == 32) gs = f;
    }
    const long long blocks = (pairs * gs + 255) / 256;
    if (blocks > 0x7fffffffLL) return fail(AGF_E_UNSUPPORTED, "panel too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    // columns per load: the widest of 16 / 8 / sizeof(X) bytes that divides the row pitch (d_x itself
    // must then be aligned to it; torch allocations are 256-byte aligned)
    const int esz = x_dtype == AGF_F64 ? 8 : 4;
    int vb = esz;
    for (int cand = 16; cand > esz; cand >>= 1)
        if (((long long)n_cols * esz) % cand == 0 && ((uintptr_t)d_x % cand) == 0) {
            vb = cand;
            break;
        }
    const int ev = vb / esz;
    const bool wide = n_cols > 4;
#define AGF_SPMM_W(TX, GS, EV, W)                                                                                    \
    agf_spmm<TX, GS, EV, W><<<(unsigned)blocks, 256, 0, st>>>(c->d_row_ptr, c->d_cell_idx, c->d_w, (const TX *)d_x, \
                                                              d_valid, c->n_cells, n_groups, n_cols, c->n_regions,  \
                                                              d_panel, d_den)
#define AGF_SPMM(TX, GS, EV)                \
    do {                                    \
        if (wide) AGF_SPMM_W(TX, GS, EV, true); \
        else AGF_SPMM_W(TX, GS, EV, false); \
    } while (0)
#define AGF_SPMM_GS(TX, EV)                    \
    do {                                       \
        if (gs == 1) AGF_SPMM(TX, 1, EV);      \
        else if (gs == 8) AGF_SPMM(TX, 8, EV); \
        else if (gs == 16) AGF_SPMM(TX, 16, EV); \
        else AGF_SPMM(TX, 32, EV);             \
    } while (0)
    if (x_dtype == AGF_F64) {
        if (ev == 2) AGF_SPMM_GS(double, 2);
        else AGF_SPMM_GS(double, 1);
    } else {
        if (ev == 4) AGF_SPMM_GS(float, 4);
        else if (ev == 2) AGF_SPMM_GS(float, 2);
        else AGF_SPMM_GS(float, 1);
    }
#undef AGF_SPMM_GS
#undef AGF_SPMM_W
#undef AGF_SPMM
    CU(cudaGetLastError());
    return 0;
}

extern "C" int agf_valid_mask_run(const void *d_x, int32_t x_dtype, int64_t n_groups, int32_t n_cols,
                                  int64_t n_cells, uint8_t *d_valid, uintptr_t stream) {
    if (!d_x || !d_valid) return fail(AGF_E_INVALID, "null argument");
    if (n_groups <= 0 || n_groups > 0x7fffffff || n_cols <= 0 || n_cells <= 0) return fail(AGF_E_INVALID, "bad sizes");
    if ((n_cells + 255) / 256 > 65535) return fail(AGF_E_UNSUPPORTED, "more than 16776960 cells per step");
    dim3 grid((unsigned)n_groups, (unsigned)((n_cells + 255) / 256));
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == AGF_F64)
        agf_valid_mask<double><<<grid, 256, 0, st>>>((const double *)d_x, n_cells, n_cols, d_valid);
    else
        agf_valid_mask<float><<<grid, 256, 0, st>>>((const float *)d_x, n_cells, n_cols, d_valid);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int agf_elementwise_run(const void *d_in, int32_t in_dtype, void *d_out, int32_t out_dtype, int64_t n,
                                   int32_t xform, double xparam, const void *d_other, int32_t other_dtype,
                                   uint8_t *d_valid, int32_t n_pre, const agf_pre_t *pre, uintptr_t stream) {
    if (!d_in || !d_out) return fail(AGF_E_INVALID, "null argument");
    if (n_pre < 0 || n_pre > AGF_MAX_PRE || (n_pre > 0 && !pre)) return fail(AGF_E_INVALID, "bad preprocess chain");
    EwPre ep;
    memset(&ep, 0, sizeof(ep));
    ep.n = n_pre;
    for (int i = 0; i < n_pre; ++i) {
        if (pre[i].op < AGF_PRE_ADD || pre[i].op > AGF_PRE_NEG) return fail(AGF_E_INVALID, "pre %d: bad op", i);
        ep.op[i].op = pre[i].op;
        ep.op[i].c = in_dtype == AGF_F64 ? pre[i].c : (double)(float)pre[i].c;
    }
    if (n <= 0) return fail(AGF_E_INVALID, "bad size");
    if (!d_other && !(xform >= AGF_XF_POWI && xform <= AGF_XF_SPLINE2)) return fail(AGF_E_INVALID, "bad xform");
    if (in_dtype == AGF_F64 && out_dtype != AGF_F64) return fail(AGF_E_INVALID, "float64 input needs float64 output");
    if (d_other && other_dtype == AGF_F64 && out_dtype != AGF_F64)
        return fail(AGF_E_INVALID, "float64 interaction needs float64 output");
    int dev = 0, sms = 148;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16);
    cudaStream_t st = (cudaStream_t)stream;
#define AGF_EW(TI, TO, TB)                                                                                  \
    agf_elementwise<TI, TO, TB><<<blocks, 256, 0, st>>>((const TI *)d_in, (TO *)d_out, (const TB *)d_other, \
                                                        (long long)n, xform, xparam, d_valid, ep)
    const bool of64 = d_other && other_dtype == AGF_F64;
    if (in_dtype == AGF_F64) {
        if (of64) AGF_EW(double, double, double);
        else AGF_EW(double, double, float);
    } else if (out_dtype == AGF_F64) {
        if (of64) AGF_EW(float, double, double);
        else AGF_EW(float, double, float);
    } else {
        AGF_EW(float, float, float);
    }
#undef AGF_EW
    CU(cudaGetLastError());
    return 0;
}
