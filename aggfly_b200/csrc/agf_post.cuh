// agf_post.cuh -- kernels after the temporal pass: finalize (K1f), CSR regional average (K2),
// validity mask.  Included by agf_api.cu only.
#pragma once
#include "agf_kernels.cuh"

namespace agf {

// ------------------------------------------------------------------------------------------
// K1f: finalize (two-level programs)
// ------------------------------------------------------------------------------------------
struct FinParams {
    const double *partial;
    void *out;
    unsigned char *valid;
    const int *b2;
    const int *g2_rec_ptr;  // [G2+1] -> range in g2_rec_idx
    const int *g2_rec_idx;  // record ids of each level-2 group, in stripe order
    int n_cells, n_slots, n_cols, in_f64, out_f64;
    int out_ncols, valid_and;
    SlotP slots[AGF_MAX_SLOTS];
    ColP cols[AGF_MAX_COLS];
};

__global__ void __launch_bounds__(256) agf_finalize(const __grid_constant__ FinParams p) {
    const int cell = blockIdx.x * 256 + threadIdx.x;
    if (cell >= p.n_cells) return;
    const int g2 = blockIdx.y;
    const int r0 = p.g2_rec_ptr[g2], r1 = p.g2_rec_ptr[g2 + 1];
    const int n2 = p.b2[g2 + 1] - p.b2[g2];
    bool ok = true;
    for (int c = 0; c < p.n_cols; ++c) {
        const ColP &C = p.cols[c];
        const SlotP &S = p.slots[C.src];
        double v;
        if (r1 == r0) {
            v = agf_nan();  // empty level-2 group
        } else {
            v = p.partial[((size_t)p.g2_rec_idx[r0] * p.n_slots + C.src) * p.n_cells + cell];
            for (int r = r0 + 1; r < r1; ++r)
                v = l2_merge(S.calc, v,
                             p.partial[((size_t)p.g2_rec_idx[r] * p.n_slots + C.src) * p.n_cells + cell]);
            if (S.calc == AGF_CALC_MEAN) v = v / (double)n2;
        }
        const bool slot_f64 = S.x_f64 || p.in_f64;
        if (!slot_f64) v = (double)(float)v;  // stored in the dtype of the slot's input series
        if (C.xform != AGF_XF_NONE) {
            if (slot_f64) {
                v = apply_xform64(v, C.xform, C.xparam);
            } else {
                v = apply_xform<float>(v, C.xform, C.xparam, C.x_f64);
            }
        }
        ok &= (v == v);
        const size_t idx = ((size_t)g2 * p.n_cells + cell) * p.out_ncols + C.dst;  // X[g, cell, :]
        if (p.out_f64)
            reinterpret_cast<double *>(p.out)[idx] = v;
        else
            reinterpret_cast<float *>(p.out)[idx] = (float)v;
    }
    unsigned char *vp = p.valid + (size_t)g2 * p.n_cells + cell;
    *vp = (p.valid_and ? (*vp != 0) && ok : ok) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// K2: CSR weighted regional average.  A group of GS lanes (8, 16 or 32, picked from the mean row
// length) owns one (period g, region r) pair; pairs are numbered g-major / r-fastest so that the
// warps in flight at any moment gather from the same X[g, c, :] planes at neighbouring cells
// (regions are spatially coherent in shapefile order): DRAM pages and 32-byte sectors touched by
// one region are still open / in L2 when its neighbours read them.  (r-major numbering made every
// warp of a CTA read a different 58 MB-apart plane: 40 ms for the C3b daily panel.)
// ------------------------------------------------------------------------------------------
// GS == 1 (one thread per pair, entries walked sequentially in weights-frame order like the
// reference's np.add.at) is used when there are enough pairs to fill the GPU (daily / monthly
// panels): no shuffles, 16 columns per pass, and the lanes of a warp are 32 neighbouring regions of
// the same period, so their gathers fall into a few shared sectors that stay in L1 while the
// threads walk their rows.  Measured on the C3b daily panel (45 000 x 365 pairs x 14 columns): the
// warp-per-pair form was instruction-bound (25e9 warp instructions, 43.7 ms, 5 % of DRAM peak).
// (A column-parallel form -- LP lanes per pair, each owning a few columns, one load per cell row -- was
// measured too: 10.8 ms, 8.5 ms with a 4-entry unroll, against 7.5 ms for this one; it has one row load in
// flight per lane where this form has n_cols / EV, and the walk is latency-bound.)
// columns accumulated per pass over a region's entries: WIDE = false for panels of <= 4 columns (a
// 16-wide unrolled, guarded accumulator block cost 1900 instructions per warp on the one-column C5
// panel), else 16 for the thread-per-pair form and 8 for the group forms
template <int GS, bool WIDE>
__host__ __device__ constexpr int spmm_ncb() {
    return !WIDE ? 4 : (GS == 1 ? 16 : 8);
}

// EV consecutive columns of one cell with a single load (EV * sizeof(TX) = 4, 8 or 16 bytes; the
// launcher picks the widest width the row pitch n_cols * sizeof(TX) is a multiple of)
template <typename TX, int EV>
__device__ __forceinline__ void load_cols(const TX *p, double (&x)[EV]) {
    if constexpr (EV == 1) {
        x[0] = (double)__ldg(p);
    } else if constexpr (sizeof(TX) * EV == 8) {  // float2
        const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
        x[0] = (double)v.x;
        x[1] = (double)v.y;
    } else if constexpr (sizeof(TX) == 4) {  // float4
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        x[0] = (double)v.x;
        x[1] = (double)v.y;
        x[2] = (double)v.z;
        x[3] = (double)v.w;
    } else {  // double2
        const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
        x[0] = v.x;
        x[1] = v.y;
    }
}

template <typename TX, int GS, int EV, bool WIDE>
__global__ void __launch_bounds__(256)
    agf_spmm(const int *__restrict__ row_ptr, const int *__restrict__ cell_idx,
             const double *__restrict__ w, const TX *__restrict__ X,
             const unsigned char *__restrict__ V, long long n_cells, long long G, int n_cols,
             int n_regions, double *__restrict__ panel, double *__restrict__ den_out) {
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / GS;
    const int lane = threadIdx.x % GS;
    if (gid >= (long long)n_regions * G) return;  // whole groups leave together (256 % GS == 0)
    const long long g = gid / n_regions;
    const int r = (int)(gid % n_regions);
    const int e0 = row_ptr[r], e1 = row_ptr[r + 1];
    const unsigned char *Vg = V + (size_t)g * n_cells;
    // lanes of this group inside the warp (shuffles must name exactly the participating lanes)
    const unsigned gmask = (GS == 32) ? 0xffffffffu : (((1u << GS) - 1u) << ((threadIdx.x & 31) / GS * GS));
    constexpr int SPMM_NCB = spmm_ncb<GS, WIDE>();
    const TX *Xg = X + (size_t)g * n_cols * n_cells;  // X[g, cell, c]: a cell's columns are contiguous

    for (int c0 = 0; c0 < n_cols; c0 += SPMM_NCB) {
        double acc[SPMM_NCB];
#pragma unroll
        for (int c = 0; c < SPMM_NCB; ++c) acc[c] = 0.0;
        double den = 0.0;
        // software pipeline: the next entry's index, weight and mask byte are requested before this entry's
        // rows are consumed -- the walk is a chain of dependent loads (index -> mask -> row) otherwise
        int e = e0 + lane;
        int cell_n = e < e1 ? cell_idx[e] : 0;
        double we_n = e < e1 ? w[e] : 0.0;
        unsigned char ok_n = e < e1 ? Vg[cell_n] : (unsigned char)0;
        for (; e < e1; e += GS) {
            const int cell = cell_n;
            const double we = we_n;
            const unsigned char ok = ok_n;
            if (e + GS < e1) {
                cell_n = cell_idx[e + GS];
                we_n = w[e + GS];
                ok_n = Vg[cell_n];
            }
            if (ok) {
                const TX *Xc = Xg + (size_t)cell * n_cols;
                den += we;
#pragma unroll
                for (int c = 0; c < SPMM_NCB; c += EV) {
                    if (c0 + c < n_cols) {  // n_cols is a multiple of EV
                        double x[EV];
                        load_cols<TX, EV>(Xc + c0 + c, x);
#pragma unroll
                        for (int k = 0; k < EV; ++k) acc[c + k] += we * x[k];
                    }
                }
            }
        }
#pragma unroll
        for (int off = GS / 2; off > 0; off >>= 1) {
            den += __shfl_xor_sync(gmask, den, off);
#pragma unroll
            for (int c = 0; c < SPMM_NCB; ++c) acc[c] += __shfl_xor_sync(gmask, acc[c], off);
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < SPMM_NCB; ++c)
                if (c0 + c < n_cols)
                    panel[((size_t)r * G + g) * n_cols + c0 + c] = (den != 0.0) ? acc[c] / den : agf_nan();
            if (c0 == 0 && den_out) den_out[(size_t)r * G + g] = den;
        }
    }
}

// shared validity mask of an existing X: V[g, cell] = AND_c !isnan(X[g, c, cell])  (spatial.py:114-119)
template <typename TX>
__global__ void __launch_bounds__(256)
    agf_valid_mask(const TX *__restrict__ X, long long n_cells, int n_cols, unsigned char *__restrict__ V) {
    // blockIdx.y: cell block, blockIdx.x: group (grid.x may hold 2^31 - 1 groups: an un-aggregated hourly series of any length)
    const long long cell = (long long)blockIdx.y * 256 + threadIdx.x;
    if (cell >= n_cells) return;
    const size_t g = blockIdx.x;
    bool ok = true;
    for (int c = 0; c < n_cols; ++c) {
        const TX v = X[(g * n_cells + cell) * n_cols + c];
        ok &= (v == v);
    }
    V[g * n_cells + cell] = ok ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// elementwise transform of a whole series (materialised transforms: a power / spline / interaction
// that cannot ride inside a fused program, e.g. applied to the raster before the first aggregate
// step, or two transforms in a row): out[i] = f(in[i]) (* other[i]), V[i] = !isnan(out[i]).
// Dtype rules are those of apply_xform (NumPy promotion decided by the host).
// ------------------------------------------------------------------------------------------
struct EwPre {  // preprocess chain of the raster (by value in the parameter bank)
    int n;
    PreP<double> op[AGF_MAX_PRE];  // constants already rounded to the input dtype
};

template <typename TI, typename TO, typename TB>
__global__ void __launch_bounds__(256)
    agf_elementwise(const TI *__restrict__ in, TO *__restrict__ out, const TB *__restrict__ other, long long n,
                    int xform, double xparam, unsigned char *__restrict__ valid, const __grid_constant__ EwPre pre) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        TI xin = in[i];
        for (int k = 0; k < pre.n; ++k) {  // one rounding per operation, in the input dtype
            const TI c = (TI)pre.op[k].c;
            switch (pre.op[k].op) {
                case AGF_PRE_ADD: xin = xin + c; break;
                case AGF_PRE_SUB: xin = xin - c; break;
                case AGF_PRE_RSUB: xin = c - xin; break;
                case AGF_PRE_MUL: xin = xin * c; break;
                case AGF_PRE_DIV: xin = xin / c; break;
                case AGF_PRE_RDIV: xin = c / xin; break;
                default: xin = -xin; break;
            }
        }
        const double x = (double)xin;
        double r;
        if (other != nullptr) {
            // np.multiply(array, inter) in the promoted dtype (dataset.py:547-563)
            if (sizeof(TO) == 4)
                r = (double)((float)x * (float)other[i]);
            else
                r = x * (double)other[i];
        } else if (sizeof(TI) == 4) {
            r = apply_xform<float>(x, xform, xparam, sizeof(TO) == 8);
        } else {
            r = apply_xform64(x, xform, xparam);
        }
        out[i] = (TO)r;
        if (valid != nullptr) valid[i] = (r == r) ? 1 : 0;
    }
}

}  // namespace agf
