/*
 * oracle/agf_oracle.c -- CPU restatement of aggfly's aggregation arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (aggfly_b200/) may import, link or
 * call this file; it exists so tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * `--impl reference` legs can check and time the reference algorithm without the reference
 * package (which cannot be imported in this image: no xarray/dask/geopandas).
 *
 * Parity status: PINNED.  Checked by tests/test_oracle.py against
 *   (a) tests/golden/ref_kernels.npz -- outputs of the reference's own numba kernels and
 *       spatial helpers, produced in the build container by tests/golden/make_golden.py;
 *   (b) the hard-coded expectations of the reference's tests
 *       (aggfly/tests/test_aggregate.py:275-280, 311-313, 454-466, 614-664).
 *
 * Each function follows the reference loop nest exactly (cell-major outer loops, time-ordered
 * fp64 accumulation, result stored in the input dtype), so it is both the numerical oracle
 * and a faithful model of the reference's CPU cost:
 *
 *   orc_block_stat_*     <- aggfly/aggregate/nb_kernels.py:121-155  (_block_stat)
 *   orc_block_dd_*       <- aggfly/aggregate/nb_kernels.py:158-179  (_block_dd)
 *   orc_block_bins_*     <- aggfly/aggregate/nb_kernels.py:182-199  (_block_bins)
 *   orc_block_sine_dd_*  <- aggfly/aggregate/nb_kernels.py:202-251  (_block_sine_dd)
 *   orc_scatter_block    <- aggfly/aggregate/spatial.py:181-186     (_scatter_block)
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: numba compiles with fastmath=False and no FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_STAT_MEAN 0
#define ORC_STAT_SUM 1
#define ORC_STAT_MIN 2
#define ORC_STAT_MAX 3
#define ORC_STAT_NANMEAN 4

int orc_version(void) { return 1; }

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* cube is C-contiguous [T, NY, NX]; bounds is int64[G+1]; out is [G, NY, NX] (stat) or
 * [G, NY, NX, D] (dd / bins / sine_dd), same element type as cube. */
#define DEFINE_KERNELS(SUF, TYPE)                                                              \
    void orc_block_stat_##SUF(const TYPE *cube, int64_t NY, int64_t NX, const int64_t *bounds, \
                              int64_t G, int code, TYPE *out) {                                \
        _Pragma("omp parallel for schedule(static)") for (int64_t iy = 0; iy < NY; ++iy) {     \
            for (int64_t ix = 0; ix < NX; ++ix) {                                              \
                for (int64_t g = 0; g < G; ++g) {                                              \
                    int64_t lo = bounds[g], hi = bounds[g + 1];                                \
                    int64_t n = 0;                                                             \
                    double s = 0.0;                                                            \
                    TYPE mn = (TYPE)INFINITY, mx = (TYPE)-INFINITY;                            \
                    int hasnan = 0;                                                            \
                    for (int64_t k = lo; k < hi; ++k) {                                        \
                        TYPE v = cube[(k * NY + iy) * NX + ix];                                \
                        if (isnan(v)) {                                                        \
                            hasnan = 1;                                                        \
                        } else {                                                               \
                            s += (double)v;                                                    \
                            n += 1;                                                            \
                            if (v < mn) mn = v;                                                \
                            if (v > mx) mx = v;                                                \
                        }                                                                      \
                    }                                                                          \
                    TYPE *o = &out[(g * NY + iy) * NX + ix];                                   \
                    if (hi == lo)                                                              \
                        *o = (TYPE)NAN;                                                        \
                    else if (code == ORC_STAT_NANMEAN)                                         \
                        *o = (n > 0) ? (TYPE)(s / (double)n) : (TYPE)NAN;                      \
                    else if (hasnan)                                                           \
                        *o = (TYPE)NAN;                                                        \
                    else if (code == ORC_STAT_MEAN)                                            \
                        *o = (TYPE)(s / (double)n);                                            \
                    else if (code == ORC_STAT_SUM)                                             \
                        *o = (TYPE)s;                                                          \
                    else if (code == ORC_STAT_MIN)                                             \
                        *o = mn;                                                               \
                    else                                                                       \
                        *o = mx;                                                               \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
    }                                                                                          \
                                                                                               \
    void orc_block_dd_##SUF(const TYPE *cube, int64_t NY, int64_t NX, const int64_t *bounds,   \
                            int64_t G, const double *ddargs, int64_t D, TYPE *out) {           \
        _Pragma("omp parallel for schedule(static)") for (int64_t iy = 0; iy < NY; ++iy) {     \
            for (int64_t ix = 0; ix < NX; ++ix) {                                              \
                for (int64_t g = 0; g < G; ++g) {                                              \
                    int64_t lo = bounds[g], hi = bounds[g + 1];                                \
                    for (int64_t d = 0; d < D; ++d) {                                          \
                        double t0 = ddargs[3 * d], t1 = ddargs[3 * d + 1];                     \
                        double base = (ddargs[3 * d + 2] == 0.0) ? t0 : t1;                    \
                        double acc = 0.0;                                                      \
                        int hasnan = 0;                                                        \
                        for (int64_t k = lo; k < hi; ++k) {                                    \
                            double v = (double)cube[(k * NY + iy) * NX + ix];                  \
                            if (isnan(v)) {                                                    \
                                hasnan = 1;                                                    \
                            } else if (v > t0 && v < t1) {                                     \
                                double av = v - base;                                          \
                                if (av < 0.0) av = -av;                                        \
                                acc += av;                                                     \
                            }                                                                  \
                        }                                                                      \
                        out[((g * NY + iy) * NX + ix) * D + d] =                               \
                            (hasnan || hi == lo) ? (TYPE)NAN : (TYPE)acc;                      \
                    }                                                                          \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
    }                                                                                          \
                                                                                               \
    void orc_block_bins_##SUF(const TYPE *cube, int64_t NY, int64_t NX, const int64_t *bounds, \
                              int64_t G, const double *ddargs, int64_t D, TYPE *out) {         \
        _Pragma("omp parallel for schedule(static)") for (int64_t iy = 0; iy < NY; ++iy) {     \
            for (int64_t ix = 0; ix < NX; ++ix) {                                              \
                for (int64_t g = 0; g < G; ++g) {                                              \
                    int64_t lo = bounds[g], hi = bounds[g + 1];                                \
                    for (int64_t d = 0; d < D; ++d) {                                          \
                        double t0 = ddargs[3 * d], t1 = ddargs[3 * d + 1];                     \
                        double c = 0.0;                                                        \
                        for (int64_t k = lo; k < hi; ++k) {                                    \
                            double v = (double)cube[(k * NY + iy) * NX + ix];                  \
                            if (v > t0 && v < t1) c += 1.0;                                    \
                        }                                                                      \
                        out[((g * NY + iy) * NX + ix) * D + d] =                               \
                            (hi == lo) ? (TYPE)NAN : (TYPE)c;                                  \
                    }                                                                          \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
    }                                                                                          \
                                                                                               \
    void orc_block_sine_dd_##SUF(const TYPE *cube, int64_t NY, int64_t NX,                     \
                                 const int64_t *bounds, int64_t G, const double *ddargs,       \
                                 int64_t D, TYPE *out) {                                       \
        const double PI = 3.141592653589793;                                                   \
        _Pragma("omp parallel for schedule(static)") for (int64_t iy = 0; iy < NY; ++iy) {     \
            for (int64_t ix = 0; ix < NX; ++ix) {                                              \
                for (int64_t g = 0; g < G; ++g) {                                              \
                    int64_t lo = bounds[g], hi = bounds[g + 1];                                \
                    int64_t n = 0;                                                             \
                    double s = 0.0;                                                            \
                    double tmax = -INFINITY, tmin = INFINITY;                                  \
                    int hasnan = 0;                                                            \
                    for (int64_t k = lo; k < hi; ++k) {                                        \
                        double v = (double)cube[(k * NY + iy) * NX + ix];                      \
                        if (isnan(v)) {                                                        \
                            hasnan = 1;                                                        \
                        } else {                                                               \
                            s += v;                                                            \
                            n += 1;                                                            \
                            if (v > tmax) tmax = v;                                            \
                            if (v < tmin) tmin = v;                                            \
                        }                                                                      \
                    }                                                                          \
                    for (int64_t d = 0; d < D; ++d) {                                          \
                        TYPE *o = &out[((g * NY + iy) * NX + ix) * D + d];                     \
                        if (hasnan || n == 0) {                                                \
                            *o = (TYPE)NAN;                                                    \
                            continue;                                                          \
                        }                                                                      \
                        double tavg = s / (double)n;                                           \
                        double kind = ddargs[3 * d + 2];                                       \
                        double val = 0.0;                                                      \
                        for (int j = 0; j < 2; ++j) {                                          \
                            double thr = ddargs[3 * d + j];                                    \
                            double part;                                                       \
                            if (kind == 0.0) {                                                 \
                                if (thr <= tmin) {                                             \
                                    part = tavg - thr;                                         \
                                } else if (thr < tmax && tmin < thr) {                         \
                                    double rng = tmax - tmin;                                  \
                                    double a = acos((2.0 * thr - tmax - tmin) / rng);          \
                                    part = ((tavg - thr) * a + rng * sin(a) / 2.0) / PI;       \
                                } else {                                                       \
                                    part = 0.0;                                                \
                                }                                                              \
                                val += (j == 0) ? part : -part;                                \
                            } else {                                                           \
                                if (thr >= tmax) {                                             \
                                    part = thr - tavg;                                         \
                                } else if (thr < tmax && tmin < thr) {                         \
                                    double alpha = (tmax - tmin) / 2.0;                        \
                                    double r = (thr - tavg) / alpha;                           \
                                    double at = atan(r / sqrt(1.0 - r * r));                   \
                                    part = (1.0 / PI) *                                        \
                                           ((thr - tavg) * (at + PI / 2.0) + alpha * cos(at)); \
                                } else {                                                       \
                                    part = 0.0;                                                \
                                }                                                              \
                                val += (j == 0) ? -part : part;                                \
                            }                                                                  \
                        }                                                                      \
                        *o = (TYPE)val;                                                        \
                    }                                                                          \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
    }

DEFINE_KERNELS(f32, float)
DEFINE_KERNELS(f64, double)

/* out[region_idx[e], :] += w[e] * block[cell_idx[e], :], entries applied in order (np.add.at).
 * block is [n_cells, n_t] fp64, out is [n_regions, n_t] fp64 and is zeroed here. */
void orc_scatter_block(const double *block, int64_t n_t, const int64_t *region_idx,
                       const int64_t *cell_idx, const double *w, int64_t nnz, int64_t n_regions,
                       double *out) {
    memset(out, 0, sizeof(double) * (size_t)(n_regions * n_t));
    for (int64_t e = 0; e < nnz; ++e) {
        const double *src = block + cell_idx[e] * n_t;
        double *dst = out + region_idx[e] * n_t;
        double we = w[e];
        for (int64_t t = 0; t < n_t; ++t) dst[t] += we * src[t];
    }
}
