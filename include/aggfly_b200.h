/*
 * aggfly_b200.h -- C-ABI of libaggfly_b200.so, the B200 (sm_100a) engine behind aggfly's
 * `aggregate_dataset` hot path.
 *
 * The reference (dylanhogan/aggfly, pure Python) has no FFI; its backend seam for this path is
 * the `engine=` string ("auto" | "dask" | "numba", aggfly/aggregate/aggregate.py:210-217,
 * aggfly/aggregate/nb_kernels.py:59-74) and, underneath it, two operator-level functions:
 *
 *     numba_resample(da, freq, calc, ddargs, multi_dd)        aggfly/aggregate/nb_kernels.py:271-305
 *         -> _block_stat / _block_dd / _block_bins / _block_sine_dd              :121-251
 *     _scatter_block(block, region_idx, cell_idx, w_vals, n_regions)   aggfly/aggregate/spatial.py:181-186
 *
 * This header is what a maintainer would bind (ctypes) to add
 * `engine="cuda"`:  a *program* is the lowered form of one aggregate_dataset spec (every
 * `('aggregate', ...)` / `('transform', ...)` chain of the call, common prefixes shared, so the
 * hourly raster is read once for all output names); a *CSR* is the lowered weights frame
 * (aggfly/aggregate/spatial.py:157-178).
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns int: 0 = ok, <0 = AGF_E_* library
 *     error, >0 = cudaError_t.  agf_last_error() returns a thread-local message.
 *   - the caller owns every data buffer (device pointers are e.g. torch tensors' data_ptr());
 *     the library never allocates or frees caller-visible memory.  Opaque handles own only
 *     small private tables (group bounds, stripe tables) and are destroyed explicitly.
 *   - every launch takes a cudaStream_t (as uintptr_t) and is asynchronous; no hidden syncs.
 *   - handles are bound to the device that was current at creation; distinct handles are
 *     independent and may be used from different threads.
 *   - there is no CPU fallback anywhere behind this header.
 *
 * Data layout
 *   raster   x[T, n_cells]      time-major, cell = lat_index * n_lon + lon_index fastest
 *                               (row stride `ld` elements), float32 or float64
 *   columns  X[G, n_cells, n_cols]  per-cell temporal results, G = output periods; the columns of
 *                                   a cell are contiguous (the regional average gathers whole
 *                                   cells: one or two sectors per entry instead of n_cols)
 *   valid    V[G, n_cells] uint8    1 iff every column of the call is non-NaN there
 *                                   (shared validity mask, aggfly/aggregate/spatial.py:114-119)
 *   panel    P[n_regions, G, n_cols] float64
 */
#ifndef AGGFLY_B200_H
#define AGGFLY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGF_ABI_VERSION 5

#define AGF_MAX_LANES 32  /* level-1 reducers per program */
#define AGF_MAX_SLOTS 32  /* level-2 reducers per program */
#define AGF_MAX_COLS 64   /* output columns per program   */

/* library error codes (negative) */
#define AGF_E_INVALID (-1)      /* bad argument / descriptor */
#define AGF_E_UNSUPPORTED (-2)  /* valid spec the fused kernels do not cover (host must split) */
#define AGF_E_NOMEM (-3)
#define AGF_E_STATE (-4)        /* handle used on the wrong device / after destroy */

/* calc codes: aggfly/aggregate/temporal.py:98-134, aggfly/aggregate/nb_kernels.py:33 */
enum {
    AGF_CALC_MEAN = 0,
    AGF_CALC_SUM = 1,
    AGF_CALC_MIN = 2,
    AGF_CALC_MAX = 3,
    AGF_CALC_NANMEAN = 4,
    AGF_CALC_DD = 5,       /* sum |v - base| over t0 < v < t1 (strict), nb_kernels.py:158-179 */
    AGF_CALC_BINS = 6,     /* count of t0 < v < t1 (strict),           nb_kernels.py:182-199 */
    AGF_CALC_SINE_DD = 7,  /* single-sine degree days,                 nb_kernels.py:202-251 */
    AGF_CALC_HIDDEN_SUM = 8, /* helper lanes the host inserts in front of SINE_DD lanes */
    AGF_CALC_HIDDEN_MIN = 9,
    AGF_CALC_HIDDEN_MAX = 10,
    AGF_CALC_DD_R = 11     /* dd whose every term |v - base| is rounded to the raster dtype before it is
                              added: the one-pass form of `dd` over single-row groups followed by `sum`
                              (daily data: dd/date -> sum/month), same bits as the two-step chain */
};

/* transforms between / after aggregate steps: aggfly/dataset/dataset.py:442-481, 527-543 */
enum {
    AGF_XF_NONE = 0,
    AGF_XF_POWI = 1,    /* np.power(x, e), e a small non-negative integer (repeated multiply) */
    AGF_XF_POW = 2,     /* np.power(x, e), general exponent */
    AGF_XF_SPLINE2 = 3  /* (x > 20) * (x - 20), in the value's own dtype */
};

enum { AGF_F32 = 0, AGF_F64 = 1,
       /* storage dtypes of chunked sources (agf_tile_place_run only) */
       AGF_I16 = 2, AGF_I32 = 3, AGF_U8 = 4, AGF_I8 = 5, AGF_U16 = 6 };

/* Elementwise preprocess of every raster value, applied in the raster dtype before any reducer
 * sees it (the reference's Dataset(preprocess=...): aggfly/cli/preprocess.py:24-30 named
 * conversions such as kelvin_to_celsius = x - 273.15, :33-113 arithmetic in x).  One IEEE
 * operation per entry, constants rounded to the raster dtype, no FMA contraction -- the same
 * bits NumPy produces -- fused into the scan instead of a separate pass over the raster. */
enum {
    AGF_PRE_ADD = 0,  /* x + c */
    AGF_PRE_SUB = 1,  /* x - c */
    AGF_PRE_RSUB = 2, /* c - x */
    AGF_PRE_MUL = 3,  /* x * c */
    AGF_PRE_DIV = 4,  /* x / c */
    AGF_PRE_RDIV = 5, /* c / x */
    AGF_PRE_NEG = 6   /* -x    */
};
#define AGF_MAX_PRE 4
typedef struct {
    int32_t op;
    int32_t pad_;
    double c;
} agf_pre_t;

/* One level-1 reducer over the raw time axis (one `('aggregate', {...})` step applied to the
 * raster; one lane per ddargs row for multi-ddargs). */
typedef struct {
    int32_t calc;
    int32_t flag;  /* ddargs[2]: dd base selector (0 -> t0, else t1) / sine_dd kind */
    double t0, t1; /* ddargs[0], ddargs[1] (compared in fp64, like the reference) */
} agf_lane_t;

/* One level-2 reducer: consumes the per-group values of lane `src` (rounded to the raster
 * dtype, nb_kernels.py:260), optionally transformed, grouped by bounds2. */
typedef struct {
    int32_t src;    /* level-1 lane index */
    int32_t xform;  /* AGF_XF_* applied to the lane value before reducing */
    double xparam;  /* exponent for POWI / POW */
    int32_t x_f64;  /* dtype of the transformed value (NumPy promotion decided by the host):
                       0 = raster dtype, 1 = float64 */
    int32_t calc;   /* MEAN, SUM, MIN, MAX, DD, BINS */
    int32_t flag;
    int32_t pad_;
    double t0, t1;
} agf_slot_t;

/* One output column: value of a slot (or, when n_slots == 0, of a lane), optionally transformed
 * after the last aggregate step. */
typedef struct {
    int32_t src;
    int32_t xform;
    double xparam;
    int32_t x_f64; /* dtype of the column after the transform: 0 = raster dtype, 1 = float64 */
    int32_t dst;   /* column index in the destination X (several programs may share one X) */
} agf_col_t;

typedef struct {
    int32_t in_dtype;  /* AGF_F32 | AGF_F64: raster dtype */
    int32_t out_dtype; /* dtype of X: AGF_F64 if any column is float64, else raster dtype */
    int32_t n_lanes;
    int32_t n_slots;   /* 0: single-level program, columns come straight from lanes */
    int32_t n_cols;
    int32_t pad_;
    int64_t n_time;    /* T: rows of the whole time axis the bounds refer to */
    int64_t n_groups1; /* G1 */
    int64_t n_groups2; /* G2 (ignored when n_slots == 0) */
    const int32_t *bounds1; /* host, int32[G1+1], row positions; nb_kernels.py:80-115 */
    const int32_t *bounds2; /* host, int32[G2+1], positions on the level-1 group axis */
    agf_lane_t lanes[AGF_MAX_LANES];
    agf_slot_t slots[AGF_MAX_SLOTS];
    agf_col_t cols[AGF_MAX_COLS];
    int32_t n_pre; /* 0..AGF_MAX_PRE preprocess operations, applied in order */
    int32_t pad2_;
    agf_pre_t pre[AGF_MAX_PRE];
} agf_program_desc_t;

typedef struct {
    int32_t n_stripes;  /* time stripes (cut at level-1 group boundaries) */
    int32_t n_recs;     /* partial records: one per (stripe, level-2 group it touches) */
    int32_t n_cols;
    int32_t out_dtype;
    int64_t n_out_groups;   /* G of X / V / the panel: G2, or G1 when n_slots == 0 */
    int64_t partial_bytes;  /* size of the partial buffer the caller must provide (0 if none) */
    int64_t out_bytes;      /* size of X */
    int64_t valid_bytes;    /* size of V */
    int32_t kernel_lanes, kernel_slots, kernel_mode; /* which instantiation will run */
    int32_t uses_tma;     /* 16-byte aligned rasters run the TMA / shared-memory-ring variant */
    int32_t kernel_kinds; /* compile-time lane-kind set of the chosen instantiation */
    int32_t direct_out;   /* two-level program cut into ONE stripe: agf_temporal_run writes X / V itself
                             (partial_bytes == 0, agf_temporal_finalize is a no-op) */
} agf_program_info_t;

typedef struct agf_program agf_program_t;
typedef struct agf_csr agf_csr_t;

int agf_version(void);
const char *agf_last_error(void);

/* ---- programs (replace numba_resample + Dataset.power/spline chains) ------------------- */

/* Lower a descriptor for a raster of n_cells cells on the current device.  `target_stripes`
 * == 0 lets the library choose how many time stripes to cut (enough CTAs to fill 148 SMs);
 * < 0 means "the library's choice, but at least -target_stripes" (a raster that is streamed from
 * the host wants stripes of a few copy chunks, so that stripes can start before the copy ends). */
int agf_program_create(agf_program_t **out, const agf_program_desc_t *desc, int64_t n_cells,
                       int32_t target_stripes);
/* Destroy a handle.  The launches that used it must have completed (synchronise their stream first): its
 * device tables go back to a stream-ordered pool and may be handed to the next program at once. */
int agf_program_destroy(agf_program_t *prog);
/* Host-only planning (no device needed): validates the descriptor, picks the kernel
 * instantiation and cuts the time axis into stripes.  stripes_out (may be NULL) receives
 * (g1_begin, g1_end, g2_first, rec0) per stripe.  Used by agf_program_create. */
int agf_program_plan(const agf_program_desc_t *desc, int64_t n_cells, int32_t target_stripes,
                     int32_t sm_count, int32_t *stripes_out, int32_t max_stripes,
                     int32_t *n_stripes, int32_t *n_recs, int32_t *kernel_lanes,
                     int32_t *kernel_slots, int32_t *kernel_diag);
int agf_program_info(const agf_program_t *prog, agf_program_info_t *info);
/* rows [row_begin, row_end) of the time axis covered by stripe s */
int agf_program_stripe_rows(const agf_program_t *prog, int32_t stripe, int64_t *row_begin,
                            int64_t *row_end);

/* Run stripes [stripe_begin, stripe_end) of the fused temporal kernel.  d_x points at the
 * raster row `row0` (so a streamed chunk can be passed on its own), `ld` is the row stride in
 * elements.  Single-level programs write X / V directly; two-level programs write partial
 * records that agf_temporal_finalize merges -- unless the program has a single stripe
 * (info.direct_out), in which case there is nothing to merge and this call writes X / V.
 * Several programs of one call may share X / V: column c of this program goes to
 * X[:, :, cols[c].dst] of an X with `out_ncols` columns, and with valid_and != 0 its validity
 * is AND-ed into V instead of overwriting it. */
int agf_temporal_run(const agf_program_t *prog, const void *d_x, int64_t ld, int64_t row0,
                     int32_t stripe_begin, int32_t stripe_end, double *d_partial, void *d_out,
                     uint8_t *d_valid, int32_t out_ncols, int32_t valid_and, uintptr_t stream);
/* Merge partial records in stripe order, apply mean division / dtype rounding / trailing
 * transforms, write X and V.  No-op for single-level programs. */
int agf_temporal_finalize(const agf_program_t *prog, const double *d_partial, void *d_out,
                          uint8_t *d_valid, int32_t out_ncols, int32_t valid_and, uintptr_t stream);

/* ---- CSR weights + weighted regional average (replace _weight_triplets/_scatter_block) -- */

/* d_row_ptr int32[n_regions+1], d_cell_idx int32[nnz] (raster memory order), d_w fp64[nnz]:
 * caller-owned device arrays that must outlive the handle. */
int agf_csr_create(agf_csr_t **out, int32_t n_regions, int64_t n_cells, int64_t nnz,
                   const int32_t *d_row_ptr, const int32_t *d_cell_idx, const double *d_w);
int agf_csr_destroy(agf_csr_t *csr);

/* P[r, g, c] = sum_e w_e X[g, c, cell_e] V[g, cell_e] / sum_e w_e V[g, cell_e]  (NaN if the
 * denominator is 0), aggfly/aggregate/spatial.py:114-133.  Also writes the denominators
 * D[r, g] when d_den != NULL. */
int agf_spmm_run(const agf_csr_t *csr, const void *d_x, int32_t x_dtype, const uint8_t *d_valid,
                 int64_t n_groups, int32_t n_cols, double *d_panel, double *d_den,
                 uintptr_t stream);

/* ---- temporal scan + regional average in ONE kernel (daily / many-period panels) ------------------------ */

/* For panels with many periods the per-cell columns X are as large as the raster (hourly -> daily: 21.6 GB written
 * and read back next to a 36.4 GB raster).  agf_temporal_regional_run produces the panel of
 *     numba_resample (aggfly/aggregate/nb_kernels.py:253-305)  ->  _scatter_block + divide
 *                                                    (aggfly/aggregate/spatial.py:114-133, 181-186)
 * without materialising X: cells are scanned in 8 x 32 (lat x lon) tiles, every period's columns are contracted
 * with the weights inside the tile (entries in weights-frame order); a region that lies inside one tile gets its
 * panel row from that tile -- the same bits as agf_temporal_run + agf_spmm_run -- and a region that straddles tiles
 * gets per-tile partial sums that a second small kernel adds in ascending tile order (rel ~1e-16 from the two-kernel
 * path: one sum re-associated).  No atomics: results are bit-identical from run to run.  Tiles without a weighted
 * cell are never read.  Covered: single-level float32 programs whose periods have 24 rows (hourly -> date). */

/* A CSR lowered onto the cell tiles of a n_lat x n_lon grid (cell = lat * n_lon + lon, the raster's memory order).
 * row_ptr / cell_idx / w are HOST arrays (the arrays agf_csr_create takes on the device); the handle owns its
 * device tables (about 16 bytes per entry). */
typedef struct agf_rplan agf_rplan_t;
typedef struct {
    int32_t n_tiles, n_active_tiles; /* tiles of the grid / tiles that hold at least one entry (only those are read) */
    int32_t n_slots;                 /* (tile, region) pairs */
    int32_t max_slots_per_tile;
    int64_t n_entries;
    int32_t tile_lat, tile_lon;      /* 8, 32 */
    int32_t n_empty_regions;         /* regions without any entry on this grid (their rows are NaN) */
    int32_t n_partial_rows;          /* slots of regions that straddle tiles (one scratch row per slot and period) */
    int64_t table_bytes;
} agf_rplan_info_t;
int agf_rplan_create(agf_rplan_t **out, int32_t n_regions, int32_t n_lat, int32_t n_lon, int64_t nnz,
                     const int32_t *row_ptr, const int32_t *cell_idx, const double *w);
int agf_rplan_destroy(agf_rplan_t *plan);
/* Host-only twin of agf_rplan_create (no device needed): fills *info and copies the tables into the non-NULL
 * outputs -- tile_ids[n_active_tiles], tile_slot_ptr[n_active_tiles + 1], slot_region[n_slots],
 * slot_ent_ptr[n_slots + 1], entry_cell[n_entries] (cell inside its tile: (lat % 8) * 32 + lon % 32),
 * entry_w[n_entries], region_slot_ptr[n_regions + 1], region_slots[n_slots], slot_dst[n_slots] (region r >= 0 when
 * the slot holds the whole region, else -(partial row + 1)).  Call once with NULL outputs for the sizes. */
int agf_rplan_tables(int32_t n_regions, int32_t n_lat, int32_t n_lon, int64_t nnz, const int32_t *row_ptr,
                     const int32_t *cell_idx, const double *w, agf_rplan_info_t *info, int32_t *tile_ids,
                     int32_t *tile_slot_ptr, int32_t *slot_region, int32_t *slot_ent_ptr, int32_t *entry_cell,
                     double *entry_w, int32_t *region_slot_ptr, int32_t *region_slots, int32_t *slot_dst);
int agf_rplan_info(const agf_rplan_t *plan, agf_rplan_info_t *info);
/* Host-only self check of the balanced walk tables the kernel variant with `lps` (4 | 8 | 16) lanes per slot uses: builds
 * them from the CSR and verifies that the lane groups' segments visit every entry of every (tile, region) slot exactly
 * once and in the slot's own order.  stats (int64[7], may be NULL): segments, padded entries, largest group load,
 * mean group load, active tiles, most segments / padded entries in one tile. */
int agf_rplan_check_segments(int32_t n_regions, int32_t n_lat, int32_t n_lon, int64_t nnz, const int32_t *row_ptr,
                             const int32_t *cell_idx, const double *w, int32_t lps, int64_t *stats);

typedef struct {
    int32_t supported;        /* 0: this program has no regional instantiation -- use agf_temporal_run + agf_spmm_run */
    int32_t lanes_per_slot;
    int32_t kernel_lanes, smem_bytes, ctas_per_sm;
    int32_t pad_;
    int64_t workspace_bytes;  /* size of d_workspace for a panel of panel_groups periods */
} agf_regional_info_t;
int agf_temporal_regional_plan(const agf_program_t *prog, const agf_rplan_t *plan, int64_t panel_groups,
                               agf_regional_info_t *info);
/* P[r, g, cols[c].dst] for g in [group_begin, group_end) (and D[r, g] when d_den != NULL) of a panel
 * P[n_regions, panel_groups, out_ncols].  d_x points at raster row `row0` (row stride ld elements) and must hold
 * the rows of the period range; d_workspace is caller-owned scratch of workspace_bytes (partial rows of straddling
 * regions, indexed by period: ranges of one panel may share it).  Stream-ordered, no host synchronisation. */
int agf_temporal_regional_run(const agf_program_t *prog, const agf_rplan_t *plan, const void *d_x, int64_t ld,
                              int64_t row0, int64_t group_begin, int64_t group_end, void *d_workspace,
                              int64_t workspace_bytes, double *d_panel, int64_t panel_groups, int32_t out_ncols,
                              double *d_den, uintptr_t stream);

/* V[g, cell] = 1 iff no column of X[g, :, cell] is NaN -- for callers that bring their own
 * temporally-reduced X (aggregate_space / SpatialAggregator, aggfly/aggregate/spatial.py:114-119). */
int agf_valid_mask_run(const void *d_x, int32_t x_dtype, int64_t n_groups, int32_t n_cols,
                       int64_t n_cells, uint8_t *d_valid, uintptr_t stream);

/* out[i] = f(in[i]) for a whole series of n values: AGF_XF_POWI / AGF_XF_POW / AGF_XF_SPLINE2, or, when
 * d_other != NULL, in[i] * other[i] (the reference's `inter` transform, aggfly/dataset/dataset.py:483-563).
 * For transforms that cannot be fused into a program (applied to the raster before the first aggregate
 * step, two in a row, interactions).  out_dtype follows NumPy's promotion, decided by the caller; d_valid
 * (may be NULL) receives !isnan(out[i]).  pre[0..n_pre) (host array, may be NULL) is the preprocess chain
 * applied to in[i] first, in in_dtype -- for transforms that read the raster of a Dataset(preprocess=...). */
int agf_elementwise_run(const void *d_in, int32_t in_dtype, void *d_out, int32_t out_dtype, int64_t n,
                        int32_t xform, double xparam, const void *d_other, int32_t other_dtype,
                        uint8_t *d_valid, int32_t n_pre, const agf_pre_t *pre, uintptr_t stream);

/* ---- chunked sources: storage chunk -> time-major raster ------------------------------------------ */

/* Place one decoded storage chunk (zarr / netCDF chunk, any axis order) into the device raster:
 *     dst[(t0 + t) * ld + (y0 + y) * n_lon + (x0 + x)] = decode(src[t * st + y * sy + x * sx])
 * for t < nt, y < ny, x < nx; st / sy / sx are ELEMENT strides of the stored chunk (a time-contiguous
 * store written by the reference's dataset_to_zarr, aggfly/dataset/zarr_convert.py:31-47, has st = 1;
 * the reference transposes such arrays to (time, y, x) in dask, aggfly/aggregate/nb_kernels.py:280).
 * decode: if has_fill and src == fill -> NaN; if packed -> (double)src * scale + offset (CF
 * scale_factor / add_offset, what xarray's decode_cf does on open, aggfly/dataset/dataset.py:700-707);
 * then converted to dst_dtype (AGF_F32 | AGF_F64).  src_dtype: any AGF_* dtype; float64 sources need a
 * float64 raster.  d_src and d_dst are caller-owned device buffers; dst_rows = rows of the raster d_dst points at
 * (t0 + nt must not exceed it: a wrong time offset is rejected instead of writing past the raster, ABI 5). */
int agf_tile_place_run(const void *d_src, int32_t src_dtype, int64_t nt, int64_t ny, int64_t nx,
                       int64_t st, int64_t sy, int64_t sx, void *d_dst, int32_t dst_dtype, int64_t ld,
                       int64_t n_lon, int64_t t0, int64_t y0, int64_t x0, int32_t packed, double scale,
                       double offset, int32_t has_fill, double fill, int64_t dst_rows, uintptr_t stream);

/* ---- compressed chunks: Blackwell decompression engine --------------------------------------------- */

/* Bit mask of the algorithms the device's hardware decompression engine offers (1 deflate, 2 snappy,
 * 4 lz4; 0: none / driver too old) and the largest number of bytes one operation may read or write. */
int agf_decompress_caps(int32_t *algo_mask, int64_t *max_length);

/* Inflate n independent raw-LZ4 streams with the decompression engine, stream-ordered:
 *     d_dst[dst_off[i] .. + dst_len[i]) = lz4_block_decode(d_src[src_off[i] .. + src_len[i]))
 * The offset / length arrays are HOST arrays; d_actual[n] (device, uint32) receives the bytes each
 * operation produced, for the caller to compare with dst_len.  d_src / d_dst / d_actual must come from
 * cudaMalloc (torch's default allocator qualifies).  This is what Blosc-compressed zarr chunks need: one
 * stream per (block, byte plane); see aggfly_b200/zarrio.py.  AGF_E_UNSUPPORTED without an LZ4 engine. */
int agf_decompress_lz4_run(const void *d_src, const int64_t *src_off, const int64_t *src_len, void *d_dst,
                           const int64_t *dst_off, const int64_t *dst_len, int64_t n, uint32_t *d_actual,
                           uintptr_t stream);

/* d_dst[dst_off[i] .. + len[i]) = d_src[src_off[i] .. + len[i]) for n byte segments; d_table is a DEVICE
 * array int64[3 * n] = src_off[n], dst_off[n], len[n].  For the streams Blosc stored uncompressed (the noisy
 * low-mantissa byte planes of float data), which bypass the decompression engine. */
int agf_copy_segments_run(const void *d_src, void *d_dst, const int64_t *d_table, int64_t n, uintptr_t stream);

/* Undo Blosc's byte shuffle of a decoded chunk (blocks of `blocksize` bytes, the last one shorter):
 * out[b * blocksize + i * typesize + j] = in[b * blocksize + j * n_b + i].  typesize 2, 4 or 8. */
int agf_unshuffle_run(const void *d_src, void *d_dst, int64_t nbytes, int32_t typesize, int64_t blocksize,
                      uintptr_t stream);

/* ---- weights builder geometry (host only; replaces the GEOS work of calculate_weights) ---------- */

/* For every region the fraction of each grid cell's rectangle that it covers -- what
 * aggfly/weights/grid_weights.py:238-421 obtains from two buffered centroid joins plus shapely
 * intersections of the border cells (interior cells: exactly 1; cells with no overlap: absent).
 *   regions   region r owns rings [region_ring_ptr[r], region_ring_ptr[r+1]); ring k owns vertices
 *             [ring_ptr[k], ring_ptr[k+1]) of xy (x0, y0, x1, y1, ...; closing vertex optional).
 *             Holes must be oriented opposite to their shell (shapefile / OGC convention).
 *   grid      cell (i, j) is the rectangle lon[j] +- dlon/2, lat[i] +- dlat/2 (centres in any
 *             monotonic order); cell_id = i * n_lon + j (aggfly/dataset/grid.py:74-80).
 * Pairs come out grouped by region, cell_id ascending.  Pure host code, no device needed. */
typedef struct agf_overlap agf_overlap_t;
int agf_overlap_create(agf_overlap_t **out, int32_t n_regions, const int64_t *region_ring_ptr,
                       const int64_t *ring_ptr, const double *xy, int32_t n_lon, const double *lon,
                       double dlon, int32_t n_lat, const double *lat, double dlat, int64_t *n_pairs);
int agf_overlap_fetch(const agf_overlap_t *h, int32_t *region, int64_t *cell_id, double *fraction);
int agf_overlap_destroy(agf_overlap_t *h);

#ifdef __cplusplus
}
#endif
#endif /* AGGFLY_B200_H */
