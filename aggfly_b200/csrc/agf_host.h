// agf_host.h -- host-side declarations shared by the translation units of libaggfly_b200.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "agf_kernels.cuh"

struct agf_program {
    agf_program_desc_t desc;  // bounds pointers are NOT kept (copied below)
    std::vector<int32_t> b1, b2;
    std::vector<agf::Stripe> stripes;
    std::vector<int32_t> g2_rec_ptr, g2_rec_idx;
    int64_t n_cells = 0;
    int n_recs = 0;
    int device = -1;
    int kernel_lanes = 0, kernel_slots = 0, kernel_diag = 0;  // instantiation picked for the TMA variant
    unsigned kernel_kinds = 0;
    unsigned kinds = 0;       // lane kinds the program uses
    unsigned slot_kinds = 0;  // slot kinds the program uses
    int n_bin_slots = 0;      // slots of kind SK_BINS
    int uniform_gl = 0;       // rows per level-1 group when every group has the same size, else 0
    int max_group_rows = 0;   // longest level-1 group
    int direct_out = 0;       // two-level, one stripe, no empty level-2 group: agf_temporal_run writes X / V itself
    int diag_ok = 0;     // columns (single-level) / slots (two-level) map 1:1 onto lanes
    int need_nan = 0, need_cnt = 0, has_sine = 0;
    // device copies
    int *d_b1 = nullptr, *d_b2 = nullptr, *d_g2_rec_ptr = nullptr, *d_g2_rec_idx = nullptr;
    agf::Stripe *d_stripes = nullptr;
};

int agf_fail(int code, const char *fmt, ...);
int agf_cuda_fail(cudaError_t e, const char *what);

#define CU(call)                                                \
    do {                                                        \
        cudaError_t e_ = (call);                                \
        if (e_ != cudaSuccess) return agf_cuda_fail(e_, #call); \
    } while (0)

// Encode a 2-D tiled tensor map over the raster view [n_rows, n_cells] (row stride ld elements),
// box = [box_rows, TMA_CW].  Returns 0, or AGF_E_UNSUPPORTED when the view cannot be described
// (alignment) -- the caller then uses the direct-load kernel.
int agf_make_tensor_map(agf::TensorMap *out, const void *base, int elem_size, uint64_t n_cells, uint64_t n_rows,
                        uint64_t ld, int box_rows);
bool agf_tma_eligible(const void *base, int elem_size, uint64_t ld);

struct K1Launch {
    const agf_program *p;
    const void *d_x;
    int64_t ld, row0;
    int s0, s1;
    double *d_partial;
    void *d_out;
    uint8_t *d_valid;
    int ncols, vand;
    cudaStream_t stream;
    int use_tma;
};

// ---- K1R: temporal scan + regional average in one kernel (agf_regional.cuh, agf_rplan.cu) ----
// A CSR lowered onto 8 x 32 cell tiles of a lat x lon grid: per-tile slots (one per region the tile touches) and their
// entries, plus for every region the slots that contribute to it.  Device tables are owned by the handle.
struct agf_rplan {
    int n_regions = 0, n_lat = 0, n_lon = 0;
    int tiles_x = 0, tiles_y = 0, n_active = 0, n_gslots = 0, max_slots = 0;
    int n_partial_rows = 0;  // slots of regions that straddle tiles
    int n_multi = 0;         // such regions
    int64_t n_entries = 0;
    int n_empty_regions = 0;
    int device = -1;
    int *d_tile_ids = nullptr, *d_tile_slot_ptr = nullptr, *d_slot_dst = nullptr, *d_slot_ent_ptr = nullptr;
    int *d_region_slot_ptr = nullptr, *d_region_slots = nullptr, *d_multi_regions = nullptr;
    void *d_entries = nullptr;
    int64_t table_bytes = 0;
    // Balanced walk tables, one set per lanes-per-slot variant of the kernel (256 / LPS lane groups per tile), built on
    // first use from the host copies below (agf_rplan_segments).
    struct SegTables {
        int ng = 0;                       // lane groups per tile (0: not built yet)
        int max_segs = 0, max_pent = 0;   // largest tile: segments / padded entries
        int64_t n_segs = 0, n_pent = 0;
        int *d_tile_seg_ptr = nullptr;    // [n_active + 1] segments of the tile
        int *d_tile_pent_ptr = nullptr;   // [n_active + 1] padded entries of the tile
        int *d_grp = nullptr;             // [n_active][ng + 1][2]: (first segment record, first padded entry), tile-relative
        int *d_seg = nullptr;             // [n_segs][2]: (end of the segment in the tile's padded entries, segment id), group-major
        int *d_slot_q = nullptr;          // [n_gslots][2]: segment ids [q0, q1) of the slot (slot-major numbering)
        void *d_pent = nullptr;           // [n_pent] RgEntry, group-major; pads have w = 0 and cell = -1
    };
    mutable SegTables seg[3];             // LPS 4, 8, 16
    std::vector<int32_t> h_tile_slot_ptr, h_slot_ent_ptr;
    std::vector<double> h_ent_w;
    std::vector<int32_t> h_ent_cell;
};
// Builds (once) and returns the balanced walk tables for `lps` lanes per slot; nullptr + agf_fail on error.
const agf_rplan::SegTables *agf_rplan_segments(const agf_rplan *plan, int lps);

struct RegionalLaunch {
    K1Launch k;              // program, raster view, stream (stripe / X / V fields unused)
    const agf_rplan *plan;
    int64_t g_begin, g_end;  // periods of this launch
    void *d_workspace;
    int64_t workspace_bytes;
    double *d_panel, *d_den;
    int64_t G;               // periods of the whole panel (row pitch)
    int out_ncols;
};

struct RegionalChoice {
    int lanes, typed_bins, lps;   // instantiation: kernel lanes, NB, lanes per slot
    int smem_bytes, ctas_per_sm;
};

// mode 0: launch; mode 1: only report the instantiation in *choice.  Returns 1 if nothing fits.
int agf_k1_f32_regional(const RegionalLaunch &a, int mode, RegionalChoice *choice, int *rc);
int agf_make_tensor_map3(agf::TensorMap *out, const void *base, int elem_size, uint64_t n_lon, uint64_t n_lat,
                         uint64_t n_rows, uint64_t ld, int box_lon, int box_lat, int box_rows);

struct K1Choice {
    int lanes, slots, diag;
    unsigned kinds;
    int typed_bins;  // NB of the instantiation (-1: general slots)
    int uniform_gl;  // GL of the instantiation (0: general group bounds)
};

// One per translation unit (agf_k1_*.cu).  mode 0: launch the first instantiation of that unit
// that fits the program; mode 1: only report it in *choice.  Returns 1 if nothing in the unit fits.
int agf_k1_f32_tma_single(const K1Launch &a, int mode, K1Choice *choice, int *rc);
int agf_k1_f32_tma_two(const K1Launch &a, int mode, K1Choice *choice, int *rc);
int agf_k1_f32_tma_uni(const K1Launch &a, int mode, K1Choice *choice, int *rc);
int agf_k1_f32_ldg(const K1Launch &a, int mode, K1Choice *choice, int *rc);
int agf_k1_f64_tma(const K1Launch &a, int mode, K1Choice *choice, int *rc);
int agf_k1_f64_ldg(const K1Launch &a, int mode, K1Choice *choice, int *rc);
