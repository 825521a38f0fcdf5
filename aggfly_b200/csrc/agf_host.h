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
