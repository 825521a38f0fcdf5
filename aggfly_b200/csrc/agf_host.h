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
    int kernel_lanes = 0, kernel_slots = 0, kernel_diag = 0;
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
};

// one per (dtype, part) translation unit; return 1 if the instantiation is not in that unit
int agf_k1_f32_single(const K1Launch &a, int *rc);
int agf_k1_f32_two(const K1Launch &a, int *rc);
int agf_k1_f64_single(const K1Launch &a, int *rc);
int agf_k1_f64_two(const K1Launch &a, int *rc);
