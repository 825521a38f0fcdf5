// k1_sweep.cu -- tuning harness for the fused temporal kernel (development tool, not product).
//
// Runs the C3 program (hourly f32 -> daily mean -> 13 yearly bins + sum of x, x^2) over a
// synthetic [T, cells] raster with several compile-time configurations of agf_k1_tma, next to
// two probes that bound what the memory system can deliver for this access pattern:
//   probe_read   : plain grid-stride float4 read + sum (read-only STREAM)
//   probe_ring   : the same TMA ring as the kernel with a trivial consumer (f32 add)
// Prints one line per variant: ms, GB/s.  Build: see tools/Makefile.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../aggfly_b200/csrc/agf_kernels.cuh"

using namespace agf;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TensorMap make_map(const float *base, uint64_t n_cells, uint64_t n_rows, int box_cols, int box_rows, int l2promo) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
    EncodeTiledFn fn = (EncodeTiledFn)ptr;
    TensorMap tm;
    cuuint64_t dims[2] = {n_cells, n_rows};
    cuuint64_t strides[1] = {n_cells * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn((CUtensorMap *)&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2promo,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "encode failed %d\n", (int)r);
        exit(1);
    }
    return tm;
}

__global__ void fill_kernel(float *x, size_t n, int n_cells) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned h = (unsigned)(i * 2654435761u) ^ (unsigned)(i >> 32);
        h ^= h >> 15;
        h *= 0x2c1b3c6du;
        h ^= h >> 12;
        const size_t t = i / n_cells;
        x[i] = 12.0f + 20.0f * __sinf((float)(t % 8760) * (6.2831853f / 8760.f)) + 6.0f * __sinf((float)(t % 24) * 0.2618f) +
               (float)(h & 0xffff) * (8.0f / 65536.0f);
    }
}

__global__ void __launch_bounds__(512) probe_read(const float4 *x, size_t n4, float *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = x[i], b = x[i + stride], c = x[i + 2 * stride], d = x[i + 3 * stride];
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
    }
    for (; i < n4; i += stride) {
        float4 a = x[i];
        acc += a.x + a.y + a.z + a.w;
    }
    if (acc == 123.456f) out[0] = acc;
}

// TMA ring with a trivial consumer: what the ring alone can stream
template <int TT, int STAGES, int MINB>
__global__ void __launch_bounds__(TMA_THREADS, MINB) probe_ring(const __grid_constant__ TensorMap tmap, int n_rows, float *out) {
    constexpr int TILE = TT * TMA_CW * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + STAGES * TILE);
    uint64_t *empty = full + STAGES;
    const int n_tiles = (n_rows + TT - 1) / TT;
    const int cell0 = blockIdx.x * TMA_CW;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], TMA_CW / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= TMA_CW) {
        if (threadIdx.x == TMA_CW) {
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % STAGES;
                if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) - 1) & 1);
                mbar_expect_tx(&full[s], TILE);
                tma_load_2d(smem_raw + s * TILE, &tmap, cell0, i * TT, &full[s]);
            }
        }
        return;
    }
    float acc = 0.f;
    for (int i = 0; i < n_tiles; ++i) {
        const int stg = i % STAGES;
        mbar_wait(&full[stg], (i / STAGES) & 1);
        const float *col = tiles + (size_t)stg * (TILE / 4) + threadIdx.x;
#pragma unroll
        for (int r = 0; r < TT; ++r) acc += col[r * TMA_CW];
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stg]);
    }
    if (acc == 123.456f) out[0] = acc;
}

struct Ctx {
    float *x;
    int T, n_cells;
    int *b1, *b2;
    Stripe *stripes;
    double *partial;
    cudaEvent_t e0, e1;
    int reps;
    const char *filter;
};

template <typename F>
static float time_it(Ctx &c, F launch) {
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(c.e0));
    for (int i = 0; i < c.reps; ++i) launch();
    CK(cudaEventRecord(c.e1));
    CK(cudaEventSynchronize(c.e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, c.e0, c.e1));
    return ms / c.reps;
}

static void report(const char *name, Ctx &c, float ms) {
    const double bytes = (double)c.T * c.n_cells * 4.0;
    printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / (ms * 1e-3) / 1e9);
    fflush(stdout);
}

template <int NS, int NB, int TT, int STAGES, int MINB, int GL = 0>
static void run_k1(Ctx &c, const char *name, int n_bins, int n_pow) {
    if (c.filter && !strstr(name, c.filter)) return;
    K1Params<float, 1, NS> kp;
    memset(&kp, 0, sizeof(kp));
    kp.x = c.x;
    kp.ld = c.n_cells;
    kp.row0 = 0;
    kp.n_cells = c.n_cells;
    kp.stripe0 = 0;
    kp.b1 = c.b1;
    kp.b2 = c.b2;
    kp.stripes = c.stripes;
    kp.partial = c.partial;
    kp.n_lanes = 1;
    kp.n_slots = n_bins + n_pow;
    kp.lanes[0].calc = AGF_CALC_MEAN;
    int j = 0;
    for (int b = 0; b < NB; ++b, ++j) {
        SlotP &S = kp.slots[j];
        S.calc = AGF_CALC_BINS;
        S.ip = 1;
        if (b < n_bins) {
            S.t0 = S.flo = -20.0 + 5.0 * b;
            S.t1 = S.fhi = -15.0 + 5.0 * b;
            S.dst = b;
        } else {
            S.t0 = S.flo = INFINITY;
            S.t1 = S.fhi = -INFINITY;
            S.dst = -1;
        }
    }
    for (int q = 0; q < NS - NB; ++q, ++j) {
        SlotP &S = kp.slots[j];
        S.calc = AGF_CALC_SUM;
        S.x_f64 = 1;
        if (q < n_pow) {
            S.xform = q == 0 ? AGF_XF_NONE : AGF_XF_POWI;
            S.ip = q + 1;
            S.xparam = q + 1;
            S.dst = n_bins + q;
        } else {
            S.ip = 1;
            S.dst = -1;
        }
    }
    TensorMap tm = make_map(c.x, c.n_cells, c.T, TMA_CW, TT, 2);
    constexpr int smem = STAGES * TT * TMA_CW * 4 + 2 * STAGES * 8;
    void (*kern)(const K1Params<float, 1, NS>, const TensorMap);
    if constexpr (GL > 0)
        kern = agf_k1_tma_uni<float, 1, NS, false, KIND_SUM, NB, GL, TT, STAGES, MINB>;
    else
        kern = agf_k1_tma<float, 1, NS, false, KIND_SUM, NB, TT, STAGES, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid((c.n_cells + TMA_CW - 1) / TMA_CW, 1);
    float ms = time_it(c, [&] { kern<<<grid, TMA_THREADS, smem>>>(kp, tm); });
    report(name, c, ms);
}

template <int TT, int STAGES, int MINB>
static void run_ring(Ctx &c, const char *name, float *out) {
    if (c.filter && !strstr(name, c.filter)) return;
    TensorMap tm = make_map(c.x, c.n_cells, c.T, TMA_CW, TT, 2);
    constexpr int smem = STAGES * TT * TMA_CW * 4 + 2 * STAGES * 8;
    auto kern = probe_ring<TT, STAGES, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid((c.n_cells + TMA_CW - 1) / TMA_CW);
    float ms = time_it(c, [&] { kern<<<grid, TMA_THREADS, smem>>>(tm, c.T, out); });
    report(name, c, ms);
}

int main(int argc, char **argv) {
    Ctx c;
    c.T = argc > 1 ? atoi(argv[1]) : 8760;
    c.n_cells = argc > 2 ? atoi(argv[2]) : 721 * 1440;
    c.reps = argc > 3 ? atoi(argv[3]) : 5;
    c.filter = argc > 4 ? argv[4] : nullptr;
    const size_t n = (size_t)c.T * c.n_cells;
    CK(cudaMalloc(&c.x, n * 4));
    fill_kernel<<<148 * 16, 512>>>(c.x, n, c.n_cells);
    CK(cudaDeviceSynchronize());
    const int G1 = c.T / 24;
    std::vector<int> b1(G1 + 1), b2 = {0, G1};
    for (int g = 0; g <= G1; ++g) b1[g] = g * 24;
    Stripe st{0, G1, 0, 0};
    CK(cudaMalloc(&c.b1, b1.size() * 4));
    CK(cudaMalloc(&c.b2, 8));
    CK(cudaMalloc(&c.stripes, sizeof(Stripe)));
    CK(cudaMemcpy(c.b1, b1.data(), b1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.b2, b2.data(), 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.stripes, &st, sizeof(st), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c.partial, (size_t)32 * c.n_cells * 8));
    CK(cudaEventCreate(&c.e0));
    CK(cudaEventCreate(&c.e1));
    float *out;
    CK(cudaMalloc(&out, 4));
    printf("raster %d x %d f32 = %.2f GB, reps %d\n", c.T, c.n_cells, n * 4 / 1e9, c.reps);

    if (!c.filter) {
        float ms = time_it(c, [&] { probe_read<<<148 * 4, 512>>>((const float4 *)c.x, n / 4, out); });
        report("probe_read float4 (148x4 CTAs x 512)", c, ms);
    }
    run_ring<24, 3, 3>(c, "probe_ring TT=24 ST=3 B=3", out);
    run_ring<8, 9, 3>(c, "probe_ring TT=8 ST=9 B=3", out);

#define C3(TT, ST, B) run_k1<18, 16, TT, ST, B>(c, "k1 general C3 (13 bins + 2 pow) NS=18 NB=16 TT=" #TT " ST=" #ST " B=" #B, 13, 2)
#define U3(ST, B) run_k1<18, 16, 24, ST, B, 24>(c, "k1 uni24 C3 (13 bins + 2 pow) NS=18 NB=16 ST=" #ST " B=" #B, 13, 2)
    C3(24, 3, 3);
    U3(4, 2);
    U3(3, 3);
    U3(2, 4);
    U3(8, 1);
    run_k1<32, 24, 24, 4, 2, 24>(c, "k1 uni24 24 bins + 4 pow NS=32 NB=24 ST=4 B=2", 24, 4);
    run_k1<32, 24, 24, 3, 3, 24>(c, "k1 uni24 24 bins + 4 pow NS=32 NB=24 ST=3 B=3", 24, 4);
    run_k1<16, 16, 24, 3, 3, 24>(c, "k1 uni24 13 bins only NS=16 NB=16 ST=3 B=3", 13, 0);
    run_k1<4, 0, 24, 3, 3, 24>(c, "k1 uni24 2 pow only NS=4 NB=0 ST=3 B=3", 0, 2);
    run_k1<1, 0, 24, 3, 3, 24>(c, "k1 uni24 mean->sum NS=1 ST=3 B=3", 0, 1);
    run_k1<1, 0, 24, 3, 3>(c, "k1 general mean->sum NS=1 TT=24 ST=3 B=3", 0, 1);
    return 0;
}
