// agf_tile.cu -- storage chunks of a chunked raster (zarr / netCDF chunk) -> the time-major device raster.
//
// The reference reads its rasters lazily through xarray/dask (aggfly/dataset/dataset.py:636-740); its own
// conversion tool writes "time-contiguous" zarr stores whose arrays are laid out (latitude, longitude,
// time) in chunks of [s, s, T] (aggfly/dataset/zarr_convert.py:31-47, :109), i.e. TIME is the fastest
// axis of every chunk, while the temporal kernel scans a raster x[T, lat * lon] with the cell index
// fastest (nb_kernels.py:280 transposes to (time, y, x) for the same reason).  Transposing 36 GB per year
// on the host would cost far more than the PCIe copy, so the decoded chunk is copied to the device as it
// is stored and this kernel places it:
//
//     dst[(t0 + t) * ld + (y0 + y) * n_lon + (x0 + x)] = decode(src[t * st + y * sy + x * sx])
//
// for any permutation of the chunk's axes (st / sy / sx are element strides of the stored chunk).
//   * sx == 1  (time-major chunks, e.g. [24, lat, lon]): straight row copy, coalesced on both sides;
//   * otherwise the stride-1 axis f (time or latitude) is exchanged with x through a 32 x 33 shared-memory
//     tile: reads run along f, writes along x, both in full 128-byte lines.
// decode(): CF packing (value * scale + offset, evaluated in double like NumPy on a float64 result) for
// integer sources, and `fill` -> NaN masking, so packed int16 ERA5 files cross PCIe at 2 bytes per value.
// HBM-bound byte work: bytes = extent * (sizeof(src) + sizeof(dst)); no tensor cores.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include <cuda.h>

#include "agf_host.h"

namespace {

struct TileArgs {
    long long nt, ny, nx;      // extent of the placed block
    long long st, sy, sx;      // source strides (elements)
    long long ld, n_lon;       // destination: row pitch (elements), columns per latitude row
    long long t0, y0, x0;      // destination offset
    double scale, offset, fill;
    int has_fill, packed;
};

template <typename TS, typename TD>
__device__ __forceinline__ TD tile_decode(TS v, const TileArgs &a) {
    if (a.has_fill) {
        // a NaN fill value never compares equal; NaN sources stay NaN through the conversion below
        if ((double)v == a.fill) return (TD)__longlong_as_double(0x7ff8000000000000LL);
    }
    if (a.packed) return (TD)((double)v * a.scale + a.offset);
    return (TD)v;
}

// sx == 1 (or no unit-stride axis at all): one thread per destination element, x fastest.
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) agf_tile_rows(const TS *__restrict__ src, TD *__restrict__ dst, TileArgs a) {
    const long long n = a.nt * a.ny * a.nx;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const long long x = i % a.nx, r = i / a.nx, y = r % a.ny, t = r / a.ny;
        const TS v = src[t * a.st + y * a.sy + x * a.sx];
        dst[(a.t0 + t) * a.ld + (a.y0 + y) * a.n_lon + (a.x0 + x)] = tile_decode<TS, TD>(v, a);
    }
}

// Whole raster rows stored contiguously (a row chunk of a time-major source, dataset.PackedRaster): the source and the
// destination are the same linear index apart, so no div / mod per element; 16 bytes of source per thread and trip.
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) agf_tile_linear(const TS *__restrict__ src, TD *__restrict__ dst, long long n, TileArgs a) {
    constexpr int V = 16 / (int)sizeof(TS);
    const long long nv = n / V;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += step) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(src) + i);
        TS v[V];
        memcpy(v, &raw, 16);
        TD o[V];
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = tile_decode<TS, TD>(v[k], a);
        constexpr int OV = 16 / (int)sizeof(TD);   // destination elements per 16-byte store
#pragma unroll
        for (int k = 0; k < V / OV; ++k) {
            uint4 w;
            memcpy(&w, o + k * OV, 16);
            reinterpret_cast<uint4 *>(dst + i * V)[k] = w;
        }
    }
    for (long long i = nv * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        dst[i] = tile_decode<TS, TD>(src[i], a);
}

// The stride-1 source axis is f in {t, y}; o is the remaining axis.  Block (32, 8) moves a 32 (f) x 32 (x)
// tile; blockIdx.z strides over o.
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) agf_tile_transpose(const TS *__restrict__ src, TD *__restrict__ dst, TileArgs a,
                                                          int f_is_t) {
    __shared__ TS tile[32][33];
    const long long nf = f_is_t ? a.nt : a.ny, no = f_is_t ? a.ny : a.nt;
    const long long so = f_is_t ? a.sy : a.st;
    const long long f0 = (long long)blockIdx.x * 32, x0 = (long long)blockIdx.y * 32;
    for (long long o = blockIdx.z; o < no; o += gridDim.z) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {                       // rows of the tile = x, threads along f
            const long long x = x0 + threadIdx.y + j, f = f0 + threadIdx.x;
            if (x < a.nx && f < nf) tile[threadIdx.y + j][threadIdx.x] = src[f + x * a.sx + o * so];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {                       // rows of the tile = f, threads along x
            const long long f = f0 + threadIdx.y + j, x = x0 + threadIdx.x;
            if (x < a.nx && f < nf) {
                const long long t = f_is_t ? f : o, y = f_is_t ? o : f;
                dst[(a.t0 + t) * a.ld + (a.y0 + y) * a.n_lon + (a.x0 + x)] =
                    tile_decode<TS, TD>(tile[threadIdx.x][threadIdx.y + j], a);
            }
        }
        __syncthreads();
    }
}

template <typename TS, typename TD>
int launch_tile(const void *d_src, void *d_dst, const TileArgs &a, int sms, cudaStream_t st) {
    const long long n = a.nt * a.ny * a.nx;
    const int f_is_t = a.st == 1 && a.nt > 1, f_is_y = a.sy == 1 && a.ny > 1;
    TD *dbase = (TD *)d_dst + a.t0 * a.ld;
    if (a.sx == 1 && a.sy == a.nx && a.st == a.ny * a.nx && a.x0 == 0 && a.y0 == 0 && a.nx == a.n_lon && a.ny * a.n_lon == a.ld &&
        sizeof(TS) <= sizeof(TD) && (uintptr_t)d_src % 16 == 0 && (uintptr_t)dbase % 16 == 0) {
        constexpr int V = 16 / (int)sizeof(TS);
        const unsigned blocks = (unsigned)std::min<long long>((n / V + 255) / 256 + 1, (long long)sms * 16);
        agf_tile_linear<TS, TD><<<blocks, 256, 0, st>>>((const TS *)d_src, dbase, n, a);
    } else if (a.sx == 1 || a.nx == 1 || (!f_is_t && !f_is_y)) {
        const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, (long long)sms * 32);
        agf_tile_rows<TS, TD><<<blocks, 256, 0, st>>>((const TS *)d_src, (TD *)d_dst, a);
    } else {
        const long long nf = f_is_t ? a.nt : a.ny, no = f_is_t ? a.ny : a.nt;
        dim3 grid((unsigned)((nf + 31) / 32), (unsigned)((a.nx + 31) / 32), (unsigned)std::min<long long>(no, 65535));
        if (grid.y > 65535) return agf_fail(AGF_E_INVALID, "tile wider than 2097120 columns");
        agf_tile_transpose<TS, TD><<<grid, dim3(32, 8), 0, st>>>((const TS *)d_src, (TD *)d_dst, a, f_is_t);
    }
    return 0;
}

}  // namespace

extern "C" int agf_tile_place_run(const void *d_src, int32_t src_dtype, int64_t nt, int64_t ny, int64_t nx,
                                  int64_t st, int64_t sy, int64_t sx, void *d_dst, int32_t dst_dtype, int64_t ld,
                                  int64_t n_lon, int64_t t0, int64_t y0, int64_t x0, int32_t packed, double scale,
                                  double offset, int32_t has_fill, double fill, int64_t dst_rows, uintptr_t stream) {
    if (!d_src || !d_dst) return agf_fail(AGF_E_INVALID, "null argument");
    if (nt < 0 || ny < 0 || nx < 0 || st < 0 || sy < 0 || sx < 0) return agf_fail(AGF_E_INVALID, "negative extent / stride");
    if (t0 < 0 || y0 < 0 || x0 < 0 || n_lon <= 0 || x0 + nx > n_lon || (y0 + ny) * n_lon > ld)
        return agf_fail(AGF_E_INVALID, "tile [%lld+%lld, %lld+%lld] does not fit a row of %lld x %lld cells (pitch %lld)",
                        (long long)y0, (long long)ny, (long long)x0, (long long)nx, (long long)(ld / n_lon),
                        (long long)n_lon, (long long)ld);
    if (dst_rows <= 0 || t0 + nt > dst_rows)
        return agf_fail(AGF_E_INVALID, "tile rows [%lld, %lld) do not fit a raster of %lld rows", (long long)t0, (long long)(t0 + nt),
                        (long long)dst_rows);
    if (dst_dtype != AGF_F32 && dst_dtype != AGF_F64) return agf_fail(AGF_E_INVALID, "destination dtype %d", dst_dtype);
    if (nt == 0 || ny == 0 || nx == 0) return 0;
    TileArgs a{nt, ny, nx, st, sy, sx, ld, n_lon, t0, y0, x0, scale, offset, fill, has_fill != 0, packed != 0};
    int dev = 0, sms = 148;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t s = (cudaStream_t)stream;
    int rc = AGF_E_INVALID;
    const bool d64 = dst_dtype == AGF_F64;
    switch (src_dtype) {
    case AGF_F32: rc = d64 ? launch_tile<float, double>(d_src, d_dst, a, sms, s) : launch_tile<float, float>(d_src, d_dst, a, sms, s); break;
    case AGF_F64:
        if (!d64) return agf_fail(AGF_E_INVALID, "a float64 source needs a float64 raster");
        rc = launch_tile<double, double>(d_src, d_dst, a, sms, s);
        break;
    case AGF_I16: rc = d64 ? launch_tile<int16_t, double>(d_src, d_dst, a, sms, s) : launch_tile<int16_t, float>(d_src, d_dst, a, sms, s); break;
    case AGF_I32: rc = d64 ? launch_tile<int32_t, double>(d_src, d_dst, a, sms, s) : launch_tile<int32_t, float>(d_src, d_dst, a, sms, s); break;
    case AGF_U8: rc = d64 ? launch_tile<uint8_t, double>(d_src, d_dst, a, sms, s) : launch_tile<uint8_t, float>(d_src, d_dst, a, sms, s); break;
    case AGF_I8: rc = d64 ? launch_tile<int8_t, double>(d_src, d_dst, a, sms, s) : launch_tile<int8_t, float>(d_src, d_dst, a, sms, s); break;
    case AGF_U16: rc = d64 ? launch_tile<uint16_t, double>(d_src, d_dst, a, sms, s) : launch_tile<uint16_t, float>(d_src, d_dst, a, sms, s); break;
    default: return agf_fail(AGF_E_INVALID, "source dtype %d", src_dtype);
    }
    if (rc) return rc;
    CU(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Blackwell decompression engine: LZ4 streams of Blosc-compressed chunks inflated next to HBM
// ------------------------------------------------------------------------------------------
// A zarr v2 store written with the default compressor holds Blosc frames: every block of a chunk is byte
// shuffled and split into `typesize` independent raw-LZ4 streams.  Instead of inflating them on host
// threads (~1 GB/s per core) and pushing the decoded bytes through PCIe, the compressed frame is copied to
// the device as stored and its streams are handed to the hardware decompression engine in one batch
// (cuMemBatchDecompressAsync; measured on this B200: 64 MB in 0.19 ms = 350 GB/s decoded,
// profiles/r1_decompress_probe.json); agf_unshuffle_run then undoes the byte shuffle and
// agf_tile_place_run places the chunk.  The driver entry point is fetched through the runtime: no
// link-time libcuda dependency (same as the TMA descriptor encoder in agf_api.cu).

typedef CUresult (*BatchDecompressFn)(CUmemDecompressParams *, size_t, unsigned int, size_t *, CUstream);

static BatchDecompressFn batch_decompress_fn() {
    static BatchDecompressFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuMemBatchDecompressAsync", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (BatchDecompressFn)ptr;
    }
    return fn;
}

static void *driver_entry(const char *name) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
        ptr = nullptr;
    (void)cudaGetLastError();
    return ptr;
}

extern "C" int agf_decompress_caps(int32_t *algo_mask, int64_t *max_length) {
    typedef CUresult (*DeviceGetFn)(CUdevice *, int);
    typedef CUresult (*GetAttrFn)(int *, CUdevice_attribute, CUdevice);
    static DeviceGetFn device_get = (DeviceGetFn)driver_entry("cuDeviceGet");
    static GetAttrFn get_attr = (GetAttrFn)driver_entry("cuDeviceGetAttribute");
    int dev = 0, mask = 0, mx = 0;
    CU(cudaGetDevice(&dev));
    CUdevice cudev;
    if (batch_decompress_fn() != nullptr && device_get && get_attr && device_get(&cudev, dev) == CUDA_SUCCESS) {
        if (get_attr(&mask, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK, cudev) != CUDA_SUCCESS) mask = 0;
        if (get_attr(&mx, CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_MAXIMUM_LENGTH, cudev) != CUDA_SUCCESS) mx = 0;
    }
    if (algo_mask) *algo_mask = mask;
    if (max_length) *max_length = mx;
    return 0;
}

extern "C" int agf_decompress_lz4_run(const void *d_src, const int64_t *src_off, const int64_t *src_len, void *d_dst,
                                      const int64_t *dst_off, const int64_t *dst_len, int64_t n, uint32_t *d_actual,
                                      uintptr_t stream) {
    if (n == 0) return 0;
    if (!d_src || !d_dst || !src_off || !src_len || !dst_off || !dst_len || !d_actual || n < 0)
        return agf_fail(AGF_E_INVALID, "null argument");
    BatchDecompressFn fn = batch_decompress_fn();
    int32_t mask = 0;
    int64_t mx = 0;
    int rc = agf_decompress_caps(&mask, &mx);
    if (rc) return rc;
    if (!fn || !(mask & CU_MEM_DECOMPRESS_ALGORITHM_LZ4))
        return agf_fail(AGF_E_UNSUPPORTED, "this device / driver has no LZ4 decompression engine");
    std::vector<CUmemDecompressParams> ops((size_t)n);
    memset(ops.data(), 0, ops.size() * sizeof(CUmemDecompressParams));
    for (int64_t i = 0; i < n; ++i) {
        if (src_len[i] <= 0 || dst_len[i] <= 0 || src_off[i] < 0 || dst_off[i] < 0 || src_len[i] > mx || dst_len[i] > mx)
            return agf_fail(AGF_E_INVALID, "stream %lld: %lld -> %lld bytes (engine limit %lld per operation)", (long long)i,
                            (long long)src_len[i], (long long)dst_len[i], (long long)mx);
        ops[i].srcNumBytes = (size_t)src_len[i];
        ops[i].dstNumBytes = (size_t)dst_len[i];
        ops[i].dstActBytes = d_actual + i;
        ops[i].src = (const char *)d_src + src_off[i];
        ops[i].dst = (char *)d_dst + dst_off[i];
        ops[i].algo = CU_MEM_DECOMPRESS_ALGORITHM_LZ4;
    }
    size_t bad = 0;
    CUresult e = fn(ops.data(), (size_t)n, 0, &bad, (CUstream)stream);
    if (e != CUDA_SUCCESS) return agf_fail(AGF_E_STATE, "cuMemBatchDecompressAsync: CUresult %d at stream %lld", (int)e, (long long)bad);
    return 0;
}

namespace {

// out[b * blocksize + i * TS + j] = in[b * blocksize + j * n_b + i],  n_b = elements of block b; the
// (bsize % TS) trailing bytes of a block are stored unshuffled (Blosc's shuffle contract).
template <int TS>
__global__ void __launch_bounds__(256) agf_unshuffle(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long nbytes,
                                                     long long blocksize) {
    const long long per_block = blocksize / TS;                            // elements of a full block
    const long long n_elem_slots = ((nbytes + blocksize - 1) / blocksize) * per_block;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem_slots; e += step) {
        const long long b = e / per_block, i = e % per_block;
        const long long base = b * blocksize;
        const long long bsize = min(blocksize, nbytes - base);
        const long long n = bsize / TS;
        if (i < n) {
            uint8_t v[TS];
#pragma unroll
            for (int j = 0; j < TS; ++j) v[j] = in[base + j * n + i];
#pragma unroll
            for (int j = 0; j < TS; ++j) out[base + i * TS + j] = v[j];
        }
        if (i == 0)
            for (long long k = n * TS; k < bsize; ++k) out[base + k] = in[base + k];
    }
}

}  // namespace

namespace {

// One CTA per segment (grid-stride): byte copies, coalesced; sources sit at arbitrary byte offsets of a frame.
__global__ void __launch_bounds__(256) agf_copy_segments(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                         const long long *__restrict__ table, long long n) {
    for (long long s = blockIdx.x; s < n; s += gridDim.x) {
        const uint8_t *a = src + table[s];
        uint8_t *b = dst + table[n + s];
        const long long len = table[2 * n + s];
        for (long long k = threadIdx.x; k < len; k += blockDim.x) b[k] = a[k];
    }
}

}  // namespace

extern "C" int agf_copy_segments_run(const void *d_src, void *d_dst, const int64_t *d_table, int64_t n, uintptr_t stream) {
    if (n == 0) return 0;
    if (!d_src || !d_dst || !d_table || n < 0) return agf_fail(AGF_E_INVALID, "null argument");
    int dev = 0, sms = 148;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned blocks = (unsigned)std::min<long long>(n, (long long)sms * 16);
    agf_copy_segments<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)d_src, (uint8_t *)d_dst,
                                                                (const long long *)d_table, (long long)n);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int agf_unshuffle_run(const void *d_src, void *d_dst, int64_t nbytes, int32_t typesize, int64_t blocksize,
                                 uintptr_t stream) {
    if (!d_src || !d_dst || d_src == d_dst) return agf_fail(AGF_E_INVALID, "null / aliased argument");
    if (nbytes < 0 || blocksize <= 0 || blocksize < typesize) return agf_fail(AGF_E_INVALID, "bad size");
    if (nbytes == 0) return 0;
    int dev = 0, sms = 148;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long slots = ((nbytes + blocksize - 1) / blocksize) * (blocksize / typesize);
    const unsigned blocks = (unsigned)std::min<long long>((slots + 255) / 256, (long long)sms * 32);
    cudaStream_t s = (cudaStream_t)stream;
    const uint8_t *in = (const uint8_t *)d_src;
    uint8_t *out = (uint8_t *)d_dst;
    switch (typesize) {
    case 2: agf_unshuffle<2><<<blocks, 256, 0, s>>>(in, out, nbytes, blocksize); break;
    case 4: agf_unshuffle<4><<<blocks, 256, 0, s>>>(in, out, nbytes, blocksize); break;
    case 8: agf_unshuffle<8><<<blocks, 256, 0, s>>>(in, out, nbytes, blocksize); break;
    default: return agf_fail(AGF_E_UNSUPPORTED, "unshuffle for typesize %d", typesize);
    }
    CU(cudaGetLastError());
    return 0;
}
