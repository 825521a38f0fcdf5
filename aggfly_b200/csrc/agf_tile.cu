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
#include <cstdint>

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
    if (a.sx == 1 || a.nx == 1 || (!f_is_t && !f_is_y)) {
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
                                  double offset, int32_t has_fill, double fill, uintptr_t stream) {
    if (!d_src || !d_dst) return agf_fail(AGF_E_INVALID, "null argument");
    if (nt < 0 || ny < 0 || nx < 0 || st < 0 || sy < 0 || sx < 0) return agf_fail(AGF_E_INVALID, "negative extent / stride");
    if (t0 < 0 || y0 < 0 || x0 < 0 || n_lon <= 0 || x0 + nx > n_lon || (y0 + ny) * n_lon > ld)
        return agf_fail(AGF_E_INVALID, "tile [%lld+%lld, %lld+%lld] does not fit a row of %lld x %lld cells (pitch %lld)",
                        (long long)y0, (long long)ny, (long long)x0, (long long)nx, (long long)(ld / n_lon),
                        (long long)n_lon, (long long)ld);
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
