"""Probe of the B200 hardware decompression engine through the CUDA driver API (cuMemBatchDecompressAsync):
which algorithms the device reports, the maximum length of one operation, and a round trip of raw LZ4 /
Snappy / Deflate blocks compressed on the host.  Evidence for the next step of the zarr feed (SURVEY f-1):
compressed chunks could cross PCIe as stored and be inflated next to HBM instead of on host threads.
Prints one JSON object; never raises (each stage records its own error).

    python tools/probe_decompress.py [--out gpurun_out/decompress_probe.json]"""
import argparse
import json
import os
import time
import zlib

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--mb", type=int, default=64, help="decoded bytes per operation batch")
    ap.add_argument("--block-kb", type=int, default=256, help="decoded bytes per independent block")
    ap.add_argument("--cases", default="lz4", help="comma list of lz4, lz4_src_unaligned, lz4_dst_unaligned, snappy, deflate; "
                    "a failing operation poisons the CUDA context, so run doubtful cases in separate processes")
    a = ap.parse_args()
    out = {}
    try:
        import ctypes as C
        import torch
        import pyarrow as pa
        torch.cuda.init()
        torch.zeros(1, device="cuda")                                       # primary context current on this thread
        cu = C.CDLL("libcuda.so.1")
        dev = C.c_int()
        assert cu.cuCtxGetDevice(C.byref(dev)) == 0
        val = C.c_int()
        err = cu.cuDeviceGetAttribute(C.byref(val), 136, dev)               # CU_DEVICE_ATTRIBUTE_MEM_DECOMPRESS_ALGORITHM_MASK
        out["algorithm_mask"] = int(val.value) if err == 0 else f"CUresult {err}"
        err = cu.cuDeviceGetAttribute(C.byref(val), 137, dev)               # ..._MEM_DECOMPRESS_MAXIMUM_LENGTH
        out["maximum_length"] = int(val.value) if err == 0 else f"CUresult {err}"

        class Params(C.Structure):                                          # CUmemDecompressParams (cuda.h), 64 bytes
            _fields_ = [("srcNumBytes", C.c_size_t), ("dstNumBytes", C.c_size_t), ("dstActBytes", C.c_void_p),
                        ("src", C.c_void_p), ("dst", C.c_void_p), ("algo", C.c_int), ("padding", C.c_ubyte * 20)]
        assert C.sizeof(Params) == 64
        cu.cuMemBatchDecompressAsync.argtypes = [C.POINTER(Params), C.c_size_t, C.c_uint, C.POINTER(C.c_size_t), C.c_void_p]
        cu.cuMemBatchDecompressAsync.restype = C.c_int
        out["algorithms"] = [n for b, n in ((1, "deflate"), (2, "snappy"), (4, "lz4")) if isinstance(out["algorithm_mask"], int)
                             and out["algorithm_mask"] & b]
    except Exception as exc:                                               # pragma: no cover
        out["error"] = f"{type(exc).__name__}: {exc}"
        _emit(out, a.out)
        return

    rng = np.random.default_rng(0)
    block = a.block_kb << 10
    n_blocks = max(1, (a.mb << 20) // block)
    vals = (np.cumsum(rng.integers(-8, 9, n_blocks * block // 4)) / 64).astype(np.float32)       # smooth, compressible
    raw = vals.tobytes()
    # byte-shuffled like Blosc does before LZ4 (what zarr v2 stores by default)
    shuf = np.frombuffer(raw, np.uint8).reshape(n_blocks, block // 4, 4).transpose(0, 2, 1).copy().tobytes()

    def deflate_raw(b):
        c = zlib.compressobj(1, zlib.DEFLATED, -15)
        return c.compress(b) + c.flush()

    codecs = {"lz4": (4, lambda b: pa.Codec("lz4_raw").compress(b, asbytes=True)),
              "snappy": (2, lambda b: pa.Codec("snappy").compress(b, asbytes=True)),
              "deflate": (1, deflate_raw)}
    for case in a.cases.split(","):
        name = case.split("_")[0]
        algo, enc = codecs[name]
        rec = {}
        out[case] = rec
        src_shift = 1 if case.endswith("src_unaligned") else 0          # every stream starts at an odd address
        dst_shift = 1 if case.endswith("dst_unaligned") else 0
        if dst_shift:
            block -= 1                                                   # odd block size -> odd destination offsets
            shuf = shuf[: n_blocks * block]
        try:
            parts = [enc(shuf[i * block:(i + 1) * block]) for i in range(n_blocks)]
            offs = np.concatenate([[0], np.cumsum([(len(p) + 15) // 16 * 16 + 16 for p in parts])]).astype(np.int64) + src_shift
            comp = np.zeros(int(offs[-1]) + 16, np.uint8)
            for i, p in enumerate(parts):
                comp[offs[i]: offs[i] + len(p)] = np.frombuffer(p, np.uint8)
            rec.update(blocks=n_blocks, block_bytes=block, decoded_bytes=len(shuf), compressed_bytes=int(sum(len(p) for p in parts)))
            d_src = torch.from_numpy(comp).cuda()
            d_dst = torch.zeros(len(shuf), dtype=torch.uint8, device="cuda")
            d_act = torch.zeros(n_blocks, dtype=torch.int32, device="cuda")
            if not (isinstance(out["algorithm_mask"], int) and out["algorithm_mask"] & algo):
                rec["error"] = "not in the device's algorithm mask"
                continue
            params = (Params * n_blocks)()
            for i, p in enumerate(parts):
                q = params[i]
                q.srcNumBytes, q.dstNumBytes = len(p), block
                q.src, q.dst = d_src.data_ptr() + int(offs[i]), d_dst.data_ptr() + i * block
                q.dstActBytes = d_act.data_ptr() + 4 * i
                q.algo = algo
            stream = torch.cuda.current_stream().cuda_stream
            ms = []
            for rep in range(4):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                t0 = time.perf_counter()
                bad = C.c_size_t(0)
                err = cu.cuMemBatchDecompressAsync(params, n_blocks, 0, C.byref(bad), C.c_void_p(stream))
                submit_ms = (time.perf_counter() - t0) * 1e3
                e1.record()
                torch.cuda.synchronize()
                if err != 0:
                    rec["error"] = f"CUresult {err} (error index {bad.value})"
                    break
                ms.append(e0.elapsed_time(e1))
                rec["submit_ms"] = submit_ms
            if ms:
                rec["ms"] = ms
                rec["decoded_gbs"] = len(shuf) / 1e6 / min(ms[1:] or ms)
                rec["round_trip_ok"] = bool(d_dst.cpu().numpy().tobytes() == shuf)
                rec["act_bytes_ok"] = bool((d_act.cpu().numpy() == block).all())
        except Exception as exc:
            rec["error"] = f"{type(exc).__name__}: {exc}"
    _emit(out, a.out)


def _emit(out, path):
    line = json.dumps(out)
    print(line)
    if path:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
