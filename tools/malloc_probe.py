"""cudaMalloc / cudaFree cost on this box: fresh memory vs memory the process released before, and a stream-ordered pool
(cudaMallocAsync with an unlimited release threshold).  usage: malloc_probe.py [GB]"""
import ctypes as C
import sys
import time

import torch  # noqa: F401  (loads libcudart)

rt = None
for n in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = C.CDLL(n)
        break
    except OSError:
        pass
GB = int(sys.argv[1]) if len(sys.argv) > 1 else 48
rt.cudaFree(None)


def alloc_all(sizes, label):
    ps = []
    t0 = time.perf_counter()
    for s in sizes:
        p = C.c_void_p()
        rc = rt.cudaMalloc(C.byref(p), C.c_size_t(s))
        assert rc == 0, rc
        ps.append(p)
    t1 = time.perf_counter()
    print(f"{label}: cudaMalloc of {len(sizes)} buffers / {sum(sizes)/2**30:.1f} GiB: {1e3*(t1-t0):.1f} ms", flush=True)
    return ps


def free_all(ps, label):
    t0 = time.perf_counter()
    for p in ps:
        rt.cudaFree(p)
    print(f"{label}: cudaFree of {len(ps)} buffers: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)


big = [4 << 30] * (GB // 4)
small = [64 << 20] * 64 + [1 << 20] * 256 + [4096] * 256
for rep in range(3):
    a = alloc_all(big, f"rep {rep} big")
    b = alloc_all(small, f"rep {rep} small")
    rt.cudaMemset(a[0], 0, C.c_size_t(4 << 30))
    rt.cudaDeviceSynchronize()
    free_all(b, f"rep {rep} small")
    free_all(a, f"rep {rep} big")

# stream-ordered pool
pool = C.c_void_p()
assert rt.cudaDeviceGetDefaultMemPool(C.byref(pool), 0) == 0
thr = C.c_uint64(2**64 - 1)
assert rt.cudaMemPoolSetAttribute(pool, 4, C.byref(thr)) == 0     # cudaMemPoolAttrReleaseThreshold = 4
for rep in range(3):
    ps = []
    t0 = time.perf_counter()
    for s in big + small:
        p = C.c_void_p()
        assert rt.cudaMallocAsync(C.byref(p), C.c_size_t(s), None) == 0
        ps.append(p)
    rt.cudaDeviceSynchronize()
    t1 = time.perf_counter()
    for p in ps:
        rt.cudaFreeAsync(p, None)
    rt.cudaDeviceSynchronize()
    t2 = time.perf_counter()
    print(f"pool rep {rep}: cudaMallocAsync {1e3*(t1-t0):.1f} ms, cudaFreeAsync {1e3*(t2-t1):.1f} ms", flush=True)
