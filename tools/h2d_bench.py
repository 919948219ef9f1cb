"""tools/h2d_bench.py -- host->device copy rate of 128 MiB from default pinned memory and from write-combined pinned memory
(cudaHostAllocWriteCombined), alone and with a device->host copy of 64 MiB running the other way (the e2e leg of bench.py)."""
import ctypes as C

import torch

rt = C.CDLL("libcudart.so.12")
torch.cuda.init()
dev = torch.device("cuda", 0)
n = 128 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(n // 2, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def alloc(flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 1, n)
    return p


down = alloc(0)
for name, flags in (("default pinned", 0), ("write-combined", 4)):
    h = alloc(flags)
    for duplex in (False, True):
        best = 1e9
        for rep in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(s_in)
            for _ in range(4):
                assert rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), h, C.c_size_t(n), 1, C.c_void_p(s_in.cuda_stream)) == 0
                if duplex:
                    assert rt.cudaMemcpyAsync(down, C.c_void_p(d2.data_ptr()), C.c_size_t(n // 2), 2, C.c_void_p(s_out.cuda_stream)) == 0
            e1.record(s_in)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 4)
        print(f"{name:16s} {'with d2h the other way' if duplex else 'alone':24s} {n / best / 1e6:6.1f} GB/s  ({best:.3f} ms per 128 MiB)", flush=True)
