"""PCIe duplex check of a GPU box: host->device and device->host alone, together on dedicated streams, and interleaved
chunk by chunk on two streams the way a double-buffered pipeline issues them (H2D, kernel, D2H per chunk and stream)."""
import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * reps / dt / 1e9
run(True, True, 1)
print("H2D alone %.1f GB/s" % run(True, False)); print("D2H alone %.1f GB/s" % run(False, True))
print("both, one stream per direction: %.1f GB/s each direction" % run(True, True))
C = 48 << 20
def pipeline(mode, chunks=64):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    evs = []
    for k in range(chunks):
        o = (k % 16) * C
        if mode == "mixed":            # stream k & 1: H2D, kernel, D2H
            with torch.cuda.stream((s1, s2)[k & 1]):
                d_in[o:o + C].copy_(h_in[o:o + C], non_blocking=True)
                d_out[o:o + C].add_(1)
                h_out[o:o + C].copy_(d_out[o:o + C], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return chunks * C / dt / 1e9
print("two streams, H2D + kernel + D2H per chunk on the same stream: %.1f GB/s each direction" % pipeline("mixed"))

# write-combined pinned host memory (cudaHostAllocWriteCombined): no snooping on the device's reads
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
ptr = ctypes.c_void_p()
m = 1 << 30
for flags, name in ((0, "default pinned"), (4, "write-combined pinned")):
    assert rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(m), ctypes.c_uint(flags)) == 0
    ctypes.memset(ptr, 1, m)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        assert rt.cudaMemcpyAsync(ctypes.c_void_p(d_in.data_ptr()), ptr, ctypes.c_size_t(m), 1, ctypes.c_void_p(0)) == 0
    assert rt.cudaDeviceSynchronize() == 0
    print("H2D from %s (cudaHostAlloc): %.1f GB/s" % (name, m * 5 / (time.perf_counter() - t0) / 1e9))
    rt.cudaFreeHost(ptr)
