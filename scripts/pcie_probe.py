"""PCIe ceilings for the e2e path: pinned H2D, D2H, and both directions at once (chunked like the host
pipeline)."""
import torch, time
dev = torch.device("cuda:0")
n = 737_280_000 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device=dev)
d_b = torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, k=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both(): h2d(); d2h()
gb = n * 4 / 1e9
print(f"H2D alone   {gb / timeit(h2d):.1f} GB/s")
print(f"D2H alone   {gb / timeit(d2h):.1f} GB/s")
t = timeit(both)
print(f"both at once {gb / t:.1f} GB/s per direction ({t*1e3:.2f} ms for 737 MB each way)")
for chunk_mb in (8, 32, 128):
    c = chunk_mb * (1 << 20) // 4
    def chunked():
        for o in range(0, n, c):
            with torch.cuda.stream(s1): d_a[o:o + c].copy_(h_in[o:o + c], non_blocking=True)
            with torch.cuda.stream(s2): h_out[o:o + c].copy_(d_b[o:o + c], non_blocking=True)
    t = timeit(chunked)
    print(f"both, {chunk_mb} MiB chunks {gb / t:.1f} GB/s per direction ({t*1e3:.2f} ms)")
