"""How fast can this host move 160 MB each way at the same time?  Raw pinned-memory copies on two streams (no kernels):
the ceiling of the host-buffer entry point's e2e figure.   python profiles/pcie_duplex_probe.py"""
import time, torch
n = 160 * 1024 * 1024 // 8
h_in, h_out = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.float64, device="cuda"), torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, chunks=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = n // chunks
    for k in range(chunks):
        a, b = k * c, (k + 1) * c if k < chunks - 1 else n
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1): d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out[a:b].copy_(d_out[a:b], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for mode in ("h2d", "d2h", "both"):
    for chunks in (1, 8):
        run(mode, chunks)
        ts = [run(mode, chunks) for _ in range(5)]
        print(f"{mode:5s} chunks {chunks}: best {min(ts):.3f} ms  ({160 * (2 if mode == 'both' else 1) / min(ts):.1f} GB/s aggregate)")
