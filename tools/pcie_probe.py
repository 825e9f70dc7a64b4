"""Pinned host <-> device copy bandwidth on the GPU box (diagnostics for the e2e leg of bench.py)."""
import time, torch
dev = torch.device("cuda:0")
nbytes = 184 * 1024 * 1024
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev); d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
for nm, fn, mult in (("h2d", h2d, 1), ("d2h", d2h, 1), ("both directions", both, 2)):
    t = timeit(fn)
    print(f"{nm}: {t*1e3:.2f} ms for {mult} x 184 MiB -> {mult * nbytes / t / 1e9:.1f} GB/s total")
for parts in (4, 16):
    c = nbytes // parts
    def chunked():
        for p in range(parts):
            with torch.cuda.stream(s1): d_a[p*c:(p+1)*c].copy_(h_in[p*c:(p+1)*c], non_blocking=True)
            with torch.cuda.stream(s2): h_out[p*c:(p+1)*c].copy_(d_b[p*c:(p+1)*c], non_blocking=True)
    t = timeit(chunked)
    print(f"both directions in {parts} chunks: {t*1e3:.2f} ms -> {2 * nbytes / t / 1e9:.1f} GB/s total")
