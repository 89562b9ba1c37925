import torch, time
dev = torch.device("cuda:0")
h_in = torch.empty(800_000_000 // 8, dtype=torch.float64).pin_memory()
h_out = torch.empty(2_500_000_000 // 8, dtype=torch.float64).pin_memory()
d_in = torch.empty_like(h_in, device=dev); d_out = torch.empty_like(h_out, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D 0.8 GB: {a:.1f} ms ({0.8/a*1e3:.1f} GB/s) | D2H 2.5 GB: {b:.1f} ms ({2.5/b*1e3:.1f} GB/s) | both: {c:.1f} ms (sum {a+b:.1f}, max {max(a,b):.1f})")
