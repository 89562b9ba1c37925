"""Timing of one pooled (reference-shaped) recursion round at BASELINE C4 sizes (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import hmm, synth
from honeybadger_b200 import honeybadger as hb

def member_from_clones(N, heights, rng):
    clone = rng.choice(3, size=N, p=[0.5, 0.3, 0.2])
    sub = rng.integers(0, 2, size=N)
    labs = []
    for h in heights:
        if h == 1: lab = np.ones(N, int)
        elif h == 2: lab = np.where(clone == 0, 1, 2)
        elif h == 3: lab = clone + 1
        elif h == 4: lab = np.where(clone == 0, 1 + sub, clone + 2)
        else: lab = np.where(clone == 0, 1 + sub, np.where(clone == 1, 3 + sub, 5))
        labs.append(lab)
    return hb._membership(labs, N)[0]

rng = np.random.default_rng(0)
dev = torch.device("cuda:0")
N, G = 10000, 20000
x = hmm.alloc_rows(G, N, torch.float64, dev)
x.copy_(torch.randn((G, N), dtype=torch.float64, device=dev) * 2.5)
mem = member_from_clones(N, [1, 2, 3, 4, 5], rng)
print("expr members", mem.shape, flush=True)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    y, px, psd = hb.gexp_round_dev(x, mem, [-1.0157, 0, 1.0157], 1e-6)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"gexp pooled round  K={mem.shape[0]} G={G} N={N}: {dt*1e3:.1f} ms  ({8*N*G/dt/1e9:.1f} GB/s of the 8NG bytes)", flush=True)
del x
G2 = 200000
n = hmm.alloc_rows(G2, N, torch.int32, dev); r = hmm.alloc_rows(G2, N, torch.int32, dev)
n.copy_(((torch.rand((G2, N), device=dev) < 0.108).int() * torch.randint(1, 6, (G2, N), device=dev, dtype=torch.int32)))
r.copy_((torch.rand((G2, N), device=dev) * (n + 1).float()).floor().int().clamp(max=n))
mem = member_from_clones(N, [1, 2, 3], rng)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    y, mafl, sizel = hb.allele_round_dev(r, n, mem, 0.1, 0.45, 1e-6)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"allele pooled round K={mem.shape[0]} G={G2} N={N}: {dt*1e3:.1f} ms  ({8*N*G2/dt/1e9:.1f} GB/s of the 2x4NG bytes)", flush=True)
