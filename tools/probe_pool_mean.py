"""Timing of the pooled expression mean (R `mean`, 80-bit replay) at C4 size (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import _lib, hmm
lib = _lib.load()
G, N = 20000, 10000
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = hmm.alloc_rows(G, N, torch.float64, "cuda")
x.copy_(torch.randn((G, N), dtype=torch.float64, device="cuda", generator=g) * 2.6)
rng = np.random.default_rng(0)
clone = rng.choice(3, size=N, p=[0.5, 0.3, 0.2]); idx = np.arange(N)
labs = [np.zeros(N, int), (clone == 0) * 1, clone, clone * 2 + (idx % 2) * (clone == 0), clone * 3 + (idx % 3) * (clone == 0)]
rows = [l == v for l in labs for v in np.unique(l)]
member = torch.from_numpy(np.stack(rows).astype(np.uint8)).cuda()
K = member.shape[0]
out = torch.empty((K, G), dtype=torch.float64, device="cuda")
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.check(lib.hb_pool_mean_dev(x.data_ptr(), x.stride(0), G, N, member.data_ptr(), K, out.data_ptr(), None))
    torch.cuda.synchronize(); print(f"pool_mean K={K}: {(time.perf_counter() - t0) * 1e3:.2f} ms")
# exactness spot check against numpy longdouble (x87 on the host) for a few (group, gene) pairs
xh = x[:64].cpu().numpy(); mh = member.cpu().numpy().astype(bool)
bad = 0
for k in range(K):
    for gi in range(0, 64, 7):
        v = xh[gi, mh[k]].astype(np.longdouble)
        s = np.longdouble(0)
        for a in v: s += a
        s /= len(v)
        t = np.longdouble(0)
        for a in v: t += (a - s)
        ref = float(s + t / len(v))
        bad += ref != float(out[k, gi])
print("mismatches vs host long double replay:", bad)
