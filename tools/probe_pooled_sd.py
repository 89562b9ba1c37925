"""sd of K pooled vectors (R's sd: three sequential 80-bit accumulations per vector), time and exactness (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import _lib
from oracle import oracle as O
O.build()
lib = _lib.load()
K, G = 15, 20000
rng = np.random.default_rng(2)
x = rng.normal(size=(K, G)) * 0.3 + rng.normal(size=(K, 1)) * 0.02
xd = torch.from_numpy(x).cuda()
sd = torch.empty(K, dtype=torch.float64, device="cuda")
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.check(lib.hb_pooled_sd_dev(xd.data_ptr(), G, K, sd.data_ptr(), None))
    torch.cuda.synchronize(); print(f"pooled_sd K={K} G={G}: {(time.perf_counter() - t0) * 1e3:.2f} ms")
got = sd.cpu().numpy()
print("mismatches vs R sd (oracle):", sum(got[k] != O.r_sd(x[k]) for k in range(K)))
# adversarial rows of one length: wide exponents, integers, tie-heavy, powers of two pulled down, zeros, cancelling
G2 = 6000
rows = [rng.normal(size=G2) * 10.0 ** rng.integers(-12, 12, G2), rng.integers(0, 50, G2).astype(float),
        np.ldexp(rng.integers(1, 1 << 20, G2).astype(float), rng.integers(-70, 70, G2)), rng.gamma(2.0, 1.5, G2),
        np.concatenate([np.zeros(1000), rng.normal(size=G2 - 1500), np.zeros(500)]), np.zeros(G2) + 0.1,
        np.concatenate([[float(1 << 23)], np.ldexp(rng.integers(1, 1 << 30, G2 - 1).astype(float), -40)])]
for k in (0, 5, 11, 30):
    rows.append(np.concatenate([[2.0 ** k], np.ldexp(rng.integers(-(1 << 14), 1 << 14, G2 - 1).astype(float),
                                                     k - 64 - rng.integers(0, 4, G2 - 1))]))
pw = []
for _ in range(G2 // 3):
    k = int(rng.integers(-20, 20))
    pw += [2.0 ** k, -2.0 ** (k - int(rng.integers(50, 70))), -2.0 ** k * (1 - 2.0 ** -int(rng.integers(1, 53)))]
rows.append(np.array(pw))
y = rng.normal(size=G2 // 2)
z = np.empty(G2); z[0::2] = y; z[1::2] = -y + np.ldexp(rng.integers(-8, 8, G2 // 2).astype(float), -70)
rows.append(z)
X = np.ascontiguousarray(np.stack(rows))
xd2 = torch.from_numpy(X).cuda()
sd2 = torch.empty(X.shape[0], dtype=torch.float64, device="cuda")
_lib.check(lib.hb_pooled_sd_dev(xd2.data_ptr(), G2, X.shape[0], sd2.data_ptr(), None))
g2 = sd2.cpu().numpy()
bad = [k for k in range(X.shape[0]) if not (g2[k] == O.r_sd(X[k]) or (g2[k] != g2[k] and O.r_sd(X[k]) != O.r_sd(X[k])))]
print("adversarial rows differing from R sd:", bad, "of", X.shape[0])
for G3 in (2, 3, 31, 32, 33, 65):
    X3 = np.ascontiguousarray(rng.normal(size=(4, G3)) * 2.6 + 7.5)
    x3 = torch.from_numpy(X3).cuda(); s3 = torch.empty(4, dtype=torch.float64, device="cuda")
    _lib.check(lib.hb_pooled_sd_dev(x3.data_ptr(), G3, 4, s3.data_ptr(), None))
    print("G =", G3, "mismatches:", sum(float(s3[k]) != O.r_sd(X3[k]) for k in range(4)))
