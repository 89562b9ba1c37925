"""Ward linkage: single block vs thread-block cluster (cluster size forced through HB_WARD_R), time and
equality (development aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import _lib, hmm
lib = _lib.load()
for N in (256, 1024, 2048, 3000, 5000, 10000):
    g = torch.Generator(device="cuda"); g.manual_seed(N)
    x = torch.randn((N, 64), dtype=torch.float64, device="cuda", generator=g)
    x[: N // 3] += 1.0
    D0 = torch.cdist(x, x) ** 2
    D0 = (D0 + D0.T) / 2
    D0.fill_diagonal_(0.0)
    res = {}
    for tag, mode, r in (("single", 1, None), ("r1", 2, 1), ("r2", 2, 2), ("r4", 2, 4), ("r8", 2, 8), ("r16", 2, 16), ("auto", 0, None)):
        if r is None:
            os.environ.pop("HB_WARD_R", None)
        else:
            os.environ["HB_WARD_R"] = str(r)
        hmm.set_ward_mode(mode)
        best = 1e9
        for it in range(2):
            D = D0.clone()
            ma = torch.empty(N - 1, dtype=torch.int32, device="cuda"); mb = torch.empty_like(ma)
            hh = torch.empty(N - 1, dtype=torch.float64, device="cuda")
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _lib.check(lib.hb_ward_linkage_dev(D.data_ptr(), N, ma.data_ptr(), mb.data_ptr(), hh.data_ptr(), None))
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        res[tag] = (best, ma.cpu().numpy(), mb.cpu().numpy(), hh.cpu().numpy())
    same = all(np.array_equal(res["single"][i], res[t][i]) for i in (1, 2, 3) for t in res)
    print(f"N={N:6d}  " + "  ".join(f"{t} {res[t][0]*1e3:7.2f}" for t in res) + f"  ms  identical={same}", flush=True)
os.environ.pop("HB_WARD_R", None)
hmm.set_ward_mode(0)
