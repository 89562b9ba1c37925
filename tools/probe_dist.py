"""dist between cells at C4 size: plain vs pipelined kernel, time and equality (development aid)."""
import sys, time
import torch
sys.path.insert(0, ".")
from honeybadger_b200 import _lib, hmm
lib = _lib.load()
for G, N, miss in ((20000, 10000, 0.0), (20000, 5000, 0.0), (20000, 2000, 0.0), (50000, 5000, 0.3), (20003, 3001, 0.02)):
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    s = hmm.alloc_rows(G, N, torch.float64, "cuda")
    s.copy_(torch.randn((G, N), dtype=torch.float64, device="cuda", generator=g))
    if miss:
        s[:, :N][torch.rand((G, N), device="cuda", generator=g) < miss] = float("nan")
    out = {}
    for mode in (1, 0):
        hmm.set_dist_mode(mode)
        D = torch.empty((N, N), dtype=torch.float64, device="cuda")
        best = 1e9
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _lib.check(lib.hb_dist_sq_dev(s.data_ptr(), s.stride(0), G, N, D.data_ptr(), None))
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        out[mode] = (best, D)
    same = bool(torch.equal(out[0][1].view(torch.int64), out[1][1].view(torch.int64)))
    ops = 3.0 * G * N * (N + 64) / 2
    print(f"G={G} N={N} missing={miss}: plain {out[1][0]*1e3:.1f} ms  pipelined {out[0][0]*1e3:.1f} ms  ({ops/out[0][0]/1e12:.2f} T fp64 op/s)  identical={same}", flush=True)
hmm.set_dist_mode(0)
