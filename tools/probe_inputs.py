"""Input construction at C4 size: setGexpMats (10k cells x 20k genes) and setAlleleMats (10k x 200k SNPs),
device-resident forms, CUDA-event timing (development aid)."""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import _lib
from honeybadger_b200.honeybadger import set_gexp_mats_dev, set_allele_mats_dev
lib = _lib.load()
N, G, GS = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, 20000, 200000
g = torch.Generator(device="cuda").manual_seed(1)
sc = torch.empty((G, N), dtype=torch.float64, device="cuda")
sc.copy_(torch.rand((G, N), generator=g, device="cuda") * 8)
sc *= (torch.rand((G, N), generator=g, device="cuda") > 0.4)
ref = sc[:, :3].mean(1, keepdim=True).contiguous() + 0.1
def timed(label, f, reps=3):
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); out = f(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"{label}: {dt*1e3:.1f} ms", flush=True)
    return out
timed("setGexpMats filter=F scale=F", lambda: set_gexp_mats_dev(sc, ref, False, 0, None, None, False))
timed("setGexpMats filter=T scale=F", lambda: set_gexp_mats_dev(sc, ref, True, 0, None, None, False))
o = timed("setGexpMats filter=T scale=T", lambda: set_gexp_mats_dev(sc, ref, True, 0, None, None, True))
print("kept", o[1].size, "info", o[2])
out = torch.empty(G, dtype=torch.float64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
timed("  rowMeans kernel", lambda: lib.hb_row_means_dev(sc.data_ptr(), N, G, N, out.data_ptr(), st))
cen = torch.empty(N, dtype=torch.float64, device="cuda"); scl = torch.empty(N, dtype=torch.float64, device="cuda")
timed("  column stats kernel", lambda: lib.hb_col_scale_stats_dev(sc.data_ptr(), N, G, N, None, cen.data_ptr(), scl.data_ptr(), st))
m, mg = C.c_double(0), C.c_double(0)
timed("  mean nonzero (fast)", lambda: lib.hb_mean_nonzero_dev(sc.data_ptr(), N, G, N, 0, C.byref(m), C.byref(mg), st))
print("   mean", m.value, "margin", mg.value)
if len(sys.argv) > 2:
    timed("  mean nonzero (exact replay)", lambda: lib.hb_mean_nonzero_dev(sc.data_ptr(), N, G, N, 1, C.byref(m), None, st), reps=1)
    print("   mean", m.value)
del sc, o, out
torch.cuda.empty_cache()
n = (torch.rand((GS, N), generator=g, device="cuda") < 0.108).to(torch.int32) * torch.randint(1, 12, (GS, N), generator=g, device="cuda", dtype=torch.int32)
r = (n.to(torch.float32) * torch.rand((GS, N), generator=g, device="cuda")).round().to(torch.int32)
o = timed("setAlleleMats filter=T", lambda: set_allele_mats_dev(r, n, het_deviance_threshold=0.05))
print("kept", o["keep"].size, "of", GS, "n_het", o["n_het"])
