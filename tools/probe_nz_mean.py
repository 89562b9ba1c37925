"""mean(x[x != 0]) replayed in R's serial order (hb_mean_nonzero_dev, exact = 1) at growing sizes (development aid)."""
import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import _lib, hmm
from oracle import oracle as O
O.build()
lib = _lib.load()
rng = np.random.default_rng(3)
for G, N in ((311, 77), (5000, 500), (20000, 2000)):
    m = rng.gamma(2.0, 1.5, size=(G, N)) * (rng.random((G, N)) < 0.4)
    d = hmm.to_rows(torch.from_numpy(m).cuda())
    mean, margin = C.c_double(0), C.c_double(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.check(lib.hb_mean_nonzero_dev(d.data_ptr(), d.stride(0), G, N, 1, C.byref(mean), C.byref(margin), None))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    want = O.mean_nonzero(m) if G * N <= 3_000_000 else None
    print(f"G={G} N={N}: {dt*1e3:.1f} ms ({dt / (G * N) * 1e9:.2f} ns per element, 2 passes)  equals R: {None if want is None else mean.value == want}", flush=True)
