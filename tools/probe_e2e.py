"""Host-buffer entry point: pinned vs ordinary (pageable) host memory (development aid)."""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import _lib, hmm, synth
lib = _lib.load()
N, G = 5000, 20000
seg = synth.segments(G)
rng = np.random.default_rng(0)
x_page = rng.standard_normal((N, G))            # C-order [cells, positions] == R column-major [G x N]
sd = np.full(N, 1.0)
Pi, mean, delta = hmm.sym_Pi(3, 1e-6), np.array([-1.0, 0, 1.0]), np.array([0.0, 1, 0])
vp = lambda a: a.ctypes.data_as(C.c_void_p)
def run(xbuf, ybuf, gbuf, label):
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        rc = lib.hb_hmm_norm_host(xbuf, G, N, vp(seg), seg.size - 1, vp(mean), vp(sd), 0, vp(Pi), vp(delta), 3, ybuf, gbuf, None)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        _lib.check(rc)
    print(f"{label}: {dt*1e3:.0f} ms per {N}x{G} ({N*G/dt/1e9:.2f} G steps/s, {(N*G*36)/dt/1e9:.1f} GB/s over PCIe)", flush=True)
y_page = np.empty((N, G), np.int32); g_page = np.empty((3, N, G))
run(vp(x_page), vp(y_page), vp(g_page), "pageable")
xp = torch.empty((N, G), dtype=torch.float64).pin_memory(); xp.copy_(torch.from_numpy(x_page))
yp = torch.empty((N, G), dtype=torch.int32).pin_memory(); gp = torch.empty((3, N, G), dtype=torch.float64).pin_memory()
run(C.c_void_p(xp.data_ptr()), C.c_void_p(yp.data_ptr()), C.c_void_p(gp.data_ptr()), "pinned  ")
assert (yp.numpy() == y_page).all() and (gp.numpy() == g_page).all()
