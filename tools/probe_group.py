"""Timing of the device grouping (runmean + dist + Ward.D2) at C4 size (development aid)."""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import _lib, hmm
from honeybadger_b200 import honeybadger as hb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
clone = torch.multinomial(torch.tensor([0.5, 0.3, 0.2], device=dev), N, True, generator=g)
x = hmm.alloc_rows(G, N, torch.float64, dev)
x.copy_(torch.randn((G, N), dtype=torch.float64, device=dev, generator=g) * 2.5)
x[2000:4000, clone == 0] += 1.0
x[9000:11000, clone == 1] -= 1.0
lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
t0 = sync()
S = hmm.alloc_rows(G, N, torch.float64, dev)
_lib.check(lib.hb_runmean_dev(x.data_ptr(), x.stride(0), G, N, 101, S.data_ptr(), S.stride(0), st))
t1 = sync()
D = torch.empty((N, N), dtype=torch.float64, device=dev)
_lib.check(lib.hb_dist_sq_dev(S.data_ptr(), S.stride(0), G, N, D.data_ptr(), st))
t2 = sync()
ma = torch.empty(N - 1, dtype=torch.int32, device=dev); mb = torch.empty_like(ma); hh = torch.empty(N - 1, dtype=torch.float64, device=dev)
_lib.check(lib.hb_ward_linkage_dev(D.data_ptr(), N, ma.data_ptr(), mb.data_ptr(), hh.data_ptr(), st))
t3 = sync()
labs = hb.cut_merges(ma.cpu().numpy(), mb.cpu().numpy(), hh.cpu().numpy(), N, [1, 2, 3, 4, 5])
t4 = time.perf_counter()
print(f"N={N} G={G}: runmean {1e3*(t1-t0):.1f} ms | dist {1e3*(t2-t1):.1f} ms ({N*N/2*G*3/(t2-t1)/1e12:.2f} Tflop/s) | ward {1e3*(t3-t2):.1f} ms | cutree(host) {1e3*(t4-t3):.1f} ms")
l3 = labs[2]; c = clone.cpu().numpy()
print("k=3 purity:", sum(np.bincount(c[l3 == v]).max() for v in np.unique(l3)) / N)
