"""binom2 V+FB timing for library variants: python tools/probe_binom.py lib1.so ..."""
import os, subprocess, sys
if len(sys.argv) > 2 or (len(sys.argv) == 2 and not os.environ.get("HB_CHILD")):
    for lib in sys.argv[1:]:
        env = dict(os.environ, HB_B200_LIB=os.path.abspath(lib), HB_CHILD="1")
        subprocess.run([sys.executable, __file__, lib], env=env)
    sys.exit(0)
import numpy as np, torch
sys.path.insert(0, ".")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from honeybadger_b200 import _lib, hmm, synth
_lib.check(_lib.load().hb_set_binom2_mode(int(os.environ.get("HB_MODE", "0"))))
N, G = int(os.environ.get("HB_N", 10000)), int(os.environ.get("HB_G", 50000))
d = bench.device_inputs("allele", N, G, 3, torch.device("cuda:0"))
r, n, seg = d["r"], d["n"], d["seg"]
Pi2 = hmm.sym_Pi(2, 1e-6)
TILED = int(os.environ.get("HB_TILED", "0"))
if TILED:
    rt, nt = hmm.tile_rows(r), hmm.tile_rows(n)
    run = lambda out=None, **kw: hmm.hmm_binom_tiled_dev(rt, nt, N, [0.1, 0.45], Pi2, [0., 1.], seg_off=seg, out=out, **kw)
else:
    run = lambda out=None, **kw: hmm.hmm_binom_dev(r, n, [0.1, 0.45], Pi2, [0., 1.], seg_off=seg, out=out, **kw)
out = run()
res = []
for mode, kw in (("V+FB", dict(viterbi=True, gamma=True)), ("FB", dict(viterbi=False, gamma=True)), ("V", dict(viterbi=True, gamma=False))):
    fn = lambda: run(out=out, **kw)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[2]
    res.append(f"{mode} {ms:.3f} ms {N*G/ms/1e6:.1f} G/s")
hmm.status_dev()
print(os.path.basename(sys.argv[1]), "tiled" if TILED else "rows", "mode", os.environ.get("HB_MODE", "0"), "general rerun", hmm.binom2_general_rerun(), f"n>0 frac {(n>0).float().mean().item():.3f} max n {int(n.max())}", " | ".join(res), flush=True)
