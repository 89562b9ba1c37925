"""norm3 V+FB timing for a list of library variants: python tools/probe_norm3.py lib1.so lib2.so ..."""
import os, subprocess, sys
if len(sys.argv) > 2 or (len(sys.argv) == 2 and not os.environ.get("HB_CHILD")):
    for lib in sys.argv[1:]:
        env = dict(os.environ, HB_B200_LIB=os.path.abspath(lib), HB_CHILD="1")
        subprocess.run([sys.executable, __file__, lib], env=env)
    sys.exit(0)
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import _lib, hmm, synth
_lib.check(_lib.load().hb_set_norm3_mode(int(os.environ.get("HB_MODE", "0"))))
N, G = int(os.environ.get("HB_N", 10000)), int(os.environ.get("HB_G", 20000))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(2)
LD = int(os.environ.get("HB_LD", N))
x = torch.empty((G, LD), dtype=torch.float64, device=dev)[:, :N]
x.copy_(torch.randn((G, N), dtype=torch.float64, device=dev, generator=g))
x.mul_(torch.empty(N, dtype=torch.float64, device=dev).uniform_(2.0, 3.1, generator=g))
sd = hmm.chain_sd_dev(x)
Pi, delta, mean = hmm.sym_Pi(3, 1e-6), [0., 1., 0.], [-synth.DEV, 0., synth.DEV]
seg = synth.segments(G)
TILED = int(os.environ.get("HB_TILED", "0"))
if TILED:
    xt = hmm.tile_rows(x)
    run = lambda out=None, **kw: hmm.hmm_norm_tiled_dev(xt, N, sd, mean, Pi, delta, seg_off=seg, out=out, **kw)
    out = run(viterbi=True, gamma=True)
else:
    run = lambda out=None, **kw: hmm.hmm_norm_dev(x, sd, mean, Pi, delta, seg_off=seg, out=out, **kw)
    out = dict(y=torch.empty((G, LD), dtype=torch.uint8, device=dev)[:, :N],
               gamma=torch.empty((3, G, LD), dtype=torch.float64, device=dev)[:, :, :N])
    out = run(out=out, viterbi=True, gamma=True)
res = []
for mode, kw in (("V+FB", dict(viterbi=True, gamma=True)), ("FB", dict(viterbi=False, gamma=True)), ("V", dict(viterbi=True, gamma=False))):
    fn = lambda: run(out=out, **kw)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[3]
    res.append(f"{mode} {ms:.3f} ms {N*G/ms/1e6:.1f} G/s")
hmm.status_dev()
run(out=out, viterbi=True, gamma=True)
print(os.path.basename(sys.argv[1]), "tiled" if TILED else f"ld {LD}", "mode", os.environ.get("HB_MODE", "0"), " | ".join(res),
      "| rerun chains", hmm.norm3_rerun_chains(), flush=True)
