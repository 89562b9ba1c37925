"""Quick device timing probe (development aid; bench.py is the contract)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import hmm, synth

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[n // 2]

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(2)
x = torch.randn((G, N), dtype=torch.float64, device=dev, generator=g) * 2.5
sd = hmm.chain_sd_dev(x)
Pi, delta, mean = hmm.sym_Pi(3, 1e-6), [0., 1., 0.], [-synth.DEV, 0., synth.DEV]
for nseg in (22, 1):
    seg = synth.segments(G) if nseg == 22 else np.array([0, G], dtype=np.int64)
    out = hmm.hmm_norm_dev(x, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True)
    for mode, kw in (("V+FB", dict(viterbi=True, gamma=True)), ("V", dict(viterbi=True, gamma=False)), ("FB", dict(viterbi=False, gamma=True))):
        ms = timeit(lambda: hmm.hmm_norm_dev(x, sd, mean, Pi, delta, seg_off=seg, out=out, **kw))
        steps = N * G
        bps = {"V+FB": 33, "V": 9, "FB": 32}[mode]
        print(f"norm3 nseg={nseg} {mode}: {ms:.3f} ms  {steps/ms/1e6:.1f} Gsteps/s  alg {steps*bps/ms/1e6:.0f} GB/s", flush=True)
    hmm.status_dev()
del out, x
Gs = G * 4
n = (torch.rand((Gs, N), device=dev, generator=g) < 0.108).to(torch.int32) * torch.randint(1, 6, (Gs, N), device=dev, generator=g, dtype=torch.int32)
r = (torch.rand((Gs, N), device=dev, generator=g) * (n + 1).double()).floor().to(torch.int32).clamp(max=n)
Pi2 = hmm.sym_Pi(2, 1e-6)
for nseg in (22, 1):
    seg = synth.segments(Gs) if nseg == 22 else np.array([0, Gs], dtype=np.int64)
    out = hmm.hmm_binom_dev(r, n, [0.1, 0.45], Pi2, [0., 1.], seg_off=seg)
    for mode, kw in (("V+FB", dict(viterbi=True, gamma=True)), ("V", dict(viterbi=True, gamma=False))):
        ms = timeit(lambda: hmm.hmm_binom_dev(r, n, [0.1, 0.45], Pi2, [0., 1.], seg_off=seg, out=out, **kw))
        steps = N * Gs
        bps = {"V+FB": 25, "V": 9}[mode]
        print(f"binom2 nseg={nseg} {mode}: {ms:.3f} ms  {steps/ms/1e6:.1f} Gsteps/s  alg {steps*bps/ms/1e6:.0f} GB/s", flush=True)
    hmm.status_dev()
