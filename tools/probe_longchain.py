"""Timing of the gene-axis parallel Viterbi for a few long chains (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from honeybadger_b200 import hmm
from test_longchain import binom_lb, norm_lb, sym_Pi

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

rng = np.random.default_rng(0)
for S, T, K in [(2, 200000, 6), (3, 20000, 15), (2, 4550, 6), (3, 6082, 15)]:
    lb = torch.from_numpy(binom_lb(rng, K, T) if S == 2 else norm_lb(rng, K, T)).cuda()
    Pi, delta = sym_Pi(S, 1e-6), np.array([0., 1.] if S == 2 else [0., 1., 0.])
    for L in (0, 16, 32, 64, 128, 256):
        ms = timeit(lambda: hmm.viterbi_long_logb_dev(lb, Pi, delta, L))
        print(f"S={S} T={T} K={K} L={L}: {ms:.3f} ms  {hmm.longchain_stats()}", flush=True)

print("posteriors (lc_fb_*)")
for S, T, K in [(2, 200000, 6), (3, 20000, 15), (3, 6082, 15)]:
    lb = torch.from_numpy(binom_lb(rng, K, T) if S == 2 else norm_lb(rng, K, T)).cuda()
    Pi, delta = sym_Pi(S, 1e-6), np.array([0., 1.] if S == 2 else [0., 1., 0.])
    for L in (0, 64, 128, 512):
        ms = timeit(lambda: hmm.forwardback_long_logb_dev(lb, Pi, delta, L))
        print(f"FB S={S} T={T} K={K} L={L}: {ms:.3f} ms", flush=True)
