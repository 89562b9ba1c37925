"""One long-chain launch sequence (for an ncu launch list): S T K L from argv."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from honeybadger_b200 import hmm
from test_longchain import binom_lb, norm_lb, sym_Pi
S, T, K, L = (int(v) for v in sys.argv[1:5])
rng = np.random.default_rng(0)
lb = torch.from_numpy(binom_lb(rng, K, T) if S == 2 else norm_lb(rng, K, T)).cuda()
Pi, delta = sym_Pi(S, 1e-6), np.array([0., 1.] if S == 2 else [0., 1., 0.])
for _ in range(3):
    y = hmm.viterbi_long_logb_dev(lb, Pi, delta, L)
torch.cuda.synchronize()
print(hmm.longchain_stats())
