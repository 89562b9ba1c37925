"""One pooled allele round at C4 sizes (ncu target)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from honeybadger_b200 import hmm
from honeybadger_b200 import honeybadger as hb
dev = torch.device("cuda:0")
N, G = 10000, 200000
n = hmm.alloc_rows(G, N, torch.int32, dev); r = hmm.alloc_rows(G, N, torch.int32, dev)
n.copy_(((torch.rand((G, N), device=dev) < 0.108).int() * torch.randint(1, 6, (G, N), device=dev, dtype=torch.int32)))
r.copy_((torch.rand((G, N), device=dev) * (n + 1).float()).floor().int().clamp(max=n))
idx = np.arange(N)
member = np.stack([np.ones(N, bool)] + [idx % 2 == k for k in range(2)] + [idx % 3 == k for k in range(3)])
for _ in range(3):
    y, mafl, sizel = hb.allele_round_dev(r, n, member, 0.1, 0.45, 1e-6)
print(hmm.longchain_stats())
