"""BASELINE config 4 through the RefClass mirror: synthetic 10k cells x 20k genes with planted clones,
calcGexpCnvBoundaries(init=True) end to end (device grouping, pooled rounds, device retest, recursion).
Development aid / evidence for DESIGN.md; the parity tests use the bundled data instead."""
import sys, time
import numpy as np, pandas as pd, torch
sys.path.insert(0, ".")
from honeybadger_b200 import honeybadger as hb, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
rng = np.random.default_rng(4)
seg = synth.segments(G)
clone = rng.choice(3, size=N, p=[0.5, 0.3, 0.2])
x = rng.standard_normal((G, N)).astype(np.float64)
x *= rng.uniform(2.0, 3.1, size=N)
dev = synth.DEV
x[seg[6]:seg[7], clone == 0] += dev       # clone A: +chr7
x[seg[9]:seg[10], clone == 0] -= dev      #          -chr10
x[seg[12]:seg[13], clone == 1] -= dev     # clone B: -chr13
genes_names = [f"g{i}" for i in range(G)]
chrom = np.empty(G, dtype=object)
for c in range(22):
    chrom[seg[c]:seg[c + 1]] = f"chr{c + 1}"
pos = np.concatenate([np.arange(seg[c + 1] - seg[c]) * 1000.0 for c in range(22)])
genes = hb.GRanges(chrom, pos, pos + 500, genes_names)
torch.zeros(1, device="cuda").sum().item()   # CUDA context + torch lazy init outside the timed region
t0 = time.perf_counter()
h = hb.HoneyBADGER("c4")
h.genes = genes
h.gexp_norm = pd.DataFrame(x, index=genes_names, columns=[f"c{i}" for i in range(N)])
h.setGexpDev(dev=dev)
t1 = time.perf_counter()
if len(sys.argv) > 3 and sys.argv[3] == "profile":
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    h.calcGexpCnvBoundaries(init=True)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
else:
    h.calcGexpCnvBoundaries(init=True)
    torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"N={N} G={G}: setup {t1-t0:.1f} s, calcGexpCnvBoundaries(init=TRUE) {t2-t1:.1f} s, "
      f"{len(h.bound_genes_final)} regions recorded")
for k, reg in enumerate(h.bound_genes_final):
    names = reg[0]
    idx = [int(n[1:]) for n in (names[0], names[-1])]
    cells = (h.pred_genes_r.loc[names[0]] > 0).values
    comp = np.bincount(clone[cells], minlength=3)
    ch = np.searchsorted(seg, idx[0], side="right")
    print(f"  region {k}: genes {idx[0]}..{idx[1]} (chr{ch}), {cells.sum()} cells, clone mix {comp.tolist()}")
