"""BASELINE config 3 through the RefClass mirror: synthetic 10k cells x 200k het SNPs with planted clones,
calcAlleleCnvBoundaries(init=True) end to end (device grouping, pooled rounds on the long-chain path,
snpModel-marginal retest, recursion).  Development aid / evidence for DESIGN.md."""
import sys, time
import numpy as np, pandas as pd, torch
sys.path.insert(0, ".")
from honeybadger_b200 import honeybadger as hb, synth, hmm

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(3)
torch.manual_seed(3)  # torch._standard_gamma below draws from the DEFAULT generator: without this the data differ run to run
seg = synth.segments(G)
# tumour-like mix: clone A 60 %, clone B 20 %, normal cells 20 %.  A truncal deletion (chr14, clones A and B)
# is visible in the all-cell pooled chain (pooled lesser-allele rate 0.17); the sub-clonal one (chr10, clone A)
# only after the first split — the recursion has to find it.
clone = torch.multinomial(torch.tensor([0.6, 0.2, 0.2], device=dev), N, True, generator=g)
p = torch.full((G, 3), 0.45, dtype=torch.float32, device=dev)
# the chains see cells WITH the lesser allele among covered cells (rowSums(r.maf > 0), R/HoneyBADGER.R:1404):
# with 1-5 reads per covered site that is ~0.65 at a neutral het SNP and ~0.07 where the allele is lost
PDEL = 0.02
p[seg[13]:seg[14], 0] = PDEL     # clones A, B: -chr14
p[seg[13]:seg[14], 1] = PDEL
p[seg[9]:seg[10], 0] = PDEL      # clone A: -chr10
r = torch.empty((G, N), dtype=torch.int16, device=dev); n = torch.empty((G, N), dtype=torch.int16, device=dev)
rows = max(1, (1 << 26) // N)
for a in range(0, G, rows):
    b = min(G, a + rows); shp = (b - a, N)
    cov = torch.rand(shp, device=dev, generator=g) < 0.108
    nb = torch.poisson(torch._standard_gamma(torch.full(shp, 0.3, device=dev)) * 10.0, generator=g)
    nn = torch.where(cov, (1 + nb).clamp(max=1245), torch.zeros_like(nb))
    rr = torch.binomial(nn, p[a:b][:, clone], generator=g)
    # lesser-allele convention of setAlleleMats: r <= n - r in the bulk is not needed for the chain itself
    n[a:b] = nn.to(torch.int16); r[a:b] = rr.to(torch.int16)
clone_h = clone.cpu().numpy()
torch.zeros(1, device="cuda").sum().item()   # CUDA context + torch lazy init outside the timed region
t0 = time.perf_counter()
names = np.array([f"chr{np.searchsorted(seg, i, side='right')}:{1000 * (i + 1)}:{1000 * (i + 1)}" for i in range(G)], dtype=object)
cells = [f"c{i}" for i in range(N)]
h = hb.HoneyBADGER("c3")
h.r_maf = pd.DataFrame(r.cpu().numpy(), index=names, columns=cells)
h.n_sc = pd.DataFrame(n.cpu().numpy(), index=names, columns=cells)
h.l_maf = (r > 0).sum(1).double().cpu().numpy()
h.n_bulk = (n > 0).sum(1).double().cpu().numpy()
del r, n
torch.cuda.empty_cache()
t1 = time.perf_counter()
if len(sys.argv) > 3 and sys.argv[3] == "profile":
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    h.calcAlleleCnvBoundaries(init=True)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
else:
    h.calcAlleleCnvBoundaries(init=True)
    torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"N={N} G={G}: frames {t1-t0:.1f} s, calcAlleleCnvBoundaries(init=TRUE) {t2-t1:.1f} s, "
      f"{len(h.bound_snps_final)} regions recorded", flush=True)
for k, reg in enumerate(h.bound_snps_final):
    nm = reg[0]
    cellsel = (h.pred_snps_r.loc[nm[0]] > 0).values
    print(f"  region {k}: {nm[0]} .. {nm[-1]} ({len(nm)} SNPs), {cellsel.sum()} cells, clone mix "
          f"{np.bincount(clone_h[cellsel], minlength=3).tolist()}")
print("last long-chain launch:", hmm.longchain_stats())
