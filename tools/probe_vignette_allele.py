"""Getting_Started vignette, allele half, through the RefClass mirror on the bundled data (raw fixture):
setAlleleMats(het.deviance.threshold=0.1) -> calcAlleleCnvBoundaries(init=TRUE) -> range(snps[unlist(bsf)]).
docs/Getting_Started.md:287-293 prints chr10 [3180316, 126670356], chr13 [21949267, 114175032],
chr14 [20216560, 73465003], chr19 [20727463, 48598823] (development aid)."""
import sys
import numpy as np, pandas as pd
sys.path.insert(0, ".")
from honeybadger_b200 import honeybadger as hb
d = np.load("tests/golden/inputs_raw_mgh31.npz")
r, cov = d["r"].astype(np.int32), d["cov"].astype(np.int32)
names = [f"{c}:{p}" for c, p in zip(d["chrom"], d["pos"])]
cells = [f"c{i}" for i in range(r.shape[1])]
o = hb.HoneyBADGER("MGH31")
o.setAlleleMats(pd.DataFrame(r, index=names, columns=cells), pd.DataFrame(cov, index=names, columns=cells),
                het_deviance_threshold=0.1, verbose=False)
for hook in (None, hb.retest_allele_lr):
    o.retest_allele = hook
    o.calcAlleleCnvBoundaries(init=True)
    found = [s for lst in o.bound_snps_final for s in lst[0]]
    print("retest:", "model" if hook is None else "lr", "| levels recorded:", len(o.bound_snps_final))
    print(o.snps[found].range())
    for lst in o.bound_snps_final:
        bs = lst[0]
        ncell = int((o.pred_snps_r.loc[bs[0]].values > 0).sum())
        print("   ", bs[0], "..", bs[-1], len(bs), "snps,", ncell, "cells")
