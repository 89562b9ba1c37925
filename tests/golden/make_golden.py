#!/usr/bin/env python
"""Generate the committed fixtures in tests/golden/ from the reference's bundled example data.

Run in the build container only (needs /root/reference/data/*.rda, absent on the GPU box):

    python tests/golden/make_golden.py

What is pinned, and by what:
  * densities.json      — dnorm/dbinom log-densities computed with mpmath (50 digits): an
                          arithmetic-independent anchor for the nmath restatements.
  * allele_mgh31.npz    — setAlleleMats (R/HoneyBADGER.R:485-614) output for het threshold 0.1 on
                          data/r.rda + data/cov.sc.rda.  The reference's own printed known answers
                          (docs/Getting_Started.md:230 "5359", docs/Integrating.md:27 "5552") are
                          asserted here.  Holds r.maf / n.sc (uint16), SNP coordinates, group labels
                          for 3 cut heights, pooled (mafl, sizel), the Viterbi state paths of every
                          pooled chain and the function-variant boundary runs.
  * gexp_mgh31.npz      — gexp.norm = gexp - ref (R/HoneyBADGER.R:160-161, filter/scale off) for the
                          first 40 cells rounded to float32, labels for 5 cut heights, pooled means,
                          sd, Viterbi paths, sampled posteriors with dev = 1.015658
                          (docs/Getting_Started.md:108).  Gene coordinates are not bundled (biomaRt),
                          so genes stay in stored order, split into 22 synthetic segments.
  * inputs_raw_mgh31.npz — raw r / cov.sc (uint16) and gexp[:800] / ref[:800] for the input
                          construction (`python tests/golden/make_golden.py inputs` writes only this).
State paths / posteriors are produced by oracle/hb_oracle.c and REQUIRED to equal the independent
pure-Python restatement (oracle/pyref.py) before they are written.  Group labels come from scipy's
Ward linkage on a simplified smoothing — they are inputs of the path, not a claim about R's hclust.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle import pyref as P  # noqa: E402
from oracle.rda import as_matrix, read_rda  # noqa: E402

REF = "/root/reference/data"
DEV = 1.015658  # docs/Getting_Started.md:108


def rle(y):
    y = np.asarray(y)
    starts = np.concatenate([[0], np.nonzero(np.diff(y))[0] + 1])
    return starts.astype(np.int32), y[starts].astype(np.int8)


def runmean_nan(m, k):
    """Centered window mean along axis 0, shrinking at the ends, NaN skipped."""
    G = m.shape[0]
    ok = ~np.isnan(m)
    v = np.where(ok, m, 0.0)
    cs = np.vstack([np.zeros((1,) + m.shape[1:]), np.cumsum(v, 0)])
    cn = np.vstack([np.zeros((1,) + m.shape[1:]), np.cumsum(ok, 0)])
    h = k // 2
    lo = np.clip(np.arange(G) - h, 0, G)
    hi = np.clip(np.arange(G) + h + 1, 0, G)
    with np.errstate(all="ignore"):
        return (cs[hi] - cs[lo]) / (cn[hi] - cn[lo])


def ward_labels(mat_smooth, heights):
    from scipy.cluster.hierarchy import cut_tree, linkage
    X = np.nan_to_num(mat_smooth.T, nan=0.0, posinf=0.0, neginf=0.0)
    Z = linkage(X, method="ward")
    out = []
    for h in heights:
        lab = cut_tree(Z, n_clusters=h)[:, 0]
        # cutree numbers clusters by first appearance
        remap, nxt = {}, 1
        res = np.zeros_like(lab)
        for i, v in enumerate(lab):
            if v not in remap:
                remap[v] = nxt
                nxt += 1
            res[i] = remap[v]
        out.append(res.astype(np.int8))
    return np.stack(out)


def make_densities():
    pts_b, pts_n = [], []
    for n in [1, 2, 3, 4, 5, 7, 10, 15, 16, 20, 31, 32, 35, 36, 50, 80, 81, 137, 500, 501, 1245,
              10000]:
        for x in sorted({0, 1, 2, n // 4, n // 2, n - 1, n}):
            if 0 <= x <= n:
                for p in (0.1, 0.45):
                    pts_b.append([x, n, p, float(P.mp_dbinom_log(x, n, p))])
    rng = np.random.default_rng(11)
    for _ in range(200):
        x = float(rng.normal() * 4)
        mu = float(rng.choice([-DEV, 0.0, DEV]))
        sd = float(rng.uniform(0.4, 3.2))
        pts_n.append([x, mu, sd, float(P.mp_dnorm_log(x, mu, sd))])
    json.dump({"dbinom_log": pts_b, "dnorm_log": pts_n, "source": "mpmath 50 digits"},
              open(os.path.join(HERE, "densities.json"), "w"))


def set_allele_mats(r, cov, thr, min_cell=3):
    """R/HoneyBADGER.R:485-614 with an in-silico bulk (filter=TRUE)."""
    l = (r > 0).sum(1)
    nb = (cov > 0).sum(1)
    with np.errstate(all="ignore"):
        E = l / nb
    vi = (E > thr) & (E < 1 - thr)
    printed = int(vi.sum())
    r, cov, l, nb = r[vi], cov[vi], l[vi], nb[vi]
    keep_idx = np.nonzero(vi)[0]
    vi2 = (cov > 0).sum(1) >= min_cell
    r, cov, l, nb = r[vi2], cov[vi2], l[vi2], nb[vi2]
    keep_idx = keep_idx[vi2]
    E = l / nb
    flip = E > 0.5
    r_maf = np.where(flip[:, None], cov - r, r)
    l_maf = np.where(flip, nb - l, l)
    return r_maf, cov, l_maf, nb, keep_idx, printed


def make_allele():
    r, rn, cn = as_matrix(read_rda(f"{REF}/r.rda")["r"])
    cov, _, _ = as_matrix(read_rda(f"{REF}/cov.sc.rda")["cov.sc"])
    known = {}
    for thr in (0.05, 0.1):
        out = set_allele_mats(r, cov, thr)
        known[str(thr)] = [out[5], int(out[0].shape[0])]
    assert known["0.05"] == [5552, 4550], known  # docs/Integrating.md:27
    assert known["0.1"] == [5359, 4357], known   # docs/Getting_Started.md:230
    r_maf, n_sc, l_maf, n_bulk, keep_idx, _ = set_allele_mats(r, cov, 0.1)
    names = [rn[i] for i in keep_idx]
    chrom = np.array([int(s.split(":")[0]) for s in names], dtype=np.uint8)
    pos = np.array([int(s.split(":")[1]) for s in names], dtype=np.int32)
    with np.errstate(all="ignore"):
        mat_tot = r_maf / n_sc
    labels = ward_labels(runmean_nan(mat_tot, 31), [1, 2, 3])
    res = P.calc_allele_cnv_boundaries_fn(r_maf, n_sc, labels)
    # paths: oracle C == independent Python, per pooled chain
    Pi = O.sym_Pi(2, 1e-6)
    paths_s, paths_v, paths_off, mafl, sizel, chain_hg = [], [], [0], [], [], []
    for h, per_group in enumerate(res["calls"]):
        for gi, c in enumerate(per_group):
            if c is None:
                continue
            y2 = P.viterbi_binom(c["mafl"], c["sizel"], [0.1, 0.45], Pi, [0.0, 1.0])
            assert (y2 == c["y"]).all(), "oracle C != pyref on pooled allele chain"
            s, v = rle(c["y"])
            paths_s.append(s)
            paths_v.append(v)
            paths_off.append(paths_off[-1] + s.size)
            mafl.append(c["mafl"])
            sizel.append(c["sizel"])
            chain_hg.append([h + 1, gi + 1])
    # the survey's probe: all-cell chain has state-1 runs on chr1 (9 SNPs) and chr13..14 (54 SNPs)
    y_all = res["calls"][0][0]["y"]
    s, v = rle(y_all)
    ends = np.concatenate([s[1:], [y_all.size]])
    runs1 = [(names[a], names[b - 1], int(b - a)) for a, b, vv in zip(s, ends, v) if vv == 1]
    print("all-cell allele chain state-1 runs:", runs1)
    info = res["info"] or []
    np.savez_compressed(
        os.path.join(HERE, "allele_mgh31.npz"),
        r_maf=r_maf.astype(np.uint16), n_sc=n_sc.astype(np.uint16), chrom=chrom, pos=pos,
        l_maf=l_maf.astype(np.int32), n_bulk=n_bulk.astype(np.int32),
        labels=labels, chain_hg=np.array(chain_hg, dtype=np.int8),
        mafl=np.stack(mafl).astype(np.int32), sizel=np.stack(sizel).astype(np.int32),
        path_starts=np.concatenate(paths_s), path_vals=np.concatenate(paths_v),
        path_off=np.array(paths_off, dtype=np.int64),
        info_idx=np.concatenate(info).astype(np.int32) if info else np.zeros(0, np.int32),
        info_off=np.cumsum([0] + [a.size for a in info]).astype(np.int64),
        known_counts=np.array([known["0.05"], known["0.1"]], dtype=np.int32),
    )
    return runs1


def make_gexp():
    gexp, gn, cn = as_matrix(read_rda(f"{REF}/gexp.rda")["gexp"])
    ref = np.asarray(read_rda(f"{REF}/ref.rda")["ref"]["value"])
    assert gexp.shape == (6082, 75)  # docs/Getting_Started.md:79
    gexp_norm = (gexp - ref[:, None])[:, :40].astype(np.float32).astype(np.float64)
    G = gexp_norm.shape[0]
    labels = ward_labels(runmean_nan(gexp_norm, 101), [1, 2, 3, 4, 5])
    # 22 synthetic segments ∝ bundled per-chromosome SNP counts
    snp_counts = np.array([704, 392, 379, 268, 279, 431, 388, 194, 223, 94, 328, 331, 35, 82, 165,
                           218, 311, 81, 338, 150, 77, 119], dtype=np.float64)
    seg_off = np.concatenate([[0], np.round(np.cumsum(snp_counts) / snp_counts.sum() * G)])
    seg_off = seg_off.astype(np.int64)
    calls = P.gexp_hmm_calls(gexp_norm, labels, DEV)
    Pi = O.sym_Pi(3, 1e-6)
    xs, sds, ps, pv, poff, gam, lls = [], [], [], [], [0], [], []
    samp = np.arange(0, G, 53)
    for c in calls:
        y2 = P.viterbi_norm(c["x"], [-DEV, 0.0, DEV], [c["sd"]] * 3, Pi, [0.0, 1.0, 0.0])
        assert (y2 == c["y"]).all(), "oracle C != pyref on pooled expression chain"
        assert c["sd"] == P.r_sd(c["x"])
        g, ll = O.forwardback_norm(c["x"], [-DEV, 0.0, DEV], [c["sd"]] * 3, Pi, [0.0, 1.0, 0.0])
        s, v = rle(c["y"])
        xs.append(c["x"]); sds.append(c["sd"]); ps.append(s); pv.append(v)
        poff.append(poff[-1] + s.size); gam.append(g[samp]); lls.append(ll)
    res = P.calc_gexp_cnv_boundaries_fn(gexp_norm, labels[:3], m=DEV, min_num_genes=3, trim=0.1)

    def pack(lst):
        lst = lst or []
        return (np.concatenate(lst).astype(np.int32) if lst else np.zeros(0, np.int32),
                np.cumsum([0] + [a.size for a in lst]).astype(np.int64))

    amp_idx, amp_off = pack(res["amp"])
    del_idx, del_off = pack(res["del"])
    np.savez_compressed(
        os.path.join(HERE, "gexp_mgh31.npz"),
        gexp_norm_f32=gexp_norm.astype(np.float32), labels=labels, seg_off=seg_off,
        dev=np.float64(DEV), pooled_x=np.stack(xs), pooled_sd=np.array(sds),
        path_starts=np.concatenate(ps), path_vals=np.concatenate(pv),
        path_off=np.array(poff, dtype=np.int64), gamma_samp_idx=samp.astype(np.int32),
        gamma_samp=np.stack(gam), ll=np.array(lls),
        fn_amp_idx=amp_idx, fn_amp_off=amp_off, fn_del_idx=del_idx, fn_del_off=del_off,
    )
    print("gexp chains:", len(calls), "amp runs:", len(res["amp"] or []),
          "del runs:", len(res["del"] or []))


def make_inputs_raw():
    """Raw inputs of setAlleleMats / setGexpMats: the bundled r / cov.sc count matrices in
    full and the first 800 genes of gexp / ref, so that the device construction can be checked against
    the reference's printed known answers (5359/4357 at 0.1, 5552/4550 at 0.05) on the GPU box."""
    r, rn, cn = as_matrix(read_rda(f"{REF}/r.rda")["r"])
    cov, _, _ = as_matrix(read_rda(f"{REF}/cov.sc.rda")["cov.sc"])
    assert (r == np.rint(r)).all() and (cov == np.rint(cov)).all() and cov.max() < 65536
    gexp, gn, _ = as_matrix(read_rda(f"{REF}/gexp.rda")["gexp"])
    ref = np.asarray(read_rda(f"{REF}/ref.rda")["ref"]["value"])
    chrom = np.array([int(s.split(":")[0]) for s in rn], dtype=np.uint8)
    pos = np.array([int(s.split(":")[1]) for s in rn], dtype=np.int32)
    # docs/figure/Integrating/unnamed-chunk-6-1.png: the panels plotAlleleProfile draws for
    # region = calcAlleleCnvBoundaries(...)$region on this data (setAlleleMats defaults, i.e. threshold
    # 0.05) are region@seqnames@values (R/HoneyBADGER_allele.R:248) = chr1 chr3 chr9 chr10 chr12 chr13
    # chr14 chr15 chr19 — the reference's own rendered answer for the stateless allele path
    np.savez_compressed(os.path.join(HERE, "inputs_raw_mgh31.npz"), r=r.astype(np.uint16),
                        cov=cov.astype(np.uint16), chrom=chrom, pos=pos, gexp=gexp[:800], ref=ref[:800],
                        known_counts=np.array([[5552, 4550], [5359, 4357]], dtype=np.int32),
                        figure_region_chroms=np.array([1, 3, 9, 10, 12, 13, 14, 15, 19], dtype=np.uint8))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "inputs":
        make_inputs_raw()
        sys.exit(0)
    make_inputs_raw()
    make_densities()
    runs = make_allele()
    make_gexp()
    json.dump({"allele_all_cell_state1_runs": runs,
               # docs/Getting_Started.md:287-293, printed by the reference after the RefClass allele run
               "getting_started_allele_regions": [["chr10", 3180316, 126670356], ["chr13", 21949267, 114175032],
                                                  ["chr14", 20216560, 73465003], ["chr19", 20727463, 48598823]],
               "getting_started_allele_regions_source":
                   "docs/Getting_Started.md:287-293: print(range(snps[unlist(bound.snps.final)])) after "
                   "hb$calcAlleleCnvBoundaries(init=TRUE) on the bundled data (het.deviance.threshold=0.1); "
                   "the recursion's splits come from JAGS, so cell sets are stochastic, the HMM boundaries are not"},
              open(os.path.join(HERE, "known_answers.json"), "w"), indent=1)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
