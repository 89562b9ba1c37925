"""GPU parity of the pooled (reference-shaped) rounds and of the driver, against goldens + oracle."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pandas as pd
import pytest

from honeybadger_b200 import synth
from oracle import pyref as P

pytestmark = pytest.mark.gpu


def unrle(starts, vals, n):
    y = np.zeros(n, dtype=np.int32)
    for a, b, v in zip(starts, list(starts[1:]) + [n], vals):
        y[a:b] = v
    return y


def test_gexp_pooled_round_on_bundled_data(golden_gexp, oracle):
    """15 pooled chains of the MGH31 example (5 cut heights): pooled means and sd bit-identical to
    R's long-double arithmetic, Viterbi paths identical."""
    from honeybadger_b200 import honeybadger as hb
    g = golden_gexp
    x = g["gexp_norm_f32"].astype(np.float64)
    member, _ = hb._membership(list(g["labels"]), x.shape[1])
    assert member.shape[0] == 15
    dev = float(g["dev"])
    y, px, psd = hb.gexp_round_dev(hb._to_dev_f64(x), member, [-dev, 0.0, dev], 1e-6)
    assert (px == g["pooled_x"]).all(), "pooled means differ from R mean() semantics"
    assert (psd == g["pooled_sd"]).all(), "sd differs from R sd() semantics"
    for k in range(15):
        a, b = g["path_off"][k], g["path_off"][k + 1]
        assert (y[k] == unrle(g["path_starts"][a:b], g["path_vals"][a:b], y.shape[1])).all()


def test_allele_pooled_round_on_bundled_data(golden_allele):
    from honeybadger_b200 import honeybadger as hb
    g = golden_allele
    r, n = g["r_maf"].astype(np.float64), g["n_sc"].astype(np.float64)
    member, tags = hb._membership(list(g["labels"]), r.shape[1])
    y, mafl, sizel = hb.allele_round_dev(hb._to_dev_i32(r), hb._to_dev_i32(n), member, 0.1, 0.45, 1e-6)
    assert (mafl == g["mafl"]).all() and (sizel == g["sizel"]).all()
    for k in range(member.shape[0]):
        a, b = g["path_off"][k], g["path_off"][k + 1]
        assert (y[k] == unrle(g["path_starts"][a:b], g["path_vals"][a:b], y.shape[1])).all()
    # the vignette's call: all-cell chain, state 1 on chr1 (9 SNPs) and chr13..14 (54 SNPs)
    s = np.nonzero(np.diff(np.concatenate([[0], (y[0] == 1).astype(int), [0]])))[0]
    assert (s[1::2] - s[0::2]).tolist() == [9, 54]


def test_host_wrappers_match_dev_rounds(golden_gexp, golden_allele):
    from honeybadger_b200 import _lib, honeybadger as hb
    from honeybadger_b200.hmm import sym_Pi
    lib = _lib.load()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    g = golden_gexp
    x = np.asfortranarray(g["gexp_norm_f32"].astype(np.float64))
    member, _ = hb._membership(list(g["labels"][:2]), x.shape[1])
    K, (G, N) = member.shape[0], x.shape
    y = np.zeros((K, G), np.int32)
    px, psd = np.zeros((K, G)), np.zeros(K)
    dev = float(g["dev"])
    mean, Pi, delta = np.array([-dev, 0, dev]), sym_Pi(3, 1e-6), np.array([0., 1, 0])
    _lib.check(lib.hb_gexp_pooled_viterbi_host(vp(x), G, N, vp(member), K, vp(mean), vp(Pi), vp(delta),
                                               vp(y), vp(px), vp(psd)))
    assert (px == g["pooled_x"][:K]).all() and (psd == g["pooled_sd"][:K]).all()
    ga = golden_allele
    r = np.asfortranarray(ga["r_maf"].astype(np.float64))
    n = np.asfortranarray(ga["n_sc"].astype(np.float64))
    member, _ = hb._membership(list(ga["labels"]), r.shape[1])
    K, (G, N) = member.shape[0], r.shape
    y = np.zeros((K, G), np.int32)
    mafl, sizel = np.zeros((K, G), np.int32), np.zeros((K, G), np.int32)
    prob, Pi2, d2 = np.array([0.1, 0.45]), sym_Pi(2, 1e-6), np.array([0., 1.])
    _lib.check(lib.hb_allele_pooled_viterbi_host(vp(r), vp(n), G, N, vp(member), K, vp(prob), vp(Pi2),
                                                 vp(d2), vp(y), vp(mafl), vp(sizel)))
    assert (mafl == ga["mafl"]).all() and (sizel == ga["sizel"]).all()


def test_pool_mean_random_groups_bit_exact(oracle):
    """x87 replay on the device == R mean() for ragged random groups (incl. a 2-cell group)."""
    import torch
    from honeybadger_b200 import _lib, honeybadger as hb
    rng = np.random.default_rng(12)
    G, N = 777, 203
    x = rng.normal(size=(G, N)) * 2.7 + rng.normal(size=(G, 1))
    labels = rng.integers(1, 6, N)
    labels[:2] = 9
    member, _ = hb._membership([np.ones(N, int), labels], N)
    xd = hb._to_dev_f64(x)
    md = torch.as_tensor(member).cuda()
    out = torch.empty((member.shape[0], G), dtype=torch.float64, device="cuda")
    lib = _lib.load()
    _lib.check(lib.hb_pool_mean_dev(xd.data_ptr(), xd.stride(0), G, N, md.data_ptr(), member.shape[0],
                                    out.data_ptr(), None))
    sd = torch.empty(member.shape[0], dtype=torch.float64, device="cuda")
    _lib.check(lib.hb_pooled_sd_dev(out.data_ptr(), G, member.shape[0], sd.data_ptr(), None))
    got, gsd = out.cpu().numpy(), sd.cpu().numpy()
    for k in range(member.shape[0]):
        ref = oracle.pool_mean(x, np.nonzero(member[k])[0])
        assert (got[k] == ref).all()
        assert gsd[k] == oracle.r_sd(ref)


def test_pool_mean_zero_inflated_rows_bit_exact(oracle):
    """Rows of zeros, rows that start with zeros, zero-inflated rows, integer counts, a constant row: the
    double-pair accumulation (zero operands return the other operand, as the integer routine does) must give R's
    mean down to the sign of a zero."""
    import torch
    from honeybadger_b200 import _lib, honeybadger as hb
    rng = np.random.default_rng(13)
    G, N = 300, 157
    x = rng.gamma(2.0, 1.5, size=(G, N)) * (rng.random((G, N)) < 0.3)
    x[:20] = 0.0
    x[20:40, :70] = 0.0
    x[40:60] = rng.integers(0, 4, size=(20, N))
    x[60:70] = 1.25
    x[70:80] = -x[70:80]
    labels = rng.integers(1, 4, N)
    member, _ = hb._membership([np.ones(N, int), labels], N)
    xd = hb._to_dev_f64(x)
    md = torch.as_tensor(member).cuda()
    out = torch.empty((member.shape[0], G), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().hb_pool_mean_dev(xd.data_ptr(), xd.stride(0), G, N, md.data_ptr(), member.shape[0],
                                            out.data_ptr(), None))
    got = out.cpu().numpy()
    for k in range(member.shape[0]):
        ref = np.asarray(oracle.pool_mean(x, np.nonzero(member[k])[0]), dtype=np.float64)
        assert np.array_equal(got[k].view(np.int64), ref.view(np.int64))


def test_pooled_sd_adversarial_rows_bit_exact(oracle):
    """hb_pooled_sd_dev (running 80-bit sums added block-wise as integer prefix sums across the warp, literal
    x80_add for ties / oversized terms / binade changes) == R's sd() on rows built to hit every branch: wide
    exponent ranges, integers, tie-heavy increments at four accumulator scales, powers of two pulled just below
    themselves, runs of zeros, a constant row, sums that cancel — and on very short vectors."""
    import torch
    from honeybadger_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(43)
    G = 6000
    rows = [rng.normal(size=G) * 0.3, rng.normal(size=G) * 10.0 ** rng.integers(-12, 12, G),
            rng.integers(0, 50, G).astype(float),
            np.ldexp(rng.integers(1, 1 << 20, G).astype(float), rng.integers(-70, 70, G)), rng.gamma(2.0, 1.5, G),
            np.concatenate([np.zeros(1000), rng.normal(size=G - 1500), np.zeros(500)]), np.zeros(G) + 0.1,
            np.concatenate([[float(1 << 23)], np.ldexp(rng.integers(1, 1 << 30, G - 1).astype(float), -40)])]
    for k in (0, 5, 11, 30):
        rows.append(np.concatenate([[2.0 ** k], np.ldexp(rng.integers(-(1 << 14), 1 << 14, G - 1).astype(float),
                                                         k - 64 - rng.integers(0, 4, G - 1))]))
    pw = []
    for _ in range(G // 3):
        k = int(rng.integers(-20, 20))
        pw += [2.0 ** k, -2.0 ** (k - int(rng.integers(50, 70))), -2.0 ** k * (1 - 2.0 ** -int(rng.integers(1, 53)))]
    rows.append(np.array(pw))
    y = rng.normal(size=G // 2)
    z = np.empty(G)
    z[0::2] = y
    z[1::2] = -y + np.ldexp(rng.integers(-8, 8, G // 2).astype(float), -70)
    rows.append(z)
    X = np.ascontiguousarray(np.stack(rows))

    def run(M):
        d = torch.from_numpy(np.ascontiguousarray(M)).cuda()
        out = torch.empty(M.shape[0], dtype=torch.float64, device="cuda")
        _lib.check(lib.hb_pooled_sd_dev(d.data_ptr(), M.shape[1], M.shape[0], out.data_ptr(), None))
        return out.cpu().numpy()

    got = run(X)
    for k in range(X.shape[0]):
        ref = oracle.r_sd(X[k])
        assert got[k] == ref or (got[k] != got[k] and ref != ref), k
    for g in (2, 3, 31, 32, 33, 65):
        M = rng.normal(size=(4, g)) * 2.6 + 7.5
        got = run(M)
        assert all(got[k] == oracle.r_sd(M[k]) for k in range(4))


def test_function_variants_match_reference_restatement(golden_gexp, golden_allele):
    """calcGexpCnvBoundaries / calcAlleleCnvBoundaries (stateless) == the literal R restatement."""
    from honeybadger_b200 import honeybadger as hb
    ga = golden_allele
    names = [f"{c}:{p}" for c, p in zip(ga["chrom"], ga["pos"])]
    cells = [f"c{i}" for i in range(ga["r_maf"].shape[1])]
    r = pd.DataFrame(ga["r_maf"].astype(float), index=names, columns=cells)
    n = pd.DataFrame(ga["n_sc"].astype(float), index=names, columns=cells)
    labels = list(ga["labels"])
    res = hb.calcAlleleCnvBoundaries(r, n, grouping=lambda R, Nn, hs: labels)
    ref = P.calc_allele_cnv_boundaries_fn(ga["r_maf"].astype(float), ga["n_sc"].astype(float), labels)
    assert [list(np.asarray(names, object)[i]) for i in ref["info"]] == res["info"]
    assert {rg[0] for rg in res["regions"]} >= {"chr13", "chr14"}  # the vignette's deletion calls
    # expression: genes in stored order on 22 synthetic chromosomes
    gg = golden_gexp
    x = gg["gexp_norm_f32"].astype(np.float64)
    G = x.shape[0]
    seg = gg["seg_off"]
    chrom = np.empty(G, dtype=object)
    for s in range(22):
        chrom[seg[s]:seg[s + 1]] = f"chr{s + 1}"
    gnames = [f"g{i}" for i in range(G)]
    genes = hb.GRanges(chrom, np.arange(G) * 10.0, np.arange(G) * 10.0 + 5, gnames)
    xdf = pd.DataFrame(x, index=gnames, columns=[f"c{i}" for i in range(x.shape[1])])
    labels3 = list(gg["labels"][:3])
    res = hb.calcGexpCnvBoundaries(xdf, genes, m=float(gg["dev"]), grouping=lambda M, hs: labels3)
    for key, idx, off in (("amp", gg["fn_amp_idx"], gg["fn_amp_off"]), ("del", gg["fn_del_idx"], gg["fn_del_off"])):
        want = [[gnames[i] for i in idx[a:b]] for a, b in zip(off[:-1], off[1:])]
        got = res[key]["info"] if res[key] else []
        assert got == want


def test_refclass_driver_finds_planted_clones():
    """End-to-end recursion on synthetic clones (labels from the planted clone tree, retest hooks on
    the device): the planted deletion/amplification segments come out and clones get split."""
    from honeybadger_b200 import honeybadger as hb
    G, N = 2200, 90
    x, seg, clone = synth.expression(G, N, seed=11)
    x = x / 2.5  # per-cell noise sd ~1 so the pooled chains see the signal at dev/2.5
    chrom = np.empty(G, dtype=object)
    for s in range(22):
        chrom[seg[s]:seg[s + 1]] = f"chr{s + 1}"
    gnames = [f"g{i}" for i in range(G)]
    cells = [f"c{i}" for i in range(N)]
    genes = hb.GRanges(chrom, np.arange(G) * 10.0, np.arange(G) * 10.0 + 5, gnames)
    obj = hb.HoneyBADGER("synthetic")
    obj.gexp_norm = pd.DataFrame(x, index=gnames, columns=cells)
    obj.genes = genes
    obj.setGexpDev(dev=synth.DEV / 2.5)
    obj.calcGexpCnvBoundaries(init=True, min_num_genes=10)
    assert obj.bound_genes_final, "no CNV found"
    found = {g for lst in obj.bound_genes_final for g in lst[0]}
    planted = set(gnames[seg[6]:seg[7]]) | set(gnames[seg[9]:seg[10]]) | set(gnames[seg[12]:seg[13]])
    assert len(found & planted) > 0.5 * len(found)
    assert obj.pred_genes_r.values.sum() > 0


def test_pool_mean_relay_equals_single_pass(oracle):
    """The accumulator relay (cells split into shards handed from 'rank' to 'rank') gives the same
    bits as the one-GPU mean and as R's long-double mean, for uneven shards and an empty shard."""
    import torch
    from honeybadger_b200 import _lib, dist as hd, hmm
    rng = np.random.default_rng(5)
    G, N, K = 700, 211, 7
    x = rng.standard_normal((G, N)) * 2.5
    member = (rng.random((K, N)) < 0.4).astype(np.uint8)
    member[0] = 1
    member[:, 0] = 1            # no empty group
    sizes = member.sum(1).astype(np.int32)
    xd = hmm.to_rows(torch.from_numpy(x).cuda())
    md = torch.from_numpy(member).cuda()
    full = torch.empty((K, G), dtype=torch.float64, device="cuda")
    lib = _lib.load()
    _lib.check(lib.hb_pool_mean_dev(xd.data_ptr(), xd.stride(0), G, N, md.data_ptr(), K, full.data_ptr(), None))
    cuts = [0, 50, 50, 131, N]   # includes an empty shard
    shards = [hmm.to_rows(xd[:, a:b].contiguous()) if b > a else xd[:, a:a] for a, b in zip(cuts[:-1], cuts[1:])]
    mems = [member[:, a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    keep = [i for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])) if b > a]
    relay = hd.pooled_mean_relay_shards([shards[i] for i in keep], [mems[i] for i in keep], sizes)
    assert torch.equal(relay, full)
    for k in (0, 3, 6):
        assert (relay[k].cpu().numpy() == oracle.pool_mean(x, np.nonzero(member[k])[0])).all()


def test_retest_gexp_closed_form_vs_oracle(oracle):
    """hb_retest_gexp_dev (exact marginal of expressionModel.bug) vs the numpy restatement, with and
    without per-cell precisions, on a region carried by a third of the cells."""
    from honeybadger_b200 import honeybadger as hb
    rng = np.random.default_rng(17)
    n, N, m = 60, 500, 0.8
    x = rng.standard_normal((n, N)) * 1.1
    x[:, : N // 3] += m
    for sigma0 in (None, rng.uniform(0.5, 1.5, N)):
        pa, pd_ = hb.retest_gexp_closed_form(x, m, sigma0)
        ra, rd = oracle.retest_gexp(x, m, sigma0)
        assert np.abs(pa - ra).max() < 1e-12 and np.abs(pd_ - rd).max() < 1e-12
        assert pa[: N // 3].mean() > 0.9 and pa[N // 3:].mean() < 0.1 and pd_.max() < 1e-6
    # deletion direction, weak evidence: probabilities stay proper
    y = rng.standard_normal((5, 40)) * 2.0
    y[:, :10] -= 0.4
    pa, pd_ = hb.retest_gexp_closed_form(y, 0.4)
    ra, rd = oracle.retest_gexp(y, 0.4)
    assert np.abs(pa - ra).max() < 1e-12 and np.abs(pd_ - rd).max() < 1e-12
    assert ((pa + pd_) <= 1 + 1e-12).all() and (pa >= 0).all() and (pd_ >= 0).all()


def test_retest_allele_lr_vs_numpy():
    from honeybadger_b200 import honeybadger as hb
    rng = np.random.default_rng(19)
    n = np.where(rng.random((80, 300)) < 0.3, rng.integers(1, 9, (80, 300)), 0)
    p = np.full(300, 0.45)
    p[:90] = 0.1
    r = rng.binomial(n, p)
    got = hb.retest_allele_lr(r, n, 0.1, 0.45)
    llr = (r * np.log(0.1 / 0.45) + (n - r) * np.log(0.9 / 0.55)).sum(0)
    ref = 1 / (1 + np.exp(-llr))
    assert np.abs(got - ref).max() < 1e-12
    assert got[:90].mean() > 0.9 and got[90:].mean() < 0.1


def test_device_grouping_matches_scipy_ward():
    """runmean + dist + Ward.D2 + cutree on the device vs the host form (numpy runmean, scipy ward):
    identical memberships for k = 1..6 on clustered data; dist's NA rule vs a numpy restatement."""
    import ctypes as C
    import torch
    from honeybadger_b200 import _lib, hmm
    from honeybadger_b200 import honeybadger as hb
    rng = np.random.default_rng(41)
    G, N = 900, 157
    clone = rng.choice(4, size=N, p=[0.4, 0.3, 0.2, 0.1])
    x = rng.standard_normal((G, N)) * 1.5
    for c, (a, b, s) in enumerate([(100, 300, 1.2), (400, 650, -1.0), (700, 850, 1.5), (0, 0, 0)]):
        x[a:b, clone == c] += s
    xd = hmm.to_rows(torch.from_numpy(x).cuda())
    heights = [1, 2, 3, 4, 5, 6]
    got = hb.ward_groups_dev(xd, heights, 101)
    from oracle import oracle as O
    ref = O.ward_groups(O.runmean(x, 101), heights)
    for g, r in zip(got, ref):
        assert (g == r).all()
    # the four planted clones are what k = 4 finds
    lab4 = got[3]
    assert len({(a, b) for a, b in zip(lab4, clone)}) == 4
    # runmean on the device == the numpy restatement (NaN skipping, shrinking ends)
    xn = x.copy()
    xn[rng.random(x.shape) < 0.05] = np.nan
    xnd = hmm.to_rows(torch.from_numpy(xn).cuda())
    S = hmm.alloc_rows(G, N, torch.float64, xnd.device)
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.hb_runmean_dev(xnd.data_ptr(), xnd.stride(0), G, N, 31, S.data_ptr(), S.stride(0), st))
    assert np.allclose(S.cpu().numpy(), O.runmean(xn, 31), rtol=1e-12, atol=1e-12, equal_nan=True)
    # dist with R's NA rule: drop unusable pairs, scale by G / pairs used, nothing usable -> 0
    Sn = S.cpu().numpy().copy()
    Sn[:, 5] = np.nan
    Sn[::2, 7] = np.nan
    Sd = hmm.to_rows(torch.from_numpy(Sn).cuda())
    D = torch.empty((N, N), dtype=torch.float64, device="cuda")
    _lib.check(lib.hb_dist_sq_dev(Sd.data_ptr(), Sd.stride(0), G, N, D.data_ptr(), st))
    Dh = D.cpu().numpy()
    for i, j in [(0, 1), (3, 7), (7, 9), (5, 2), (5, 5), (7, 7)]:
        d = Sn[:, i] - Sn[:, j]
        ok = np.isfinite(d)
        want = 0.0 if (i == j or not ok.any()) else (d[ok] ** 2).sum() * G / ok.sum()
        assert abs(Dh[i, j] - want) <= 1e-10 * max(1.0, want) and Dh[i, j] == Dh[j, i]


def test_device_grouping_allele_form():
    import torch
    from honeybadger_b200 import honeybadger as hb
    r, n, _, clone = synth.allele(3000, 90, seed=43)
    got = hb.ward_groups_dev(None, [1, 2, 3], 31, counts=(torch.from_numpy(r).cuda(), torch.from_numpy(n).cuda()))
    assert [len(set(g)) for g in got] == [1, 2, 3] and all(g[0] == 1 for g in got)


def test_host_forms_of_grouping_and_retest(oracle):
    """hb_group_*_host / hb_retest_gexp_host (what the .Call shim binds) vs the device forms."""
    import ctypes as C
    import torch
    from honeybadger_b200 import _lib, hmm
    from honeybadger_b200 import honeybadger as hb
    lib = _lib.load()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rng = np.random.default_rng(47)
    G, N = 700, 93
    clone = rng.choice(3, size=N)
    x = rng.standard_normal((G, N)) * 1.4
    x[100:300, clone == 0] += 1.3
    x[400:600, clone == 1] -= 1.1
    ks = np.array([1, 2, 3, 4], dtype=np.int32)
    lab = np.zeros((4, N), dtype=np.int32)
    xf = np.asfortranarray(x)                         # R's layout
    _lib.check(lib.hb_group_gexp_host(vp(xf), G, N, 101, vp(ks), 4, vp(lab)))
    ref = hb.ward_groups_dev(hmm.to_rows(torch.from_numpy(x).cuda()), list(ks), 101)
    for a, b in zip(lab, ref):
        assert (a == b).all()
    r, n, _, _ = synth.allele(2500, 60, seed=49)
    lab2 = np.zeros((3, 60), dtype=np.int32)
    k3 = np.array([1, 2, 3], dtype=np.int32)
    _lib.check(lib.hb_group_allele_host(vp(np.asfortranarray(r, dtype=np.float64)),
                                        vp(np.asfortranarray(n, dtype=np.float64)), 2500, 60, 31, vp(k3), 3, vp(lab2)))
    ref2 = hb.ward_groups_dev(None, [1, 2, 3], 31, counts=(torch.from_numpy(r).cuda(), torch.from_numpy(n).cuda()))
    for a, b in zip(lab2, ref2):
        assert (a == b).all()
    rows = np.asfortranarray(x[100:160])
    sig = rng.uniform(0.5, 1.5, N)
    pa, pd_, mu = np.zeros(N), np.zeros(N), np.zeros(N)
    _lib.check(lib.hb_retest_gexp_host(vp(rows), 60, N, 1.3, vp(sig), vp(pa), vp(pd_), vp(mu)))
    ra, rd = oracle.retest_gexp(x[100:160], 1.3, sig)
    assert np.abs(pa - ra).max() < 1e-12 and np.abs(pd_ - rd).max() < 1e-12
    assert np.allclose(mu, x[100:160].mean(0), rtol=1e-13, atol=1e-13)


def test_retest_allele_model_vs_oracle(oracle):
    """hb_retest_allele_model_dev (per-cell marginal of inst/bug/snpModel.bug) vs the numpy / scipy.quad
    restatement: with and without a gene factor, small and large coverage, decisive and ambiguous bulk."""
    from honeybadger_b200 import honeybadger as hb
    rng = np.random.default_rng(23)
    m, N = 40, 64
    n = (rng.random((m, N)) < 0.35) * rng.integers(1, 30, (m, N))
    n[3, 5], n[7, 9] = 400, 1245  # sharply peaked integrands (bundled maximum coverage)
    p = np.where(np.arange(N) < N // 3, 0.1, 0.45)
    r = rng.binomial(n, p[None, :])
    l, nb = (r > 0).sum(1).astype(float), (n > 0).sum(1).astype(float)
    nb[::5] = 2 * l[::5]          # bulk exactly balanced: q = 1/2
    for gene in (None, np.arange(m) // 4, rng.integers(0, 6, m)):
        for pe, mono in ((0.1, 0.7), (0.01, 0.3)):
            got = hb.retest_allele_model(r, n, l, nb, gene=gene, pe=pe, mono=mono)
            want = oracle.retest_allele_model(r, n, l, nb, gene=gene, pseudo=pe, mono=mono)
            assert np.abs(got - want).max() < 1e-9
    # lesser-allele counts as setAlleleMats leaves them (l.maf well below n.bulk / 2): the planted cells,
    # whose r / n sits at pseudo, come out as deleted
    got = hb.retest_allele_model(r, n, np.floor(0.2 * nb), nb, pe=0.1)
    assert got[: N // 3].mean() > 0.9 and got[N // 3:].mean() < 0.1
    # host form in R's layout, gene codes in arbitrary order
    import ctypes as C
    from honeybadger_b200 import _lib
    gene = rng.integers(0, 6, m).astype(np.int32)
    rf, nf = np.asfortranarray(r.astype(np.int32)), np.asfortranarray(n.astype(np.int32))
    out = np.zeros(N)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.load().hb_retest_allele_model_host(vp(rf), vp(nf), m, N, vp(l), vp(nb), vp(gene), 0.1, 0.7, vp(out)))
    assert np.abs(out - oracle.retest_allele_model(r, n, l, nb, gene=gene)).max() < 1e-9


def test_refclass_allele_recursion_uses_the_model_retest():
    """calcAlleleCnvBoundaries(init=TRUE) end to end with the default (snpModel marginal) retest on planted
    deletions: the deleted clone's cells are split off and the region is recorded."""
    import pandas as pd
    from honeybadger_b200 import honeybadger as hb
    rng = np.random.default_rng(31)
    G, N = 3000, 120
    chrom = np.repeat(np.arange(1, 11), G // 10)
    names = [f"{c}:{1000 + 50 * i}" for i, c in enumerate(chrom)]
    cells = [f"c{i}" for i in range(N)]
    # the pooled chain sees presence / absence (rowSums(r.maf > 0) of rowSums(n.sc > 0)): shallow coverage,
    # lesser allele seen in ~half of the covered cells when neutral, almost never in the deleted clone
    n = (rng.random((G, N)) < 0.3) * rng.integers(1, 3, (G, N))
    p = np.full((G, N), 0.45)
    dele = slice(900, 1200)          # chromosome 4
    p[dele, :50] = 0.02              # clone of 50 cells
    r = rng.binomial(n, p)
    obj = hb.HoneyBADGER("allele")
    obj.setAlleleMats(pd.DataFrame(r, index=names, columns=cells), pd.DataFrame(n, index=names, columns=cells),
                      filter=False, verbose=False)
    obj.calcAlleleCnvBoundaries(init=True)
    assert obj.bound_snps_final, "no deletion found"
    found = [s for lst in obj.bound_snps_final for s in lst[0]]
    inside = [s for s in found if s in set(names[dele])]
    assert len(inside) > 0.8 * len(found) and len(inside) > 100
    hit = obj.pred_snps_r.loc[inside[0]].values > 0
    assert hit[:50].mean() > 0.9 and hit[50:].mean() < 0.05


def test_retest_comb_vs_oracle(oracle):
    """hb_retest_comb_dev (marginal of inst/bug/combinedModel.bug: expression retest with the allele model's
    per-cell log-odds on the deletion branch) vs the numpy restatement; alleles rescue a deletion that the
    expression data alone leave undecided."""
    from honeybadger_b200 import honeybadger as hb
    rng = np.random.default_rng(41)
    J, m, N, mag = 25, 30, 96, 0.5
    x = rng.standard_normal((J, N)) * 1.5
    x[:, :32] -= mag
    n = (rng.random((m, N)) < 0.4) * rng.integers(1, 6, (m, N))
    p = np.where(np.arange(N) < 32, 0.1, 0.45)
    r = rng.binomial(n, p[None, :])
    nb = (n > 0).sum(1).astype(float)
    l = np.floor(0.2 * nb)
    gene = np.arange(m) // 3
    for sigma0 in (None, rng.uniform(0.3, 0.8, N)):
        pa, pd_ = hb.retest_comb_closed_form(x, mag, r, n, l, nb, sigma0=sigma0, gene=gene)
        ra, rd = oracle.retest_comb(x, mag, r, n, l, nb, sigma0=sigma0, gene=gene)
        assert np.abs(pa - ra).max() < 1e-9 and np.abs(pd_ - rd).max() < 1e-9
    ea, ed = hb.retest_gexp_closed_form(x, mag, sigma0)
    assert pd_[:32].mean() > 0.9 and pd_[32:].mean() < 0.05 and pd_[:32].mean() > ed[:32].mean()


def test_stateless_allele_function_reproduces_the_vignette_figure(golden_inputs):
    """The whole deterministic allele path on the device — setAlleleMats, runmean / dist / Ward.D2 / cutree,
    pooled counts, binomial Viterbi, vote, runs, trim, range() — against the reference's own rendered answer:
    the panels of docs/figure/Integrating/unnamed-chunk-6-1.png are chr1 chr3 chr9 chr10 chr12 chr13 chr14
    chr15 chr19 (region@seqnames@values, R/HoneyBADGER_allele.R:248)."""
    import pandas as pd
    from honeybadger_b200 import honeybadger as hb
    r, cov = golden_inputs["r"].astype(np.int32), golden_inputs["cov"].astype(np.int32)
    names = [f"{c}:{p}" for c, p in zip(golden_inputs["chrom"], golden_inputs["pos"])]
    cells = [f"c{i}" for i in range(r.shape[1])]
    am = hb.set_allele_mats_dev(r, cov, het_deviance_threshold=0.05)
    assert am["keep"].size == 4550 and am["n_het"] == 5552  # docs/Integrating.md:27
    kept = [names[i] for i in am["keep"]]
    res = hb.calcAlleleCnvBoundaries(pd.DataFrame(am["r_maf"].cpu().numpy(), index=kept, columns=cells),
                                     pd.DataFrame(am["n_sc"].cpu().numpy(), index=kept, columns=cells))
    chroms = []
    for seqname, _, _ in res["regions"]:
        c = int(str(seqname).replace("chr", ""))
        if not chroms or chroms[-1] != c:
            chroms.append(c)
    assert chroms == list(golden_inputs["figure_region_chroms"])


def test_refclass_allele_run_reproduces_the_vignette_regions(golden_inputs):
    """docs/Getting_Started.md:287-293 prints the regions the reference's RefClass run finds on the bundled
    data: chr10 [3180316, 126670356], chr13 [21949267, 114175032], chr14 [20216560, 73465003],
    chr19 [20727463, 48598823].  The mirror's whole device flow (setAlleleMats 0.1 -> Ward grouping -> pooled
    binomial Viterbi -> vote / runs / trim -> snpModel-marginal retest -> split -> recursion) finds the same
    four; chr10 / chr13 / chr14 to the base pair.  The chr19 run is found in a sub-population whose cell set
    the reference draws by MCMC, so its far end may sit a SNP or two away."""
    import json
    import os
    import pandas as pd
    from honeybadger_b200 import honeybadger as hb
    want = {c: (a, b) for c, a, b in json.load(open(os.path.join(
        os.path.dirname(__file__), "golden", "known_answers.json")))["getting_started_allele_regions"]}
    r, cov = golden_inputs["r"].astype(np.int32), golden_inputs["cov"].astype(np.int32)
    names = [f"{c}:{p}" for c, p in zip(golden_inputs["chrom"], golden_inputs["pos"])]
    cells = [f"c{i}" for i in range(r.shape[1])]
    obj = hb.HoneyBADGER("MGH31")
    obj.setAlleleMats(pd.DataFrame(r, index=names, columns=cells), pd.DataFrame(cov, index=names, columns=cells),
                      het_deviance_threshold=0.1, verbose=False)
    obj.calcAlleleCnvBoundaries(init=True)
    found = [s for lst in obj.bound_snps_final for s in lst[0]]
    got = {c: (int(a), int(b)) for c, a, b in obj.snps[found].range()}
    assert set(got) == set(want)
    for c in ("chr10", "chr13", "chr14"):
        assert got[c] == tuple(want[c])
    assert got["chr19"][0] == want["chr19"][0] and abs(got["chr19"][1] - want["chr19"][1]) < 100_000


def test_large_count_uploads_convert_on_the_device():
    """_to_dev_i32 on matrices past the slab threshold: C / Fortran order (a DataFrame's .values), integer and
    double inputs, half-even rounding — identical to the small-matrix host path."""
    from honeybadger_b200 import honeybadger as hb
    rng = np.random.default_rng(8)
    a = rng.integers(0, 40, size=(4200, 4100)).astype(np.int16)
    want = a.astype(np.int32)
    for arr in (a, np.asfortranarray(a), a.astype(np.float64), np.asfortranarray(a.astype(np.float64))):
        assert arr.size >= 1 << 24
        assert np.array_equal(hb._to_dev_i32(arr).cpu().numpy(), want)
    f = a.astype(np.float64) + 0.5                       # ties: R's rint / numpy rint round half to even
    assert np.array_equal(hb._to_dev_i32(np.asfortranarray(f)).cpu().numpy(), np.rint(f).astype(np.int32))
    df = pd.DataFrame(a)
    assert np.array_equal(hb._to_dev_i32(df.values).cpu().numpy(), want)
    x = rng.standard_normal((4200, 4100))
    for arr in (x, np.asfortranarray(x), np.asfortranarray(x.astype(np.float32))):
        assert np.array_equal(hb._to_dev_f64(arr).cpu().numpy(), arr.astype(np.float64))


@pytest.mark.parametrize("N,K,ld_extra", [(75, 6, 0), (20037, 32, 0), (20037, 40, 0), (1000, 3, 1), (333, 33, 3),
                                          (60000, 32, 0)])
def test_pool_count_kernels_against_numpy(N, K, ld_extra):
    """rowSums(m[, grp] > 0) for K groups (R/HoneyBADGER.R:1404-1405): the 16-byte-load stream (aligned rows,
    membership words in shared memory, > 48 KB of them for many groups x cells), its ragged last slab, the
    passes of 32 groups, and the per-group ballot kernel that serves unaligned rows and oversized tables —
    exact integers either way."""
    import torch
    from honeybadger_b200 import _lib
    rng = np.random.default_rng(N + K)
    G = 97
    ld = (N + 3) // 4 * 4 + ld_extra
    v = np.full((G, ld), 7, np.int32)                 # padding holds positives: must not be counted
    v[:, :N] = rng.integers(-1, 3, size=(G, N))
    member = (rng.random((K, N)) < 0.4).astype(np.uint8)
    member[:, 0] = 1
    vd = torch.from_numpy(v).cuda()
    out = torch.zeros((K, G), dtype=torch.int32, device="cuda")
    lib = _lib.load()
    md = torch.from_numpy(member).cuda()
    _lib.check(lib.hb_pool_count_pos_dev(vd.data_ptr(), ld, G, N, md.data_ptr(), K, out.data_ptr(), None))
    torch.cuda.synchronize()
    want = (member.astype(np.int64) @ (v[:, :N] > 0).T.astype(np.int64)).astype(np.int32)
    assert np.array_equal(out.cpu().numpy(), want)


def test_pool_count_on_column_slice_views():
    """A rank's shard is a column slice r[:, lo:hi] of the parent matrix: ld is the parent's stride and the
    tail of the last row beyond N is NOT part of the allocation — the 16-byte stream must stay inside N
    (run under compute-sanitizer memcheck in profiles/r2_sanitizer.txt)."""
    import torch
    from honeybadger_b200 import _lib
    rng = np.random.default_rng(8)
    G, Nfull, K = 61, 1024, 5
    v = rng.integers(-1, 3, size=(G, Nfull)).astype(np.int32)
    vd = torch.from_numpy(v).cuda()
    lib = _lib.load()
    for lo, hi in [(0, 1024), (512, 1024), (768, 1021), (4, 1024), (1000, 1024)]:
        N = hi - lo
        member = (rng.random((K, N)) < 0.5).astype(np.uint8)
        member[:, 0] = 1
        md = torch.from_numpy(member).cuda()
        out = torch.zeros((K, G), dtype=torch.int32, device="cuda")
        view = vd[:, lo:hi]
        _lib.check(lib.hb_pool_count_pos_dev(view.data_ptr(), vd.stride(0), G, N, md.data_ptr(), K,
                                             out.data_ptr(), None))
        torch.cuda.synchronize()
        want = (member.astype(np.int64) @ (v[:, lo:hi] > 0).T.astype(np.int64)).astype(np.int32)
        assert np.array_equal(out.cpu().numpy(), want), (lo, hi)
