"""CPU: the oracle (oracle/hb_oracle.c) against the committed golden vectors and its cross-checks."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

from oracle import pyref as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def unrle(starts, vals, n):
    y = np.zeros(n, dtype=np.int32)
    ends = list(starts[1:]) + [n]
    for a, b, v in zip(starts, ends, vals):
        y[a:b] = v
    return y


def test_densities_against_mpmath_golden(oracle):
    d = json.load(open(os.path.join(GOLDEN, "densities.json")))
    for x, n, p, ref in d["dbinom_log"]:
        got = oracle.dbinom(x, n, p, True)
        assert got == P.dbinom_log(x, n, p)  # C restatement == Python restatement, bit for bit
        assert abs(got - ref) <= 4e-15 * max(1.0, abs(ref)), (x, n, p, got, ref)
    for x, mu, sd, ref in d["dnorm_log"]:
        got = oracle.dnorm_log(x, mu, sd)
        assert got == P.dnorm_log(x, mu, sd)
        assert abs(got - ref) <= 2e-15 * max(1.0, abs(ref))


def test_zero_coverage_emission_is_zero_for_both_states(oracle):
    # dbinom(0, 0, p, log=TRUE) == 0 — the source of the structural ties
    assert oracle.dbinom(0, 0, 0.1, True) == 0.0 and oracle.dbinom(0, 0, 0.45, True) == 0.0
    assert oracle.dbinom(3, 2, 0.45, True) == -np.inf


def test_r_mean_sd_match_longdouble_python(oracle):
    rng = np.random.default_rng(3)
    for n in (2, 3, 17, 1000, 6082):
        x = rng.normal(size=n) * 3 + rng.choice([0, 5])
        assert oracle.r_mean(x) == P.r_mean(x)
        assert oracle.r_sd(x) == P.r_sd(x)
        assert abs(oracle.r_sd(x) - np.std(x, ddof=1)) < 1e-12


def test_known_answers_from_vignettes(golden_allele):
    # docs/Integrating.md:27 prints 5552 (4550 kept); docs/Getting_Started.md:230 prints 5359 (4357 kept)
    assert golden_allele["known_counts"].tolist() == [[5552, 4550], [5359, 4357]]
    assert golden_allele["r_maf"].shape == (4357, 75)
    ka = json.load(open(os.path.join(GOLDEN, "known_answers.json")))["allele_all_cell_state1_runs"]
    # the all-cell pooled chain calls chr1 (9 SNPs) and chr13..chr14 (54 SNPs): the vignette's chr13/14 deletions
    assert [r[2] for r in ka] == [9, 54] and ka[1][0].startswith("13:") and ka[1][1].startswith("14:")


def test_allele_golden_paths(oracle, golden_allele):
    g = golden_allele
    r_maf, n_sc = g["r_maf"].astype(np.float64), g["n_sc"].astype(np.float64)
    Pi = oracle.sym_Pi(2, 1e-6)
    for k, (h, gi) in enumerate(g["chain_hg"]):
        labels = g["labels"][h - 1]
        cells = P.group_members(labels)[gi - 1][1]
        mafl, sizel = oracle.pool_count_pos(r_maf, cells), oracle.pool_count_pos(n_sc, cells)
        assert (mafl == g["mafl"][k]).all() and (sizel == g["sizel"][k]).all()
        y = oracle.viterbi_binom(mafl, sizel, [0.1, 0.45], Pi, [0.0, 1.0])
        a, b = g["path_off"][k], g["path_off"][k + 1]
        assert (y == unrle(g["path_starts"][a:b], g["path_vals"][a:b], y.size)).all()


def test_allele_golden_boundaries(golden_allele):
    g = golden_allele
    res = P.calc_allele_cnv_boundaries_fn(g["r_maf"].astype(float), g["n_sc"].astype(float), g["labels"])
    info = res["info"] or []
    assert np.concatenate(info).tolist() == g["info_idx"].tolist() if info else g["info_idx"].size == 0
    assert np.cumsum([0] + [a.size for a in info]).tolist() == g["info_off"].tolist()


def test_gexp_golden_paths_and_posteriors(oracle, golden_gexp):
    g = golden_gexp
    dev = float(g["dev"])
    Pi = oracle.sym_Pi(3, 1e-6)
    x_all = g["gexp_norm_f32"].astype(np.float64)
    k = 0
    for h in range(5):
        for _, cells in P.group_members(g["labels"][h]):
            if cells.size < 2:
                continue
            x = oracle.pool_mean(x_all, cells)
            assert (x == g["pooled_x"][k]).all()
            sd = oracle.r_sd(x)
            assert sd == g["pooled_sd"][k]
            y = oracle.viterbi_norm(x, [-dev, 0, dev], [sd] * 3, Pi, [0, 1, 0])
            a, b = g["path_off"][k], g["path_off"][k + 1]
            assert (y == unrle(g["path_starts"][a:b], g["path_vals"][a:b], y.size)).all()
            gam, ll = oracle.forwardback_norm(x, [-dev, 0, dev], [sd] * 3, Pi, [0, 1, 0])
            assert np.abs(gam[g["gamma_samp_idx"]] - g["gamma_samp"][k]).max() < 1e-12
            assert abs(ll - g["ll"][k]) <= 1e-12 * abs(ll)
            k += 1
    assert k == g["pooled_x"].shape[0] == 15


def test_viterbi_c_vs_python_with_structural_ties(oracle):
    """Small groups: zero-coverage SNPs give (0, 0) emissions and exact ties."""
    rng = np.random.default_rng(5)
    Pi = oracle.sym_Pi(2, 1e-6)
    for trial in range(6):
        T = 700
        n = np.where(rng.random(T) < 0.25, rng.integers(1, 4, T), 0)
        p = np.full(T, 0.45)
        p[200:420] = 0.1
        x = rng.binomial(n, p)
        y1 = oracle.viterbi_binom(x, n, [0.1, 0.45], Pi, [0.0, 1.0])
        y2 = P.viterbi_binom(x, n, [0.1, 0.45], Pi, [0.0, 1.0])
        assert (y1 == y2).all()


def test_forwardback_matches_exact_arithmetic(oracle):
    rng = np.random.default_rng(8)
    x = rng.normal(size=120)
    x[40:70] += 1.3
    Pi = oracle.sym_Pi(3, 1e-3)
    sd = oracle.r_sd(x)
    g, _ = oracle.forwardback_norm(x, [-1, 0, 1], [sd] * 3, Pi, [0, 1, 0])
    logb = [[P.dnorm_log(float(v), m, sd) for m in (-1, 0, 1)] for v in x]
    gm = P.mp_posteriors(logb, Pi, [0, 1, 0])
    assert np.abs(g - gm).max() < 1e-12


def test_underflow_error_like_hiddenmarkov(oracle):
    with pytest.raises(FloatingPointError, match="Problems With Underflow"):
        oracle.viterbi_norm([0.3], [-1, 0, 1], [1.0] * 3, oracle.sym_Pi(3, 1e-6), [0, 1, 0])


def test_betabinom_logpmf_against_mpmath_and_its_binomial_limit(oracle):
    """hbo_dbetabinom_log (binomial + correction; an extension — the reference is binomial,
    R/HoneyBADGER.R:1409) against the exact log-gamma formulation in mpmath, over dispersion from 1e-14 to
    0.3 and counts up to the bundled maximum 1245; rho = 0 IS dbinom, and rho -> 0 converges to it without the
    cancellation a lgamma-of-(p/rho) form suffers (error stays ~1e-15 relative at rho = 1e-12)."""
    from oracle import pyref
    rng = np.random.default_rng(0)
    cases = [(0, 0), (0, 5), (5, 5), (1, 2), (3, 17), (40, 100), (12, 511), (600, 1245), (0, 1245), (1245, 1245)]
    cases += [(int(rng.integers(0, n + 1)), int(n)) for n in rng.integers(1, 400, size=40)]
    worst = 0.0
    for rho in (1e-14, 1e-12, 1e-9, 1e-6, 1e-3, 0.05, 0.3):
        for p in (0.1, 0.45):
            for x, n in cases:
                got = oracle.dbetabinom_log(x, n, p, rho)
                want = float(pyref.betabinom_logpmf_exact(x, n, p, rho))
                # the value is binomial + correction: both can be hundreds of nats apart and cancel, so the
                # error scales with the binomial term, not with the (possibly small) result
                scale = max(1.0, abs(want), abs(oracle.dbinom(x, n, p, log=True)))
                err = abs(got - want) / scale
                worst = max(worst, err)
                assert err < 2e-14, (x, n, p, rho, got, want)
    # exact reduction at rho = 0, and a correction that vanishes linearly
    for x, n in cases:
        b = oracle.dbinom(x, n, 0.1, log=True)
        assert oracle.dbetabinom_log(x, n, 0.1, 0.0) == b
        d12, d6 = oracle.dbetabinom_log(x, n, 0.1, 1e-12) - b, oracle.dbetabinom_log(x, n, 0.1, 1e-6) - b
        assert abs(d12) <= 1e-12 * (n * n + 1) * 6 and abs(d6) <= 1e-6 * (n * n + 1) * 6
    # zero coverage carries no information under any dispersion (the structural ties of SURVEY 7.2-1 survive)
    assert oracle.dbetabinom_log(0, 0, 0.1, 0.05) == 0.0 and oracle.dbetabinom_log(0, 0, 0.45, 0.05) == 0.0


def test_betabinom_table_of_the_product_equals_the_oracle_bit_for_bit(oracle, emul):
    """The product's emission table builder (hb_host_math.h: BetaBinomCorr prefix sums) against the oracle's
    direct loops: the same long-double sums in the same order, so identical doubles for every (x, n)."""
    import ctypes as C
    emul.emul_betabinom_logpmf.restype = C.c_double
    emul.emul_betabinom_logpmf.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
    for rho in (1e-12, 1e-6, 0.05, 0.3):
        for p in (0.1, 0.45):
            for n in list(range(0, 40)) + [97, 200, 511, 1245]:
                for x in sorted({0, 1, n // 3, n // 2, max(0, n - 1), n}):
                    a = emul.emul_betabinom_logpmf(x, n, p, rho, 1245)
                    b = oracle.dbetabinom_log(x, n, p, rho)
                    assert a == b, (x, n, p, rho, a, b)


def test_betabinom_chains_reduce_to_binomial_chains(oracle):
    """Oracle level: Viterbi paths and posteriors of beta-binomial chains at rho = 1e-14 / 1e-12 against the
    binomial chains on the same counts (zero-coverage ties included): states identical, posteriors within 1e-9;
    rho = 0.05 gives different posteriors (the extension does something)."""
    from honeybadger_b200 import synth
    r, n, seg, _ = synth.allele(6000, 48, seed=3)
    Pi, delta, prob = oracle.sym_Pi(2, 1e-6), [0.0, 1.0], [0.1, 0.45]
    y0, g0, _, rc0 = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta)
    assert rc0 == 0
    for rho in (1e-14, 1e-12):
        y, g, _, rc = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta, rho=rho)
        assert rc == 0 and (y == y0).all(), f"rho={rho}: {(y != y0).sum()} states differ"
        assert np.abs(g - g0).max() < 1e-9
    y5, g5, _, rc = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta, rho=0.05)
    assert rc == 0 and np.abs(g5 - g0).max() > 1e-3
