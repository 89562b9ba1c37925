"""Few long chains: the gene-axis parallel Viterbi (csrc/hb_longchain.cuh) against the oracle.

CPU part: the block bodies compiled for the host (tests/emul) walk the phases serially; the states must
equal `hbo_viterbi_logb` (the restatement of HiddenMarkov::Viterbi) on every input, whatever the integer
prediction does — binade crossings, half-way addends, exact ties, -Inf emissions, zero transition
probabilities, positive and sign-changing nu.  GPU part: the same through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st_


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def sym_Pi(S, t):
    P = np.full((S, S), t)
    np.fill_diagonal(P, 1 - (S - 1) * t)
    return P


def ref_paths(oracle, lb, Pi, delta):
    O = oracle.lib()
    K, T, S = lb.shape
    y = np.zeros((K, T), np.int32)
    rcs = []
    for c in range(K):
        yc = np.zeros(T, np.int32)
        rcs.append(O.hbo_viterbi_logb(S, C.c_int64(T), ptr(np.ascontiguousarray(lb[c])), ptr(Pi), ptr(delta),
                                      ptr(yc), None))
        y[c] = yc
    return rcs, y


def run_emul(emul, lb, Pi, delta, L=0):
    K, T, S = lb.shape
    y = np.zeros((K, T), np.int32)
    st = np.zeros((K, 4), np.int32)
    emul.emul_longchain.restype = C.c_int
    rc = emul.emul_longchain(S, C.c_int64(T), K, ptr(lb), ptr(np.ascontiguousarray(Pi, dtype=np.float64)),
                             ptr(np.ascontiguousarray(delta, dtype=np.float64)), L, ptr(y), ptr(st))
    return rc, y, st


def binom_lb(rng, K, T, depth=5000, cover=0.108):
    from scipy.stats import binom
    n = rng.binomial(depth, cover, size=(K, T))
    dele = np.zeros((K, T), bool)
    for c in range(K):
        for _ in range(5):
            a = rng.integers(0, max(T - 1, 1))
            dele[c, a:a + rng.integers(1, max(T // 10, 2))] = True
    x = rng.binomial(n, np.where(dele, 0.1, 0.45))
    return np.ascontiguousarray(np.stack([binom.logpmf(x, n, 0.1), binom.logpmf(x, n, 0.45)], -1))


def norm_lb(rng, K, T, sd=0.25, dev=1.0157):
    x = rng.normal(0, sd, size=(K, T))
    for c in range(K):
        for _ in range(4):
            a = rng.integers(0, max(T - 1, 1))
            x[c, a:a + rng.integers(1, max(T // 10, 2))] += rng.choice([-1, 1]) * dev
    mean = np.array([-dev, 0, dev])
    return np.ascontiguousarray(-0.5 * ((x[..., None] - mean) / sd) ** 2 - np.log(sd) - 0.9189385332046727)


@pytest.mark.parametrize("S,T,K,L", [(2, 20000, 3, 0), (3, 20000, 3, 0), (2, 200000, 1, 0), (3, 6082, 4, 0),
                                     (2, 2, 2, 0), (3, 17, 3, 0), (2, 130, 2, 16), (2, 129, 2, 128),
                                     (3, 4097, 2, 64), (2, 5000, 2, 1)])
def test_pooled_shaped_chains_equal_the_oracle(emul, oracle, S, T, K, L):
    rng = np.random.default_rng(S * 1000 + T)
    lb = binom_lb(rng, K, T) if S == 2 else norm_lb(rng, K, T)
    Pi, delta = sym_Pi(S, 1e-6), np.array([0., 1.] if S == 2 else [0., 1., 0.])
    rc, y, st = run_emul(emul, lb, Pi, delta, L)
    rcs, yr = ref_paths(oracle, lb, Pi, delta)
    assert rc == 0 and not any(rcs)
    assert np.array_equal(y, yr)
    assert (st[:, 1] > 0).all()


def test_prediction_carries_deep_binomial_chains(emul, oracle):
    """Pooled allele shape: after the first binades the integer prediction must hold (no repair rounds,
    no sequential finish, only a few blocks run literally by the sequential walker)."""
    rng = np.random.default_rng(7)
    lb = binom_lb(rng, 2, 100000)
    Pi, delta = sym_Pi(2, 1e-6), np.array([0., 1.])
    rc, y, st = run_emul(emul, lb, Pi, delta)
    _, yr = ref_paths(oracle, lb, Pi, delta)
    assert np.array_equal(y, yr)
    assert (st[:, 1] == 1).all() and (st[:, 2] == 0).all()
    assert (st[:, 3] < st[:, 0] // 10).all(), st


def test_halfway_addends_and_exact_ties(emul, oracle):
    """Emissions on a coarse binary grid: c / q is a half-integer at most positions of some binades
    (round-half-even depends on the parity of nu), and zero-evidence positions tie exactly."""
    rng = np.random.default_rng(11)
    for S, grid in [(2, 2.0 ** -40), (2, 2.0 ** -36), (3, 2.0 ** -38), (3, 2.0 ** -44)]:
        T, K = 6000, 3
        lb = -np.round(rng.gamma(2.0, 3.0, size=(K, T, S)) / grid) * grid - grid / 2
        lb[:, ::3, :] = 0.0                                   # zero coverage: exact ties
        lb = np.ascontiguousarray(lb)
        Pi, delta = sym_Pi(S, 1e-6), np.array([0., 1.] if S == 2 else [0., 1., 0.])
        for L in (0, 16, 64):
            rc, y, st = run_emul(emul, lb, Pi, delta, L)
            rcs, yr = ref_paths(oracle, lb, Pi, delta)
            assert rc == 0 and not any(rcs)
            assert np.array_equal(y, yr), (S, grid, L)


def test_minus_inf_emissions_zero_transitions_and_underflow(emul, oracle):
    rng = np.random.default_rng(5)
    T, K = 3000, 2
    lb = binom_lb(rng, K, T, depth=40)
    lb[0, 100:140, 0] = -np.inf                                # impossible under state 1
    lb[1, 2000, 1] = -np.inf
    Pi, delta = sym_Pi(2, 1e-6), np.array([0.3, 0.7])
    rc, y, _ = run_emul(emul, lb, Pi, delta, 32)
    rcs, yr = ref_paths(oracle, lb, Pi, delta)
    assert rc == 0 and not any(rcs) and np.array_equal(y, yr)
    Pi0 = np.array([[1.0, 0.0], [1e-4, 1 - 1e-4]])             # absorbing state: log(0) transitions
    rc, y, _ = run_emul(emul, lb, Pi0, delta, 32)              # chain 1 is driven into the absorbing state:
    rcs, yr = ref_paths(oracle, lb, Pi0, delta)                # its other nu stays -Inf -> underflow, as in R
    assert rc == 1 and rcs[0] == 0 and rcs[1] != 0 and np.array_equal(y[0], yr[0])
    lb1 = lb.copy()
    lb1[1, 2000] = lb1[1, 2000, ::-1]
    rc, y, _ = run_emul(emul, lb1, Pi0, delta, 32)
    rcs, yr = ref_paths(oracle, lb1, Pi0, delta)
    assert rc == 0 and not any(rcs) and np.array_equal(y, yr)
    lb2 = lb.copy()
    lb2[:, -1, :] = -np.inf                                    # HiddenMarkov: "Problems With Underflow"
    rc, _, _ = run_emul(emul, lb2, Pi, delta, 32)
    rcs, _ = ref_paths(oracle, lb2, Pi, delta)
    assert rc == 1 and all(r != 0 for r in rcs)
    rc, _, _ = run_emul(emul, lb[:, :1, :].copy(), Pi, np.array([0., 1.]), 0)   # length 1, zero prior
    assert rc == 1


def test_positive_and_sign_changing_nu(emul, oracle):
    """Gaussian log-densities with a small sd are positive: nu grows upwards, or hovers around zero."""
    rng = np.random.default_rng(3)
    for sd in (0.01, 0.2420, 0.3):
        lb = norm_lb(rng, 3, 8000, sd=sd)
        Pi, delta = sym_Pi(3, 1e-6), np.array([0., 1., 0.])
        rc, y, _ = run_emul(emul, lb, Pi, delta)
        rcs, yr = ref_paths(oracle, lb, Pi, delta)
        assert rc == 0 and not any(rcs) and np.array_equal(y, yr), sd


@settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck))
@given(st_.sampled_from([2, 3]), st_.integers(1, 700), st_.integers(1, 3), st_.sampled_from([0, 1, 2, 7, 16, 128]),
       st_.sampled_from([1e-8, 1e-6, 1e-3, 0.2]), st_.sampled_from([1e-3, 1.0, 40.0, 3000.0]),
       st_.sampled_from([0.0, 2.0 ** -30, 2.0 ** -45]), st_.integers(0, 2 ** 31 - 1))
def test_random_chains(emul, oracle, S, T, K, L, t, scale, grid, seed):
    rng = np.random.default_rng(seed)
    lb = -rng.gamma(1.5, scale, size=(K, T, S))
    if grid:
        lb = np.round(lb / grid) * grid
    if rng.random() < 0.3:
        lb += rng.normal(0, scale, size=(K, 1, 1)) + scale    # positive stretches
    lb = np.ascontiguousarray(lb)
    Pi = sym_Pi(S, t)
    delta = rng.dirichlet(np.ones(S)) if rng.random() < 0.5 else np.eye(S)[S // 2]
    rc, y, _ = run_emul(emul, lb, Pi, delta, L)
    rcs, yr = ref_paths(oracle, lb, Pi, delta)
    if any(rcs):
        assert rc == 1
        return
    assert rc == 0 and np.array_equal(y, yr)


# ------------------------------------------------------------------------------ posteriors (lc_fb_*)
def ref_posteriors(oracle, lb, Pi, delta):
    """HiddenMarkov::forwardback + exp(logalpha + logbeta - LL) on the linear densities exp(logb)."""
    O = oracle.lib()
    K, T, S = lb.shape
    g, ll = np.zeros((K, T, S)), np.zeros(K)
    for c in range(K):
        prob, gc, L = np.ascontiguousarray(np.exp(lb[c])), np.zeros((T, S)), C.c_double(0)
        assert O.hbo_forwardback_prob(S, C.c_int64(T), ptr(prob), ptr(Pi), ptr(delta), None, None, ptr(gc),
                                      C.byref(L)) == 0
        g[c], ll[c] = gc, L.value
    return g, ll


def longdouble_posteriors(lb, Pi, delta):
    """Scaled forward-backward in 80-bit arithmetic (independent of the oracle's log-space roundings)."""
    T, S = lb.shape
    ld = np.longdouble
    P, b = Pi.astype(ld), np.exp((lb - lb.max(1, keepdims=True)).astype(ld))
    al = np.zeros((T, S), ld)
    a = delta.astype(ld) * b[0]
    al[0] = a / a.sum()
    for t in range(1, T):
        a = (al[t - 1] @ P) * b[t]
        al[t] = a / a.sum()
    g = np.zeros((T, S), ld)
    bt = np.ones(S, ld)
    for t in range(T - 1, -1, -1):
        v = al[t] * bt
        g[t] = v / v.sum()
        bt = P @ (b[t] * bt)
        bt = bt / bt.sum()
    return g.astype(np.float64)


def run_emul_fb(emul, lb, Pi, delta, L=0):
    K, T, S = lb.shape
    g, ll = np.zeros((K, T, S)), np.zeros(K)
    emul.emul_longchain_fb.restype = C.c_int
    assert emul.emul_longchain_fb(S, C.c_int64(T), K, ptr(lb), ptr(np.ascontiguousarray(Pi, dtype=np.float64)),
                                  ptr(np.ascontiguousarray(delta, dtype=np.float64)), L, ptr(g), ptr(ll)) == 0
    return g, ll


@pytest.mark.parametrize("S,T,K,L", [(2, 5000, 2, 0), (3, 5000, 2, 0), (3, 1, 2, 0), (2, 2, 2, 0), (3, 700, 2, 16),
                                     (2, 129, 1, 128), (2, 1000, 2, 1), (3, 6082, 3, 0)])
def test_posteriors_match_the_oracle(emul, oracle, S, T, K, L):
    """north_star: forward-backward posteriors within 1e-9 absolute in fp64."""
    rng = np.random.default_rng(S * 31 + T)
    lb = binom_lb(rng, K, T, depth=60) if S == 2 else norm_lb(rng, K, T, sd=1.0)
    Pi, delta = sym_Pi(S, 1e-6), np.array([0.3, 0.7] if S == 2 else [0., 1., 0.])
    g, ll = run_emul_fb(emul, lb, Pi, delta, L)
    gr, llr = ref_posteriors(oracle, lb, Pi, delta)
    assert np.abs(g - gr).max() < 1e-9
    assert np.abs(g.sum(-1) - 1).max() < 1e-14
    assert np.allclose(ll, llr, rtol=1e-12, atol=1e-9)


def test_posteriors_in_r_plane_layout(emul, oracle):
    """gamma written straight into R's [T x K x S] column-major array (what hb_hmm_*_host hands back)."""
    rng = np.random.default_rng(9)
    lb = norm_lb(rng, 3, 2500, sd=1.0)
    Pi, delta = sym_Pi(3, 1e-6), np.array([0., 1., 0.])
    g0, ll0 = run_emul_fb(emul, lb, Pi, delta)
    emul.emul_set_fb_planes(1)
    try:
        g1, ll1 = run_emul_fb(emul, lb, Pi, delta)
    finally:
        emul.emul_set_fb_planes(0)
    planes = g1.reshape(-1).reshape(3, 3, 2500)            # [S][K][T]
    assert np.array_equal(planes.transpose(1, 2, 0), g0) and np.array_equal(ll0, ll1)


def test_posteriors_of_a_long_deep_chain(emul):
    """40 000 positions with pooled-scale evidence (hundreds of nats per position: exp(logb) underflows,
    the reference's linear-density forwardback returns NaN there) against 80-bit arithmetic."""
    rng = np.random.default_rng(2)
    lb = binom_lb(rng, 1, 40000)
    Pi, delta = sym_Pi(2, 1e-6), np.array([0., 1.])
    g, _ = run_emul_fb(emul, lb, Pi, delta)
    assert np.abs(g[0] - longdouble_posteriors(lb[0], Pi, delta)).max() < 1e-10


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("S,T,K,L", [(2, 200000, 6, 0), (3, 20000, 15, 0), (2, 4550, 6, 0), (3, 6082, 15, 16),
                                     (2, 1, 2, 0), (3, 2, 1, 0), (2, 70000, 3, 64)])
def test_gpu_long_chains_equal_the_oracle(oracle, S, T, K, L):
    import torch
    from honeybadger_b200 import hmm
    rng = np.random.default_rng(S * 77 + T)
    lb = binom_lb(rng, K, T) if S == 2 else norm_lb(rng, K, T)
    Pi, delta = sym_Pi(S, 1e-6), np.array([0.2, 0.8] if S == 2 else [0., 1., 0.])
    y = hmm.viterbi_long_logb_dev(torch.from_numpy(lb).cuda(), Pi, delta, L)
    hmm.status_dev()
    rcs, yr = ref_paths(oracle, lb, Pi, delta)
    assert not any(rcs)
    assert np.array_equal(y.cpu().numpy(), yr)
    s = hmm.longchain_stats()
    assert s["chains"] == K
    if S == 2 and T >= 70000:
        assert s["sequential_chains"] == 0 and s["literal_blocks"] < K * s["blocks"] // 10, s


@pytest.mark.gpu
def test_gpu_adversarial_emissions(oracle):
    import torch
    from honeybadger_b200 import _lib, hmm
    rng = np.random.default_rng(19)
    for S, grid in [(2, 2.0 ** -40), (3, 2.0 ** -38)]:
        lb = -np.round(rng.gamma(2.0, 3.0, size=(4, 9000, S)) / grid) * grid - grid / 2
        lb[:, ::3, :] = 0.0
        lb[0, 500:520, 0] = -np.inf
        lb = np.ascontiguousarray(lb)
        Pi, delta = sym_Pi(S, 1e-6), np.array([0.5, 0.5] if S == 2 else [0., 1., 0.])
        y = hmm.viterbi_long_logb_dev(torch.from_numpy(lb).cuda(), Pi, delta, 32)
        hmm.status_dev()
        _, yr = ref_paths(oracle, lb, Pi, delta)
        assert np.array_equal(y.cpu().numpy(), yr)
    lb[:, -1, :] = -np.inf
    hmm.viterbi_long_logb_dev(torch.from_numpy(lb).cuda(), Pi, delta, 0)
    with pytest.raises(_lib.HBError):
        hmm.status_dev()


@pytest.mark.gpu
def test_pooled_rounds_same_states_in_every_mode(golden_gexp, golden_allele):
    """The pooled rounds of the bundled data: the long-chain path and the one-thread-per-chain kernels
    give the same states (both equal the golden paths in test_gpu_pooled)."""
    from honeybadger_b200 import hmm, honeybadger as hb
    g = golden_allele
    r, n = hb._to_dev_i32(g["r_maf"].astype(np.float64)), hb._to_dev_i32(g["n_sc"].astype(np.float64))
    member, _ = hb._membership(list(g["labels"]), g["r_maf"].shape[1])
    gx = hb._to_dev_f64(golden_gexp["gexp_norm_f32"].astype(np.float64))
    mem_x, _ = hb._membership(list(golden_gexp["labels"]), gx.shape[1])
    dev = float(golden_gexp["dev"])
    out = {}
    try:
        for mode in (0, 1, 2):
            hmm.set_longchain_mode(mode)
            out[mode] = (hb.allele_round_dev(r, n, member, 0.1, 0.45, 1e-6)[0],
                         hb.gexp_round_dev(gx, mem_x, [-dev, 0.0, dev], 1e-6)[0])
    finally:
        hmm.set_longchain_mode(0)
    for mode in (1, 2):
        assert np.array_equal(out[0][0], out[mode][0]) and np.array_equal(out[0][1], out[mode][1])


@pytest.mark.gpu
@pytest.mark.parametrize("S,T,K,L", [(2, 6000, 6, 0), (3, 6082, 15, 0), (3, 1, 2, 0), (2, 3000, 3, 16)])
def test_gpu_posteriors_match_the_oracle(oracle, S, T, K, L):
    import torch
    from honeybadger_b200 import hmm
    rng = np.random.default_rng(S * 5 + T)
    lb = binom_lb(rng, K, T, depth=60) if S == 2 else norm_lb(rng, K, T, sd=1.0)
    Pi, delta = sym_Pi(S, 1e-6), np.array([0.3, 0.7] if S == 2 else [0., 1., 0.])
    g, ll = hmm.forwardback_long_logb_dev(torch.from_numpy(lb).cuda(), Pi, delta, L)
    gr, llr = ref_posteriors(oracle, lb, Pi, delta)
    assert np.abs(g.cpu().numpy() - gr).max() < 1e-9
    assert np.allclose(ll.cpu().numpy(), llr, rtol=1e-12, atol=1e-9)


@pytest.mark.gpu
def test_gpu_posteriors_long_deep_chains():
    import torch
    from honeybadger_b200 import hmm
    rng = np.random.default_rng(4)
    lb = binom_lb(rng, 6, 200000)
    Pi, delta = sym_Pi(2, 1e-6), np.array([0., 1.])
    g, ll = hmm.forwardback_long_logb_dev(torch.from_numpy(lb).cuda(), Pi, delta)
    g = g.cpu().numpy()
    assert np.isfinite(g).all() and np.abs(g.sum(-1) - 1).max() < 1e-14 and np.isfinite(ll.cpu().numpy()).all()
    ref = longdouble_posteriors(lb[2], Pi, delta)
    assert np.abs(g[2] - ref).max() < 1e-10
    # posteriors and Viterbi states agree where the posterior is decisive
    y = hmm.viterbi_long_logb_dev(torch.from_numpy(lb).cuda(), Pi, delta).cpu().numpy()
    sure = g.max(-1) > 0.999
    assert (g.argmax(-1)[sure] + 1 == y[sure]).mean() > 0.999


@pytest.mark.gpu
def test_gpu_dropin_seam_uses_the_long_chain_path(oracle, golden_gexp, golden_allele):
    """Viterbi(dthmm(x, ...)) on ONE pooled vector — the reference's literal call (R/HoneyBADGER.R:1150-1151,
    :1409-1410) through hb_hmm_*_host — runs along the gene axis; states equal the oracle's and the
    one-thread-per-chain kernels' (mode 1)."""
    import time
    from honeybadger_b200 import hmm
    g = golden_gexp
    x, sd, dev = g["pooled_x"], g["pooled_sd"], float(g["dev"])
    Pi3, d3, mean = sym_Pi(3, 1e-6), np.array([0., 1., 0.]), np.array([-dev, 0.0, dev])
    O = oracle.lib()
    for k in (0, 7):
        y = hmm.hmm_norm_host(x[k], mean, float(sd[k]), Pi3, d3)["y"]
        assert hmm.longchain_stats()["chains"] == 1
        yo = np.zeros(x.shape[1], np.int32)
        assert O.hbo_viterbi_norm(3, C.c_int64(x.shape[1]), ptr(np.ascontiguousarray(x[k])), ptr(mean),
                                  ptr(np.full(3, float(sd[k]))), ptr(Pi3), ptr(d3), ptr(yo)) == 0
        assert np.array_equal(y, yo)
    a = golden_allele
    Pi2, d2, prob = sym_Pi(2, 1e-6), np.array([0., 1.]), np.array([0.1, 0.45])
    mafl, sizel = a["mafl"].astype(np.int32), a["sizel"].astype(np.int32)
    y0 = hmm.hmm_binom_host(mafl[0], sizel[0], prob, Pi2, d2)["y"]
    yo = np.zeros(mafl.shape[1], np.int32)
    assert O.hbo_viterbi_binom(2, C.c_int64(mafl.shape[1]), ptr(np.ascontiguousarray(mafl[0])),
                               ptr(np.ascontiguousarray(sizel[0])), ptr(prob), ptr(Pi2), ptr(d2), ptr(yo)) == 0
    assert np.array_equal(y0, yo)
    # a 200 000-SNP pooled vector: same states either way, and the seam is no longer latency-bound
    rng = np.random.default_rng(21)
    n = rng.binomial(5000, 0.108, size=200000).astype(np.int32)
    dele = np.zeros(200000, bool)
    dele[50000:58000] = True
    xx = rng.binomial(n, np.where(dele, 0.1, 0.45)).astype(np.int32)
    res = {}
    try:
        for mode in (0, 1):
            hmm.set_longchain_mode(mode)
            hmm.hmm_binom_host(xx, n, prob, Pi2, d2)
            t0 = time.perf_counter()
            res[mode] = hmm.hmm_binom_host(xx, n, prob, Pi2, d2)["y"]
            res[mode, "ms"] = (time.perf_counter() - t0) * 1e3
    finally:
        hmm.set_longchain_mode(0)
    assert np.array_equal(res[0], res[1])
    assert (res[0][52000:57000] == 1).all() and (res[0][:40000] == 2).all()
    print(f"one 200k-SNP chain through hb_hmm_binom_host: {res[0, 'ms']:.2f} ms along the gene axis, "
          f"{res[1, 'ms']:.2f} ms one thread")
    assert res[0, "ms"] < res[1, "ms"]


@pytest.mark.gpu
def test_gpu_seam_posteriors(oracle, golden_gexp):
    """forwardback(dthmm(x, ...)) on pooled vectors through hb_hmm_*_host: long-chain posteriors in R's layout,
    equal to the thread-per-chain kernels' (mode 1) and the oracle's within 1e-9; LL too."""
    from honeybadger_b200 import hmm
    g = golden_gexp
    x, sd, dev = g["pooled_x"][:4].T.copy(), g["pooled_sd"][:4], float(g["dev"])      # [T, B]
    Pi3, d3, mean = sym_Pi(3, 1e-6), np.array([0., 1., 0.]), np.array([-dev, 0.0, dev])
    res = {}
    try:
        for mode in (0, 1):
            hmm.set_longchain_mode(mode)
            res[mode] = hmm.hmm_norm_host(x, mean, sd, Pi3, d3, viterbi=True, gamma=True, ll=True)
    finally:
        hmm.set_longchain_mode(0)
    assert np.array_equal(res[0]["y"], res[1]["y"])
    assert np.abs(res[0]["gamma"] - res[1]["gamma"]).max() < 1e-9
    assert np.allclose(res[0]["ll"], res[1]["ll"], rtol=1e-12)
    yo, go, llo, rc = oracle.fbv_norm_batch(x, np.array([0, x.shape[0]], dtype=np.int64), mean, sd, Pi3, d3)
    assert rc == 0 and np.array_equal(res[0]["y"], yo)
    assert np.abs(res[0]["gamma"].transpose(2, 0, 1) - go).max() < 1e-9
    rng = np.random.default_rng(6)
    n = rng.binomial(60, 0.3, size=(5000, 2)).astype(np.int32)
    r = rng.binomial(n, 0.45).astype(np.int32)
    r[2000:2600, 0] = rng.binomial(n[2000:2600, 0], 0.1)
    Pi2, d2 = sym_Pi(2, 1e-6), np.array([0., 1.])
    try:
        for mode in (0, 1):
            hmm.set_longchain_mode(mode)
            res[mode] = hmm.hmm_binom_host(r, n, [0.1, 0.45], Pi2, d2, viterbi=True, gamma=True, ll=True)
    finally:
        hmm.set_longchain_mode(0)
    assert np.array_equal(res[0]["y"], res[1]["y"]) and (res[0]["y"][2100:2500, 0] == 1).all()
    assert np.abs(res[0]["gamma"] - res[1]["gamma"]).max() < 1e-9
    assert np.allclose(res[0]["ll"], res[1]["ll"], rtol=1e-12)


@pytest.mark.gpu
def test_gpu_random_shapes(oracle):
    """Odd chain counts and lengths around the staging-window and block sizes (nb = 1, 63, 64, 65, 129 ...),
    forced block lengths, coarse emission grids: device states equal the oracle's; posteriors 1e-9."""
    import torch
    from honeybadger_b200 import hmm
    rng = np.random.default_rng(77)
    cases = [(2, 1 + 64 * 16, 3, 16), (3, 2 + 64 * 16, 2, 16), (2, 65 * 16, 5, 16), (3, 129 * 32 + 7, 1, 32),
             (2, 63 * 16 + 1, 33, 16), (3, 4097, 40, 0), (2, 2, 1, 0), (3, 3, 7, 1)]
    for _ in range(24):
        cases.append((int(rng.choice([2, 3])), int(rng.integers(1, 6000)), int(rng.integers(1, 9)),
                      int(rng.choice([0, 1, 3, 16, 64, 512]))))
    for S, T, K, L in cases:
        scale = float(rng.choice([0.01, 1.0, 50.0]))
        lb = -rng.gamma(1.5, scale, size=(K, T, S))
        if rng.random() < 0.5:
            grid = 2.0 ** -int(rng.integers(20, 46))
            lb = np.round(lb / grid) * grid - grid / 2
        lb = np.ascontiguousarray(lb)
        Pi = sym_Pi(S, float(rng.choice([1e-8, 1e-6, 1e-2])))
        delta = rng.dirichlet(np.ones(S))
        dl = torch.from_numpy(lb).cuda()
        y = hmm.viterbi_long_logb_dev(dl, Pi, delta, L)
        hmm.status_dev()
        rcs, yr = ref_paths(oracle, lb, Pi, delta)
        assert not any(rcs) and np.array_equal(y.cpu().numpy(), yr), (S, T, K, L)
        if scale <= 1.0:
            g, ll = hmm.forwardback_long_logb_dev(dl, Pi, delta, L)
            gr, llr = ref_posteriors(oracle, lb, Pi, delta)
            assert np.abs(g.cpu().numpy() - gr).max() < 1e-9, (S, T, K, L)
            assert np.allclose(ll.cpu().numpy(), llr, rtol=1e-11, atol=1e-9)
