"""CPU: the CUDA chain bodies (compiled for the host, tests/emul) against the oracle.

This exercises the kernels' logic — back-pointer packing, checkpoint/replay, cp.async group
accounting (emulated with deferred visibility), power-of-two rescaling, tie handling — without a
GPU.  The GPU parity tests (tests/test_gpu_*.py) run the real thing through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from honeybadger_b200 import synth
from oracle import pyref as P


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run_norm3(L, x, seg, mean, sd, Pi, delta, flags, force_general=0):
    T, B = x.shape
    nseg = seg.size - 1
    y = np.zeros((T, B), np.uint8)
    g = np.zeros((3, T, B))
    ll = np.zeros((nseg, B))
    st = L.emul_norm3(ptr(x), C.c_int64(B), C.c_int64(T), C.c_int64(B), ptr(seg), nseg, ptr(mean),
                      ptr(sd), 0, ptr(Pi), ptr(delta), flags, force_general, ptr(y), ptr(g), ptr(ll))
    return st, y, g, ll


def run_norm3_fast(L, x, seg, mean, sd, Pi, delta, flags, eps_scale=1.0):
    T, B = x.shape
    nseg = seg.size - 1
    y = np.zeros((T, B), np.uint8)
    g = np.zeros((3, T, B))
    ll = np.zeros((nseg, B))
    nflag = C.c_int32(0)
    st = L.emul_norm3_fast(ptr(x), C.c_int64(B), C.c_int64(T), C.c_int64(B), ptr(seg), nseg, ptr(mean),
                           ptr(sd), 0, ptr(Pi), ptr(delta), flags, C.c_double(eps_scale), ptr(y), ptr(g),
                           ptr(ll), C.byref(nflag))
    run_norm3_fast.ll = ll
    return st, y, g, nflag.value


def run_binom2(L, x, n, seg, prob, Pi, delta, flags, lut):
    T, B = x.shape
    nseg = seg.size - 1
    y = np.zeros((T, B), np.uint8)
    g = np.zeros((2, T, B))
    ll = np.zeros((nseg, B))
    st = L.emul_binom2(ptr(x), ptr(n), C.c_int64(B), C.c_int64(T), C.c_int64(B), ptr(seg), nseg,
                       ptr(prob), ptr(Pi), ptr(delta), flags, lut, ptr(y), ptr(g), ptr(ll))
    return st, y, g, ll


def run_binom2_fast(L, x, n, seg, prob, Pi, delta, flags, lut):
    T, B = x.shape
    nseg = seg.size - 1
    y = np.zeros((T, B), np.uint8)
    g = np.zeros((2, T, B))
    ll = np.zeros((nseg, B))
    gate = C.c_int32(0)
    st = L.emul_binom2_fast(ptr(x), ptr(n), C.c_int64(B), C.c_int64(T), C.c_int64(B), ptr(seg), nseg,
                            ptr(prob), ptr(Pi), ptr(delta), flags, lut, ptr(y), ptr(g), ptr(ll),
                            C.byref(gate))
    return st, y, g, ll, gate.value


@pytest.mark.parametrize("force_general", [0, 1])
@pytest.mark.parametrize("segs", [[0, 1203], [0, 2, 9, 17, 500, 508, 1203]])
def test_norm3_chain_vs_oracle(emul, oracle, force_general, segs):
    x, _, _ = synth.expression(1203, 9, seed=4)
    seg = np.array(segs, dtype=np.int64)
    sd = np.array([oracle.r_sd(x[:, b]) for b in range(x.shape[1])])
    mean, Pi, delta = np.array([-synth.DEV, 0, synth.DEV]), oracle.sym_Pi(3, 1e-6), np.array([0., 1, 0])
    yo, go, llo, rc = oracle.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    assert rc == 0
    for flags in (7, 3, 1, 2):
        st, y, g, ll = run_norm3(emul, x, seg, mean, sd, Pi, delta, flags, force_general)
        assert st == 0
        if flags & 1:
            assert (y == yo).all()
        if flags & 2:
            assert np.abs(g - go).max() < 1e-9
        if flags & 4:
            assert np.allclose(ll, llo, rtol=1e-12, atol=0)


@pytest.mark.parametrize("eps_scale", [1.0, float("inf")])
@pytest.mark.parametrize("segs", [[0, 1203], [0, 2, 9, 17, 500, 508, 1203], [0, 2, 4, 7, 1203]])
def test_norm3_fast_chain_vs_oracle(emul, oracle, eps_scale, segs):
    """Speculate-and-verify body: same paths as the oracle whether a chain is verified by its margin
    (eps_scale 1) or re-run through the exact fallback (eps_scale inf); posteriors within 1e-9."""
    x, _, _ = synth.expression(1203, 9, seed=4)
    seg = np.array(segs, dtype=np.int64)
    sd = np.array([oracle.r_sd(x[:, b]) for b in range(x.shape[1])])
    mean, Pi, delta = np.array([-synth.DEV, 0, synth.DEV]), oracle.sym_Pi(3, 1e-6), np.array([0., 1, 0])
    yo, go, llo, rc = oracle.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    nchains = x.shape[1] * (seg.size - 1)
    for flags in (7, 3, 1, 2, 6):
        st, y, g, nflag = run_norm3_fast(emul, x, seg, mean, sd, Pi, delta, flags, eps_scale)
        assert st == 0 and rc == 0
        if flags & 1:
            assert (y == yo).all()
            assert nflag == (nchains if eps_scale > 1 else 0)
        if flags & 2:
            assert np.abs(g - go).max() < 1e-9
            assert np.abs(g.sum(0) - 1).max() < 1e-11
        if flags & 4:
            assert np.allclose(run_norm3_fast.ll, llo, rtol=1e-12, atol=0)


def test_norm3_fast_flags_ties_and_near_ties(emul, oracle):
    """Adversarial chains: exact ties (x = 0 keeps both outer states level), decisions within a few
    ulps, extreme outliers.  Every such chain must either be verified or take the exact fallback, and
    the path must equal the oracle's."""
    rng = np.random.default_rng(11)
    T, B = 700, 24
    x = rng.standard_normal((T, B)) * 2.3
    x[:, 0] = 0.0                                   # symmetric: nu_0 == nu_2 exactly at every step
    x[:, 1] = np.where(np.arange(T) % 2 == 0, 1.5, -1.5)
    x[100:300, 2] += 3.0
    x[350:500, 3] -= 3.0
    x[:, 4] = np.round(x[:, 4], 1)                  # coarse grid: repeated values, structural ties
    x[123, 5] = 400.0                               # absurd outlier: exponent clamps, huge margins
    x[:, 6] = 1e-300
    seg = np.array([0, 350, 700], dtype=np.int64)
    sd = np.full(B, 2.3)
    mean, Pi, delta = np.array([-1.0, 0, 1.0]), oracle.sym_Pi(3, 1e-6), np.array([0., 1, 0])
    yo, go, _, rc = oracle.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    st, y, g, nflag = run_norm3_fast(emul, x, seg, mean, sd, Pi, delta, 3)
    assert rc == 0 and st == 0
    assert (y == yo).all()
    assert nflag >= 2          # the two all-zero chains tie at every step
    # posteriors: the linear-domain reference underflows for the outlier column; compare the rest
    keep = [b for b in range(B) if b != 5]
    assert np.abs(g[:, :, keep] - go[:, :, keep]).max() < 1e-9
    # with verification switched off the all-zero chain is NOT guaranteed: show the hook matters
    st0, y0, _, nflag0 = run_norm3_fast(emul, x, seg, mean, sd, Pi, delta, 1, eps_scale=0.0)
    assert nflag0 <= nflag


def test_norm3_fast_other_priors(emul, oracle):
    """delta with all states reachable, asymmetric in value (the differences start finite)."""
    x, _, _ = synth.expression(900, 6, seed=9)
    seg = np.array([0, 450, 900], dtype=np.int64)
    sd = np.array([oracle.r_sd(x[:, b]) for b in range(x.shape[1])])
    mean, Pi, delta = np.array([-0.8, 0.1, 1.0]), oracle.sym_Pi(3, 1e-4), np.array([0.2, 0.5, 0.3])
    yo, go, _, rc = oracle.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    yo, go, llo, rc = oracle.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    st, y, g, _ = run_norm3_fast(emul, x, seg, mean, sd, Pi, delta, 7)
    assert st == 0 and rc == 0 and (y == yo).all()
    assert np.abs(g - go).max() < 1e-9
    assert np.allclose(run_norm3_fast.ll, llo, rtol=1e-12, atol=0)


def test_norm3_general_parameters(emul, oracle):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((800, 5)) * 1.2
    x[200:400] += 1.0
    sd = np.array([oracle.r_sd(x[:, b]) for b in range(5)])
    Pi = np.array([[0.97, 0.02, 0.01], [0.005, 0.99, 0.005], [0.03, 0.01, 0.96]])
    mean, delta = np.array([-1.3, 0.2, 0.8]), np.array([0.3, 0.4, 0.3])
    seg = np.array([0, 333, 800], dtype=np.int64)
    yo, go, llo, rc = oracle.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    st, y, g, ll = run_norm3(emul, x, seg, mean, sd, Pi, delta, 7)
    assert st == 0 and rc == 0 and (y == yo).all()
    assert np.abs(g - go).max() < 1e-9 and np.allclose(ll, llo, rtol=1e-12, atol=0)


def test_norm3_fast_length_one_chain_flags_underflow(emul, oracle):
    x = np.array([[0.3], [0.1]])
    for eps_scale in (1.0, float("inf")):
        st, *_ = run_norm3_fast(emul, x, np.array([0, 1, 2], dtype=np.int64), np.array([-1., 0, 1]),
                                np.array([1.0]), oracle.sym_Pi(3, 1e-6), np.array([0., 1, 0]), 1, eps_scale)
        assert st == 1


def test_norm3_length_one_chain_flags_underflow(emul, oracle):
    x = np.array([[0.3], [0.1]])
    st, *_ = run_norm3(emul, x, np.array([0, 1, 2], dtype=np.int64), np.array([-1., 0, 1]),
                       np.array([1.0]), oracle.sym_Pi(3, 1e-6), np.array([0., 1, 0]), 1)
    assert st == 1  # nu = (-Inf, v, -Inf) at the only step: HiddenMarkov stops with "Underflow"


@pytest.mark.parametrize("lut", [511, 31, 0])
@pytest.mark.parametrize("segs", [[0, 3001], [0, 700, 1301, 1305, 3001]])
def test_binom2_chain_vs_oracle(emul, oracle, lut, segs):
    r, n, _, _ = synth.allele(3001, 10, seed=6, deep=0.002)
    seg = np.array(segs, dtype=np.int64)
    prob, Pi, delta = np.array([0.1, 0.45]), oracle.sym_Pi(2, 1e-6), np.array([0., 1.])
    yo, go, llo, rc = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta)
    st, y, g, ll = run_binom2(emul, r, n, seg, prob, Pi, delta, 7, lut)
    assert st == 0 and rc == 0
    assert (y == yo).all()
    assert np.abs(g - go).max() < 1e-9
    assert np.allclose(ll, llo, rtol=1e-10, atol=1e-10)
    st, y2, _, _ = run_binom2(emul, r, n, seg, prob, Pi, delta, 1, lut)
    assert (y2 == yo).all()


def test_binom2_pooled_scale_counts_vs_exact_arithmetic(emul, oracle):
    """Counts of thousands (pooled groups): per-site evidence of hundreds of nats.  HiddenMarkov's
    own linear-domain recursion underflows here; the reference for the posteriors is mpmath."""
    rng = np.random.default_rng(7)
    T = 300
    n = rng.integers(2000, 6000, size=(T, 1)).astype(np.int32)
    p = np.full((T, 1), 0.45)
    p[80:160] = 0.1
    x = rng.binomial(n, p).astype(np.int32)
    prob, Pi, delta = np.array([0.1, 0.45]), oracle.sym_Pi(2, 1e-6), np.array([0., 1.])
    st, y, g, _ = run_binom2(emul, x, n, np.array([0, T], dtype=np.int64), prob, Pi, delta, 7, 511)
    logb = [[P.dbinom_log(float(a), float(c), q) for q in prob] for a, c in zip(x[:, 0], n[:, 0])]
    assert np.abs(g[:, :, 0].T - P.mp_posteriors(logb, Pi, delta, dps=80)).max() < 1e-12
    assert (y[:, 0] == oracle.viterbi_binom(x[:, 0], n[:, 0], prob, Pi, delta)).all()


@pytest.mark.parametrize("lut", [511, 31, 0])
@pytest.mark.parametrize("segs", [[0, 3001], [0, 700, 1301, 1305, 3001], [0, 2, 4, 3001]])
def test_binom2_fast_chain_vs_oracle(emul, oracle, lut, segs):
    """Lean binomial body: literal Viterbi (paths equal the oracle's, structural ties included),
    normalisation-free posteriors within 1e-9, log-likelihood, for table and off-table counts."""
    r, n, _, _ = synth.allele(3001, 10, seed=6, deep=0.002)
    seg = np.array(segs, dtype=np.int64)
    prob, Pi, delta = np.array([0.1, 0.45]), oracle.sym_Pi(2, 1e-6), np.array([0., 1.])
    yo, go, llo, rc = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta)
    assert rc == 0
    for flags in (7, 3, 1, 2, 6):
        st, y, g, ll, gate = run_binom2_fast(emul, r, n, seg, prob, Pi, delta, flags, lut)
        assert st == 0 and gate == 0
        if flags & 1:
            assert (y == yo).all()
        if flags & 2:
            assert np.abs(g - go).max() < 1e-9
            assert np.abs(g.sum(0) - 1).max() < 1e-12
        if flags & 4:
            assert np.allclose(ll, llo, rtol=1e-10, atol=1e-10)


def test_binom2_fast_general_Pi_and_prior(emul, oracle):
    r, n, _, _ = synth.allele(1500, 7, seed=8)
    seg = np.array([0, 600, 1500], dtype=np.int64)
    prob, delta = np.array([0.2, 0.5]), np.array([0.3, 0.7])
    Pi = np.array([[0.98, 0.02], [0.001, 0.999]])
    yo, go, llo, rc = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta)
    st, y, g, ll, gate = run_binom2_fast(emul, r, n, seg, prob, Pi, delta, 7, 64)
    assert st == 0 and rc == 0 and gate == 0
    assert (y == yo).all() and np.abs(g - go).max() < 1e-9
    assert np.allclose(ll, llo, rtol=1e-10, atol=1e-10)


def test_binom2_fast_gate_on_extreme_evidence(emul, oracle):
    """Pooled-scale counts (thousands of reads per position) exceed what a double ratio can hold: the
    lean body must raise its gate, and the general body then produces the exact-arithmetic answer."""
    rng = np.random.default_rng(7)
    T = 300
    n = rng.integers(2000, 6000, size=(T, 1)).astype(np.int32)
    p = np.full((T, 1), 0.45)
    p[80:160] = 0.1
    x = rng.binomial(n, p).astype(np.int32)
    prob, Pi, delta = np.array([0.1, 0.45]), oracle.sym_Pi(2, 1e-6), np.array([0., 1.])
    st, y, g, _, gate = run_binom2_fast(emul, x, n, np.array([0, T], dtype=np.int64), prob, Pi, delta, 3, 511)
    assert gate == 1 and st == 0
    logb = [[P.dbinom_log(float(a), float(c), q) for q in prob] for a, c in zip(x[:, 0], n[:, 0])]
    assert np.abs(g[:, :, 0].T - P.mp_posteriors(logb, Pi, delta, dps=80)).max() < 1e-12
    assert (y[:, 0] == oracle.viterbi_binom(x[:, 0], n[:, 0], prob, Pi, delta)).all()
    # moderate depth (hundreds of reads): within the plain-double range, no gate, same accuracy
    n2 = rng.integers(100, 400, size=(T, 1)).astype(np.int32)
    x2 = rng.binomial(n2, p).astype(np.int32)
    st, y2, g2, _, gate2 = run_binom2_fast(emul, x2, n2, np.array([0, T], dtype=np.int64), prob, Pi, delta, 3, 511)
    logb2 = [[P.dbinom_log(float(a), float(c), q) for q in prob] for a, c in zip(x2[:, 0], n2[:, 0])]
    assert gate2 == 0 and st == 0
    assert np.abs(g2[:, :, 0].T - P.mp_posteriors(logb2, Pi, delta, dps=80)).max() < 1e-11
    assert (y2[:, 0] == oracle.viterbi_binom(x2[:, 0], n2[:, 0], prob, Pi, delta)).all()


# ---------------------------------------------------------------------------- warp-tiled layout
def np_tile(a):
    """[T, B] -> [ntile, T, 32] (zero padded), numpy restatement of hb_tile_dev."""
    T, B = a.shape
    nt = (B + 31) // 32
    p = np.zeros((T, nt * 32), dtype=a.dtype)
    p[:, :B] = a
    return np.ascontiguousarray(p.reshape(T, nt, 32).transpose(1, 0, 2))


def np_untile(t, B):
    nt, T, _ = t.shape
    return np.ascontiguousarray(t.transpose(1, 0, 2).reshape(T, nt * 32)[:, :B])


@pytest.mark.parametrize("which", ["fast", "literal"])
def test_norm3_tiled_layout(emul, oracle, which):
    """The chain bodies on the warp-tiled layout (B not a multiple of 32) give the oracle's results."""
    x, _, _ = synth.expression(803, 45, seed=21)
    T, B = x.shape
    seg = np.array([0, 300, 311, 803], dtype=np.int64)
    sd = np.array([oracle.r_sd(x[:, b]) for b in range(B)])
    mean, Pi, delta = np.array([-synth.DEV, 0, synth.DEV]), oracle.sym_Pi(3, 1e-6), np.array([0., 1, 0])
    yo, go, llo, rc = oracle.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    xt = np_tile(x)
    nt = xt.shape[0]
    y = np.zeros((nt, T, 32), np.uint8)
    g = np.zeros((3, nt, T, 32))
    ll = np.zeros((seg.size - 1, B))
    emul.emul_set_tiled(1)
    try:
        if which == "fast":
            nflag = C.c_int32(0)
            st = emul.emul_norm3_fast(ptr(xt), C.c_int64(32), C.c_int64(T), C.c_int64(B), ptr(seg), seg.size - 1,
                                      ptr(mean), ptr(sd), 0, ptr(Pi), ptr(delta), 7, C.c_double(1.0), ptr(y),
                                      ptr(g), ptr(ll), C.byref(nflag))
        else:
            st = emul.emul_norm3(ptr(xt), C.c_int64(32), C.c_int64(T), C.c_int64(B), ptr(seg), seg.size - 1,
                                 ptr(mean), ptr(sd), 0, ptr(Pi), ptr(delta), 7, 0, ptr(y), ptr(g), ptr(ll))
    finally:
        emul.emul_set_tiled(0)
    assert st == 0 and rc == 0
    assert (np_untile(y, B) == yo).all()
    gu = np.stack([np_untile(g[k], B) for k in range(3)])
    assert np.abs(gu - go).max() < 1e-9
    assert np.allclose(ll, llo, rtol=1e-12, atol=0)


@pytest.mark.parametrize("which", ["fast", "general"])
def test_binom2_tiled_layout(emul, oracle, which):
    r, n, _, _ = synth.allele(1203, 40, seed=23, deep=0.002)
    T, B = r.shape
    seg = np.array([0, 500, 507, 1203], dtype=np.int64)
    prob, Pi, delta = np.array([0.1, 0.45]), oracle.sym_Pi(2, 1e-6), np.array([0., 1.])
    yo, go, llo, rc = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta)
    rt, nt_ = np_tile(r), np_tile(n)
    nt = rt.shape[0]
    y = np.zeros((nt, T, 32), np.uint8)
    g = np.zeros((2, nt, T, 32))
    ll = np.zeros((seg.size - 1, B))
    emul.emul_set_tiled(1)
    try:
        if which == "fast":
            gate = C.c_int32(0)
            st = emul.emul_binom2_fast(ptr(rt), ptr(nt_), C.c_int64(32), C.c_int64(T), C.c_int64(B), ptr(seg),
                                       seg.size - 1, ptr(prob), ptr(Pi), ptr(delta), 7, 64, ptr(y), ptr(g), ptr(ll),
                                       C.byref(gate))
            assert gate.value == 0
        else:
            st = emul.emul_binom2(ptr(rt), ptr(nt_), C.c_int64(32), C.c_int64(T), C.c_int64(B), ptr(seg),
                                  seg.size - 1, ptr(prob), ptr(Pi), ptr(delta), 7, 64, ptr(y), ptr(g), ptr(ll))
    finally:
        emul.emul_set_tiled(0)
    assert st == 0 and rc == 0
    assert (np_untile(y, B) == yo).all()
    gu = np.stack([np_untile(g[k], B) for k in range(2)])
    assert np.abs(gu - go).max() < 1e-9
    assert np.allclose(ll, llo, rtol=1e-10, atol=1e-10)


def test_binom2_packed_counts(emul, oracle):
    """One (n << 16) | x word per position: same results as the two-plane form, lean and general
    bodies (the general one behind the gate for pooled-scale counts up to 65535)."""
    r, n, _, _ = synth.allele(1503, 12, seed=29, deep=0.002)
    n[:, 3] = np.minimum(n[:, 3] * 300, 60000)         # deep column: raises the gate
    r[:, 3] = np.minimum(r[:, 3] * 300, n[:, 3])
    seg = np.array([0, 700, 1503], dtype=np.int64)
    prob, Pi, delta = np.array([0.1, 0.45]), oracle.sym_Pi(2, 1e-6), np.array([0., 1.])
    yo, go, llo, rc = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta)
    xn = ((n.astype(np.uint32) << 16) | r.astype(np.uint32)).view(np.int32)
    emul.emul_set_packed(1)
    try:
        st, y, g, ll, gate = run_binom2_fast(emul, xn, np.zeros_like(xn), seg, prob, Pi, delta, 3, 64)
    finally:
        emul.emul_set_packed(0)
    assert st == 0 and rc == 0 and gate == 1
    assert (y == yo).all()
    keep = [b for b in range(12) if b != 3]   # (the linear-domain oracle underflows on the deep column)
    assert np.abs(g[:, :, keep] - go[:, :, keep]).max() < 1e-9
    emul.emul_set_packed(1)
    try:
        st, y2, g2, _, gate2 = run_binom2_fast(emul, np.ascontiguousarray(xn[:, keep]), np.zeros((1503, 11), np.int32),
                                               seg, prob, Pi, delta, 3, 64)
    finally:
        emul.emul_set_packed(0)
    assert gate2 == 0 and (y2 == yo[:, keep]).all() and np.abs(g2 - go[:, :, keep]).max() < 1e-9


# ---------------------------------------------------------------------------- randomised shapes / parameters
from hypothesis import HealthCheck, given, settings, strategies as st_  # noqa: E402


@st_.composite
def _chain_case(draw):
    T = draw(st_.integers(1, 90))
    B = draw(st_.integers(1, 37))
    ncut = draw(st_.integers(0, 4))
    cuts = sorted(draw(st_.lists(st_.integers(0, T), min_size=ncut, max_size=ncut)))
    seg = np.array([0] + cuts + [T], dtype=np.int64)
    t = draw(st_.sampled_from([1e-8, 1e-6, 1e-3, 0.05, 0.2]))
    seed = draw(st_.integers(0, 2 ** 31 - 1))
    tiled = draw(st_.booleans())
    return T, B, seg, t, seed, tiled


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck))
@given(_chain_case(), st_.sampled_from([(0.0, 1.0, 0.0), (0.2, 0.5, 0.3), (0.0, 0.4, 0.6)]),
       st_.floats(0.3, 2.0), st_.sampled_from([0.0, 1.0]))
def test_norm3_fast_random_cases(emul, oracle, case, delta, m, coarse):
    """Random lengths (incl. 1 and chunk boundaries), ragged / empty segments, transition rates, priors,
    coarse-grid observations (many exact ties) and both layouts: paths equal the oracle's, always."""
    T, B, seg, t, seed, tiled = case
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((T, B)) * 1.7
    x[T // 3: 2 * T // 3, : B // 2] += m
    if coarse:
        x = np.round(x, 0)
    sd = rng.uniform(0.7, 2.5, B)
    mean, Pi, dl = np.array([-m, 0.0, m]), oracle.sym_Pi(3, t), np.array(delta)
    yo, go, _, rc = oracle.fbv_norm_batch(x, seg, mean, sd, Pi, dl)
    if rc != 0:                       # some chain ends with nu = -Inf (length-1 chain, zero prior)
        return
    if tiled:
        xt = np_tile(x)
        y = np.zeros((xt.shape[0], T, 32), np.uint8)
        g = np.zeros((3, xt.shape[0], T, 32))
        emul.emul_set_tiled(1)
    else:
        xt, y, g = x, np.zeros((T, B), np.uint8), np.zeros((3, T, B))
    ll = np.zeros((seg.size - 1, B))
    nflag = C.c_int32(0)
    try:
        st = emul.emul_norm3_fast(ptr(xt), C.c_int64(32 if tiled else B), C.c_int64(T), C.c_int64(B), ptr(seg),
                                  seg.size - 1, ptr(mean), ptr(sd), 0, ptr(Pi), ptr(dl), 3, C.c_double(1.0),
                                  ptr(y), ptr(g), ptr(ll), C.byref(nflag))
    finally:
        emul.emul_set_tiled(0)
    assert st == 0
    yy = np_untile(y, B) if tiled else y
    gg = np.stack([np_untile(g[k], B) for k in range(3)]) if tiled else g
    live = np.zeros(T, bool)          # positions covered by a segment
    for a, b in zip(seg[:-1], seg[1:]):
        live[a:b] = True
    assert (yy[live] == yo[live]).all()
    assert np.abs(gg[:, live] - go[:, live]).max() < 1e-9


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck))
@given(_chain_case(), st_.sampled_from([(0.0, 1.0), (0.3, 0.7)]), st_.sampled_from([0, 5, 64]),
       st_.booleans())
def test_binom2_fast_random_cases(emul, oracle, case, delta, lut, packed):
    T, B, seg, t, seed, tiled = case
    rng = np.random.default_rng(seed)
    n = np.where(rng.random((T, B)) < 0.3, rng.integers(1, 12, (T, B)), 0).astype(np.int32)
    p = np.full((T, B), 0.45)
    p[T // 4: T // 2, : B // 2] = 0.1
    r = rng.binomial(n, p).astype(np.int32)
    prob, Pi, dl = np.array([0.1, 0.45]), oracle.sym_Pi(2, t), np.array(delta)
    yo, go, _, rc = oracle.fbv_binom_batch(r, n, seg, prob, Pi, dl)
    if rc != 0:
        return
    xin = ((n.astype(np.uint32) << 16) | r.astype(np.uint32)).view(np.int32) if packed else r
    nin = np.zeros_like(n) if packed else n
    if tiled:
        xin, nin = np_tile(xin), np_tile(nin)
        y = np.zeros((xin.shape[0], T, 32), np.uint8)
        g = np.zeros((2, xin.shape[0], T, 32))
    else:
        y, g = np.zeros((T, B), np.uint8), np.zeros((2, T, B))
    ll = np.zeros((seg.size - 1, B))
    gate = C.c_int32(0)
    emul.emul_set_tiled(int(tiled))
    emul.emul_set_packed(int(packed))
    try:
        st = emul.emul_binom2_fast(ptr(xin), ptr(nin), C.c_int64(32 if tiled else B), C.c_int64(T), C.c_int64(B),
                                   ptr(seg), seg.size - 1, ptr(prob), ptr(Pi), ptr(dl), 3, lut, ptr(y), ptr(g),
                                   ptr(ll), C.byref(gate))
    finally:
        emul.emul_set_tiled(0)
        emul.emul_set_packed(0)
    assert st == 0 and gate.value == 0
    yy = np_untile(y, B) if tiled else y
    gg = np.stack([np_untile(g[k], B) for k in range(2)]) if tiled else g
    live = np.zeros(T, bool)
    for a, b in zip(seg[:-1], seg[1:]):
        live[a:b] = True
    assert (yy[live] == yo[live]).all()
    assert np.abs(gg[:, live] - go[:, live]).max() < 1e-9
