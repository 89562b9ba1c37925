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


def run_binom2(L, x, n, seg, prob, Pi, delta, flags, lut):
    T, B = x.shape
    nseg = seg.size - 1
    y = np.zeros((T, B), np.uint8)
    g = np.zeros((2, T, B))
    ll = np.zeros((nseg, B))
    st = L.emul_binom2(ptr(x), ptr(n), C.c_int64(B), C.c_int64(T), C.c_int64(B), ptr(seg), nseg,
                       ptr(prob), ptr(Pi), ptr(delta), flags, lut, ptr(y), ptr(g), ptr(ll))
    return st, y, g, ll


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
