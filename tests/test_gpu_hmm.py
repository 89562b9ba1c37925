"""GPU parity: CUDA chains (through the C ABI) vs the CPU oracle on seeded inputs."""
from __future__ import annotations

import numpy as np
import pytest

from honeybadger_b200 import synth

pytestmark = pytest.mark.gpu


def _sd_cols(O, x):
    return np.array([O.r_sd(x[:, b]) for b in range(x.shape[1])])


@pytest.mark.parametrize("nseg", [1, 22])
def test_norm_host_vs_oracle(oracle, nseg):
    from honeybadger_b200 import hmm
    O = oracle
    G, N = 3000, 300
    x, seg22, _ = synth.expression(G, N, seed=2)
    seg = seg22 if nseg == 22 else np.array([0, G], dtype=np.int64)
    sd = _sd_cols(O, x)
    Pi, delta, mean = O.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0], [-synth.DEV, 0.0, synth.DEV]
    r = hmm.hmm_norm_host(x, mean, sd, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    yo, go, llo, rc = O.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    assert rc == 0
    assert r["y"].dtype == np.int32 and r["y"].shape == (G, N)
    assert (r["y"] == yo).all(), f"{(r['y'] != yo).sum()} Viterbi states differ"
    assert np.abs(r["gamma"].transpose(2, 0, 1) - go).max() < 1e-9  # north_star tolerance
    assert np.allclose(r["ll"].T, llo, rtol=1e-10, atol=0)
    assert set(np.unique(r["y"])) <= {1, 2, 3} and (r["y"] != 2).any()


def test_norm_general_path_vs_oracle(oracle):
    """Asymmetric means + arbitrary Pi take the general kernel."""
    from honeybadger_b200 import hmm
    O = oracle
    rng = np.random.default_rng(9)
    G, N = 1500, 130
    x = rng.standard_normal((G, N)) * 1.3
    x[400:700, :60] += 0.9
    x[900:1200, 40:] -= 1.4
    sd = _sd_cols(O, x)
    Pi = np.array([[0.98, 0.015, 0.005], [0.001, 0.998, 0.001], [0.02, 0.01, 0.97]])
    delta, mean = [0.2, 0.5, 0.3], [-1.4, 0.1, 0.9]
    seg = np.array([0, 700, G], dtype=np.int64)
    r = hmm.hmm_norm_host(x, mean, sd, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    yo, go, llo, rc = O.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    assert rc == 0 and (r["y"] == yo).all()
    assert np.abs(r["gamma"].transpose(2, 0, 1) - go).max() < 1e-9
    assert np.allclose(r["ll"].T, llo, rtol=1e-10, atol=0)


@pytest.mark.parametrize("nseg", [1, 22])
def test_binom_host_vs_oracle(oracle, nseg):
    from honeybadger_b200 import hmm
    O = oracle
    G, N = 6000, 200
    r_maf, n_sc, seg22, _ = synth.allele(G, N, seed=3)
    seg = seg22 if nseg == 22 else np.array([0, G], dtype=np.int64)
    Pi, delta, prob = O.sym_Pi(2, 1e-6), [0.0, 1.0], [0.1, 0.45]
    r = hmm.hmm_binom_host(r_maf, n_sc, prob, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    yo, go, llo, rc = O.fbv_binom_batch(r_maf, n_sc, seg, prob, Pi, delta)
    assert rc == 0
    assert (r["y"] == yo).all(), f"{(r['y'] != yo).sum()} Viterbi states differ"
    assert np.abs(r["gamma"].transpose(2, 0, 1) - go).max() < 1e-9
    assert np.allclose(r["ll"].T, llo, rtol=1e-10, atol=0)
    assert (r["y"] == 1).any()


def test_vector_api_and_underflow(oracle):
    """dthmm/Viterbi mirror on one chain; a length-1 chain reproduces R's underflow error."""
    from honeybadger_b200 import _lib, hmm
    O = oracle
    rng = np.random.default_rng(1)
    x = rng.standard_normal(500)
    x[100:200] += 1.5
    sd = O.r_sd(x)
    z = hmm.dthmm(x, O.sym_Pi(3, 1e-6), [0, 1, 0], "norm", dict(mean=[-1, 0, 1], sd=[sd] * 3))
    y = hmm.Viterbi(z)
    assert (y == O.viterbi_norm(x, [-1, 0, 1], [sd] * 3, O.sym_Pi(3, 1e-6), [0, 1, 0])).all()
    fb = hmm.forwardback(z)
    g, ll = O.forwardback_norm(x, [-1, 0, 1], [sd] * 3, O.sym_Pi(3, 1e-6), [0, 1, 0])
    assert np.abs(fb["u"] - g).max() < 1e-9 and abs(fb["LL"][0] - ll) < 1e-9 * abs(ll)
    with pytest.raises(_lib.UnderflowError, match="Problems With Underflow"):
        hmm.Viterbi(hmm.dthmm(x[:1], O.sym_Pi(3, 1e-6), [0, 1, 0], "norm",
                              dict(mean=[-1, 0, 1], sd=[1.0] * 3)))


def test_dev_api_matches_host_api(oracle):
    import torch
    from honeybadger_b200 import hmm
    O = oracle
    G, N = 2000, 260
    x, seg, _ = synth.expression(G, N, seed=5)
    Pi, delta, mean = O.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0], [-synth.DEV, 0.0, synth.DEV]
    xd = torch.from_numpy(x).cuda()
    sd = hmm.chain_sd_dev(xd)
    sd_ref = _sd_cols(O, x)
    assert np.allclose(sd.cpu().numpy(), sd_ref, rtol=1e-14, atol=0)
    out = hmm.hmm_norm_dev(xd, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    hmm.status_dev()
    sdh = sd.cpu().numpy()
    yo, go, llo, rc = O.fbv_norm_batch(x, seg, mean, sdh, Pi, delta)
    assert (out["y"].cpu().numpy() == yo).all()
    assert np.abs(out["gamma"].cpu().numpy() - go).max() < 1e-9
    assert np.allclose(out["ll"].cpu().numpy(), llo, rtol=1e-10, atol=0)


@pytest.mark.parametrize("nseg", [1, 22])
def test_norm_kernel_modes_agree(oracle, nseg):
    """The speculate-and-verify kernel (mode 3), the literal-arithmetic kernel (mode 1) and the
    speculative kernel with every chain re-run exactly (mode 2) give the same paths as the oracle;
    posteriors within 1e-9 of it and within 1e-11 of each other."""
    import torch
    from honeybadger_b200 import _lib, hmm
    O = oracle
    G, N = 3500, 400
    x, seg22, _ = synth.expression(G, N, seed=12)
    x[:, 7] = 0.0
    x[:, 8] = np.round(x[:, 8], 1)     # coarse values: repeated decisions
    x[1000, 9] = 300.0                 # absurd outlier
    seg = seg22 if nseg == 22 else np.array([0, G], dtype=np.int64)
    sd = np.full(N, 2.5)
    Pi, delta, mean = O.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0], [-synth.DEV, 0.0, synth.DEV]
    yo, go, llo, rc = O.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    assert rc == 0
    xd, sdd = torch.from_numpy(x).cuda(), torch.from_numpy(sd).cuda()
    lib = _lib.load()
    res = {}
    try:
        for mode in (3, 1, 2):
            _lib.check(lib.hb_set_norm3_mode(mode))
            out = hmm.hmm_norm_dev(xd, sdd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
            hmm.status_dev()
            res[mode] = (out["y"].cpu().numpy(), out["gamma"].cpu().numpy(), out["ll"].cpu().numpy(),
                         hmm.norm3_rerun_chains())
    finally:
        lib.hb_set_norm3_mode(0)
    keep = np.array([b for b in range(N) if b != 9])  # the linear-domain oracle underflows at the outlier
    for mode, (y, g, ll, nre) in res.items():
        assert (y == yo).all(), f"mode {mode}: {(y != yo).sum()} states differ"
        assert np.abs(g[:, :, keep] - go[:, :, keep]).max() < 1e-9
        assert np.allclose(ll[:, keep], llo[:, keep], rtol=1e-10, atol=0)
    assert np.abs(res[3][1] - res[1][1]).max() < 1e-11
    nchains = N * (seg.size - 1)
    if nseg == 22:   # chains short enough for the speculative kernel
        assert res[3][3] < nchains // 20           # almost nothing needs the exact re-run
        assert res[2][3] == nchains
    assert res[1][3] == 0


@pytest.mark.parametrize("nseg", [1, 22])
def test_binom_kernel_modes_agree(oracle, nseg):
    """Lean binomial kernel (auto mode) vs general kernel (mode 1) vs the oracle, with counts beyond
    the emission table and a deep-coverage column; the gate stays closed for per-cell data."""
    import torch
    from honeybadger_b200 import _lib, hmm
    O = oracle
    G, N = 5000, 300
    r_maf, n_sc, seg22, _ = synth.allele(G, N, seed=14, deep=0.002)
    n_sc[:, 5] = np.minimum(n_sc[:, 5] * 40, 600)      # counts above the 511-entry table
    r_maf[:, 5] = np.minimum(r_maf[:, 5] * 40, n_sc[:, 5])
    seg = seg22 if nseg == 22 else np.array([0, G], dtype=np.int64)
    Pi, delta, prob = O.sym_Pi(2, 1e-6), [0.0, 1.0], [0.1, 0.45]
    yo, go, llo, rc = O.fbv_binom_batch(r_maf, n_sc, seg, prob, Pi, delta)
    assert rc == 0
    rd, nd = torch.from_numpy(r_maf).cuda(), torch.from_numpy(n_sc).cuda()
    lib = _lib.load()
    res = {}
    try:
        for mode in (0, 1):
            _lib.check(lib.hb_set_binom2_mode(mode))
            out = hmm.hmm_binom_dev(rd, nd, prob, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
            hmm.status_dev()
            res[mode] = (out["y"].cpu().numpy(), out["gamma"].cpu().numpy(), out["ll"].cpu().numpy(),
                         hmm.binom2_general_rerun())
    finally:
        lib.hb_set_binom2_mode(0)
    for mode, (y, g, ll, rerun) in res.items():
        assert (y == yo).all(), f"mode {mode}: {(y != yo).sum()} states differ"
        assert np.abs(g - go).max() < 1e-9
        assert np.allclose(ll, llo, rtol=1e-10, atol=1e-9)
    assert res[1][3] == 0   # (the general-only mode never reports a re-run; auto may, for the deep column)
    assert np.abs(res[0][1] - res[1][1]).max() < 1e-11


def test_binom_gate_reruns_general_kernel(oracle):
    """Pooled-scale counts: the lean kernel raises its gate and the general kernel redoes the launch."""
    import torch
    from honeybadger_b200 import hmm
    from oracle import pyref as P
    rng = np.random.default_rng(7)
    T = 300
    n = rng.integers(2000, 6000, size=(T, 40)).astype(np.int32)
    p = np.full((T, 1), 0.45)
    p[80:160] = 0.1
    x = rng.binomial(n, p).astype(np.int32)
    prob, Pi, delta = [0.1, 0.45], hmm.sym_Pi(2, 1e-6), [0.0, 1.0]
    out = hmm.hmm_binom_dev(torch.from_numpy(x).cuda(), torch.from_numpy(n).cuda(), prob, Pi, delta)
    hmm.status_dev()
    assert hmm.binom2_general_rerun() == 1
    g = out["gamma"].cpu().numpy()
    for b in (0, 39):
        logb = [[P.dbinom_log(float(a), float(c), q) for q in prob] for a, c in zip(x[:, b], n[:, b])]
        assert np.abs(g[:, :, b].T - P.mp_posteriors(logb, Pi, delta, dps=80)).max() < 1e-12
        assert (out["y"][:, b].cpu().numpy() == oracle.viterbi_binom(x[:, b], n[:, b], prob, Pi, delta)).all()


def test_tiled_layout_matches_gene_major(oracle):
    """Warp-tiled resident layout (hb_tile_dev / hb_hmm_*_tiled_dev / hb_untile_dev): bit-identical to
    the gene-major entry points, for B not a multiple of 32, both chain families, and vs the oracle."""
    import torch
    from honeybadger_b200 import hmm
    O = oracle
    G, N = 2600, 333
    x, seg, _ = synth.expression(G, N, seed=31)
    Pi, delta, mean = O.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0], [-synth.DEV, 0.0, synth.DEV]
    xd = hmm.to_rows(torch.from_numpy(x).cuda())
    sd = hmm.chain_sd_dev(xd)
    ref = hmm.hmm_norm_dev(xd, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    xt = hmm.tile_rows(xd)
    assert xt.shape == ((N + 31) // 32, G, 32)
    assert torch.equal(hmm.untile_rows(xt, N), xd)
    out = hmm.hmm_norm_tiled_dev(xt, N, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    hmm.status_dev()
    assert torch.equal(hmm.untile_rows(out["y"], N), ref["y"])
    assert torch.equal(hmm.untile_rows(out["gamma"], N), ref["gamma"])
    assert torch.equal(out["ll"], ref["ll"])
    yo, go, llo, rc = O.fbv_norm_batch(x, seg, mean, sd.cpu().numpy(), Pi, delta)
    assert rc == 0 and (hmm.untile_rows(out["y"], N).cpu().numpy() == yo).all()
    assert np.abs(hmm.untile_rows(out["gamma"], N).cpu().numpy() - go).max() < 1e-9
    # general parameters take the literal kernel on the tiled layout as well
    Pig = np.array([[0.98, 0.015, 0.005], [0.001, 0.998, 0.001], [0.02, 0.01, 0.97]])
    og = hmm.hmm_norm_tiled_dev(xt, N, sd, [-1.4, 0.1, 0.9], Pig, [0.2, 0.5, 0.3], seg_off=seg)
    rg = hmm.hmm_norm_dev(xd, sd, [-1.4, 0.1, 0.9], Pig, [0.2, 0.5, 0.3], seg_off=seg)
    assert torch.equal(hmm.untile_rows(og["y"], N), rg["y"])
    assert torch.equal(hmm.untile_rows(og["gamma"], N), rg["gamma"])

    r_maf, n_sc, seg2, _ = synth.allele(4000, 205, seed=33, deep=0.002)
    Pi2, prob = O.sym_Pi(2, 1e-6), [0.1, 0.45]
    rd, nd = hmm.to_rows(torch.from_numpy(r_maf).cuda()), hmm.to_rows(torch.from_numpy(n_sc).cuda())
    refb = hmm.hmm_binom_dev(rd, nd, prob, Pi2, [0.0, 1.0], seg_off=seg2, viterbi=True, gamma=True, ll=True)
    ob = hmm.hmm_binom_tiled_dev(hmm.tile_rows(rd), hmm.tile_rows(nd), 205, prob, Pi2, [0.0, 1.0],
                                 seg_off=seg2, viterbi=True, gamma=True, ll=True)
    hmm.status_dev()
    assert torch.equal(hmm.untile_rows(ob["y"], 205), refb["y"])
    assert torch.equal(hmm.untile_rows(ob["gamma"], 205), refb["gamma"])
    assert torch.equal(ob["ll"], refb["ll"])
    yo, go, llo, rc = O.fbv_binom_batch(r_maf, n_sc, seg2, prob, Pi2, [0.0, 1.0])
    assert rc == 0 and (refb["y"].cpu().numpy() == yo).all()
    # packed counts: one (n << 16) | x word per position, same bits out
    xn = hmm.pack_counts(hmm.tile_rows(rd), hmm.tile_rows(nd))
    op = hmm.hmm_binom_tiled_dev(xn, None, 205, prob, Pi2, [0.0, 1.0], seg_off=seg2, viterbi=True, gamma=True,
                                 ll=True)
    hmm.status_dev()
    assert torch.equal(hmm.untile_rows(op["y"], 205), refb["y"])          # (padding lanes are not written)
    assert torch.equal(hmm.untile_rows(op["gamma"], 205), refb["gamma"])
    assert torch.equal(op["ll"], ob["ll"])
    big = nd.clone()
    big[5, 7] = 70000
    with pytest.raises(Exception, match="16 bits"):
        hmm.pack_counts(rd.contiguous(), big.contiguous())


def test_random_small_cases_on_the_device(oracle):
    """60 random small problems per chain family through the device entry points (both layouts, packed
    counts, every kernel mode): lengths 1..120 incl. chunk boundaries, ragged / empty segments, coarse
    observations with exact ties.  Paths equal the oracle's; posteriors within 1e-9."""
    import torch
    from honeybadger_b200 import _lib, hmm
    O = oracle
    lib = _lib.load()
    rng = np.random.default_rng(2024)
    try:
        for it in range(60):
            T, B = int(rng.integers(1, 121)), int(rng.integers(1, 70))
            cuts = sorted(rng.integers(0, T + 1, size=int(rng.integers(0, 4))).tolist())
            seg = np.array([0] + cuts + [T], dtype=np.int64)
            live = np.zeros(T, bool)
            for a, b in zip(seg[:-1], seg[1:]):
                live[a:b] = True
            t = float(rng.choice([1e-8, 1e-6, 1e-3, 0.05]))
            tiled = bool(rng.integers(0, 2))
            # ---- Gaussian
            m = float(rng.uniform(0.3, 2.0))
            x = rng.standard_normal((T, B)) * 1.7
            x[T // 3: 2 * T // 3, : B // 2] += m
            if it % 3 == 0:
                x = np.round(x, 0)
            sd = rng.uniform(0.7, 2.5, B)
            delta = [(0.0, 1.0, 0.0), (0.2, 0.5, 0.3)][it % 2]
            mean, Pi = [-m, 0.0, m], O.sym_Pi(3, t)
            yo, go, _, rc = O.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
            if rc == 0:
                _lib.check(lib.hb_set_norm3_mode([3, 1, 2][it % 3]))
                xd, sdd = hmm.to_rows(torch.from_numpy(x).cuda()), torch.from_numpy(sd).cuda()
                if tiled:
                    out = hmm.hmm_norm_tiled_dev(hmm.tile_rows(xd), B, sdd, mean, Pi, delta, seg_off=seg)
                    y, g = hmm.untile_rows(out["y"], B), hmm.untile_rows(out["gamma"], B)
                else:
                    out = hmm.hmm_norm_dev(xd, sdd, mean, Pi, delta, seg_off=seg)
                    y, g = out["y"], out["gamma"]
                hmm.status_dev()
                assert (y.cpu().numpy()[live] == yo[live]).all(), f"case {it}"
                assert np.abs(g.cpu().numpy()[:, live] - go[:, live]).max() < 1e-9, f"case {it}"
            # ---- binomial
            n = np.where(rng.random((T, B)) < 0.3, rng.integers(1, 12, (T, B)), 0).astype(np.int32)
            p = np.full((T, B), 0.45)
            p[T // 4: T // 2, : B // 2] = 0.1
            r = rng.binomial(n, p).astype(np.int32)
            d2 = [(0.0, 1.0), (0.3, 0.7)][it % 2]
            prob, Pi2 = [0.1, 0.45], O.sym_Pi(2, t)
            yo, go, _, rc = O.fbv_binom_batch(r, n, seg, prob, Pi2, d2)
            if rc == 0:
                _lib.check(lib.hb_set_binom2_mode(it % 2))
                rd, nd = hmm.to_rows(torch.from_numpy(r).cuda()), hmm.to_rows(torch.from_numpy(n).cuda())
                if tiled:
                    rt, nt = hmm.tile_rows(rd), hmm.tile_rows(nd)
                    if it % 4 == 1:
                        out = hmm.hmm_binom_tiled_dev(hmm.pack_counts(rt, nt), None, B, prob, Pi2, d2, seg_off=seg)
                    else:
                        out = hmm.hmm_binom_tiled_dev(rt, nt, B, prob, Pi2, d2, seg_off=seg)
                    y, g = hmm.untile_rows(out["y"], B), hmm.untile_rows(out["gamma"], B)
                else:
                    out = hmm.hmm_binom_dev(rd, nd, prob, Pi2, d2, seg_off=seg)
                    y, g = out["y"], out["gamma"]
                hmm.status_dev()
                assert (y.cpu().numpy()[live] == yo[live]).all(), f"case {it}"
                assert np.abs(g.cpu().numpy()[:, live] - go[:, live]).max() < 1e-9, f"case {it}"
    finally:
        lib.hb_set_norm3_mode(0)
        lib.hb_set_binom2_mode(0)


def test_host_entry_pinned_and_pageable_buffers_agree():
    """hb_hmm_*_host stages ordinary host memory (R vectors, numpy) through its own pinned bounce
    buffers and copies page-locked buffers directly; both must give the same bytes, over several
    sub-batches and a ragged last one."""
    import ctypes as C

    import torch

    from honeybadger_b200 import _lib, hmm

    lib = _lib.load()
    T, B = 700, 333
    rng = np.random.default_rng(5)
    seg = np.array([0, 250, 253, 700], dtype=np.int64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    tp = lambda t: C.c_void_p(t.data_ptr())

    def pinned(a):
        t = torch.from_numpy(a.copy()).pin_memory()
        assert t.is_pinned()
        return t

    # Gaussian
    x = rng.standard_normal((B, T))
    x[:40, 300:500] += 1.0
    sd = np.full(B, 0.7)
    Pi, mean, delta = hmm.sym_Pi(3, 1e-6), np.array([-1.0, 0.0, 1.0]), np.array([0.0, 1.0, 0.0])
    outs = []
    for pin in (False, True):
        y = np.full((B, T), -1, np.int32)
        g = np.full((3, B, T), np.nan)
        l = np.full((3, B), np.nan)
        if pin:
            xb, yb, gb, lb = pinned(x), pinned(y), pinned(g), pinned(l)
            rc = lib.hb_hmm_norm_host(tp(xb), T, B, vp(seg), 3, vp(mean), vp(sd), 0, vp(Pi), vp(delta), 7,
                                      tp(yb), tp(gb), tp(lb))
            outs.append((yb.numpy().copy(), gb.numpy().copy(), lb.numpy().copy()))
        else:
            rc = lib.hb_hmm_norm_host(vp(x), T, B, vp(seg), 3, vp(mean), vp(sd), 0, vp(Pi), vp(delta), 7,
                                      vp(y), vp(g), vp(l))
            outs.append((y, g, l))
        _lib.check(rc)
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    assert outs[0][0].min() >= 1 and np.isfinite(outs[0][1]).all()

    # binomial
    n = rng.integers(0, 40, (B, T)).astype(np.int32)
    k = rng.binomial(n, 0.3).astype(np.int32)
    P2, prob, d2 = hmm.sym_Pi(2, 1e-5), np.array([0.5, 0.1]), np.array([0.5, 0.5])
    outs = []
    for pin in (False, True):
        y = np.full((B, T), -1, np.int32)
        g = np.full((2, B, T), np.nan)
        l = np.full((3, B), np.nan)
        if pin:
            kb, nb, yb, gb, lb = pinned(k), pinned(n), pinned(y), pinned(g), pinned(l)
            rc = lib.hb_hmm_binom_host(tp(kb), tp(nb), T, B, vp(seg), 3, vp(prob), 0.0, vp(P2), vp(d2), 7,
                                       tp(yb), tp(gb), tp(lb))
            outs.append((yb.numpy().copy(), gb.numpy().copy(), lb.numpy().copy()))
        else:
            rc = lib.hb_hmm_binom_host(vp(k), vp(n), T, B, vp(seg), 3, vp(prob), 0.0, vp(P2), vp(d2), 7,
                                       vp(y), vp(g), vp(l))
            outs.append((y, g, l))
        _lib.check(rc)
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    assert outs[0][0].min() >= 1 and np.isfinite(outs[0][1]).all()


def test_ll_alone_runs_the_forward_sweep_only(oracle):
    """HB_WANT_LL without HB_WANT_GAMMA (header: 'alone: the forward sweep only'): same LL as the full run on
    every kernel family, with and without Viterbi; no posterior buffer is needed or touched."""
    import torch
    from honeybadger_b200 import _lib, hmm
    lib = _lib.load()
    x, seg, _ = synth.expression(2600, 150, seed=5)
    xd = hmm.to_rows(torch.from_numpy(x).cuda())
    sd = hmm.chain_sd_dev(xd)
    mean, Pi, delta = [-synth.DEV, 0.0, synth.DEV], hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0]
    full = hmm.hmm_norm_dev(xd, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    _, _, llo, _ = oracle.fbv_norm_batch(x, seg, mean, sd.cpu().numpy(), Pi, delta)
    for mode in (0, 1):          # speculate-and-verify kernel (padded rows), literal kernel
        lib.hb_set_norm3_mode(mode)
        try:
            for vit in (False, True):
                o = hmm.hmm_norm_dev(xd, sd, mean, Pi, delta, seg_off=seg, viterbi=vit, gamma=False, ll=True)
                hmm.status_dev()
                assert "gamma" not in o
                assert np.allclose(o["ll"].cpu().numpy(), llo, rtol=1e-10, atol=0)
                if vit:
                    assert torch.equal(o["y"], full["y"])
        finally:
            lib.hb_set_norm3_mode(0)
    xt = hmm.tile_rows(xd)
    o = hmm.hmm_norm_tiled_dev(xt, 150, sd, mean, Pi, delta, seg_off=seg, viterbi=False, gamma=False, ll=True)
    assert torch.equal(o["ll"], full["ll"])
    # binomial: lean kernel and general kernel, host entry point (pipeline) as well
    r, n, seg2, _ = synth.allele(5000, 100, seed=6)
    Pi2, prob = hmm.sym_Pi(2, 1e-6), [0.1, 0.45]
    yo, _, llo2, _ = oracle.fbv_binom_batch(r, n, seg2, prob, Pi2, [0.0, 1.0])
    for mode in (0, 1):
        lib.hb_set_binom2_mode(mode)
        try:
            res = hmm.hmm_binom_host(r, n, prob, Pi2, [0.0, 1.0], seg_off=seg2, viterbi=True, gamma=False, ll=True)
            assert (res["y"] == yo).all() and np.allclose(res["ll"].T, llo2, rtol=1e-10, atol=1e-12)
            res = hmm.hmm_binom_host(r, n, prob, Pi2, [0.0, 1.0], seg_off=seg2, viterbi=False, gamma=False, ll=True)
            assert np.allclose(res["ll"].T, llo2, rtol=1e-10, atol=1e-12)
        finally:
            lib.hb_set_binom2_mode(0)
    # a pooled-shaped single chain (long-chain route): LL alone
    one = hmm.hmm_norm_host(x[:, 0], mean, float(sd[0]), Pi, delta, viterbi=False, gamma=False, ll=True)
    _, _, ll1, _ = oracle.fbv_norm_batch(x[:, :1], np.array([0, x.shape[0]]), mean, sd[:1].cpu().numpy(), Pi, delta)
    assert np.allclose(one["ll"], ll1.ravel(), rtol=1e-9, atol=0)


def test_two_streams_share_the_workspace_safely(oracle):
    """Calls alternating between two non-blocking streams (different inputs, different segment tables) must
    each produce the single-stream result: the per-device scratch is handed over with an event."""
    import torch
    from honeybadger_b200 import hmm
    mean, Pi, delta = [-synth.DEV, 0.0, synth.DEV], hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0]
    xa, sega, _ = synth.expression(4000, 700, seed=11)
    xb, _, _ = synth.expression(3000, 900, seed=12)
    segb = np.array([0, 1000, 3000], dtype=np.int64)
    xa_d, xb_d = hmm.to_rows(torch.from_numpy(xa).cuda()), hmm.to_rows(torch.from_numpy(xb).cuda())
    sda, sdb = hmm.chain_sd_dev(xa_d), hmm.chain_sd_dev(xb_d)
    ra = hmm.hmm_norm_dev(xa_d, sda, mean, Pi, delta, seg_off=sega)
    ra = {k: v.clone() for k, v in ra.items()}
    rb = hmm.hmm_norm_dev(xb_d, sdb, mean, Pi, delta, seg_off=segb)
    rb = {k: v.clone() for k, v in rb.items()}
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for i in range(6):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2):
            if i % 2 == 0:
                outs.append(("a", hmm.hmm_norm_dev(xa_d, sda, mean, Pi, delta, seg_off=sega)))
            else:
                outs.append(("b", hmm.hmm_norm_dev(xb_d, sdb, mean, Pi, delta, seg_off=segb)))
    torch.cuda.synchronize()
    for tag, o in outs:
        ref = ra if tag == "a" else rb
        assert torch.equal(o["y"], ref["y"]) and torch.equal(o["gamma"], ref["gamma"])


def test_betabinom_emissions_reduce_to_the_binomial_and_match_the_oracle(oracle):
    """north_star's beta-binomial allele emissions (an extension: the reference is binomial,
    R/HoneyBADGER.R:1409).  rho = 1e-12: the binomial's exact states, posteriors within 1e-9 of the binomial's;
    rho = 1e-6 and 0.05: states exactly the oracle's at the same rho, posteriors within 1e-9; both kernels
    (lean + general), tiled packed layout as well; counts beyond the emission table are an error, not a
    silent binomial."""
    import torch
    from honeybadger_b200 import _lib, hmm
    lib = _lib.load()
    r, n, seg, _ = synth.allele(6000, 96, seed=3)
    Pi, delta, prob = hmm.sym_Pi(2, 1e-6), [0.0, 1.0], [0.1, 0.45]
    y0, g0, ll0, rc = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta)
    assert rc == 0
    assert int(n.max()) <= 511
    for rho in (1e-12, 1e-6, 0.05):
        yo, go, llo, rc = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta, rho=rho)
        assert rc == 0
        for mode in (0, 1):
            lib.hb_set_binom2_mode(mode)
            try:
                res = hmm.hmm_binom_host(r, n, prob, Pi, delta, seg_off=seg, rho=rho, viterbi=True, gamma=True,
                                         ll=True)
            finally:
                lib.hb_set_binom2_mode(0)
            assert (res["y"] == yo).all(), f"rho={rho} mode={mode}: {(res['y'] != yo).sum()} states differ"
            assert np.abs(res["gamma"].transpose(2, 0, 1) - go).max() < 1e-9
            assert np.allclose(res["ll"].T, llo, rtol=1e-10, atol=1e-12)
            if rho == 1e-12:
                assert (res["y"] == y0).all()
                assert np.abs(res["gamma"].transpose(2, 0, 1) - g0).max() < 1e-9
        if rho == 0.05:
            assert np.abs(go - g0).max() > 1e-3  # the dispersion changes the posteriors
    # the resident form bench.py uses: packed counts, warp-tiled
    rd, nd = torch.from_numpy(r).cuda(), torch.from_numpy(n).cuda()
    xn = hmm.pack_counts(hmm.tile_rows(rd), hmm.tile_rows(nd))
    yo, go, _, _ = oracle.fbv_binom_batch(r, n, seg, prob, Pi, delta, rho=0.05)
    o = hmm.hmm_binom_tiled_dev(xn, None, r.shape[1], prob, Pi, delta, seg_off=seg, rho=0.05)
    hmm.status_dev()
    assert (hmm.untile_rows(o["y"], r.shape[1]).cpu().numpy() == yo).all()
    assert np.abs(hmm.untile_rows(o["gamma"], r.shape[1]).cpu().numpy() - go).max() < 1e-9
    # counts above the table: error under rho > 0; fine once the table is enlarged
    r2, n2 = r.copy(), n.copy()
    n2[100, 3], r2[100, 3] = 900, 130
    with pytest.raises(_lib.HBError):
        hmm.hmm_binom_host(r2, n2, prob, Pi, delta, seg_off=seg, rho=0.05)
    _lib.check(lib.hb_set_binom_lut_max(1245))
    try:
        res = hmm.hmm_binom_host(r2, n2, prob, Pi, delta, seg_off=seg, rho=0.05, viterbi=True, gamma=True)
        yo, go, _, _ = oracle.fbv_binom_batch(r2, n2, seg, prob, Pi, delta, rho=0.05)
        assert (res["y"] == yo).all() and np.abs(res["gamma"].transpose(2, 0, 1) - go).max() < 1e-9
    finally:
        _lib.check(lib.hb_set_binom_lut_max(511))
