"""GPU: size-independent properties at BASELINE.json's full sizes, and spot parity on sampled chains."""
from __future__ import annotations

import numpy as np
import pytest

from honeybadger_b200 import synth

pytestmark = pytest.mark.gpu


def _expr_device(G, N, seed):
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    seg = synth.segments(G)
    x = torch.randn((G, N), dtype=torch.float64, device="cuda", generator=g)
    sig = torch.empty(N, dtype=torch.float64, device="cuda").uniform_(2.0, 3.1, generator=g)
    x.mul_(sig)
    x[seg[6]:seg[7], : N // 2] += synth.DEV
    x[seg[9]:seg[10], : N // 2] -= synth.DEV
    return x, seg


def _tiled_cols(t, cols):
    """Columns `cols` of a warp-tiled [ntile, T, 32] (or [planes, ntile, T, 32]) tensor -> numpy [T, n] / [planes, T, n]."""
    import torch
    c = torch.as_tensor(np.asarray(cols), device=t.device)
    if t.dim() == 4:
        return np.stack([_tiled_cols(p, cols) for p in t])
    return t[c // 32, :, c % 32].t().cpu().numpy()  # advanced indices around a slice: [n, T]


def _valid_parts(t, N):
    """The real chains of a warp-tiled tensor ([.., ntile, T, 32]): all full tiles, and the filled lanes of
    the last one (lanes past B in the last tile are padding the kernels never write)."""
    rem = N % 32
    if rem == 0:
        return [t]
    return [t[..., :-1, :, :], t[..., -1:, :, :rem]]


def test_c4_expression_full_size_on_the_timed_path(oracle):
    """BASELINE config 4, expression half, exactly as bench.py runs it: 10k cells x 20k genes x 22 segments
    from bench.device_inputs, warp-tiled resident layout, hb_hmm_norm_tiled_dev in auto mode.  Asserts that the
    speculate-and-verify kernel is the one that ran, then compares 512 sampled cells (11 264 chains) PLUS every
    chain the kernel re-ran exactly with the oracle: states exact, posteriors 1e-9.  Size-independent
    properties on the whole output."""
    import torch
    import bench
    from honeybadger_b200 import hmm
    G, N = 20000, 10000
    dev = torch.device("cuda", torch.cuda.current_device())
    d = bench.device_inputs("expr", N, G, seed=100, dev=dev)
    x, seg = d["x"], d["seg"]
    sd = hmm.chain_sd_dev(x)
    mean, Pi, delta = [-synth.DEV, 0.0, synth.DEV], hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0]
    xt = hmm.tile_rows(x)
    out = hmm.hmm_norm_tiled_dev(xt, N, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    hmm.status_dev()
    assert hmm.last_kernels()["norm3_fast"] == 1, "auto mode did not pick the speculate-and-verify kernel"
    nre = hmm.norm3_rerun_chains()
    flagged = hmm.norm3_flagged_chains(N)
    assert nre == len(flagged) and nre < 220000 // 500, f"{nre} of 220000 chains needed the exact fallback"
    y, g = out["y"], out["gamma"]
    for yp, gp in zip(_valid_parts(y, N), _valid_parts(g, N)):
        assert int(yp.min()) >= 1 and int(yp.max()) <= 3
        assert float((gp.sum(0) - 1).abs().max()) < 1e-12 and float(gp.min()) >= 0.0
    out2 = hmm.hmm_norm_tiled_dev(xt, N, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    for a, b in zip(_valid_parts(out2["y"], N) + _valid_parts(out2["gamma"], N), _valid_parts(y, N) + _valid_parts(g, N)):
        assert torch.equal(a, b)  # deterministic
    # forcing EVERY chain through the literal re-run gives the same states (mode 2), and so does the literal kernel
    lib = __import__("honeybadger_b200._lib", fromlist=["load"]).load()
    try:
        lib.hb_set_norm3_mode(2)
        o2 = hmm.hmm_norm_tiled_dev(xt, N, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=False)
        assert hmm.norm3_rerun_chains() == 220000
        assert all(torch.equal(a, b) for a, b in zip(_valid_parts(o2["y"], N), _valid_parts(y, N)))
        lib.hb_set_norm3_mode(1)
        o1 = hmm.hmm_norm_tiled_dev(xt, N, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True)
        assert hmm.last_kernels()["norm3_fast"] == 0
        assert all(torch.equal(a, b) for a, b in zip(_valid_parts(o1["y"], N), _valid_parts(y, N)))
        assert all(float((a - b).abs().max()) < 1e-10 for a, b in zip(_valid_parts(o1["gamma"], N), _valid_parts(g, N)))
    finally:
        lib.hb_set_norm3_mode(0)
    rng = np.random.default_rng(4)
    cols = np.unique(np.concatenate([rng.choice(N, 512, replace=False), [0, 31, 32, N - 1], flagged[:, 1]]))
    xs = x[:, torch.as_tensor(cols, device=dev)].cpu().numpy()
    sds = sd[torch.as_tensor(cols, device=dev)].cpu().numpy()
    yo, go, llo, rc = oracle.fbv_norm_batch(xs, seg, mean, sds, Pi, delta)
    assert rc == 0
    ys = _tiled_cols(y, cols)
    assert (ys == yo).all(), f"{(ys != yo).sum()} states differ on {len(cols)} sampled cells"
    assert np.abs(_tiled_cols(g, cols) - go).max() < 1e-9
    assert np.allclose(out["ll"][:, torch.as_tensor(cols, device=dev)].cpu().numpy(), llo, rtol=1e-10, atol=1e-9)
    # amplified / deleted segments are called in the carrier cells (clone labels are inside device_inputs:
    # at least a third of all cells carry them)
    yg = hmm.untile_rows(y, N)
    assert float((yg[seg[6]:seg[7]] == 3).double().mean()) > 0.3
    assert float((yg[seg[9]:seg[10]] == 1).double().mean()) > 0.3


def test_c2_expression_at_its_stated_size(oracle):
    """BASELINE config 2: 2k cells x 10k genes, fp64, Viterbi parity — ALL 2000 chains against the oracle, in
    the reference's own shape (nseg = 1: one 10k-position chain per cell, R/HoneyBADGER.R:1117-1121; longer
    than the speculative kernel accepts, so the literal kernel) and per chromosome (nseg = 22, speculative
    kernel on the tiled layout)."""
    import torch
    from honeybadger_b200 import hmm
    G, N = 10000, 2000
    x, seg22, _ = synth.expression(G, N, seed=2)
    xd = hmm.to_rows(torch.from_numpy(x).cuda())
    sd = hmm.chain_sd_dev(xd)
    sdh = sd.cpu().numpy()
    mean, Pi, delta = [-synth.DEV, 0.0, synth.DEV], hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0]
    xt = hmm.tile_rows(xd)
    for seg, fast in ((np.array([0, G], dtype=np.int64), 0), (seg22, 1)):
        out = hmm.hmm_norm_tiled_dev(xt, N, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True)
        hmm.status_dev()
        assert hmm.last_kernels()["norm3_fast"] == fast
        yo, go, _, rc = oracle.fbv_norm_batch(x, seg, mean, sdh, Pi, delta)
        assert rc == 0
        y = hmm.untile_rows(out["y"], N).cpu().numpy()
        assert (y == yo).all(), f"nseg={len(seg) - 1}: {(y != yo).sum()} of {y.size} states differ"
        g = hmm.untile_rows(out["gamma"], N).cpu().numpy()
        assert np.abs(g - go).max() < 1e-9
        # gene-major padded rows give the same answer as the tiled layout
        o2 = hmm.hmm_norm_dev(xd, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=False)
        assert (o2["y"].cpu().numpy() == yo).all()


def test_c3_allele_full_size_on_the_timed_path(oracle):
    """BASELINE config 3 as bench.py runs its allele half: 10k cells x 200k SNPs from bench.device_inputs,
    PACKED 16-bit counts on the warp-tiled layout, hb_hmm_binom_packed_tiled_dev (lean kernel + gated general
    kernel).  512 sampled cells (11 264 chains, 1e8 steps) against the oracle: states exact, posteriors 1e-9."""
    import torch
    import bench
    from honeybadger_b200 import hmm
    G, N = 200000, 10000
    dev = torch.device("cuda", torch.cuda.current_device())
    d = bench.device_inputs("allele", N, G, seed=300, dev=dev)
    seg = d["seg"]
    rng = np.random.default_rng(5)
    cols = np.unique(np.concatenate([rng.choice(N, 512, replace=False), [0, 31, 32, N - 1]]))
    ct = torch.as_tensor(cols, device=dev)
    r_np, n_np = d["r"][:, ct].cpu().numpy(), d["n"][:, ct].cpu().numpy()
    xn = hmm.pack_counts(hmm.tile_rows(d["r"]), hmm.tile_rows(d["n"]))
    del d
    torch.cuda.empty_cache()
    Pi, delta, prob = hmm.sym_Pi(2, 1e-6), [0.0, 1.0], [0.1, 0.45]
    out = hmm.hmm_binom_tiled_dev(xn, None, N, prob, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    hmm.status_dev()
    assert hmm.last_kernels()["binom2_lean"] == 1 and hmm.binom2_general_rerun() == 0
    y, gm = out["y"], out["gamma"]
    for yp, gp in zip(_valid_parts(y, N), _valid_parts(gm, N)):
        assert int(yp.min()) >= 1 and int(yp.max()) <= 2
        assert float((gp[0] + gp[1] - 1).abs().max()) < 1e-12 and float(gp.min()) >= 0.0
    yo, go, llo, rc = oracle.fbv_binom_batch(r_np, n_np, seg, prob, Pi, delta)
    assert rc == 0
    ys = _tiled_cols(y, cols)
    assert (ys == yo).all(), f"{(ys != yo).sum()} states differ on {len(cols)} sampled cells"
    assert np.abs(_tiled_cols(gm, cols) - go).max() < 1e-9
    assert np.allclose(out["ll"][:, ct].cpu().numpy(), llo, rtol=1e-10, atol=1e-9)
    assert float((y == 1).double().mean()) > 0.005  # deleted segments are called


def test_ragged_and_tiny_shapes(oracle):
    """B not a multiple of the block, 1 cell, chains shorter than a chunk, empty segments."""
    from honeybadger_b200 import hmm
    rng = np.random.default_rng(3)
    Pi, delta, mean = hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0], [-1.0, 0.0, 1.0]
    for T, B, seg in [(37, 1, [0, 37]), (9, 131, [0, 2, 2, 9]), (64, 129, [0, 8, 16, 17 + 3, 64]),
                      (5, 3, [0, 5])]:
        x = rng.standard_normal((T, B)) * 1.5
        sd = np.full(B, 1.3)
        s = np.array(seg, dtype=np.int64)
        r = hmm.hmm_norm_host(x, mean, sd, Pi, delta, seg_off=s, viterbi=True, gamma=True, ll=True)
        yo, go, llo, rc = oracle.fbv_norm_batch(x, s, mean, sd, Pi, delta)
        assert rc == 0 and (r["y"] == yo).all()
        assert np.abs(r["gamma"].transpose(2, 0, 1) - go).max() < 1e-9


@pytest.mark.parametrize("T", [4096, 4097, 9000])
def test_norm_long_single_chains_cross_the_kernel_switch(oracle, T):
    """One concatenated chain per cell (the reference's shape): up to 4096 positions the
    speculate-and-verify kernel runs, beyond that the literal one; both equal the oracle."""
    from honeybadger_b200 import hmm
    x, _, _ = synth.expression(T, 40, seed=T)
    sd = np.array([oracle.r_sd(x[:, b]) for b in range(x.shape[1])])
    Pi, delta, mean = hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0], [-synth.DEV, 0.0, synth.DEV]
    seg = np.array([0, T], dtype=np.int64)
    r = hmm.hmm_norm_host(x, mean, sd, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    nre = hmm.norm3_rerun_chains()
    yo, go, llo, rc = oracle.fbv_norm_batch(x, seg, mean, sd, Pi, delta)
    assert rc == 0 and (r["y"] == yo).all()
    assert np.abs(r["gamma"].transpose(2, 0, 1) - go).max() < 1e-9
    assert np.allclose(r["ll"].T, llo, rtol=1e-10, atol=0)
    assert nre <= 40 and (T <= 4096 or nre == 0)
