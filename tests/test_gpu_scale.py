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


def test_c4_expression_full_size_properties(oracle):
    """10k cells x 20k genes: posteriors are distributions, states valid, re-running is idempotent,
    sampled chains equal the oracle, segment results do not depend on the batch around them."""
    import torch
    from honeybadger_b200 import hmm
    G, N = 20000, 10000
    x, seg = _expr_device(G, N, 21)
    sd = hmm.chain_sd_dev(x)
    mean, Pi, delta = [-synth.DEV, 0.0, synth.DEV], hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0]
    out = hmm.hmm_norm_dev(x, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    hmm.status_dev()
    y, g = out["y"], out["gamma"]
    assert int(y.min()) >= 1 and int(y.max()) <= 3
    assert float((g.sum(0) - 1).abs().max()) < 1e-12
    assert float(g.min()) >= 0.0
    chk = (int(y.to(torch.int64).sum()), float(g[0].sum()), float(g[2].sum()))
    out2 = hmm.hmm_norm_dev(x, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    assert torch.equal(out2["y"], y) and torch.equal(out2["gamma"], g)  # deterministic
    nre = hmm.norm3_rerun_chains()
    assert nre < 220000 // 500, f"{nre} of 220000 chains needed the exact fallback"
    assert chk == (int(out2["y"].to(torch.int64).sum()), float(out2["gamma"][0].sum()), float(out2["gamma"][2].sum()))
    cols = np.array([0, 1, 4999, 5000, 9998, 9999, 1234, 7777])
    xs = x[:, cols].cpu().numpy()
    sds = sd[cols].cpu().numpy()
    yo, go, llo, rc = oracle.fbv_norm_batch(xs, seg, mean, sds, Pi, delta)
    assert rc == 0 and (y[:, cols].cpu().numpy() == yo).all()
    assert np.abs(g[:, :, cols].cpu().numpy() - go).max() < 1e-9
    assert np.allclose(out["ll"][:, cols].cpu().numpy(), llo, rtol=1e-10, atol=1e-9)
    # a sub-batch gives the same chains (no cross-chain coupling, any B / alignment)
    sub = x[:, 100:357].contiguous()
    o3 = hmm.hmm_norm_dev(sub, sd[100:357].contiguous(), mean, Pi, delta, seg_off=seg)
    assert torch.equal(o3["y"], y[:, 100:357]) and torch.equal(o3["gamma"], g[:, :, 100:357])
    # amplified / deleted segments are called in the carrier cells
    assert float((y[seg[6]:seg[7], : N // 2] == 3).double().mean()) > 0.5
    assert float((y[seg[9]:seg[10], : N // 2] == 1).double().mean()) > 0.5


def test_c3_allele_large_properties(oracle):
    """2000 cells x 200k SNPs (C3's chain length): same properties for the binomial chains."""
    import torch
    from honeybadger_b200 import hmm
    G, N = 200000, 2000
    r_np, n_np, seg, _ = synth.allele(G, 64, seed=31)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    n = ((torch.rand((G, N), device="cuda", generator=g) < 0.108).to(torch.int32)
         * torch.randint(1, 9, (G, N), device="cuda", generator=g, dtype=torch.int32))
    r = (torch.rand((G, N), device="cuda", generator=g) * (n + 1)).floor().to(torch.int32).clamp(max=n)
    r[:, :64] = torch.from_numpy(r_np).cuda()
    n[:, :64] = torch.from_numpy(n_np).cuda()
    Pi, delta, prob = hmm.sym_Pi(2, 1e-6), [0.0, 1.0], [0.1, 0.45]
    out = hmm.hmm_binom_dev(r, n, prob, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    hmm.status_dev()
    y, gm = out["y"], out["gamma"]
    assert int(y.min()) >= 1 and int(y.max()) <= 2
    assert float((gm.sum(0) - 1).abs().max()) < 1e-12
    yo, go, llo, rc = oracle.fbv_binom_batch(r_np, n_np, seg, prob, Pi, delta)
    assert rc == 0 and (y[:, :64].cpu().numpy() == yo).all()
    assert np.abs(gm[:, :, :64].cpu().numpy() - go).max() < 1e-9
    assert np.allclose(out["ll"][:, :64].cpu().numpy(), llo, rtol=1e-10, atol=1e-9)
    # single concatenated chain (the reference's shape, nseg = 1) on a few cells
    one = np.array([0, G], dtype=np.int64)
    o1 = hmm.hmm_binom_dev(r[:, :32].contiguous(), n[:, :32].contiguous(), prob, Pi, delta, seg_off=one)
    y1, g1, _, rc = oracle.fbv_binom_batch(r_np[:, :32], n_np[:, :32], one, prob, Pi, delta)
    assert rc == 0 and (o1["y"].cpu().numpy() == y1).all()
    assert np.abs(o1["gamma"].cpu().numpy() - g1).max() < 1e-9


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
