"""GPU: red zones around every output buffer of the chain kernels (a stand-in for compute-sanitizer memcheck,
which this pool refuses to run — profiles/r2_sanitizer.txt).  Each output lives inside a larger allocation whose
surroundings (and, for padded rows, the padding columns) hold a sentinel; after the launch the sentinel must be
intact.  Catches out-of-bounds WRITES of the kernels at awkward shapes (partial tiles, partial blocks, ragged
segments, one-position chains); out-of-bounds reads are covered by construction tests elsewhere
(test_pool_count_on_column_slice_views: views that end exactly at the end of an allocation)."""
from __future__ import annotations

import numpy as np
import pytest

from honeybadger_b200 import synth

pytestmark = pytest.mark.gpu
GUARD = 4096  # elements on each side


def guarded(shape, dtype, sentinel):
    """A contiguous CUDA tensor of `shape` embedded in a sentinel-filled buffer -> (view, check)."""
    import torch
    n = int(np.prod(shape))
    esz = torch.empty(0, dtype=dtype).element_size()
    pad = GUARD + (-GUARD * esz) % 256 // esz          # keep the view 256-byte aligned
    buf = torch.full((pad + n + GUARD,), sentinel, dtype=dtype, device="cuda")
    view = buf[pad:pad + n].view(*shape)
    assert view.data_ptr() % 16 == 0

    def check(what):
        assert bool((buf[:pad] == sentinel).all()) and bool((buf[pad + n:] == sentinel).all()), \
            f"{what}: a kernel wrote outside its output buffer"
    return view, check


@pytest.mark.parametrize("T,B,segs", [(257, 33, [0, 257]), (1000, 130, [0, 1, 9, 9, 500, 1000]),
                                      (96, 1, [0, 96]), (2500, 257, [0, 700, 2500])])
def test_tiled_gaussian_outputs_stay_inside_their_buffers(T, B, segs):
    import torch
    from honeybadger_b200 import hmm
    x, _, _ = synth.expression(max(T, 40), B, seed=T + B)
    x = np.ascontiguousarray(x[:T])
    xd = hmm.to_rows(torch.from_numpy(x).cuda())
    sd = hmm.chain_sd_dev(xd)
    xt = hmm.tile_rows(xd)
    ntile = (B + 31) // 32
    seg = np.array(segs, dtype=np.int64)
    # (a proper initial distribution: with R's delta = (0, 1, 0) a one-position chain is HiddenMarkov's
    # "Problems With Underflow" error, which is not what this test is about)
    mean, Pi, delta = [-synth.DEV, 0.0, synth.DEV], hmm.sym_Pi(3, 1e-6), [0.2, 0.5, 0.3]
    ref = hmm.hmm_norm_tiled_dev(xt, B, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True)
    y, cy = guarded((ntile, T, 32), torch.uint8, 0xAB)
    g, cg = guarded((3, ntile, T, 32), torch.float64, -7.25)
    ll, cl = guarded((seg.size - 1, B), torch.float64, -7.25)
    for mode in (0, 1):     # speculate-and-verify and literal kernels
        __import__("honeybadger_b200._lib", fromlist=["load"]).load().hb_set_norm3_mode(mode)
        try:
            hmm.hmm_norm_tiled_dev(xt, B, sd, mean, Pi, delta, seg_off=seg, viterbi=True, gamma=True, ll=True,
                                   out={"y": y, "gamma": g, "ll": ll})
            hmm.status_dev()
        finally:
            __import__("honeybadger_b200._lib", fromlist=["load"]).load().hb_set_norm3_mode(0)
        cy("y"), cg("gamma"), cl("ll")
        rem = B % 32 or 32
        assert torch.equal(y[:-1], ref["y"][:-1]) and torch.equal(y[-1, :, :rem], ref["y"][-1, :, :rem])


@pytest.mark.parametrize("T,B", [(300, 100), (1025, 129), (40, 5)])
def test_gene_major_padding_columns_are_not_written(T, B):
    """Padded rows (ld > B): the kernels must leave the padding columns of y / gamma alone, both families."""
    import torch
    from honeybadger_b200 import hmm
    x, _, _ = synth.expression(max(T, 40), B, seed=T)
    xd = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(x[:T])).cuda())
    sd = hmm.chain_sd_dev(xd)
    ld = xd.stride(0)
    ybuf = torch.full((T, ld), 0xAB, dtype=torch.uint8, device="cuda")
    gbuf = torch.full((3, T, ld), -7.25, dtype=torch.float64, device="cuda")
    hmm.hmm_norm_dev(xd, sd, [-synth.DEV, 0.0, synth.DEV], hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0],
                     out={"y": ybuf[:, :B], "gamma": gbuf[:, :, :B]})
    hmm.status_dev()
    assert bool((ybuf[:, B:] == 0xAB).all()) and bool((gbuf[:, :, B:] == -7.25).all())
    assert int(ybuf[:, :B].min()) >= 1 and int(ybuf[:, :B].max()) <= 3
    r, n, _, _ = synth.allele(max(T, 64), B, seed=T)
    rd = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(r[:T])).cuda())
    nd = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(n[:T])).cuda())
    ybuf.fill_(0xAB)
    g2 = torch.full((2, T, ld), -7.25, dtype=torch.float64, device="cuda")
    hmm.hmm_binom_dev(rd, nd, [0.1, 0.45], hmm.sym_Pi(2, 1e-6), [0.0, 1.0], out={"y": ybuf[:, :B], "gamma": g2[:, :, :B]})
    hmm.status_dev()
    assert bool((ybuf[:, B:] == 0xAB).all()) and bool((g2[:, :, B:] == -7.25).all())


def test_packed_tiled_binomial_and_resident_kernel_outputs_stay_inside():
    import torch
    from honeybadger_b200 import hmm
    T, B = 3000, 75
    r, n, seg, _ = synth.allele(T, B, seed=2)
    xn = hmm.pack_counts(hmm.tile_rows(torch.from_numpy(r).cuda()), hmm.tile_rows(torch.from_numpy(n).cuda()))
    ntile = (B + 31) // 32
    y, cy = guarded((ntile, T, 32), torch.uint8, 0xAB)
    g, cg = guarded((2, ntile, T, 32), torch.float64, -7.25)
    hmm.hmm_binom_tiled_dev(xn, None, B, [0.1, 0.45], hmm.sym_Pi(2, 1e-6), [0.0, 1.0], seg_off=seg, out={"y": y, "gamma": g})
    hmm.status_dev()
    cy("y"), cg("gamma")
    # chain-resident posteriors: [3, B, T] written by blocks of 1 / 2 / 4 warps
    x, segx, _ = synth.expression(5000, 21, seed=3)
    for so in (segx, np.array([0, 1500, 5000], dtype=np.int64), np.array([0, 5000], dtype=np.int64)):
        xd = hmm.to_rows(torch.from_numpy(x).cuda())
        sd = hmm.chain_sd_dev(xd)
        x_cm = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
        gc, cgc = guarded((3, 21, 5000), torch.float64, -7.25)
        hmm.forwardback_norm_cm_dev(x_cm, sd, [-synth.DEV, 0.0, synth.DEV], hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0], seg_off=so,
                                    out=gc)
        torch.cuda.synchronize()
        cgc("resident gamma")
        assert float(gc.min()) >= 0.0 and float((gc.sum(0) - 1).abs().max()) < 1e-12
