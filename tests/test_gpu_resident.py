"""GPU: the chain-resident forward-backward kernel (csrc/hb_norm3_resident.cuh) against the oracle and against
the thread-per-chain kernel — same posteriors (1e-9 vs the oracle, 1e-11 between the kernels)."""
from __future__ import annotations

import numpy as np
import pytest

from honeybadger_b200 import synth

pytestmark = pytest.mark.gpu


def _run(x, seg, mean, Pi, delta, oracle):
    import torch
    from honeybadger_b200 import hmm
    xd = hmm.to_rows(torch.from_numpy(x).cuda())
    sd = hmm.chain_sd_dev(xd)
    x_cm = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    g = hmm.forwardback_norm_cm_dev(x_cm, sd, mean, Pi, delta, seg_off=seg)
    assert hmm.fbr_gate() == 0
    _, go, _, rc = oracle.fbv_norm_batch(x, seg, mean, sd.cpu().numpy(), Pi, delta)
    assert rc == 0
    got = g.permute(0, 2, 1).cpu().numpy()          # [3, B, T] -> [3, T, B]
    assert np.abs(got - go).max() < 1e-9
    ref = hmm.hmm_norm_dev(xd, sd, mean, Pi, delta, seg_off=seg, viterbi=False, gamma=True)["gamma"].cpu().numpy()
    assert np.abs(got - ref).max() < 1e-11
    assert np.abs(got.sum(0) - 1).max() < 1e-12 and got.min() >= 0.0
    return got


def test_resident_posteriors_22_segments(oracle):
    from honeybadger_b200 import hmm
    x, seg, _ = synth.expression(3000, 203, seed=2)
    _run(x, seg, [-synth.DEV, 0.0, synth.DEV], hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0], oracle)


@pytest.mark.parametrize("T", [2, 5, 33, 97, 1024, 1025, 2049, 4100, 6000])
def test_resident_posteriors_single_chains_of_every_block_shape(oracle, T):
    """One segment of T positions: 1, 2 and 4 warps per chain, pieces of 3 .. 47 positions, chains shorter than
    a warp (empty pieces), the first-position rule (delta, no transition)."""
    from honeybadger_b200 import hmm
    x, _, _ = synth.expression(max(T, 40), 37, seed=T)
    x = np.ascontiguousarray(x[:T])
    if T >= 40:
        x[T // 3: T // 2, :12] += 1.3
    _run(x, np.array([0, T], dtype=np.int64), [-synth.DEV, 0.0, synth.DEV], hmm.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0],
         oracle)


def test_resident_ragged_segments_and_general_delta(oracle):
    """Empty and one-position segments, a non-degenerate initial distribution, shifted (still symmetric) means."""
    from honeybadger_b200 import hmm
    rng = np.random.default_rng(5)
    x = rng.standard_normal((700, 65)) * 1.4 + 0.3
    x[200:420, :30] += 1.1
    seg = np.array([0, 1, 1, 40, 333, 334, 700], dtype=np.int64)
    _run(x, seg, [-0.7, 0.3, 1.3], hmm.sym_Pi(3, 1e-4), [0.2, 0.5, 0.3], oracle)


def test_resident_rejects_what_it_does_not_serve():
    import torch
    from honeybadger_b200 import _lib, hmm
    x = torch.zeros((4, 50), dtype=torch.float64, device="cuda")
    sd = torch.ones(4, dtype=torch.float64, device="cuda")
    with pytest.raises(_lib.HBError):   # asymmetric means
        hmm.forwardback_norm_cm_dev(x, sd, [-1.0, 0.0, 2.0], hmm.sym_Pi(3, 1e-6), [0, 1, 0])
    big = torch.zeros((1, 7000), dtype=torch.float64, device="cuda")
    with pytest.raises(_lib.HBError):   # a chain that does not fit one block's shared memory
        hmm.forwardback_norm_cm_dev(big, sd[:1], [-1.0, 0.0, 1.0], hmm.sym_Pi(3, 1e-6), [0, 1, 0])
