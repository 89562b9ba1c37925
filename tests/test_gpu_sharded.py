"""GPU: one recursion level of the expression driver with the CELLS SHARDED over two ranks
(`dist.sharded_gexp_level`) against the same level on one rank holding every cell.  The two ranks share
cuda:0 and talk over gloo, so this runs on a single-GPU box (the NCCL form of the same code runs in
tests/test_gpu_multi.py on >= 2 GPUs and in `bench.py --gpus N`, which prints the same comparison)."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_case(G=3000, N=241, seed=4):
    from honeybadger_b200 import synth
    x, _, clone = synth.expression(G, N, seed=seed)
    idx = np.arange(N)
    labels = [np.zeros(N, np.int64), (clone == 0).astype(np.int64), clone.astype(np.int64),
              clone * 2 + (idx % 2) * (clone == 0), clone * 3 + (idx % 3) * (clone == 0)]
    return x, labels


def compare_levels(a, b, tol=1e-12):
    """sharded result `a` vs single-rank result `b`: exact except the retest posteriors (one fp64 all-reduce)."""
    assert a["K"] == b["K"]
    assert np.array_equal(a["pooled"], b["pooled"]), "pooled means differ (the relay must be exact)"
    assert np.array_equal(a["sd"], b["sd"]) and np.array_equal(a["y"], b["y"])
    assert len(a["runs"]) == len(b["runs"]) and all(np.array_equal(p, q) for p, q in zip(a["runs"], b["runs"]))
    assert a["kinds"] == b["kinds"]
    if a["runs"]:
        assert np.abs(a["prob"] - b["prob"]).max() < tol
        assert a["chosen"] == b["chosen"] and np.array_equal(a["g1"], b["g1"]) and np.array_equal(a["g2"], b["g2"])
    return True


def _worker(rank, world, port, q):
    try:
        import torch
        import torch.distributed as dist
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from honeybadger_b200 import _lib, dist as hd, hmm, synth
        torch.cuda.set_device(0)
        _lib.check(_lib.load().hb_set_device(0))
        x, labels = make_case()
        G, N = x.shape
        ref = hd.sharded_gexp_level(hmm.to_rows(torch.from_numpy(x).cuda()), labels, N, synth.DEV)
        ra, na, _, clone_a = synth.allele(4000, 150, seed=5)
        labs_a = [np.zeros(150, np.int64), (clone_a != clone_a[0]).astype(np.int64), clone_a.astype(np.int64)]
        ref_a = hd.sharded_allele_level(hmm.to_rows(torch.from_numpy(ra).cuda()), hmm.to_rows(torch.from_numpy(na).cuda()),
                                        labs_a, 150)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lo, hi = hd.shard_range(N, rank, world)
        xl = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(x[:, lo:hi])).cuda())
        got = hd.sharded_gexp_level(xl, [l[lo:hi] for l in labels], N, synth.DEV)
        ok = compare_levels(got, ref) and len(got["runs"]) > 0 and len(got["collectives"]) == 4
        ok = ok and got["g1"].size > 10
        # the allele level: integer exchanges only, so EVERYTHING is exact at any rank count
        r, n, _, clone = synth.allele(4000, 150, seed=5)
        idx = np.arange(150)
        labs = [np.zeros(150, np.int64), (clone != clone[0]).astype(np.int64), clone.astype(np.int64)]
        lo, hi = hd.shard_range(150, rank, world)
        rd = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(r[:, lo:hi])).cuda())
        nd = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(n[:, lo:hi])).cuda())
        ga = hd.sharded_allele_level(rd, nd, [l[lo:hi] for l in labs], 150)
        dist.barrier()
        ok = ok and len(ga["collectives"]) == 3 and np.array_equal(ga["y"], ref_a["y"]) \
            and np.array_equal(ga["mafl"], ref_a["mafl"]) and np.array_equal(ga["sizel"], ref_a["sizel"]) \
            and len(ga["runs"]) == len(ref_a["runs"]) and len(ga["runs"]) > 0 \
            and all(np.array_equal(p, q) for p, q in zip(ga["runs"], ref_a["runs"])) \
            and np.array_equal(ga["prob"], ref_a["prob"]) and np.array_equal(ga["g1"], ref_a["g1"])
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:]))


def test_sharded_level_equals_single_rank_level():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(60)
    for rank, ok, msg in res:
        assert ok, f"rank {rank}: {msg}"


def test_single_rank_level_matches_the_resident_round(oracle):
    """world = 1: the level's pooled vectors, sd and states are those of hb_gexp_pooled_viterbi_dev (and of the
    oracle), its runs those of runs_from_vote on the votes."""
    import torch
    from honeybadger_b200 import dist as hd, hmm, synth
    from honeybadger_b200 import honeybadger as hb
    x, labels = make_case(G=2000, N=120, seed=9)
    G, N = x.shape
    xd = hmm.to_rows(torch.from_numpy(x).cuda())
    lvl = hd.sharded_gexp_level(xd, labels, N, synth.DEV)
    member, _ = hb._membership([l + 1 for l in labels], N)
    # same groups in (height, id) order as _membership's (height, first appearance) order? compare as sets of rows
    y, px, psd = hb.gexp_round_dev(xd, member, [-synth.DEV, 0.0, synth.DEV], 1e-6)
    key = lambda a: a.tobytes()  # noqa: E731
    assert sorted(map(key, lvl["pooled"])) == sorted(map(key, px))
    assert sorted(map(key, lvl["y"])) == sorted(map(key, y))
    for k in range(lvl["K"]):
        yo = oracle.viterbi_norm(lvl["pooled"][k], [-synth.DEV, 0.0, synth.DEV], [lvl["sd"][k]] * 3,
                                 oracle.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0])
        assert np.array_equal(yo, lvl["y"][k])
