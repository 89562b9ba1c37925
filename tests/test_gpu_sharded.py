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


def allele_case(G=16000, N=220, seed=5):
    """counts with the deleted clone's cells first (the vote counts only the group of the first cell,
    R/HoneyBADGER.R:1424-1426), labels = nested cuts of the clone tree"""
    from honeybadger_b200 import synth
    _, n, seg, clone = synth.allele(G, N, seed=seed)
    # single-cell depth: one read per covered SNP (the pooled statistic of R/HoneyBADGER.R:1404-1405 counts CELLS
    # with a lesser-allele read; with several reads per cell it saturates and the binomial chain sees nothing)
    rng = np.random.default_rng(seed + 1)
    n = (n > 0).astype(np.int32)
    p = np.full((G, 3), 0.45)
    p[seg[9]:seg[10], 0] = 0.1
    p[seg[12]:seg[13], 1] = 0.1
    r = rng.binomial(n, p[:, clone]).astype(np.int32)
    order = np.argsort(clone, kind="stable")
    r, n, clone = np.ascontiguousarray(r[:, order]), np.ascontiguousarray(n[:, order]), clone[order]
    labels = [np.zeros(N, np.int64), (clone != 0).astype(np.int64), clone.astype(np.int64)]
    return r, n, labels


def labels_fn_for(labels):
    """labels_fn of ShardedGexpDriver from full-population label arrays (nested cuts of the clone tree)"""
    def fn(cells, heights):
        return [np.asarray(labels[h - 1])[cells] for h in heights]
    return fn


def compare_drivers(a, b):
    assert a.levels == b.levels and len(a.bound_genes_final) == len(b.bound_genes_final) >= 2
    for p, q in zip(a.bound_genes_final, b.bound_genes_final):
        assert np.array_equal(p, q)
    for p, q in zip(a.regions, b.regions):
        assert np.array_equal(p["cells"], q["cells"]) and p["kind"] == q["kind"] and p["depth"] == q["depth"]
    assert np.array_equal(a.bound_genes_old, b.bound_genes_old)
    return True


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


def _worker(rank, world, port, q, backend="gloo"):
    try:
        import torch
        import torch.distributed as dist
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from honeybadger_b200 import _lib, dist as hd, hmm, synth
        devno = rank if backend == "nccl" else 0    # gloo: the ranks share cuda:0 (single-GPU boxes)
        torch.cuda.set_device(devno)
        _lib.check(_lib.load().hb_set_device(devno))
        x, labels = make_case()
        G, N = x.shape
        ref = hd.sharded_gexp_level(hmm.to_rows(torch.from_numpy(x).cuda()), labels, N, synth.DEV)
        ref_drv = hd.ShardedGexpDriver(hmm.to_rows(torch.from_numpy(x).cuda()), N, synth.DEV, labels_fn_for(labels)).run()
        ra, na, labs_a = allele_case()
        ref_a = hd.sharded_allele_level(hmm.to_rows(torch.from_numpy(ra).cuda()), hmm.to_rows(torch.from_numpy(na).cuda()),
                                        labs_a, ra.shape[1])
        # grouping on one rank: labels, the distance matrix, and the recursion that computes its own groups
        from honeybadger_b200 import honeybadger as hbm
        xfull = hmm.to_rows(torch.from_numpy(x).cuda())
        ref_labs, ref_D = hd.sharded_ward_groups(xfull, [N], [1, 2, 3, 4, 5], return_matrix=True)
        dev_labs = hbm.ward_groups_dev(xfull, [1, 2, 3, 4, 5], 101)
        assert all(np.array_equal(a, b) for a, b in zip(ref_labs, dev_labs))
        ref_auto = hd.ShardedGexpDriver(xfull, N, synth.DEV).run()
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        if backend == "nccl":
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", devno))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
        lo, hi = hd.shard_range(N, rank, world)
        xl = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(x[:, lo:hi])).cuda())
        got = hd.sharded_gexp_level(xl, [l[lo:hi] for l in labels], N, synth.DEV)
        ok = compare_levels(got, ref) and len(got["runs"]) > 0 and len(got["collectives"]) == 4
        ok = ok and got["g1"].size > 10
        # the whole recursion on sharded cells (sub-populations unevenly spread over the ranks) == one rank
        drv = hd.ShardedGexpDriver(xl, N, synth.DEV, labels_fn_for(labels)).run()
        ok = ok and compare_drivers(drv, ref_drv)
        # the grouping itself on sharded cells: dist split by blocks over the ranks, every entry of the assembled
        # matrix bit-identical to the single-GPU one, hence the same tree and the same labels
        labs, Dm = hd.sharded_ward_groups(xl, hd.shard_sizes(N, world), [1, 2, 3, 4, 5], return_matrix=True)
        ok = ok and all(np.array_equal(a, b) for a, b in zip(labs, ref_labs))
        if rank == 0:
            ok = ok and bool(torch.equal(Dm.view(torch.int64), ref_D.view(torch.int64)))
        auto = hd.ShardedGexpDriver(xl, N, synth.DEV).run()
        ok = ok and compare_drivers(auto, ref_auto)
        # the allele level: integer exchanges only, so EVERYTHING is exact at any rank count
        r, n, labs = allele_case()
        lo, hi = hd.shard_range(r.shape[1], rank, world)
        rd = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(r[:, lo:hi])).cuda())
        nd = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(n[:, lo:hi])).cuda())
        ga = hd.sharded_allele_level(rd, nd, [l[lo:hi] for l in labs], r.shape[1])
        dist.barrier()
        checks = {"collectives": len(ga["collectives"]) == 3, "y": np.array_equal(ga["y"], ref_a["y"]),
                  "mafl": np.array_equal(ga["mafl"], ref_a["mafl"]), "sizel": np.array_equal(ga["sizel"], ref_a["sizel"]),
                  "nruns": len(ga.get("runs", [])) == len(ref_a.get("runs", [])) and len(ga.get("runs", [])) > 0}
        if checks["nruns"]:
            checks["runs"] = all(np.array_equal(p, q) for p, q in zip(ga["runs"], ref_a["runs"]))
            checks["prob"] = bool(np.abs(ga["prob"] - ref_a["prob"]).max() < 1e-12)
            checks["g1"] = np.array_equal(ga["g1"], ref_a["g1"])
        bad = [k for k, v in checks.items() if not v]
        if bad:
            raise AssertionError(f"allele level differs: {bad}; collectives {ga['collectives']}; runs "
                                 f"{len(ga.get('runs', []))} vs {len(ref_a.get('runs', []))}")
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:]))


@pytest.mark.parametrize("backend", ["gloo", "nccl"])
def test_sharded_level_equals_single_rank_level(backend):
    """levels (expression + allele) and the whole expression recursion, sharded over two ranks vs one rank;
    gloo: both ranks on cuda:0; nccl: one GPU per rank (skipped on a single-GPU box)"""
    import torch
    import torch.multiprocessing as mp
    if backend == "nccl" and torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q, backend)) for r in range(2)]
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


def _worker_groups(rank, world, port, q, sizes):
    try:
        import torch
        import torch.distributed as dist
        sys.path.insert(0, ROOT)
        from honeybadger_b200 import _lib, dist as hd, hmm
        torch.cuda.set_device(0)
        _lib.check(_lib.load().hb_set_device(0))
        N = int(sum(sizes))
        x, _ = make_case(G=1500, N=N, seed=21)
        ref_labs, ref_D = hd.sharded_ward_groups(hmm.to_rows(torch.from_numpy(x).cuda()), [N], [1, 2, 3, 4], k=31,
                                                 return_matrix=True)
        ra, na, _ = allele_case(G=6000, N=N, seed=23)
        ref_labs_a, ref_Da = hd.sharded_ward_groups(
            None, [N], [1, 2, 3], k=31, return_matrix=True,
            counts=(hmm.to_rows(torch.from_numpy(ra).cuda()), hmm.to_rows(torch.from_numpy(na).cuda())))
        assert bool(torch.isfinite(ref_Da).all())
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        lo, hi = int(offs[rank]), int(offs[rank + 1])
        xl = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(x[:, lo:hi])).cuda()) if hi > lo else \
            torch.zeros((x.shape[0], 0), dtype=torch.float64, device="cuda")
        labs, Dm = hd.sharded_ward_groups(xl, sizes, [1, 2, 3, 4], k=31, return_matrix=True)
        ok = all(np.array_equal(a, b) for a, b in zip(labs, ref_labs))
        if rank == 0:
            ok = ok and bool(torch.equal(Dm.view(torch.int64), ref_D.view(torch.int64)))
        # the allele form: r / n with NaN wherever a cell has no coverage (the masked path of the distance kernel)
        cl = (hmm.to_rows(torch.from_numpy(np.ascontiguousarray(ra[:, lo:hi])).cuda()),
              hmm.to_rows(torch.from_numpy(np.ascontiguousarray(na[:, lo:hi])).cuda())) if hi > lo else \
            (torch.zeros((ra.shape[0], 0), dtype=torch.int32, device="cuda"),) * 2
        labs_a, Da = hd.sharded_ward_groups(None, sizes, [1, 2, 3], k=31, return_matrix=True, counts=cl)
        ok = ok and all(np.array_equal(a, b) for a, b in zip(labs_a, ref_labs_a))
        if rank == 0:
            ok = ok and bool(torch.equal(Da.view(torch.int64), ref_Da.view(torch.int64)))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()[-1500:]))


@pytest.mark.parametrize("sizes", [[40, 41, 39], [5, 0, 60], [1, 70, 2, 33], [64, 1]])
def test_sharded_grouping_with_uneven_shards(sizes):
    """`dist` split by blocks over 2-4 ranks (odd and even rings, empty and one-cell shards): the assembled matrix is
    bit-identical to one rank's, the labels are the same."""
    import torch.multiprocessing as mp
    world = len(sizes)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_groups, args=(r, world, port, q, sizes)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(60)
    for rank, ok, msg in res:
        assert ok, f"rank {rank}: {msg}"
