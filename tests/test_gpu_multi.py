"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the exchanges of SURVEY.md §8e on real devices —
cells sharded over two GPUs, pooled expression mean by accumulator relay, pooled allele counts by
int32 all-reduce, posterior all-gather; results identical to one GPU holding every cell."""
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


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from honeybadger_b200 import _lib, dist as hd, hmm, synth
    from honeybadger_b200 import honeybadger as hb
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    _lib.check(_lib.load().hb_set_device(rank))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = True
    rng = np.random.default_rng(3)  # same data on every rank, each rank uploads its own columns
    G, N, K = 900, 301, 6
    x = rng.standard_normal((G, N)) * 2.5
    member = (rng.random((K, N)) < 0.45).astype(np.uint8)
    member[0] = 1
    member[:, 0] = 1
    sizes = member.sum(1).astype(np.int32)
    lo, hi = hd.shard_range(N, rank, world)
    xl = hmm.to_rows(torch.from_numpy(np.ascontiguousarray(x[:, lo:hi])).cuda())
    pm = hd.pooled_mean_relay(xl, member[:, lo:hi], sizes)
    # reference: one GPU with all the cells
    xd = hmm.to_rows(torch.from_numpy(x).cuda())
    full = torch.empty((K, G), dtype=torch.float64, device="cuda")
    md = torch.from_numpy(member).cuda()
    lib = _lib.load()
    _lib.check(lib.hb_pool_mean_dev(xd.data_ptr(), xd.stride(0), G, N, md.data_ptr(), K, full.data_ptr(), None))
    ok = ok and bool(torch.equal(pm, full))
    # allele counts: per-rank partials + all-reduce == counts over all cells
    n = np.where(rng.random((G, N)) < 0.2, rng.integers(1, 5, (G, N)), 0).astype(np.int32)
    r = rng.binomial(n, 0.45).astype(np.int32)
    memb = np.ones((1, N), np.uint8)
    y_l, mafl_l, sizel_l = hb.allele_round_dev(torch.from_numpy(np.ascontiguousarray(r[:, lo:hi])).cuda(),
                                               torch.from_numpy(np.ascontiguousarray(n[:, lo:hi])).cuda(),
                                               memb[:, lo:hi], 0.1, 0.45, 1e-6)
    mafl, sizel = hd.allreduce_pooled_counts(mafl_l[0], sizel_l[0])
    _, mafl_f, sizel_f = hb.allele_round_dev(torch.from_numpy(r).cuda(), torch.from_numpy(n).cuda(), memb,
                                             0.1, 0.45, 1e-6)
    ok = ok and (mafl.cpu().numpy() == mafl_f[0]).all() and (sizel.cpu().numpy() == sizel_f[0]).all()
    # the whole pooled round on sharded cells == the single-GPU round (states, counts), on every rank
    memb3 = np.stack([np.ones(N, bool), np.arange(N) % 2 == 0, np.arange(N) % 2 == 1]).astype(np.uint8)
    nn = np.where(rng.random((2600, N)) < 0.3, rng.integers(1, 4, (2600, N)), 0).astype(np.int32)
    pp = np.full((2600, 1), 0.45)
    pp[700:1100] = 0.02
    rr = rng.binomial(nn, pp).astype(np.int32)
    y_s, mafl_s, sizel_s = hd.pooled_allele_round(torch.from_numpy(np.ascontiguousarray(rr[:, lo:hi])).cuda(),
                                                  torch.from_numpy(np.ascontiguousarray(nn[:, lo:hi])).cuda(),
                                                  memb3[:, lo:hi])
    y_1, mafl_1, sizel_1 = hb.allele_round_dev(torch.from_numpy(rr).cuda(), torch.from_numpy(nn).cuda(), memb3,
                                               0.1, 0.45, 1e-6)
    ok = ok and np.array_equal(y_s, y_1) and np.array_equal(mafl_s, mafl_1) and np.array_equal(sizel_s, sizel_1)
    ok = ok and bool((y_1[0, 750:1050] == 1).all())
    post = torch.arange(4 * N, dtype=torch.float64, device="cuda").reshape(4, N)
    got = hd.allgather_region_posteriors(post[:, lo:hi].contiguous(), N)
    ok = ok and bool(torch.equal(got, post))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_gpu_exchanges_over_nccl():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
