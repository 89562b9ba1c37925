"""CPU, world_size 2 over gloo: the N > 1 host logic (sharding, count all-reduce, posterior gather)."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from honeybadger_b200 import dist as hd
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # same data on every rank; each rank keeps its own columns
    G, N = 400, 37
    n = np.where(rng.random((G, N)) < 0.2, rng.integers(1, 5, (G, N)), 0)
    r = rng.binomial(n, 0.45)
    lo, hi = hd.shard_range(N, rank, world)
    cells = np.arange(lo, hi)
    mafl, sizel = hd.allreduce_pooled_counts(O.pool_count_pos(r[:, lo:hi].astype(float), cells - lo),
                                             O.pool_count_pos(n[:, lo:hi].astype(float), cells - lo))
    full_m = O.pool_count_pos(r.astype(float), np.arange(N))
    full_s = O.pool_count_pos(n.astype(float), np.arange(N))
    ok = (mafl.numpy() == full_m).all() and (sizel.numpy() == full_s).all()
    # the pooled chain is then identical on every rank, whatever the GPU count
    y = O.viterbi_binom(mafl.numpy(), sizel.numpy(), [0.1, 0.45], O.sym_Pi(2, 1e-6), [0, 1])
    ok = ok and (y == O.viterbi_binom(full_m, full_s, [0.1, 0.45], O.sym_Pi(2, 1e-6), [0, 1])).all()
    post = torch.arange(3 * N, dtype=torch.float64).reshape(3, N)
    got = hd.allgather_region_posteriors(post[:, lo:hi].contiguous(), N)
    ok = ok and torch.equal(got, post)
    # pooled expression mean: the 80-bit accumulator relay (device call replaced by np.longdouble,
    # which is the x87 extended format on x86-64) reproduces R's mean over ALL cells on every rank
    xg = rng.standard_normal((G, N)) * 2.5
    memb = (rng.random((3, N)) < 0.5).astype(np.uint8)
    memb[:, 0] = 1
    sizes = memb.sum(1)

    def accumulate(x, member, phase, state, mean80, out=None):
        st = state.numpy().view(np.longdouble).reshape(3, -1)     # one block of genes of the pipelined relay
        mn = mean80.numpy().view(np.longdouble).reshape(3, -1)
        x = x.numpy() if torch.is_tensor(x) else x
        if phase in (0, 2):
            for k in range(3):
                for c in np.nonzero(member[k])[0]:
                    st[k] += (x[:, c].astype(np.longdouble) - mn[k]) if phase == 2 else x[:, c].astype(np.longdouble)
        elif phase == 1:
            for k in range(3):
                mn[k] = st[k] / np.longdouble(sizes[k])
        else:
            for k in range(3):
                out[k] = torch.from_numpy((mn[k] + st[k] / np.longdouble(sizes[k])).astype(np.float64))

    pm = hd.pooled_mean_relay(torch.from_numpy(xg[:, lo:hi]), memb[:, lo:hi], sizes, accumulate=accumulate, chunks=5)
    for k in range(3):
        ok = ok and (pm[k].numpy() == O.pool_mean(xg, np.nonzero(memb[k])[0])).all()
    v = hd.merge_votes(np.full(G, rank + 1))
    ok = ok and int(v[0]) == sum(range(1, world + 1))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_exchange_over_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
