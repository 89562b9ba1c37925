"""Cell sharding across the GPUs of one box.

Per-cell chains are independent, so the data path needs no collective: every rank owns a contiguous
block of cells (columns) with all genes.  The only exchanges are the two small ones the driver
needs per recursion level — integer pooled counts (all-reduce, exact and invariant to the GPU
count) and per-region posteriors (all-gather).  One process per GPU, `torch.distributed` with the
NCCL backend on the box and gloo in the CPU tests; device work stays in libhb_b200.so.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_cells: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the contiguous block of cells owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_cells, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_cells: int, world: int) -> list[int]:
    return [shard_range(n_cells, r, world)[1] - shard_range(n_cells, r, world)[0] for r in range(world)]


# ---- collectives that also work on the gloo backend with CUDA tensors (the single-GPU CI path runs two
# ranks on one device over gloo): there the payload is staged through the host; NCCL moves it directly.
def _staged(t):
    import torch.distributed as dist
    return t.is_cuda and dist.get_backend() == "gloo"


def _all_reduce(t, op=None):
    import torch.distributed as dist
    op = dist.ReduceOp.SUM if op is None else op
    if _staged(t):
        c = t.cpu()
        dist.all_reduce(c, op=op)
        t.copy_(c)
    else:
        dist.all_reduce(t, op=op)


def _broadcast(t, src):
    import torch.distributed as dist
    if _staged(t):
        c = t.cpu()
        dist.broadcast(c, src=src)
        t.copy_(c)
    else:
        dist.broadcast(t, src=src)


def _send(t, dst):
    import torch.distributed as dist
    dist.send(t.cpu() if _staged(t) else t, dst=dst)


def _recv(t, src):
    import torch.distributed as dist
    if _staged(t):
        c = t.cpu()
        dist.recv(c, src=src)
        t.copy_(c)
    else:
        dist.recv(t, src=src)


def allreduce_pooled_counts(mafl_local, sizel_local):
    """Sum the per-rank partial counts of R/HoneyBADGER.R:1404-1405 (`rowSums(. > 0)` over the
    rank's cells): int32 all-reduce — exact, so pooled allele chains are identical at any GPU count."""
    import torch
    import torch.distributed as dist
    t = torch.stack([torch.as_tensor(mafl_local), torch.as_tensor(sizel_local)]).to(torch.int32)
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t[0], t[1]


def pooled_allele_round(r_local, n_local, member_local, pd_=0.1, pn=0.45, t=1e-6):
    """One pooled round of calcAlleleCnvBoundaries (R/HoneyBADGER.R:1395-1417) on cells sharded over the ranks.

    r_local, n_local: this rank's columns of the resident count matrices (gene-major int32 CUDA tensors);
    member_local [K, cells of this rank].  Every rank counts its own cells (`hb_pool_count_pos_dev`), the
    integer partials are all-reduced — exact and invariant to the GPU count — and the K pooled chains then
    run on every rank through the reference-shaped host entry point, i.e. parallel along the gene axis
    (csrc/hb_longchain.cuh; 0.75 ms for 6 x 200 000), which is cheaper than shipping states around.
    Returns (y int32 [K, G] 1-based, mafl, sizel) — identical on all ranks and to a single-GPU round."""
    import ctypes as C
    import torch
    from . import _lib, hmm
    lib = _lib.load()
    G = r_local.shape[0]
    member_local = np.ascontiguousarray(member_local, dtype=np.uint8)
    K, n_loc = member_local.shape
    md = torch.from_numpy(member_local).to(r_local.device)
    cnt = torch.zeros((2, K, G), dtype=torch.int32, device=r_local.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if n_loc > 0:
        for i, m in enumerate((r_local, n_local)):
            _lib.check(lib.hb_pool_count_pos_dev(m.data_ptr(), m.stride(0), G, n_loc, md.data_ptr(), K,
                                                 cnt[i].data_ptr(), st))
    mafl, sizel = allreduce_pooled_counts(cnt[0], cnt[1])
    mafl, sizel = mafl.cpu().numpy(), sizel.cpu().numpy()
    out = hmm.hmm_binom_host(mafl.T, sizel.T, [pd_, pn], hmm.sym_Pi(2, t), [0.0, 1.0])
    return np.ascontiguousarray(out["y"].T), mafl, sizel


def allgather_region_posteriors(local, n_cells: int, sizes=None):
    """local: [regions, cells_of_this_rank] -> [regions, n_cells] on every rank (cells in global order).

    Shards may differ in size (by one cell for the contiguous sharding; arbitrarily for the sub-populations of a
    recursion, `sizes` = cells of every rank), so blocks are padded to the largest shard for the fixed-size
    all-gather and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    local = torch.as_tensor(local)
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local
    world = dist.get_world_size()
    sizes = shard_sizes(n_cells, world) if sizes is None else [int(v) for v in sizes]
    mx = max(sizes)
    pad = torch.zeros((local.shape[0], mx), dtype=local.dtype, device=local.device)
    pad[:, :local.shape[1]] = local
    src = pad.cpu() if _staged(pad) else pad
    out = [torch.empty_like(src) for _ in range(world)]
    dist.all_gather(out, src)
    return torch.cat([o[:, :s] for o, s in zip(out, sizes)], dim=1).to(local.device)


def allgather_states(y_local, n_cells: int):
    """uint8 Viterbi states [positions, cells_of_this_rank] -> [positions, n_cells] (debug / tests;
    production keeps per-cell outputs sharded)."""
    import torch
    t = allgather_region_posteriors(torch.as_tensor(y_local).t().contiguous().t(), n_cells)
    return t


def merge_votes(vote_local):
    """Votes are integer counts per position: all-reduce(sum)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(vote_local)).to(torch.int64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


# --------------------------------------------------------------------------- pooled expression mean
# R's mean() (R/HoneyBADGER.R:1141) is an order-dependent 80-bit accumulation over the cells of a
# group, so per-rank partial sums cannot be merged afterwards.  The exact multi-GPU form is a relay:
# an 80-bit state per (group, gene) is continued by each rank with ITS cells and handed to the next
# rank (cell order = rank order), twice (R's two passes).  hb_pool_mean_relay_dev does the per-rank
# work; this module moves the state.
def _relay_phase(lib, x, member_dev, K, n_total_dev, phase, state, mean80, out):
    import ctypes as C
    import torch
    from . import _lib
    G = state.numel() // (16 * K)
    N = 0 if x is None else x.shape[1]
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.hb_pool_mean_relay_dev(
        None if x is None or N == 0 else x.data_ptr(), 0 if x is None else x.stride(0), G, N,
        None if member_dev is None or N == 0 else member_dev.data_ptr(), K,
        None if n_total_dev is None else n_total_dev.data_ptr(), phase, state.data_ptr(),
        None if mean80 is None else mean80.data_ptr(), None if out is None else out.data_ptr(), stream))


def pooled_mean_relay_shards(shards, members, group_sizes):
    """Single-process form (tests, and a GPU that holds several shards): `shards` is a list of
    gene-major CUDA tensors [G, N_i] in cell order, `members` the matching uint8 [K, N_i] arrays.
    Returns the [K, G] pooled means — identical to hb_pool_mean_dev on the concatenated matrix."""
    import torch
    from . import _lib
    lib = _lib.load()
    K, G = int(members[0].shape[0]), int(shards[0].shape[0])
    dev = shards[0].device
    n_tot = torch.as_tensor(np.asarray(group_sizes, dtype=np.int32), device=dev)
    state = torch.zeros(K * G * 16, dtype=torch.uint8, device=dev)
    mean80 = torch.zeros(K * G * 16, dtype=torch.uint8, device=dev)
    out = torch.empty((K, G), dtype=torch.float64, device=dev)
    mems = [torch.as_tensor(np.ascontiguousarray(m, dtype=np.uint8), device=dev) for m in members]
    for x, m in zip(shards, mems):
        _relay_phase(lib, x, m, K, None, 0, state, None, None)
    _relay_phase(lib, None, None, K, n_tot, 1, state, mean80, None)
    state.zero_()
    for x, m in zip(shards, mems):
        _relay_phase(lib, x, m, K, None, 2, state, mean80, None)
    _relay_phase(lib, None, None, K, n_tot, 3, state, mean80, out)
    return out


def pooled_mean_relay(x_local, member_local, group_sizes, accumulate=None, chunks=None):
    """Multi-process form: every rank calls this with its gene-major shard [G, N_local] and its
    uint8 [K, N_local] membership slice; `group_sizes` are the TOTAL group sizes.  The 16-byte
    states travel rank -> rank+1 with send/recv, the last rank closes each pass and broadcasts.
    Returns the [K, G] pooled means on every rank.  `accumulate(x, member, phase, state, mean80)` can
    replace the device call (the gloo CPU test passes a long-double emulation).

    The relay is PIPELINED over blocks of genes (accumulators of different genes are independent; only the
    cell order inside one accumulator is fixed): rank r works on block j while rank r+1 works on block j-1, so a
    pass over W ranks costs (W + J - 1) / J rank-times instead of W in the ideal case.  The longest accumulation
    of a block (the largest group's cells, hundreds of cycles per emulated 80-bit addition) is a latency floor that
    does not shrink with the block, so small blocks lose: measured (10 k cells per rank, 15 groups, whole sharded
    level) 2 ranks: no pipelining 97 ms, J = 2 112 ms, J = 16 295 ms; 8 ranks: no pipelining 428 ms, J = 8 295 ms.
    `chunks` = J; default W (at most 8) from 3 ranks up, 1 below; env HB_RELAY_CHUNKS overrides."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    K, G = int(member_local.shape[0]), int(x_local.shape[0])
    if accumulate is None:
        from . import _lib
        lib = _lib.load()
        dev = x_local.device
        mem = torch.as_tensor(np.ascontiguousarray(member_local, dtype=np.uint8), device=dev)
        n_tot = torch.as_tensor(np.asarray(group_sizes, dtype=np.int32), device=dev)

        def accumulate(x, member, phase, state, mean80, out=None):
            if phase in (0, 2):
                _relay_phase(lib, x, mem, K, None, phase, state, mean80 if phase == 2 else None, None)
            else:
                _relay_phase(lib, None, None, K, n_tot, phase, state, mean80, out)
    else:
        dev = torch.device("cpu")
    import os
    env = int(os.environ.get("HB_RELAY_CHUNKS", "0") or 0)
    J = int(chunks) if chunks else (env if env > 0 else (min(world, 8) if world > 2 else 1))
    J = max(1, min(J, G // 32 if G >= 32 else 1))
    bounds = [int(round(j * G / J)) for j in range(J + 1)]
    sizes = [bounds[j + 1] - bounds[j] for j in range(J)]
    offs = np.concatenate([[0], np.cumsum([K * g * 16 for g in sizes])]).astype(np.int64)
    state_all = torch.zeros(int(offs[-1]), dtype=torch.uint8, device=dev)
    mean_all = torch.zeros(int(offs[-1]), dtype=torch.uint8, device=dev)
    out_all = torch.empty(K * G, dtype=torch.float64, device=dev)
    ooffs = np.concatenate([[0], np.cumsum([K * g for g in sizes])]).astype(np.int64)
    st = [state_all[int(offs[j]):int(offs[j + 1])] for j in range(J)]
    mn = [mean_all[int(offs[j]):int(offs[j + 1])] for j in range(J)]
    ou = [out_all[int(ooffs[j]):int(ooffs[j + 1])].view(K, sizes[j]) for j in range(J)]
    for phase in (0, 2):
        if phase == 2:
            state_all.zero_()
        for j in range(J):
            if rank > 0:
                _recv(st[j], rank - 1)
            accumulate(x_local[bounds[j]:bounds[j + 1]], member_local, phase, st[j], mn[j])
            if rank < world - 1:
                _send(st[j], rank + 1)
            else:
                accumulate(None, None, phase + 1, st[j], mn[j], ou[j])  # close the pass on the last rank
        if world > 1:
            _broadcast(mean_all if phase == 0 else out_all, world - 1)
    return torch.cat(ou, dim=1) if J > 1 else ou[0]


# --------------------------------------------------------------------------- one sharded recursion level
def sharded_gexp_level(x_local, labels_local, n_cells, dev_mag, t=1e-6, min_num_genes=10, sigma0_local=None,
                       mean_local=None, local_sizes=None):
    """One level of HoneyBADGER$calcGexpCnvBoundaries (R/HoneyBADGER.R:1135-1298) with the CELLS SHARDED over the
    ranks of the default process group (one process per GPU; world = 1 runs the same code without exchanges).

    x_local [G, N_local]: this rank's columns of the resident gene-major matrix (genes already ordered);
    labels_local: one int array [N_local] per cut height with GLOBAL group ids (the benchmark sizes take the
    labels as input — dist + hclust over 200 k cells is 160 GB).  Exchanges, all inside this call:

      group sizes            int64 all-reduce
      pooled means           80-bit accumulator relay rank -> rank+1, twice (`pooled_mean_relay`): R's `mean` bit
                             for bit at any GPU count
      K pooled chains        no exchange: every rank runs them on the identical pooled vectors (hmm_norm_host on
                             [G, K]: the reference-shaped call)
      retest of every run    per-cell weights locally (hb_retest_gexp_dev phase 1), ONE fp64 all-reduce of the
                             shared direction evidence per run, phase 2 locally
      per-region posteriors  all-gather [runs, N_local] -> [runs, n_cells]  (`allgather_region_posteriors`)

    Returns a dict identical on every rank: y [K, G] states, pooled [K, G], sd [K], runs (index arrays), lists
    of amp / del runs, prob [runs, n_cells], chosen region index, g1 / g2 global cell indices, and the names of
    the collectives that ran."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _lib, hmm
    from .honeybadger import _vp, runs_from_vote
    lib = _lib.load()
    world = dist.get_world_size() if dist.is_initialized() else 1
    G, n_loc = x_local.shape
    device = x_local.device
    used = []
    # ---- groups with more than one cell (globally), in (height, first id) order — the same on every rank
    ids = [np.asarray(l, dtype=np.int64) for l in labels_local]
    nid = int(max((int(l.max()) if l.size else 0) for l in ids)) + 1
    if world > 1:
        mx = torch.tensor([nid], dtype=torch.int64, device=device)
        _all_reduce(mx, dist.ReduceOp.MAX)
        nid = int(mx.item())
    counts = torch.zeros((len(ids), nid), dtype=torch.int64, device=device)
    for h, l in enumerate(ids):
        if l.size:
            counts[h] += torch.bincount(torch.as_tensor(l, device=device), minlength=nid)
    if world > 1:
        _all_reduce(counts)
        used.append("all_reduce(int64 group sizes)")
    counts = counts.cpu().numpy()
    groups = [(h, g) for h in range(len(ids)) for g in range(nid) if counts[h, g] > 1]
    K = len(groups)
    out = {"collectives": used, "K": K}
    if K == 0:
        return out
    member = np.stack([(ids[h] == g) for h, g in groups]).astype(np.uint8).reshape(K, n_loc)
    sizes = np.array([counts[h, g] for h, g in groups], dtype=np.int32)
    # ---- pooled observation of every group (R/HoneyBADGER.R:1141): exact at any GPU count
    if world > 1:
        pooled = pooled_mean_relay(x_local, member, sizes)
        used.append("send/recv relay of 80-bit accumulators (x2) + broadcast")
    else:
        pooled = torch.empty((K, G), dtype=torch.float64, device=device)
        md = torch.from_numpy(member).to(device)
        _lib.check(lib.hb_pool_mean_dev(x_local.data_ptr(), x_local.stride(0), G, n_loc, md.data_ptr(), K,
                                        pooled.data_ptr(), _cur_stream()))
    sd = torch.empty(K, dtype=torch.float64, device=device)
    _lib.check(lib.hb_pooled_sd_dev(pooled.data_ptr(), G, K, sd.data_ptr(), _cur_stream()))
    pooled_h, sd_h = pooled.cpu().numpy(), sd.cpu().numpy()
    # ---- the K chains (R/HoneyBADGER.R:1144-1151), redundantly on every rank: identical inputs, identical paths
    res = hmm.hmm_norm_host(np.ascontiguousarray(pooled_h.T), [-dev_mag, 0.0, dev_mag], sd_h, hmm.sym_Pi(3, t),
                            [0.0, 1.0, 0.0])
    y = np.ascontiguousarray(res["y"].T)
    out.update({"y": y, "pooled": pooled_h, "sd": sd_h})
    amp_v, del_v = (y == 3).sum(0), (y == 1).sum(0)
    runs, kinds = [], []
    for vote, kind in ((del_v, "del"), (amp_v, "amp")):        # deletion rows first (R/HoneyBADGER.R:1245-1246)
        for idx in runs_from_vote(vote, min_num_genes, None) or []:
            runs.append(idx)
            kinds.append(kind)
    out.update({"runs": runs, "kinds": kinds})
    if not runs:
        return out
    # ---- retest of every run on the sharded cells: phase 1, one scalar all-reduce, phase 2
    st = _cur_stream()
    pa = torch.empty(n_loc, dtype=torch.float64, device=device)
    pdl = torch.empty(n_loc, dtype=torch.float64, device=device)
    d_loc = torch.zeros(1, dtype=torch.float64, device=device)
    sig = None if sigma0_local is None else torch.as_tensor(sigma0_local, dtype=torch.float64, device=device)
    rows_local = []
    for idx, kind in zip(runs, kinds):
        rows = torch.as_tensor(np.asarray(idx, dtype=np.int64), device=device)
        args = (x_local.data_ptr(), x_local.stride(0), G, n_loc, rows.data_ptr(), rows.numel(), float(dev_mag),
                None if sig is None else sig.data_ptr(), pa.data_ptr(), pdl.data_ptr(), None)
        if world > 1:
            d_loc.zero_()
            if n_loc > 0:
                _lib.check(lib.hb_retest_gexp_dev(*args, d_loc.data_ptr(), None, 1, st))
            _all_reduce(d_loc)
            if n_loc > 0:
                _lib.check(lib.hb_retest_gexp_dev(*args, None, d_loc.data_ptr(), 2, st))
        else:
            _lib.check(lib.hb_retest_gexp_dev(*args, None, None, 0, st))
        rows_local.append((pdl if kind == "del" else pa).clone())
    if world > 1:
        used.append("all_reduce(fp64 direction evidence) per run")
    prob = allgather_region_posteriors(torch.stack(rows_local), n_cells, sizes=local_sizes)
    if world > 1:
        used.append("all_gather(per-region posteriors)")
    prob = prob.cpu().numpy()
    dps = (prob > 0.75).sum(1)
    dpsi = int(np.nonzero(dps == dps.max())[0][0])
    out.update({"prob": prob, "chosen": dpsi, "g1": np.nonzero(prob[dpsi] > 0.75)[0],
                "g2": np.nonzero(prob[dpsi] <= 0.25)[0]})
    return out


def _cur_stream():
    import ctypes as C
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def sharded_allele_level(r_local, n_local, labels_local, n_cells, l_maf=None, n_bulk=None, pd_=0.1, pn=0.45, t=1e-6,
                         min_num_snps=5, trim=0.1, pe=0.1, mono=0.7):
    """One level of HoneyBADGER$calcAlleleCnvBoundaries (R/HoneyBADGER.R:1395-1551) with the CELLS SHARDED over the
    ranks of the default process group.  r_local, n_local: this rank's columns of the resident lesser-allele /
    coverage count matrices (gene-major int32 CUDA tensors); labels_local: one int array per cut height with
    GLOBAL group ids.  Exchanges:

      group sizes            int64 all-reduce
      pooled counts          mafl / sizel of every group: per-rank `rowSums(. > 0)` + ONE int32 all-reduce — integers,
                             so the K pooled chains are identical at any GPU count (R/HoneyBADGER.R:1404-1405)
      bulk counts            l.maf / n.bulk of the level (rowSums over ALL its cells, :1563-1566) ride in the same
                             all-reduce as an extra "all cells" group when the caller does not pass them
      K pooled chains        none (every rank runs them, parallel along the gene axis, hb_longchain.cuh)
      retest of every run    none: the snpModel marginal is per cell given the bulk counts (hb_retest_allele_model_dev)
      per-region posteriors  all-gather [runs, N_local] -> [runs, n_cells]

    Returns a dict identical on every rank (y, mafl, sizel, runs, prob, chosen, g1, g2, collectives)."""
    import torch
    import torch.distributed as dist
    from . import _lib, hmm
    from .honeybadger import retest_allele_model, runs_from_vote
    lib = _lib.load()
    world = dist.get_world_size() if dist.is_initialized() else 1
    G, n_loc = r_local.shape
    device = r_local.device
    used = []
    ids = [np.asarray(l, dtype=np.int64) for l in labels_local]
    nid = int(max((int(l.max()) if l.size else 0) for l in ids)) + 1
    if world > 1:
        mx = torch.tensor([nid], dtype=torch.int64, device=device)
        _all_reduce(mx, dist.ReduceOp.MAX)
        nid = int(mx.item())
    counts = torch.zeros((len(ids), nid), dtype=torch.int64, device=device)
    for h, l in enumerate(ids):
        if l.size:
            counts[h] += torch.bincount(torch.as_tensor(l, device=device), minlength=nid)
    if world > 1:
        _all_reduce(counts)
        used.append("all_reduce(int64 group sizes)")
    counts = counts.cpu().numpy()
    groups = [(h, g) for h in range(len(ids)) for g in range(nid) if counts[h, g] > 1]
    first = {}                     # vote counts only b[[1]]: the first group of each height (:1424-1426) — the
    for h in range(len(ids)):      # group of the level's FIRST cell, which lives on rank 0
        f0 = torch.tensor([int(ids[h][0]) if (n_loc and (not dist.is_initialized() or dist.get_rank() == 0)) else 0],
                          dtype=torch.int64, device=device)
        if world > 1:
            _broadcast(f0, 0)
        first[h] = int(f0.item())
    K = len(groups)
    out = {"collectives": used, "K": K}
    if K == 0:
        return out
    want_bulk = l_maf is None or n_bulk is None
    member = np.stack([(ids[h] == g) for h, g in groups] + ([np.ones(n_loc, bool)] if want_bulk else []))
    member = np.ascontiguousarray(member.astype(np.uint8).reshape(-1, n_loc))
    Kx = member.shape[0]
    md = torch.from_numpy(member).to(device)
    cnt = torch.zeros((2, Kx, G), dtype=torch.int32, device=device)
    if n_loc > 0:
        for i, m in enumerate((r_local, n_local)):
            _lib.check(lib.hb_pool_count_pos_dev(m.data_ptr(), m.stride(0), G, n_loc, md.data_ptr(), Kx,
                                                 cnt[i].data_ptr(), _cur_stream()))
    if world > 1:
        _all_reduce(cnt)
        used.append("all_reduce(int32 pooled counts)")
    cnt_h = cnt.cpu().numpy()
    mafl, sizel = cnt_h[0, :K], cnt_h[1, :K]
    if want_bulk:
        l_maf, n_bulk = cnt_h[0, K].astype(np.float64), cnt_h[1, K].astype(np.float64)
    res = hmm.hmm_binom_host(np.ascontiguousarray(mafl.T), np.ascontiguousarray(sizel.T), [pd_, pn], hmm.sym_Pi(2, t),
                             [0.0, 1.0])
    y = np.ascontiguousarray(res["y"].T)
    out.update({"y": y, "mafl": mafl, "sizel": sizel})
    vote = np.zeros(G, np.int64)
    for k, (h, g) in enumerate(groups):
        if g == first[h]:
            vote += y[k] == 1
    runs = runs_from_vote(vote, min_num_snps, trim) or []
    out["runs"] = runs
    if not runs:
        return out
    rows_local = []
    for idx in runs:
        it = torch.as_tensor(np.asarray(idx, dtype=np.int64), device=device)
        rows_local.append(retest_allele_model(r_local.index_select(0, it), n_local.index_select(0, it),
                                              np.asarray(l_maf)[idx], np.asarray(n_bulk)[idx], pe=pe, mono=mono,
                                              keep_on_device=True) if n_loc > 0 else
                          torch.zeros(0, dtype=torch.float64, device=device))
    prob = allgather_region_posteriors(torch.stack(rows_local), n_cells)
    if world > 1:
        used.append("all_gather(per-region posteriors)")
    prob = prob.cpu().numpy()
    dps = (prob > 0.75).sum(1)
    dpsi = int(np.nonzero(dps == dps.max())[0][0]) if len(runs) > 1 else 0
    out.update({"prob": prob, "chosen": dpsi, "g1": np.nonzero(prob[dpsi] > 0.75)[0],
                "g2": np.nonzero(prob[dpsi] <= 0.25)[0]})
    return out


def _pair_owner(r, q, world):
    """Which rank computes the block (cells of r) x (cells of q) of the distance matrix: the symmetric pair
    {r, q} is computed once, by the rank from which the other is at most half a ring ahead.  (Exactly opposite
    ranks of an even ring share their block, see _block_rows.)"""
    if r == q:
        return r
    d = (q - r) % world
    if 2 * d < world:
        return r
    if 2 * d > world:
        return q
    return min(r, q)   # even world, exactly opposite: the lower rank


def _block_rows(r, q, world, sizes):
    """What rank r computes of the pair {r, q}: ("rows", a, b) = rows [a, b) of the block (r, q), every column;
    ("cols", a, b) = the block (r, q) restricted to columns [a, b) of q's cells; None = nothing (q computes it).
    Opposite ranks of an even ring split their block in two so that every rank carries the same number of pairs:
    the lower rank takes the first h rows of (lo, hi), the higher rank the columns [h, n_lo) of (hi, lo)."""
    if r == q:
        return ("rows", 0, sizes[r])
    d = (q - r) % world
    if 2 * d != world:
        return ("rows", 0, sizes[r]) if 2 * d < world else None
    lo, hi = min(r, q), max(r, q)
    h = (sizes[lo] // 2) & ~1                      # even: the sub-matrix stays 16-byte aligned
    if h < 2 or sizes[lo] - h < 2 or sizes[hi] < 2:   # too small to split: the lower rank computes it all
        return ("rows", 0, sizes[r]) if r == lo else None
    return ("rows", 0, h) if r == lo else ("cols", h, sizes[lo])


def sharded_ward_groups(x_local, sizes, heights, k=101, return_matrix=False, counts=None):
    """cutree(hclust(dist(t(runmean(x, k))), "ward.D2"), h) for each h (R/HoneyBADGER.R:1122-1136) with the CELLS
    sharded over the ranks — the grouping step of a recursion level without any rank holding every cell's
    expression.  x_local: gene-major fp64 CUDA tensor [G, N_local] (this rank's cells of the level, rank-major
    order = the level's cell order); sizes: cells of every rank.  counts=(r_local, n_local) (int32, gene-major)
    instead of x_local: the allele form, r / n smoothed with k = 31 (R/HoneyBADGER.R:1378-1397).  Steps:
      runmean                 local (windows run along the genes of one cell)
      smoothed matrix         all-gather (G x N doubles in total; padded to the largest shard)
      dist                    split by BLOCKS of the symmetric matrix: rank r computes (r, q) for the q at most half
                              a ring ahead (hb_dist_sq_rect_dev; its own diagonal block with hb_dist_sq_dev; the
                              block of two exactly opposite ranks is halved between them) — N^2 / (2 * world) pairs
                              per rank, every entry bit-identical to the single-GPU kernel's
      blocks -> rank 0        send / recv of each rank's row panel; rank 0 fills the other triangle by transposition
      Ward.D2 + cutree        rank 0 (the nearest-neighbour chain is sequential), labels broadcast
    Returns one int32 label array over ALL the level's cells per height (identical on every rank) — and the
    assembled matrix on rank 0 when return_matrix (tests).  world = 1 is honeybadger.ward_groups_dev."""
    import torch
    import torch.distributed as dist
    from . import _lib, hmm
    from .honeybadger import _stream, cut_merges
    lib = _lib.load()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    sizes = [int(v) for v in sizes]
    N = int(sum(sizes))
    G, n_loc = (counts[0] if counts is not None else x_local).shape
    assert len(sizes) == world and n_loc == sizes[rank]
    dev = (counts[0] if counts is not None else x_local).device
    heights = [int(h) for h in heights]
    if N == 1:
        return [np.ones(1, dtype=np.int32) for _ in heights]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    mx = max(sizes)
    # ---- runmean of this rank's cells, then everybody's smoothed cells (padded blocks, gene-major)
    S = torch.zeros((G, mx), dtype=torch.float64, device=dev)
    if n_loc:
        tmp = hmm.alloc_rows(G, n_loc, torch.float64, dev)
        if counts is not None:
            rr, nn = counts
            if rr.stride(0) != nn.stride(0) or rr.stride(1) != 1 or nn.stride(1) != 1:
                rr, nn = rr.contiguous(), nn.contiguous()
            _lib.check(lib.hb_runmean_ratio_dev(rr.data_ptr(), nn.data_ptr(), rr.stride(0), G, n_loc, int(k),
                                                tmp.data_ptr(), tmp.stride(0), _stream()))
        else:
            xr = x_local if x_local.stride(1) == 1 else x_local.contiguous()
            _lib.check(lib.hb_runmean_dev(xr.data_ptr(), xr.stride(0), G, n_loc, int(k), tmp.data_ptr(), tmp.stride(0),
                                          _stream()))
        S[:, :n_loc] = tmp
        del tmp
    if world > 1:
        src = S.cpu() if _staged(S) else S
        parts = [torch.empty_like(src) for _ in range(world)]
        dist.all_gather(parts, src)
    else:
        parts = [S]

    def smoothed(q):   # rank q's cells as an aligned gene-major matrix on this device
        t = hmm.alloc_rows(G, max(sizes[q], 2), torch.float64, dev)
        t[:, :sizes[q]] = parts[q][:, :sizes[q]].to(dev)
        return t

    # ---- this rank's row panel of the distance matrix: the blocks it owns, zeros elsewhere
    panel = torch.zeros((max(n_loc, 1), N), dtype=torch.float64, device=dev)
    mine = smoothed(rank) if n_loc else None
    def rect(a, na, b, nb):   # distances of the na cells of `a` (rows) against the nb cells of `b`, [na, nb]
        blk = torch.empty((na, nb), dtype=torch.float64, device=dev)
        _lib.check(lib.hb_dist_sq_rect_dev(a.data_ptr(), a.stride(0), na, b.data_ptr(), b.stride(0), nb, G,
                                           blk.data_ptr(), blk.stride(0), _stream()))
        return blk

    for q in range(world):
        plan = _block_rows(rank, q, world, sizes)
        if plan is None or not n_loc or not sizes[q]:
            continue
        kind, lo_, hi_ = plan
        if q == rank:
            if n_loc == 1:
                continue   # a single cell: its distance to itself is 0
            blk = torch.empty((n_loc, n_loc), dtype=torch.float64, device=dev)
            _lib.check(lib.hb_dist_sq_dev(mine.data_ptr(), mine.stride(0), G, n_loc, blk.data_ptr(), _stream()))
            panel[:n_loc, offs[q]:offs[q + 1]] = blk
            continue
        other = smoothed(q)
        if kind == "cols":     # every cell of this rank against cells [lo_, hi_) of rank q
            panel[:n_loc, offs[q] + lo_:offs[q] + hi_] = rect(mine, n_loc, other[:, lo_:], hi_ - lo_)
        elif hi_ - lo_ >= 2 and sizes[q] >= 2:
            panel[lo_:hi_, offs[q]:offs[q + 1]] = rect(mine[:, lo_:], hi_ - lo_, other, sizes[q])
        else:
            # the rectangular kernel wants two cells on either side; a one-cell shard is padded with a copy of itself
            a = mine
            if n_loc == 1:
                a = mine.clone()
                a[:, 1] = a[:, 0]
            if sizes[q] == 1:
                other[:, 1] = other[:, 0]
            panel[:n_loc, offs[q]:offs[q + 1]] = rect(a, max(n_loc, 2), other, max(sizes[q], 2))[:n_loc, :sizes[q]]
    # ---- row panels to rank 0, which completes the other triangle
    labels = torch.zeros((len(heights), N), dtype=torch.int32, device=dev)
    D = None
    if rank == 0:
        D = torch.empty((N, N), dtype=torch.float64, device=dev)
        D[offs[0]:offs[1]] = panel[:sizes[0]]
        for r in range(1, world):
            if sizes[r]:
                buf = torch.empty((sizes[r], N), dtype=torch.float64, device=dev)
                _recv(buf, r)
                D[offs[r]:offs[r + 1]] = buf
        for r in range(world):
            for q in range(world):
                if r == q:
                    continue
                plan = _block_rows(r, q, world, sizes)
                rows_r, cols_q = slice(offs[r], offs[r + 1]), slice(offs[q], offs[q + 1])
                if plan is None:               # all of (r, q) is the transpose of what q computed
                    other = _block_rows(q, r, world, sizes)
                    if other[0] == "rows" and (other[1], other[2]) == (0, sizes[q]):
                        D[rows_r, cols_q] = D[cols_q, rows_r].t()
                elif plan[0] == "rows" and (plan[1], plan[2]) != (0, sizes[r]):
                    # r computed its first h rows; the rest is the transpose of q's columns [h, n_r) of (q, r)
                    h = plan[2]
                    D[offs[r] + h:offs[r + 1], cols_q] = D[cols_q, offs[r] + h:offs[r + 1]].t()
                elif plan[0] == "cols":
                    # r computed columns [h, n_q) of (r, q); the first h columns are the transpose of q's rows [0, h)
                    h = plan[1]
                    D[rows_r, offs[q]:offs[q] + h] = D[offs[q]:offs[q] + h, rows_r].t()
        keep = D.clone() if return_matrix else None
        ma = torch.empty(N - 1, dtype=torch.int32, device=dev)
        mb = torch.empty(N - 1, dtype=torch.int32, device=dev)
        hh = torch.empty(N - 1, dtype=torch.float64, device=dev)
        _lib.check(lib.hb_ward_linkage_dev(D.data_ptr(), N, ma.data_ptr(), mb.data_ptr(), hh.data_ptr(), _stream()))
        labs = cut_merges(ma.cpu().numpy(), mb.cpu().numpy(), hh.cpu().numpy(), N, heights)
        labels.copy_(torch.as_tensor(np.stack(labs).astype(np.int32), device=dev))
        D = keep
    elif n_loc:
        _send(panel[:n_loc].contiguous(), 0)
    if world > 1:
        _broadcast(labels, 0)
    out = [a for a in labels.cpu().numpy()]
    return (out, D) if return_matrix else out


class ShardedGexpDriver:
    """HoneyBADGER$calcGexpCnvBoundaries(init = TRUE) with its recursion (R/HoneyBADGER.R:1091-1316) on cells SHARDED
    over the ranks: every level is `sharded_gexp_level` on the level's cells (each rank gathers ITS members of the
    sub-population on the device, ranks may hold none), the split sets g1 / g2 come back as global cell ids and the
    recursion continues on them with the found genes removed (`vi <- !(rownames %in% bound.genes.old)`, :1110).
    Group labels: `labels_fn=None` (default) computes them on the sharded cells (`sharded_ward_groups`: runmean local,
    `dist` split by blocks over the ranks, Ward + cutree on rank 0); a callable `labels_fn(cell_ids, heights)` -> one
    int array per height over those cells supplies them from outside.  Results — identical on
    every rank and to a single rank holding every cell (tests/test_gpu_sharded.py): `bound_genes_final` (list of gene
    index arrays, the reference's bound.genes.final), `regions` (genes, g1 cells), `bound_genes_old`."""

    def __init__(self, x_local, n_cells, dev_mag, labels_fn=None, t=1e-6, min_traverse=5, min_num_genes=10, max_depth=32,
                 window=101):
        import torch.distributed as dist
        self.x, self.n_cells, self.dev, self.window = x_local, int(n_cells), float(dev_mag), int(window)
        self.labels_fn, self.t, self.H, self.min_genes, self.max_depth = labels_fn, t, min_traverse, min_num_genes, max_depth
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = world
        self.lo, self.hi = shard_range(self.n_cells, rank, world)
        assert x_local.shape[1] == self.hi - self.lo, "x_local must hold this rank's contiguous block of cells"
        self.bound_genes_old = np.zeros(0, dtype=np.int64)
        self.bound_genes_final, self.regions, self.levels = [], [], 0

    def run(self):
        self._level(np.arange(self.n_cells, dtype=np.int64), 0)
        return self

    def _level(self, cells, depth):
        import torch
        import torch.distributed as dist
        from .hmm import to_rows
        self.levels += 1
        G = self.x.shape[0]
        genes = np.setdiff1d(np.arange(G, dtype=np.int64), self.bound_genes_old)   # keeps genomic order
        mine = cells[(cells >= self.lo) & (cells < self.hi)]
        dev = self.x.device
        xs = self.x.index_select(0, torch.as_tensor(genes, device=dev))
        xs = to_rows(xs.index_select(1, torch.as_tensor(mine - self.lo, device=dev)).contiguous()) if mine.size else \
            torch.zeros((genes.size, 0), dtype=torch.float64, device=dev)
        heights = list(range(1, min(self.H, cells.size) + 1))
        sel = (cells >= self.lo) & (cells < self.hi)
        sizes = None
        if self.world > 1:
            cnt = torch.zeros(self.world, dtype=torch.int64, device=dev)
            cnt[dist.get_rank()] = int(mine.size)
            _all_reduce(cnt)
            sizes = cnt.cpu().numpy().tolist()
        if self.labels_fn is None:   # grouping of the level's cells across the ranks (dist split by blocks)
            labels = sharded_ward_groups(xs, sizes if sizes is not None else [int(mine.size)], heights, k=self.window)
        else:
            labels = self.labels_fn(cells, heights)
        lvl = sharded_gexp_level(xs, [np.asarray(l)[sel] for l in labels], cells.size, self.dev, t=self.t,
                                 min_num_genes=self.min_genes, local_sizes=sizes)
        if not lvl.get("runs"):
            return
        bound_new = genes[lvl["runs"][lvl["chosen"]]]
        g1, g2 = cells[lvl["g1"]], cells[lvl["g2"]]
        if g1.size:
            self.bound_genes_final.append(bound_new)
            self.regions.append({"genes": bound_new, "cells": g1, "kind": lvl["kinds"][lvl["chosen"]], "depth": depth})
        self.bound_genes_old = np.union1d(self.bound_genes_old, bound_new)
        if depth + 1 >= self.max_depth:
            return
        for g in (g1, g2):
            if g.size >= 3:
                self._level(g, depth + 1)
