"""CPU: host-side bookkeeping of the mirror (vote/runs/trim, ordering, ranges, sharding) and the ABI."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from honeybadger_b200 import _lib, dist, honeybadger as hb
from oracle import pyref as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@settings(max_examples=300, deadline=None)
@given(st.lists(st.integers(0, 3), min_size=2, max_size=120), st.integers(1, 6),
       st.sampled_from([None, 0.1, 0.25, 0.5, 1.0]))
def test_runs_from_vote_equals_literal_r_loop(vote, min_num, trim):
    a = hb.runs_from_vote(np.array(vote), min_num, trim)
    b = P.runs_filter_trim(np.array(vote), min_num, trim)
    assert (a is None) == (b is None)
    if a is not None:
        assert len(a) == len(b) and all((x == y).all() for x, y in zip(a, b))


def test_runs_quirks():
    # the first element of a run is dropped; a run of L voted positions yields L-1 labelled ones
    v = np.array([0, 1, 1, 1, 1, 0, 1, 1, 0])
    runs = hb.runs_from_vote(v, 1, None)
    assert [r.tolist() for r in runs] == [[2, 3, 4], [7]]
    assert hb.runs_from_vote(v, 3, None)[0].tolist() == [2, 3, 4] and len(hb.runs_from_vote(v, 3, None)) == 1
    assert hb.runs_from_vote(np.zeros(5), 1, 0.1) is None
    # trim: names[1:round(L - L*trim)], half-even: L=5 -> round(4.5)=4 ; L=15 -> round(13.5)=14
    v5 = np.array([0] + [1] * 6)
    assert hb.runs_from_vote(v5, 1, 0.1)[0].size == 4
    v15 = np.array([0] + [1] * 16)
    assert hb.runs_from_vote(v15, 1, 0.1)[0].size == 14
    # position 0 can never be labelled (the R loop starts at i = 2)
    assert hb.runs_from_vote(np.array([1, 1, 1]), 1, None)[0].tolist() == [1, 2]


def test_group_members_order_is_first_appearance():
    lab = np.array([2, 2, 1, 3, 1, 2])
    gm = hb.group_members(lab)
    assert [g for g, _ in gm] == [2, 1, 3]
    assert gm[0][1].tolist() == [0, 1, 5]
    m, tags = hb._membership([np.array([1, 1, 1, 1, 1, 1]), lab], 6)
    assert m.shape == (3, 6) and tags == [(0, 0), (1, 0), (1, 1)]  # the singleton group 3 is skipped


def test_order_genes_and_range():
    g = hb.GRanges(["chr2", "chr1", "chr1", "chrX", "chr2"], [50, 300, 100, 5, 10], [60, 400, 200, 9, 20],
                   ["a", "b", "c", "d", "e"])
    o = hb.order_genes(["a", "b", "c", "d", "e"], g, ["chr1", "chr2"])
    assert o.tolist() == [2, 1, 4, 0]  # chr1 by midpoint, then chr2; chrX dropped
    o = hb.order_genes(["a", "b", "c", "d", "e"], g, ["chr1", "chr2"], has_na=[False, True, False, False, False])
    assert o.tolist() == [2, 4, 0]
    assert g[["b", "c", "a"]].range() == [("chr1", 100.0, 400.0), ("chr2", 50.0, 60.0)]
    s = hb.snps_from_names(["1:762273", "chr13:21949361", "14:5-9"])
    assert list(s.seqnames) == ["chr1", "chr13", "chr14"] and s.end[2] == 9


def test_ward_groups_numbering_and_runmean():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(50, 6))
    a[:, 3:] += 4
    from oracle import oracle as O
    labs = O.ward_groups(a, [1, 2, 3])
    assert labs[0].tolist() == [1] * 6 and labs[1].tolist() == [1, 1, 1, 2, 2, 2] and labs[2][0] == 1
    m = np.arange(10, dtype=float).reshape(-1, 1)
    rm = O.runmean(m, 3)[:, 0]
    assert rm[0] == 0.5 and rm[1] == 1.0 and rm[-1] == 8.5
    m[4] = np.nan
    assert O.runmean(m, 3)[4, 0] == 4.0


def test_shard_ranges_cover_cells():
    for n, w in [(10, 3), (200000, 8), (7, 8), (25000, 1)]:
        r = [dist.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(dist.shard_sizes(n, w)) - min(dist.shard_sizes(n, w)) <= 1


def test_abi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} missing from libhb_b200.so"
    assert _lib.load().hb_version() >= 100


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_device():
    from honeybadger_b200 import hmm
    with pytest.raises(_lib.HBError) as e:
        hmm.hmm_norm_host(np.zeros(8), [-1, 0, 1], 1.0, hmm.sym_Pi(3, 1e-6), [0, 1, 0])
    assert e.value.code in (_lib.HB_ERR_NODEVICE, _lib.HB_ERR_CUDA)
    with pytest.raises(_lib.HBError):
        hmm.hmm_binom_host(np.zeros(8), np.zeros(8), [0.1, 0.45], hmm.sym_Pi(2, 1e-6), [0, 1])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "honeybadger_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "hb_oracle" not in txt and "libhb_emul" not in txt, f


def test_cut_merges_matches_scipy_cut_tree():
    """cutree from a merge list of leaf representatives: same partitions as scipy for every k, numbered by
    first appearance in cell order (R's cutree), for any order of the requested cuts."""
    from scipy.cluster.hierarchy import cut_tree, linkage
    rng = np.random.default_rng(12)
    N = 300
    Z = linkage(rng.standard_normal((N, 5)), "ward")
    rep = list(range(N)) + [0] * (N - 1)
    ma, mb = [], []
    for i, (a, b, _, _) in enumerate(Z):
        ma.append(rep[int(a)])
        mb.append(rep[int(b)])
        rep[N + i] = rep[int(a)]
    ks = [3, 1, 5, 2, 4, 300, 299]
    labs = hb.cut_merges(ma, mb, Z[:, 2], N, ks)
    for k, lab in zip(ks, labs):
        ref = cut_tree(Z, n_clusters=k)[:, 0]
        assert len(set(lab)) == k
        assert np.array_equal(lab[:, None] == lab[None, :], ref[:, None] == ref[None, :])
        first = [int(np.nonzero(lab == g)[0][0]) for g in range(1, k + 1)]
        assert first == sorted(first)          # clusters numbered in order of first appearance


def test_cutree_entry_point_edges_and_errors():
    """hb_cutree_host: k above the number of cells (every cell its own cluster), k = 1, two cells, equal heights
    kept in list order (children before parents), and a merge list naming a leaf outside [0, N) is refused."""
    from honeybadger_b200 import _lib
    ma, mb, hh = [0, 2, 0], [1, 3, 2], [1.0, 1.0, 2.0]      # ((0,1),(2,3)) then the two pairs
    labs = hb.cut_merges(ma, mb, hh, 4, [1, 2, 3, 4, 9])
    assert labs[0].tolist() == [1, 1, 1, 1]
    assert labs[1].tolist() == [1, 1, 2, 2]
    assert labs[2].tolist() == [1, 1, 2, 3]                  # equal heights: the first merge in the list is cut last
    assert labs[3].tolist() == [1, 2, 3, 4] and labs[4].tolist() == [1, 2, 3, 4]
    assert hb.cut_merges([0], [1], [0.5], 2, [1, 2])[1].tolist() == [1, 2]
    with pytest.raises(Exception):
        hb.cut_merges([0, 7], [1, 2], [1.0, 2.0], 3, [1])
    assert "leaf" in _lib.load().hb_last_error().decode()


def test_r_shim_is_well_formed_against_the_c_abi():
    """R is not in this image, so nobody links r/hb_shim.c here; at least it must be valid C against
    include/hb_b200.h (argument counts and pointer types of every hb_* call) with a stand-in for the R API
    (tests/r_stub), its registration table must carry each entry point's real arity, and every .Call in
    r/hb_b200.R must name a registered routine."""
    import subprocess
    shim = os.path.join(ROOT, "r", "hb_shim.c")
    cmd = ["/usr/bin/gcc", "-std=c11", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type",
           "-I" + os.path.join(ROOT, "tests", "r_stub"), "-I" + os.path.join(ROOT, "include"), shim]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    src = open(shim).read()
    defs = {m.group(1): m.group(2).count("SEXP") for m in re.finditer(r"^SEXP (C_hb_\w+)\(([^)]*)\)", src, flags=re.M)}
    table = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(C_hb_\w+)", \(DL_FUNC\)&\1, (\d+)\}', src)}
    assert defs and defs == table, (defs, table)
    rsrc = open(os.path.join(ROOT, "r", "hb_b200.R")).read()
    called = set(re.findall(r'\.Call\("(C_hb_\w+)"', rsrc))
    assert called and called <= set(table), called - set(table)
    for name in called:                              # argument count of each .Call == registered arity
        for m in re.finditer(r'\.Call\("%s"((?:[^()]|\([^()]*(?:\([^()]*\)[^()]*)*\))*)\)' % name, rsrc):
            depth, nargs = 0, 0
            for ch in m.group(1):
                depth += ch in "([" 
                depth -= ch in ")]"
                nargs += ch == "," and depth == 0
            assert nargs == table[name], (name, nargs, table[name])


def test_sharded_dist_block_plan_covers_every_pair_exactly_once():
    """dist._block_rows: over any ring size and shard sizes (empty, one-cell, odd) every unordered pair of cells is
    computed by exactly one rank, and the ranks' pair counts differ by little (opposite ranks halve their block)."""
    from honeybadger_b200 import dist as hd
    rng = np.random.default_rng(0)
    cases = [[7], [5, 4], [64, 1], [1, 1], [40, 41, 39], [5, 0, 60], [1, 70, 2, 33], [10] * 8, [9, 0, 3, 12, 1, 8]]
    cases += [rng.integers(0, 30, size=w).tolist() for w in (2, 3, 4, 5, 6, 7, 8) for _ in range(5)]
    for sizes in cases:
        W, N = len(sizes), int(sum(sizes))
        offs = np.concatenate([[0], np.cumsum(sizes)])
        cover = np.zeros((N, N), dtype=np.int32)
        work = np.zeros(W)
        for r in range(W):
            for q in range(W):
                plan = hd._block_rows(r, q, W, sizes)
                if plan is None or not sizes[r] or not sizes[q]:
                    continue
                kind, a, b = plan
                if r == q:
                    assert (kind, a, b) == ("rows", 0, sizes[r])
                    blk = np.triu(np.ones((sizes[r], sizes[r]), dtype=np.int32))      # the symmetric kernel: each pair once
                    cover[offs[r]:offs[r + 1], offs[r]:offs[r + 1]] += blk
                    work[r] += sizes[r] * (sizes[r] + 1) / 2
                elif kind == "rows":
                    cover[offs[r] + a:offs[r] + b, offs[q]:offs[q + 1]] += 1
                    work[r] += (b - a) * sizes[q]
                else:
                    assert (a % 2 == 0) and a >= 2   # the column sub-matrix stays 16-byte aligned
                    cover[offs[r]:offs[r + 1], offs[q] + a:offs[q] + b] += 1
                    work[r] += sizes[r] * (b - a)
        both = cover + cover.T - np.diag(np.diag(cover))
        assert (both == 1).all(), sizes
        if W > 1 and len(set(sizes)) == 1 and sizes[0] >= 8:
            assert work.max() <= 1.3 * work.min(), (sizes, work)


def test_integration_md_names_every_exported_symbol():
    """INTEGRATION.md's index must name each entry point include/hb_b200.h declares (the header is the contract, the
    document says what each symbol replaces in the reference)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "hb_b200.h")).read()
    names = sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", hdr)))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in names if n not in doc]
    assert not missing, missing
