"""Parity against the REFERENCE ITSELF, when its golden vectors are available.

`tools/r_parity.R` (run once on any machine with R + HiddenMarkov + caTools) writes
tests/golden/r_parity_mgh31.txt: for the bundled MGH31 data, every (height, group) of the four literal call
sites (R/HoneyBADGER.R:1122-1151, :1378-1410 and their function twins) with the exact inputs and outputs of
`HiddenMarkov::dthmm -> Viterbi` (+ forwardback posteriors), R's dnorm / dbinom log-densities, sd(), the
pooled vectors and cutree labels.  R does not exist in this project's containers, so the file cannot be
produced here: until somebody commits it these tests SKIP (and DESIGN.md keeps saying "parity unpinned" for
the per-step arithmetic).  The checker itself is exercised on every run by a dump in the same format written
from the oracle's own values (`test_checker_on_a_synthetic_dump`), so a committed file is a one-command fix.
"""
from __future__ import annotations

import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DUMP = os.path.join(HERE, "golden", "r_parity_mgh31.txt")


# --------------------------------------------------------------------------- parser
def parse_dump(path):
    """-> dict(header=str, cases=[dict], cutree={(kind, h): labels}, meta={...}, dbinom=[(n, xs, pd, pn)])"""
    out = {"header": "", "cases": [], "cutree": {}, "meta": {}, "dbinom": []}
    cur = None
    with open(path) as fh:
        for ln in fh:
            tok = ln.split()
            if not tok:
                continue
            if tok[0] == "#":
                out["header"] = ln.strip()
            elif tok[0] == "CASE":
                cur = {"kind": tok[1], "h": int(tok[3]), "group": int(tok[5]), "ncell": int(tok[7]), "n": int(tok[9])}
                out["cases"].append(cur)
            elif tok[0] == "CUTREE":
                out["cutree"][(tok[1], int(tok[3]))] = np.array(tok[5:], dtype=np.int32)
            elif tok[0] in ("GEXP", "ALLELE"):
                for k, v in zip(tok[1::2], tok[2::2]):
                    out["meta"][f"{tok[0].lower()}_{k}"] = float(v)
            elif tok[0] == "SNPNAMES":
                out["meta"]["snpnames"] = tok[1:]
            elif tok[0] in ("RUNMEAN_COL1", "HCLUST_HEIGHT"):
                out["meta"].setdefault(tok[0].lower(), []).append(np.array(tok[1:], dtype=np.float64))
            elif tok[0] == "DBINOM_TABLE":
                cur = None
            elif tok[0] == "dbinom":
                n = int(tok[2])
                i, j = tok.index("pd"), tok.index("pn")
                out["dbinom"].append((n, np.arange(n + 1), np.array(tok[i + 1:j], float), np.array(tok[j + 1:], float)))
            elif tok[0] == "dbinom_deep":
                n = int(tok[2])
                i, j, k = tok.index("x"), tok.index("pd"), tok.index("pn")
                out["dbinom"].append((n, np.array(tok[i + 1:j], int), np.array(tok[j + 1:k], float),
                                      np.array(tok[k + 1:], float)))
            elif cur is not None:
                key = tok[0]
                if key in ("y", "mafl", "sizel"):
                    cur[key] = np.array(tok[1:], dtype=np.int32)
                elif key in ("sd", "LL"):
                    cur[key] = float(tok[1])
                else:
                    cur[key] = np.array(tok[1:], dtype=np.float64)
    return out


# --------------------------------------------------------------------------- checker
DEVV, T_SW, PD, PN = 1.015658, 1e-6, 0.1, 0.45


def check_dump(d, O, gpu_hmm=None):
    """Every CASE of a dump against the oracle `O` (and, when given, the CUDA front end `gpu_hmm`)."""
    nchk = 0
    Pi3, Pi2 = O.sym_Pi(3, T_SW), O.sym_Pi(2, T_SW)
    for c in d["cases"]:
        if c["kind"] == "gexp":
            x, sd = c["x"], c["sd"]
            assert O.r_sd(x) == sd, "sd() differs from R"
            mean = [-DEVV, 0.0, DEVV]
            for k, mu in enumerate(mean):
                lb = np.array([O.dnorm_log(v, mu, sd) for v in x])
                assert np.array_equal(lb, c[f"logb{k + 1}"]), f"dnorm(log=TRUE) differs from R, state {k + 1}"
            y = O.viterbi_norm(x, mean, [sd] * 3, Pi3, [0.0, 1.0, 0.0])
            assert np.array_equal(y, c["y"]), f"Viterbi path differs from R (gexp h={c['h']} group={c['group']})"
            g, ll = O.forwardback_norm(x, mean, [sd] * 3, Pi3, [0.0, 1.0, 0.0])
            G = np.stack([c["gamma1"], c["gamma2"], c["gamma3"]], axis=1)
            assert np.abs(g - G).max() < 1e-9 and abs(ll - c["LL"]) <= 1e-9 * max(1.0, abs(c["LL"]))
            if gpu_hmm is not None:
                r = gpu_hmm.hmm_norm_host(x, mean, sd, Pi3, [0.0, 1.0, 0.0], viterbi=True, gamma=True, ll=True)
                assert np.array_equal(r["y"], c["y"]), "CUDA Viterbi path differs from R"
                assert np.abs(r["gamma"] - G).max() < 1e-9
        else:
            x, n = c["mafl"], c["sizel"]
            for k, p in enumerate((PD, PN)):
                lb = np.array([O.dbinom(a, b, p, log=True) for a, b in zip(x, n)])
                assert np.array_equal(lb, c[f"logb{k + 1}"]), f"dbinom(log=TRUE) differs from R, state {k + 1}"
            y = O.viterbi_binom(x, n, [PD, PN], Pi2, [0.0, 1.0])
            assert np.array_equal(y, c["y"]), f"Viterbi path differs from R (allele h={c['h']} group={c['group']})"
            G = np.stack([c["gamma1"], c["gamma2"]], axis=1)
            ok = np.isfinite(G).all()          # linear densities of deep pooled counts underflow to NaN in R
            if ok:
                g, ll = O.forwardback_binom(x, n, [PD, PN], Pi2, [0.0, 1.0])
                assert np.abs(g - G).max() < 1e-9
            if gpu_hmm is not None:
                r = gpu_hmm.hmm_binom_host(x, n, [PD, PN], Pi2, [0.0, 1.0], viterbi=True, gamma=True)
                assert np.array_equal(r["y"], c["y"]), "CUDA Viterbi path differs from R"
                if ok:
                    assert np.abs(r["gamma"] - G).max() < 1e-9
        nchk += 1
    for n, xs, lpd, lpn in d["dbinom"]:
        for xv, a, b in zip(xs, lpd, lpn):
            assert O.dbinom(xv, n, PD, log=True) == a and O.dbinom(xv, n, PN, log=True) == b, (xv, n)
    return nchk


def check_pooling(d, O, golden_allele, golden_inputs):
    """Pooled vectors of the dump against the oracle's R-mean / rowSums restatements on the bundled matrices,
    with R's own cutree labels."""
    n = 0
    names = d["meta"].get("snpnames")
    if names is not None:
        ours = [f"{c}:{p}" for c, p in zip(golden_allele["chrom"], golden_allele["pos"])]
        assert list(names) == ours, "setAlleleMats keeps a different SNP set than R"
    r_maf, n_sc = golden_allele["r_maf"].astype(np.float64), golden_allele["n_sc"].astype(np.float64)
    gx = golden_inputs["gexp"] - golden_inputs["ref"][:, None]       # first 800 genes, all 75 cells
    for c in d["cases"]:
        lab = d["cutree"][(c["kind"], c["h"])]
        cols = np.nonzero(lab == c["group"])[0].astype(np.int32)
        if c["kind"] == "allele":
            assert np.array_equal(O.pool_count_pos(r_maf, cols), c["mafl"])
            assert np.array_equal(O.pool_count_pos(n_sc, cols), c["sizel"])
        else:
            assert np.array_equal(O.pool_mean(gx, cols), c["x"][: gx.shape[0]]), "apply(., 1, mean) differs from R"
        n += 1
    return n


# --------------------------------------------------------------------------- tests
needs_dump = pytest.mark.skipif(not os.path.exists(DUMP), reason="tests/golden/r_parity_mgh31.txt absent: run "
                                "`Rscript tools/r_parity.R <HoneyBADGER checkout> tests/golden/r_parity_mgh31.txt` "
                                "on a machine with R (none in this project's containers)")


@needs_dump
def test_oracle_against_the_reference_dump(oracle, golden_allele, golden_inputs):
    d = parse_dump(DUMP)
    assert check_dump(d, oracle) >= 20
    assert check_pooling(d, oracle, golden_allele, golden_inputs) >= 20


@needs_dump
@pytest.mark.gpu
def test_cuda_against_the_reference_dump(oracle):
    from honeybadger_b200 import hmm
    d = parse_dump(DUMP)
    assert check_dump(d, oracle, gpu_hmm=hmm) >= 20


@needs_dump
@pytest.mark.gpu
def test_device_grouping_against_r_cutree(golden_allele):
    """runmean(k = 31) -> dist -> hclust(ward.D2) -> cutree on the device == R's labels on the bundled data."""
    import torch
    from honeybadger_b200 import honeybadger as hb
    d = parse_dump(DUMP)
    r = torch.from_numpy(golden_allele["r_maf"].astype(np.int32)).cuda()
    n = torch.from_numpy(golden_allele["n_sc"].astype(np.int32)).cuda()
    labs = hb.ward_groups_dev(None, [1, 2, 3], 31, counts=(r, n))
    for h, lab in zip((1, 2, 3), labs):
        assert np.array_equal(lab, d["cutree"][("allele", h)])


def _write_synthetic_dump(path, O):
    """A dump in tools/r_parity.R's format whose values come from the oracle — exercises parser + checker."""
    num = lambda v: " ".join("%.17g" % float(a) for a in np.atleast_1d(v))  # noqa: E731
    rng = np.random.default_rng(2)
    with open(path, "w") as fh:
        fh.write("# hb_r_parity v1  synthetic (oracle-generated; NOT reference output)\n")
        x = rng.standard_normal(300) * 1.1
        x[100:160] += 1.4
        sd = O.r_sd(x)
        mean = [-DEVV, 0.0, DEVV]
        Pi3, Pi2 = O.sym_Pi(3, T_SW), O.sym_Pi(2, T_SW)
        fh.write("CUTREE gexp h 1 labels 1 1 1\n")
        fh.write("CASE gexp h 1 group 1 ncell 3 n 300\n")
        fh.write(f"sd {num(sd)}\nx {num(x)}\n")
        for k, mu in enumerate(mean):
            fh.write(f"logb{k + 1} {num([O.dnorm_log(v, mu, sd) for v in x])}\n")
        fh.write("y " + " ".join(map(str, O.viterbi_norm(x, mean, [sd] * 3, Pi3, [0.0, 1.0, 0.0]))) + "\n")
        g, ll = O.forwardback_norm(x, mean, [sd] * 3, Pi3, [0.0, 1.0, 0.0])
        fh.write(f"LL {num(ll)}\n")
        for k in range(3):
            fh.write(f"gamma{k + 1} {num(g[:, k])}\n")
        n = (rng.random(400) < 0.3) * rng.integers(1, 20, 400)
        r = rng.binomial(n, 0.45)
        r[150:220] = rng.binomial(n[150:220], 0.1)
        fh.write("CASE allele h 1 group 1 ncell 3 n 400\n")
        fh.write("mafl " + " ".join(map(str, r)) + "\nsizel " + " ".join(map(str, n)) + "\n")
        for k, p in enumerate((PD, PN)):
            fh.write(f"logb{k + 1} {num([O.dbinom(a, b, p, log=True) for a, b in zip(r, n)])}\n")
        fh.write("y " + " ".join(map(str, O.viterbi_binom(r, n, [PD, PN], Pi2, [0.0, 1.0]))) + "\n")
        g, ll = O.forwardback_binom(r, n, [PD, PN], Pi2, [0.0, 1.0])
        fh.write(f"LL {num(ll)}\n")
        for k in range(2):
            fh.write(f"gamma{k + 1} {num(g[:, k])}\n")
        fh.write("DBINOM_TABLE nmax 3\n")
        for nn in range(4):
            fh.write(f"dbinom n {nn} pd {num([O.dbinom(v, nn, PD, log=True) for v in range(nn + 1)])} "
                     f"pn {num([O.dbinom(v, nn, PN, log=True) for v in range(nn + 1)])}\n")
        fh.write(f"dbinom_deep n 1245 x 0 60 1245 pd {num([O.dbinom(v, 1245, PD, log=True) for v in (0, 60, 1245)])} "
                 f"pn {num([O.dbinom(v, 1245, PN, log=True) for v in (0, 60, 1245)])}\n")


def test_checker_on_a_synthetic_dump(oracle, tmp_path):
    p = str(tmp_path / "dump.txt")
    _write_synthetic_dump(p, oracle)
    d = parse_dump(p)
    assert len(d["cases"]) == 2 and len(d["dbinom"]) == 5
    assert check_dump(d, oracle) == 2
    # a corrupted state must be caught
    d["cases"][0]["y"] = d["cases"][0]["y"].copy()
    d["cases"][0]["y"][120] = 2 if d["cases"][0]["y"][120] != 2 else 3
    with pytest.raises(AssertionError):
        check_dump(d, oracle)


@pytest.mark.gpu
def test_checker_on_a_synthetic_dump_through_cuda(oracle, tmp_path):
    from honeybadger_b200 import hmm
    p = str(tmp_path / "dump.txt")
    _write_synthetic_dump(p, oracle)
    assert check_dump(parse_dump(p), oracle, gpu_hmm=hmm) == 2
