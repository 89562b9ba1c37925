"""The `.Call` shim a HoneyBADGER maintainer adds (r/hb_shim.c) LINKED AND RUN against libhb_b200.so.

R does not exist in this project's containers, so the shim could only be syntax-checked against stand-in
headers (tests/test_host_logic.py).  Here it is linked with tests/r_stub/fake_r.c — a miniature of the part of
R's C API it uses (vectors, matrices, lists, external pointers with finalizers, PROTECT as a counter, Rf_error
unwinding with longjmp) — and every routine is called the way r/hb_b200.R calls it: argument coercions, R's
column-major conventions (Pi, member matrices, [G x K] results), PROTECT balance on success and on error, the
error text HiddenMarkov's callers expect, the handle's finalizer.  Results are compared with the ctypes path
through the same C ABI."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "r_stub")
SO = os.path.join(STUB, "libhbshim_test.so")
REALSXP, INTSXP, LGLSXP, VECSXP, EXTPTRSXP = 14, 13, 10, 19, 22


@pytest.fixture(scope="module")
def R():
    from honeybadger_b200 import build as hb_build
    lib_dir = os.path.dirname(hb_build.build())
    srcs = [os.path.join(ROOT, "r", "hb_shim.c"), os.path.join(STUB, "fake_r.c")]
    deps = srcs + [os.path.join(STUB, "Rinternals.h"), os.path.join(ROOT, "include", "hb_b200.h"),
                   os.path.join(lib_dir, "libhb_b200.so")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type",
                               "-fPIC", "-shared", "-I" + STUB, "-I" + os.path.join(ROOT, "include"), *srcs,
                               "-L" + lib_dir, "-lhb_b200", "-Wl,-rpath," + lib_dir, "-o", SO])
    L = C.CDLL(SO)
    p = C.c_void_p
    for name, res, args in [("fr_init", C.c_int, []), ("fr_arity", C.c_int, [C.c_char_p]), ("fr_nil", p, []),
                            ("fr_real", p, [C.c_ssize_t, p]), ("fr_int", p, [C.c_ssize_t, p]),
                            ("fr_lgl", p, [C.c_ssize_t, p]), ("fr_set_dim", None, [p, C.c_int, C.c_int]),
                            ("fr_type", C.c_int, [p]), ("fr_len", C.c_longlong, [p]), ("fr_nrow", C.c_int, [p]),
                            ("fr_ncol", C.c_int, [p]), ("fr_data", p, [p]), ("fr_elt", p, [p, C.c_longlong]),
                            ("fr_error_message", C.c_char_p, []), ("fr_protect_depth", C.c_int, []),
                            ("fr_depth_at_error", C.c_int, []), ("fr_finalize", None, [p]),
                            ("fr_call", p, [C.c_char_p, C.c_int, C.POINTER(p)]), ("fr_reset", None, [])]:
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    assert L.fr_init() == 1
    yield Shim(L)
    L.fr_reset()


class RError(RuntimeError):
    pass


class Shim:
    """R values <-> numpy through the fake runtime."""

    def __init__(self, L):
        self.L = L

    def real(self, a):
        a = np.asarray(a, dtype=np.float64)
        s = self.L.fr_real(a.size, np.asfortranarray(a).ravel(order="F").ctypes.data_as(C.c_void_p))
        if a.ndim == 2:
            self.L.fr_set_dim(s, a.shape[0], a.shape[1])
        return s

    def int(self, a, logical=False):
        a = np.asarray(a, dtype=np.int32)
        mk = self.L.fr_lgl if logical else self.L.fr_int
        s = mk(a.size, np.asfortranarray(a).ravel(order="F").ctypes.data_as(C.c_void_p))
        if a.ndim == 2:
            self.L.fr_set_dim(s, a.shape[0], a.shape[1])
        return s

    def nil(self):
        return self.L.fr_nil()

    def value(self, s):
        t, n = self.L.fr_type(s), self.L.fr_len(s)
        if t == VECSXP:
            return [self.value(self.L.fr_elt(s, i)) for i in range(n)]
        ct = C.c_double if t == REALSXP else C.c_int
        buf = np.ctypeslib.as_array(C.cast(self.L.fr_data(s), C.POINTER(ct)), shape=(n,)).copy() if n else np.zeros(0)
        nr, nc = self.L.fr_nrow(s), self.L.fr_ncol(s)
        return buf.reshape((nr, nc), order="F") if nr >= 0 else buf

    def call(self, name, *args):
        """.Call(name, ...): raises RError with R's message; checks that PROTECT is balanced either way."""
        arr = (C.c_void_p * len(args))(*args)
        d0 = self.L.fr_protect_depth()
        r = self.L.fr_call(name.encode(), len(args), arr)
        if not r:
            msg = self.L.fr_error_message().decode()
            assert self.L.fr_depth_at_error() in (0, -1), f"{name}: PROTECT stack not released before Rf_error"
            raise RError(msg)
        assert self.L.fr_protect_depth() == d0, f"{name}: unbalanced PROTECT / UNPROTECT"
        return r


def test_every_call_in_the_r_file_resolves_with_its_arity(R):
    import re
    rsrc = open(os.path.join(ROOT, "r", "hb_b200.R")).read()
    names = set(re.findall(r'\.Call\("(C_hb_\w+)"', rsrc))
    assert len(names) >= 20
    for n in names:
        assert R.L.fr_arity(n.encode()) > 0, n


def test_errors_surface_as_r_errors_with_a_balanced_protect_stack(R):
    """Without a device every routine must end in Rf_error("libhb_b200: ...") AFTER releasing its PROTECTs;
    with one, a bad argument does the same."""
    from honeybadger_b200 import _lib
    Pi = R.real(np.full((3, 3), 1e-6) + np.eye(3) * (1 - 3e-6))
    if _lib.device_count() == 0:
        with pytest.raises(RError, match="libhb_b200"):
            R.call("C_hb_viterbi_norm", R.real(np.zeros(10)), R.real([-1, 0, 1]), R.real([1.0]), Pi, R.real([0, 1, 0]))
        with pytest.raises(RError, match="libhb_b200"):
            R.call("C_hb_upload_gexp", R.real(np.zeros((4, 3))))
    with pytest.raises(RError, match="invalid or released device handle"):
        R.call("C_hb_handle_bytes", R.real([1.0]))          # not an external pointer
    assert R.L.fr_protect_depth() == 0


@pytest.mark.gpu
def test_the_four_call_sites_through_the_shim(R, oracle):
    from honeybadger_b200 import hmm, synth
    x, _, _ = synth.expression(1500, 7, seed=3)
    sd = np.array([oracle.r_sd(x[:, b]) for b in range(x.shape[1])])
    mean, delta = [-synth.DEV, 0.0, synth.DEV], [0.0, 1.0, 0.0]
    Pi = hmm.sym_Pi(3, 1e-6)
    Pi[0, 1] = 2e-6           # asymmetric on purpose: catches a row/column-major mix-up of Pi
    Pi[0, 0] -= 1e-6
    # vector form (the reference's shape) and matrix form
    y = R.value(R.call("C_hb_viterbi_norm", R.real(x[:, 0]), R.real(mean), R.real([sd[0]]), R.real(Pi), R.real(delta)))
    assert np.array_equal(y, hmm.hmm_norm_host(x[:, 0], mean, sd[0], Pi, delta)["y"])
    assert np.array_equal(y, oracle.viterbi_norm(x[:, 0], mean, [sd[0]] * 3, Pi, delta))
    ym = R.value(R.call("C_hb_viterbi_norm", R.real(x), R.real(mean), R.real(sd), R.real(Pi), R.real(delta)))
    assert ym.shape == x.shape and np.array_equal(ym, hmm.hmm_norm_host(x, mean, sd, Pi, delta)["y"])
    # posteriors: list(y, gamma [T x B x 3], LL)
    yy, g, ll = R.value(R.call("C_hb_posteriors_norm", R.real(x), R.real(mean), R.real(sd), R.real(Pi), R.real(delta)))
    ref = hmm.hmm_norm_host(x, mean, sd, Pi, delta, viterbi=True, gamma=True, ll=True)
    assert np.array_equal(yy, ref["y"])
    assert np.array_equal(np.asarray(g).reshape(x.shape + (3,), order="F"), ref["gamma"])
    assert np.array_equal(np.ravel(ll), np.ravel(ref["ll"]))
    # binomial
    r, n, _, _ = synth.allele(2000, 5, seed=4)
    Pi2 = hmm.sym_Pi(2, 1e-6)
    Pi2[0, 1], Pi2[0, 0] = 3e-6, 1 - 3e-6
    yb = R.value(R.call("C_hb_viterbi_binom", R.real(r[:, 1]), R.real(n[:, 1]), R.real([0.1, 0.45]), R.real(Pi2),
                        R.real([0, 1])))   # R hands counts over as doubles: the shim coerces
    assert np.array_equal(yb, oracle.viterbi_binom(r[:, 1], n[:, 1], [0.1, 0.45], Pi2, [0, 1]))
    # HiddenMarkov's own error text for an underflowing chain (length 1 with delta = (0, 1, 0))
    with pytest.raises(RError, match="^Problems With Underflow$"):
        R.call("C_hb_viterbi_norm", R.real([0.3]), R.real(mean), R.real([1.0]), R.real(hmm.sym_Pi(3, 1e-6)), R.real(delta))


@pytest.mark.gpu
def test_pooled_round_and_handle_through_the_shim(R):
    from honeybadger_b200 import hmm, synth
    from honeybadger_b200.handle import DeviceMatrix
    G, N = 1800, 60
    x, _, clone = synth.expression(G, N, seed=6)
    mean, delta, Pi = [-synth.DEV, 0.0, synth.DEV], [0.0, 1.0, 0.0], hmm.sym_Pi(3, 1e-6)
    member = np.stack([np.ones(N, bool), clone == 0, clone != 0], axis=1)          # [N x K], as R builds it
    y = R.value(R.call("C_hb_gexp_pooled_viterbi", R.real(x), R.int(member, logical=True), R.real(mean), R.real(Pi),
                       R.real(delta)))
    dev = DeviceMatrix.gexp(x)
    yref, _, _ = dev.gexp_round(member.T.astype(np.uint8), mean, Pi, delta)
    assert y.shape == (G, 3) and np.array_equal(y, yref.T)
    # the handle: upload once, then rounds / grouping / retest on 1-based index sets
    h = R.call("C_hb_upload_gexp", R.real(x))
    assert R.L.fr_type(h) == EXTPTRSXP
    assert R.value(R.call("C_hb_handle_bytes", h))[0] == G * N * 8
    cells = np.nonzero(clone == 0)[0]
    genes = np.arange(100, 1500)
    ct = R.value(R.call("C_hb_group_dev", h, R.int(genes + 1), R.int(cells + 1), R.int([101]), R.int([1, 2, 3])))
    ref = dev.group([1, 2, 3], 101, rows=genes, cols=cells)
    assert ct.shape == (cells.size, 3) and all(np.array_equal(ct[:, q], ref[q]) for q in range(3))
    mem2 = np.stack([np.ones(cells.size, bool), ct[:, 1] == 1], axis=1)
    y2 = R.value(R.call("C_hb_gexp_round_dev", h, R.int(genes + 1), R.int(cells + 1), R.int(mem2, logical=True),
                        R.real(mean), R.real(Pi), R.real(delta)))
    y2ref, _, _ = dev.gexp_round(mem2.T.astype(np.uint8), mean, Pi, delta, rows=genes, cols=cells)
    assert np.array_equal(y2, y2ref.T)
    pa, pdl, mu = R.value(R.call("C_hb_retest_gexp_dev", h, R.int(genes[:50] + 1), R.nil(), R.real([synth.DEV]), R.nil()))
    ra, rd, rm = dev.retest_gexp(genes[:50], synth.DEV)
    assert np.array_equal(pa, ra) and np.array_equal(pdl, rd) and np.array_equal(mu, rm)
    # allele counts: upload, one round, the snpModel retest of a region — against the ctypes handle
    r, n, _, _ = synth.allele(2500, 40, seed=8)
    ha = R.call("C_hb_upload_counts", R.real(r), R.real(n))
    da = DeviceMatrix.counts(r, n)
    Pi2 = hmm.sym_Pi(2, 1e-6)
    mema = np.stack([np.ones(40, bool), np.arange(40) % 2 == 0], axis=1)
    ya, mafl, sizel = R.value(R.call("C_hb_allele_round_dev", ha, R.nil(), R.nil(), R.int(mema, logical=True),
                                     R.real([0.1, 0.45]), R.real(Pi2), R.real([0, 1])))
    yr, mr, sr = da.allele_round(mema.T.astype(np.uint8), [0.1, 0.45], Pi2, [0, 1])
    assert np.array_equal(ya, yr.T) and np.array_equal(mafl, mr.T) and np.array_equal(sizel, sr.T)
    reg = np.arange(300, 340)
    lm, nb = (r[reg] > 0).sum(1).astype(float), (n[reg] > 0).sum(1).astype(float)
    pm = R.value(R.call("C_hb_retest_allele_dev", ha, R.int(reg + 1), R.nil(), R.real(lm), R.real(nb), R.nil(),
                        R.real([0.1]), R.real([0.7])))
    assert pm.shape == (1, 40) and np.array_equal(pm[0], da.retest_allele(reg, lm, nb))
    da.close()
    # garbage collection: the finalizer frees the device copy; the pointer is dead afterwards
    R.L.fr_finalize(h)
    with pytest.raises(RError, match="invalid or released device handle"):
        R.call("C_hb_handle_bytes", h)
    dev.close()
