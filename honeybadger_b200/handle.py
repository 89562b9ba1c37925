"""Device-resident handle (include/hb_b200.h, "device-resident handle"): the matrices cross PCIe once, every
later call names genes / cells by 0-based index sets.  Thin ctypes mirror of the C ABI the R shim binds
(r/hb_shim.c: C_hb_upload_gexp ... C_hb_retest_gexp_dev); used by the tests and by anything that holds R-layout
host matrices."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _idx(a):
    if a is None:
        return None, 0, None
    v = np.ascontiguousarray(a, dtype=np.int32)
    return v.ctypes.data_as(C.c_void_p), v.size, v


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceMatrix:
    """hb_upload_gexp_host / hb_upload_counts_host -> opaque handle; freed on close() or garbage collection."""

    def __init__(self, ptr, kind):
        self._h = ptr
        self.kind = kind

    @classmethod
    def gexp(cls, gexp_norm):
        """gexp_norm: [G, N] array (R matrix semantics); uploaded as R holds it (column-major)."""
        x = np.asfortranarray(gexp_norm, dtype=np.float64)
        h = C.c_void_p()
        _lib.check(_lib.load().hb_upload_gexp_host(_vp(x), x.shape[0], x.shape[1], C.byref(h)))
        return cls(h, 1)

    @classmethod
    def counts(cls, r_maf, n_sc):
        r = np.asfortranarray(r_maf, dtype=np.float64)
        n = np.asfortranarray(n_sc, dtype=np.float64)
        if r.shape != n.shape:
            raise ValueError("r_maf and n_sc must have the same shape")
        h = C.c_void_p()
        _lib.check(_lib.load().hb_upload_counts_host(_vp(r), _vp(n), r.shape[0], r.shape[1], C.byref(h)))
        return cls(h, 2)

    def info(self):
        k, G, N, b = C.c_int32(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _lib.check(_lib.load().hb_handle_info(self._h, C.byref(k), C.byref(G), C.byref(N), C.byref(b)))
        return {"kind": k.value, "G": G.value, "N": N.value, "h2d_bytes": b.value}

    def close(self):
        if self._h:
            _lib.load().hb_free_handle(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ one recursion level
    def _shape(self, rows, cols):
        i = self.info()
        return (i["G"] if rows is None else len(rows)), (i["N"] if cols is None else len(cols))

    def group(self, heights, window, rows=None, cols=None):
        """cutree labels for each height on the [rows x cols] sub-matrix: list of int32 [ncols]."""
        G2, N2 = self._shape(rows, cols)
        pr, nr, _kr = _idx(rows)
        pc, nc, _kc = _idx(cols)
        ks = np.ascontiguousarray(heights, dtype=np.int32)
        lab = np.zeros((ks.size, N2), dtype=np.int32)
        _lib.check(_lib.load().hb_group_handle(self._h, pr, nr, pc, nc, int(window), _vp(ks), ks.size, _vp(lab)))
        return [lab[i] for i in range(ks.size)]

    def gexp_round(self, member, mean, Pi, delta, rows=None, cols=None):
        """member [K, ncols] (0/1) -> (y int32 [K, nrows], pooled_x [K, nrows], pooled_sd [K])."""
        G2, N2 = self._shape(rows, cols)
        m = np.ascontiguousarray(member, dtype=np.uint8)
        K = m.shape[0]
        if m.shape != (K, N2):
            raise ValueError(f"member must be [K, {N2}]")
        pr, nr, _kr = _idx(rows)
        pc, nc, _kc = _idx(cols)
        y = np.zeros((K, G2), dtype=np.int32)
        px, psd = np.zeros((K, G2)), np.zeros(K)
        mean, Pi, delta = (np.ascontiguousarray(a, dtype=np.float64) for a in (mean, Pi, delta))
        _lib.check(_lib.load().hb_gexp_round_handle(self._h, pr, nr, pc, nc, _vp(m), K, _vp(mean), _vp(Pi),
                                                    _vp(delta), _vp(y), _vp(px), _vp(psd)))
        return y, px, psd

    def allele_round(self, member, prob, Pi, delta, rows=None, cols=None):
        G2, N2 = self._shape(rows, cols)
        m = np.ascontiguousarray(member, dtype=np.uint8)
        K = m.shape[0]
        if m.shape != (K, N2):
            raise ValueError(f"member must be [K, {N2}]")
        pr, nr, _kr = _idx(rows)
        pc, nc, _kc = _idx(cols)
        y = np.zeros((K, G2), dtype=np.int32)
        mafl, sizel = np.zeros((K, G2), dtype=np.int32), np.zeros((K, G2), dtype=np.int32)
        prob, Pi, delta = (np.ascontiguousarray(a, dtype=np.float64) for a in (prob, Pi, delta))
        _lib.check(_lib.load().hb_allele_round_handle(self._h, pr, nr, pc, nc, _vp(m), K, _vp(prob), _vp(Pi),
                                                      _vp(delta), _vp(y), _vp(mafl), _vp(sizel)))
        return y, mafl, sizel

    def retest_gexp(self, rows, mag0, sigma0=None, cols=None):
        _, N2 = self._shape(rows, cols)
        pr, nr, _kr = _idx(rows)
        pc, nc, _kc = _idx(cols)
        s0 = None if sigma0 is None else np.ascontiguousarray(sigma0, dtype=np.float64)
        pa, pd_, mu = np.zeros(N2), np.zeros(N2), np.zeros(N2)
        _lib.check(_lib.load().hb_retest_gexp_handle(self._h, pr, nr, pc, nc, float(mag0), _vp(s0), _vp(pa),
                                                     _vp(pd_), _vp(mu)))
        return pa, pd_, mu

    def retest_allele(self, rows, l_maf, n_bulk, gene_end=None, pseudo=0.1, mono=0.7, cols=None):
        _, N2 = self._shape(rows, cols)
        pr, nr, _kr = _idx(rows)
        pc, nc, _kc = _idx(cols)
        l_maf = np.ascontiguousarray(l_maf, dtype=np.float64)
        n_bulk = np.ascontiguousarray(n_bulk, dtype=np.float64)
        ge = None if gene_end is None else np.ascontiguousarray(gene_end, dtype=np.uint8)
        out = np.zeros(N2)
        _lib.check(_lib.load().hb_retest_allele_handle(self._h, pr, nr, pc, nc, _vp(l_maf), _vp(n_bulk), _vp(ge),
                                                       float(pseudo), float(mono), _vp(out)))
        return out
