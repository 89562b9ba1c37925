"""Minimal reader for R `save()` files (bzip2/gzip RDX2 + XDR serialization v2).

TEST INFRASTRUCTURE ONLY (see oracle/README.md): used by tests/golden/make_golden.py to read
the reference's bundled example data (`/root/reference/data/*.rda`, described in
`R/data.R:1-22`) in a container that has no R.  Supports exactly what those files contain:
a pairlist of named REALSXP/INTSXP/STRSXP vectors with `dim`, `dimnames`, `names` attributes.
"""
from __future__ import annotations

import bz2
import gzip
import struct

import numpy as np

NILVALUE_SXP = 254
REFSXP = 255
SYMSXP = 1
LISTSXP = 2
CHARSXP = 9
LGLSXP = 10
INTSXP = 13
REALSXP = 14
STRSXP = 16
VECSXP = 19


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs: list = []

    def i32(self) -> int:
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def take(self, n: int) -> bytes:
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def length(self) -> int:
        n = self.i32()
        if n == -1:  # long vector
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == NILVALUE_SXP:
            return None
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if t == CHARSXP:
            n = self.i32()
            return None if n == -1 else self.take(n).decode("utf-8", "replace")
        if t == LISTSXP:
            out = []
            while True:
                attr = self.item() if has_attr else None  # noqa: F841
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t = flags & 0xFF
                if t == NILVALUE_SXP:
                    break
                if t != LISTSXP:
                    raise ValueError(f"unexpected pairlist cdr type {t}")
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if t in (LGLSXP, INTSXP):
            n = self.length()
            val = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif t == REALSXP:
            n = self.length()
            val = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif t == STRSXP:
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif t == VECSXP:
            n = self.length()
            val = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {t} at byte {self.p}")
        attrs = dict(self.item()) if has_attr else {}
        return {"value": val, "attr": attrs}


def read_rda(path: str) -> dict:
    """Return {object name: {"value": ndarray|list, "attr": {...}}} for an .rda file."""
    raw = open(path, "rb").read()
    if raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    if raw[:5] != b"RDX2\n":
        raw_head = raw[:5]
        raise ValueError(f"not an RDX2 file: {raw_head!r}")
    if raw[5:7] != b"X\n":
        raise ValueError("only XDR serialization supported")
    r = _Reader(raw)
    r.p = 7
    r.i32()  # serialization version
    r.i32()  # writer R version
    r.i32()  # min reader version
    top = r.item()
    return {tag: car for tag, car in top}


def as_matrix(obj) -> tuple[np.ndarray, list | None, list | None]:
    """(values as [nrow, ncol] C-order array, rownames, colnames) of an R matrix object."""
    v = np.asarray(obj["value"])
    dim = obj["attr"]["dim"]["value"]
    m = v.reshape((int(dim[1]), int(dim[0]))).T.copy()  # R is column-major
    dn = obj["attr"].get("dimnames")
    rn = cn = None
    if dn is not None:
        rn_o, cn_o = dn["value"]
        rn = rn_o["value"] if rn_o is not None else None
        cn = cn_o["value"] if cn_o is not None else None
    return m, rn, cn
