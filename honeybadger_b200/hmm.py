"""HiddenMarkov-shaped front end of the CUDA chains.

Mirrors the three calls HoneyBADGER makes into (or, for forwardback, could make into) the CRAN
package `HiddenMarkov` — same names, argument meaning and error behaviour:

    z <- HiddenMarkov::dthmm(x, Pi, delta, "norm", list(mean=c(pd,pn,pa), sd=c(sd,sd,sd)))
    y <- HiddenMarkov::Viterbi(z)                        # R/HoneyBADGER.R:1150-1151, _gexp.R:503-504
    z <- HiddenMarkov::dthmm(mafl, Pi, delta, "binom", list(prob=c(pd,pn)), list(size=sizel), discrete=TRUE)
    y <- HiddenMarkov::Viterbi(z)                        # R/HoneyBADGER.R:1409-1410, _allele.R:579-580

plus the batched, device-resident entry points the segmentation driver and bench.py use
(`hmm_norm_dev`, `hmm_binom_dev`: torch CUDA tensors in the gene-major layout, current stream).
Everything runs in libhb_b200.so; there is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib

__all__ = ["dthmm", "Viterbi", "forwardback", "hmm_norm_host", "hmm_binom_host", "hmm_norm_dev",
           "hmm_binom_dev", "sym_Pi"]


def sym_Pi(S: int, t: float) -> np.ndarray:
    """matrix(c(1-2t,t,t, t,1-2t,t, t,t,1-2t), byrow=TRUE, nrow=3) / matrix(c(1-t,t,t,1-t), ...)."""
    Pi = np.full((S, S), t, dtype=np.float64)
    np.fill_diagonal(Pi, 1 - (S - 1) * t)
    return Pi


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _seg(seg_off, T):
    if seg_off is None:
        seg_off = [0, T]
    s = np.ascontiguousarray(seg_off, dtype=np.int64)
    if s.ndim != 1 or s.size < 2:
        raise ValueError("seg_off must hold nseg+1 offsets")
    return s


def _flags(viterbi, gamma, ll):
    f = (_lib.WANT_VITERBI if viterbi else 0) | (_lib.WANT_GAMMA if gamma else 0)
    if ll:
        f |= _lib.WANT_LL  # alone: the forward sweep only
    if not f:
        raise ValueError("nothing requested")
    return f


# --------------------------------------------------------------------------- host buffers
def hmm_norm_host(x, mean, sd, Pi, delta, seg_off=None, viterbi=True, gamma=False, ll=False):
    """x: [T] or [T, B] (R matrix semantics: one chain per column).  sd: scalar, [B] or [nseg, B].

    Returns dict(y=int32 [T,B] 1-based, gamma=[T,B,3], ll=[B,nseg]) with the requested keys.
    """
    lib = _lib.load()
    x = np.asarray(x, dtype=np.float64)
    vec = x.ndim == 1
    xm = x.reshape(-1, 1) if vec else x
    T, B = xm.shape
    xf = np.asfortranarray(xm)
    seg = _seg(seg_off, T)
    nseg = seg.size - 1
    sd = np.asarray(sd, dtype=np.float64)
    if sd.ndim == 0:
        sd = np.full(B, float(sd))
    per_seg = int(sd.ndim == 2)
    if sd.shape != ((nseg, B) if per_seg else (B,)):
        raise ValueError(f"sd has shape {sd.shape}, expected ({B},) or ({nseg}, {B})")
    sd = np.ascontiguousarray(sd)
    mean, Pi, delta = _f64(mean), _f64(Pi), _f64(delta)
    if mean.size != 3 or Pi.size != 9 or delta.size != 3:
        raise ValueError("the Gaussian chain has 3 states: mean[3], Pi[3x3], delta[3]")
    want_g = gamma
    y = np.zeros((T, B), dtype=np.int32, order="F") if viterbi else None
    g = np.zeros((T, B, 3), dtype=np.float64, order="F") if want_g else None
    l = np.zeros((B, nseg), dtype=np.float64, order="F") if ll else None
    rc = lib.hb_hmm_norm_host(_vp(xf), T, B, _vp(seg), nseg, _vp(mean), _vp(sd), per_seg, _vp(Pi),
                              _vp(delta), _flags(viterbi, gamma, ll), _vp(y), _vp(g), _vp(l))
    _lib.check(rc)
    out = {}
    if viterbi:
        out["y"] = y[:, 0] if vec else y
    if gamma:
        out["gamma"] = g[:, 0, :] if vec else g
    if ll:
        out["ll"] = l[0] if vec else l
    return out


def hmm_binom_host(x, size, prob, Pi, delta, seg_off=None, rho=0.0, viterbi=True, gamma=False,
                   ll=False):
    lib = _lib.load()
    x = np.asarray(x)
    size = np.asarray(size)
    if x.shape != size.shape:
        raise ValueError("x and size must have the same shape")
    vec = x.ndim == 1
    xm = (x.reshape(-1, 1) if vec else x)
    sm = (size.reshape(-1, 1) if vec else size)
    xr, sr = np.rint(xm), np.rint(sm)
    xf = np.asfortranarray(xr, dtype=np.int32)
    sf = np.asfortranarray(sr, dtype=np.int32)
    T, B = xf.shape
    seg = _seg(seg_off, T)
    nseg = seg.size - 1
    prob, Pi, delta = _f64(prob), _f64(Pi), _f64(delta)
    if prob.size != 2 or Pi.size != 4 or delta.size != 2:
        raise ValueError("the binomial chain has 2 states: prob[2], Pi[2x2], delta[2]")
    want_g = gamma
    y = np.zeros((T, B), dtype=np.int32, order="F") if viterbi else None
    g = np.zeros((T, B, 2), dtype=np.float64, order="F") if want_g else None
    l = np.zeros((B, nseg), dtype=np.float64, order="F") if ll else None
    rc = lib.hb_hmm_binom_host(_vp(xf), _vp(sf), T, B, _vp(seg), nseg, _vp(prob), float(rho),
                               _vp(Pi), _vp(delta), _flags(viterbi, gamma, ll), _vp(y), _vp(g),
                               _vp(l))
    _lib.check(rc)
    out = {}
    if viterbi:
        out["y"] = y[:, 0] if vec else y
    if gamma:
        out["gamma"] = g[:, 0, :] if vec else g
    if ll:
        out["ll"] = l[0] if vec else l
    return out


# --------------------------------------------------------------------------- HiddenMarkov mirror
@dataclass
class dthmm:
    """HiddenMarkov::dthmm(x, Pi, delta, distn, pm, pn = NULL, discrete = NULL, nonstat = FALSE).

    Supported on the device: distn "norm" with 3 states and a shared sd (pm = dict(mean=[3], sd=[3]
    equal entries)), distn "binom" with 2 states (pm = dict(prob=[2]), pn = dict(size=[T])).
    `x` may also be a [T, B] matrix: one chain per column (batch extension).
    """
    x: np.ndarray
    Pi: np.ndarray
    delta: np.ndarray
    distn: str
    pm: dict
    pn: dict | None = None
    discrete: bool | None = None
    nonstat: bool = False
    seg_off: np.ndarray | None = field(default=None)

    def __post_init__(self):
        self.Pi = np.asarray(self.Pi, dtype=np.float64)
        self.delta = np.asarray(self.delta, dtype=np.float64)
        m = self.Pi.shape[0]
        if self.Pi.shape != (m, m) or self.delta.shape != (m,):
            raise ValueError("Pi must be m x m and delta of length m")
        if self.distn == "norm":
            sd = np.asarray(self.pm["sd"], dtype=np.float64)
            if m != 3 or np.asarray(self.pm["mean"]).size != 3:
                raise NotImplementedError('distn "norm" is implemented for the 3-state chain')
            if not (sd.ravel()[0] == sd).all():
                raise NotImplementedError("the device chain shares one sd across the states")
        elif self.distn == "binom":
            if m != 2 or np.asarray(self.pm["prob"]).size != 2 or not self.pn or "size" not in self.pn:
                raise NotImplementedError('distn "binom" needs 2 states, pm$prob[2] and pn$size')
        else:
            raise NotImplementedError(f'distn "{self.distn}" is not on the HoneyBADGER path')


def _run(obj: dthmm, viterbi, gamma, ll):
    if obj.distn == "norm":
        sd = float(np.asarray(obj.pm["sd"], dtype=np.float64).ravel()[0])
        return hmm_norm_host(obj.x, obj.pm["mean"], sd, obj.Pi, obj.delta, obj.seg_off,
                             viterbi=viterbi, gamma=gamma, ll=ll)
    return hmm_binom_host(obj.x, obj.pn["size"], obj.pm["prob"], obj.Pi, obj.delta, obj.seg_off,
                          rho=float(obj.pm.get("rho", 0.0)), viterbi=viterbi, gamma=gamma, ll=ll)


def Viterbi(obj: dthmm) -> np.ndarray:
    """HiddenMarkov::Viterbi(object): most likely state path, 1-based integers.

    Raises `_lib.UnderflowError` ("Problems With Underflow") exactly when the reference does: some
    final nu equals -Inf.
    """
    return _run(obj, True, False, False)["y"]


def forwardback(obj: dthmm) -> dict:
    """State posteriors u = exp(logalpha + logbeta - LL) and LL of HiddenMarkov::forwardback/Estep."""
    r = _run(obj, False, True, True)
    return {"u": r["gamma"], "LL": r["ll"]}


# --------------------------------------------------------------------------- device-resident (torch)
def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_gm(t, dtype, name):
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.dim() == 2
            and t.stride(1) == 1):
        raise ValueError(f"{name} must be a 2-D CUDA {dtype} tensor with unit column stride")


def alloc_rows(T, B, dtype, device, planes=None):
    """[T, B] (or [planes, T, B]) view of a buffer whose rows are padded to a multiple of 128 elements:
    the resident layout the chain kernels like best (every row starts on a 1 KB boundary for fp64)."""
    import torch
    ld = (B + 127) // 128 * 128
    if planes is None:
        return torch.empty((T, ld), dtype=dtype, device=device)[:, :B]
    return torch.empty((planes, T, ld), dtype=dtype, device=device)[:, :, :B]


def to_rows(t):
    """Copy a [T, B] CUDA tensor into the padded resident layout (no-op if it already is)."""
    if t.stride(0) % 128 == 0 and t.stride(1) == 1:
        return t
    out = alloc_rows(t.shape[0], t.shape[1], t.dtype, t.device)
    out.copy_(t)
    return out


# ---- warp-tiled resident layout: tensors of shape [ceil(B/32), T, 32] (tile, position, lane)
def tile_rows(t):
    """Gene-major [T, B] CUDA tensor -> warp-tiled [ntile, T, 32] (libhb_b200's hb_tile_dev)."""
    import torch
    if not (t.is_cuda and t.dim() == 2 and t.stride(1) == 1):
        raise ValueError("tile_rows needs a 2-D CUDA tensor with unit column stride")
    T, B = t.shape
    out = torch.empty(((B + 31) // 32, T, 32), dtype=t.dtype, device=t.device)
    _lib.check(_lib.load().hb_tile_dev(t.data_ptr(), t.stride(0), T, B, t.element_size(),
                                       out.data_ptr(), _stream()))
    return out


def untile_rows(t, B):
    """Warp-tiled [ntile, T, 32] (or [planes, ntile, T, 32]) -> gene-major [T, B] ([planes, T, B])."""
    import torch
    if t.dim() == 4:
        return torch.stack([untile_rows(p, B) for p in t])
    ntile, T, w = t.shape
    if w != 32 or not t.is_contiguous() or ntile != (B + 31) // 32:
        raise ValueError("not a warp-tiled tensor for this B")
    out = alloc_rows(T, B, t.dtype, t.device)
    _lib.check(_lib.load().hb_untile_dev(t.data_ptr(), T, B, t.element_size(), out.data_ptr(),
                                         out.stride(0), _stream()))
    return out


def _check_tiled(t, dtype, name):
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.dim() == 3
            and t.shape[2] == 32 and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous warp-tiled CUDA {dtype} tensor [ntile, T, 32]")


def hmm_norm_tiled_dev(x, B, sd, mean, Pi, delta, seg_off=None, viterbi=True, gamma=True, ll=False,
                       out=None):
    """As hmm_norm_dev on the warp-tiled layout: x [ntile, T, 32] (from tile_rows), B real columns.
    Returns dict(y=uint8 [ntile,T,32], gamma=[3,ntile,T,32], ll=[nseg,B])."""
    import torch
    lib = _lib.load()
    _check_tiled(x, torch.float64, "x")
    ntile, T, _ = x.shape
    if ntile != (B + 31) // 32:
        raise ValueError("x does not have ceil(B/32) tiles")
    seg = _seg(seg_off, T)
    nseg = seg.size - 1
    per_seg = int(sd.dim() == 2)
    mean, Pi, delta = _f64(mean), _f64(Pi), _f64(delta)
    out = dict(out or {})
    want_g = gamma
    if viterbi and "y" not in out:
        out["y"] = torch.empty((ntile, T, 32), dtype=torch.uint8, device=x.device)
    if want_g and "gamma" not in out:
        out["gamma"] = torch.empty((3, ntile, T, 32), dtype=torch.float64, device=x.device)
    if ll and "ll" not in out:
        out["ll"] = torch.empty((nseg, B), dtype=torch.float64, device=x.device)
    y, g, l = out.get("y"), out.get("gamma"), out.get("ll")
    rc = lib.hb_hmm_norm_tiled_dev(
        x.data_ptr(), T, B, _vp(seg), nseg, _vp(mean), sd.data_ptr(), per_seg, _vp(Pi), _vp(delta),
        _flags(viterbi, gamma, ll), y.data_ptr() if viterbi else None,
        g.data_ptr() if want_g else None, l.data_ptr() if ll else None, _stream())
    _lib.check(rc)
    return out


def forwardback_norm_cm_dev(x_cm, sd, mean, Pi, delta, seg_off=None, out=None):
    """Posteriors by the chain-resident kernel on CHAIN-MAJOR data: x_cm [B, T] contiguous fp64 CUDA tensor
    (R's column-major [T x B]); returns gamma [3, B, T] (R's [T x B x 3]).  `fbr_gate()` != 0 afterwards
    means a transfer matrix degenerated (extreme emission ratios): use hmm_norm_dev for that input."""
    import torch
    if not (x_cm.is_cuda and x_cm.dtype == torch.float64 and x_cm.dim() == 2 and x_cm.is_contiguous()):
        raise ValueError("x_cm must be a contiguous CUDA float64 tensor [B, T]")
    B, T = x_cm.shape
    seg = _seg(seg_off, T)
    per_seg = int(sd.dim() == 2)
    mean, Pi, delta = _f64(mean), _f64(Pi), _f64(delta)
    g = out if out is not None else torch.empty((3, B, T), dtype=torch.float64, device=x_cm.device)
    _lib.check(_lib.load().hb_forwardback_norm_cm_dev(x_cm.data_ptr(), T, B, _vp(seg), seg.size - 1, _vp(mean),
                                                      sd.data_ptr(), per_seg, _vp(Pi), _vp(delta), g.data_ptr(),
                                                      _stream()))
    return g


def fbr_gate() -> int:
    """1 if the last chain-resident launch met a degenerate transfer matrix; waits for the current stream."""
    g = C.c_int32(0)
    _lib.check(_lib.load().hb_fbr_stats_dev(_stream(), C.byref(g)))
    return int(g.value)


def pack_counts(x, size):
    """(size << 16) | x per element as an int32 tensor of the same shape (hb_pack_counts_dev); raises
    if a count does not fit 16 bits.  x, size: contiguous int32 CUDA tensors in any (same) layout."""
    import torch
    if not (x.is_cuda and x.dtype == torch.int32 and x.is_contiguous() and size.is_contiguous()
            and x.shape == size.shape and size.dtype == torch.int32):
        raise ValueError("pack_counts needs two contiguous int32 CUDA tensors of the same shape")
    out = torch.empty_like(x)
    _lib.check(_lib.load().hb_pack_counts_dev(x.data_ptr(), size.data_ptr(), x.numel(), out.data_ptr(),
                                              _stream()))
    return out


def hmm_binom_tiled_dev(x, size, B, prob, Pi, delta, seg_off=None, rho=0.0, viterbi=True, gamma=True,
                        ll=False, out=None):
    """As hmm_binom_dev on the warp-tiled layout: x, size int32 [ntile, T, 32]; size=None means x
    holds PACKED counts (pack_counts of the tiled planes): one 4-byte word per position."""
    import torch
    lib = _lib.load()
    _check_tiled(x, torch.int32, "x")
    packed = size is None
    if not packed:
        _check_tiled(size, torch.int32, "size")
    ntile, T, _ = x.shape
    if (not packed and x.shape != size.shape) or ntile != (B + 31) // 32:
        raise ValueError("x and size must be [ceil(B/32), T, 32]")
    seg = _seg(seg_off, T)
    nseg = seg.size - 1
    prob, Pi, delta = _f64(prob), _f64(Pi), _f64(delta)
    out = dict(out or {})
    want_g = gamma
    if viterbi and "y" not in out:
        out["y"] = torch.empty((ntile, T, 32), dtype=torch.uint8, device=x.device)
    if want_g and "gamma" not in out:
        out["gamma"] = torch.empty((2, ntile, T, 32), dtype=torch.float64, device=x.device)
    if ll and "ll" not in out:
        out["ll"] = torch.empty((nseg, B), dtype=torch.float64, device=x.device)
    y, g, l = out.get("y"), out.get("gamma"), out.get("ll")
    tail = (T, B, _vp(seg), nseg, _vp(prob), float(rho), _vp(Pi), _vp(delta), _flags(viterbi, gamma, ll),
            y.data_ptr() if viterbi else None, g.data_ptr() if want_g else None,
            l.data_ptr() if ll else None, _stream())
    if packed:
        rc = lib.hb_hmm_binom_packed_tiled_dev(x.data_ptr(), *tail)
    else:
        rc = lib.hb_hmm_binom_tiled_dev(x.data_ptr(), size.data_ptr(), *tail)
    _lib.check(rc)
    return out


def hmm_norm_dev(x, sd, mean, Pi, delta, seg_off=None, viterbi=True, gamma=True, ll=False,
                 out=None):
    """Gene-major x[T, B] (fp64 CUDA tensor, row stride = ld), sd [B] or [nseg, B] on the device.

    Enqueues on the current stream and returns dict(y=uint8 [T,B], gamma=[3,T,B], ll=[nseg,B])
    of device tensors (reused from `out` when given).  Does not synchronise; call
    `status_dev()` to surface HiddenMarkov's underflow error.
    """
    import torch
    lib = _lib.load()
    _check_gm(x, torch.float64, "x")
    T, B = x.shape
    seg = _seg(seg_off, T)
    nseg = seg.size - 1
    per_seg = int(sd.dim() == 2)
    if not (sd.is_cuda and sd.dtype == torch.float64 and sd.is_contiguous()):
        raise ValueError("sd must be a contiguous CUDA float64 tensor")
    mean, Pi, delta = _f64(mean), _f64(Pi), _f64(delta)
    out = dict(out or {})
    want_g = gamma
    if viterbi and "y" not in out:
        out["y"] = alloc_rows(T, B, torch.uint8, x.device)
    if want_g and "gamma" not in out:
        out["gamma"] = alloc_rows(T, B, torch.float64, x.device, planes=3)
    if ll and "ll" not in out:
        out["ll"] = torch.empty((nseg, B), dtype=torch.float64, device=x.device)
    y, g, l = out.get("y"), out.get("gamma"), out.get("ll")
    rc = lib.hb_hmm_norm_dev(
        x.data_ptr(), x.stride(0), T, B, _vp(seg), nseg, _vp(mean), sd.data_ptr(), per_seg,
        _vp(Pi), _vp(delta), _flags(viterbi, gamma, ll),
        y.data_ptr() if viterbi else None, y.stride(0) if viterbi else 0,
        g.data_ptr() if want_g else None, g.stride(1) if want_g else 0,
        l.data_ptr() if ll else None, _stream())
    _lib.check(rc)
    return out


def hmm_binom_dev(x, size, prob, Pi, delta, seg_off=None, rho=0.0, viterbi=True, gamma=True,
                  ll=False, out=None):
    """Gene-major int32 x[T, B], size[T, B] CUDA tensors with the same row stride."""
    import torch
    lib = _lib.load()
    _check_gm(x, torch.int32, "x")
    _check_gm(size, torch.int32, "size")
    if x.shape != size.shape or x.stride(0) != size.stride(0):
        raise ValueError("x and size must share shape and row stride")
    T, B = x.shape
    seg = _seg(seg_off, T)
    nseg = seg.size - 1
    prob, Pi, delta = _f64(prob), _f64(Pi), _f64(delta)
    out = dict(out or {})
    want_g = gamma
    if viterbi and "y" not in out:
        out["y"] = alloc_rows(T, B, torch.uint8, x.device)
    if want_g and "gamma" not in out:
        out["gamma"] = alloc_rows(T, B, torch.float64, x.device, planes=2)
    if ll and "ll" not in out:
        out["ll"] = torch.empty((nseg, B), dtype=torch.float64, device=x.device)
    y, g, l = out.get("y"), out.get("gamma"), out.get("ll")
    rc = lib.hb_hmm_binom_dev(
        x.data_ptr(), size.data_ptr(), x.stride(0), T, B, _vp(seg), nseg, _vp(prob), float(rho),
        _vp(Pi), _vp(delta), _flags(viterbi, gamma, ll),
        y.data_ptr() if viterbi else None, y.stride(0) if viterbi else 0,
        g.data_ptr() if want_g else None, g.stride(1) if want_g else 0,
        l.data_ptr() if ll else None, _stream())
    _lib.check(rc)
    return out


def chain_sd_dev(x, seg_off=None, per_seg=False):
    """sd() of every chain column (and segment) of a gene-major fp64 CUDA tensor."""
    import torch
    lib = _lib.load()
    _check_gm(x, torch.float64, "x")
    T, B = x.shape
    seg = _seg(seg_off, T)
    nseg = seg.size - 1 if per_seg else 1
    sd = torch.empty((nseg, B) if per_seg else (B,), dtype=torch.float64, device=x.device)
    _lib.check(lib.hb_chain_sd_dev(x.data_ptr(), x.stride(0), T, B, _vp(seg), seg.size - 1,
                                   int(per_seg), sd.data_ptr(), _stream()))
    return sd


def norm3_rerun_chains() -> int:
    """Chains of the last Gaussian launch whose Viterbi recursion was re-run with the literal
    reference arithmetic (speculate-and-verify kernel); waits for the current stream."""
    n = C.c_int64(0)
    _lib.check(_lib.load().hb_norm3_stats_dev(_stream(), C.byref(n)))
    return int(n.value)


def norm3_flagged_chains(B: int, cap: int = 16384):
    """(segment, column) pairs of the chains the last Gaussian launch re-ran with the literal arithmetic
    (first `cap` of them); waits for the current stream."""
    buf = np.zeros(cap, dtype=np.int64)
    n = C.c_int64(0)
    _lib.check(_lib.load().hb_norm3_flagged_dev(_stream(), _vp(buf), cap, C.byref(n)))
    ids = buf[: n.value]
    return np.stack([ids // B, ids % B], axis=1)


def last_kernels() -> dict:
    """Which kernels the last chain launches used: norm3_fast (1 speculate-and-verify, 0 literal),
    binom2_lean (1 lean kernel + gated general kernel, 0 general only); -1 = none yet."""
    a, b = C.c_int32(-1), C.c_int32(-1)
    _lib.check(_lib.load().hb_last_kernels(C.byref(a), C.byref(b)))
    return {"norm3_fast": int(a.value), "binom2_lean": int(b.value)}


def binom2_general_rerun() -> int:
    """1 if the last binomial launch was redone by the general (relative-exponent) kernel."""
    n = C.c_int32(0)
    _lib.check(_lib.load().hb_binom2_stats_dev(_stream(), C.byref(n)))
    return int(n.value)


def viterbi_long_logb_dev(logb, Pi, delta, block_len=0):
    """HiddenMarkov::Viterbi for a few LONG chains on precomputed log-emissions, parallel along the
    chain and bit-exact (csrc/hb_longchain.cuh).  logb: CUDA float64 [K, T, S] (S = 2 or 3), the
    reference's own log-densities; returns CUDA int32 [K, T], 1-based states.  No synchronisation."""
    import torch
    if logb.dtype != torch.float64 or not logb.is_cuda or logb.dim() != 3 or not logb.is_contiguous():
        raise ValueError("logb must be a contiguous CUDA float64 tensor [K, T, S]")
    K, T, S = logb.shape
    Pi = _f64(Pi).reshape(S * S)
    delta = _f64(delta).reshape(S)
    y = torch.empty((K, T), dtype=torch.int32, device=logb.device)
    _lib.check(_lib.load().hb_viterbi_long_logb_dev(logb.data_ptr(), S, T, K, _vp(Pi), _vp(delta),
                                                    int(block_len), y.data_ptr(), _stream()))
    return y


def forwardback_long_logb_dev(logb, Pi, delta, block_len=0, ll=True):
    """State posteriors (HiddenMarkov::forwardback semantics) of a few LONG chains, parallel along the
    chain.  logb: CUDA float64 [K, T, S]; returns (gamma [K, T, S], ll [K] or None) on the device."""
    import torch
    if logb.dtype != torch.float64 or not logb.is_cuda or logb.dim() != 3 or not logb.is_contiguous():
        raise ValueError("logb must be a contiguous CUDA float64 tensor [K, T, S]")
    K, T, S = logb.shape
    Pi = _f64(Pi).reshape(S * S)
    delta = _f64(delta).reshape(S)
    gamma = torch.empty((K, T, S), dtype=torch.float64, device=logb.device)
    llv = torch.empty(K, dtype=torch.float64, device=logb.device) if ll else None
    _lib.check(_lib.load().hb_forwardback_long_logb_dev(logb.data_ptr(), S, T, K, _vp(Pi), _vp(delta),
                                                        int(block_len), gamma.data_ptr(),
                                                        llv.data_ptr() if ll else None, _stream()))
    return gamma, llv


def longchain_stats() -> dict:
    """Counters of the last long-chain launch; waits for the current stream."""
    out = (C.c_int64 * 5)()
    _lib.check(_lib.load().hb_longchain_stats_dev(_stream(), out))
    return {"chains": out[0], "blocks": out[1], "sequential_chains": out[2], "repair_rounds": out[3],
            "literal_blocks": out[4]}


def set_longchain_mode(mode: int):
    """Pooled rounds: 0 allele round on the long-chain path (default), 1 neither, 2 both."""
    _lib.check(_lib.load().hb_set_longchain_mode(int(mode)))


def set_ward_mode(mode: int):
    """Ward linkage kernel: 0 cluster of CTAs from 256 cells up (default), 1 one block, 2 always the cluster."""
    _lib.check(_lib.load().hb_set_ward_mode(int(mode)))


def set_dist_mode(mode: int):
    """dist kernel: 0 pipelined (default), 1 plain."""
    _lib.check(_lib.load().hb_set_dist_mode(int(mode)))


def status_dev():
    """Wait for the current stream and raise UnderflowError if the last launch hit nu == -Inf."""
    _lib.check(_lib.load().hb_status_dev(_stream()))
