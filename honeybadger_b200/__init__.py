"""honeybadger_b200 — B200 (sm_100a) HMM segmentation for HoneyBADGER.

One hot path, built from scratch: the Viterbi / forward-backward segmentation behind
`calcGexpCnvBoundaries` (R/HoneyBADGER.R:1091-1316, R/HoneyBADGER_gexp.R:462-584) and
`calcAlleleCnvBoundaries` (R/HoneyBADGER.R:1336-1578, R/HoneyBADGER_allele.R:543-651), as
hand-written CUDA behind a plain C ABI (include/hb_b200.h, libhb_b200.so).  The Python modules
mirror the reference's R interface for that path; nothing here computes on the CPU.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
