"""CPU: numerics of the device helpers, through the host build of the same headers (tests/emul)."""
from __future__ import annotations

import ctypes as C
import random
from fractions import Fraction as F

import numpy as np


def test_div_by_const_is_correctly_rounded_on_hardest_cases(emul):
    """Markstein's q' = q + (d - b q) y with y = RN(1/b): the quotients closest to rounding
    midpoints (found with modular inverses) must still round correctly."""
    rnd = random.Random(1)
    n = fails = naive = 0
    for trial in range(600):
        B = rnd.randrange(2 ** 52, 2 ** 53) | 1
        if trial % 10 == 0:
            B = 2 ** 53 - 1 - 2 * rnd.randrange(0, 50)
        b = B * 2.0 ** -52 * 2.0 ** rnd.randrange(-3, 3)
        inv = pow(B, -1, 2 ** 53)
        for delta in range(-9, 10, 2):
            M = (delta * inv) % 2 ** 53 + 2 ** 53
            num = B * M - delta
            if num % 2 ** 53:
                continue
            A = num >> 53
            if A >= 2 ** 53 and A % 2:
                continue
            a = float(A) * 2.0 ** -52 * (b / (B * 2.0 ** -52))
            if F(a) != F(A, 2 ** 52) * F(b) / F(B, 2 ** 52):
                continue
            exact = float(F(a) / F(b))
            n += 1
            fails += emul.emul_div_by_const(a, b) != exact
            naive += a * (1.0 / b) != exact
    assert n > 2000 and fails == 0 and naive > n // 4
    rng = np.random.default_rng(0)
    for a, b in zip(rng.uniform(-20, 20, 20000), rng.uniform(0.2, 3.5, 20000)):
        assert emul.emul_div_by_const(float(a), float(b)) == float(a) / float(b)


def test_exp_pair_accuracy(emul):
    import mpmath as mp
    mp.mp.dps = 40
    ep, em = C.c_double(), C.c_double()
    worst = 0.0
    for u in np.concatenate([np.linspace(-40, 40, 801), [1e-9, -1e-9, 0.0, 300.0, -300.0, 345.0]]):
        emul.emul_exp_pair(float(u), C.byref(ep), C.byref(em))
        for got, ref in ((ep.value, mp.e ** mp.mpf(float(u))), (em.value, mp.e ** mp.mpf(-float(u)))):
            worst = max(worst, abs(float((mp.mpf(got) - ref) / ref)))
    assert worst < 4e-16


def test_x87_emulation_reproduces_r_mean_and_sd(emul, oracle):
    rng = np.random.default_rng(4)
    for n in (2, 3, 5, 64, 1001, 6082):
        for shift in (0.0, 7.5):
            x = np.ascontiguousarray(rng.normal(size=n) * 2.6 + shift)
            p = x.ctypes.data_as(C.c_void_p)
            assert emul.emul_x87_mean(p, n) == oracle.r_mean(x)
            assert emul.emul_x87_sd(p, n) == oracle.r_sd(x)


def test_x87_add_matches_the_hardware_step_by_step(emul):
    """x80_add (fast path and general path) against the host's x87 unit after every single addition:
    same-sign growth, mixed signs, cancellation, exact ties, terms far below / above the accumulator."""
    rng = np.random.default_rng(17)
    cases = []
    cases.append(rng.gamma(2.0, 1.5, 20000))                                  # all positive (rowMeans)
    cases.append(rng.normal(size=20000) * 2.6)                                # mixed signs
    cases.append(rng.normal(size=5000) * 10.0 ** rng.integers(-12, 12, 5000))  # wide exponent range
    cases.append(np.ldexp(rng.integers(1, 1 << 20, 5000).astype(float), rng.integers(-70, 70, 5000)))
    cases.append(np.array([1.0, 2.0 ** -64, 2.0 ** -64, -1.0, 2.0 ** -63, 1.0, -(1.0 - 2.0 ** -53), 3.0, -3.0]))
    t = np.ldexp(rng.integers(1, 1 << 30, 4000).astype(float), -40)            # few significant bits: many ties
    cases.append(np.concatenate([[float(1 << 23)], t, -t]))
    cases.append(np.concatenate([np.full(3000, 0.1), np.full(3000, -0.1), [1e300, -1e300, 1e-300]]))
    for x in cases:
        x = np.ascontiguousarray(x, dtype=np.float64)
        p = x.ctypes.data_as(C.c_void_p)
        assert emul.emul_x87_addcheck(p, x.size, 0.0, 0) == 0
        assert emul.emul_x87_addcheck(p, x.size, float(x.mean()), 1) == 0


def _pair_cover(emul, reset=True):
    cov = (C.c_int64 * 8)()
    emul.emul_x87_pair_cover(cov, 1 if reset else 0)
    return list(cov)


def test_x87_pair_accumulation_matches_the_hardware_step_by_step(emul):
    """The double-pair form of the 80-bit running sums (hb_x87_pair.cuh, what pool_accum_*_kernel runs)
    against the host's x87 unit after every addition, both passes of R's mean: s += x and s += (x - c80)
    with an 80-bit centre.  Same cases as the integer form above plus tie-heavy and cancelling ones."""
    rng = np.random.default_rng(23)
    cases = [(rng.gamma(2.0, 1.5, 100000), 1.3, 1e-18),
             (rng.normal(size=100000) * 2.6, 0.01, 3e-20),
             (rng.normal(size=30000) * 10.0 ** rng.integers(-12, 12, 30000), 1.0, 1e-17),
             (np.ldexp(rng.integers(1, 1 << 20, 30000).astype(float), rng.integers(-70, 70, 30000)), 3.0, 2.0 ** -62),
             (np.array([1.0, 2.0 ** -64, 2.0 ** -64, -1.0, 2.0 ** -63, 1.0, -(1.0 - 2.0 ** -53), 3.0, -3.0]), 1.0, 2.0 ** -60),
             (rng.integers(0, 50, 50000).astype(float), 3.0, 0.0),
             (np.log2(rng.gamma(2, 2, 50000) + 1) - 2.0, 0.3, 1e-19),
             (np.concatenate([np.full(3000, 0.1), np.full(3000, -0.1), [1e300, -1e300, 1e-300]]), 0.1, 0.0)]
    t = np.ldexp(rng.integers(1, 1 << 30, 20000).astype(float), -40)
    cases.append((np.concatenate([[float(1 << 23)], t, -t]), 0.5, 2.0 ** -64))
    for k in (0, 5, 11, 30):  # accumulator near 2^k, terms with bits right at half a unit of the result
        adds = np.ldexp(rng.integers(-(1 << 14), 1 << 14, 30000).astype(float), k - 64 - rng.integers(0, 4, 30000))
        cases.append((np.concatenate([[2.0 ** k], adds]), 2.0 ** k, 2.0 ** (k - 63)))
    pw = []
    for _ in range(10000):  # powers of two pulled just below themselves (the exponent of the sum drops)
        k = int(rng.integers(-20, 20))
        pw += [2.0 ** k, -2.0 ** (k - int(rng.integers(50, 70))), -2.0 ** k * (1 - 2.0 ** -int(rng.integers(1, 53)))]
    cases.append((np.array(pw), 1.0, -2.0 ** -64))
    y = rng.normal(size=20000)
    z = np.empty(40000)
    z[0::2] = y
    z[1::2] = -y + np.ldexp(rng.integers(-8, 8, 20000).astype(float), -70)  # sums that cancel to (almost) nothing
    cases.append((z, 0.0, 2.0 ** -70))
    for x, c, c2 in cases:
        x = np.ascontiguousarray(x, dtype=np.float64)
        for mode in (0, 1):
            assert emul.emul_x87_paircheck(x.ctypes.data_as(C.c_void_p), x.size, c, c2, mode, None) == 0


def test_x87_pair_add_on_explicit_64_bit_operands(emul):
    """x80p_add / x80p_add_double on operands given as 64-bit significands: sparse bit patterns (exact sums,
    ties, long carries), full ones, exponents from equal to 130 apart, both signs."""
    rng = np.random.default_rng(29)
    N = 60000
    _pair_cover(emul)

    def sparse(keep):
        m = np.zeros(N, dtype=np.uint64)
        for _ in range(keep):
            m |= np.uint64(1) << rng.integers(0, 64, N).astype(np.uint64)
        return m

    def full():
        return rng.integers(0, 1 << 63, N, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, N).astype(np.uint64)

    def go(ms, es, md, ed):
        ns = rng.integers(0, 2, N).astype(np.uint8)
        nd = rng.integers(0, 2, N).astype(np.uint8)
        ms, md = np.ascontiguousarray(ms, np.uint64), np.ascontiguousarray(md, np.uint64)
        es, ed = np.ascontiguousarray(es, np.int32), np.ascontiguousarray(ed, np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        for dbl in (0, 1):
            assert emul.emul_x87_pair_add80(p(ms), p(es), p(ns), p(md), p(ed), p(nd), N, dbl) == 0

    for spread in (2, 12, 40, 64, 70, 130):
        for k in (1, 2, 3, 5):
            go(sparse(k), rng.integers(-3, 3, N), sparse(k), rng.integers(-spread, 1, N))
        go(full(), rng.integers(-3, 3, N), full(), rng.integers(-spread, spread // 8 + 1, N))
        go(full(), rng.integers(-3, 3, N), sparse(3), rng.integers(-spread, 1, N))
        go(sparse(2), rng.integers(-3, 3, N), full(), rng.integers(-spread, 1, N))
    low = (np.uint64(1) << np.uint64(63)) | rng.integers(0, 4096, N).astype(np.uint64)
    dm = rng.integers(0, 1 << 16, N).astype(np.uint64) << rng.integers(0, 48, N).astype(np.uint64)
    go(low, np.zeros(N), dm, rng.integers(-70, -40, N))
    cov = _pair_cover(emul)
    assert all(v > 0 for v in cov), cov  # every case of the rounding step and both integer fall-backs were taken


def test_x87_pair_fast_path_covers_expression_like_data(emul):
    """On data shaped like the pooled-mean inputs no step leaves the branch-free path (a fall-back is
    correct but serialises a warp, so this is the performance contract of the kernel)."""
    rng = np.random.default_rng(31)
    zi = rng.gamma(2.0, 1.5, 200000) * (rng.random(200000) < 0.3)   # zero-inflated, as un-normalised expression is
    zi[:5000] = 0.0                                                # and starting with a long run of zeros
    for x in (rng.normal(size=200000) * 2.6, rng.gamma(2.0, 1.5, 200000), np.log2(rng.gamma(2, 2, 200000) + 1) - 2, zi,
              np.zeros(1000)):
        x = np.ascontiguousarray(x)
        m = float(np.mean(x))
        for mode in (0, 1):
            _pair_cover(emul)
            assert emul.emul_x87_paircheck(x.ctypes.data_as(C.c_void_p), x.size, m, m * 2.0 ** -55, mode, None) == 0
            cov = _pair_cover(emul)
            assert cov[6] == 0 and cov[7] == 0, cov


def test_x87_block_accumulation_matches_the_literal_one(emul, oracle):
    """x80_align_term / x80_block_add (pooled_sd_kernel's running sums: one integer addition per term while the sum
    stays in its binade, x80_add for the terms that fail a condition) against the host's x87 unit block by block,
    against the term-by-term emulation, and against R's sd() — and most terms must take the integer path."""
    rng = np.random.default_rng(41)
    cases = [rng.normal(size=20000) * 0.3, rng.normal(size=20000) * 0.3 + 0.05, rng.gamma(2.0, 1.5, 20000),
             rng.normal(size=6082) * 2.6, rng.normal(size=30000) * 10.0 ** rng.integers(-12, 12, 30000),
             rng.integers(0, 50, 20000).astype(float),
             np.ldexp(rng.integers(1, 1 << 20, 20000).astype(float), rng.integers(-70, 70, 20000)),
             np.concatenate([np.zeros(100), rng.normal(size=1000), np.zeros(50), rng.normal(size=1000)])]
    t = np.ldexp(rng.integers(1, 1 << 30, 10000).astype(float), -40)
    cases.append(np.concatenate([[float(1 << 23)], t, -t]))
    for k in (0, 11, 30):
        adds = np.ldexp(rng.integers(-(1 << 14), 1 << 14, 20000).astype(float), k - 64 - rng.integers(0, 4, 20000))
        cases.append(np.concatenate([[2.0 ** k], adds]))
    pw = []
    for _ in range(5000):
        k = int(rng.integers(-20, 20))
        pw += [2.0 ** k, -2.0 ** (k - int(rng.integers(50, 70))), -2.0 ** k * (1 - 2.0 ** -int(rng.integers(1, 53)))]
    cases.append(np.array(pw))
    cases += [rng.normal(size=n) * 2.6 + 7.5 for n in (2, 3, 5, 31, 32, 33, 64, 65, 1001)]
    for x in cases:
        x = np.ascontiguousarray(x, dtype=np.float64)
        p = x.ctypes.data_as(C.c_void_p)
        f = C.c_int64(0)
        got = emul.emul_x87_sd_blocked(p, x.size, C.byref(f))
        assert got == emul.emul_x87_sd(p, x.size) == oracle.r_sd(x)
        assert emul.emul_x87_blockcheck(p, x.size, 0.0, 0, None) == 0
        assert emul.emul_x87_blockcheck(p, x.size, float(x.mean()), 1, None) == 0
    x = np.ascontiguousarray(rng.normal(size=20000) * 0.3)     # a pooled expression vector: the case the kernel serves
    f = C.c_int64(0)
    emul.emul_x87_sd_blocked(x.ctypes.data_as(C.c_void_p), x.size, C.byref(f))
    assert f.value > 0.85 * 3 * x.size


def test_binom_logpmf_host_and_device_restatements_match_oracle(emul, oracle):
    for n in list(range(0, 70)) + [100, 255, 256, 511, 512, 1245, 5000]:
        for x in sorted({0, 1, n // 3, n // 2, n - 1, n}):
            if 0 <= x <= n:
                for p in (0.1, 0.45):
                    ref = oracle.dbinom(x, n, p, True)
                    assert emul.emul_binom_logpmf_host(x, n, p) == ref
                    assert emul.emul_binom_logpmf_dev(x, n, p) == ref  # same libm on the host build


def test_emission_table_against_exact_arithmetic_for_every_pair(emul):
    """The product's emission table (hb_host_math.h: binom_logpmf, what prepare_lut uploads) for EVERY
    (x, n <= 511) pair and both states against the exact log-pmf from mpmath — an arithmetic-independent anchor
    that a shared mis-recollection of nmath (the oracle restates the same published algorithm) could not pass.
    Loader's saddle point is not correctly rounded: the error is measured in ulps of the exact value, its
    distribution printed, and bounded (a wrong Stirling constant or bd0 branch shows up as >= 1e-10)."""
    import mpmath as mp
    NMAX = 511
    worst = {}
    hist = np.zeros(6, dtype=np.int64)  # <=0.5, <=1, <=2, <=4, <=16, >16 ulp
    with mp.workdps(40):
        lfact = [mp.mpf(0)]
        for i in range(1, NMAX + 1):
            lfact.append(lfact[-1] + mp.log(i))
        for p in (0.1, 0.45):
            lp, lq = mp.log(mp.mpf(p)), mp.log(1 - mp.mpf(p))
            for n in range(NMAX + 1):
                for x in range(n + 1):
                    exact = lfact[n] - lfact[x] - lfact[n - x] + x * lp + (n - x) * lq
                    got = emul.emul_binom_logpmf_host(float(x), float(n), p)
                    ex = float(exact)
                    if ex == 0.0:
                        assert got == 0.0
                        hist[0] += 1
                        continue
                    ulps = abs(float((mp.mpf(got) - exact) / mp.mpf(np.spacing(abs(ex)))))
                    b = 0 if ulps <= 0.5 else 1 if ulps <= 1 else 2 if ulps <= 2 else 3 if ulps <= 4 else 4 if ulps <= 16 else 5
                    hist[b] += 1
                    if ulps > worst.get(p, (0,))[0]:
                        worst[p] = (ulps, x, n)
    total = int(hist.sum())
    print(f"\nemission table vs exact: {total} entries; <=0.5 ulp {hist[0]}, <=1 {hist[1]}, <=2 {hist[2]}, "
          f"<=4 {hist[3]}, <=16 {hist[4]}, >16 {hist[5]}; worst {worst}")
    assert total == 2 * (NMAX + 1) * (NMAX + 2) // 2
    # measured: worst 38 ulp (8e-15 relative, at the mode of deep counts where stirlerr / bd0 terms cancel);
    # a wrong constant or branch would be off by >= 1e5 ulp
    assert all(w[0] <= 64 for w in worst.values()), worst
    assert hist[5] < 0.01 * total and hist[0] + hist[1] + hist[2] > 0.6 * total, hist
