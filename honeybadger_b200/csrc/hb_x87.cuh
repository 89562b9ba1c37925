// hb_x87.cuh — exact integer emulation of x87 80-bit extended arithmetic (64-bit significand,
// round-to-nearest-even), enough of it to reproduce base R's `mean()` and `var()`/`sd()` bit for
// bit on a GPU.
//
// R accumulates in LDOUBLE (`long double`, x87 extended on x86-64 Linux): summary.c real_mean and
// cov.c cov_complete1, reached from R/HoneyBADGER.R:1141 (`apply(.,1,mean)`)
// and :1149 (`sd(mat.smooth)`).  Every add / subtract / multiply / divide there rounds to 64
// significand bits, so the result depends on the summation order and cannot be reproduced by
// "more accurate" arithmetic — it has to be replayed.  Values are finite and far from the
// exponent limits on this path; infinities, NaNs and denormal extended results are not handled.
#pragma once
#include "hb_device.cuh"

namespace hb {

struct x80 {
    uint64_t m;  // significand, msb set unless zero
    int32_t e;   // value = (-1)^neg * m * 2^(e - 63)
    uint32_t neg;
};

HB_HD int clz64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}

HB_HD x80 x80_zero()
{
    x80 r;
    r.m = 0;
    r.e = 0;
    r.neg = 0;
    return r;
}

HB_HD x80 x80_from_double(double d)
{
    uint64_t u = ((uint64_t)(uint32_t)hi32(d) << 32) | (uint32_t)lo32(d);
    x80 r;
    r.neg = (uint32_t)(u >> 63);
    int32_t ef = (int32_t)((u >> 52) & 0x7ff);
    uint64_t fr = u & 0xfffffffffffffull;
    if (ef == 0) {
        if (fr == 0) {
            r.m = 0;
            r.e = 0;
            return r;
        }
        int lz = clz64(fr);
        r.m = fr << lz;
        r.e = -1022 - 52 + (63 - lz);
        return r;
    }
    r.m = (1ull << 63) | (fr << 11);
    r.e = ef - 1023;
    return r;
}

HB_HD x80 x80_from_u32(uint32_t n)
{
    x80 r;
    r.neg = 0;
    if (n == 0) {
        r.m = 0;
        r.e = 0;
        return r;
    }
    int lz = clz64((uint64_t)n);
    r.m = (uint64_t)n << lz;
    r.e = 63 - lz;
    return r;
}

// Round (hi:lo + sticky) to 64 bits, hi has its msb set.  Returns the significand and bumps e on carry.
HB_HD uint64_t x80_round(uint64_t hi, uint64_t lo, bool sticky, int32_t &e)
{
    const uint64_t half = 1ull << 63;
    bool up = (lo > half) || (lo == half && (sticky || (hi & 1ull)));
    if (up) {
        hi += 1;
        if (hi == 0) {
            hi = half;
            e += 1;
        }
    }
    return hi;
}

// General case of a + b (any magnitudes, cancellation, far-apart exponents).
HB_HD x80 x80_add_general(x80 a, x80 b);

// a + b.  The accumulations this file exists for add a term to a much larger running sum, so the
// common case is handled first and in 64-bit registers: a's exponent exceeds b's by d in [0, 63]
// (d >= 2 when the signs differ, so at most one leading bit cancels).  b's significand is cut at the
// 2^-d point into an integer part (added to / subtracted from a.m) and a 64-bit remainder that holds
// every shifted-out bit, i.e. the result is still the exactly rounded one.
HB_HD x80 x80_add(x80 a, x80 b)
{
    const int32_t d = a.e - b.e;
    const bool same = a.neg == b.neg;
    if (a.m != 0 && b.m != 0 && d < 64 && d >= (same ? 0 : 2)) {
        const uint64_t half = 1ull << 63;
        const uint64_t bhi = b.m >> d, blo = d ? (b.m << (64 - d)) : 0ull;
        x80 r;
        r.neg = a.neg;
        r.e = a.e;
        uint64_t m, lo;
        bool sticky = false;
        if (same) {
            m = a.m + bhi;
            lo = blo;
            if (m < a.m) { // carry out of bit 63: renormalise by one bit
                sticky = (lo & 1ull) != 0;
                lo = (lo >> 1) | (m << 63);
                m = half | (m >> 1);
                r.e += 1;
            }
        } else {
            lo = 0ull - blo;
            m = a.m - bhi - (blo != 0 ? 1ull : 0ull);
            if (!(m >> 63)) { // one leading bit cancelled
                m = (m << 1) | (lo >> 63);
                lo <<= 1;
                r.e -= 1;
            }
        }
        if (lo > half || (lo == half && (sticky || (m & 1ull)))) {
            m += 1;
            if (m == 0) {
                m = half;
                r.e += 1;
            }
        }
        r.m = m;
        return r;
    }
    return x80_add_general(a, b);
}

HB_HD x80 x80_add_general(x80 a, x80 b)
{
    if (a.m == 0) return b;
    if (b.m == 0) return a;
    if (a.e < b.e || (a.e == b.e && a.m < b.m)) {
        x80 t = a;
        a = b;
        b = t;
    }
    // 128-bit significands with the msb at bit 126: A = a.m << 63
    uint64_t ahi = a.m >> 1, alo = a.m << 63;
    uint64_t bhi, blo;
    bool sticky = false;
    int32_t d = a.e - b.e;
    {
        uint64_t h = b.m >> 1, l = b.m << 63;
        if (d == 0) {
            bhi = h;
            blo = l;
        } else if (d < 64) {
            sticky = (l << (64 - d)) != 0;
            blo = (l >> d) | (h << (64 - d));
            bhi = h >> d;
        } else if (d < 128) {
            int s = d - 64;
            sticky = (l != 0) || (s ? ((h << (64 - s)) != 0) : false);
            blo = s ? (h >> s) : h;
            bhi = 0;
        } else {
            sticky = true;
            blo = 0;
            bhi = 0;
        }
    }
    x80 r;
    r.neg = a.neg;
    r.e = a.e;
    uint64_t rhi, rlo;
    if (a.neg == b.neg) {
        rlo = alo + blo;
        rhi = ahi + bhi + (rlo < alo ? 1 : 0);
        if (rhi >> 63) { // carry into bit 127
            r.e += 1;
        } else {
            rhi = (rhi << 1) | (rlo >> 63);
            rlo <<= 1;
        }
        r.m = x80_round(rhi, rlo, sticky, r.e);
        return r;
    }
    // |a| >= |b|: A - B_trunc - sticky, remembering that the true value is a bit above
    uint64_t bl2 = blo + (sticky ? 1 : 0);
    uint64_t bh2 = bhi + ((sticky && bl2 == 0) ? 1 : 0);
    rlo = alo - bl2;
    rhi = ahi - bh2 - (alo < bl2 ? 1 : 0);
    if (rhi == 0 && rlo == 0) {
        if (!sticky) return x80_zero();
    }
    int lz;
    if (rhi) {
        lz = clz64(rhi);
        if (lz) {
            rhi = (rhi << lz) | (rlo >> (64 - lz));
            rlo <<= lz;
        }
    } else {
        lz = 64 + clz64(rlo);
        int s = lz - 64;
        rhi = s ? (rlo << s) : rlo;
        rlo = 0;
    }
    r.e = a.e + 1 - lz; // msb was at bit 126 <=> lz == 1
    r.m = x80_round(rhi, rlo, sticky, r.e);
    return r;
}

HB_HD x80 x80_neg(x80 a)
{
    a.neg ^= 1u;
    return a;
}
HB_HD x80 x80_sub(x80 a, x80 b) { return x80_add(a, x80_neg(b)); }

HB_HD uint64_t umul64hi_(uint64_t a, uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

HB_HD x80 x80_mul(x80 a, x80 b)
{
    if (a.m == 0 || b.m == 0) return x80_zero();
    uint64_t hi = umul64hi_(a.m, b.m), lo = a.m * b.m;
    x80 r;
    r.neg = a.neg ^ b.neg;
    if (hi >> 63) {
        r.e = a.e + b.e + 1;
    } else {
        hi = (hi << 1) | (lo >> 63);
        lo <<= 1;
        r.e = a.e + b.e;
    }
    r.m = x80_round(hi, lo, false, r.e);
    return r;
}

// a / n for an integer 1 <= n < 2^32 (exactly what `s /= n` does after the int -> LDOUBLE convert).
HB_HD x80 x80_div_u32(x80 a, uint32_t n)
{
    if (a.m == 0) return a;
    // Q = floor(a.m * 2^64 / n) by schoolbook division in 32-bit limbs
    uint64_t q2 = a.m / n, r = a.m % n;
    uint64_t t = r << 32;
    uint64_t q1 = t / n;
    r = t % n;
    t = r << 32;
    uint64_t q0 = t / n;
    r = t % n;
    uint64_t qhi = q2, qlo = (q1 << 32) | q0; // q1 < 2^32 because r < n <= 2^32
    bool sticky = r != 0;
    int lz = clz64(qhi); // qhi >= 2^31 since a.m >= 2^63 and n < 2^32
    if (lz) {
        qhi = (qhi << lz) | (qlo >> (64 - lz));
        qlo <<= lz;
    }
    x80 res;
    res.neg = a.neg;
    res.e = a.e - lz;
    res.m = x80_round(qhi, qlo, sticky, res.e);
    return res;
}

// (double) of an extended value: one RNE rounding 64 -> 53 bits.
// ---- a block of terms added to a running sum as an integer prefix sum ---------------------------------
// While the running sum stays inside one binade [2^e, 2^(e+1)) it is an integer multiple m of q = 2^(e-63),
// and sum <- RN64(sum + term) is m <- m +- RN(term / q) — provided the term is not exactly half-way
// between two multiples of q (then the parity of the sum decides) and the exact sum stays inside the
// binade (otherwise the result is rounded on another grid).  So a block of terms can be rounded to the
// grid independently (x80_align_term: any lane can do that for its own term, knowing only the sum's
// exponent and sign) and added with an integer PREFIX SUM; the prefix sums say at which term, if any, the
// sum leaves the binade, and that term (or a tie, or a term too large for the grid) is added literally
// with x80_add, after which the rest of the block is aligned again.  Same results as x80_add term by term
// (tests/test_numerics.py).  pooled_sd_kernel runs the prefix sum across the lanes of a warp
// (x80_block_add_warp below); x80_block_add is the same procedure written as loops, for the host tests.
HB_HD bool x80_align_term(x80 t, int32_t acc_e, uint32_t acc_neg, int64_t &s)
{
    s = 0;
    if (t.m == 0) return true;
    const int32_t d = acc_e - t.e; // term / q = t.m / 2^d
    if (d < 1) return false;
    if (d >= 65) return true; // below q / 2: rounds to nothing, cannot be a tie
    uint64_t ip, rem, half;
    if (d == 64) {
        ip = 0;
        rem = t.m;
        half = 1ull << 63;
    } else {
        ip = t.m >> d;
        rem = t.m & ((1ull << d) - 1);
        half = 1ull << (d - 1);
    }
    if (rem == half) return false; // a tie on the term alone
    if (rem > half) ip += 1;
    if (ip >> 62) return false; // fits the signed prefix sum (whose magnitude is guarded separately)
    s = (t.neg == acc_neg) ? (int64_t)ip : -(int64_t)ip;
    return true;
}

// Is the prefix `p` (sum of the aligned terms so far) a legitimate state of a sum that started at m?  The exact
// sum is within half a unit of (m + p) * q: strictly above 2^63 units it is inside the binade, at 2^63 itself it
// may lie below it (finer grid); a carry out of 64 bits has left the binade upwards.
HB_HD bool x80_prefix_ok(uint64_t m, int64_t p)
{
    const uint64_t nm = m + (uint64_t)p;
    return !(p >= 0 && nm < m) && nm > (1ull << 63);
}

// The prefix sums are taken in wrapping 64-bit arithmetic; a coarse sum of the terms >> 8 beside them (which
// cannot wrap) says whether the true prefix is below 2^63 in magnitude, i.e. whether the wrapped value is the
// true one.
HB_HD bool x80_coarse_ok(int64_t c) { return c > -(1ll << 54) && c < (1ll << 54); }

// acc + terms[0..n), n <= 32, in order (loop form of the warp procedure, same decisions).  *nfast: terms added
// as integers.  An exact-zero term never changes the sum (x80_add returns the other operand) and is skipped.
HB_HD x80 x80_block_add(x80 acc, const x80 *t, int n, int *nfast)
{
    int start = 0, nf = 0;
    while (start < n) {
        uint64_t p = 0; // wrapping
        int64_t c = 0;
        int g = start;
        uint64_t m_ok = acc.m;
        for (; g < n; g++) { // the prefix sum, up to the first term that breaks a condition
            if (t[g].m == 0) continue;
            int64_t sj;
            if (acc.m == 0 || !x80_align_term(t[g], acc.e, acc.neg, sj)) break;
            const uint64_t pn = p + (uint64_t)sj;
            const int64_t cn = c + (sj >> 8);
            if (!x80_coarse_ok(cn) || !x80_prefix_ok(acc.m, (int64_t)pn)) break;
            p = pn;
            c = cn;
            m_ok = acc.m + p;
        }
        nf += g - start;
        acc.m = m_ok;
        if (g < n) acc = x80_add(acc, t[g]);
        start = g + 1;
    }
    if (nfast) *nfast = nf;
    return acc;
}

#if defined(__CUDACC__)
// acc + the terms of lanes [0, lim) in lane order; every lane holds the same acc and gets the same result.
// The warp form of x80_block_add (hb_x87.cuh): each round aligns the unprocessed terms to the sum's grid, takes
// their integer prefix sums with shuffles, applies the longest legitimate prefix in one addition, and adds the
// term that stopped it (a tie, a term too large for the grid, the sum leaving its binade) literally.
__device__ __forceinline__ x80 x80_block_add_warp(x80 acc, x80 term, int lane, int lim)
{
    const unsigned full = 0xffffffffu;
    int start = 0;
    while (start < lim) {
        const bool mine = lane >= start && lane < lim && term.m != 0; // exact zeros never change the sum
        int64_t sj = 0;
        const bool okj = !mine || (acc.m != 0 && x80_align_term(term, acc.e, acc.neg, sj));
        const unsigned bad = __ballot_sync(full, !okj);
        const int f = bad ? __ffs((int)bad) - 1 : lim; // first term that cannot go on the grid
        if (!mine || lane >= f) sj = 0;
        uint64_t p = (uint64_t)sj; // wrapping prefix sum, and a coarse one that says whether it wrapped
        int64_t c = sj >> 8;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t up = __shfl_up_sync(full, p, o);
            const int64_t uc = __shfl_up_sync(full, c, o);
            if (lane >= o) {
                p += up;
                c += uc;
            }
        }
        const bool v = !mine || lane >= f || (x80_coarse_ok(c) && x80_prefix_ok(acc.m, (int64_t)p));
        const unsigned inv = __ballot_sync(full, !v);
        const int g = inv ? __ffs((int)inv) - 1 : f; // first term whose prefix is not a legitimate state
        const uint64_t pg = __shfl_sync(full, p, g > 0 ? g - 1 : 0);
        if (g > start) acc.m += pg;
        if (g < lim) {
            x80 tg;
            tg.m = __shfl_sync(full, term.m, g);
            tg.e = __shfl_sync(full, term.e, g);
            tg.neg = __shfl_sync(full, term.neg, g);
            acc = x80_add(acc, tg);
        }
        start = g + 1;
    }
    return acc;
}
#endif

HB_HD double x80_to_double(x80 a)
{
    if (a.m == 0) return a.neg ? -0.0 : 0.0;
    uint64_t mant = a.m >> 11, rem = a.m & 0x7ffull;
    int32_t e = a.e;
    if (rem > 0x400ull || (rem == 0x400ull && (mant & 1ull))) {
        mant += 1;
        if (mant >> 53) {
            mant >>= 1;
            e += 1;
        }
    }
    int32_t ef = e + 1023;
    if (ef >= 1 && ef <= 2046) {
        uint64_t u = ((uint64_t)a.neg << 63) | ((uint64_t)ef << 52) | (mant & 0xfffffffffffffull);
        return mk64((int32_t)(u >> 32), (int32_t)(u & 0xffffffffu));
    }
    double v = ldexp((double)mant, e - 52); // out of the normal range (not reached on this path)
    return a.neg ? -v : v;
}

} // namespace hb
