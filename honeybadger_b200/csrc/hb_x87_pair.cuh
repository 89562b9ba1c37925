// hb_x87_pair.cuh — the 80-bit running sums of hb_x87.cuh carried as a PAIR of doubles.
//
// An x87 extended value has a 64-bit significand, so it is h + l exactly with h = RN53(value) and
// l the remainder (at most 11 significant bits).  One step of R's `mean` accumulation, the
// running sum plus a term rounded to 64 bits, is then: build the exact sum as a few doubles with
// error-free transformations (Knuth TwoSum, Shewchuk's GROW-EXPANSION) and round that to 64 bits
// by hand.  A few dozen fp64 additions and selects with no data-dependent branch, against the
// ~150 integer instructions (and per-lane branches) of x80_add — the accumulation in
// pool_accum_cm_kernel is instruction-bound, so this is what it costs.
//
// Every step either takes the fast path below or falls back to the integer routine for that one
// step; both produce the exactly rounded 64-bit result, so states are bit-identical to
// hb_x87.cuh's on any input (tests/test_numerics.py drives both against the host's x87 unit).
// Compile with FMA contraction off (-fmad=false / -ffp-contract=off): the transformations
// need each operation rounded on its own.
#pragma once
#include "hb_x87.cuh"

#ifndef HB_X80P_COVER
#define HB_X80P_COVER(i) // tests/emul counts the cases taken (host build only)
#endif

namespace hb {

struct x80p {
    double h, l;
};

HB_HD void two_sum(double a, double b, double &s, double &e)
{
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}

HB_HD x80p x80p_from_x80(x80 a)
{
    x80p r;
    if (a.m == 0) { // zero keeps its sign in h; l is +0 (the subtraction below would hand back a zero of the other sign)
        r.h = a.neg ? -0.0 : 0.0;
        r.l = 0.0;
        return r;
    }
    r.h = x80_to_double(a);
    r.l = x80_to_double(x80_sub(a, x80_from_double(r.h))); // exact: at most 11 bits are left
    return r;
}

HB_HD x80 x80_from_x80p(x80p a)
{
    if (a.l == 0.0) return x80_from_double(a.h);
    return x80_add(x80_from_double(a.h), x80_from_double(a.l)); // exact: the sum has 64 bits
}

// RN64(a + z0 + z1 + z2), the four doubles being ANY exact decomposition of the value whose terms
// happen to fall off quickly (what the error-free sums below produce).  Written without
// data-dependent branches — the 32 lanes of a warp carry 32 different genes, and a branch per
// rounding case serialises them — so every case is computed and selected:
//   q = one unit of the 64-bit result, r = y0 to the nearest multiple of q (the 1.5 * 2^52 * q
//   constant trick; ties to even, which is the right parity because a / q is even), rem = y0 - r;
//   0 < |rem| < q/2 : the terms below y0 change nothing if they are smaller than q/2 - |rem|;
//   |rem| = q/2     : a tie on y0 alone, the sign of the next term pushes it off the tie;
//   rem = 0         : y0 is on the grid and the next term is rounded on its own, same rules.
// Nothing is assumed about the inputs beyond a + z0 + z1 + z2 being the exact value: every
// inequality the case analysis needs (y0 at most half an ulp of a, each further term at most half
// the one before, the tail inside the gap to the nearest decision point) is CHECKED, and the
// function returns false when one fails — a == 0, exponents near the limits, a second-level tie —
// so that the caller redoes that step in integers.  A true return is therefore the exactly
// rounded sum by construction, not by the statistics of a fuzz test.
HB_HD bool x80p_round(double a, double z0, double z1, double z2, x80p &out)
{
    const double y0 = z0, y1 = z1, y2 = z2;
    const double f0 = fabs(y0), f1 = fabs(y1), f2 = fabs(y2);
    const int32_t ah = hi32(a);
    const int32_t E2 = ((ah >> 20) & 0x7ff) - 1023;
    bool ok = (uint32_t)(E2 + 900) <= 1800u; // also rejects a == 0
    ok = ok && f0 <= mk64((int32_t)((uint32_t)(E2 - 53 + 1023) << 20), 0) && f1 + f1 <= f0 && f2 + f2 <= f1;
    // exponent of the exact sum: one less than a's when a is a power of two and the tail pulls it down
    const bool down = (((ah & 0xfffff) | lo32(a)) == 0) && y0 != 0.0 && ((hi32(y0) ^ ah) < 0);
    const int32_t eb = (int32_t)((uint32_t)(E2 - (down ? 1 : 0) + 1023) << 20);
    const double q = mk64(eb - (63 << 20), 0), half = mk64(eb - (64 << 20), 0);
    const double C = mk64((eb - (11 << 20)) | 0x80000, 0); // 1.5 * 2^52 * q
    const double r = (y0 + C) - C, rem = y0 - r;
    const double r1 = (y1 + C) - C, rem1 = y1 - r1;
    const bool tie0 = fabs(rem) == half, grid0 = rem == 0.0;
    const bool push = tie0 && y1 != 0.0 && ((hi32(y1) ^ hi32(rem)) >= 0);
    double adj = push ? copysign(q, rem) : 0.0;
    adj = grid0 ? r1 : adj;
    // the tail must stay inside the gap to the nearest point where the decision would change
    const double tail = (grid0 ? f2 : f1 + f2) * 1.0000000000000009; // (1 + 2^-50): covers the rounding of f1 + f2
    const double gap = half - (grid0 ? fabs(rem1) : tie0 ? 0.0 : fabs(rem));
    ok = ok && (tail < gap || (grid0 && rem1 == 0.0 && y2 == 0.0));
    HB_X80P_COVER(tie0 ? (push ? 0 : 1) : grid0 ? (y1 != 0.0 ? 2 : 3) : 4);
    if (down) HB_X80P_COVER(5);
    const double rr = r + adj;
    out.h = a + rr;
    out.l = rr - (out.h - a);
    return ok;
}

// the steps the fast path declines, redone in integers (out of line: they are rare)
HB_HD_NOINLINE x80p x80p_add_double_slow(x80p s, double v)
{
    HB_X80P_COVER(6);
    return x80p_from_x80(x80_add(x80_from_x80p(s), x80_from_double(v)));
}
HB_HD_NOINLINE x80p x80p_add_slow(x80p s, x80p d)
{
    HB_X80P_COVER(7);
    return x80p_from_x80(x80_add(x80_from_x80p(s), x80_from_x80p(d)));
}

// s + v rounded to 64 bits.  GROW-EXPANSION gives the exact sum as three doubles; two more TwoSums
// bring it to the form x80p_round wants (leading term = the rounded whole) whatever cancels.
HB_HD x80p x80p_add_double(x80p s, double v)
{
    double Q, c, a, b, s1, t1, a2, s2;
    two_sum(s.l, v, Q, c);
    two_sum(s.h, Q, a, b);
    two_sum(b, c, s1, t1);
    two_sum(a, s1, a2, s2);
    x80p r;
    if (x80p_round(a2, s2, t1, 0.0, r)) return r;
    // Declined.  A zero operand first: the other one comes back as it is, exactly what x80_add does (0 + 0 has no
    // exponent to round at, and rows that start with zeros would otherwise call the integer routine at every step).
    if (s.h == 0.0) {
        r.h = v;
        r.l = 0.0;
        return r;
    }
    if (v == 0.0) return s;
    return x80p_add_double_slow(s, v);
}

// s + d rounded to 64 bits, both 64-bit values
HB_HD x80p x80p_add(x80p s, x80p d)
{
    double Q1, c1, a1, b1, Q2, e0, Q3, e1, e2, e3, s1, t1, s2, t2, a2, s3;
    two_sum(s.l, d.l, Q1, c1);
    two_sum(s.h, Q1, a1, b1);
    two_sum(c1, d.h, Q2, e0);
    two_sum(b1, Q2, Q3, e1);
    two_sum(a1, Q3, e3, e2);
    two_sum(e1, e0, s1, t1);
    two_sum(e2, s1, s2, t2);
    two_sum(e3, s2, a2, s3);
    x80p r;
    if (x80p_round(a2, s3, t2, t1, r)) return r;
    if (s.h == 0.0) return d; // as above
    if (d.h == 0.0) return s;
    return x80p_add_slow(s, d);
}

} // namespace hb
