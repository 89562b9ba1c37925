// hb_longchain.cuh — Viterbi for a FEW, LONG chains, parallel along the gene axis and bit-exact.
//
// The pooled rounds of the reference (R/HoneyBADGER.R:1141-1151, :1404-1410) run K <= a few dozen chains
// of G = 6 000 ... 200 000 positions.  One thread per chain leaves the GPU idle there (40 ms for
// 6 x 200 000 steps).  HiddenMarkov::Viterbi's recursion
//     nu_t(k) = max_j( nu_{t-1}(j) + logPi[j][k] ) + logb_k(t)            (every "+" rounded)
// is not associative in floating point, so a plain max-plus scan cannot reproduce it.  This file cuts
// the chain into blocks of L positions and uses two facts:
//
//  1. While every value of a block lies in ONE binade [2^E, 2^(E+1)), all of them are integer
//     multiples of q = 2^(E-52), and RN(v + c) = v + rint(c / q) q for any addend c — the rounded
//     recursion is an INTEGER max-plus recursion (translation invariant, monotone, associative).  A
//     half-way addend (c/q = m + 1/2) rounds to the even neighbour, which depends only on the parity
//     of v: still monotone, so "max" still distributes; one transfer matrix per entry parity covers it.
//     Each block therefore has an exact integer transfer matrix, computable without knowing the
//     block's entry value (only its binade, taken from an approximate fp64 pass).
//  2. Whatever predicted a block's entry, running the block with the LITERAL fp64 recursion from that
//     entry and finding its exit bit-identical to the next block's entry proves, by induction from
//     position 0, that every entry — hence every back-pointer written by the literal runs — is the
//     sequential recursion's.  The integer model is never trusted, only checked.
//
// Phases (each a small kernel; "parallel" = one thread per (chain, block), "sequential" = one thread
// per chain walking over the blocks):
//   P0 parallel    fp64 max-plus transfer matrix of every block                     lc_fmat
//   P1 sequential  approximate entries -> binade guess per block                    lc_prefix_approx
//   P2 parallel    integer transfer matrices (entry parity 0 / 1) for that binade   lc_imat
//   P3 sequential  exact entries: integer prediction where the entry really lies in
//                  the guessed binade, literal run of the block otherwise            lc_prefix_exact
//   P4 parallel    literal run of every block from its entry: back-pointers, the
//                  block's exit->entry state map, exit compared with the next entry  lc_literal
//   (P3, P4 repeat from the first block whose exit did not match — P4 records it with an atomicMin;
//    after the last round the rest of such a chain is finished sequentially — always literal, always
//    exact).  The sequential walkers run as one warp per chain: the warp stages the per-block tables of
//    HB_LC_WIN blocks in shared memory with coalesced loads, lane 0 walks them.
//   P5 sequential  final state, states at the block boundaries through the maps     lc_resolve
//   P6 parallel    states inside every block                                         lc_write_y
//
// Everything is HB_HD so tests/emul can run the same bodies on the host.
#pragma once
#include "hb_device.cuh"

namespace hb {

#define HB_LC_NONUNIFORM (-100000)
#define HB_LC_NEG (-(int64_t)1 << 61) // "unreachable" in the integer matrices
#define HB_LC_ROUNDS 3
#define HB_LC_WIN 64 // blocks per staging window of the sequential walkers
#define HB_RESTRICT __restrict__

struct LcArgs {
    const double *lb; // [K][T][S] log-emissions, chain-major (literal values of the reference)
    int64_t T, nb;    // positions per chain; blocks per chain = ceil((T-1)/L)
    int32_t K, S, L;
    double lp[9]; // log Pi, row-major j -> k
    double ld[3]; // log delta
    double *fmat;     // [K][nb][S*S]
    int32_t *bexp;    // [K][nb]   binade of the block, or HB_LC_NONUNIFORM
    int64_t *imat;    // [K][nb][2][S*S]
    uint8_t *iok;     // [K][nb]   integer matrices valid
    double *entry;    // [K][nb][S]
    double *lexit;    // [K][nb][S] literal exits
    uint8_t *psi;     // [K][T]    back-pointers of position t, 2 bits per state
    uint8_t *bmap;    // [K][nb]   state at the block's entry boundary per exit state, 2 bits each
    int32_t *cstate;  // [K][8]: 0 first block whose entry is not yet proven | 1 done (1 proven, 2 finished
                      //         sequentially) | 2 repair rounds used | 3 blocks run literally by the
                      //         sequential walker | 4 first block of the last round whose exit did not match
    uint8_t *ybound;  // [K][nb]   state at the last position of the block (0-based)
    int32_t *y;       // [K][T]    1-based states (R integer)
    int32_t *status;  // [0] = 1: a chain ended with nu == -Inf
    // Staging window of the sequential walkers (device: shared memory filled by the whole warp; host:
    // NULL = read the global tables).  Rows of blocks [win_b0, win_b0 + HB_LC_WIN) of one chain.
    int64_t win_b0;
    const double *wf;
    const int64_t *wi;
    const int32_t *wbe;
    const uint8_t *wio;
    // forward-backward posteriors of the same chains (lc_fb_*): linear-domain block products
    double pi[9], dl[3]; // Pi, delta
    double *ffm;         // [K][nb][S*S] row j0: e_j0 (prod_t Pi diag(b_t)), each row scaled by 2^-ffe
    int32_t *ffe;        // [K][nb][S]
    double *fae;         // [K][nb][S]  filter at the position before the block (sums to 1)
    double *fbe;         // [K][nb][S]  beta at the block's last position (sums to 1)
    double *fls;         // [K][nb]     sum over the block of log c_t + max_k logb_t
    double *gamma;       // out: gamma_t(k) of chain c at gamma[c * g_sc + t * g_st + k * g_sk]
    int64_t g_sc, g_st, g_sk; // chain-major [K][T][S]: T*S, S, 1; R's [T x K x S] column-major: T, 1, K*T
    double *ll;          // [K]         out (or NULL)
    const double *wff;
    const int32_t *wfe;
};
#define HB_LC_CS 8
#define HB_LC_NOFAIL 0x7fffffff

HB_HD int64_t lc_t0(const LcArgs &a, int64_t b) { return 1 + b * a.L; }
HB_HD int64_t lc_t1(const LcArgs &a, int64_t b)
{
    const int64_t e = 1 + (b + 1) * a.L;
    return e < a.T ? e : a.T;
}
template <int S> HB_HD const double *lc_fm(const LcArgs &a, int32_t c, int64_t b)
{
    return a.wf ? a.wf + (b - a.win_b0) * S * S : a.fmat + ((int64_t)c * a.nb + b) * S * S;
}
template <int S> HB_HD const int64_t *lc_im(const LcArgs &a, int32_t c, int64_t b)
{
    return a.wi ? a.wi + (b - a.win_b0) * 2 * S * S : a.imat + ((int64_t)c * a.nb + b) * 2 * S * S;
}
HB_HD int32_t lc_be(const LcArgs &a, int32_t c, int64_t b)
{
    return a.wbe ? a.wbe[b - a.win_b0] : a.bexp[(int64_t)c * a.nb + b];
}
HB_HD int32_t lc_io(const LcArgs &a, int32_t c, int64_t b)
{
    return a.wio ? a.wio[b - a.win_b0] : a.iok[(int64_t)c * a.nb + b];
}
HB_HD void lc_mark_fail(int32_t *slot, int32_t b)
{
#if defined(__CUDA_ARCH__)
    atomicMin(slot, b);
#else
    if (b < *slot) *slot = b;
#endif
}

// One literal step (HiddenMarkov::Viterbi: apply(matrix(nu) + logPi, 2, max) + logb; which.max takes
// the first maximum).  Returns the back-pointers, 2 bits per state.
template <int S>
HB_HD uint32_t lc_lit_step(const double *lp, double *nu, const double *l)
{
    double nn[S];
    uint32_t pk = 0;
#pragma unroll
    for (int k = 0; k < S; k++) {
        double m = nu[0] + lp[k];
        uint32_t p = 0;
#pragma unroll
        for (int j = 1; j < S; j++) {
            const double c = nu[j] + lp[j * S + k];
            if (c > m) {
                m = c;
                p = j;
            }
        }
        nn[k] = m + l[k];
        pk |= p << (2 * k);
    }
#pragma unroll
    for (int k = 0; k < S; k++) nu[k] = nn[k];
    return pk;
}

// The block walkers are a handful of threads each following a dependent recursion: nothing else hides
// the latency of their loads, so the emissions are fetched one chunk of positions ahead into registers
// (independent loads in flight while the previous chunk is being worked on).
template <int S> struct LcChunk {
    static constexpr int C = S == 2 ? 16 : 8;
    double v[C * S];
};
template <int S>
HB_HD void lc_fetch(LcChunk<S> &ch, const double *HB_RESTRICT l, int64_t n)
{
#pragma unroll
    for (int i = 0; i < LcChunk<S>::C * S; i++) ch.v[i] = i < n * S ? l[i] : 0.0;
}

// Literal run of block b from `nu` (in: entry, out: exit).  WR: write psi, the block map and the exit.
template <int S, bool WR>
HB_HD void lc_literal_block(const LcArgs &a, int32_t c, int64_t b, double *nu)
{
    constexpr int C = LcChunk<S>::C;
    const int64_t t0 = lc_t0(a, b), n = lc_t1(a, b) - t0;
    const double *HB_RESTRICT l = a.lb + ((int64_t)c * a.T + t0) * S;
    uint8_t *HB_RESTRICT ps = a.psi + (int64_t)c * a.T + t0;
    uint32_t o[S];
#pragma unroll
    for (int k = 0; k < S; k++) o[k] = k;
    LcChunk<S> cur, nxt;
    lc_fetch<S>(cur, l, n);
    for (int64_t i0 = 0; i0 < n; i0 += C) {
        lc_fetch<S>(nxt, l + (i0 + C) * S, n - i0 - C);
#pragma unroll
        for (int j = 0; j < C; j++) {
            if (i0 + j < n) {
                const uint32_t pk = lc_lit_step<S>(a.lp, nu, cur.v + j * S);
                if (WR) {
                    ps[i0 + j] = (uint8_t)pk;
                    uint32_t on[S];
#pragma unroll
                    for (int k = 0; k < S; k++) {
                        const uint32_t p = (pk >> (2 * k)) & 3u;
                        uint32_t v = o[0];
#pragma unroll
                        for (int jj = 1; jj < S; jj++) v = (p == (uint32_t)jj) ? o[jj] : v;
                        on[k] = v;
                    }
#pragma unroll
                    for (int k = 0; k < S; k++) o[k] = on[k];
                }
            }
        }
        cur = nxt;
    }
    if (WR) {
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < S; k++) m |= o[k] << (2 * k);
        a.bmap[(int64_t)c * a.nb + b] = (uint8_t)m;
        double *le = a.lexit + ((int64_t)c * a.nb + b) * S;
#pragma unroll
        for (int k = 0; k < S; k++) le[k] = nu[k];
    }
}

// ---- P0: fp64 transfer matrix of a block (approximate: used only to guess binades)
template <int S>
HB_HD void lc_fmat(const LcArgs &a, int32_t c, int64_t b)
{
    constexpr int C = LcChunk<S>::C;
    const int64_t t0 = lc_t0(a, b), n = lc_t1(a, b) - t0;
    const double *HB_RESTRICT l = a.lb + ((int64_t)c * a.T + t0) * S;
    double w[S][S];
#pragma unroll
    for (int j = 0; j < S; j++)
#pragma unroll
        for (int k = 0; k < S; k++) w[j][k] = j == k ? 0.0 : -INFINITY;
    LcChunk<S> cur, nxt;
    lc_fetch<S>(cur, l, n);
    for (int64_t i0 = 0; i0 < n; i0 += C) {
        lc_fetch<S>(nxt, l + (i0 + C) * S, n - i0 - C);
#pragma unroll
        for (int j = 0; j < C; j++)
            if (i0 + j < n) {
#pragma unroll
                for (int j0 = 0; j0 < S; j0++) lc_lit_step<S>(a.lp, w[j0], cur.v + j * S);
            }
        cur = nxt;
    }
    double *f = a.fmat + ((int64_t)c * a.nb + b) * S * S;
#pragma unroll
    for (int j = 0; j < S; j++)
#pragma unroll
        for (int k = 0; k < S; k++) f[j * S + k] = w[j][k];
}

// Common binade of n values: all finite, normal, same sign, same exponent -> that exponent.
HB_HD int32_t lc_binade(const double *v, int n)
{
    const int32_t h0 = hi32(v[0]);
    const int32_t e0 = (h0 >> 20) & 0x7ff;
    if (e0 == 0 || e0 == 0x7ff) return HB_LC_NONUNIFORM;
    for (int i = 1; i < n; i++)
        if (((hi32(v[i]) ^ h0) >> 20) != 0) return HB_LC_NONUNIFORM; // sign or exponent differs
    const int32_t E = e0 - 1023;
    if (E < -400 || E > 900) return HB_LC_NONUNIFORM;
    return E;
}

// ---- P1: approximate entries, binade guess per block.  v[S]: running entry (in/out); `be` receives the
// guesses of blocks [bb, be_) at be[b - bb].
template <int S>
HB_HD void lc_prefix_approx_begin(const LcArgs &a, int32_t c, double *v)
{
    const double *l0 = a.lb + (int64_t)c * a.T * S;
#pragma unroll
    for (int k = 0; k < S; k++) v[k] = a.ld[k] + l0[k];
}
template <int S>
HB_HD void lc_prefix_approx_range(const LcArgs &a, int32_t c, int64_t bb, int64_t be_, double *v, int32_t *be)
{
    for (int64_t b = bb; b < be_; b++) {
        const double *f = lc_fm<S>(a, c, b);
        double x[2 * S];
#pragma unroll
        for (int k = 0; k < S; k++) x[k] = v[k];
#pragma unroll
        for (int k = 0; k < S; k++) {
            double m = v[0] + f[k];
#pragma unroll
            for (int j = 1; j < S; j++) {
                const double y = v[j] + f[j * S + k];
                m = y > m ? y : m;
            }
            x[S + k] = m;
        }
        be[b - bb] = lc_binade(x, 2 * S);
#pragma unroll
        for (int k = 0; k < S; k++) v[k] = x[S + k];
    }
}

// ---- grid arithmetic
struct LcInc {
    int64_t v;
    int32_t tie;
};
// c / q as an integer increment; false when it cannot be represented (non-finite, too large).
HB_HD bool lc_to_grid(double c, double scale, LcInc &o)
{
    o.tie = 0;
    if (c == -INFINITY) {
        o.v = HB_LC_NEG;
        return true;
    }
    const double d = c * scale; // exact: scale is a power of two
    if (!(fabs(d) < 9007199254740992.0)) return false;
    const double r = rint(d);
    o.tie = fabs(d - r) == 0.5 ? 1 : 0;
    o.v = (int64_t)(o.tie ? floor(d) : r);
    return true;
}
// v + inc on the grid; `par0` is the parity of the block's entry value (v is relative to it).
HB_HD int64_t lc_add(int64_t v, int64_t par0, const LcInc &i)
{
    int64_t s = v + i.v;
    if (i.tie) s += (s + par0) & 1; // of s, s + 1 the one whose absolute value is even
    return s < HB_LC_NEG / 2 ? HB_LC_NEG : s;
}

// ---- P2: integer transfer matrices of a block for its guessed binade
template <int S>
HB_HD void lc_imat(const LcArgs &a, int32_t c, int64_t b)
{
    constexpr int C = LcChunk<S>::C;
    const int64_t ib = (int64_t)c * a.nb + b;
    const int32_t E = a.bexp[ib];
    a.iok[ib] = 0;
    if (E == HB_LC_NONUNIFORM) return;
    const double scale = pow2i(52 - E);
    LcInc ap[S * S];
    bool good = true;
#pragma unroll
    for (int i = 0; i < S * S; i++) good = lc_to_grid(a.lp[i], scale, ap[i]) && good;
    if (!good) return;
    const int64_t t0 = lc_t0(a, b), n = lc_t1(a, b) - t0;
    const double *HB_RESTRICT l = a.lb + ((int64_t)c * a.T + t0) * S;
    int64_t w[2][S][S]; // [entry parity][entry state][state]
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int j = 0; j < S; j++)
#pragma unroll
            for (int k = 0; k < S; k++) w[p][j][k] = j == k ? 0 : HB_LC_NEG;
    LcChunk<S> cur, nxt;
    lc_fetch<S>(cur, l, n);
    for (int64_t i0 = 0; i0 < n; i0 += C) {
        lc_fetch<S>(nxt, l + (i0 + C) * S, n - i0 - C);
#pragma unroll
        for (int jc = 0; jc < C; jc++) {
            if (i0 + jc >= n) continue;
            LcInc e[S];
#pragma unroll
            for (int k = 0; k < S; k++) good = lc_to_grid(cur.v[jc * S + k], scale, e[k]) && good;
#pragma unroll
            for (int p = 0; p < 2; p++)
#pragma unroll
                for (int j0 = 0; j0 < S; j0++) {
                    int64_t nn[S];
#pragma unroll
                    for (int k = 0; k < S; k++) {
                        int64_t m = lc_add(w[p][j0][0], p, ap[k]);
#pragma unroll
                        for (int j = 1; j < S; j++) {
                            const int64_t x = lc_add(w[p][j0][j], p, ap[j * S + k]);
                            m = x > m ? x : m;
                        }
                        nn[k] = lc_add(m, p, e[k]);
                    }
#pragma unroll
                    for (int k = 0; k < S; k++) w[p][j0][k] = nn[k];
                }
        }
        cur = nxt;
    }
    if (!good) return;
    int64_t *im = a.imat + ib * 2 * S * S;
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int j = 0; j < S; j++)
#pragma unroll
            for (int k = 0; k < S; k++) im[(p * S + j) * S + k] = w[p][j][k];
    a.iok[ib] = 1;
}

// Exit of block b from the exact entry `nu` through the integer matrices; false when the entry or
// the predicted exit is not in the block's binade (the caller then runs the block literally).
template <int S>
HB_HD bool lc_predict(const LcArgs &a, int32_t c, int64_t b, double *nu)
{
    if (!lc_io(a, c, b)) return false;
    const int32_t E = lc_be(a, c, b);
    const int32_t h0 = hi32(nu[0]);
    if (((h0 >> 20) & 0x7ff) != E + 1023) return false;
#pragma unroll
    for (int k = 1; k < S; k++)
        if (((hi32(nu[k]) ^ h0) >> 20) != 0) return false;
    const double scale = pow2i(52 - E), q = pow2i(E - 52);
    const int64_t *im = lc_im<S>(a, c, b);
    int64_t n[S], x[S];
#pragma unroll
    for (int k = 0; k < S; k++) n[k] = (int64_t)(nu[k] * scale); // exact: nu is a multiple of q
    const int64_t lo = (int64_t)1 << 52, hi = (int64_t)1 << 53;
    const bool neg = h0 < 0;
#pragma unroll
    for (int k = 0; k < S; k++) {
        int64_t m = HB_LC_NEG * 2;
#pragma unroll
        for (int j = 0; j < S; j++) {
            const int64_t v = n[j] + im[((n[j] & 1) * S + j) * S + k];
            m = v > m ? v : m;
        }
        const int64_t am = neg ? -m : m;
        if (am < lo || am >= hi) return false;
        x[k] = m;
    }
#pragma unroll
    for (int k = 0; k < S; k++) nu[k] = (double)x[k] * q;
    return true;
}

// ---- P3: exact entries of the blocks from the first one whose entry is not yet proven.
// begin: where to start (false: nothing to do); range: blocks [bb, be_) with `nu` carried along.
template <int S>
HB_HD bool lc_prefix_exact_begin(const LcArgs &a, int32_t c, int32_t round, double *nu, int64_t &b0,
                                 int64_t &forced_end)
{
    int32_t *cs = a.cstate + c * HB_LC_CS;
    if (cs[1]) return false;
    b0 = 0;
    forced_end = 0;
    if (round == 0) {
        const double *l0 = a.lb + (int64_t)c * a.T * S;
#pragma unroll
        for (int k = 0; k < S; k++) nu[k] = a.ld[k] + l0[k];
    } else {
        const int64_t f = cs[4];
        if (f >= a.nb) { // every exit matched: the chain is proven
            cs[0] = (int32_t)a.nb;
            cs[1] = 1;
            return false;
        }
        const double *le = a.lexit + ((int64_t)c * a.nb + f) * S;
#pragma unroll
        for (int k = 0; k < S; k++) nu[k] = le[k]; // block f had a true entry: its literal exit is true
        b0 = f + 1;
        cs[0] = (int32_t)b0;
        cs[2] = round;
        forced_end = b0 + ((int64_t)1 << (2 * round)); // literal right after a miss
    }
    cs[4] = HB_LC_NOFAIL;
    return true;
}
template <int S>
HB_HD void lc_prefix_exact_range(const LcArgs &a, int32_t c, int64_t bb, int64_t be_, double *nu,
                                 int64_t forced_end, int32_t &nlit)
{
    for (int64_t b = bb; b < be_; b++) {
        double *en = a.entry + ((int64_t)c * a.nb + b) * S;
#pragma unroll
        for (int k = 0; k < S; k++) en[k] = nu[k];
        if (b < forced_end || !lc_predict<S>(a, c, b, nu)) {
            lc_literal_block<S, false>(a, c, b, nu);
            nlit++;
        }
    }
}

// ---- P4: literal run of a block from its entry, exit checked against the next entry
template <int S>
HB_HD void lc_literal(const LcArgs &a, int32_t c, int64_t b)
{
    int32_t *cs = a.cstate + c * HB_LC_CS;
    if (cs[1] || b < cs[0]) return;
    const int64_t ib = (int64_t)c * a.nb + b;
    double nu[S];
#pragma unroll
    for (int k = 0; k < S; k++) nu[k] = a.entry[ib * S + k];
    lc_literal_block<S, true>(a, c, b, nu);
    bool same = true;
    if (b + 1 < a.nb) {
        const double *nx = a.entry + (ib + 1) * S;
#pragma unroll
        for (int k = 0; k < S; k++)
            same = same && hi32(nx[k]) == hi32(nu[k]) && lo32(nx[k]) == lo32(nu[k]);
    }
    if (!same) lc_mark_fail(cs + 4, (int32_t)b);
}

// ---- P5: finish unproven chains sequentially; final state and boundary states
template <int S>
HB_HD void lc_resolve(const LcArgs &a, int32_t c)
{
    int32_t *cs = a.cstate + c * HB_LC_CS;
    double nu[S];
    if (a.nb == 0) {
        const double *l0 = a.lb + (int64_t)c * a.T * S;
#pragma unroll
        for (int k = 0; k < S; k++) nu[k] = a.ld[k] + l0[k];
    } else {
        if (!cs[1]) {
            const int64_t f = cs[4];
            if (f < a.nb) { // literal from the last proven exit to the end of the chain
                const double *le = a.lexit + ((int64_t)c * a.nb + f) * S;
#pragma unroll
                for (int k = 0; k < S; k++) nu[k] = le[k];
                for (int64_t b = f + 1; b < a.nb; b++) lc_literal_block<S, true>(a, c, b, nu);
                cs[3] += (int32_t)(a.nb - 1 - f);
                cs[1] = 2;
            } else {
                cs[1] = 1;
            }
            cs[0] = (int32_t)a.nb;
        }
        const double *le = a.lexit + ((int64_t)c * a.nb + a.nb - 1) * S;
#pragma unroll
        for (int k = 0; k < S; k++) nu[k] = le[k];
    }
    bool under = false;
    uint32_t st = 0;
#pragma unroll
    for (int k = 0; k < S; k++) under = under || nu[k] == -INFINITY;
#pragma unroll
    for (int k = 1; k < S; k++)
        if (nu[k] > nu[st]) st = k;
    if (under) *a.status = 1;
    if (a.nb == 0) {
        a.y[(int64_t)c * a.T] = (int32_t)st + 1;
        return;
    }
    const uint8_t *HB_RESTRICT bm = a.bmap + (int64_t)c * a.nb;
    uint8_t *HB_RESTRICT yb = a.ybound + (int64_t)c * a.nb;
    for (int64_t b1 = a.nb; b1 > 0; b1 -= 16) { // 16 maps in flight per round trip
        uint8_t m[16];
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = b1 - 1 - i >= 0 ? bm[b1 - 1 - i] : 0;
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (b1 - 1 - i >= 0) {
                yb[b1 - 1 - i] = (uint8_t)st;
                st = (m[i] >> (2 * st)) & 3u;
            }
    }
}

// ---- P6: states inside a block, from the state at its last position
template <int S>
HB_HD void lc_write_y(const LcArgs &a, int32_t c, int64_t b)
{
    const int64_t t0 = lc_t0(a, b), t1 = lc_t1(a, b);
    const uint8_t *HB_RESTRICT ps = a.psi + (int64_t)c * a.T;
    int32_t *HB_RESTRICT y = a.y + (int64_t)c * a.T;
    uint32_t st = a.ybound[(int64_t)c * a.nb + b];
    for (int64_t t = t1 - 1; t >= t0; t -= 16) {
        uint8_t m[16];
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = t - i >= t0 ? ps[t - i] : 0;
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (t - i >= t0) {
                y[t - i] = (int32_t)st + 1;
                st = (m[i] >> (2 * st)) & 3u;
            }
    }
    if (b == 0) y[0] = (int32_t)st + 1;
}


// =====================================================================================================
// Forward-backward posteriors for the same few long chains (semantics of HiddenMarkov::forwardback,
// gamma_t = exp(logalpha_t + logbeta_t - LL); not bit-level: 1e-9, north_star).  Linear domain:
// alpha_t = (alpha_{t-1} Pi) o b_t, beta_{t-1} = Pi (b_t o beta_t), b_t = exp(logb_t - max_k logb_t).
// With A_t = Pi diag(b_t) one block product F_b = prod_t A_t serves both sweeps (alpha_exit = alpha_entry F_b,
// beta_entry = F_b beta_exit); it is associative up to rounding, all terms non-negative (no cancellation).
//   F0 parallel    F_b, one row per entry state, each row with its own binary exponent      lc_fb_mat
//   F1 sequential  filter before every block (forward), beta at every block end (backward)  lc_fb_fwd/bwd
//   F2 parallel    inside the block: filter forward (parked in gamma), beta backward, gamma  lc_fb_block
//   F3 sequential  LL = sum of the blocks' log normalisers                                   lc_fb_ll
template <int S>
HB_HD void lc_emis(const double *l, double *b, double &m)
{
    m = l[0];
#pragma unroll
    for (int k = 1; k < S; k++) m = l[k] > m ? l[k] : m;
#pragma unroll
    for (int k = 0; k < S; k++) b[k] = m == -INFINITY ? 0.0 : exp(l[k] - m);
}
// w <- (w Pi) o b
template <int S>
HB_HD void lc_fwd_vec(const double *pi, double *w, const double *b)
{
    double n[S];
#pragma unroll
    for (int k = 0; k < S; k++) {
        double acc = w[0] * pi[k];
#pragma unroll
        for (int j = 1; j < S; j++) acc += w[j] * pi[j * S + k];
        n[k] = acc * b[k];
    }
#pragma unroll
    for (int k = 0; k < S; k++) w[k] = n[k];
}
// w <- Pi (b o w)
template <int S>
HB_HD void lc_bwd_vec(const double *pi, double *w, const double *b)
{
    double u[S], n[S];
#pragma unroll
    for (int k = 0; k < S; k++) u[k] = b[k] * w[k];
#pragma unroll
    for (int j = 0; j < S; j++) {
        double acc = pi[j * S] * u[0];
#pragma unroll
        for (int k = 1; k < S; k++) acc += pi[j * S + k] * u[k];
        n[j] = acc;
    }
#pragma unroll
    for (int k = 0; k < S; k++) w[k] = n[k];
}
template <int S>
HB_HD double lc_normalise(double *w)
{
    double s = w[0];
#pragma unroll
    for (int k = 1; k < S; k++) s += w[k];
    const double r = 1.0 / s;
#pragma unroll
    for (int k = 0; k < S; k++) w[k] *= r;
    return s;
}
// power-of-two rescale of a non-negative vector to max ~ 1; returns the exponent taken out
template <int S>
HB_HD int32_t lc_rescale(double *w)
{
    double m = w[0];
#pragma unroll
    for (int k = 1; k < S; k++) m = w[k] > m ? w[k] : m;
    if (!(m > 0.0) || !(m < INFINITY)) return 0;
    const int32_t e = ((hi32(m) >> 20) & 0x7ff) - 1023;
    if (e == 0 || e < -1000) return 0; // already ~1, or denormal (leave)
    const double f = pow2i(-e);
#pragma unroll
    for (int k = 0; k < S; k++) w[k] *= f;
    return e;
}

// ---- F0
template <int S>
HB_HD void lc_fb_mat(const LcArgs &a, int32_t c, int64_t b)
{
    constexpr int C = LcChunk<S>::C;
    const int64_t t0 = lc_t0(a, b), n = lc_t1(a, b) - t0;
    const double *HB_RESTRICT l = a.lb + ((int64_t)c * a.T + t0) * S;
    double w[S][S];
    int32_t e[S];
#pragma unroll
    for (int j = 0; j < S; j++) {
        e[j] = 0;
#pragma unroll
        for (int k = 0; k < S; k++) w[j][k] = j == k ? 1.0 : 0.0;
    }
    LcChunk<S> cur, nxt;
    lc_fetch<S>(cur, l, n);
    for (int64_t i0 = 0; i0 < n; i0 += C) {
        lc_fetch<S>(nxt, l + (i0 + C) * S, n - i0 - C);
#pragma unroll
        for (int j = 0; j < C; j++)
            if (i0 + j < n) {
                double bv[S], m;
                lc_emis<S>(cur.v + j * S, bv, m);
#pragma unroll
                for (int j0 = 0; j0 < S; j0++) lc_fwd_vec<S>(a.pi, w[j0], bv);
            }
#pragma unroll
        for (int j0 = 0; j0 < S; j0++) e[j0] += lc_rescale<S>(w[j0]);
        cur = nxt;
    }
    const int64_t ib = (int64_t)c * a.nb + b;
#pragma unroll
    for (int j = 0; j < S; j++) {
        a.ffe[ib * S + j] = e[j];
#pragma unroll
        for (int k = 0; k < S; k++) a.ffm[ib * S * S + j * S + k] = w[j][k];
    }
}
template <int S> HB_HD const double *lc_ff(const LcArgs &a, int32_t c, int64_t b)
{
    return a.wff ? a.wff + (b - a.win_b0) * S * S : a.ffm + ((int64_t)c * a.nb + b) * S * S;
}
template <int S> HB_HD const int32_t *lc_fe(const LcArgs &a, int32_t c, int64_t b)
{
    return a.wfe ? a.wfe + (b - a.win_b0) * S : a.ffe + ((int64_t)c * a.nb + b) * S;
}
HB_HD double lc_pow2_clamped(int32_t k) { return k < -1020 ? 0.0 : pow2i(k > 1000 ? 1000 : k); }

// ---- F1 forward: filter at the position before each block of [bb, be_); `al` carried (sums to 1)
template <int S>
HB_HD void lc_fb_begin(const LcArgs &a, int32_t c, double *al)
{
    const double *l0 = a.lb + (int64_t)c * a.T * S;
    double bv[S], m;
    lc_emis<S>(l0, bv, m);
#pragma unroll
    for (int k = 0; k < S; k++) al[k] = a.dl[k] * bv[k];
    lc_normalise<S>(al);
}
template <int S>
HB_HD void lc_fb_fwd_range(const LcArgs &a, int32_t c, int64_t bb, int64_t be_, double *al)
{
    for (int64_t b = bb; b < be_; b++) {
        double *o = a.fae + ((int64_t)c * a.nb + b) * S;
#pragma unroll
        for (int k = 0; k < S; k++) o[k] = al[k];
        const double *f = lc_ff<S>(a, c, b);
        const int32_t *e = lc_fe<S>(a, c, b);
        int32_t em = -100000;
#pragma unroll
        for (int j = 0; j < S; j++) em = (al[j] > 0.0 && e[j] > em) ? e[j] : em;
        double n[S];
#pragma unroll
        for (int k = 0; k < S; k++) n[k] = 0.0;
#pragma unroll
        for (int j = 0; j < S; j++) {
            const double wgt = al[j] > 0.0 ? al[j] * lc_pow2_clamped(e[j] - em) : 0.0;
#pragma unroll
            for (int k = 0; k < S; k++) n[k] += wgt * f[j * S + k];
        }
        lc_normalise<S>(n);
#pragma unroll
        for (int k = 0; k < S; k++) al[k] = n[k];
    }
}
// ---- F1 backward: beta at the last position of each block, blocks be_-1 down to bb; `be` carried
template <int S>
HB_HD void lc_fb_bwd_range(const LcArgs &a, int32_t c, int64_t bb, int64_t be_, double *bt)
{
    for (int64_t b = be_ - 1; b >= bb; b--) {
        double *o = a.fbe + ((int64_t)c * a.nb + b) * S;
#pragma unroll
        for (int k = 0; k < S; k++) o[k] = bt[k];
        const double *f = lc_ff<S>(a, c, b);
        const int32_t *e = lc_fe<S>(a, c, b);
        double n[S];
        int32_t em = -100000;
#pragma unroll
        for (int j = 0; j < S; j++) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < S; k++) acc += f[j * S + k] * bt[k];
            n[j] = acc;
            em = (acc > 0.0 && e[j] > em) ? e[j] : em;
        }
#pragma unroll
        for (int j = 0; j < S; j++) n[j] = n[j] > 0.0 ? n[j] * lc_pow2_clamped(e[j] - em) : 0.0;
        lc_normalise<S>(n);
#pragma unroll
        for (int k = 0; k < S; k++) bt[k] = n[k];
    }
}

// ---- F2: posteriors inside a block
template <int S>
HB_HD void lc_fb_block(const LcArgs &a, int32_t c, int64_t b)
{
    constexpr int C = LcChunk<S>::C;
    const int64_t t0 = lc_t0(a, b), n = lc_t1(a, b) - t0;
    const int64_t ib = (int64_t)c * a.nb + b;
    const double *HB_RESTRICT l = a.lb + ((int64_t)c * a.T + t0) * S;
    double *HB_RESTRICT g = a.gamma + (int64_t)c * a.g_sc + t0 * a.g_st;
    const int64_t gt = a.g_st, gk = a.g_sk;
    double al[S], bt[S], ls = 0.0;
#pragma unroll
    for (int k = 0; k < S; k++) al[k] = a.fae[ib * S + k];
    LcChunk<S> cur, nxt;
    lc_fetch<S>(cur, l, n);
    for (int64_t i0 = 0; i0 < n; i0 += C) { // filter forward, parked in gamma
        lc_fetch<S>(nxt, l + (i0 + C) * S, n - i0 - C);
#pragma unroll
        for (int j = 0; j < C; j++)
            if (i0 + j < n) {
                double bv[S], m;
                lc_emis<S>(cur.v + j * S, bv, m);
                lc_fwd_vec<S>(a.pi, al, bv);
                ls += log(lc_normalise<S>(al)) + m;
#pragma unroll
                for (int k = 0; k < S; k++) g[(i0 + j) * gt + k * gk] = al[k];
            }
        cur = nxt;
    }
    a.fls[ib] = ls;
#pragma unroll
    for (int k = 0; k < S; k++) bt[k] = a.fbe[ib * S + k];
    // beta backward, gamma = alpha o beta / sum; emissions and parked filter fetched a chunk ahead
    constexpr int CB = S == 2 ? 8 : 4;
    double ec[CB * S], en[CB * S], gc[CB * S], gn[CB * S];
    int64_t i0 = n > 0 ? (n - 1) / CB * CB : 0;
#pragma unroll
    for (int i = 0; i < CB * S; i++) {
        ec[i] = i < (n - i0) * S ? l[i0 * S + i] : 0.0;
        gc[i] = i < (n - i0) * S ? g[(i0 + i / S) * gt + (i % S) * gk] : 0.0;
    }
    for (; i0 >= 0 && n > 0; i0 -= CB) {
#pragma unroll
        for (int i = 0; i < CB * S; i++) {
            en[i] = i0 >= CB ? l[(i0 - CB) * S + i] : 0.0;
            gn[i] = i0 >= CB ? g[(i0 - CB + i / S) * gt + (i % S) * gk] : 0.0;
        }
#pragma unroll
        for (int j = CB - 1; j >= 0; j--)
            if (i0 + j < n) {
                double gv[S], bv[S], m;
#pragma unroll
                for (int k = 0; k < S; k++) gv[k] = gc[j * S + k] * bt[k];
                lc_normalise<S>(gv);
#pragma unroll
                for (int k = 0; k < S; k++) g[(i0 + j) * gt + k * gk] = gv[k];
                lc_emis<S>(ec + j * S, bv, m);
                lc_bwd_vec<S>(a.pi, bt, bv);
                lc_normalise<S>(bt);
            }
#pragma unroll
        for (int i = 0; i < CB * S; i++) {
            ec[i] = en[i];
            gc[i] = gn[i];
        }
    }
    if (b == 0) { // position 0 belongs to no block: bt is now beta_0
        double gv[S];
        lc_fb_begin<S>(a, c, gv);
#pragma unroll
        for (int k = 0; k < S; k++) gv[k] *= bt[k];
        lc_normalise<S>(gv);
#pragma unroll
        for (int k = 0; k < S; k++) a.gamma[(int64_t)c * a.g_sc + k * gk] = gv[k];
    }
}

// ---- F3: log-likelihood; chains of one position
template <int S>
HB_HD void lc_fb_ll(const LcArgs &a, int32_t c)
{
    const double *l0 = a.lb + (int64_t)c * a.T * S;
    double bv[S], m, al[S];
    lc_emis<S>(l0, bv, m);
#pragma unroll
    for (int k = 0; k < S; k++) al[k] = a.dl[k] * bv[k];
    double ll = log(lc_normalise<S>(al)) + m;
    if (a.nb == 0) {
#pragma unroll
        for (int k = 0; k < S; k++) a.gamma[(int64_t)c * a.g_sc + k * a.g_sk] = al[k];
    }
    for (int64_t b = 0; b < a.nb; b++) ll += a.fls[(int64_t)c * a.nb + b];
    if (a.ll) a.ll[c] = ll;
}

// Block length: the sequential passes cost ~ T / L, the parallel ones ~ L.
inline int32_t lc_block_len(int64_t T)
{
    int32_t L = 16;
    while (L < 256 && (int64_t)L * L * 4 < T) L *= 2;
    return L;
}

} // namespace hb
