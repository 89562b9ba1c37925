// hb_emul.cpp — TEST-ONLY host build of the CUDA chain bodies.
//
// The per-thread bodies in honeybadger_b200/csrc/*.cuh are `__host__ __device__`; this file compiles
// them with g++ and walks the (block, thread) space serially so the CPU test-suite can check the
// kernels' LOGIC (packing, checkpoints, replay, scaling, tie handling) against the oracle without a
// GPU.  It is built into tests/emul/libhb_emul.so by tests/conftest.py and is never loaded by the
// honeybadger_b200 package: the product has no CPU path.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "hb_binom2.cuh"
#include "hb_binom2_fast.cuh"
#include "hb_host_math.h"
#include "hb_longchain.cuh"
#include "hb_norm3.cuh"
#include "hb_norm3_fast.cuh"
#include "hb_x87.cuh"
static int64_t g_x80p_cover[8];
#define HB_X80P_COVER(i) (g_x80p_cover[(i)]++)
#include "hb_x87_pair.cuh"

using namespace hb;

namespace {
struct Seg {
    std::vector<int64_t> off, chunk;
    std::vector<int32_t> order;
    int64_t nchunks;
};
Seg make_seg(const int64_t *seg_off, int nseg)
{
    Seg s;
    s.off.assign(seg_off, seg_off + nseg + 1);
    s.chunk.resize(nseg + 1);
    int64_t acc = 0;
    for (int i = 0; i <= nseg; i++) {
        s.chunk[i] = acc;
        if (i < nseg) acc += (seg_off[i + 1] - seg_off[i] + HB_CHUNK - 1) / HB_CHUNK;
    }
    s.nchunks = acc;
    s.order.resize(nseg);
    std::iota(s.order.begin(), s.order.end(), 0);
    std::stable_sort(s.order.begin(), s.order.end(), [&](int a, int b) {
        return (seg_off[a + 1] - seg_off[a]) > (seg_off[b + 1] - seg_off[b]);
    });
    return s;
}
} // namespace

static int g_tiled = 0; // 1: the buffers passed in use the warp-tiled layout (B padded to 32)
static int64_t pad32(int64_t B) { return (B + 31) / 32 * 32; }

extern "C" {

void emul_set_tiled(int t) { g_tiled = t; }
static int g_packed = 0; // 1: `x` holds (n << 16) | x words, `n` is ignored
void emul_set_packed(int p) { g_packed = p; }

// flags: 1 = Viterbi, 2 = gamma, 4 = ll ; returns status flag (1 = underflow)
int emul_norm3(const double *x, int64_t ld, int64_t T, int64_t B, const int64_t *seg_off, int nseg,
               const double *mean, const double *sd, int sd_per_seg, const double *Pi,
               const double *delta, int flags, int force_general, uint8_t *y, double *gamma,
               double *ll)
{
    Seg sg = make_seg(seg_off, nseg);
    Norm3Args a;
    memset(&a, 0, sizeof a);
    for (int i = 0; i < 3; i++) {
        a.c.mean[i] = mean[i];
        a.c.delta[i] = delta[i];
        a.c.logdelta[i] = log(delta[i]);
    }
    for (int i = 0; i < 9; i++) {
        a.c.Pi[i] = Pi[i];
        a.c.logPi[i] = log(Pi[i]);
    }
    bool sym_pi = Pi[0] == Pi[4] && Pi[0] == Pi[8] && Pi[1] == Pi[2] && Pi[1] == Pi[3] &&
                  Pi[1] == Pi[5] && Pi[1] == Pi[6] && Pi[1] == Pi[7] && Pi[1] > 0;
    bool sym = sym_pi && (mean[0] - mean[1]) == -(mean[2] - mean[1]) && !force_general;
    a.c.la = a.c.logPi[0];
    a.c.lt = a.c.logPi[1];
    a.c.amt = Pi[0] - Pi[1];
    a.c.tt = Pi[1];
    norm3_fill_k(a.c);
    int64_t nsd = sd_per_seg ? (int64_t)nseg * B : B;
    std::vector<double> inv(nsd), lg(nsd);
    for (int64_t i = 0; i < nsd; i++) {
        inv[i] = 1.0 / sd[i];
        lg[i] = log(sd[i]);
    }
    std::vector<uint64_t> psi((size_t)sg.nchunks * pad32(B));
    std::vector<double> ckpt((size_t)sg.nchunks * 3 * pad32(B));
    std::vector<double2> sva(HB_CHUNK), svb(HB_CHUNK);
    std::vector<double> se2(HB_CHUNK), sxs(HB_CHUNK);
    Norm3Smem sm;
    sm.va = sva.data(); sm.vb = svb.data(); sm.e2 = se2.data(); sm.xs = sxs.data(); sm.st = 1;
    int32_t status = 0;
    a.x = x; a.ld = ld; a.T = T; a.B = B;
    a.seg_off = sg.off.data(); a.chunk_off = sg.chunk.data(); a.seg_order = sg.order.data();
    a.nseg = nseg; a.nblk_per_seg = 1;
    a.sd = sd; a.inv_sd = inv.data(); a.log_sd = lg.data(); a.sd_per_seg = sd_per_seg;
    a.y = y; a.ldy = ld; a.gamma = gamma; a.ldg = ld; a.nchunks = sg.nchunks;
    a.tiled = g_tiled; a.gplane = g_tiled ? T * pad32(B) : T * ld; a.ll = (flags & 4) ? ll : nullptr;
    a.psi = psi.data(); a.ckpt = ckpt.data(); a.status = &status;
    const bool vit = flags & 1, fb = flags & 6;
    for (int s = 0; s < nseg; s++)
        for (int64_t b = 0; b < B; b++) {
            int seg = sg.order[s];
            const bool ll_on = a.ll != nullptr;
            if (sym) {
                if (vit && fb && ll_on) norm3_chain<true, true, true, true>(a, seg, b, sm);
                else if (vit && fb) norm3_chain<true, true, true, false>(a, seg, b, sm);
                else if (fb && ll_on) norm3_chain<true, false, true, true>(a, seg, b, sm);
                else if (fb) norm3_chain<true, false, true, false>(a, seg, b, sm);
                else norm3_chain<true, true, false, false>(a, seg, b, sm);
            } else {
                if (vit && fb && ll_on) norm3_chain<false, true, true, true>(a, seg, b, sm);
                else if (vit && fb) norm3_chain<false, true, true, false>(a, seg, b, sm);
                else if (fb && ll_on) norm3_chain<false, false, true, true>(a, seg, b, sm);
                else if (fb) norm3_chain<false, false, true, false>(a, seg, b, sm);
                else norm3_chain<false, true, false, false>(a, seg, b, sm);
            }
        }
    return status;
}

// The speculate-and-verify kernel body (hb_norm3_fast.cuh).  eps_scale: 1 = product setting,
// +Inf = every chain takes the exact fallback, 0 = never verify (shows what speculation alone gives).
// Returns status; *nflag receives the number of chains re-run exactly.
int emul_norm3_fast(const double *x, int64_t ld, int64_t T, int64_t B, const int64_t *seg_off,
                    int nseg, const double *mean, const double *sd, int sd_per_seg,
                    const double *Pi, const double *delta, int flags, double eps_scale, uint8_t *y,
                    double *gamma, double *ll, int32_t *nflag)
{
    Seg sg = make_seg(seg_off, nseg);
    Norm3Args a;
    memset(&a, 0, sizeof a);
    for (int i = 0; i < 3; i++) {
        a.c.mean[i] = mean[i];
        a.c.delta[i] = delta[i];
        a.c.logdelta[i] = log(delta[i]);
    }
    for (int i = 0; i < 9; i++) {
        a.c.Pi[i] = Pi[i];
        a.c.logPi[i] = log(Pi[i]);
    }
    a.c.la = a.c.logPi[0];
    a.c.lt = a.c.logPi[1];
    a.c.amt = Pi[0] - Pi[1];
    a.c.tt = Pi[1];
    a.c.tau = a.c.lt - a.c.la;
    a.c.dl0 = a.c.logdelta[0] - a.c.logdelta[1];
    a.c.dl2 = a.c.logdelta[2] - a.c.logdelta[1];
    a.c.mumax2 = std::max(mean[0] * mean[0], std::max(mean[1] * mean[1], mean[2] * mean[2]));
    a.c.dmu2 = std::max((mean[0] - mean[1]) * (mean[0] - mean[1]), (mean[2] - mean[1]) * (mean[2] - mean[1]));
    a.c.ldabs = 0;
    for (int i = 0; i < 3; i++)
        if (std::isfinite(a.c.logdelta[i])) a.c.ldabs = std::max(a.c.ldabs, fabs(a.c.logdelta[i]));
    a.c.eps_scale = eps_scale;
    norm3_fill_k(a.c);
    int64_t nsd = sd_per_seg ? (int64_t)nseg * B : B;
    std::vector<double> inv(nsd), lg(nsd);
    for (int64_t i = 0; i < nsd; i++) {
        inv[i] = 1.0 / sd[i];
        lg[i] = log(sd[i]);
    }
    std::vector<uint64_t> psi((size_t)sg.nchunks * pad32(B));
    std::vector<double> ckpt((size_t)sg.nchunks * 3 * pad32(B));
    std::vector<double2> sva(HB_CHUNK), svb(HB_CHUNK);
    std::vector<double> se2(HB_CHUNK), sxs(HB_CHUNK);
    std::vector<int32_t> sph(HB_CHUNK);
    Norm3Smem sm;
    sm.va = sva.data(); sm.vb = svb.data(); sm.e2 = se2.data(); sm.xs = sxs.data();
    sm.ph = sph.data(); sm.st = 1;
    int32_t status[4] = {0, 0, 0, 0};
    a.x = x; a.ld = ld; a.T = T; a.B = B;
    a.seg_off = sg.off.data(); a.chunk_off = sg.chunk.data(); a.seg_order = sg.order.data();
    a.nseg = nseg; a.nblk_per_seg = 1;
    a.sd = sd; a.inv_sd = inv.data(); a.log_sd = lg.data(); a.sd_per_seg = sd_per_seg;
    a.y = y; a.ldy = ld; a.gamma = gamma; a.ldg = ld; a.nchunks = sg.nchunks;
    a.tiled = g_tiled; a.gplane = g_tiled ? T * pad32(B) : T * ld; a.ll = (flags & 4) ? ll : nullptr;
    a.psi = psi.data(); a.ckpt = ckpt.data(); a.status = status;
    const bool vit = flags & 1, fb = flags & 6, llon = flags & 4;
    for (int s = 0; s < nseg; s++)
        for (int64_t b = 0; b < B; b++) {
            int seg = sg.order[s];
            BulkRing br;
            br.par[0] = br.par[1] = 0u;
            if (g_tiled) { // the tiled layout takes the bulk-staging form of the body, as on the device
                if (vit && fb && llon) norm3_fast_chain<true, true, true, true>(a, seg, b, sm, br);
                else if (vit && fb) norm3_fast_chain<true, true, false, true>(a, seg, b, sm, br);
                else if (fb && llon) norm3_fast_chain<false, true, true, true>(a, seg, b, sm, br);
                else if (fb) norm3_fast_chain<false, true, false, true>(a, seg, b, sm, br);
                else norm3_fast_chain<true, false, false, true>(a, seg, b, sm, br);
            } else {
                if (vit && fb && llon) norm3_fast_chain<true, true, true, false>(a, seg, b, sm, br);
                else if (vit && fb) norm3_fast_chain<true, true, false, false>(a, seg, b, sm, br);
                else if (fb && llon) norm3_fast_chain<false, true, true, false>(a, seg, b, sm, br);
                else if (fb) norm3_fast_chain<false, true, false, false>(a, seg, b, sm, br);
                else norm3_fast_chain<true, false, false, false>(a, seg, b, sm, br);
            }
        }
    if (nflag) *nflag = status[1];
    return status[0];
}

int emul_binom2(const int32_t *x, const int32_t *n, int64_t ld, int64_t T, int64_t B,
                const int64_t *seg_off, int nseg, const double *prob, const double *Pi,
                const double *delta, int flags, int lut_nmax, uint8_t *y, double *gamma, double *ll)
{
    Seg sg = make_seg(seg_off, nseg);
    Binom2Args a;
    memset(&a, 0, sizeof a);
    for (int k = 0; k < 2; k++) {
        a.c.prob[k] = prob[k];
        a.c.q[k] = 1 - prob[k];
        a.c.logp[k] = log(prob[k]);
        a.c.logq[k] = log(1 - prob[k]);
        a.c.delta[k] = delta[k];
        a.c.logdelta[k] = log(delta[k]);
    }
    for (int i = 0; i < 4; i++) {
        a.c.Pi[i] = Pi[i];
        a.c.logPi[i] = log(Pi[i]);
    }
    size_t cnt = (size_t)(lut_nmax + 1) * (lut_nmax + 2) / 2;
    std::vector<double2> lb(cnt);
    std::vector<RhoME> rr(cnt);
    for (int nn = 0; nn <= lut_nmax; nn++)
        for (int xx = 0; xx <= nn; xx++) {
            size_t ix = (size_t)nn * (nn + 1) / 2 + xx;
            lb[ix].x = hbhost::binom_logpmf(xx, nn, prob[0]);
            lb[ix].y = hbhost::binom_logpmf(xx, nn, prob[1]);
            rr[ix] = rho_from_logratio(lb[ix].x - lb[ix].y);
        }
    std::vector<uint64_t> psi((size_t)sg.nchunks * pad32(B));
    std::vector<double> ckpt((size_t)sg.nchunks * 3 * pad32(B));
    std::vector<double> sm(HB_CHUNK * HB_BINOM2_SLOT);
    int32_t status = 0;
    a.x = x; a.n = n; a.ld = ld; a.T = T; a.B = B;
    a.seg_off = sg.off.data(); a.chunk_off = sg.chunk.data(); a.seg_order = sg.order.data();
    a.nseg = nseg; a.nblk_per_seg = 1;
    a.lut_lb = lb.data(); a.lut_rho = rr.data(); a.lut_nmax = lut_nmax;
    a.y = y; a.ldy = ld; a.gamma = gamma; a.ldg = ld; a.nchunks = sg.nchunks;
    a.tiled = g_tiled; a.gplane = g_tiled ? T * pad32(B) : T * ld; a.ll = (flags & 4) ? ll : nullptr;
    a.psi = psi.data(); a.ckpt = ckpt.data(); a.status = &status;
    const bool vit = flags & 1, fb = flags & 6;
    for (int s = 0; s < nseg; s++)
        for (int64_t b = 0; b < B; b++) {
            int seg = sg.order[s];
            if (vit && fb) binom2_chain<true, true>(a, seg, b, sm.data(), 1);
            else if (fb) binom2_chain<false, true>(a, seg, b, sm.data(), 1);
            else binom2_chain<true, false>(a, seg, b, sm.data(), 1);
        }
    return status;
}

// The lean binomial kernel body (hb_binom2_fast.cuh) followed, when it raises its gate, by the general
// body — the sequence binom2_run() enqueues.  *gate_out tells whether the general body had to run.
int emul_binom2_fast(const int32_t *x, const int32_t *n, int64_t ld, int64_t T, int64_t B,
                     const int64_t *seg_off, int nseg, const double *prob, const double *Pi,
                     const double *delta, int flags, int lut_nmax, uint8_t *y, double *gamma,
                     double *ll, int32_t *gate_out)
{
    Seg sg = make_seg(seg_off, nseg);
    Binom2Args a;
    memset(&a, 0, sizeof a);
    for (int k = 0; k < 2; k++) {
        a.c.prob[k] = prob[k];
        a.c.q[k] = 1 - prob[k];
        a.c.logp[k] = log(prob[k]);
        a.c.logq[k] = log(1 - prob[k]);
        a.c.delta[k] = delta[k];
        a.c.logdelta[k] = log(delta[k]);
    }
    for (int i = 0; i < 4; i++) {
        a.c.Pi[i] = Pi[i];
        a.c.logPi[i] = log(Pi[i]);
    }
    size_t cnt = (size_t)(lut_nmax + 1) * (lut_nmax + 2) / 2;
    std::vector<double2> lb(cnt);
    std::vector<RhoME> rr(cnt);
    std::vector<double> rd(cnt);
    for (int nn = 0; nn <= lut_nmax; nn++)
        for (int xx = 0; xx <= nn; xx++) {
            size_t ix = (size_t)nn * (nn + 1) / 2 + xx;
            lb[ix].x = hbhost::binom_logpmf(xx, nn, prob[0]);
            lb[ix].y = hbhost::binom_logpmf(xx, nn, prob[1]);
            rr[ix] = rho_from_logratio(lb[ix].x - lb[ix].y);
            rd[ix] = (lb[ix].x == -INFINITY && lb[ix].y == -INFINITY) ? 1.0 : b2f_rho_double(rr[ix]);
        }
    std::vector<uint64_t> psi((size_t)sg.nchunks * pad32(B));
    std::vector<double> ckpt((size_t)sg.nchunks * 3 * pad32(B));
    std::vector<double2> spr(HB_CHUNK);
    std::vector<int32_t> sph(HB_CHUNK), sxs(HB_CHUNK), sns(HB_CHUNK);
    std::vector<double> sm(HB_CHUNK * HB_BINOM2_SLOT);
    B2fSmem m;
    m.pr = spr.data(); m.ph = sph.data(); m.xs = sxs.data(); m.ns = sns.data(); m.st = 1;
    m.slb = lb.data(); m.srho = rd.data(); m.slut_n = std::min(lut_nmax, HB_B2F_SLUT_N);
    int32_t status[4] = {0, 0, 0, 0};
    a.x = x; a.n = n; a.ld = ld; a.T = T; a.B = B;
    a.seg_off = sg.off.data(); a.chunk_off = sg.chunk.data(); a.seg_order = sg.order.data();
    a.nseg = nseg; a.nblk_per_seg = 1;
    a.lut_lb = lb.data(); a.lut_rho = rr.data(); a.lut_rho_d = rd.data(); a.lut_nmax = lut_nmax;
    a.y = y; a.ldy = ld; a.gamma = gamma; a.ldg = ld; a.nchunks = sg.nchunks;
    a.tiled = g_tiled; a.gplane = g_tiled ? T * pad32(B) : T * ld; a.ll = (flags & 4) ? ll : nullptr;
    a.psi = psi.data(); a.ckpt = ckpt.data(); a.status = status; a.gate = status + 2;
    a.packed = g_packed;
    const bool vit = flags & 1, fb = flags & 6, llon = flags & 4;
    for (int s = 0; s < nseg; s++)
        for (int64_t b = 0; b < B; b++) {
            int seg = sg.order[s];
            if (g_packed) {
                if (vit && fb && llon) binom2_fast_chain<true, true, true, true>(a, seg, b, m);
                else if (vit && fb) binom2_fast_chain<true, true, false, true>(a, seg, b, m);
                else if (fb && llon) binom2_fast_chain<false, true, true, true>(a, seg, b, m);
                else if (fb) binom2_fast_chain<false, true, false, true>(a, seg, b, m);
                else binom2_fast_chain<true, false, false, true>(a, seg, b, m);
            } else {
                if (vit && fb && llon) binom2_fast_chain<true, true, true, false>(a, seg, b, m);
                else if (vit && fb) binom2_fast_chain<true, true, false, false>(a, seg, b, m);
                else if (fb && llon) binom2_fast_chain<false, true, true, false>(a, seg, b, m);
                else if (fb) binom2_fast_chain<false, true, false, false>(a, seg, b, m);
                else binom2_fast_chain<true, false, false, false>(a, seg, b, m);
            }
        }
    if (gate_out) *gate_out = status[2];
    if (status[2]) {
        for (int s = 0; s < nseg; s++)
            for (int64_t b = 0; b < B; b++) {
                int seg = sg.order[s];
                if (vit && fb) binom2_chain<true, true>(a, seg, b, sm.data(), 1);
                else if (fb) binom2_chain<false, true>(a, seg, b, sm.data(), 1);
                else binom2_chain<true, false>(a, seg, b, sm.data(), 1);
            }
    }
    return status[0];
}

// R mean() / sd() through the integer x87 emulation (the arithmetic of pool_mean_kernel / pooled_sd_kernel)
double emul_x87_mean(const double *x, int64_t n)
{
    x80 S = x80_zero(), Tt = x80_zero();
    for (int64_t i = 0; i < n; i++) S = x80_add(S, x80_from_double(x[i]));
    S = x80_div_u32(S, (uint32_t)n);
    for (int64_t i = 0; i < n; i++) Tt = x80_add(Tt, x80_sub(x80_from_double(x[i]), S));
    S = x80_add(S, x80_div_u32(Tt, (uint32_t)n));
    return x80_to_double(S);
}

// Every step of a running sum through x80_add against the host's own x87 unit (long double):
// returns the number of steps whose 80-bit result differs.  mode 0: s += x[i]; mode 1: s += x[i] - c
// (the second pass of mean); mode 2: s += x[i] * x[i+1] style products are not needed on this path.
int64_t emul_x87_addcheck(const double *x, int64_t n, double c, int mode)
{
    x80 S = x80_zero();
    const x80 C80 = x80_from_double(c);
    volatile long double s = 0.0L;
    const long double cl = c;
    int64_t bad = 0;
    for (int64_t i = 0; i < n; i++) {
        if (mode == 0) {
            S = x80_add(S, x80_from_double(x[i]));
            s = s + (long double)x[i];
        } else {
            S = x80_add(S, x80_sub(x80_from_double(x[i]), C80));
            volatile long double t = (long double)x[i] - cl;
            s = s + t;
        }
        long double mine = S.m ? ldexpl((long double)S.m, S.e - 63) : 0.0L;
        if (S.neg) mine = -mine;
        if (mine != s) bad++;
    }
    return bad;
}

// The same check for the double-pair form (hb_x87_pair.cuh).  mode 0: s += x[i]; mode 1: s += (x[i] - c80)
// with c80 = c + c2 (an 80-bit centre).  Returns the number of differing steps (emul_x87_pair_cover tells
// which cases of the rounding step and how many integer fall-backs were taken).
int64_t emul_x87_paircheck(const double *x, int64_t n, double c, double c2, int mode, int64_t *fallbacks)
{
    x80p S = {0.0, 0.0};
    volatile long double cl = (long double)c + (long double)c2;
    const x80p Mneg = x80p_from_x80(x80_neg(x80_add(x80_from_double(c), x80_from_double(c2))));
    volatile long double s = 0.0L;
    int64_t bad = 0;
    for (int64_t i = 0; i < n; i++) {
        if (mode == 0) {
            S = x80p_add_double(S, x[i]);
            s = s + (long double)x[i];
        } else {
            const x80p D = x80p_add_double(Mneg, x[i]);
            volatile long double t = (long double)x[i] - cl;
            if ((long double)D.h + (long double)D.l != t) bad++;
            S = x80p_add(S, D);
            s = s + t;
        }
        if ((long double)S.h + (long double)S.l != s) bad++;
        if (S.h != (double)s) bad++;
    }
    (void)fallbacks;
    return bad;
}

// x80p_add / x80p_add_double on explicit 64-bit operands (significand, exponent, sign as hb::x80 holds them)
// against the host's x87 unit; returns the number of mismatches.  dbl != 0: the second operand is first
// rounded to double and added with x80p_add_double.
int64_t emul_x87_pair_add80(const uint64_t *ms, const int32_t *es, const uint8_t *ns, const uint64_t *md,
                            const int32_t *ed, const uint8_t *nd, int64_t n, int dbl)
{
    int64_t bad = 0;
    for (int64_t i = 0; i < n; i++) {
        x80 A, B;
        A.m = ms[i], A.e = es[i], A.neg = ns[i];
        B.m = md[i], B.e = ed[i], B.neg = nd[i];
        if (A.m) A.m |= 1ull << 63;
        if (B.m) B.m |= 1ull << 63;
        const x80p a = x80p_from_x80(A);
        volatile long double la = (long double)a.h + (long double)a.l, want;
        x80p r;
        if (dbl) {
            const double v = x80_to_double(B);
            r = x80p_add_double(a, v);
            want = la + (long double)v;
        } else {
            const x80p b = x80p_from_x80(B);
            volatile long double lb = (long double)b.h + (long double)b.l;
            r = x80p_add(a, b);
            want = la + lb;
        }
        if ((long double)r.h + (long double)r.l != want || r.h != (double)want) bad++;
    }
    return bad;
}

void emul_x87_pair_cover(int64_t *out, int reset)
{
    for (int i = 0; i < 8; i++) {
        out[i] = g_x80p_cover[i];
        if (reset) g_x80p_cover[i] = 0;
    }
}

double emul_x87_sd(const double *x, int64_t n)
{
    x80 xm = x80_from_double(emul_x87_mean(x, n)), SS = x80_zero();
    for (int64_t i = 0; i < n; i++) {
        x80 d = x80_sub(x80_from_double(x[i]), xm);
        SS = x80_add(SS, x80_mul(d, d));
    }
    return sqrt(x80_to_double(x80_div_u32(SS, (uint32_t)(n - 1))));
}

// R sd() with the running sums taken block by block (x80_align_term / x80_block_add, what pooled_sd_kernel does):
// must equal emul_x87_sd bit for bit.  *fast receives the number of terms added on the integer path, of 3 n.
double emul_x87_sd_blocked(const double *x, int64_t n, int64_t *fast)
{
    x80 mean80 = x80_zero(), xm = x80_zero(), acc = x80_zero();
    int64_t nf = 0;
    for (int pass = 0; pass < 3; pass++) {
        acc = x80_zero();
        for (int64_t i0 = 0; i0 < n; i0 += 32) {
            const int lim = (int)((n - i0) < 32 ? (n - i0) : 32);
            x80 tt[32];
            for (int j = 0; j < lim; j++) {
                x80 v = x80_from_double(x[i0 + j]), term;
                if (pass == 0)
                    term = v;
                else if (pass == 1)
                    term = x80_sub(v, mean80);
                else {
                    x80 d = x80_sub(v, xm);
                    term = x80_mul(d, d);
                }
                tt[j] = term;
            }
            int f = 0;
            acc = x80_block_add(acc, tt, lim, &f);
            nf += f;
        }
        if (pass == 0)
            mean80 = x80_div_u32(acc, (uint32_t)n);
        else if (pass == 1) {
            mean80 = x80_add(mean80, x80_div_u32(acc, (uint32_t)n));
            xm = x80_from_double(x80_to_double(mean80));
        }
    }
    if (fast) *fast = nf;
    return sqrt(x80_to_double(x80_div_u32(acc, (uint32_t)(n - 1))));
}

// A running sum taken block by block against the host's x87 unit after every block (mode 0: s += x[i];
// mode 1: s += x[i] - c); returns the number of blocks whose 80-bit result differs.
int64_t emul_x87_blockcheck(const double *x, int64_t n, double c, int mode, int64_t *fast)
{
    x80 acc = x80_zero();
    const x80 C80 = x80_from_double(c);
    volatile long double sl = 0.0L;
    const long double cl = c;
    int64_t bad = 0, nf = 0;
    for (int64_t i0 = 0; i0 < n; i0 += 32) {
        const int lim = (int)((n - i0) < 32 ? (n - i0) : 32);
        x80 tt[32];
        for (int j = 0; j < lim; j++) {
            tt[j] = mode == 0 ? x80_from_double(x[i0 + j]) : x80_sub(x80_from_double(x[i0 + j]), C80);
            if (mode == 0)
                sl = sl + (long double)x[i0 + j];
            else {
                volatile long double t = (long double)x[i0 + j] - cl;
                sl = sl + t;
            }
        }
        int f = 0;
        acc = x80_block_add(acc, tt, lim, &f);
        nf += f;
        long double mine = acc.m ? ldexpl((long double)acc.m, acc.e - 63) : 0.0L;
        if (acc.neg) mine = -mine;
        if (mine != sl) bad++;
    }
    if (fast) *fast = nf;
    return bad;
}

double emul_div_by_const(double d, double b) { return div_by_const(d, b, 1.0 / b); }

void emul_exp_pair(double u, double *ep, double *em)
{
    double Pp, Pm;
    int32_t n;
    Norm3Const c;
    norm3_fill_k(c);
    norm3_exp_pair(c, u, Pp, Pm, n); // the kernel's exp(u), exp(-u)
    *ep = ldexp(Pp, n);
    *em = ldexp(Pm, -n);
}

double emul_binom_logpmf_host(double x, double n, double p) { return hbhost::binom_logpmf(x, n, p); }
// the product's beta-binomial table entry (prefix-sum builder, as prepare_lut uses it)
double emul_betabinom_logpmf(int x, int n, double p, double rho, int nmax)
{
    static hbhost::BetaBinomCorr *c = nullptr;
    static double cp = -1, cr = -1;
    static int cn = -1;
    if (!c || cp != p || cr != rho || cn != nmax) {
        delete c;
        c = new hbhost::BetaBinomCorr(p, rho, nmax);
        cp = p, cr = rho, cn = nmax;
    }
    return hbhost::betabinom_logpmf(*c, x, n, p);
}
double emul_binom_logpmf_dev(double x, double n, double p)
{
    return dev_dbinom_log(x, n, p, 1 - p, log(p), log(1 - p));
}

} // extern "C"

// Gene-axis parallel Viterbi for long chains (hb_longchain.cuh), phases walked serially; the
// sequential walkers go through staging windows (copies of the per-block tables) like the device's.
// lb [K][T][S]; y [K][T]; stats [K][4] (first four cstate words).  L <= 0: the library's block length.
template <int S>
static int emul_longchain_t(int64_t T, int K, const double *lb, const double *Pi, const double *delta,
                            int L, int32_t *y, int32_t *stats)
{
    LcArgs a;
    memset(&a, 0, sizeof a);
    a.lb = lb; a.T = T; a.K = K; a.S = S; a.L = L > 0 ? L : lc_block_len(T);
    a.nb = T > 1 ? (T - 1 + a.L - 1) / a.L : 0;
    for (int i = 0; i < S * S; i++) a.lp[i] = log(Pi[i]);
    for (int i = 0; i < S; i++) a.ld[i] = log(delta[i]);
    const size_t nkb = (size_t)K * a.nb + 1;
    std::vector<double> fmat(nkb * S * S), entry(nkb * S), lexit(nkb * S);
    std::vector<int32_t> bexp(nkb), cstate((size_t)K * HB_LC_CS, 0);
    std::vector<int64_t> imat(nkb * 2 * S * S);
    std::vector<uint8_t> iok(nkb), psi((size_t)K * T), bmap(nkb), ybound(nkb);
    int32_t status = 0;
    a.fmat = fmat.data(); a.bexp = bexp.data(); a.imat = imat.data(); a.iok = iok.data();
    a.entry = entry.data(); a.lexit = lexit.data(); a.psi = psi.data(); a.bmap = bmap.data();
    a.cstate = cstate.data(); a.ybound = ybound.data(); a.y = y; a.status = &status;
    for (int c = 0; c < K; c++)
        for (int64_t b = 0; b < a.nb; b++) lc_fmat<S>(a, c, b);
    for (int c = 0; c < K && a.nb; c++) {
        double v[S];
        lc_prefix_approx_begin<S>(a, c, v);
        for (int64_t w0 = 0; w0 < a.nb; w0 += HB_LC_WIN) {
            const int64_t n = std::min<int64_t>(HB_LC_WIN, a.nb - w0);
            std::vector<double> wf(fmat.begin() + ((size_t)c * a.nb + w0) * S * S,
                                   fmat.begin() + ((size_t)c * a.nb + w0 + n) * S * S);
            int32_t be[HB_LC_WIN];
            LcArgs w = a;
            w.win_b0 = w0; w.wf = wf.data();
            lc_prefix_approx_range<S>(w, c, w0, w0 + n, v, be);
            for (int64_t i = 0; i < n; i++) bexp[(size_t)c * a.nb + w0 + i] = be[i];
        }
    }
    for (int c = 0; c < K; c++)
        for (int64_t b = 0; b < a.nb; b++) lc_imat<S>(a, c, b);
    for (int r = 0; r < HB_LC_ROUNDS && a.nb; r++) {
        for (int c = 0; c < K; c++) {
            double nu[S];
            int64_t b0, forced_end;
            if (!lc_prefix_exact_begin<S>(a, c, r, nu, b0, forced_end)) continue;
            int32_t nlit = 0;
            for (int64_t w0 = b0; w0 < a.nb; w0 += HB_LC_WIN) {
                const int64_t n = std::min<int64_t>(HB_LC_WIN, a.nb - w0);
                const size_t o = (size_t)c * a.nb + w0;
                std::vector<int64_t> wi(imat.begin() + o * 2 * S * S, imat.begin() + (o + n) * 2 * S * S);
                std::vector<int32_t> wbe(bexp.begin() + o, bexp.begin() + o + n);
                std::vector<uint8_t> wio(iok.begin() + o, iok.begin() + o + n);
                LcArgs w = a;
                w.win_b0 = w0; w.wi = wi.data(); w.wbe = wbe.data(); w.wio = wio.data();
                lc_prefix_exact_range<S>(w, c, w0, w0 + n, nu, forced_end, nlit);
            }
            cstate[(size_t)c * HB_LC_CS + 3] += nlit;
        }
        for (int c = 0; c < K; c++)
            for (int64_t b = 0; b < a.nb; b++) lc_literal<S>(a, c, b);
    }
    for (int c = 0; c < K; c++) lc_resolve<S>(a, c);
    for (int c = 0; c < K; c++)
        for (int64_t b = 0; b < a.nb; b++) lc_write_y<S>(a, c, b);
    if (stats)
        for (int c = 0; c < K; c++) memcpy(stats + 4 * c, cstate.data() + (size_t)c * HB_LC_CS, 4 * sizeof(int32_t));
    return status;
}

extern "C" int emul_longchain(int S, int64_t T, int K, const double *lb, const double *Pi,
                              const double *delta, int L, int32_t *y, int32_t *stats)
{
    if (S == 2) return emul_longchain_t<2>(T, K, lb, Pi, delta, L, y, stats);
    if (S == 3) return emul_longchain_t<3>(T, K, lb, Pi, delta, L, y, stats);
    return -1;
}

// Posteriors of long chains (lc_fb_*), phases walked serially.  gamma [K][T][S] (or, after
// emul_set_fb_planes(1), R's [T x K x S] column-major), ll [K].
static int g_fb_planes = 0;
extern "C" void emul_set_fb_planes(int v) { g_fb_planes = v; }
template <int S>
static void emul_longchain_fb_t(int64_t T, int K, const double *lb, const double *Pi, const double *delta,
                                int L, double *gamma, double *ll)
{
    LcArgs a;
    memset(&a, 0, sizeof a);
    a.lb = lb; a.T = T; a.K = K; a.S = S; a.L = L > 0 ? L : lc_block_len(T);
    a.nb = T > 1 ? (T - 1 + a.L - 1) / a.L : 0;
    for (int i = 0; i < S * S; i++) a.pi[i] = Pi[i];
    for (int i = 0; i < S; i++) a.dl[i] = delta[i];
    const size_t nkb = (size_t)K * a.nb + 1;
    std::vector<double> ffm(nkb * S * S), fae(nkb * S), fbe(nkb * S), fls(nkb);
    std::vector<int32_t> ffe(nkb * S);
    a.ffm = ffm.data(); a.ffe = ffe.data(); a.fae = fae.data(); a.fbe = fbe.data(); a.fls = fls.data();
    a.gamma = gamma; a.ll = ll;
    if (g_fb_planes) { a.g_sc = T; a.g_st = 1; a.g_sk = (int64_t)K * T; }   // R's [T x K x S] column-major
    else { a.g_sc = T * S; a.g_st = S; a.g_sk = 1; }
    for (int c = 0; c < K; c++)
        for (int64_t b = 0; b < a.nb; b++) lc_fb_mat<S>(a, c, b);
    for (int c = 0; c < K && a.nb; c++) {
        double al[S], bt[S];
        lc_fb_begin<S>(a, c, al);
        for (int64_t w0 = 0; w0 < a.nb; w0 += HB_LC_WIN) {
            const int64_t n = std::min<int64_t>(HB_LC_WIN, a.nb - w0);
            const size_t o = (size_t)c * a.nb + w0;
            std::vector<double> wf(ffm.begin() + o * S * S, ffm.begin() + (o + n) * S * S);
            std::vector<int32_t> we(ffe.begin() + o * S, ffe.begin() + (o + n) * S);
            LcArgs w = a;
            w.win_b0 = w0; w.wff = wf.data(); w.wfe = we.data();
            lc_fb_fwd_range<S>(w, c, w0, w0 + n, al);
        }
        for (int k = 0; k < S; k++) bt[k] = 1.0 / S;
        for (int64_t w1 = a.nb; w1 > 0; w1 -= HB_LC_WIN) {
            const int64_t w0 = std::max<int64_t>(0, w1 - HB_LC_WIN);
            lc_fb_bwd_range<S>(a, c, w0, w1, bt);
        }
    }
    for (int c = 0; c < K; c++)
        for (int64_t b = 0; b < a.nb; b++) lc_fb_block<S>(a, c, b);
    for (int c = 0; c < K; c++) lc_fb_ll<S>(a, c);
}

extern "C" int emul_longchain_fb(int S, int64_t T, int K, const double *lb, const double *Pi,
                                 const double *delta, int L, double *gamma, double *ll)
{
    if (S == 2) emul_longchain_fb_t<2>(T, K, lb, Pi, delta, L, gamma, ll);
    else if (S == 3) emul_longchain_fb_t<3>(T, K, lb, Pi, delta, L, gamma, ll);
    else return -1;
    return 0;
}
