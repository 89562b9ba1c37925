/*
 * hb_oracle.c — CPU restatement of the arithmetic on HoneyBADGER's HMM segmentation path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The shipped path (honeybadger_b200/) never
 * links, imports or calls anything in oracle/.
 *
 * PARITY UNPINNED.  The reference (an R package) delegates this arithmetic to third-party code that
 * is not under /root/reference and cannot be run in this environment (no R):
 *   - CRAN `HiddenMarkov` (DESCRIPTION:21, no version pin; 1.8-x era): dthmm, Viterbi, forwardback
 *   - base-R `stats`/nmath (R 3.3.x era, classic stirlerr): dnorm, dbinom, mean, var/sd
 * called from R/HoneyBADGER.R:1141-1151, :1404-1410, R/HoneyBADGER_gexp.R:494-504,
 * R/HoneyBADGER_allele.R:574-580.  The functions below restate the published algorithms of those
 * dependencies operation for operation, in the same evaluation order, with
 * no FMA contraction (build with -O2 -ffp-contract=off -mno-fma) and x87 `long double` where R
 * accumulates in LDOUBLE.  They are cross-checked against mpmath and an independent pure-Python
 * restatement (oracle/pyref.py) and against the bundled-data known answers (tests/golden).
 * What the reference's own rendered outputs do pin (tests/test_inputs_oracle.py): the setAlleleMats SNP
 * counts printed in docs/Getting_Started.md:230 and docs/Integrating.md:27, and the chromosomes of the
 * regions the stateless calcAlleleCnvBoundaries finds on the bundled data
 * (docs/figure/Integrating/unnamed-chunk-6-1.png), and — through the GPU suite — the regions printed
 * after the RefClass allele run (docs/Getting_Started.md:287-293; three of four to the base pair) —
 * end-to-end answers, not per-step arithmetic.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HBO_OK 0
#define HBO_ERR_ARG (-1)
#define HBO_ERR_UNDERFLOW (-3) /* HiddenMarkov::Viterbi: stop("Problems With Underflow") */
#define HBO_ERR_NOMEM (-4)

#define HBO_MAX_S 8

#define M_LN_SQRT_2PI_ 0.918938533204672741780329736406 /* log(sqrt(2*pi)) */
#define M_LN_2PI_ 1.837877066409345483560659472811     /* log(2*pi) */
#define M_1_SQRT_2PI_ 0.398942280401432677939946059934  /* 1/sqrt(2pi) */
#define M_LN2_ 0.693147180559945309417232121458

/* ---------------------------------------------------------------- nmath: dnorm */

/* stats::dnorm(x, mean, sd, log=TRUE) — nmath dnorm4, log branch. */
double hbo_dnorm_log(double x, double mu, double sigma)
{
    if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
    if (!isfinite(sigma)) return -INFINITY;
    if (!isfinite(x) && mu == x) return NAN;
    if (sigma <= 0) {
        if (sigma < 0) return NAN;
        return (x == mu) ? INFINITY : -INFINITY;
    }
    x = (x - mu) / sigma;
    if (!isfinite(x)) return -INFINITY;
    x = fabs(x);
    if (x >= 2 * sqrt(DBL_MAX)) return -INFINITY;
    return -(M_LN_SQRT_2PI_ + 0.5 * x * x + log(sigma));
}

/* stats::dnorm(x, mean, sd, log=FALSE) — nmath dnorm4, accurate (R >= 3.1) linear branch. */
double hbo_dnorm(double x, double mu, double sigma)
{
    if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
    if (!isfinite(sigma)) return 0.;
    if (!isfinite(x) && mu == x) return NAN;
    if (sigma <= 0) {
        if (sigma < 0) return NAN;
        return (x == mu) ? INFINITY : 0.;
    }
    x = (x - mu) / sigma;
    if (!isfinite(x)) return 0.;
    x = fabs(x);
    if (x >= 2 * sqrt(DBL_MAX)) return 0.;
    if (x < 5) return M_1_SQRT_2PI_ * exp(-0.5 * x * x) / sigma;
    if (x > sqrt(-2 * M_LN2_ * (DBL_MIN_EXP + 1 - DBL_MANT_DIG))) return 0.;
    double x1 = ldexp(nearbyint(ldexp(x, 16)), -16);
    double x2 = x - x1;
    return M_1_SQRT_2PI_ / sigma * (exp(-0.5 * x1 * x1) * exp((-0.5 * x2 - x1) * x2));
}

/* --------------------------------------------------------------- nmath: dbinom */

static const double sferr_halves[31] = {
    0.0,                           /* n=0 - wrong, place holder only */
    0.1534264097200273452913848,   /* 0.5 */
    0.0810614667953272582196702,   /* 1.0 */
    0.0548141210519176538961390,   /* 1.5 */
    0.0413406959554092940938221,   /* 2.0 */
    0.03316287351993628748511048,  /* 2.5 */
    0.02767792568499833914878929,  /* 3.0 */
    0.02374616365629749597132920,  /* 3.5 */
    0.02079067210376509311152277,  /* 4.0 */
    0.01848845053267318523077934,  /* 4.5 */
    0.01664469118982119216319487,  /* 5.0 */
    0.01513497322191737887351255,  /* 5.5 */
    0.01387612882307074799874573,  /* 6.0 */
    0.01281046524292022692424986,  /* 6.5 */
    0.01189670994589177009505572,  /* 7.0 */
    0.01110455975820691732662991,  /* 7.5 */
    0.010411265261972096497478567, /* 8.0 */
    0.009799416126158803298389475, /* 8.5 */
    0.009255462182712732917728637, /* 9.0 */
    0.008768700134139385462952823, /* 9.5 */
    0.008330563433362871256469318, /* 10.0 */
    0.007934114564314020547248100, /* 10.5 */
    0.007573675487951840794972024, /* 11.0 */
    0.007244554301320383179543912, /* 11.5 */
    0.006942840107209529865664152, /* 12.0 */
    0.006665247032707682442354394, /* 12.5 */
    0.006408994188004207068439631, /* 13.0 */
    0.006171712263039457647532867, /* 13.5 */
    0.005951370112758847735624416, /* 14.0 */
    0.005746216513010115682023589, /* 14.5 */
    0.005554733551962801371038690  /* 15.0 */
};

/* nmath stirlerr(n) = log(n!) - log(sqrt(2*pi*n)*(n/e)^n), classic (R <= 4.3) form. */
double hbo_stirlerr(double n)
{
    const double S0 = 0.083333333333333333333;        /* 1/12 */
    const double S1 = 0.00277777777777777777778;      /* 1/360 */
    const double S2 = 0.00079365079365079365079365;   /* 1/1260 */
    const double S3 = 0.000595238095238095238095238;  /* 1/1680 */
    const double S4 = 0.0008417508417508417508417508; /* 1/1188 */
    double nn;
    if (n <= 15.0) {
        nn = n + n;
        if (nn == (int)nn) return sferr_halves[(int)nn];
        return lgamma(n + 1.) - (n + 0.5) * log(n) + n - M_LN_SQRT_2PI_;
    }
    nn = n * n;
    if (n > 500) return (S0 - S1 / nn) / n;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

/* nmath bd0(x, np) = x log(x/np) + np - x, Loader's series near x == np. */
double hbo_bd0(double x, double np)
{
    double ej, s, s1, v;
    int j;
    if (!isfinite(x) || !isfinite(np) || np == 0.0) return NAN;
    if (fabs(x - np) < 0.1 * (x + np)) {
        v = (x - np) / (x + np);
        s = (x - np) * v;
        if (fabs(s) < DBL_MIN) return s;
        ej = 2 * x * v;
        v = v * v;
        for (j = 1; j < 1000; j++) {
            ej *= v;
            s1 = s + ej / ((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

/* nmath dbinom_raw(x, n, p, q, give_log). */
double hbo_dbinom_raw(double x, double n, double p, double q, int give_log)
{
    double lf, lc;
    if (p == 0) return (x == 0) ? (give_log ? 0. : 1.) : (give_log ? -INFINITY : 0.);
    if (q == 0) return (x == n) ? (give_log ? 0. : 1.) : (give_log ? -INFINITY : 0.);
    if (x == 0) {
        if (n == 0) return give_log ? 0. : 1.;
        lc = (p < 0.1) ? -hbo_bd0(n, n * q) - n * p : n * log(q);
        return give_log ? lc : exp(lc);
    }
    if (x == n) {
        lc = (q < 0.1) ? -hbo_bd0(n, n * p) - n * q : n * log(p);
        return give_log ? lc : exp(lc);
    }
    if (x < 0 || x > n) return give_log ? -INFINITY : 0.;
    lc = hbo_stirlerr(n) - hbo_stirlerr(x) - hbo_stirlerr(n - x) - hbo_bd0(x, n * p) -
         hbo_bd0(n - x, n * q);
    lf = M_LN_2PI_ + log(x) + log1p(-x / n);
    return give_log ? (lc - 0.5 * lf) : exp(lc - 0.5 * lf);
}

/* stats::dbinom(x, size, prob, log) for integer-valued x, size. */
double hbo_dbinom(double x, double n, double p, int give_log)
{
    if (isnan(x) || isnan(n) || isnan(p)) return x + n + p;
    if (p < 0 || p > 1 || n < 0 || fabs(n - nearbyint(n)) > 1e-7 * fmax(1., fabs(n))) return NAN;
    if (fabs(x - nearbyint(x)) > 1e-7 * fmax(1., fabs(x)) || x < 0 || !isfinite(x))
        return give_log ? -INFINITY : 0.;
    n = nearbyint(n);
    x = nearbyint(x);
    return hbo_dbinom_raw(x, n, p, 1 - p, give_log);
}

/* ------------------------------------------------- base R: mean / sd / sum */

/* mean(x) for a double vector (summary.c real_mean): LDOUBLE two-pass. */
double hbo_r_mean(const double *x, int64_t n)
{
    long double s = 0.0L;
    for (int64_t i = 0; i < n; i++) s += x[i];
    s /= n;
    if (isfinite((double)s)) {
        long double t = 0.0L;
        for (int64_t i = 0; i < n; i++) t += (x[i] - s);
        s += t / n;
    }
    return (double)s;
}

/* mean of a strided row: apply(M[, cols], 1, mean) on a column-major matrix. */
double hbo_r_mean_gather(const double *m, int64_t stride, const int32_t *cols, int64_t ncols)
{
    long double s = 0.0L;
    for (int64_t i = 0; i < ncols; i++) s += m[(int64_t)cols[i] * stride];
    s /= ncols;
    if (isfinite((double)s)) {
        long double t = 0.0L;
        for (int64_t i = 0; i < ncols; i++) t += (m[(int64_t)cols[i] * stride] - s);
        s += t / ncols;
    }
    return (double)s;
}

/* sd(x) = sqrt(var(x)) (cov.c cov_complete1: LDOUBLE mean rounded to double, LDOUBLE SSQ, n-1). */
double hbo_r_sd(const double *x, int64_t n)
{
    if (n < 2) return NAN;
    double xm = hbo_r_mean(x, n);
    long double xxm = xm, sum = 0.0L;
    for (int64_t k = 0; k < n; k++) sum += (x[k] - xxm) * (x[k] - xxm);
    double var = (double)(sum / (long double)(n - 1));
    return sqrt(var);
}

static double r_sum(const double *x, int n)
{
    long double s = 0.0L; /* summary.c rsum: LDOUBLE accumulator */
    for (int i = 0; i < n; i++) s += x[i];
    return (double)s;
}

/* -------------------------------------------- HiddenMarkov::Viterbi.dthmm */

/* First index of the maximum, NaN skipped — R's which.max. Returns -1 if all NaN. */
static int which_max(const double *v, int m)
{
    int best = -1;
    double s = 0;
    for (int i = 0; i < m; i++) {
        if (isnan(v[i])) continue;
        if (best < 0 || v[i] > s) {
            s = v[i];
            best = i;
        }
    }
    return best;
}

/*
 * Generic core on a precomputed log-density table logb[T][S] (row-major).
 * Pi is row-major Pi[j*S+k] = P(j -> k).  y is 1-based as in R.  nu_out (optional) is [T][S].
 */
int hbo_viterbi_logb(int S, int64_t T, const double *logb, const double *Pi, const double *delta,
                     int32_t *y, double *nu_out)
{
    if (S < 1 || S > HBO_MAX_S || T < 1) return HBO_ERR_ARG;
    double logPi[HBO_MAX_S * HBO_MAX_S], tmp[HBO_MAX_S];
    for (int i = 0; i < S * S; i++) logPi[i] = log(Pi[i]);
    double *nu = nu_out ? nu_out : (double *)malloc(sizeof(double) * (size_t)T * S);
    if (!nu) return HBO_ERR_NOMEM;
    for (int k = 0; k < S; k++) nu[k] = log(delta[k]) + logb[k];
    for (int64_t i = 1; i < T; i++) {
        const double *prev = nu + (i - 1) * S;
        double *cur = nu + i * S;
        for (int k = 0; k < S; k++) {
            /* apply(matrixnu + logPi, 2, max): plain max over j of nu[j] + logPi[j,k] */
            double mx = prev[0] + logPi[0 * S + k];
            for (int j = 1; j < S; j++) {
                double c = prev[j] + logPi[j * S + k];
                if (c > mx) mx = c; /* NaN never arises on finite inputs */
            }
            cur[k] = mx + logb[i * S + k];
        }
    }
    int rc = HBO_OK;
    for (int k = 0; k < S; k++)
        if (nu[(T - 1) * S + k] == -INFINITY) rc = HBO_ERR_UNDERFLOW;
    if (rc == HBO_OK) {
        y[T - 1] = which_max(nu + (T - 1) * S, S) + 1;
        for (int64_t i = T - 2; i >= 0; i--) {
            int nxt = y[i + 1] - 1;
            for (int j = 0; j < S; j++) tmp[j] = logPi[j * S + nxt] + nu[i * S + j];
            y[i] = which_max(tmp, S) + 1;
        }
    }
    if (!nu_out) free(nu);
    return rc;
}

/* dthmm(x, Pi, delta, "norm", list(mean=, sd=)) -> Viterbi */
int hbo_viterbi_norm(int S, int64_t T, const double *x, const double *mean, const double *sd,
                     const double *Pi, const double *delta, int32_t *y)
{
    if (S < 1 || S > HBO_MAX_S || T < 1) return HBO_ERR_ARG;
    double *logb = (double *)malloc(sizeof(double) * (size_t)T * S);
    if (!logb) return HBO_ERR_NOMEM;
    for (int64_t i = 0; i < T; i++)
        for (int k = 0; k < S; k++) logb[i * S + k] = hbo_dnorm_log(x[i], mean[k], sd[k]);
    int rc = hbo_viterbi_logb(S, T, logb, Pi, delta, y, NULL);
    free(logb);
    return rc;
}

/* Beta-binomial log-pmf, mean p, intra-class correlation rho in [0, 1) (alpha = p(1-rho)/rho,
 * beta = (1-p)(1-rho)/rho).  NOT in the reference — HoneyBADGER's allele chain is binomial
 * (R/HoneyBADGER.R:1409, R/HoneyBADGER_allele.R:579); north_star names it as an extension, so there is no
 * reference arithmetic to restate and THIS is the definition the product's emission table follows
 * (honeybadger_b200/csrc/hb_host_math.h builds the same sums as prefix arrays):
 *     th   = rho / (1 - rho)                                              (long double)
 *     corr = [ sum_{i=0}^{x-1} log1pl(i th / p) + sum_{i=0}^{n-x-1} log1pl(i th / q) ] - sum_{i=0}^{n-1} log1pl(i th)
 *            each sum accumulated from 0 in increasing i, long double, q = 1 - p rounded to double first
 *     log BB(x; n, p, rho) = dbinom(x, n, p, log = TRUE) + (double) corr
 * i.e. the identity BB = C(n,x) prod(p + i th) prod(q + i th) / prod(1 + i th) written relative to the binomial,
 * so that rho = 0 IS the reference's binomial and rho -> 0 converges without cancellation.
 * tests/test_oracle.py pins it on mpmath's exact log-pmf (oracle/pyref.py: betabinom_logpmf_exact). */
double hbo_dbetabinom_log(double x, double n, double p, double rho)
{
    if (x < 0 || x > n) return -INFINITY;
    const double base = hbo_dbinom(x, n, p, 1);
    const double q = 1 - p;
    if (!(rho > 0) || p == 0 || q == 0) return base;
    const long double th = (long double)rho / (1.0L - (long double)rho);
    long double sa = 0.0L, sb = 0.0L, sd = 0.0L;
    const int64_t xi = (int64_t)x, ni = (int64_t)n;
    for (int64_t i = 0; i < xi; i++) sa += log1pl((long double)i * th / (long double)p);
    for (int64_t i = 0; i < ni - xi; i++) sb += log1pl((long double)i * th / (long double)q);
    for (int64_t i = 0; i < ni; i++) sd += log1pl((long double)i * th);
    return base + (double)((sa + sb) - sd);
}

static double binom_like_log(double x, double n, double p, double rho)
{
    return rho > 0 ? hbo_dbetabinom_log(x, n, p, rho) : hbo_dbinom(x, n, p, 1);
}

/* dthmm(x, Pi, delta, "binom", list(prob=), list(size=), discrete=TRUE) -> Viterbi; rho > 0: beta-binomial */
int hbo_viterbi_binom_rho(int S, int64_t T, const int32_t *x, const int32_t *size, const double *prob,
                          double rho, const double *Pi, const double *delta, int32_t *y)
{
    if (S < 1 || S > HBO_MAX_S || T < 1) return HBO_ERR_ARG;
    double *logb = (double *)malloc(sizeof(double) * (size_t)T * S);
    if (!logb) return HBO_ERR_NOMEM;
    for (int64_t i = 0; i < T; i++)
        for (int k = 0; k < S; k++)
            logb[i * S + k] = binom_like_log((double)x[i], (double)size[i], prob[k], rho);
    int rc = hbo_viterbi_logb(S, T, logb, Pi, delta, y, NULL);
    free(logb);
    return rc;
}

int hbo_viterbi_binom(int S, int64_t T, const int32_t *x, const int32_t *size, const double *prob,
                      const double *Pi, const double *delta, int32_t *y)
{
    return hbo_viterbi_binom_rho(S, T, x, size, prob, 0.0, Pi, delta, y);
}

/* ---------------------------------- HiddenMarkov::forwardback.dthmm + Estep u */

/*
 * prob[T][S] are LINEAR densities.  Outputs (any may be NULL): logalpha[T][S], logbeta[T][S],
 * gamma[T][S] = exp(logalpha + logbeta - LL), *LL.
 */
int hbo_forwardback_prob(int S, int64_t T, const double *prob, const double *Pi,
                         const double *delta, double *logalpha, double *logbeta, double *gamma,
                         double *LL)
{
    if (S < 1 || S > HBO_MAX_S || T < 1) return HBO_ERR_ARG;
    double *la = logalpha ? logalpha : (double *)malloc(sizeof(double) * (size_t)T * S);
    double *lb = logbeta ? logbeta : (double *)malloc(sizeof(double) * (size_t)T * S);
    if (!la || !lb) return HBO_ERR_NOMEM;
    double phi[HBO_MAX_S], nxt[HBO_MAX_S];
    double lscale = 0.0;
    for (int k = 0; k < S; k++) phi[k] = delta[k];
    for (int64_t i = 0; i < T; i++) {
        if (i > 0) { /* phi %*% Pi */
            for (int k = 0; k < S; k++) {
                double a = 0.0;
                for (int j = 0; j < S; j++) a += phi[j] * Pi[j * S + k];
                nxt[k] = a;
            }
            memcpy(phi, nxt, sizeof(double) * S);
        }
        for (int k = 0; k < S; k++) phi[k] = phi[k] * prob[i * S + k];
        double sumphi = r_sum(phi, S);
        for (int k = 0; k < S; k++) phi[k] = phi[k] / sumphi;
        lscale = lscale + log(sumphi);
        for (int k = 0; k < S; k++) la[i * S + k] = log(phi[k]) + lscale;
    }
    double ll = lscale;
    for (int k = 0; k < S; k++) {
        phi[k] = 1.0 / S;
        lb[(T - 1) * S + k] = 0.0;
    }
    lscale = log((double)S);
    for (int64_t i = T - 2; i >= 0; i--) {
        double w[HBO_MAX_S];
        for (int k = 0; k < S; k++) w[k] = prob[(i + 1) * S + k] * phi[k];
        for (int j = 0; j < S; j++) { /* Pi %*% w */
            double a = 0.0;
            for (int k = 0; k < S; k++) a += Pi[j * S + k] * w[k];
            phi[j] = a;
        }
        for (int k = 0; k < S; k++) lb[i * S + k] = log(phi[k]) + lscale;
        double sumphi = r_sum(phi, S);
        for (int k = 0; k < S; k++) phi[k] = phi[k] / sumphi;
        lscale = lscale + log(sumphi);
    }
    if (gamma)
        for (int64_t i = 0; i < T * S; i++) gamma[i] = exp(la[i] + lb[i] - ll);
    if (LL) *LL = ll;
    if (!logalpha) free(la);
    if (!logbeta) free(lb);
    return HBO_OK;
}

int hbo_forwardback_norm(int S, int64_t T, const double *x, const double *mean, const double *sd,
                         const double *Pi, const double *delta, double *gamma, double *LL)
{
    double *prob = (double *)malloc(sizeof(double) * (size_t)T * S);
    if (!prob) return HBO_ERR_NOMEM;
    for (int64_t i = 0; i < T; i++)
        for (int k = 0; k < S; k++) prob[i * S + k] = hbo_dnorm(x[i], mean[k], sd[k]);
    int rc = hbo_forwardback_prob(S, T, prob, Pi, delta, NULL, NULL, gamma, LL);
    free(prob);
    return rc;
}

int hbo_forwardback_binom_rho(int S, int64_t T, const int32_t *x, const int32_t *size,
                              const double *prob_k, double rho, const double *Pi, const double *delta,
                              double *gamma, double *LL)
{
    double *prob = (double *)malloc(sizeof(double) * (size_t)T * S);
    if (!prob) return HBO_ERR_NOMEM;
    for (int64_t i = 0; i < T; i++)
        for (int k = 0; k < S; k++)
            prob[i * S + k] = rho > 0 ? exp(hbo_dbetabinom_log((double)x[i], (double)size[i], prob_k[k], rho))
                                      : hbo_dbinom((double)x[i], (double)size[i], prob_k[k], 0);
    int rc = hbo_forwardback_prob(S, T, prob, Pi, delta, NULL, NULL, gamma, LL);
    free(prob);
    return rc;
}

int hbo_forwardback_binom(int S, int64_t T, const int32_t *x, const int32_t *size,
                          const double *prob_k, const double *Pi, const double *delta,
                          double *gamma, double *LL)
{
    return hbo_forwardback_binom_rho(S, T, x, size, prob_k, 0.0, Pi, delta, gamma, LL);
}

/* ------------------------------------------------------------------ batched drivers
 * Gene-major batches (x[t*ld + b], the layout the CUDA path uses) of independent chains, one chain
 * per (column b, segment s) with segments seg_off[0..nseg] on the gene axis.  Used by the parity
 * tests and timed by bench.py as the CPU baseline (OpenMP over chains when built with -fopenmp).
 * Outputs: y[t*ldy + b] uint8 1-based; gamma planar [S][T][ldg] (may be NULL); ll[s*B + b].
 * Returns HBO_OK or the first error (HBO_ERR_UNDERFLOW marks the chain, others still run).
 */
int hbo_fbv_norm_batch(int64_t T, int64_t B, int64_t ld, const double *x, const int64_t *seg_off,
                       int nseg, const double *mean /*[3]*/, const double *sd /*[nseg*B] or [B]*/,
                       int sd_per_seg, const double *Pi, const double *delta, uint8_t *y,
                       double *gamma, double *ll, int nthreads)
{
    const int S = 3;
    int rc_all = HBO_OK;
    int64_t nchain = B * nseg;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int64_t c = 0; c < nchain; c++) {
        int s = (int)(c / B);
        int64_t b = c % B;
        int64_t t0 = seg_off[s], L = seg_off[s + 1] - seg_off[s];
        if (L <= 0) continue;
        double *xs = (double *)malloc(sizeof(double) * L);
        int32_t *ys = (int32_t *)malloc(sizeof(int32_t) * L);
        double *gs = gamma ? (double *)malloc(sizeof(double) * L * S) : NULL;
        for (int64_t i = 0; i < L; i++) xs[i] = x[(t0 + i) * ld + b];
        double sdv = sd[sd_per_seg ? s * B + b : b];
        double sds[3] = {sdv, sdv, sdv};
        int rc = hbo_viterbi_norm(S, L, xs, mean, sds, Pi, delta, ys);
        for (int64_t i = 0; i < L; i++) y[(t0 + i) * ld + b] = rc == HBO_OK ? (uint8_t)ys[i] : 0;
        if (gamma || ll) {
            double llv = 0;
            hbo_forwardback_norm(S, L, xs, mean, sds, Pi, delta, gs, &llv);
            if (ll) ll[s * B + b] = llv;
            if (gamma)
                for (int64_t i = 0; i < L; i++)
                    for (int k = 0; k < S; k++)
                        gamma[((int64_t)k * T + t0 + i) * ld + b] = gs[i * S + k];
        }
        if (rc != HBO_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
            rc_all = rc;
        }
        free(xs);
        free(ys);
        free(gs);
    }
    return rc_all;
}

int hbo_fbv_binom_batch_rho(int64_t T, int64_t B, int64_t ld, const int32_t *x, const int32_t *n,
                            const int64_t *seg_off, int nseg, const double *prob /*[2]*/, double rho,
                            const double *Pi, const double *delta, uint8_t *y, double *gamma,
                            double *ll, int nthreads);

int hbo_fbv_binom_batch(int64_t T, int64_t B, int64_t ld, const int32_t *x, const int32_t *n,
                        const int64_t *seg_off, int nseg, const double *prob /*[2]*/,
                        const double *Pi, const double *delta, uint8_t *y, double *gamma,
                        double *ll, int nthreads)
{
    return hbo_fbv_binom_batch_rho(T, B, ld, x, n, seg_off, nseg, prob, 0.0, Pi, delta, y, gamma, ll, nthreads);
}

int hbo_fbv_binom_batch_rho(int64_t T, int64_t B, int64_t ld, const int32_t *x, const int32_t *n,
                            const int64_t *seg_off, int nseg, const double *prob /*[2]*/, double rho,
                            const double *Pi, const double *delta, uint8_t *y, double *gamma,
                            double *ll, int nthreads)
{
    const int S = 2;
    int rc_all = HBO_OK;
    int64_t nchain = B * nseg;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int64_t c = 0; c < nchain; c++) {
        int s = (int)(c / B);
        int64_t b = c % B;
        int64_t t0 = seg_off[s], L = seg_off[s + 1] - seg_off[s];
        if (L <= 0) continue;
        int32_t *xs = (int32_t *)malloc(sizeof(int32_t) * L);
        int32_t *ns = (int32_t *)malloc(sizeof(int32_t) * L);
        int32_t *ys = (int32_t *)malloc(sizeof(int32_t) * L);
        double *gs = gamma ? (double *)malloc(sizeof(double) * L * S) : NULL;
        for (int64_t i = 0; i < L; i++) {
            xs[i] = x[(t0 + i) * ld + b];
            ns[i] = n[(t0 + i) * ld + b];
        }
        int rc = hbo_viterbi_binom_rho(S, L, xs, ns, prob, rho, Pi, delta, ys);
        for (int64_t i = 0; i < L; i++) y[(t0 + i) * ld + b] = rc == HBO_OK ? (uint8_t)ys[i] : 0;
        if (gamma || ll) {
            double llv = 0;
            hbo_forwardback_binom_rho(S, L, xs, ns, prob, rho, Pi, delta, gs, &llv);
            if (ll) ll[s * B + b] = llv;
            if (gamma)
                for (int64_t i = 0; i < L; i++)
                    for (int k = 0; k < S; k++)
                        gamma[((int64_t)k * T + t0 + i) * ld + b] = gs[i * S + k];
        }
        if (rc != HBO_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
            rc_all = rc;
        }
        free(xs);
        free(ns);
        free(ys);
        free(gs);
    }
    return rc_all;
}

/* Pooled statistics of one cell group (R/HoneyBADGER.R:1141, :1404-1405) on a COLUMN-MAJOR
 * [G x N] matrix, cells given as 0-based column indices in increasing order. */
void hbo_pool_mean(const double *m, int64_t G, const int32_t *cols, int64_t ncols, double *out)
{
    for (int64_t g = 0; g < G; g++) out[g] = hbo_r_mean_gather(m + g, G, cols, ncols);
}

void hbo_pool_count_pos(const double *m, int64_t G, const int32_t *cols, int64_t ncols,
                        int32_t *out)
{
    for (int64_t g = 0; g < G; g++) {
        int32_t c = 0;
        for (int64_t i = 0; i < ncols; i++) c += m[(int64_t)cols[i] * G + g] > 0;
        out[g] = c;
    }
}

/* ------------------------------------------------------------ input construction
 * setGexpMats (R/HoneyBADGER.R:124-188) and setAlleleMats (R/HoneyBADGER.R:486-614), arithmetic part,
 * on COLUMN-MAJOR matrices exactly as R evaluates them. */

/* rowMeans(m) (array.c do_colsum, OP 3): one LDOUBLE accumulator per row, columns in order, / p. */
void hbo_row_means(const double *m, int64_t G, int64_t N, double *out)
{
    for (int64_t g = 0; g < G; g++) {
        long double s = 0.0L;
        for (int64_t c = 0; c < N; c++) s += m[c * G + g];
        s /= N;
        out[g] = (double)s;
    }
}

/* scale(m) in place (scale.default): center <- colMeans(m, na.rm=TRUE); m <- sweep(m, 2, center);
 * f(v) = sqrt(sum(v^2) / max(1, length(v) - 1)); m <- sweep(m, 2, apply(m, 2, f), "/"). */
void hbo_scale(double *m, int64_t G, int64_t N)
{
    for (int64_t c = 0; c < N; c++) {
        double *v = m + c * G;
        long double s = 0.0L;
        for (int64_t g = 0; g < G; g++) s += v[g];
        s /= G;
        const double center = (double)s;
        for (int64_t g = 0; g < G; g++) v[g] = v[g] - center;
        long double ss = 0.0L; /* sum(): rsum accumulates the double vector v^2 in LDOUBLE */
        for (int64_t g = 0; g < G; g++) {
            const double sq = v[g] * v[g]; /* R_pow(x, 2.0) = x * x */
            ss += sq;
        }
        const double den = (double)(G - 1 > 1 ? G - 1 : 1);
        const double scl = sqrt((double)ss / den);
        for (int64_t g = 0; g < G; g++) v[g] = v[g] / scl;
    }
}

/* mean(m[m != 0]): the subset vector in column-major order, then real_mean. */
double hbo_mean_nonzero(const double *m, int64_t count)
{
    double *nz = (double *)malloc((size_t)(count > 0 ? count : 1) * sizeof(double));
    int64_t n = 0;
    for (int64_t i = 0; i < count; i++)
        if (m[i] != 0.0) nz[n++] = m[i];
    double r = n ? hbo_r_mean(nz, n) : NAN;
    free(nz);
    return r;
}

/* Returns the number of kept rows Gk; norm is the Gk x N column-major result, keep_rows their indices. */
int64_t hbo_set_gexp_mats(const double *sc, int64_t G, int64_t N, const double *ref, int64_t R, int filter,
                          double min_both, const double *min_test, const double *min_ref, int scale,
                          double *norm, int32_t *keep_rows)
{
    int64_t Gk = 0;
    if (filter) {
        const double mt = min_test ? *min_test : hbo_mean_nonzero(sc, G * N);
        const double mr = min_ref ? *min_ref : hbo_mean_nonzero(ref, G * R);
        double *rs = (double *)malloc((size_t)G * 2 * sizeof(double)), *rr = rs + G;
        hbo_row_means(sc, G, N, rs);
        hbo_row_means(ref, G, R, rr);
        for (int64_t g = 0; g < G; g++)
            if ((rs[g] > min_both && rr[g] > min_both) || rs[g] > mt || rr[g] > mr) keep_rows[Gk++] = (int32_t)g;
        free(rs);
    } else {
        for (int64_t g = 0; g < G; g++) keep_rows[Gk++] = (int32_t)g;
    }
    if (Gk == 0) return 0;
    double *r2 = (double *)malloc((size_t)Gk * R * sizeof(double));
    double *rm = (double *)malloc((size_t)Gk * sizeof(double));
    for (int64_t c = 0; c < N; c++)
        for (int64_t k = 0; k < Gk; k++) norm[c * Gk + k] = sc[c * G + keep_rows[k]];
    for (int64_t c = 0; c < R; c++)
        for (int64_t k = 0; k < Gk; k++) r2[c * Gk + k] = ref[c * G + keep_rows[k]];
    if (scale) {
        hbo_scale(norm, Gk, N);
        hbo_scale(r2, Gk, R);
    }
    hbo_row_means(r2, Gk, R, rm);
    for (int64_t c = 0; c < N; c++)
        for (int64_t k = 0; k < Gk; k++) norm[c * Gk + k] = norm[c * Gk + k] - rm[k];
    free(r2);
    free(rm);
    return Gk;
}

/* setAlleleMats on column-major integer-valued matrices.  l_init / n_bulk_init NULL = in-silico bulk. */
int64_t hbo_set_allele_mats(const int32_t *r, const int32_t *n, int64_t G, int64_t N, const double *l_init,
                            const double *n_bulk_init, int filter, double thr, int min_cell, int32_t *r_maf,
                            int32_t *n_sc, double *l_out, double *n_bulk_out, double *l_maf,
                            int32_t *keep_rows)
{
    int64_t Gk = 0;
    const int silico = !l_init || !n_bulk_init;
    for (int64_t g = 0; g < G; g++) {
        int64_t cr = 0, cn = 0;
        for (int64_t c = 0; c < N; c++) {
            cr += r[c * G + g] > 0;
            cn += n[c * G + g] > 0;
        }
        const double l = silico ? (double)cr : l_init[g], nb = silico ? (double)cn : n_bulk_init[g];
        const double E = l / nb;
        if (filter) {
            if (!(E > thr && E < 1 - thr)) continue; /* which() drops NA */
            if (cn < min_cell) continue;
        }
        keep_rows[Gk] = (int32_t)g;
        l_out[Gk] = l;
        n_bulk_out[Gk] = nb;
        l_maf[Gk] = isnan(E) ? 0.0 : (E <= 0.5 ? l : nb - l);
        Gk++;
    }
    for (int64_t c = 0; c < N; c++)
        for (int64_t k = 0; k < Gk; k++) {
            const int64_t g = keep_rows[k];
            const double E = l_out[k] / n_bulk_out[k];
            const int32_t rv = r[c * G + g], nv = n[c * G + g];
            r_maf[c * Gk + k] = isnan(E) ? 0 : (E <= 0.5 ? rv : nv - rv);
            n_sc[c * Gk + k] = nv;
        }
    return Gk;
}

int hbo_num_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
