// hb_host_math.h — host-side parameter preparation for the kernels (product code, not the oracle).
//
// The binomial emission table and the logs of the chain constants are evaluated on the host with
// the host libm so that they carry exactly the values R's nmath (dbinom_raw / stirlerr / bd0,
// SURVEY.md A.3) produces on the same machine; the kernels then only add and compare them.
// Compile without FMA contraction (x86-64 default; nvcc passes -ffp-contract=off to the host
// compiler through -Xcompiler in build.py).
#pragma once
#include <float.h>
#include <math.h>

#include <vector>

namespace hbhost {

static const double kSferrHalves[31] = {
    0.0, /* placeholder for n = 0 */
    0.1534264097200273452913848,   0.0810614667953272582196702,   0.0548141210519176538961390,
    0.0413406959554092940938221,   0.03316287351993628748511048,  0.02767792568499833914878929,
    0.02374616365629749597132920,  0.02079067210376509311152277,  0.01848845053267318523077934,
    0.01664469118982119216319487,  0.01513497322191737887351255,  0.01387612882307074799874573,
    0.01281046524292022692424986,  0.01189670994589177009505572,  0.01110455975820691732662991,
    0.010411265261972096497478567, 0.009799416126158803298389475, 0.009255462182712732917728637,
    0.008768700134139385462952823, 0.008330563433362871256469318, 0.007934114564314020547248100,
    0.007573675487951840794972024, 0.007244554301320383179543912, 0.006942840107209529865664152,
    0.006665247032707682442354394, 0.006408994188004207068439631, 0.006171712263039457647532867,
    0.005951370112758847735624416, 0.005746216513010115682023589, 0.005554733551962801371038690};

inline double stirling_error(double n)
{
    const double c0 = 0.083333333333333333333, c1 = 0.00277777777777777777778,
                 c2 = 0.00079365079365079365079365, c3 = 0.000595238095238095238095238,
                 c4 = 0.0008417508417508417508417508;
    if (n <= 15.0) {
        double twice = n + n;
        if (twice == (int)twice) return kSferrHalves[(int)twice];
        return lgamma(n + 1.) - (n + 0.5) * log(n) + n - 0.918938533204672741780329736406;
    }
    double sq = n * n;
    if (n > 500) return (c0 - c1 / sq) / n;
    if (n > 80) return (c0 - (c1 - c2 / sq) / sq) / n;
    if (n > 35) return (c0 - (c1 - (c2 - c3 / sq) / sq) / sq) / n;
    return (c0 - (c1 - (c2 - (c3 - c4 / sq) / sq) / sq) / sq) / n;
}

inline double deviance_term(double x, double np)
{
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < DBL_MIN) return s;
        double ej = 2 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            double s1 = s + ej / ((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

// log dbinom(x; n, p) for integers 0 <= x <= n (Loader's saddle-point form, as in nmath).
inline double binom_logpmf(double x, double n, double p)
{
    const double q = 1 - p;
    if (p == 0) return x == 0 ? 0. : -INFINITY;
    if (q == 0) return x == n ? 0. : -INFINITY;
    if (x == 0) {
        if (n == 0) return 0.;
        return (p < 0.1) ? -deviance_term(n, n * q) - n * p : n * log(q);
    }
    if (x == n) return (q < 0.1) ? -deviance_term(n, n * p) - n * q : n * log(p);
    if (x < 0 || x > n) return -INFINITY;
    double lc = stirling_error(n) - stirling_error(x) - stirling_error(n - x) -
                deviance_term(x, n * p) - deviance_term(n - x, n * q);
    double lf = 1.837877066409345483560659472811 + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

// Beta-binomial log-pmf with mean p and intra-class correlation rho in [0, 1): alpha = p (1-rho)/rho,
// beta = (1-p)(1-rho)/rho.  An EXTENSION over the reference, which is binomial (R/HoneyBADGER.R:1409,
// R/HoneyBADGER_allele.R:579); it reduces to binom_logpmf EXACTLY at rho = 0 and smoothly
// as rho -> 0, because it is evaluated as the binomial value plus a correction:
//     BB(x; n, p, rho) = C(n,x) prod_{i<x}(p + i th) prod_{i<n-x}(q + i th) / prod_{i<n}(1 + i th),  th = rho/(1-rho)
//     log BB = log dbinom(x; n, p) + [ sum_{i<x} log1p(i th / p) + sum_{i<n-x} log1p(i th / q) - sum_{i<n} log1p(i th) ]
// (no lgamma of arguments ~ p/rho, whose cancellation destroys the limit).  The three sums are accumulated
// in long double in increasing i; BetaBinomCorr keeps them as prefix arrays so that a whole emission table
// costs O(nmax) log1p evaluations.  The CPU checker's hbo_dbetabinom_log (under oracle/) follows the same
// definition with direct loops and must agree bit for bit; oracle/pyref.py checks both against mpmath's exact value.
struct BetaBinomCorr {
    std::vector<long double> la, lb, ld; // la[k] = sum_{i<k} log1pl(i th / p), lb: / q, ld: / 1
    bool degenerate = true;
    BetaBinomCorr(double p, double rho, int nmax)
    {
        const double q = 1 - p;
        if (!(rho > 0) || p == 0 || q == 0) return;
        degenerate = false;
        const long double th = (long double)rho / (1.0L - (long double)rho);
        la.assign((size_t)nmax + 2, 0.0L);
        lb.assign((size_t)nmax + 2, 0.0L);
        ld.assign((size_t)nmax + 2, 0.0L);
        for (int k = 0; k <= nmax; k++) {
            const long double it = (long double)k * th;
            la[(size_t)k + 1] = la[k] + log1pl(it / (long double)p);
            lb[(size_t)k + 1] = lb[k] + log1pl(it / (long double)q);
            ld[(size_t)k + 1] = ld[k] + log1pl(it);
        }
    }
    double corr(int x, int n) const
    {
        if (degenerate) return 0.0;
        return (double)((la[x] + lb[n - x]) - ld[n]);
    }
};

inline double betabinom_logpmf(const BetaBinomCorr &c, int x, int n, double p)
{
    if (x < 0 || x > n) return -INFINITY;
    const double b = binom_logpmf((double)x, (double)n, p);
    return c.degenerate ? b : b + c.corr(x, n);
}

} // namespace hbhost
