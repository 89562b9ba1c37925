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

// Beta-binomial log-pmf with mean p and intra-class correlation rho in (0, 1):
// alpha = p (1-rho)/rho, beta = (1-p)(1-rho)/rho.  Extension over the reference;
// tends to binom_logpmf as rho -> 0.
inline double betabinom_logpmf(double x, double n, double p, double rho)
{
    if (x < 0 || x > n) return -INFINITY;
    const double th = (1 - rho) / rho, al = p * th, be = (1 - p) * th;
    return lgamma(n + 1) - lgamma(x + 1) - lgamma(n - x + 1) + lgamma(x + al) + lgamma(n - x + be) -
           lgamma(n + al + be) + lgamma(al + be) - lgamma(al) - lgamma(be);
}

} // namespace hbhost
