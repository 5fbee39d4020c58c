/*
 * oracle/pal_numerics.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the third-party numerics the reference's hot path calls into:
 * PAL 1.51 (`/root/reference/lib/pal.jar`, bytecode only, pinned by sha256 6eebc187…a92f1d5;
 * no source in the reference tree). Each function names the PAL method it follows
 * (`Class::method`, as read from the bytecode; SURVEY.md Appendix B has the pc ranges).
 *
 * Pinning: `tests/test_oracle_pal_golden.py` compares every function here against golden
 * vectors in `tests/golden/pal_golden.json`, which were produced by EXECUTING the jar's own
 * bytecode in `oracle/jvm/interp.py` (script: `tests/golden/make_pal_golden.py`).
 *
 * Build with -O2 -ffp-contract=off (no FMA contraction: Java evaluates a*b+c with two roundings).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define NS 20

/* AbstractRateMatrix::setFrequencies -> checkFrequencies.
 * Mutates f[20] in place: floor at 1e-10, residual into the first maximum, break exact ties. */
void pal_check_frequencies(double *f)
{
    const double eps = 1e-10;
    double sum = 0.0, maxfreq = 0.0;
    int maxi = 0;
    for (int i = 0; i < NS; ++i) {
        double v = f[i];
        if (v < eps) f[i] = eps;
        if (v > maxfreq) { maxfreq = v; maxi = i; }
        sum += f[i];
    }
    f[maxi] += 1.0 - sum;
    for (int i = 0; i < NS - 1; ++i)
        for (int j = i + 1; j < NS; ++j)
            if (f[i] == f[j]) { f[i] += eps; f[j] -= eps; }
}

/* <Model>::rebuildRateMatrix (fills the strict upper triangle from the table) ->
 * AbstractRateMatrix::fromQToR -> makeValid -> normalize. R is row-major 20x20. */
void pal_rate_matrix(const double *q_upper /*190, rows i<j*/, const double *freq, double *R)
{
    double (*r)[NS] = (double (*)[NS])R;
    memset(R, 0, sizeof(double) * NS * NS);
    int p = 0;
    for (int i = 0; i < NS; ++i)
        for (int j = i + 1; j < NS; ++j) r[i][j] = q_upper[p++];
    /* fromQToR */
    for (int i = 0; i < NS; ++i)
        for (int j = i + 1; j < NS; ++j) {
            double q = r[i][j];
            r[i][j] = q * freq[j];
            r[j][i] = q * freq[i];
        }
    /* makeValid: diagonal = -(sum of off-diagonals), j ascending */
    for (int i = 0; i < NS; ++i) {
        double s = 0.0;
        for (int j = 0; j < NS; ++j)
            if (j != i) s += r[i][j];
        r[i][i] = -s;
    }
    /* normalize: mean rate 1 */
    double subst = 0.0;
    for (int i = 0; i < NS; ++i) subst += -r[i][i] * freq[i];
    for (int i = 0; i < NS; ++i)
        for (int j = 0; j < NS; ++j) r[i][j] = r[i][j] / subst;
}

/* MatrixExponential::elmhes — EISPACK ELMHES (stabilised elementary similarity to Hessenberg). */
static void elmhes(double a[NS][NS], int *ordr, int n)
{
    for (int i = 0; i < n; ++i) ordr[i] = 0;
    for (int m = 1; m < n - 1; ++m) {
        double x = 0.0;
        int i = m;
        for (int j = m; j < n; ++j)
            if (fabs(a[j][m - 1]) > fabs(x)) { x = a[j][m - 1]; i = j; }
        ordr[m] = i;
        if (i != m) {
            for (int j = m - 1; j < n; ++j) { double y = a[i][j]; a[i][j] = a[m][j]; a[m][j] = y; }
            for (int j = 0; j < n; ++j)     { double y = a[j][i]; a[j][i] = a[j][m]; a[j][m] = y; }
        }
        if (x != 0.0) {
            for (i = m + 1; i < n; ++i) {
                double y = a[i][m - 1];
                if (y != 0.0) {
                    y /= x;
                    a[i][m - 1] = y;
                    for (int j = m; j < n; ++j) a[i][j] -= y * a[m][j];
                    for (int j = 0; j < n; ++j) a[j][m] += y * a[j][i];
                }
            }
        }
    }
}

/* MatrixExponential::eltran — EISPACK ELTRAN (accumulate the ELMHES transformations). */
static void eltran(double a[NS][NS], double zz[NS][NS], const int *ordr, int n)
{
    for (int i = 0; i < n; ++i) {
        for (int j = i + 1; j < n; ++j) { zz[i][j] = 0.0; zz[j][i] = 0.0; }
        zz[i][i] = 1.0;
    }
    if (n <= 2) return;
    for (int m = n - 1; m >= 2; --m) {
        for (int i = m; i < n; ++i) zz[i][m - 1] = a[i][m - 2];
        int i = ordr[m - 1];
        if (i != m - 1) {
            for (int j = m - 1; j < n; ++j) { zz[m - 1][j] = zz[i][j]; zz[i][j] = 0.0; }
            zz[i][m - 1] = 1.0;
        }
    }
}

/* MatrixExponential::hqr2 — EISPACK HQR2 (QR on the Hessenberg form, eigenvectors by back
 * substitution), real-eigenvalue path; complex pairs cannot arise for a reversible model and
 * return an error. 30-iteration cap and exceptional shifts at 10 and 20 as in PAL. */
static int hqr2(int n, double h[NS][NS], double zz[NS][NS], double *wr, double *wi)
{
    const int low = 0, hgh = n - 1;
    double norm = 0.0, p = 0, q = 0, r = 0, s = 0, t, w, x, y, z = 0;
    int k = 0;
    for (int i = 0; i < n; ++i) {
        for (int j = k; j < n; ++j) norm += fabs(h[i][j]);
        k = i;
    }
    int en = hgh;
    t = 0.0;
    while (en >= low) {
        int its = 0, na = en - 1, l, m;
        for (;;) {
            /* look for single small sub-diagonal element */
            for (l = en; l > low; --l) {
                s = fabs(h[l - 1][l - 1]) + fabs(h[l][l]);
                if (s == 0.0) s = norm;
                if ((fabs(h[l][l - 1]) + s) == s) break;
            }
            x = h[en][en];
            if (l == en) break;                       /* one root found */
            y = h[na][na];
            w = h[en][na] * h[na][en];
            if (l == na) break;                       /* two roots found */
            if (its == 30) return -1;                 /* "Eigenvalues not converged" */
            if (its == 10 || its == 20) {             /* exceptional shift */
                t += x;
                for (int i = low; i <= en; ++i) h[i][i] -= x;
                s = fabs(h[en][na]) + fabs(h[na][en - 2]);
                x = 0.75 * s;
                y = x;
                w = -0.4375 * s * s;
            }
            ++its;
            /* look for two consecutive small sub-diagonal elements */
            for (m = en - 2; m >= l; --m) {
                z = h[m][m];
                r = x - z;
                s = y - z;
                p = (r * s - w) / h[m + 1][m] + h[m][m + 1];
                q = h[m + 1][m + 1] - z - r - s;
                r = h[m + 2][m + 1];
                s = fabs(p) + fabs(q) + fabs(r);
                p /= s; q /= s; r /= s;
                if (m == l) break;
                double u = fabs(h[m][m - 1]) * (fabs(q) + fabs(r));
                double v = fabs(p) * (fabs(h[m - 1][m - 1]) + fabs(z) + fabs(h[m + 1][m + 1]));
                if ((u + v) == v) break;
            }
            for (int i = m + 2; i <= en; ++i) {
                h[i][i - 2] = 0.0;
                if (i != m + 2) h[i][i - 3] = 0.0;
            }
            /* double QR step on rows l..en, columns m..en */
            for (k = m; k <= na; ++k) {
                int notlas = (k != na);
                if (k != m) {
                    p = h[k][k - 1];
                    q = h[k + 1][k - 1];
                    r = notlas ? h[k + 2][k - 1] : 0.0;
                    x = fabs(p) + fabs(q) + fabs(r);
                    if (x == 0.0) continue;
                    p /= x; q /= x; r /= x;
                }
                s = sqrt(p * p + q * q + r * r);
                if (p < 0.0) s = -s;
                if (k != m) h[k][k - 1] = -s * x;
                else if (l != m) h[k][k - 1] = -h[k][k - 1];
                p += s;
                x = p / s; y = q / s; z = r / s;
                q /= p; r /= p;
                /* row modification */
                for (int j = k; j < n; ++j) {
                    p = h[k][j] + q * h[k + 1][j];
                    if (notlas) { p += r * h[k + 2][j]; h[k + 2][j] -= p * z; }
                    h[k + 1][j] -= p * y;
                    h[k][j] -= p * x;
                }
                int jmax = (en < k + 3) ? en : k + 3;
                /* column modification */
                for (int i = 0; i <= jmax; ++i) {
                    p = x * h[i][k] + y * h[i][k + 1];
                    if (notlas) { p += z * h[i][k + 2]; h[i][k + 2] -= p * r; }
                    h[i][k + 1] -= p * q;
                    h[i][k] -= p;
                }
                /* accumulate transformations */
                for (int i = low; i <= hgh; ++i) {
                    p = x * zz[i][k] + y * zz[i][k + 1];
                    if (notlas) { p += z * zz[i][k + 2]; zz[i][k + 2] -= p * r; }
                    zz[i][k + 1] -= p * q;
                    zz[i][k] -= p;
                }
            }
        }
        if (l == en) {                                /* one root */
            h[en][en] = x + t;
            wr[en] = h[en][en];
            wi[en] = 0.0;
            en = na;
            continue;
        }
        /* two roots */
        p = (y - x) / 2.0;
        q = p * p + w;
        z = sqrt(fabs(q));
        h[en][en] = x + t;
        x = h[en][en];
        h[na][na] = y + t;
        if (q < 0.0) return -2;                       /* complex pair: not a reversible model */
        z = (p < 0.0) ? (p - z) : (p + z);
        wr[na] = x + z;
        wr[en] = wr[na];
        if (z != 0.0) wr[en] = x - w / z;
        wi[na] = 0.0;
        wi[en] = 0.0;
        x = h[en][na];
        s = fabs(x) + fabs(z);
        p = x / s;
        q = z / s;
        r = sqrt(p * p + q * q);
        p /= r; q /= r;
        for (int j = na; j < n; ++j) {                /* row modification */
            z = h[na][j];
            h[na][j] = q * z + p * h[en][j];
            h[en][j] = q * h[en][j] - p * z;
        }
        for (int i = 0; i <= en; ++i) {               /* column modification */
            z = h[i][na];
            h[i][na] = q * z + p * h[i][en];
            h[i][en] = q * h[i][en] - p * z;
        }
        for (int i = low; i <= hgh; ++i) {            /* accumulate */
            z = zz[i][na];
            zz[i][na] = q * z + p * zz[i][en];
            zz[i][en] = q * zz[i][en] - p * z;
        }
        en -= 2;
    }
    if (norm == 0.0) return 0;
    /* back-substitute for the vectors of the upper-triangular form (all roots real) */
    for (en = n - 1; en >= 0; --en) {
        p = wr[en];
        int m = en;
        h[en][en] = 1.0;
        for (int i = en - 1; i >= 0; --i) {
            w = h[i][i] - p;
            r = 0.0;
            for (int j = m; j <= en; ++j) r += h[i][j] * h[j][en];
            m = i;
            t = w;
            if (t == 0.0) {
                double tst1 = norm;
                t = tst1;
                do { t *= 0.01; } while (norm + t > tst1);
            }
            h[i][en] = -r / t;
            /* overflow control */
            t = fabs(h[i][en]);
            if (t != 0.0 && (t + 1.0 / t) <= t)
                for (int j = i; j <= en; ++j) h[j][en] /= t;
        }
    }
    /* multiply by the transformation matrix to give vectors of the original full matrix */
    for (int j = n - 1; j >= low; --j) {
        int mm = (j < hgh) ? j : hgh;
        for (int i = low; i <= hgh; ++i) {
            z = 0.0;
            for (k = low; k <= mm; ++k) z += zz[i][k] * h[k][j];
            zz[i][j] = z;
        }
    }
    return 0;
}

/* MatrixExponential::luinverse — Crout LU with implicit-scaling partial pivoting, then the
 * inverse column by column. A zero pivot is replaced by MachineAccuracy.EPSILON. */
static int luinverse(double om[NS][NS], double im[NS][NS], int n)
{
    const double EPSILON = 2.220446049250313e-16;
    int index[NS];
    double ws[NS];
    int imax = 0;
    for (int i = 0; i < n; ++i) {
        double amax = 0.0;
        for (int j = 0; j < n; ++j)
            if (fabs(om[i][j]) > amax) amax = fabs(om[i][j]);
        if (amax == 0.0) return -3;                   /* "Singular matrix encountered" */
        ws[i] = 1.0 / amax;
    }
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < j; ++i) {
            double sum = om[i][j];
            for (int k = 0; k < i; ++k) sum -= om[i][k] * om[k][j];
            om[i][j] = sum;
        }
        double amax = 0.0;
        for (int i = j; i < n; ++i) {
            double sum = om[i][j];
            for (int k = 0; k < j; ++k) sum -= om[i][k] * om[k][j];
            om[i][j] = sum;
            double dum = ws[i] * fabs(sum);
            if (dum >= amax) { imax = i; amax = dum; }
        }
        if (j != imax) {
            for (int k = 0; k < n; ++k) { double dum = om[imax][k]; om[imax][k] = om[j][k]; om[j][k] = dum; }
            ws[imax] = ws[j];
        }
        index[j] = imax;
        if (om[j][j] == 0.0) om[j][j] = EPSILON;
        if (j != n - 1) {
            double dum = 1.0 / om[j][j];
            for (int i = j + 1; i < n; ++i) om[i][j] *= dum;
        }
    }
    for (int jx = 0; jx < n; ++jx) {
        for (int i = 0; i < n; ++i) ws[i] = 0.0;
        ws[jx] = 1.0;
        int ix = -1;
        for (int i = 0; i < n; ++i) {
            int ip = index[i];
            double sum = ws[ip];
            ws[ip] = ws[i];
            if (ix != -1) {
                for (int j = ix; j < i; ++j) sum -= om[i][j] * ws[j];
            } else if (sum != 0.0) {
                ix = i;
            }
            ws[i] = sum;
        }
        for (int i = n - 1; i >= 0; --i) {
            double sum = ws[i];
            for (int j = i + 1; j < n; ++j) sum -= om[i][j] * ws[j];
            ws[i] = sum / om[i][i];
        }
        for (int i = 0; i < n; ++i) im[i][jx] = ws[i];
    }
    return 0;
}

/* MatrixExponential::updateByRelativeRates: copy R, elmhes, eltran, hqr2, luinverse. */
int pal_eigensystem(const double *R, double *Eval, double *Evec, double *Ievc)
{
    double amat[NS][NS], evec[NS][NS], ievc[NS][NS], tmp[NS][NS], evali[NS];
    int ordr[NS];
    memcpy(amat, R, sizeof(amat));
    elmhes(amat, ordr, NS);
    eltran(amat, evec, ordr, NS);
    int rc = hqr2(NS, amat, evec, Eval, evali);
    if (rc) return rc;
    memcpy(tmp, evec, sizeof(tmp));
    rc = luinverse(tmp, ievc, NS);
    if (rc) return rc;
    memcpy(Evec, evec, sizeof(evec));
    memcpy(Ievc, ievc, sizeof(ievc));
    return 0;
}

/* MatrixExponential::setDistance(D)V: clamp, iexp = Ievc * exp(t*Eval), P = |Evec*iexp|. */
void pal_transition_probs(const double *Eval, const double *Evec, const double *Ievc, double t, double *P)
{
    double iexp[NS][NS];
    if (t < 1e-9) t = 1e-9;
    for (int k = 0; k < NS; ++k) {
        double e = exp(t * Eval[k]);
        for (int j = 0; j < NS; ++j) iexp[k][j] = Ievc[k * NS + j] * e;
    }
    for (int i = 0; i < NS; ++i)
        for (int j = 0; j < NS; ++j) {
            double s = 0.0;
            for (int k = 0; k < NS; ++k) s += Evec[i * NS + k] * iexp[k][j];
            P[i * NS + j] = fabs(s);
        }
}

/* ---- discrete gamma ---------------------------------------------------------- */

/* GammaFunction::lnGamma — Pike & Hill (1966) Algorithm 291 as in PAML. */
static double pal_ln_gamma(double alpha)
{
    double x = alpha, f = 0.0, z;
    if (x < 7) {
        f = 1.0;
        z = x - 1.0;
        while (++z < 7.0) f *= z;
        x = z;
        f = -log(f);
    }
    z = 1.0 / (x * x);
    return f + (x - 0.5) * log(x) - x + 0.918938533204673 +
           (((-0.000595238095238 * z + 0.000793650793651) * z - 0.002777777777778) * z + 0.083333333333333) / x;
}

/* GammaFunction::incompleteGamma(x, alpha, lnGammaAlpha) — AS 32 as in PAML, accurate=1e-8, overflow=1e30. */
static double pal_incomplete_gamma(double x, double alpha, double ln_gamma_alpha)
{
    const double accurate = 1e-8, overflow = 1e30;
    double factor, gin, rn, a, b, an, dif, term;
    double pn0, pn1, pn2, pn3, pn4, pn5;
    if (x == 0.0) return 0.0;
    if (x < 0.0 || alpha <= 0.0) return NAN;
    factor = exp(alpha * log(x) - x - ln_gamma_alpha);
    if (x > 1 && x >= alpha) {
        /* continued fraction */
        a = 1 - alpha;
        b = a + x + 1;
        term = 0;
        pn0 = 1; pn1 = x; pn2 = x + 1; pn3 = x * b;
        gin = pn2 / pn3;
        do {
            a++;
            b += 2;
            term++;
            an = a * term;
            pn4 = b * pn2 - an * pn0;
            pn5 = b * pn3 - an * pn1;
            if (pn5 != 0) {
                rn = pn4 / pn5;
                dif = fabs(gin - rn);
                if (dif <= accurate) {
                    if (dif <= accurate * rn) break;
                }
                gin = rn;
            }
            pn0 = pn2; pn1 = pn3; pn2 = pn4; pn3 = pn5;
            if (fabs(pn4) >= overflow) {
                pn0 /= overflow; pn1 /= overflow; pn2 /= overflow; pn3 /= overflow;
            }
        } while (1);
        gin = 1 - factor * gin;
    } else {
        /* series expansion */
        gin = 1; term = 1; rn = alpha;
        do {
            rn++;
            term *= x / rn;
            gin += term;
        } while (term > accurate);
        gin *= factor / alpha;
    }
    return gin;
}

/* ErrorFunction::pointNormal — Odeh & Evans (1974) AS 70. */
static double pal_point_normal(double prob)
{
    const double a0 = -0.322232431088, a1 = -1, a2 = -0.342242088547, a3 = -0.0204231210245,
                 a4 = -0.453642210148e-4, b0 = 0.0993484626060, b1 = 0.588581570495, b2 = 0.531103462366,
                 b3 = 0.103537752850, b4 = 0.0038560700634;
    double p1 = (prob < 0.5) ? prob : 1 - prob;
    double y = sqrt(log(1 / (p1 * p1)));
    double z = y + ((((y * a4 + a3) * y + a2) * y + a1) * y + a0) / ((((y * b4 + b3) * y + b2) * y + b1) * y + b0);
    return (prob < 0.5) ? -z : z;
}

/* NormalDistribution::quantile(z, m, sd) = m + sqrt(2)*sd*ErrorFunction::inverseErf(2z-1);
 * inverseErf(z) = pointNormal(0.5*z + 0.5) / sqrt(2). */
static double pal_normal_quantile(double z, double m, double sd)
{
    double ierf = pal_point_normal(0.5 * (2.0 * z - 1.0) + 0.5) / sqrt(2.0);
    return m + sqrt(2.0) * sd * ierf;
}

/* GammaDistribution::pointChi2(prob, v) — AS 91 as in PAML. */
static double pal_point_chi2(double prob, double v)
{
    const double e = 0.5e-6, aa = 0.6931471805;
    double p = prob, g, xx, c, ch, a, q, p1, p2, t, x, b, s1, s2, s3, s4, s5, s6;
    if (p < 0.000002 || p > 0.999998 || v <= 0) return NAN;   /* PAL throws IllegalArgumentException */
    g = pal_ln_gamma(v / 2);
    xx = v / 2;
    c = xx - 1;
    if (v < -1.24 * log(p)) {
        ch = pow(p * xx * exp(g + xx * aa), 1 / xx);
        if (ch - e < 0) return ch;
    } else if (v > 0.32) {
        x = pal_normal_quantile(p, 0, 1);
        p1 = 0.222222 / v;
        ch = v * pow(x * sqrt(p1) + 1 - p1, 3.0);
        if (ch > 2.2 * v + 6) ch = -2 * (log(1 - p) - c * log(0.5 * ch) + g);
    } else {
        ch = 0.4;
        a = log(1 - p);
        do {
            q = ch;
            p1 = 1 + ch * (4.67 + ch);
            p2 = ch * (6.73 + ch * (6.66 + ch));
            t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2;
            ch -= (1 - exp(a + g + 0.5 * ch + c * aa) * p2 / p1) / t;
        } while (fabs(q / ch - 1) - 0.01 > 0);
    }
    do {
        q = ch;
        p1 = 0.5 * ch;
        t = pal_incomplete_gamma(p1, xx, g);
        if (!(t >= 0)) return NAN;
        p2 = p - t;
        t = p2 * exp(xx * aa + g + p1 - c * log(ch));
        b = t / ch;
        a = 0.5 * t - b * c;
        s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420;
        s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520;
        s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520;
        s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040;
        s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520;
        s6 = (120 + c * (346 + 127 * c)) / 5040;
        ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
    } while (fabs(q / ch - 1) > e);
    return ch;
}

/* GammaRates::makeGamma: median method, K equal-probability classes, renormalised to mean 1. */
int pal_gamma_rates(int K, double alpha, double *rates)
{
    double mean = 0.0;
    for (int i = 0; i < K; ++i) {
        /* GammaDistribution::quantile(y, shape, scale) = 0.5*scale*pointChi2(y, 2*shape), scale = 1/alpha */
        rates[i] = 0.5 * (1.0 / alpha) * pal_point_chi2((2.0 * i + 1.0) / (2.0 * K), 2.0 * alpha);
        if (rates[i] != rates[i]) return -1;
        mean += rates[i];
    }
    mean = mean / (double)K;
    for (int i = 0; i < K; ++i) rates[i] /= mean;
    return 0;
}
