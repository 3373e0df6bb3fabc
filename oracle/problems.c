/*
 * problems.c — residuals of the problem zoo for the CPU ORACLE (TEST INFRASTRUCTURE).
 *
 * Twins of scipy_dae_b200/problems.py (NumPy, drives the unmodified reference when the
 * goldens are minted) and scipy_dae_b200/csrc/problems.cuh (device functors).  Problem
 * definitions: SURVEY.md App. D (reference: README.md:37-47, examples/daes/akzo_nobel.py:18-50,
 * examples/daes/andrews.py:20-225; "transamp" is the IVP-test-set two-transistor amplifier).
 */
#include "dae_oracle.h"

#include <math.h>
#include <string.h>

/* Robertson, p = (k1, k2, k3) */
static void robertson(double t, const double *y, const double *yp, const double *p, double *out)
{
    (void)t;
    double k1 = p[0], k2 = p[1], k3 = p[2];
    out[0] = yp[0] - (-k1 * y[0] + k2 * y[1] * y[2]);
    out[1] = yp[1] - (k1 * y[0] - k2 * y[1] * y[2] - k3 * y[1] * y[1]);
    out[2] = y[0] + y[1] + y[2] - 1;
}

/* Akzo Nobel, p = (k1, k2, k3, k4, kbig, kla, ks) */
static void akzo_nobel(double t, const double *y, const double *yp, const double *p, double *out)
{
    (void)t;
    const double po2 = 0.9, hen = 737.0;
    double k1 = p[0], k2 = p[1], k3 = p[2], k4 = p[3], kbig = p[4], kla = p[5], ks = p[6];
    double sq = sqrt(y[1]);
    double y0sq = y[0] * y[0];
    double r1 = k1 * (y0sq * y0sq) * sq;
    double r2 = k2 * y[2] * y[3];
    double r3 = k2 / kbig * y[0] * y[4];
    double r4 = k3 * y[0] * (y[3] * y[3]);
    double r5 = k4 * (y[5] * y[5]) * sq;
    double fin = kla * (po2 / hen - y[1]);
    out[0] = yp[0] - (-2.0 * r1 + r2 - r3 - r4);
    out[1] = yp[1] - (-0.5 * r1 - r4 - 0.5 * r5 + fin);
    out[2] = yp[2] - (r1 - r2 + r3);
    out[3] = yp[3] - (-r2 + r3 - 2.0 * r4);
    out[4] = yp[4] - (r2 - r3 + r5);
    out[5] = -(ks * y[0] * y[3] - y[5]);
}

/* two-transistor amplifier, p = (Ub, R0, beta) */
static void transamp(double t, const double *y, const double *yp, const double *p, double *out)
{
    const double UF = 0.026, ALPHA = 0.99, R = 9000.0;
    const double C1 = 1e-6, C2 = 2e-6, C3 = 3e-6, C4 = 4e-6, C5 = 5e-6;
    double Ub = p[0], R0 = p[1], beta = p[2];
    /* reciprocal forms, operation by operation as in problems.py and problems.cuh */
    const double rUF = 1.0 / UF, rR = 1.0 / R, r2R = 2.0 / R;
    double rR0 = 1.0 / R0;
    double Ue = 0.1 * sin((200.0 * M_PI) * t);
    double g23 = beta * (exp((y[1] - y[2]) * rUF) - 1.0);
    double g56 = beta * (exp((y[4] - y[5]) * rUF) - 1.0);
    double UbR = Ub * rR;
    double f0 = (y[0] - Ue) * rR0;
    double f1 = -UbR + y[1] * r2R - (ALPHA - 1.0) * g23;
    double f2 = -g23 + y[2] * rR;
    double f3 = -UbR + y[3] * rR + ALPHA * g23;
    double f4 = -UbR + y[4] * r2R - (ALPHA - 1.0) * g56;
    double f5 = -g56 + y[5] * rR;
    double f6 = -UbR + y[6] * rR + ALPHA * g56;
    double f7 = y[7] * rR;
    out[0] = (-C1 * yp[0] + C1 * yp[1]) - f0;
    out[1] = (C1 * yp[0] - C1 * yp[1]) - f1;
    out[2] = (-C2 * yp[2]) - f2;
    out[3] = (-C3 * yp[3] + C3 * yp[4]) - f3;
    out[4] = (C3 * yp[3] - C3 * yp[4]) - f4;
    out[5] = (-C4 * yp[5]) - f5;
    out[6] = (-C5 * yp[6] + C5 * yp[7]) - f6;
    out[7] = (C5 * yp[6] - C5 * yp[7]) - f7;
}

/* Andrews' squeezer, stabilised index-1 form n = 26 (examples/daes/andrews.py:79-225),
 * p = (mom, c0): driving torque and spring constant. y = (q[7], u[7], la[6], mu[6]),
 * yp = (q'[7], u'[7], lambda[6], mu_mult[6]). */
static void andrews(double t, const double *y, const double *yp, const double *p, double *out)
{
    (void)t;
    const double m1 = 0.04325, m2 = 0.00365, m3 = 0.02373, m4 = 0.00706, m5 = 0.07050,
                 m6 = 0.00706, m7 = 0.05498;
    const double I1 = 2.194e-6, I2 = 4.410e-7, I3 = 5.255e-6, I4 = 5.667e-7, I5 = 1.169e-5,
                 I6 = 5.667e-7, I7 = 1.912e-5;
    const double xa = -0.06934, ya = -0.00227, xb = -0.03635, yb = 0.03273, xc = 0.014, yc = 0.072;
    const double d = 0.028, da = 0.0115, e = 0.02, ea = 0.01421, zf = 0.02, fa = 0.01421;
    const double rr = 0.007, ra = 0.00092, ss = 0.035, sa = 0.01874, sb = 0.01043, sc = 0.018;
    const double sd = 0.02, zt = 0.04, ta = 0.02308, tb = 0.00916, u = 0.04, ua = 0.01228,
                 ub = 0.00449, l0 = 0.07785;
    double mom = p[0], c0 = p[1];
    const double *q = y, *v = y + 7;
    const double *qd = yp, *vd = yp + 7, *la = yp + 14, *mu = yp + 20;

    double sin_be = sin(q[0]), cos_be = cos(q[0]);
    double sin_Th = sin(q[1]), cos_Th = cos(q[1]);
    double sin_ga = sin(q[2]), cos_ga = cos(q[2]);
    double sin_Phi = sin(q[3]), cos_Phi = cos(q[3]);
    double sin_de = sin(q[4]), cos_de = cos(q[4]);
    double sin_Om = sin(q[5]), cos_Om = cos(q[5]);
    double sin_ep = sin(q[6]), cos_ep = cos(q[6]);
    double sin_be_Th = sin(q[0] + q[1]), cos_be_Th = cos(q[0] + q[1]);
    double sin_Ph_de = sin(q[3] + q[4]), cos_Ph_de = cos(q[3] + q[4]);
    double sin_Om_ep = sin(q[5] + q[6]), cos_Om_ep = cos(q[5] + q[6]);

    /* mass matrix (symmetric, block tridiagonal) */
    double M00 = m1 * ra * ra + m2 * (rr * rr - 2 * da * rr * cos_Th + da * da) + I1 + I2;
    double M01 = m2 * (da * da - da * rr * cos_Th) + I2;
    double M11 = m2 * da * da + I2;
    double M22 = m3 * (sa * sa + sb * sb) + I3;
    double M33 = m4 * (e - ea) * (e - ea) + I4;
    double M34 = m4 * ((e - ea) * (e - ea) + zt * (e - ea) * sin_Phi) + I4;
    double M44 = m4 * (zt * zt + 2 * zt * (e - ea) * sin_Phi + (e - ea) * (e - ea)) + m5 * (ta * ta + tb * tb) + I4 + I5;
    double M55 = m6 * (zf - fa) * (zf - fa) + I6;
    double M56 = m6 * ((zf - fa) * (zf - fa) - u * (zf - fa) * sin_Om) + I6;
    double M66 = m6 * ((zf - fa) * (zf - fa) - 2 * u * (zf - fa) * sin_Om + u * u) + m7 * (ua * ua + ub * ub) + I6 + I7;

    /* forces */
    double xd = sd * cos_ga + sc * sin_ga + xb;
    double yd = sd * sin_ga - sc * cos_ga + yb;
    double L = sqrt((xd - xc) * (xd - xc) + (yd - yc) * (yd - yc));
    double Fs = -c0 * (L - l0) / L;
    double Fx = Fs * (xd - xc), Fy = Fs * (yd - yc);
    double h[7];
    h[0] = mom - m2 * da * rr * v[1] * (v[1] + 2 * v[0]) * sin_Th;
    h[1] = m2 * da * rr * v[0] * v[0] * sin_Th;
    h[2] = Fx * (sc * cos_ga - sd * sin_ga) + Fy * (sd * cos_ga + sc * sin_ga);
    h[3] = m4 * zt * (e - ea) * v[4] * v[4] * cos_Phi;
    h[4] = -m4 * zt * (e - ea) * v[3] * (v[3] + 2 * v[4]) * cos_Phi;
    h[5] = -m6 * u * (zf - fa) * v[6] * v[6] * cos_Om;
    h[6] = m6 * u * (zf - fa) * v[5] * (v[5] + 2 * v[6]) * cos_Om;

    /* constraints and their Jacobian G[6][7] */
    double g[6];
    g[0] = rr * cos_be - d * cos_be_Th - ss * sin_ga - xb;
    g[1] = rr * sin_be - d * sin_be_Th + ss * cos_ga - yb;
    g[2] = rr * cos_be - d * cos_be_Th - e * sin_Ph_de - zt * cos_de - xa;
    g[3] = rr * sin_be - d * sin_be_Th + e * cos_Ph_de - zt * sin_de - ya;
    g[4] = rr * cos_be - d * cos_be_Th - zf * cos_Om_ep - u * sin_ep - xa;
    g[5] = rr * sin_be - d * sin_be_Th - zf * sin_Om_ep + u * cos_ep - ya;

    double G[6][7];
    memset(G, 0, sizeof(G));
    double a0 = -rr * sin_be + d * sin_be_Th, a1 = d * sin_be_Th;
    double b0 = rr * cos_be - d * cos_be_Th, b1 = -d * cos_be_Th;
    G[0][0] = a0; G[0][1] = a1; G[0][2] = -ss * cos_ga;
    G[1][0] = b0; G[1][1] = b1; G[1][2] = -ss * sin_ga;
    G[2][0] = a0; G[2][1] = a1; G[2][3] = -e * cos_Ph_de; G[2][4] = -e * cos_Ph_de + zt * sin_de;
    G[3][0] = b0; G[3][1] = b1; G[3][3] = -e * sin_Ph_de; G[3][4] = -e * sin_Ph_de - zt * cos_de;
    G[4][0] = a0; G[4][1] = a1; G[4][5] = zf * sin_Om_ep; G[4][6] = zf * sin_Om_ep - u * cos_ep;
    G[5][0] = b0; G[5][1] = b1; G[5][5] = -zf * cos_Om_ep; G[5][6] = -zf * cos_Om_ep - u * sin_ep;

    double Mvd[7];
    Mvd[0] = M00 * vd[0] + M01 * vd[1];
    Mvd[1] = M01 * vd[0] + M11 * vd[1];
    Mvd[2] = M22 * vd[2];
    Mvd[3] = M33 * vd[3] + M34 * vd[4];
    Mvd[4] = M34 * vd[3] + M44 * vd[4];
    Mvd[5] = M55 * vd[5] + M56 * vd[6];
    Mvd[6] = M56 * vd[5] + M66 * vd[6];

    for (int i = 0; i < 7; ++i) {
        double gt_mu = 0.0, gt_la = 0.0;
        for (int k = 0; k < 6; ++k) { gt_mu += G[k][i] * mu[k]; gt_la += G[k][i] * la[k]; }
        out[i] = qd[i] - v[i] + gt_mu;
        out[7 + i] = Mvd[i] - h[i] - gt_la;
    }
    for (int k = 0; k < 6; ++k) {
        double gu = 0.0;
        for (int i = 0; i < 7; ++i) gu += G[k][i] * v[i];
        out[14 + k] = gu;
        out[20 + k] = g[k];
    }
}

/* ---- the reference's own known-answer test problems (twins of problems.py) ---- */
static void rational(double t, const double *y, const double *yp, const double *p, double *out)
{
    (void)p;
    out[0] = yp[0] - y[1] / t;
    out[1] = yp[1] - y[1] * (y[0] + 2 * y[1] - 1) / (t * (y[0] - 1));
}
static void rational_jac(double t, const double *y, const double *yp, const double *p, double *Jy, double *Jyp)
{
    (void)yp; (void)p;
    Jy[0] = -0.0; Jy[1] = -(1 / t);
    Jy[2] = -(-2 * y[1] * y[1] / (t * (y[0] - 1) * (y[0] - 1)));
    Jy[3] = -((y[0] + 4 * y[1] - 1) / (t * (y[0] - 1)));
    Jyp[0] = 1; Jyp[1] = 0; Jyp[2] = 0; Jyp[3] = 1;
}
static void linear(double t, const double *y, const double *yp, const double *p, double *out)
{
    (void)t; (void)p;
    out[0] = yp[0] - (-y[0] - 5 * y[1]);
    out[1] = yp[1] - (y[0] + y[1]);
}
static void linear_jac(double t, const double *y, const double *yp, const double *p, double *Jy, double *Jyp)
{
    (void)t; (void)y; (void)yp; (void)p;
    Jy[0] = 1; Jy[1] = 5; Jy[2] = -1; Jy[3] = -1;
    Jyp[0] = 1; Jyp[1] = 0; Jyp[2] = 0; Jyp[3] = 1;
}
static void cosine(double t, const double *y, const double *yp, const double *p, double *out)
{
    (void)y;
    out[0] = yp[0] - p[0] * cos(t);
}
static void weissinger(double t, const double *y, const double *yp, const double *p, double *out)
{
    double u = y[0] / p[0], up = yp[0] / p[0];
    out[0] = t * (u * u) * (up * up * up) - (u * u * u) * (up * up) + t * (t * t + 1) * up - (t * t) * u;
}
static void brenan(double t, const double *y, const double *yp, const double *p, double *out)
{
    out[0] = yp[0] - t * yp[1] + y[0] - (1 + t) * y[1];
    out[1] = y[1] - p[0] * sin(t);
}

/* one-transistor amplifier of the reference's examples/daes/stiff_transistor_amplifier.py:20-50,
 * p = (Ub, R0, beta) */
static void transamp1(double t, const double *y, const double *yp, const double *p, double *out)
{
    const double UF = 0.026, ALPHA = 0.99, R = 9000.0;
    const double C1 = 1e-6, C2 = 2e-6, C3 = 3e-6;
    double Ub = p[0], R0 = p[1], beta = p[2];
    double Ue = 0.4 * sin((200.0 * M_PI) * t);
    double f23 = beta * (exp((y[1] - y[2]) / UF) - 1.0);
    out[0] = (Ue - y[0]) / R0 + C1 * (yp[1] - yp[0]);
    out[1] = (Ub - y[1]) / R - y[1] / R + C1 * (yp[0] - yp[1]) - (1.0 - ALPHA) * f23;
    out[2] = f23 - y[2] / R - C2 * yp[2];
    out[3] = (Ub - y[3]) / R + C3 * (yp[4] - yp[3]) - ALPHA * f23;
    out[4] = -y[4] / R + C3 * (yp[3] - yp[4]);
}
/* y' = p y^2 (finite-time blow-up: the deterministic failure-path problem) */
static void blowup(double t, const double *y, const double *yp, const double *p, double *out)
{
    (void)t;
    out[0] = yp[0] - p[0] * y[0] * y[0];
}

/* ---- analytic Jacobians (the reference's `jac=` callable, base.py:299-331), row-major n*n;
 * twins of problems.py: robertson_jacobian, akzo_jacobian, transamp_jacobian ---- */
static void robertson_jac(double t, const double *y, const double *yp, const double *p,
                          double *Jy, double *Jyp)
{
    (void)t; (void)yp;
    double k1 = p[0], k2 = p[1], k3 = p[2];
    memset(Jy, 0, sizeof(double) * 9); memset(Jyp, 0, sizeof(double) * 9);
    Jy[0] = k1;  Jy[1] = -k2 * y[2];                 Jy[2] = -k2 * y[1];
    Jy[3] = -k1; Jy[4] = k2 * y[2] + 2.0 * k3 * y[1]; Jy[5] = k2 * y[1];
    Jy[6] = 1.0; Jy[7] = 1.0;                        Jy[8] = 1.0;
    Jyp[0] = 1.0; Jyp[4] = 1.0;
}

static void akzo_nobel_jac(double t, const double *y, const double *yp, const double *p,
                           double *Jy, double *Jyp)
{
    (void)t; (void)yp;
    double k1 = p[0], k2 = p[1], k3 = p[2], k4 = p[3], kbig = p[4], kla = p[5], ks = p[6];
    double sq = sqrt(y[1]), hsq = 0.5 / sq;
    double y0c = y[0] * y[0] * y[0];
    double r1_0 = 4.0 * k1 * y0c * sq, r1_1 = k1 * (y0c * y[0]) * hsq;
    double r2_2 = k2 * y[3], r2_3 = k2 * y[2];
    double kk = k2 / kbig;
    double r3_0 = kk * y[4], r3_4 = kk * y[0];
    double r4_0 = k3 * (y[3] * y[3]), r4_3 = 2.0 * k3 * y[0] * y[3];
    double r5_1 = k4 * (y[5] * y[5]) * hsq, r5_5 = 2.0 * k4 * y[5] * sq;
    double df[36];
    memset(df, 0, sizeof(df)); memset(Jyp, 0, sizeof(double) * 36);
#define DF(i, j) df[(i) * 6 + (j)]
    DF(0, 0) = -2.0 * r1_0 - r3_0 - r4_0; DF(0, 1) = -2.0 * r1_1; DF(0, 2) = r2_2;
    DF(0, 3) = r2_3 - r4_3; DF(0, 4) = -r3_4;
    DF(1, 0) = -0.5 * r1_0 - r4_0; DF(1, 1) = -0.5 * r1_1 - 0.5 * r5_1 - kla;
    DF(1, 3) = -r4_3; DF(1, 5) = -0.5 * r5_5;
    DF(2, 0) = r1_0 + r3_0; DF(2, 1) = r1_1; DF(2, 2) = -r2_2; DF(2, 3) = -r2_3; DF(2, 4) = r3_4;
    DF(3, 0) = r3_0 - 2.0 * r4_0; DF(3, 2) = -r2_2; DF(3, 3) = -r2_3 - 2.0 * r4_3; DF(3, 4) = r3_4;
    DF(4, 0) = -r3_0; DF(4, 1) = r5_1; DF(4, 2) = r2_2; DF(4, 3) = r2_3; DF(4, 4) = -r3_4; DF(4, 5) = r5_5;
    DF(5, 0) = ks * y[3]; DF(5, 3) = ks * y[0]; DF(5, 5) = -1.0;
#undef DF
    for (int k = 0; k < 36; ++k) Jy[k] = -df[k];
    for (int i = 0; i < 5; ++i) Jyp[i * 6 + i] = 1.0;
}

static void transamp_jac(double t, const double *y, const double *yp, const double *p,
                         double *Jy, double *Jyp)
{
    (void)t; (void)yp;
    const double UF = 0.026, a = 0.99, R = 9000.0;
    const double C1 = 1e-6, C2 = 2e-6, C3 = 3e-6, C4 = 4e-6, C5 = 5e-6;
    double R0 = p[1], beta = p[2];
    double e23 = beta * exp((y[1] - y[2]) / UF) / UF;
    double e56 = beta * exp((y[4] - y[5]) / UF) / UF;
    double df[64];
    memset(df, 0, sizeof(df)); memset(Jyp, 0, sizeof(double) * 64);
#define DF(i, j) df[(i) * 8 + (j)]
#define M(i, j) Jyp[(i) * 8 + (j)]
    DF(0, 0) = 1.0 / R0;
    DF(1, 1) = 2.0 / R - (a - 1.0) * e23; DF(1, 2) = (a - 1.0) * e23;
    DF(2, 1) = -e23; DF(2, 2) = e23 + 1.0 / R;
    DF(3, 1) = a * e23; DF(3, 2) = -a * e23; DF(3, 3) = 1.0 / R;
    DF(4, 4) = 2.0 / R - (a - 1.0) * e56; DF(4, 5) = (a - 1.0) * e56;
    DF(5, 4) = -e56; DF(5, 5) = e56 + 1.0 / R;
    DF(6, 4) = a * e56; DF(6, 5) = -a * e56; DF(6, 6) = 1.0 / R;
    DF(7, 7) = 1.0 / R;
    M(0, 0) = -C1; M(0, 1) = C1; M(1, 0) = C1; M(1, 1) = -C1;
    M(2, 2) = -C2;
    M(3, 3) = -C3; M(3, 4) = C3; M(4, 3) = C3; M(4, 4) = -C3;
    M(5, 5) = -C4;
    M(6, 6) = -C5; M(6, 7) = C5; M(7, 6) = C5; M(7, 7) = -C5;
#undef DF
#undef M
    for (int k = 0; k < 64; ++k) Jy[k] = -df[k];
}

/* ---- event functions (twins of problems.py: *_EVENTS and of the device functors' event()) ---- */
static double robertson_event(int which, double t, const double *y, const double *yp, const double *p)
{
    (void)t;
    return which == 0 ? y[0] - y[2] : yp[0] + yp[1] + p[0] * y[0] * 0.5;
}
static double akzo_nobel_event(int which, double t, const double *y, const double *yp, const double *p)
{
    (void)which; (void)t; (void)yp; (void)p;
    return y[0] - y[2];
}
static double cosine_event(int which, double t, const double *y, const double *yp, const double *p)
{
    return which == 0 ? y[0] - yp[0] : y[0] - p[0] * t * 0.5;
}

static const struct {
    const char *name;
    int n, n_params;
    orc_residual_fn fn;
    orc_jacobian_fn jac;
    orc_event_fn event;
} REGISTRY[] = {
    {"robertson", 3, 3, robertson, robertson_jac, robertson_event},
    {"akzo_nobel", 6, 7, akzo_nobel, akzo_nobel_jac, akzo_nobel_event},
    {"transamp", 8, 3, transamp, transamp_jac, 0},
    {"andrews", 26, 2, andrews, 0, 0},
    {"rational", 2, 0, rational, rational_jac, 0},
    {"linear", 2, 0, linear, linear_jac, 0},
    {"cosine", 1, 1, cosine, 0, cosine_event},
    {"weissinger", 1, 1, weissinger, 0, 0},
    {"brenan", 2, 1, brenan, 0, 0},
    {"transamp1", 5, 3, transamp1, 0, 0},
    {"blowup", 1, 1, blowup, 0, 0},
};
#define N_PROBLEMS ((int)(sizeof(REGISTRY) / sizeof(REGISTRY[0])))

int orc_problem_lookup(const char *name, int *n, int *n_params)
{
    for (int i = 0; i < N_PROBLEMS; ++i)
        if (strcmp(REGISTRY[i].name, name) == 0) {
            if (n) *n = REGISTRY[i].n;
            if (n_params) *n_params = REGISTRY[i].n_params;
            return i;
        }
    return -1;
}

int orc_problem_dims(int problem_id, int *n, int *n_params)
{
    if (problem_id < 0 || problem_id >= N_PROBLEMS) return -1;
    if (n) *n = REGISTRY[problem_id].n;
    if (n_params) *n_params = REGISTRY[problem_id].n_params;
    return 0;
}

orc_residual_fn orc_problem_residual(int problem_id)
{
    if (problem_id < 0 || problem_id >= N_PROBLEMS) return 0;
    return REGISTRY[problem_id].fn;
}

orc_event_fn orc_problem_event(int problem_id)
{
    if (problem_id < 0 || problem_id >= N_PROBLEMS) return 0;
    return REGISTRY[problem_id].event;
}

orc_event_fn orc_event_of(orc_residual_fn fn)
{
    for (int i = 0; i < N_PROBLEMS; ++i)
        if (REGISTRY[i].fn == fn) return REGISTRY[i].event;
    return 0;
}

orc_jacobian_fn orc_problem_jacobian(int problem_id)
{
    if (problem_id < 0 || problem_id >= N_PROBLEMS) return 0;
    return REGISTRY[problem_id].jac;
}

orc_jacobian_fn orc_jacobian_of(orc_residual_fn fn)
{
    for (int i = 0; i < N_PROBLEMS; ++i)
        if (REGISTRY[i].fn == fn) return REGISTRY[i].jac;
    return 0;
}

int orc_problem_jac_eval(int problem_id, double t, const double *y, const double *yp,
                         const double *p, double *Jy, double *Jyp)
{
    orc_jacobian_fn jf = orc_problem_jacobian(problem_id);
    if (!jf) return -1;
    jf(t, y, yp, p, Jy, Jyp);
    return 0;
}

/* residual evaluation exported for tests (checks the C twin against problems.py) */
int orc_problem_eval(int problem_id, double t, const double *y, const double *yp,
                     const double *p, double *out)
{
    orc_residual_fn fn = orc_problem_residual(problem_id);
    if (!fn) return -1;
    fn(t, y, yp, p, out);
    return 0;
}
