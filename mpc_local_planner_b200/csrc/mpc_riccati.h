// mpc_riccati.h -- warp-cooperative Riccati factorisation of the bordered block-tridiagonal KKT system of one OCP
// instance (the replacement of MUMPS' general sparse LDL^T inside Ipopt, R/src/controller.cpp:380-421).
//
// Unknowns: dw_k = (dx_k, du_k), nu+_k, d(dt);  dx_{k+1} = A_k dx_k + B_k du_k + d_k d(dt) + e_k,  dx_0 = 0,
// dx_{N-1,j} = 0 for fixed terminal components.  Stage state y_k = (dx_k, du_{k-1}) in R^5 (the previous control
// enters through the control-rate cross block C_k), parameters theta^ = (1, d(dt), pi_0, pi_1, pi_2).
// Value function V_k = 1/2 y'P y + y' PI theta^ + 1/2 theta^' TH theta^   (DESIGN.md "Riccati").
//
// The work of one backward stage is a fixed list of tiny dot-product TASKS (<= 3 terms) grouped into passes A, B, C, D
// with no read-after-write hazard inside a pass; lane t of the warp executes task t of each pass on a per-warp
// shared-memory scratch `sm`, with a __syncwarp between passes.  The same task tables run on the CPU warp emulator
// of tests/emu (test infrastructure).
#pragma once
#include "mpc_core.h"

// ---- per-warp scratch map (doubles) ----
#define R_ZERO 0    // 4 zeros
#define R_P 4       // 5x5
#define R_PI 29     // 5x5  [row][col], col: 0 const, 1 dt, 2..4 pi
#define R_TH 54     // 5x5
#define R_EXP 79    // expanded stage record, 63 words:
#define R_HXX (R_EXP + 0)    // 3x3 (+delta on the diagonal)
#define R_HXU (R_EXP + 9)    // 3x2
#define R_HUU (R_EXP + 15)   // 2x2 (+delta)
#define R_MHX (R_EXP + 19)   // 3x5: col 0 = g_x, col 1 = hb_x
#define R_MHV (R_EXP + 34)   // 2x5: col 0 = g_u, col 1 = hb_u
#define R_CH (R_EXP + 44)    // 2x3: chat_0 = e, chat_1 = d
#define R_CD (R_EXP + 50)    // 2x2: diag(C)
#define R_A (R_EXP + 54)     // 3
#define R_B (R_EXP + 57)     // 3x2
#define R_EXP_WORDS 63
#define R_T1 142    // 3x2
#define R_T2 148    // 2x2
#define R_PA 152    // 3
#define R_PPA 155   // 2
#define R_WM 157    // 5x2
#define R_MMXX 167  // 3x3
#define R_MMXV 176  // 3x2
#define R_MMVV 182  // 2x2
#define R_NNX 186   // 3x5
#define R_NNV 201   // 2x5
#define R_TT 211    // 5x5 (entries i<2, j>=i)
#define R_LAMB 236  // 2x2
#define R_KG 240    // 2x5
#define R_KT 250    // 2x5
#define R_WORDS 264
#define RIC_WORDS 50  // per-stage gains kept for the forward pass: Px rows(15) PIx rows(15) KG(10) KT(10)

struct RTask
{
    short out, out2, z1, z2, z3, z4, xb, xs, yb, ys, L;
    short neg;  // 1: out = z - dot
};

HD inline RTask rtask_none()
{
    RTask t; t.out = -1; t.out2 = -1; t.z1 = t.z2 = t.z3 = t.z4 = R_ZERO; t.xb = t.yb = R_ZERO; t.xs = t.ys = 0; t.L = 0; t.neg = 0;
    return t;
}
HD inline int rP(int i, int j) { return R_P + i * 5 + j; }
HD inline int rPI(int i, int c) { return R_PI + i * 5 + c; }
HD inline int rTH(int i, int j) { return R_TH + i * 5 + j; }
HD inline int rTT(int i, int j) { return R_TT + i * 5 + j; }
HD inline int rWg(int r, int col) { return col < 2 ? R_WM + r * 2 + col : rPI(r, col); }
HD inline int rWgs(int col) { return col < 2 ? 2 : 5; }

HD inline void run_task(double* sm, const RTask& t)
{
    if (t.out < 0) return;
    double acc = sm[t.z1] + sm[t.z2] + sm[t.z3] + sm[t.z4];
    double d = 0.0;
#pragma unroll
    for (int l = 0; l < 3; ++l)
        if (l < t.L) d += sm[t.xb + l * t.xs] * sm[t.yb + l * t.ys];
    acc = t.neg ? acc - d : acc + d;
    sm[t.out] = acc;
    if (t.out2 >= 0) sm[t.out2] = acc;
}

// pass A: T1 = Pxx B + Pxp, T2 = Ppx B, PA = Pxx a, PPA = Ppx a, WM = P chat + PI (cols 0,1)
HD inline RTask rtask_A(int t)
{
    RTask k = rtask_none();
    if (t < 6) { int i = t / 2, j = t % 2; k.out = R_T1 + i * 2 + j; k.z1 = rP(i, 3 + j); k.xb = rP(i, 0); k.xs = 1; k.yb = R_B + j; k.ys = 2; k.L = 3; }
    else if (t < 10) { int i = (t - 6) / 2, j = (t - 6) % 2; k.out = R_T2 + i * 2 + j; k.xb = rP(3 + i, 0); k.xs = 1; k.yb = R_B + j; k.ys = 2; k.L = 3; }
    else if (t < 13) { int i = t - 10; k.out = R_PA + i; k.xb = rP(i, 0); k.xs = 1; k.yb = R_A; k.ys = 1; k.L = 3; }
    else if (t < 15) { int i = t - 13; k.out = R_PPA + i; k.xb = rP(3 + i, 0); k.xs = 1; k.yb = R_A; k.ys = 1; k.L = 3; }
    else if (t < 25) { int r = (t - 15) / 2, col = (t - 15) % 2; k.out = R_WM + r * 2 + col; k.z1 = rPI(r, col); k.xb = rP(r, 0); k.xs = 1; k.yb = R_CH + col * 3; k.ys = 1; k.L = 3; }
    return k;
}
// pass B (53 tasks)
HD inline RTask rtask_B(int t)
{
    RTask k = rtask_none();
    if (t < 4)
    {
        int i = t / 2, j = t % 2;
        k.out = R_MMVV + i * 2 + j; k.z1 = R_HUU + i * 2 + j; k.z2 = R_T2 + i * 2 + j; k.z3 = rP(3 + i, 3 + j);
        k.xb = R_B + i; k.xs = 2; k.yb = R_T1 + j; k.ys = 2; k.L = 3;
    }
    else if (t < 10)
    {
        int i = (t - 4) / 2, j = (t - 4) % 2;
        k.out = R_MMXV + i * 2 + j; k.z1 = R_HXU + i * 2 + j; k.z2 = R_T1 + i * 2 + j;
        if (i == 2) { k.xb = R_A; k.xs = 1; k.yb = R_T1 + j; k.ys = 2; k.L = 3; }
    }
    else if (t < 19)
    {
        int i = (t - 10) / 3, j = (t - 10) % 3;
        k.out = R_MMXX + i * 3 + j; k.z1 = R_HXX + i * 3 + j; k.z2 = rP(i, j);
        if (i == 2) k.z3 = R_PA + j;
        if (j == 2) k.z4 = R_PA + i;
        if (i == 2 && j == 2) { k.xb = R_A; k.xs = 1; k.yb = R_PA; k.ys = 1; k.L = 3; }
    }
    else if (t < 34)
    {
        int i = (t - 19) / 5, col = (t - 19) % 5;
        k.out = R_NNX + i * 5 + col; k.z1 = R_MHX + i * 5 + col; k.z2 = rWg(i, col);
        if (i == 2) { k.xb = R_A; k.xs = 1; k.yb = rWg(0, col); k.ys = rWgs(col); k.L = 3; }
    }
    else if (t < 44)
    {
        int i = (t - 34) / 5, col = (t - 34) % 5;
        k.out = R_NNV + i * 5 + col; k.z1 = R_MHV + i * 5 + col; k.z2 = rWg(3 + i, col);
        k.xb = R_B + i; k.xs = 2; k.yb = rWg(0, col); k.ys = rWgs(col); k.L = 3;
    }
    else if (t < 53)
    {
        int q = t - 44;
        int i = q < 5 ? 0 : 1, j = q < 5 ? q : q - 4;  // (0,0..4), (1,1..4)
        k.out = rTT(i, j); k.z1 = rTH(i, j);
        k.xb = R_CH + i * 3; k.xs = 1; k.yb = rWg(0, j); k.ys = rWgs(j); k.L = 3;
    }
    return k;
}
// pass C (23 tasks): KG = Lambda [MM_vx | diag C], KT = Lambda NN_v, second dot of TT
HD inline RTask rtask_C(int t)
{
    RTask k = rtask_none();
    if (t < 10)
    {
        int i = t / 5, j = t % 5;
        k.out = R_KG + i * 5 + j; k.xb = R_LAMB + i * 2; k.xs = 1;
        k.yb = (j < 3) ? R_MMXV + j * 2 : R_CD + (j - 3) * 2; k.ys = 1; k.L = 2;
    }
    else if (t < 20)
    {
        int i = (t - 10) / 5, col = (t - 10) % 5;
        k.out = R_KT + i * 5 + col; k.xb = R_LAMB + i * 2; k.xs = 1; k.yb = R_NNV + col; k.ys = 5; k.L = 2;
    }
    else if (t < 23)
    {
        int q = t - 20;
        int i = q == 2 ? 1 : 0, j = q == 0 ? 0 : 1;
        k.out = rTT(i, j); k.z1 = rTT(i, j); k.xb = rPI(0, i); k.xs = 5; k.yb = R_CH + j * 3; k.ys = 1; k.L = 3;
    }
    return k;
}
// pass D (55 tasks): Schur complements
HD inline void ut_pair(int q, int& i, int& j)
{
    // q-th pair (i <= j) of a 5x5 upper triangle, row-major
    int r = 0, base = 0;
    while (q >= base + (5 - r)) { base += 5 - r; ++r; }
    i = r; j = r + (q - base);
}
HD inline RTask rtask_D(int t)
{
    RTask k = rtask_none();
    if (t < 15)
    {
        int i, j; ut_pair(t, i, j);
        k.out = rP(i, j); k.out2 = rP(j, i);
        k.z1 = (i < 3 && j < 3) ? R_MMXX + i * 3 + j : R_ZERO;
        k.xb = (i < 3) ? R_MMXV + i * 2 : R_CD + (i - 3) * 2; k.xs = 1; k.yb = R_KG + j; k.ys = 5; k.L = 2; k.neg = 1;
    }
    else if (t < 40)
    {
        int i = (t - 15) / 5, col = (t - 15) % 5;
        k.out = rPI(i, col); k.z1 = (i < 3) ? R_NNX + i * 5 + col : R_ZERO;
        k.xb = (i < 3) ? R_MMXV + i * 2 : R_CD + (i - 3) * 2; k.xs = 1; k.yb = R_KT + col; k.ys = 5; k.L = 2; k.neg = 1;
    }
    else if (t < 55)
    {
        int i, j; ut_pair(t - 40, i, j);
        k.out = rTH(i, j); k.out2 = rTH(j, i);
        k.z1 = (i < 2) ? rTT(i, j) : rTH(i, j);
        k.xb = R_NNV + i; k.xs = 5; k.yb = R_KT + j; k.ys = 5; k.L = 2; k.neg = 1;
    }
    return k;
}

// value of expanded-record entry `idx` (0..62) of stage k from the [field][k] record array `rec`
HD inline double expand_entry(const double* rec, int N, int k, int idx, double delta, int dt_free)
{
#define RC(f) rec[(f) * N + k]
    if (idx < 9) { int i = idx / 3, j = idx % 3; int a = i < j ? i : j, b = i < j ? j : i; return RC(MPCB200_K_H + hidx(a, b)) + (i == j ? delta : 0.0); }
    if (idx < 15) { int q = idx - 9, i = q / 2, j = q % 2; return RC(MPCB200_K_H + hidx(i, 3 + j)); }
    if (idx < 19) { int q = idx - 15, i = q / 2, j = q % 2; int a = i < j ? i : j, b = i < j ? j : i; return RC(MPCB200_K_H + hidx(3 + a, 3 + b)) + (i == j ? delta : 0.0); }
    if (idx < 34) { int q = idx - 19, i = q / 5, col = q % 5; return col == 0 ? RC(MPCB200_K_G + i) : ((col == 1 && dt_free) ? RC(MPCB200_K_HB + i) : 0.0); }
    if (idx < 44) { int q = idx - 34, i = q / 5, col = q % 5; return col == 0 ? RC(MPCB200_K_G + 3 + i) : ((col == 1 && dt_free) ? RC(MPCB200_K_HB + 3 + i) : 0.0); }
    if (idx < 50) { int q = idx - 44, col = q / 3, l = q % 3; return col == 0 ? RC(MPCB200_K_E + l) : (dt_free ? RC(MPCB200_K_D + l) : 0.0); }
    if (idx < 54) { int q = idx - 50, p = q / 2, v = q % 2; return p == v ? RC(MPCB200_K_C + p) : 0.0; }
    if (idx < 57) return RC(MPCB200_K_A + (idx - 54));
    return RC(MPCB200_K_B + (idx - 57));
#undef RC
}

// terminal value function entry idx (0..74: P 25, PI 25, TH 25) from the record of stage N-1
HD inline double terminal_entry(const Cfg& c, const double* rec, int N, int idx, double delta, double htt, double gt)
{
    const int k = N - 1;
    if (idx < 25)
    {
        int i = idx / 5, j = idx % 5;
        if (i >= 3 || j >= 3) return 0.0;
        if (c.xf_fixed[i] || c.xf_fixed[j]) return 0.0;
        int a = i < j ? i : j, b = i < j ? j : i;
        return rec[(MPCB200_K_H + hidx(a, b)) * N + k] + (i == j ? delta : 0.0);
    }
    if (idx < 50)
    {
        int q = idx - 25, i = q / 5, col = q % 5;
        if (i >= 3) return 0.0;
        if (col == 0) return c.xf_fixed[i] ? 0.0 : rec[(MPCB200_K_G + i) * N + k];
        if (col == 2 + i) return c.xf_fixed[i] ? 1.0 : 0.0;
        return 0.0;
    }
    int q = idx - 50, i = q / 5, j = q % 5;
    if ((i == 0 && j == 1) || (i == 1 && j == 0)) return gt;
    if (i == 1 && j == 1) return htt + delta;
    return 0.0;
}

// Lambda = MM_vv^-1 with the positive-definiteness (inertia) test; returns 0 on failure
HD inline int lambda_from_mmvv(const double* sm, double* lam4)
{
    const double a = sm[R_MMVV], b = 0.5 * (sm[R_MMVV + 1] + sm[R_MMVV + 2]), d = sm[R_MMVV + 3];
    const double det = a * d - b * b;
    if (!(a > 0.0) || !(d > 0.0) || !(det > 1e-14 * a * d)) return 0;
    lam4[0] = d / det; lam4[1] = -b / det; lam4[2] = -b / det; lam4[3] = a / det;
    return 1;
}

// root: y_0 = 0 -> stationarity of 1/2 th' TH th over the active parameters.  pi block (negative definite) first,
// then d(dt) (must leave a positive pivot).  Returns 0 on wrong inertia.
HD inline int root_solve(const Cfg& c, const double* TH, double* th)
{
    th[0] = 1.0; th[1] = th[2] = th[3] = th[4] = 0.0;
    int act[3], na = 0;
    for (int j = 0; j < 3; ++j)
        if (c.xf_fixed[j]) act[na++] = 2 + j;
    double Lm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < na; ++i)
        for (int j = 0; j <= i; ++j)
        {
            double s = -0.5 * (TH[act[i] * 5 + act[j]] + TH[act[j] * 5 + act[i]]);
            for (int l = 0; l < j; ++l) s -= Lm[i][l] * Lm[j][l];
            if (i == j) { if (!(s > 0.0)) return 0; Lm[i][i] = sqrt(s); }
            else Lm[i][j] = s / Lm[j][j];
        }
    double sol0[3] = {0, 0, 0}, sol1[3] = {0, 0, 0};
    for (int pass = 0; pass < 2; ++pass)
    {
        double* sl = pass ? sol1 : sol0;
        double y[3] = {0, 0, 0};
        for (int i = 0; i < na; ++i)
        {
            double t = TH[act[i] * 5 + pass];
            for (int l = 0; l < i; ++l) t -= Lm[i][l] * y[l];
            y[i] = t / Lm[i][i];
        }
        for (int i = na - 1; i >= 0; --i)
        {
            double t = y[i];
            for (int l = i + 1; l < na; ++l) t -= Lm[l][i] * sl[l];
            sl[i] = t / Lm[i][i];
        }
    }
    double ddt = 0.0;
    if (c.variable_dt)
    {
        double htt = TH[1 * 5 + 1], gt = TH[1 * 5 + 0];
        for (int i = 0; i < na; ++i) { htt += TH[1 * 5 + act[i]] * sol1[i]; gt += TH[1 * 5 + act[i]] * sol0[i]; }
        if (!(htt > 0.0)) return 0;
        ddt = -gt / htt;
    }
    th[1] = ddt;
    for (int i = 0; i < na; ++i) th[act[i]] = sol0[i] + sol1[i] * ddt;
    return 1;
}

// forward substitution of stage k: y (in/out, 5), th (5); gains: ric[0..14] Px(k+1) rows, [15..29] PIx(k+1) rows,
// [30..39] KG, [40..49] KT; a3, Bm, e, d of the stage.  Outputs dw (5) and nu+ (3).
HD inline void forward_stage(const double* ric, const double* a3, const double* Bm, const double* e, const double* d, int dt_free,
                             const double* th, double* y, double* dw, double* nup)
{
    double v[2];
    for (int i = 0; i < 2; ++i)
    {
        double s = 0.0;
        for (int j = 0; j < 5; ++j) s -= ric[30 + i * 5 + j] * y[j] + ric[40 + i * 5 + j] * th[j];
        v[i] = s;
    }
    dw[0] = y[0]; dw[1] = y[1]; dw[2] = y[2]; dw[3] = v[0]; dw[4] = v[1];
    double yn[5];
    for (int i = 0; i < 3; ++i)
        yn[i] = y[i] + a3[i] * y[2] + Bm[2 * i] * v[0] + Bm[2 * i + 1] * v[1] + e[i] + (dt_free ? d[i] * th[1] : 0.0);
    yn[3] = v[0]; yn[4] = v[1];
    for (int i = 0; i < 3; ++i)
    {
        double s = 0.0;
        for (int j = 0; j < 5; ++j) s += ric[i * 5 + j] * yn[j] + ric[15 + i * 5 + j] * th[j];
        nup[i] = s;
    }
    for (int i = 0; i < 5; ++i) y[i] = yn[i];
}
