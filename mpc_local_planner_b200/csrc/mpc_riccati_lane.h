// mpc_riccati_lane.h -- register-resident Riccati factorisation + solve of the bordered block-tridiagonal KKT system of
// ONE OCP instance, executed by ONE thread (the replacement of MUMPS' general sparse LDL^T inside Ipopt,
// R/src/controller.cpp:380-421).  The KKT kernel runs it with one lane per instance: a warp sweeps 32 instances in
// lock-step, every load/store is a coalesced 256-byte access into 32-instance interleaved tiles, and there is no
// intra-stage synchronisation at all (the warp-cooperative variant of round-1a spent ~95% of its issue slots on
// shared-memory round trips and __syncwarp; see DESIGN.md "KKT kernel").
//
// Unknowns: dw_k = (dx_k, du_k), nu+_k, d(dt);  dx_{k+1} = A_k dx_k + B_k du_k + d_k d(dt) + e_k,  dx_0 = 0,
// dx_{N-1,j} = 0 for fixed terminal components.  Stage state y_k = (dx_k, du_{k-1}) in R^5 (the previous control
// enters through the control-rate cross block C_k), parameters theta^ = (1, d(dt), pi_0, pi_1, pi_2).
// Value function V_k = 1/2 y'P y + y' PI theta^ + 1/2 theta^' TH theta^   (DESIGN.md "Riccati").
// EXT = false: theta^ = (1) only (fixed dt, free terminal state): NC = 1 column.  EXT = true: NC = 5.
#pragma once
#include "mpc_core.h"

#define TILE 32        // instances per tile (= lanes per warp)
#define RICW_MAX 50    // gains per stage: Px rows (15), PIx rows (3*NC), KG (10), KT (2*NC)

// strided accessors into the 32-instance interleaved tiles: element (k, f) of this lane's instance
struct TileRec
{
    const double* base;  // tile base + lane
    HD double operator()(int k, int f) const { return base[((size_t)k * MPCB200_KKT_WORDS + f) * TILE]; }
};
struct TileRic
{
    double* base;
    HD double& operator()(int k, int w) const { return base[((size_t)k * RICW_MAX + w) * TILE]; }
};

template <bool EXT>
struct RicState
{
    static constexpr int NC = EXT ? 5 : 1;
    double P[5][5];
    double PI[5][NC];
    double TH[NC][NC];
};

// terminal value function from the record of stage N-1
template <bool EXT, class Rec>
HD inline void riccati_terminal(const Cfg& c, const Rec& rec, int N, double delta, double htt, double gt, RicState<EXT>& s)
{
    constexpr int NC = RicState<EXT>::NC;
#pragma unroll
    for (int i = 0; i < 5; ++i)
    {
#pragma unroll
        for (int j = 0; j < 5; ++j) s.P[i][j] = 0.0;
#pragma unroll
        for (int j = 0; j < NC; ++j) s.PI[i][j] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < NC; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) s.TH[i][j] = 0.0;
    const int k = N - 1;
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
#pragma unroll
        for (int j = i; j < 3; ++j)
        {
            double v = rec(k, MPCB200_K_H + hidx(i, j)) + (i == j ? delta : 0.0);
            if (c.xf_fixed[i] || c.xf_fixed[j]) v = 0.0;
            s.P[i][j] = v; s.P[j][i] = v;
        }
        s.PI[i][0] = c.xf_fixed[i] ? 0.0 : rec(k, MPCB200_K_G + i);
        if (EXT && c.xf_fixed[i]) s.PI[i][(EXT ? 2 + i : 0)] = 1.0;
    }
    if (EXT)
    {
        s.TH[0][EXT ? 1 : 0] = gt; s.TH[EXT ? 1 : 0][0] = gt;
        s.TH[EXT ? 1 : 0][EXT ? 1 : 0] = htt + delta;
    }
}

// one backward stage; returns 0 if M_vv is not positive definite.  Writes the gains of stage k.
// load the 42-word record of stage k into registers (one coalesced 256-byte access per word across the warp)
template <class Rec>
HD inline void riccati_load(const Rec& rec, int k, double* r)
{
#pragma unroll
    for (int f = 0; f < MPCB200_KKT_WORDS; ++f) r[f] = rec(k, f);
}

template <bool EXT, class Ric>
HD inline int riccati_stage(const double* r, const Ric& ric, int k, double delta, int dt_free, RicState<EXT>& s)
{
    constexpr int NC = RicState<EXT>::NC;
    // ---- stage record (already in registers) ----
    double H[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = i; j < 5; ++j)
        {
            const double v = r[MPCB200_K_H + hidx(i, j)] + (i == j ? delta : 0.0);
            H[i][j] = v; H[j][i] = v;
        }
    double g[5], a[3], Bm[3][2], e[3], Cc[2], hb[5], dv[3];
#pragma unroll
    for (int i = 0; i < 5; ++i) { g[i] = r[MPCB200_K_G + i]; hb[i] = (EXT && dt_free) ? r[MPCB200_K_HB + i] : 0.0; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        a[i] = r[MPCB200_K_A + i]; e[i] = r[MPCB200_K_E + i];
        dv[i] = (EXT && dt_free) ? r[MPCB200_K_D + i] : 0.0;
        Bm[i][0] = r[MPCB200_K_B + 2 * i]; Bm[i][1] = r[MPCB200_K_B + 2 * i + 1];
    }
    Cc[0] = r[MPCB200_K_C]; Cc[1] = r[MPCB200_K_C + 1];
    // ---- gains needed later: rows x of P_{k+1} and PI_{k+1} (for nu+) ----
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
#pragma unroll
        for (int j = 0; j < 5; ++j) ric(k, i * 5 + j) = s.P[i][j];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) ric(k, 15 + i * NC + cc) = s.PI[i][cc];
    }
    // ---- T1 = P Bbar (5x2), PA = Pxx a (3), W = P chat + PI (5 x NC) ----
    double T1[5][2], PA[3], Wm[5][NC];
#pragma unroll
    for (int i = 0; i < 5; ++i)
    {
#pragma unroll
        for (int j = 0; j < 2; ++j) T1[i][j] = s.P[i][0] * Bm[0][j] + s.P[i][1] * Bm[1][j] + s.P[i][2] * Bm[2][j] + s.P[i][3 + j];
        Wm[i][0] = s.P[i][0] * e[0] + s.P[i][1] * e[1] + s.P[i][2] * e[2] + s.PI[i][0];
        if (EXT)
        {
            Wm[i][EXT ? 1 : 0] = s.P[i][0] * dv[0] + s.P[i][1] * dv[1] + s.P[i][2] * dv[2] + s.PI[i][EXT ? 1 : 0];
#pragma unroll
            for (int cc = 2; cc < NC; ++cc) Wm[i][cc] = s.PI[i][cc];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) PA[i] = s.P[i][0] * a[0] + s.P[i][1] * a[1] + s.P[i][2] * a[2];
    // ---- MM = M + F'PF ----
    double Mvv[2][2], Mxv[3][2], Mxx[3][3];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            Mvv[i][j] = H[3 + i][3 + j] + Bm[0][i] * T1[0][j] + Bm[1][i] * T1[1][j] + Bm[2][i] * T1[2][j] + T1[3 + i][j];
#pragma unroll
    for (int j = 0; j < 2; ++j)
    {
        const double at = a[0] * T1[0][j] + a[1] * T1[1][j] + a[2] * T1[2][j];
        Mxv[0][j] = H[0][3 + j] + T1[0][j];
        Mxv[1][j] = H[1][3 + j] + T1[1][j];
        Mxv[2][j] = H[2][3 + j] + T1[2][j] + at;
    }
    {
        const double apa = a[0] * PA[0] + a[1] * PA[1] + a[2] * PA[2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                Mxx[i][j] = H[i][j] + s.P[i][j] + (i == 2 ? PA[j] : 0.0) + (j == 2 ? PA[i] : 0.0) + ((i == 2 && j == 2) ? apa : 0.0);
    }
    // ---- NN = Mhat + F'W ----
    double Nx[3][NC], Nv[2][NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
    {
        const double aw = a[0] * Wm[0][cc] + a[1] * Wm[1][cc] + a[2] * Wm[2][cc];
        const double mx0 = cc == 0 ? g[0] : (cc == 1 ? hb[0] : 0.0), mx1 = cc == 0 ? g[1] : (cc == 1 ? hb[1] : 0.0);
        const double mx2 = cc == 0 ? g[2] : (cc == 1 ? hb[2] : 0.0);
        const double mv0 = cc == 0 ? g[3] : (cc == 1 ? hb[3] : 0.0), mv1 = cc == 0 ? g[4] : (cc == 1 ? hb[4] : 0.0);
        Nx[0][cc] = mx0 + Wm[0][cc];
        Nx[1][cc] = mx1 + Wm[1][cc];
        Nx[2][cc] = mx2 + Wm[2][cc] + aw;
        Nv[0][cc] = mv0 + Bm[0][0] * Wm[0][cc] + Bm[1][0] * Wm[1][cc] + Bm[2][0] * Wm[2][cc] + Wm[3][cc];
        Nv[1][cc] = mv1 + Bm[0][1] * Wm[0][cc] + Bm[1][1] * Wm[1][cc] + Bm[2][1] * Wm[2][cc] + Wm[4][cc];
    }
    // ---- TT = TH + Chat'W + PI'Chat (rows/cols 0,1 only carry chat) ----
    double TT[NC][NC];
    if (EXT)
    {
#pragma unroll
        for (int i = 0; i < NC; ++i)
#pragma unroll
            for (int j = 0; j < NC; ++j)
            {
                double v = s.TH[i][j];
                if (i < 2)
                {
                    const double c0 = i == 0 ? e[0] : dv[0], c1 = i == 0 ? e[1] : dv[1], c2 = i == 0 ? e[2] : dv[2];
                    v += c0 * Wm[0][j] + c1 * Wm[1][j] + c2 * Wm[2][j];
                }
                if (j < 2)
                {
                    const double c0 = j == 0 ? e[0] : dv[0], c1 = j == 0 ? e[1] : dv[1], c2 = j == 0 ? e[2] : dv[2];
                    v += s.PI[0][i] * c0 + s.PI[1][i] * c1 + s.PI[2][i] * c2;
                }
                TT[i][j] = v;
            }
    }
    // ---- Lambda = Mvv^-1 with the inertia test ----
    const double la = Mvv[0][0], lb = 0.5 * (Mvv[0][1] + Mvv[1][0]), ld = Mvv[1][1];
    const double det = la * ld - lb * lb;
    if (!(la > 0.0) || !(ld > 0.0) || !(det > 1e-14 * la * ld)) return 0;
    const double idet = 1.0 / det;
    const double L00 = ld * idet, L01 = -lb * idet, L11 = la * idet;
    // ---- gains KG = Lambda Mvy (2x5), KT = Lambda Nv (2xNC) ----
    double KG[2][5], KT[2][NC];
#pragma unroll
    for (int j = 0; j < 3; ++j)
    {
        KG[0][j] = L00 * Mxv[j][0] + L01 * Mxv[j][1];
        KG[1][j] = L01 * Mxv[j][0] + L11 * Mxv[j][1];
    }
    KG[0][3] = L00 * Cc[0]; KG[1][3] = L01 * Cc[0];
    KG[0][4] = L01 * Cc[1]; KG[1][4] = L11 * Cc[1];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
    {
        KT[0][cc] = L00 * Nv[0][cc] + L01 * Nv[1][cc];
        KT[1][cc] = L01 * Nv[0][cc] + L11 * Nv[1][cc];
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) { ric(k, 15 + 3 * NC + j) = KG[0][j]; ric(k, 15 + 3 * NC + 5 + j) = KG[1][j]; }
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) { ric(k, 25 + 3 * NC + cc) = KT[0][cc]; ric(k, 25 + 4 * NC + cc) = KT[1][cc]; }
    // ---- Schur complements: P, PI, TH of stage k ----
    // Myv rows: x rows = Mxv, p rows = diag(C)
    double Pn[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = i; j < 5; ++j)
        {
            const double m0 = i < 3 ? Mxv[i < 3 ? i : 0][0] : (i == 3 ? Cc[0] : 0.0);
            const double m1 = i < 3 ? Mxv[i < 3 ? i : 0][1] : (i == 4 ? Cc[1] : 0.0);
            const double z = (i < 3 && j < 3) ? Mxx[i < 3 ? i : 0][j < 3 ? j : 0] : 0.0;
            const double v = z - (m0 * KG[0][j] + m1 * KG[1][j]);
            Pn[i][j] = v; Pn[j][i] = v;
        }
    // symmetrise like the oracle: P_ij <- (P_ij + P_ji)/2 with P_ji computed from the transposed formula
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = i + 1; j < 5; ++j)
        {
            const double m0 = j < 3 ? Mxv[j < 3 ? j : 0][0] : (j == 3 ? Cc[0] : 0.0);
            const double m1 = j < 3 ? Mxv[j < 3 ? j : 0][1] : (j == 4 ? Cc[1] : 0.0);
            const double z = (i < 3 && j < 3) ? Mxx[j < 3 ? j : 0][i < 3 ? i : 0] : 0.0;
            const double vt = z - (m0 * KG[0][i] + m1 * KG[1][i]);
            const double v = 0.5 * (Pn[i][j] + vt);
            Pn[i][j] = v; Pn[j][i] = v;
        }
    double PIn[5][NC];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
        {
            const double m0 = i < 3 ? Mxv[i < 3 ? i : 0][0] : (i == 3 ? Cc[0] : 0.0);
            const double m1 = i < 3 ? Mxv[i < 3 ? i : 0][1] : (i == 4 ? Cc[1] : 0.0);
            const double z = i < 3 ? Nx[i < 3 ? i : 0][cc] : 0.0;
            PIn[i][cc] = z - (m0 * KT[0][cc] + m1 * KT[1][cc]);
        }
    if (EXT)
    {
#pragma unroll
        for (int i = 0; i < NC; ++i)
#pragma unroll
            for (int j = 0; j < NC; ++j) s.TH[i][j] = TT[i][j] - (Nv[0][i] * KT[0][j] + Nv[1][i] * KT[1][j]);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i)
    {
#pragma unroll
        for (int j = 0; j < 5; ++j) s.P[i][j] = Pn[i][j];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) s.PI[i][cc] = PIn[i][cc];
    }
    return 1;
}

// root: y_0 = 0 -> stationarity of 1/2 th' TH th over the active parameters: pi block (negative definite) first,
// then d(dt) (must leave a positive pivot).  Returns 0 on wrong inertia.
HD inline int riccati_root(const Cfg& c, const double TH[5][5], double* th)
{
    th[0] = 1.0; th[1] = th[2] = th[3] = th[4] = 0.0;
    int act[3], na = 0;
    for (int j = 0; j < 3; ++j)
        if (c.xf_fixed[j]) act[na++] = 2 + j;
    double Lm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < na; ++i)
        for (int j = 0; j <= i; ++j)
        {
            double s = -0.5 * (TH[act[i]][act[j]] + TH[act[j]][act[i]]);
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
            double t = TH[act[i]][pass];
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
        double htt = TH[1][1], gt = TH[1][0];
        for (int i = 0; i < na; ++i) { htt += TH[1][act[i]] * sol1[i]; gt += TH[1][act[i]] * sol0[i]; }
        if (!(htt > 0.0)) return 0;
        ddt = -gt / htt;
    }
    th[1] = ddt;
    for (int i = 0; i < na; ++i) th[act[i]] = sol0[i] + sol1[i] * ddt;
    return 1;
}

// Full factorisation + solve of one instance with inertia-correcting regularisation (Ipopt's algorithm IC).
// Step: callable step(k, c, value) receiving dw (c = 0..4) and nu+ (c = 5..7) of stage k.
template <bool EXT, class Rec, class Ric, class Step>
HD inline int riccati_solve_lane(const Cfg& c, int N, const Rec& rec, const Ric& ric, const Step& step, double htt, double gt,
                                 double dlast, double* ddt_out, double* delta_out, int* nreg_out)
{
    constexpr int NC = RicState<EXT>::NC;
    const int dt_free = c.variable_dt;
    RicState<EXT> s;
    double th[5] = {1.0, 0, 0, 0, 0};
    // after a regularised iteration the first attempt is delta_last/3 (decays back to 0): a failed attempt costs a full sweep
    double delta = (dlast > 0.0 && dlast / 3.0 >= DELTA_FLOOR) ? dlast / 3.0 : 0.0;
    int ok = 0, nreg = 0;
    for (int tries = 0; tries < MAX_INERTIA_TRIES; ++tries)
    {
        riccati_terminal<EXT>(c, rec, N, delta, htt, gt, s);
        int good = 1;
        double cur[MPCB200_KKT_WORDS], nxt[MPCB200_KKT_WORDS];
        riccati_load(rec, N - 2, nxt);
        for (int k = N - 2; k >= 0; --k)
        {
#pragma unroll
            for (int f = 0; f < MPCB200_KKT_WORDS; ++f) cur[f] = nxt[f];
            if (k > 0) riccati_load(rec, k - 1, nxt);  // prefetch the next record while this stage is processed
            if (!riccati_stage<EXT>(cur, ric, k, delta, dt_free, s)) { good = 0; break; }
        }
        if (good && EXT)
        {
            double THf[5][5];
            for (int i = 0; i < 5; ++i)
                for (int j = 0; j < 5; ++j) THf[i][j] = s.TH[i < NC ? i : 0][j < NC ? j : 0];
            good = riccati_root(c, THf, th);
        }
        if (good) { ok = 1; break; }
        ++nreg;
        if (delta == 0.0) delta = (dlast == 0.0) ? 1e-4 : fmax(dlast / 3.0, 1e-20);
        else delta *= (dlast == 0.0 ? 100.0 : 8.0);
        if (delta > MAX_DELTA) break;
    }
    *nreg_out = nreg;
    *delta_out = delta;
    if (!ok) return 0;
    // ---- forward substitution (gains + dynamics of stage k+1 are prefetched while stage k is processed) ----
    constexpr int NG = 25 + 5 * NC;  // Px rows 15, PIx rows 3*NC, KG 10, KT 2*NC
    double y[5] = {0, 0, 0, 0, 0};
    double gc[NG], gn[NG], dc[15], dn[15];
    auto load_fwd = [&](int k, double* gg, double* dd) {
#pragma unroll
        for (int w = 0; w < NG; ++w) gg[w] = ric(k, w);
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
            dd[i] = rec(k, MPCB200_K_A + i); dd[3 + 2 * i] = rec(k, MPCB200_K_B + 2 * i); dd[4 + 2 * i] = rec(k, MPCB200_K_B + 2 * i + 1);
            dd[9 + i] = rec(k, MPCB200_K_E + i); dd[12 + i] = (EXT && dt_free) ? rec(k, MPCB200_K_D + i) : 0.0;
        }
    };
    load_fwd(0, gn, dn);
    for (int k = 0; k <= N - 2; ++k)
    {
#pragma unroll
        for (int w = 0; w < NG; ++w) gc[w] = gn[w];
#pragma unroll
        for (int w = 0; w < 15; ++w) dc[w] = dn[w];
        if (k < N - 2) load_fwd(k + 1, gn, dn);
        double v0 = 0.0, v1 = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j)
        {
            v0 -= gc[15 + 3 * NC + j] * y[j];
            v1 -= gc[15 + 3 * NC + 5 + j] * y[j];
        }
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
        {
            v0 -= gc[25 + 3 * NC + cc] * th[cc];
            v1 -= gc[25 + 4 * NC + cc] * th[cc];
        }
        step(k, 0, y[0]); step(k, 1, y[1]); step(k, 2, y[2]); step(k, 3, v0); step(k, 4, v1);
        double yn[5];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            yn[i] = y[i] + dc[i] * y[2] + dc[3 + 2 * i] * v0 + dc[4 + 2 * i] * v1 + dc[9 + i] + ((EXT && dt_free) ? dc[12 + i] * th[1] : 0.0);
        yn[3] = v0; yn[4] = v1;
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
            double sacc = 0.0;
#pragma unroll
            for (int j = 0; j < 5; ++j) sacc += gc[i * 5 + j] * yn[j];
#pragma unroll
            for (int cc = 0; cc < NC; ++cc) sacc += gc[15 + i * NC + cc] * th[cc];
            step(k, 5 + i, sacc);
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) y[i] = yn[i];
    }
    step(N - 1, 0, y[0]); step(N - 1, 1, y[1]); step(N - 1, 2, y[2]);
    for (int cc = 3; cc < 8; ++cc) step(N - 1, cc, 0.0);
    *ddt_out = th[1];
    return 1;
}
