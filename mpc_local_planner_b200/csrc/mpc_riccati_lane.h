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
        // x_{N-1} - dt cross term: end term of the trapezoidal cost rule, or the heading of the last midpoint defect
        if (EXT && c.variable_dt && (has_trapezoid(c) || is_midpoint(c)) && !c.xf_fixed[i]) s.PI[i][EXT ? 1 : 0] = rec(k, MPCB200_K_HB + i);
    }
    if (EXT)
    {
        s.TH[0][EXT ? 1 : 0] = gt; s.TH[EXT ? 1 : 0][0] = gt;
        s.TH[EXT ? 1 : 0][EXT ? 1 : 0] = htt + delta;
    }
}

// Record views: r(f) = word f of the stage record of this lane's instance.  PlainFeed reads the interleaved tile in
// global memory directly (host emulator); the CUDA kernel reads the copy a bulk-async (TMA) transfer staged in shared memory.
struct StridedView
{
    const double* p;  // word 0 of this lane's record, words TILE apart
    HD double operator()(int f) const { return p[(size_t)f * TILE]; }
};

// one backward stage; returns 0 if M_vv is not positive definite.  Writes the gains of stage k.
// Structure exploited: A = I + a e_theta', the p (= previous control) rows of M_yv are diag(C), M_pp = 0.
template <bool EXT, class RV, class Ric>
HD inline int riccati_stage(const RV& r, const Ric& ric, int k, double delta, int dt_free, RicState<EXT>& s)
{
    constexpr int NC = RicState<EXT>::NC;
    const bool dtf = EXT && dt_free;
    // ---- gains needed later: rows x of P_{k+1} and PI_{k+1} (for nu+) ----
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
#pragma unroll
        for (int j = 0; j < 5; ++j) ric(k, i * 5 + j) = s.P[i][j];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) ric(k, 15 + i * NC + cc) = s.PI[i][cc];
    }
    double a[3], Bm[3][2], e[3], dv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        a[i] = r(MPCB200_K_A + i); e[i] = r(MPCB200_K_E + i);
        dv[i] = dtf ? r(MPCB200_K_D + i) : 0.0;
        Bm[i][0] = r(MPCB200_K_B + 2 * i); Bm[i][1] = r(MPCB200_K_B + 2 * i + 1);
    }
    const double Cc0 = r(MPCB200_K_C), Cc1 = r(MPCB200_K_C + 1);
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
            Wm[i][EXT ? 1 : 0] = dtf ? s.P[i][0] * dv[0] + s.P[i][1] * dv[1] + s.P[i][2] * dv[2] + s.PI[i][EXT ? 1 : 0] : s.PI[i][EXT ? 1 : 0];
#pragma unroll
            for (int cc = 2; cc < NC; ++cc) Wm[i][cc] = s.PI[i][cc];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) PA[i] = s.P[i][0] * a[0] + s.P[i][1] * a[1] + s.P[i][2] * a[2];
    // ---- MM = M + F'PF ----
    double Mvv[3], Mxv[3][2], Mxx[3][3];  // Mvv packed (00, 01, 11)
    {
        const double h33 = r(MPCB200_K_H + hidx(3, 3)) + delta, h34 = r(MPCB200_K_H + hidx(3, 4)), h44 = r(MPCB200_K_H + hidx(4, 4)) + delta;
        Mvv[0] = h33 + Bm[0][0] * T1[0][0] + Bm[1][0] * T1[1][0] + Bm[2][0] * T1[2][0] + T1[3][0];
        Mvv[2] = h44 + Bm[0][1] * T1[0][1] + Bm[1][1] * T1[1][1] + Bm[2][1] * T1[2][1] + T1[4][1];
        const double m01 = h34 + Bm[0][0] * T1[0][1] + Bm[1][0] * T1[1][1] + Bm[2][0] * T1[2][1] + T1[3][1];
        const double m10 = h34 + Bm[0][1] * T1[0][0] + Bm[1][1] * T1[1][0] + Bm[2][1] * T1[2][0] + T1[4][0];
        Mvv[1] = 0.5 * (m01 + m10);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
    {
        const double at = a[0] * T1[0][j] + a[1] * T1[1][j] + a[2] * T1[2][j];
        Mxv[0][j] = r(MPCB200_K_H + hidx(0, 3 + j)) + T1[0][j];
        Mxv[1][j] = r(MPCB200_K_H + hidx(1, 3 + j)) + T1[1][j];
        Mxv[2][j] = r(MPCB200_K_H + hidx(2, 3 + j)) + T1[2][j] + at;
    }
    {
        const double apa = a[0] * PA[0] + a[1] * PA[1] + a[2] * PA[2];
        Mxx[0][0] = r(MPCB200_K_H + hidx(0, 0)) + delta + s.P[0][0];
        Mxx[0][1] = r(MPCB200_K_H + hidx(0, 1)) + s.P[0][1];
        Mxx[1][1] = r(MPCB200_K_H + hidx(1, 1)) + delta + s.P[1][1];
        Mxx[0][2] = r(MPCB200_K_H + hidx(0, 2)) + s.P[0][2] + PA[0];
        Mxx[1][2] = r(MPCB200_K_H + hidx(1, 2)) + s.P[1][2] + PA[1];
        Mxx[2][2] = r(MPCB200_K_H + hidx(2, 2)) + delta + s.P[2][2] + (PA[2] + PA[2]) + apa;
    }
    // ---- NN = Mhat + F'W ----
    double Nx[3][NC], Nv[2][NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
    {
        const double aw = a[0] * Wm[0][cc] + a[1] * Wm[1][cc] + a[2] * Wm[2][cc];
        const double bw0 = Bm[0][0] * Wm[0][cc] + Bm[1][0] * Wm[1][cc] + Bm[2][0] * Wm[2][cc] + Wm[3][cc];
        const double bw1 = Bm[0][1] * Wm[0][cc] + Bm[1][1] * Wm[1][cc] + Bm[2][1] * Wm[2][cc] + Wm[4][cc];
        if (cc == 0)
        {
            Nx[0][cc] = r(MPCB200_K_G + 0) + Wm[0][cc]; Nx[1][cc] = r(MPCB200_K_G + 1) + Wm[1][cc]; Nx[2][cc] = r(MPCB200_K_G + 2) + Wm[2][cc] + aw;
            Nv[0][cc] = r(MPCB200_K_G + 3) + bw0; Nv[1][cc] = r(MPCB200_K_G + 4) + bw1;
        }
        else if (cc == 1 && dtf)
        {
            Nx[0][cc] = r(MPCB200_K_HB + 0) + Wm[0][cc]; Nx[1][cc] = r(MPCB200_K_HB + 1) + Wm[1][cc]; Nx[2][cc] = r(MPCB200_K_HB + 2) + Wm[2][cc] + aw;
            Nv[0][cc] = r(MPCB200_K_HB + 3) + bw0; Nv[1][cc] = r(MPCB200_K_HB + 4) + bw1;
        }
        else
        {
            Nx[0][cc] = Wm[0][cc]; Nx[1][cc] = Wm[1][cc]; Nx[2][cc] = Wm[2][cc] + aw;
            Nv[0][cc] = bw0; Nv[1][cc] = bw1;
        }
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
                if (i == 0) v += e[0] * Wm[0][j] + e[1] * Wm[1][j] + e[2] * Wm[2][j];
                if (i == 1) v += dv[0] * Wm[0][j] + dv[1] * Wm[1][j] + dv[2] * Wm[2][j];
                if (j == 0) v += s.PI[0][i] * e[0] + s.PI[1][i] * e[1] + s.PI[2][i] * e[2];
                if (j == 1) v += s.PI[0][i] * dv[0] + s.PI[1][i] * dv[1] + s.PI[2][i] * dv[2];
                TT[i][j] = v;
            }
    }
    // ---- Lambda = Mvv^-1 with the inertia test ----
    const double la = Mvv[0], lb = Mvv[1], ld = Mvv[2];
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
    KG[0][3] = L00 * Cc0; KG[1][3] = L01 * Cc0;
    KG[0][4] = L01 * Cc1; KG[1][4] = L11 * Cc1;
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
    // ---- Schur complements (upper triangle, mirrored): P = Myy - Myv KG, PI = Ny - Myv KT, TH = TT - Nv' KT ----
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
#pragma unroll
        for (int j = i; j < 3; ++j)
        {
            const double v = Mxx[i][j] - (Mxv[i][0] * KG[0][j] + Mxv[i][1] * KG[1][j]);
            s.P[i][j] = v; s.P[j][i] = v;
        }
#pragma unroll
        for (int j = 3; j < 5; ++j)
        {
            const double v = -(Mxv[i][0] * KG[0][j] + Mxv[i][1] * KG[1][j]);
            s.P[i][j] = v; s.P[j][i] = v;
        }
    }
    s.P[3][3] = -(Cc0 * KG[0][3]);
    s.P[3][4] = -(Cc0 * KG[0][4]); s.P[4][3] = s.P[3][4];
    s.P[4][4] = -(Cc1 * KG[1][4]);
    if (EXT)
    {
#pragma unroll
        for (int i = 0; i < NC; ++i)
#pragma unroll
            for (int j = 0; j < NC; ++j) s.TH[i][j] = TT[i][j] - (Nv[0][i] * KT[0][j] + Nv[1][i] * KT[1][j]);
    }
#pragma unroll
    for (int cc = 0; cc < NC; ++cc)
    {
#pragma unroll
        for (int i = 0; i < 3; ++i) s.PI[i][cc] = Nx[i][cc] - (Mxv[i][0] * KT[0][cc] + Mxv[i][1] * KT[1][cc]);
        s.PI[3][cc] = -(Cc0 * KT[0][cc]);
        s.PI[4][cc] = -(Cc1 * KT[1][cc]);
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

// ---- feeds: how the stage records / gains reach the lane ---------------------------------------------------------
// A feed hands out the record of backward stage k (all 42 words) and the data of forward stage k (the NG gain words the
// backward sweep wrote + record words A,B,E (20..31) and D (39..41)).  PlainFeed indexes the tiles in place (host
// emulator, tests).  The CUDA kernel's TmaFeed (mpcb200.cu) streams them through a shared-memory ring with bulk-async
// copies so that no lane ever waits on a global load.  All feed calls are warp-uniform.
struct FwdViewPlain
{
    const double* g;  // gains of stage k (word 0, this lane)
    const double* r;  // record of stage k (word 0, this lane)
    HD double gain(int w) const { return g[(size_t)w * TILE]; }
    HD double rec(int f) const { return r[(size_t)f * TILE]; }
};
struct PlainFeed
{
    TileRec rec;
    TileRic ric;
    HD bool any(bool p) const { return p; }
    HD void bwd_start(int) {}
    HD StridedView bwd_acquire(int k) const { return StridedView{rec.base + (size_t)k * MPCB200_KKT_WORDS * TILE}; }
    HD void bwd_release(int) {}
    HD void bwd_abort() {}
    HD void fwd_start(int) {}
    HD FwdViewPlain fwd_acquire(int k) const { return FwdViewPlain{ric.base + (size_t)k * RICW_MAX * TILE, rec.base + (size_t)k * MPCB200_KKT_WORDS * TILE}; }
    HD void fwd_release(int) {}
};

// Regularisation schedule of one IPM iteration (Ipopt's algorithm IC, at most MAX_INERTIA_TRIES attempts per iteration):
// attempt 0 uses kkt_first_delta(), attempt t+1 uses kkt_escalate(delta_t).
HD inline double kkt_first_delta(double dlast) { return (dlast > 0.0 && dlast / 3.0 >= DELTA_FLOOR) ? dlast / 3.0 : 0.0; }
HD inline double kkt_escalate(double delta, double dlast)
{
    if (delta == 0.0) return (dlast == 0.0) ? 1e-4 : fmax(dlast / 3.0, 1e-20);
    return delta * (dlast == 0.0 ? 100.0 : 8.0);
}
// Outcome of the two attempts of an iteration when they were run speculatively side by side (ok_t, delta_t of attempt
// t): identical to running them one after the other.  Returns the winning attempt (0/1), -1 = null step (escalation
// resumes next iteration at *delta_next), -2 = give up.  *nreg = failed attempts.
HD inline int kkt_resolve(int ok0, int ok1, double d0, double d1, double dlast, int* nreg, double* delta_next)
{
    *delta_next = 0.0;
    if (ok0) { *nreg = 0; return 0; }
    *nreg = 1;
    if (d1 > MAX_DELTA) return -2;
    if (ok1) return 1;
    *nreg = 2;
    const double d2 = kkt_escalate(d1, dlast);
    if (d2 > MAX_DELTA) return -2;
    *delta_next = d2;
    return -1;
}

// Full factorisation + solve of the instances of one tile, one lane each, with inertia-correcting regularisation
// (Ipopt's algorithm IC) limited to MAX_INERTIA_TRIES sweeps.  Control flow is warp-uniform: a lane that does not
// need (another) sweep idles through it.  Step: callable step(k, c, value) receiving dw (c = 0..4) and nu+ (c = 5..7).
// Attempts try_first .. try_first + try_count - 1 of the iteration's schedule are run: (0, MAX_INERTIA_TRIES) is the
// plain serial scheme; (t, 1) is attempt t alone (run side by side with the others on small batches, kkt_resolve()).
// Returns 1 if this lane's system was solved; *delta_out is the regularisation used (serial: the next one after a failure).
template <bool EXT, class Feed, class Step>
HD inline int riccati_solve_lane(const Cfg& c, int N, Feed& feed, const Step& step, bool active, double htt, double gt, double dlast,
                                 int try_first, int try_count, double* ddt_out, double* delta_out, int* nreg_out)
{
    constexpr int NC = RicState<EXT>::NC;
    const int dt_free = c.variable_dt;
    RicState<EXT> s;
    double th[5] = {1.0, 0, 0, 0, 0};
    // after a regularised iteration the first attempt is delta_last/3 (decays back to 0): a failed attempt costs a full sweep
    double delta = kkt_first_delta(dlast);
    int ok = 0, nreg = 0;
    bool need = active;
    for (int t = 0; t < try_first; ++t) delta = kkt_escalate(delta, dlast);  // speculative later attempt: its delta is known up front
    if (delta > MAX_DELTA) need = false;
    for (int tries = 0; tries < try_count; ++tries)
    {
        if (!feed.any(need)) break;
        bool alive = need;
        if (alive) riccati_terminal<EXT>(c, feed.rec, N, delta, htt, gt, s);
        feed.bwd_start(N - 2);
        for (int k = N - 2; k >= 0; --k)
        {
            const auto r = feed.bwd_acquire(k);
            if (alive) alive = riccati_stage<EXT>(r, feed.ric, k, delta, dt_free, s) != 0;
            feed.bwd_release(k);
            if (!feed.any(alive)) { feed.bwd_abort(); break; }
        }
        if (alive && EXT)
        {
            double THf[5][5];
            for (int i = 0; i < 5; ++i)
                for (int j = 0; j < 5; ++j) THf[i][j] = s.TH[i < NC ? i : 0][j < NC ? j : 0];
            alive = riccati_root(c, THf, th) != 0;
        }
        if (need)
        {
            if (alive) { ok = 1; need = false; }
            else
            {
                ++nreg;
                if (try_count > 1)  // serial attempts: escalate for the next one (a lone speculative attempt reports the delta it used)
                {
                    delta = kkt_escalate(delta, dlast);
                    if (delta > MAX_DELTA) need = false;
                }
            }
        }
    }
    *nreg_out = nreg;
    *delta_out = delta;
    *ddt_out = th[1];
    // ---- forward substitution ----
    if (!feed.any(ok != 0)) return 0;
    double y[5] = {0, 0, 0, 0, 0};
    feed.fwd_start(N);
    for (int k = 0; k <= N - 2; ++k)
    {
        const auto f = feed.fwd_acquire(k);
        if (ok)
        {
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int j = 0; j < 5; ++j)
            {
                v0 -= f.gain(15 + 3 * NC + j) * y[j];
                v1 -= f.gain(15 + 3 * NC + 5 + j) * y[j];
            }
#pragma unroll
            for (int cc = 0; cc < NC; ++cc)
            {
                v0 -= f.gain(25 + 3 * NC + cc) * th[cc];
                v1 -= f.gain(25 + 4 * NC + cc) * th[cc];
            }
            step(k, 0, y[0]); step(k, 1, y[1]); step(k, 2, y[2]); step(k, 3, v0); step(k, 4, v1);
            double yn[5];
#pragma unroll
            for (int i = 0; i < 3; ++i)
            {
                double t = y[i] + f.rec(MPCB200_K_A + i) * y[2] + f.rec(MPCB200_K_B + 2 * i) * v0 + f.rec(MPCB200_K_B + 2 * i + 1) * v1 + f.rec(MPCB200_K_E + i);
                if (EXT && dt_free) t += f.rec(MPCB200_K_D + i) * th[1];
                yn[i] = t;
            }
            yn[3] = v0; yn[4] = v1;
#pragma unroll
            for (int i = 0; i < 3; ++i)
            {
                double sacc = 0.0;
#pragma unroll
                for (int j = 0; j < 5; ++j) sacc += f.gain(i * 5 + j) * yn[j];
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) sacc += f.gain(15 + i * NC + cc) * th[cc];
                step(k, 5 + i, sacc);
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) y[i] = yn[i];
        }
        feed.fwd_release(k);
    }
    if (ok)
    {
        step(N - 1, 0, y[0]); step(N - 1, 1, y[1]); step(N - 1, 2, y[2]);
        for (int cc = 3; cc < 8; ++cc) step(N - 1, cc, 0.0);
    }
    return ok;
}
