/*
 * mpc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code; see mpc_oracle.h for the parity statement).
 *
 * Part 1 restates the reference's OCP functions one by one (each cites the reference file:line it follows;
 * R/ = /root/reference/mpc_local_planner/).  Part 2 is a serial primal-dual interior-point solver on those
 * functions (slack form, l1-merit line search, inertia-correcting regularisation, stage-wise Riccati solve of the
 * bordered block-tridiagonal KKT system) -- the stand-in for control_box_rst + Ipopt/MUMPS (R/src/controller.cpp:380-421),
 * which are not in /root/reference.
 *
 * Formulation notes (same in the device code):
 *   - dynamics defects and control-rate rows are used in their dt-multiplied form
 *         e_k = x_k + dt f(x_k,u_k) - x_{k+1}            ( = dt * reference defect, R/include/.../fd_collocation_se2.h:54-69)
 *         du_lb*dt <= u_k - u_{k-1} <= du_ub*dt          ( = dt * reference rows, R/src/optimal_control/stage_inequality_se2.cpp:191-222)
 *     which has the same feasible set and optimum for dt > 0 but is linear in dt (hard part 1).
 *   - theta is treated as an unconstrained real during the iterations and wrapped on output; every path function is
 *     2 pi periodic in theta.
 */
#include "mpc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define IX(c, k) ((c) * N + (k))
#define KW MPCB200_KKT_WORDS

/* ======================================================================================================= */
/* Part 1: reference restatement                                                                            */
/* ======================================================================================================= */

/* normalize_theta: R/include/mpc_local_planner/utils/math_utils.h:81-91 -> [-pi, pi) */
double orc_normalize_theta(double theta)
{
    if (theta >= -M_PI && theta < M_PI) return theta;
    double multiplier = floor(theta / (2.0 * M_PI));
    theta = theta - multiplier * 2.0 * M_PI;
    if (theta >= M_PI) theta -= 2.0 * M_PI;
    if (theta < -M_PI) theta += 2.0 * M_PI;
    return theta;
}

/* interpolate_angle: R/include/mpc_local_planner/utils/math_utils.h:100-103 */
double orc_interpolate_angle(double a1, double a2, double factor)
{
    return orc_normalize_theta(a1 + factor * orc_normalize_theta(a2 - a1));
}

/*
 * Robot dynamics xdot = f(x,u):
 *   unicycle            R/include/mpc_local_planner/systems/unicycle_robot.h:59-68
 *   simple car (rear)   R/include/mpc_local_planner/systems/simple_car.h:68-77
 *   simple car (front)  R/include/mpc_local_planner/systems/simple_car.h:131-141
 *   kinematic bicycle   R/include/mpc_local_planner/systems/kinematic_bicycle_model.h:65-77
 */
void orc_dynamics(const mpcb200_config* cfg, const double* x, const double* u, double* f)
{
    switch (cfg->robot_type)
    {
        case MPCB200_ROBOT_UNICYCLE:
            f[0] = u[0] * cos(x[2]);
            f[1] = u[0] * sin(x[2]);
            f[2] = u[1];
            break;
        case MPCB200_ROBOT_SIMPLE_CAR:
            f[0] = u[0] * cos(x[2]);
            f[1] = u[0] * sin(x[2]);
            f[2] = u[0] * tan(u[1]) / cfg->wheelbase;
            break;
        case MPCB200_ROBOT_SIMPLE_CAR_FRONT:
            f[0] = u[0] * cos(x[2]);
            f[1] = u[0] * sin(x[2]);
            f[2] = u[0] * sin(u[1]) / cfg->wheelbase;
            break;
        default:
        {
            double beta = atan(cfg->length_rear / (cfg->length_front + cfg->length_rear) * tan(u[1]));
            f[0] = u[0] * cos(x[2] + beta);
            f[1] = u[0] * sin(x[2] + beta);
            f[2] = u[0] * sin(beta) / cfg->length_rear;
        }
    }
}

/*
 * Analytic first/second derivatives of f wrt q = (theta, u0, u1); replaces corbo's numeric
 * differencing.  J[j*3+i] = d f_j / d q_i.  Hc packs sum_j nu_j * d2 f_j / dq dq as (tt, t0, t1, 00, 01, 11).
 */
void orc_dynamics_derivs(const mpcb200_config* cfg, const double* x, const double* u, const double* nu, double* f,
                         double J[9], double Hc[6])
{
    const double th = x[2], v = u[0], w = u[1];
    for (int i = 0; i < 9; ++i) J[i] = 0.0;
    for (int i = 0; i < 6; ++i) Hc[i] = 0.0;
    if (cfg->robot_type != MPCB200_ROBOT_KIN_BICYCLE)
    {
        const double c = cos(th), s = sin(th);
        f[0] = v * c;
        f[1] = v * s;
        J[0] = -v * s; J[1] = c;
        J[3] = v * c;  J[4] = s;
        /* f0: tt -v c, t0 -s ; f1: tt -v s, t0 c */
        Hc[0] = nu[0] * (-v * c) + nu[1] * (-v * s);
        Hc[1] = nu[0] * (-s) + nu[1] * c;
        if (cfg->robot_type == MPCB200_ROBOT_UNICYCLE)
        {
            f[2] = w;
            J[8] = 1.0;
        }
        else if (cfg->robot_type == MPCB200_ROBOT_SIMPLE_CAR)
        {
            const double L = cfg->wheelbase, t = tan(w), sec2 = 1.0 + t * t;
            f[2] = v * t / L;
            J[7] = t / L;
            J[8] = v * sec2 / L;
            Hc[4] += nu[2] * sec2 / L;               /* d2/dv dphi */
            Hc[5] += nu[2] * 2.0 * v * t * sec2 / L; /* d2/dphi2   */
        }
        else
        {
            const double L = cfg->wheelbase, sp = sin(w), cp = cos(w);
            f[2] = v * sp / L;
            J[7] = sp / L;
            J[8] = v * cp / L;
            Hc[4] += nu[2] * cp / L;
            Hc[5] += nu[2] * (-v * sp / L);
        }
    }
    else
    {
        const double lr = cfg->length_rear, kap = lr / (cfg->length_front + lr);
        const double t = tan(w), sec2 = 1.0 + t * t, den = 1.0 + kap * kap * t * t;
        const double beta = atan(kap * t);
        const double b1 = kap * sec2 / den;                                          /* beta'  */
        const double b2 = 2.0 * kap * t * (1.0 - kap * kap) * sec2 / (den * den);     /* beta'' */
        const double c = cos(th + beta), s = sin(th + beta), sb = sin(beta), cb = cos(beta);
        f[0] = v * c; f[1] = v * s; f[2] = v * sb / lr;
        J[0] = -v * s; J[1] = c;  J[2] = -v * s * b1;
        J[3] = v * c;  J[4] = s;  J[5] = v * c * b1;
        J[6] = 0.0;    J[7] = sb / lr; J[8] = v * cb * b1 / lr;
        /* f0 */
        Hc[0] += nu[0] * (-v * c);
        Hc[1] += nu[0] * (-s);
        Hc[2] += nu[0] * (-v * c * b1);
        Hc[4] += nu[0] * (-s * b1);
        Hc[5] += nu[0] * (-v * c * b1 * b1 - v * s * b2);
        /* f1 */
        Hc[0] += nu[1] * (-v * s);
        Hc[1] += nu[1] * (c);
        Hc[2] += nu[1] * (-v * s * b1);
        Hc[4] += nu[1] * (c * b1);
        Hc[5] += nu[1] * (-v * s * b1 * b1 + v * c * b2);
        /* f2 */
        Hc[4] += nu[2] * (cb * b1 / lr);
        Hc[5] += nu[2] * (v * (-sb * b1 * b1 + cb * b2) / lr);
    }
}

/* pose at which the collocation scheme evaluates the dynamics:
 *   forward differences   x_k                                   [R/include/mpc_local_planner/optimal_control/fd_collocation_se2.h:54-69]
 *   midpoint differences  (x_k + x_{k+1})/2 with the heading interpolate_angle(theta_k, theta_{k+1}, 0.5)      [:91-108]
 * (only the heading matters: none of the models' dynamics depend on the position) */
static int is_midpoint(const mpcb200_config* c) { return c->collocation == MPCB200_COLLOC_MIDPOINT; }
static void collocation_pose(const mpcb200_config* cfg, const double* x1, const double* x2, double* xe)
{
    xe[0] = x1[0]; xe[1] = x1[1]; xe[2] = x1[2];
    if (is_midpoint(cfg))
    {
        xe[0] = 0.5 * (x1[0] + x2[0]); xe[1] = 0.5 * (x1[1] + x2[1]);
        xe[2] = x1[2] + 0.5 * orc_normalize_theta(x2[2] - x1[2]); /* = interpolate_angle up to a multiple of 2 pi */
    }
}

/* {Forward,Midpoint}DiffCollocationSE2::computeEqualityConstraint as coded: fd_collocation_se2.h:54-69, 91-108 */
void orc_defect_reference(const mpcb200_config* cfg, const double* x1, const double* u1, const double* x2, double dt,
                          double* e)
{
    double xe[3];
    collocation_pose(cfg, x1, x2, xe);
    orc_dynamics(cfg, xe, u1, e);
    e[0] -= (x2[0] - x1[0]) / dt;
    e[1] -= (x2[1] - x1[1]) / dt;
    e[2] -= orc_normalize_theta(x2[2] - x1[2]) / dt;
}

/* dt-multiplied form used by the solvers: e = dt * (reference defect) */
void orc_defect(const mpcb200_config* cfg, const double* x1, const double* u1, const double* x2, double dt, double* e)
{
    double f[3], xe[3];
    collocation_pose(cfg, x1, x2, xe);
    orc_dynamics(cfg, xe, u1, f);
    e[0] = dt * f[0] - (x2[0] - x1[0]);
    e[1] = dt * f[1] - (x2[1] - x1[1]);
    e[2] = dt * f[2] - orc_normalize_theta(x2[2] - x1[2]);
}

/* ---- footprint / obstacle geometry (teb_local_planner semantics) ---------------------- */

typedef struct { double ax, ay, bx, by, rad; } fp_seg;

/* footprint -> list of robot-frame segments with an offset radius (point = degenerate segment) */
static int footprint_segments(const mpcb200_config* cfg, fp_seg* seg)
{
    switch (cfg->footprint_type)
    {
        case MPCB200_FOOTPRINT_POINT:
            seg[0] = (fp_seg){0, 0, 0, 0, 0};
            return 1;
        case MPCB200_FOOTPRINT_CIRCULAR:
            seg[0] = (fp_seg){0, 0, 0, 0, cfg->footprint_params[0]};
            return 1;
        case MPCB200_FOOTPRINT_TWO_CIRCLES:
            seg[0] = (fp_seg){cfg->footprint_params[0], 0, cfg->footprint_params[0], 0, cfg->footprint_params[1]};
            seg[1] = (fp_seg){-cfg->footprint_params[2], 0, -cfg->footprint_params[2], 0, cfg->footprint_params[3]};
            return 2;
        case MPCB200_FOOTPRINT_LINE:
            seg[0] = (fp_seg){cfg->footprint_params[0], cfg->footprint_params[1], cfg->footprint_params[2],
                              cfg->footprint_params[3], 0};
            return 1;
        default:
        {
            int n = cfg->n_poly;
            if (n == 1)
            {
                seg[0] = (fp_seg){cfg->poly_xy[0], cfg->poly_xy[1], cfg->poly_xy[0], cfg->poly_xy[1], 0};
                return 1;
            }
            if (n == 2)
            {
                seg[0] = (fp_seg){cfg->poly_xy[0], cfg->poly_xy[1], cfg->poly_xy[2], cfg->poly_xy[3], 0};
                return 1;
            }
            for (int i = 0; i < n; ++i)
            {
                int j = (i + 1) % n; /* closing edge included; no interior test (distance to the boundary) */
                seg[i] = (fp_seg){cfg->poly_xy[2 * i], cfg->poly_xy[2 * i + 1], cfg->poly_xy[2 * j], cfg->poly_xy[2 * j + 1], 0};
            }
            return n;
        }
    }
}

/* distance from q to segment (a,b) [teb distance_point_to_segment_2d: project, clamp to [0,1]]; returns closest point */
static double point_segment(double qx, double qy, const fp_seg* sg, double* cx, double* cy, int* is_vertex)
{
    double dx = sg->bx - sg->ax, dy = sg->by - sg->ay;
    double sq = dx * dx + dy * dy;
    double t = 0.0;
    if (sq > 0.0) t = ((qx - sg->ax) * dx + (qy - sg->ay) * dy) / sq;
    *is_vertex = 0;
    if (!(sq > 0.0) || t <= 0.0) { t = 0.0; *is_vertex = 1; }
    else if (t >= 1.0) { t = 1.0; *is_vertex = 1; }
    *cx = sg->ax + t * dx;
    *cy = sg->ay + t * dy;
    double ex = qx - *cx, ey = qy - *cy;
    return sqrt(ex * ex + ey * ey);
}

/*
 * RobotFootprintModel::calculateDistance(pose, obstacle) for point / circular obstacles
 * (used at R/src/optimal_control/stage_inequality_se2.cpp:109,173).  The obstacle centre is mapped into the robot
 * frame, q = R(theta)^T (o - p); dist = min_i (dist(q, seg_i) - rad_i) - r_obst.  Derivatives are those of the
 * active (arg-min) feature; hess6 = (xx, xy, xt, yy, yt, tt).
 */
static double footprint_distance_line(const mpcb200_config* cfg, const double* pose, const double* op, double* grad3, double* hess6);
double orc_footprint_distance(const mpcb200_config* cfg, const double* pose, int obst_type, const double* op,
                              double* grad3, double* hess6)
{
    if (obst_type == MPCB200_OBST_LINE) return footprint_distance_line(cfg, pose, op, grad3, hess6);
    fp_seg seg[MPCB200_MAX_POLY + 2];
    int ns = footprint_segments(cfg, seg);
    const double c = cos(pose[2]), s = sin(pose[2]);
    const double ox = op[0] - pose[0], oy = op[1] - pose[1];
    const double qx = c * ox + s * oy, qy = -s * ox + c * oy;
    double best = 1e300, bcx = 0, bcy = 0, brho = 0;
    int bvert = 1;
    for (int i = 0; i < ns; ++i)
    {
        double cx, cy;
        int isv;
        double rho = point_segment(qx, qy, &seg[i], &cx, &cy, &isv);
        double d = rho - seg[i].rad;
        if (d < best) { best = d; bcx = cx; bcy = cy; brho = rho; bvert = isv; }
    }
    double r_obst = (obst_type == MPCB200_OBST_CIRCLE) ? op[4] : 0.0;
    double dist = best - r_obst;
    if (grad3 || hess6)
    {
        double rho = brho > 1e-12 ? brho : 1e-12;
        double nx = (qx - bcx) / rho, ny = (qy - bcy) / rho; /* grad of phi wrt q */
        /* dq/dp = -R^T ; dq/dtheta = (qy, -qx) */
        double Jq[2][3] = {{-c, -s, qy}, {s, -c, -qx}};
        if (grad3)
            for (int i = 0; i < 3; ++i) grad3[i] = nx * Jq[0][i] + ny * Jq[1][i];
        if (hess6)
        {
            /* Hphi = (I - n n^T)/rho for vertex features, 0 for edge-interior features */
            double h00 = 0, h01 = 0, h11 = 0;
            if (bvert) { h00 = (1 - nx * nx) / rho; h01 = -nx * ny / rho; h11 = (1 - ny * ny) / rho; }
            double H[3][3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    H[i][j] = Jq[0][i] * (h00 * Jq[0][j] + h01 * Jq[1][j]) + Jq[1][i] * (h01 * Jq[0][j] + h11 * Jq[1][j]);
            /* + sum_i n_i d2 q_i: d2q/dp dtheta = J R^T (J=[[0,-1],[1,0]]): columns; d2q/dtheta2 = -q */
            /* dq/dp_x = (-c, s) -> d/dtheta = (s, c);  dq/dp_y = (-s, -c) -> d/dtheta = (-c, s) */
            double mx = nx * s + ny * c;   /* n . d2q/dpx dtheta */
            double my = -nx * c + ny * s;  /* n . d2q/dpy dtheta */
            H[0][2] += mx; H[2][0] += mx;
            H[1][2] += my; H[2][1] += my;
            H[2][2] += -(nx * qx + ny * qy);
            hess6[0] = H[0][0]; hess6[1] = H[0][1]; hess6[2] = H[0][2];
            hess6[3] = H[1][1]; hess6[4] = H[1][2]; hess6[5] = H[2][2];
        }
    }
    return dist;
}

/* teb check_line_segments_intersection_2d: do (p1,p2) and (p3,p4) cross (touching counts)? */
static int seg_intersect(const double* p1, const double* p2, const double* p3, const double* p4)
{
    const double d1x = p2[0] - p1[0], d1y = p2[1] - p1[1], d2x = p4[0] - p3[0], d2y = p4[1] - p3[1];
    const double den = d1x * d2y - d1y * d2x;
    if (fabs(den) < 1e-14) return 0;
    const double rx = p3[0] - p1[0], ry = p3[1] - p1[1];
    const double t = (rx * d2y - ry * d2x) / den, u = (rx * d1y - ry * d1x) / den;
    return t >= 0.0 && t <= 1.0 && u >= 0.0 && u <= 1.0;
}

/*
 * calculateDistance(pose, LineObstacle(a,b)), evaluated in the WORLD frame: every footprint feature
 * (segment or circle centre) is moved to the world, its distance to the obstacle segment follows teb's
 * distance_segment_to_segment_2d (0 if the segments cross, else the minimum of the four end-point/segment distances),
 * circles subtract their radius.  Derivatives are those of the active pair: an obstacle end point against the footprint
 * (then exactly the point-obstacle case) or a footprint vertex against the interior of the obstacle segment
 * (d = |n.(p + R v - a)| - r, n = unit normal of the segment).
 */
static double footprint_distance_line(const mpcb200_config* cfg, const double* pose, const double* op, double* grad3, double* hess6)
{
    const double a[2] = {op[0], op[1]}, b[2] = {op[2], op[3]};
    double ux = b[0] - a[0], uy = b[1] - a[1];
    const double len = sqrt(ux * ux + uy * uy);
    if (!(len > 1e-12)) return orc_footprint_distance(cfg, pose, MPCB200_OBST_POINT, op, grad3, hess6);
    ux /= len; uy /= len;
    fp_seg seg[MPCB200_MAX_POLY + 2];
    const int ns = footprint_segments(cfg, seg);
    const double c = cos(pose[2]), s = sin(pose[2]);
    double best = 1e300;
    int kind = 0;               /* 1: obstacle end point a, 2: end point b, 3: footprint vertex vs segment interior */
    double bv[2] = {0, 0}, bsig = 1.0;
    for (int i = 0; i < ns; ++i)
    {
        /* footprint feature in the world */
        const double fr[2][2] = {{seg[i].ax, seg[i].ay}, {seg[i].bx, seg[i].by}};
        double fw[2][2];
        for (int e = 0; e < 2; ++e)
        {
            fw[e][0] = pose[0] + c * fr[e][0] - s * fr[e][1];
            fw[e][1] = pose[1] + s * fr[e][0] + c * fr[e][1];
        }
        const int degenerate = fr[0][0] == fr[1][0] && fr[0][1] == fr[1][1];
        if (!degenerate && seg_intersect(fw[0], fw[1], a, b))
        {
            if (grad3) grad3[0] = grad3[1] = grad3[2] = 0.0;
            if (hess6) for (int j = 0; j < 6; ++j) hess6[j] = 0.0;
            return 0.0;
        }
        /* obstacle end points against this footprint feature (world-frame point/segment distance) */
        const fp_seg wseg = {fw[0][0], fw[0][1], fw[1][0], fw[1][1], 0.0};
        for (int e = 0; e < 2; ++e)
        {
            double cx, cy; int isv;
            const double* q = e ? b : a;
            const double d = point_segment(q[0], q[1], &wseg, &cx, &cy, &isv) - seg[i].rad;
            if (d < best) { best = d; kind = 1 + e; }
        }
        /* footprint end points against the obstacle segment */
        for (int e = 0; e < (degenerate ? 1 : 2); ++e)
        {
            const double t = ((fw[e][0] - a[0]) * ux + (fw[e][1] - a[1]) * uy) / len;
            if (!(t > 0.0 && t < 1.0)) continue; /* clamped: that is an obstacle end point against this vertex, covered above */
            const double sd = -(fw[e][0] - a[0]) * uy + (fw[e][1] - a[1]) * ux;
            const double d = fabs(sd) - seg[i].rad;
            if (d < best) { best = d; kind = 3; bv[0] = fr[e][0]; bv[1] = fr[e][1]; bsig = sd >= 0.0 ? 1.0 : -1.0; }
        }
    }
    if (kind == 1 || kind == 2)
    {
        const double pt[5] = {kind == 1 ? a[0] : b[0], kind == 1 ? a[1] : b[1], 0, 0, 0};
        return orc_footprint_distance(cfg, pose, MPCB200_OBST_POINT, pt, grad3, hess6);
    }
    if (grad3 || hess6)
    {
        const double nx = -uy, ny = ux;
        const double rx = c * bv[0] - s * bv[1], ry = s * bv[0] + c * bv[1]; /* R v */
        if (grad3) { grad3[0] = bsig * nx; grad3[1] = bsig * ny; grad3[2] = bsig * (-nx * ry + ny * rx); }
        if (hess6) { hess6[0] = hess6[1] = hess6[2] = hess6[3] = hess6[4] = 0.0; hess6[5] = -bsig * (nx * rx + ny * ry); }
    }
    return best;
}

/* ---- config helpers ------------------------------------------------------------------------------------ */

static int xf_all_fixed(const mpcb200_config* c) { return c->xf_fixed[0] && c->xf_fixed[1] && c->xf_fixed[2]; }
static int has_quadratic(const mpcb200_config* c) { return c->objective == MPCB200_OBJ_QUADRATIC_FORM; }
/* planning/objective/quadratic_form/hybrid_cost_minimum_time: R/src/controller.cpp:595-620 installs corbo's
 * MinTimeQuadraticControls [EXT: dt per interval + the quadratic control term] only when Q is zero and R is not; in every
 * other case it logs an error and keeps the plain quadratic form. */
static int has_hybrid_mintime(const mpcb200_config* c)
{
    if (!c->hybrid_cost_minimum_time || c->objective != MPCB200_OBJ_QUADRATIC_FORM) return 0;
    int qz = 1, rz = 1;
    for (int i = 0; i < 9; ++i) qz = qz && c->Q[i] == 0.0;
    for (int i = 0; i < 4; ++i) rz = rz && c->R[i] == 0.0;
    return qz && !rz;
}
static int has_mintime(const mpcb200_config* c)
{
    return c->objective == MPCB200_OBJ_MINIMUM_TIME || c->objective == MPCB200_OBJ_MINIMUM_TIME_VIA_POINTS || has_hybrid_mintime(c);
}
static int has_viapoints(const mpcb200_config* c)
{
    return c->objective == MPCB200_OBJ_MINIMUM_TIME_VIA_POINTS ||
           (c->objective == MPCB200_OBJ_QUADRATIC_FORM && c->vp_attraction_with_quadratic);
}
/* terminal cost edge exists only if x_f is not fully fixed (R/src/optimal_control/finite_differences_grid_se2.cpp:126-131) */
static int has_terminal_cost(const mpcb200_config* c) { return c->terminal_cost && !xf_all_fixed(c); }
/* Integral form of the quadratic running cost (R/src/optimal_control/quadratic_cost_se2.cpp:54-84), integrated by the edge
 * R/src/optimal_control/finite_differences_grid_se2.cpp:57-72 selects: corbo's LeftSumCostEdge, dt l(x_k, u_k), or its
 * TrapezoidalIntegralCostEdge, dt/2 ( l(x_k, u_k) + l(x_{k+1}, u_k) ) [EXT: both ends use the control of the interval].
 * Summed over the horizon the state term of stage k carries dt * integral_state_weight, the control term of interval k dt. */
static int has_trapezoid(const mpcb200_config* c)
{ return has_quadratic(c) && c->quadratic_integral_form && c->cost_integration == MPCB200_COST_TRAPEZOIDAL; }
static double integral_state_weight(const mpcb200_config* c, int N, int k)
{
    if (c->cost_integration == MPCB200_COST_TRAPEZOIDAL) return (k == 0 || k == N - 1) ? 0.5 : 1.0;
    return k <= N - 2 ? 1.0 : 0.0;
}

/*
 * Row slots per stage (RS = 8 + K):
 *   0..3  k <= N-2: control bounds (u0 lb, u0 ub, u1 lb, u1 ub)       [R/src/controller.cpp:511,527,543]
 *         k == N-1: dt bounds (slot 0 lb, slot 1 ub) when dt is free   [R/src/controller.cpp:242-246]
 *                   slot 2: terminal ball on x_{N-1}                    [R/src/optimal_control/final_state_conditions_se2.cpp:54-64]
 *   4..7  control-rate rows of stage k (comp0 lb, comp0 ub, comp1 lb, comp1 ub), k = 0..N-1
 *         [R/src/optimal_control/stage_inequality_se2.cpp:191-222; wiring finite_differences_grid_se2.cpp:47-50,146-151]
 *   8..   obstacle rows, k = 1..N-2                                    [stage_inequality_se2.cpp:164-175]
 */
static int row_active(const orc_problem* p, const orc_ws* ws, int k, int slot)
{
    const mpcb200_config* c = p->cfg;
    const int N = ws->N;
    if (slot < 4)
    {
        if (k <= N - 2)
        {
            int i = slot >> 1;
            return (slot & 1) ? (c->u_ub[i] < MPCB200_INF) : (c->u_lb[i] > -MPCB200_INF);
        }
        if (slot == 2) return c->terminal_ball && !xf_all_fixed(c); /* TerminalBallSE2 on x_{N-1} */
        if (!c->variable_dt) return 0;
        if (slot == 0) return c->dt_lb > -MPCB200_INF;
        if (slot == 1) return c->dt_ub < MPCB200_INF;
        return 0;
    }
    if (slot < 8)
    {
        int i = (slot - 4) >> 1;
        int ub = (slot - 4) & 1;
        if (k == 0 && p->u_prev_dt == 0.0) return 0; /* rows identically zero (stage_inequality_se2.cpp:197-201) */
        return ub ? (c->du_ub[i] < MPCB200_INF) : (c->du_lb[i] > -MPCB200_INF);
    }
    if (k < 1 || k > N - 2) return 0;
    return ws->OBSIDX[IX(slot - 8, k)] >= 0.0;
}

/* dynamic obstacle (enable_dynamic_obstacles, velocity != 0) at its predicted position of stage k */
static int obstacle_is_dynamic(const mpcb200_config* c, const double* op) { return c->enable_dynamic_obstacles && (op[5] != 0.0 || op[6] != 0.0); }
static const double* obstacle_at(const mpcb200_config* c, const double* op, int k, double dt, double* buf)
{
    if (!obstacle_is_dynamic(c, op)) return op;
    const double t = (double)k * dt;
    buf[0] = op[0] + t * op[5]; buf[1] = op[1] + t * op[6]; buf[2] = op[2] + t * op[5]; buf[3] = op[3] + t * op[6];
    buf[4] = op[4]; buf[5] = op[5]; buf[6] = op[6];
    return buf;
}
/* optional extra output of row_value for rows that depend on the pose AND dt (dynamic obstacles): d2g/dx ddt (3), d2g/ddt2 */
static __thread double* g_row_hdt = NULL;

/*
 * Value of inequality row (k, slot) at (X,U,dt) and, optionally, its gradient wrt the local variables
 * loc = [x_k (0..2), u_k (3..4), u_{k-1} (5..6), dt (7)] and the Hessian wrt x_k (obstacle rows only).
 */
static double row_value(const orc_problem* p, const orc_ws* ws, int k, int slot, const double* X, const double* U,
                        double dt, double* grad8, double* hess6)
{
    const mpcb200_config* c = p->cfg;
    const int N = ws->N;
    if (grad8) for (int i = 0; i < 8; ++i) grad8[i] = 0.0;
    if (hess6) for (int i = 0; i < 6; ++i) hess6[i] = 0.0;
    if (slot < 4)
    {
        if (k <= N - 2)
        {
            int i = slot >> 1;
            if (slot & 1) { if (grad8) grad8[3 + i] = 1.0;  return U[IX(i, k)] - c->u_ub[i]; }
            if (grad8) grad8[3 + i] = -1.0;
            return c->u_lb[i] - U[IX(i, k)];
        }
        if (slot == 2)
        {
            /* terminal ball d' S d - gamma (final_state_conditions_se2.cpp:54-64) */
            const double d[3] = {X[IX(0, k)] - p->xf[0], X[IX(1, k)] - p->xf[1], orc_normalize_theta(X[IX(2, k)] - p->xf[2])};
            const double* S = c->terminal_ball_S;
            double g = -c->terminal_ball_gamma;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) g += d[i] * S[i * 3 + j] * d[j];
            if (grad8)
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) grad8[i] += (S[i * 3 + j] + S[j * 3 + i]) * d[j];
            if (hess6)
            {
                int q = 0;
                for (int i = 0; i < 3; ++i)
                    for (int j = i; j < 3; ++j, ++q) hess6[q] = S[i * 3 + j] + S[j * 3 + i];
            }
            return g;
        }
        if (slot == 0) { if (grad8) grad8[7] = -1.0; return c->dt_lb - dt; }
        if (grad8) grad8[7] = 1.0;
        return dt - c->dt_ub;
    }
    if (slot < 8)
    {
        int i = (slot - 4) >> 1;
        int ub = (slot - 4) & 1;
        /* Delta = u_k - u_{k-1}; k=0: u_{-1} = u_prev (const, dt_prev const); k=N-1: u_k := u_ref = 0 (const) */
        double uk = (k <= N - 2) ? U[IX(i, k)] : 0.0;
        double um = (k >= 1) ? U[IX(i, k - 1)] : p->u_prev[i];
        double T = (k >= 1) ? dt : p->u_prev_dt;
        double delta = uk - um;
        double sgn = ub ? 1.0 : -1.0;
        double bnd = ub ? c->du_ub[i] : c->du_lb[i];
        if (grad8)
        {
            if (k <= N - 2) grad8[3 + i] = sgn;
            if (k >= 1) grad8[5 + i] = -sgn;
            if (k >= 1 && c->variable_dt) grad8[7] = -sgn * bnd;
        }
        return sgn * (delta - bnd * T);
    }
    /* obstacle row: min_obstacle_dist - dist(footprint(x_k), obstacle) <= 0; a dynamic obstacle is taken at its predicted
       position at t = k dt (estimateSpatioTemporalDistance, stage_inequality_se2.cpp:177-189), which makes the row depend on dt */
    int j = (int)ws->OBSIDX[IX(slot - 8, k)];
    double pose[3] = {X[IX(0, k)], X[IX(1, k)], X[IX(2, k)]};
    double g3[3], h6[6], ob[MPCB200_OBST_STRIDE];
    const double* op = obstacle_at(c, p->obst_params + j * MPCB200_OBST_STRIDE, k, dt, ob);
    const int dyn = op == ob;
    double d = orc_footprint_distance(c, pose, p->obst_type[j], op, (grad8 || g_row_hdt) ? g3 : NULL, (hess6 || g_row_hdt) ? h6 : NULL);
    if (grad8) for (int i = 0; i < 3; ++i) grad8[i] = -g3[i];
    if (hess6) for (int i = 0; i < 6; ++i) hess6[i] = -h6[i];
    if (dyn && c->variable_dt)
    {
        /* the distance depends on p - (o + k dt v): d/ddt = -k v . d/dp */
        const double kk = (double)k, vx = op[5], vy = op[6];
        if (grad8) grad8[7] = -kk * (-g3[0] * vx - g3[1] * vy);
        if (g_row_hdt)
        {
            const double hx = -(h6[0] * vx + h6[1] * vy), hy = -(h6[1] * vx + h6[3] * vy), ht = -(h6[2] * vx + h6[4] * vy);
            g_row_hdt[0] = -kk * hx; g_row_hdt[1] = -kk * hy; g_row_hdt[2] = -kk * ht; g_row_hdt[3] = kk * kk * (vx * hx + vy * hy);
        }
    }
    else if (g_row_hdt) g_row_hdt[0] = g_row_hdt[1] = g_row_hdt[2] = g_row_hdt[3] = 0.0;
    return c->min_obstacle_dist - d;
}

/* quadratic form d^T W d for a full 3x3 / 2x2 row-major matrix */
static double quad3(const double* W, const double* d)
{
    double r = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r += d[i] * W[i * 3 + j] * d[j];
    return r;
}
static double quad2(const double* W, const double* d)
{
    return d[0] * (W[0] * d[0] + W[1] * d[1]) + d[1] * (W[2] * d[0] + W[3] * d[1]);
}

/*
 * Objective of an iterate:
 *   minimum_time: (N-1) dt                                [corbo::MinimumTime; R/src/optimal_control/min_time_via_points_cost.cpp:52-56,120-123]
 *   quadratic_form (non-integral): sum_{k=0}^{N-2} d^T Q d + u^T R u, d = x_k - x_goal with wrapped angle
 *                                                         [R/src/optimal_control/quadratic_cost_se2.cpp:31-52; u_ref = 0, R/src/controller.cpp:169-170]
 *   terminal quadratic cost d^T Qf d if x_f not fully fixed [R/src/optimal_control/final_state_conditions_se2.cpp:31-52]
 *   via-points: w_p |p_vp - p_k|^2 (+ w_theta wrap(theta_vp - theta_k) if w_theta > 0)  [min_time_via_points_cost.cpp:130-145]
 */
double orc_objective(const orc_problem* p, const orc_ws* ws, const double* X, const double* U, double dt)
{
    const mpcb200_config* c = p->cfg;
    const int N = ws->N;
    double J = 0.0;
    if (has_mintime(c)) J += (double)(N - 1) * dt;
    if (has_quadratic(c))
    {
        for (int k = 0; k <= N - 2; ++k)
        {
            double d[3] = {X[IX(0, k)] - p->xf[0], X[IX(1, k)] - p->xf[1], orc_normalize_theta(X[IX(2, k)] - p->xf[2])};
            double u[2] = {U[IX(0, k)], U[IX(1, k)]};
            /* integral form: dt * (w_k l_x + l_u), w_k by the integration rule (integral_state_weight) */
            const int integ = c->quadratic_integral_form != 0;
            J += (integ ? dt * integral_state_weight(c, N, k) : 1.0) * quad3(c->Q, d) + (integ ? dt : 1.0) * quad2(c->R, u);
        }
        if (has_trapezoid(c))
        {
            int k = N - 1;
            double d[3] = {X[IX(0, k)] - p->xf[0], X[IX(1, k)] - p->xf[1], orc_normalize_theta(X[IX(2, k)] - p->xf[2])};
            J += dt * integral_state_weight(c, N, k) * quad3(c->Q, d);
        }
    }
    if (has_terminal_cost(c))
    {
        int k = N - 1;
        double d[3] = {X[IX(0, k)] - p->xf[0], X[IX(1, k)] - p->xf[1], orc_normalize_theta(X[IX(2, k)] - p->xf[2])};
        J += quad3(c->Qf, d);
    }
    if (has_viapoints(c))
    {
        for (int j = 0; j < p->n_vp; ++j)
        {
            int k = ws->vp_stage[j];
            if (k < 0) continue;
            double ex = p->vp[3 * j] - X[IX(0, k)], ey = p->vp[3 * j + 1] - X[IX(1, k)];
            J += c->vp_position_weight * (ex * ex + ey * ey);
            if (c->vp_orientation_weight > 0)
                J += c->vp_orientation_weight * orc_normalize_theta(p->vp[3 * j + 2] - X[IX(2, k)]);
        }
    }
    return J;
}

/* ======================================================================================================= */
/* workspace                                                                                                */
/* ======================================================================================================= */

orc_ws* orc_ws_alloc(int N, int K)
{
    orc_ws* ws = (orc_ws*)calloc(1, sizeof(orc_ws));
    if (!ws) return NULL;
    ws->N = N; ws->K = K; ws->RS = 8 + K;
    int RS = ws->RS;
#define ALLOC(f, n) ws->f = (double*)calloc((size_t)(n), sizeof(double))
    ALLOC(X, 3 * N); ALLOC(U, 2 * N); ALLOC(NU, 3 * N); ALLOC(S, RS * N); ALLOC(LAM, RS * N);
    ALLOC(KKT, KW * N); ALLOC(STEP, 8 * N); ALLOC(OBSIDX, (K > 0 ? K : 1) * N);
    ALLOC(GL, 5 * N); ALLOC(G, RS * N); ALLOC(DS, RS * N); ALLOC(DLAM, RS * N); ALLOC(XT, 3 * N); ALLOC(UT, 2 * N);
    ALLOC(P, 25 * N); ALLOC(PI, 25 * N); ALLOC(KG, 10 * N); ALLOC(KT, 10 * N);
#undef ALLOC
    for (int i = 0; i < (K > 0 ? K : 1) * N; ++i) ws->OBSIDX[i] = -1.0;
    for (int i = 0; i < 64; ++i) ws->vp_stage[i] = -1;
    ws->cold = 1;
    return ws;
}

void orc_ws_free(orc_ws* ws)
{
    if (!ws) return;
    free(ws->X); free(ws->U); free(ws->NU); free(ws->S); free(ws->LAM); free(ws->KKT); free(ws->STEP); free(ws->OBSIDX);
    free(ws->GL); free(ws->G); free(ws->DS); free(ws->DLAM); free(ws->XT); free(ws->UT);
    free(ws->P); free(ws->PI); free(ws->KG); free(ws->KT);
    free(ws);
}

/* ======================================================================================================= */
/* initial guess, warm start, association                                                                   */
/* ======================================================================================================= */

/*
 * Cold initialisation: FullDiscretizationGridBaseSE2::initializeSequences with an xinit reference
 * [R/src/optimal_control/full_discretization_grid_base_se2.cpp:192-239]: x_0 exact, x_k = xinit(k*dt_ref) for
 * k = 1..N-2, x_{N-1} = goal, u_k = u_ref = 0, dt = dt_ref.  With x_init == NULL the two-pose initial plan of
 * Controller::step(start, goal, ...) [R/src/controller.cpp:102-109,807-857] is sampled: linear in time with
 * TimeSeriesSE2's angle-aware interpolation [R/src/utils/time_series_se2.cpp:86-102].
 */
/*
 * Cold initial guess: linear interpolation start -> goal (FullDiscretizationGridBaseSE2::initializeSequences,
 * full_discretization_grid_base_se2.cpp:192-239) or the supplied initial plan.  Solver-side addition (not in the
 * reference): without an initial plan, the straight line is replaced by the laterally bumped line
 *     p_k + A sin(pi k/(N-1)) n_perp,  A = ORC_BUMP_STEP * m,  m = -initial_guess_bumps .. initial_guess_bumps,
 * that violates the obstacle clearances (d_min + margin, all obstacles, all interior stages) least; |A| breaks ties, so a
 * clear straight line stays.  A locally convergent method inherits the homotopy class of its starting point: with the
 * straight line 40 % of the SURVEY-8d instances end in an infeasible stationary point squeezed between obstacles.
 */
#define ORC_BUMP_STEP 0.4
#define ORC_BUMP_MARGIN 0.05
static double bump_score(const orc_problem* p, const orc_ws* ws, double A, double nx, double ny)
{
    const mpcb200_config* c = p->cfg;
    const int N = ws->N;
    double score = 1e-3 * fabs(A);
    for (int k = 1; k <= N - 2; ++k)
    {
        const double o = A * sin(M_PI * (double)k / (double)(N - 1));
        const double pose[3] = {ws->X[IX(0, k)] + o * nx, ws->X[IX(1, k)] + o * ny, ws->X[IX(2, k)]};
        for (int j = 0; j < p->n_obst; ++j)
        {
            double ob[MPCB200_OBST_STRIDE];
            const double* op = obstacle_at(c, p->obst_params + j * MPCB200_OBST_STRIDE, k, c->dt_ref, ob);
            const double v = c->min_obstacle_dist + ORC_BUMP_MARGIN - orc_footprint_distance(c, pose, p->obst_type[j], op, NULL, NULL);
            if (v > 0.0) score += v;
        }
    }
    return score;
}
void orc_init_cold(const orc_problem* p, const double* x_init, orc_ws* ws)
{
    const int N = ws->N;
    for (int k = 0; k < N; ++k)
    {
        if (x_init && k > 0 && k < N - 1)
        {
            for (int i = 0; i < 3; ++i) ws->X[IX(i, k)] = x_init[3 * k + i];
        }
        else if (k == 0)
        {
            for (int i = 0; i < 3; ++i) ws->X[IX(i, k)] = p->x0[i];
        }
        else if (k == N - 1)
        {
            for (int i = 0; i < 3; ++i) ws->X[IX(i, k)] = p->xf[i];
        }
        else
        {
            double frac = (double)k / (double)(N - 1);
            ws->X[IX(0, k)] = p->x0[0] + frac * (p->xf[0] - p->x0[0]);
            ws->X[IX(1, k)] = p->x0[1] + frac * (p->xf[1] - p->x0[1]);
            ws->X[IX(2, k)] = orc_interpolate_angle(p->x0[2], p->xf[2], frac);
        }
        ws->U[IX(0, k)] = 0.0;
        ws->U[IX(1, k)] = 0.0;
    }
    ws->SCAL[MPCB200_SC_DT] = p->cfg->dt_ref;
    const int nb = p->cfg->reference_initial_guess ? 0 : p->cfg->initial_guess_bumps;   /* reference guess: the straight line itself */
    if (!x_init && nb > 0 && p->n_obst > 0)
    {
        double nx = -(p->xf[1] - p->x0[1]), ny = p->xf[0] - p->x0[0];
        const double nn = sqrt(nx * nx + ny * ny);
        if (nn > 1e-9)
        {
            nx /= nn; ny /= nn;
            double best = 1e300, best_a = 0.0;
            for (int m = -nb; m <= nb; ++m)
            {
                const double A = ORC_BUMP_STEP * (double)m;
                const double score = bump_score(p, ws, A, nx, ny);
                /* strictly better beyond rounding noise: the order of the candidates decides exact ties (symmetric scenes) */
                if (score < best - 1e-9 * (1.0 + best)) { best = score; best_a = A; }
            }
            for (int k = 1; k <= N - 2; ++k)
            {
                const double o = best_a * sin(M_PI * (double)k / (double)(N - 1));
                ws->X[IX(0, k)] += o * nx;
                ws->X[IX(1, k)] += o * ny;
            }
            /* car-like models cannot turn on the spot: on a bumped line their headings follow the path (central differences) */
            if (best_a != 0.0 && p->cfg->robot_type != MPCB200_ROBOT_UNICYCLE)
                for (int k = 1; k <= N - 2; ++k)
                    ws->X[IX(2, k)] = atan2(ws->X[IX(1, k + 1)] - ws->X[IX(1, k - 1)], ws->X[IX(0, k + 1)] - ws->X[IX(0, k - 1)]);
        }
    }
}

/*
 * MpcLocalPlannerROS::updateObstacleContainerWithCostmap [R/src/mpc_local_planner_ros.cpp:474-499] for one robot:
 * cells mx = 0..size_x-2 (outer loop), my = 0..size_y-2 (inner loop) whose cost is costmap_2d::LETHAL_OBSTACLE (254) become
 * point obstacles at Costmap2D::mapToWorld(mx, my) = origin + (m + 0.5) resolution [EXT: costmap_2d], unless
 * obs_dir . robot_orient < 0 and |obs_dir| > behind_dist (:492-493).  cost: cell (mx, my) at my*size_x + mx (Costmap2D::getIndex).
 * Writes at most max_out points (x, y pairs) in the reference's push_back order; returns how many cells qualified.
 */
int orc_costmap_obstacles(int size_x, int size_y, double resolution, const double* origin, const unsigned char* cost,
                          const double* robot_pose, double behind_dist, int max_out, double* xy)
{
    const double ox = cos(robot_pose[2]), oy = sin(robot_pose[2]); /* PoseSE2::orientationUnitVec */
    int found = 0;
    for (int i = 0; i < size_x - 1; ++i)
        for (int j = 0; j < size_y - 1; ++j)
        {
            if (cost[(size_t)j * size_x + i] != 254) continue;
            const double wx = origin[0] + (i + 0.5) * resolution, wy = origin[1] + (j + 0.5) * resolution;
            const double dx = wx - robot_pose[0], dy = wy - robot_pose[1];
            if (dx * ox + dy * oy < 0 && sqrt(dx * dx + dy * dy) > behind_dist) continue;
            if (found < max_out) { xy[2 * found] = wx; xy[2 * found + 1] = wy; }
            ++found;
        }
    return found;
}

/*
 * Controller::isPoseTrajectoryFeasible [R/src/controller.cpp:859-917], called by MpcLocalPlannerROS::computeVelocityCommands
 * [R/src/mpc_local_planner_ros.cpp:414-428]: the footprint is laid over the costmap at the first look_ahead_idx + 1 poses of
 * the trajectory and, where two consecutive poses are farther apart than the inscribed radius or turn more than
 * min_resolution_collision_check_angular, at evenly spaced poses between them (the intermediate pose is ACCUMULATED step by
 * step, :903-906).  Infeasible <=> some footprintCost == -1.
 * [EXT] base_local_planner::CostmapModel::footprintCost / lineCost / pointCost and LineIterator, costmap_2d::Costmap2D::worldToMap
 * (ROS navigation; not in /root/reference -- restated from upstream knowledge):
 *   centre outside the map -> -1;  fewer than 3 footprint points: cost of the centre cell, LETHAL (254) or INSCRIBED (253) -> -1,
 *   NO_INFORMATION (255) -> -2;  else every footprint edge (closing edge included) is rasterised with the Bresenham LineIterator:
 *   a vertex outside the map -> -3 (NOT -1: such a pose passes the check), a LETHAL cell -> -1, a NO_INFORMATION cell -> -2.
 * x_seq: [n][3] poses.  footprint: n_fp points (x, y) in the robot frame.  Returns 1 = feasible, 0 = not.
 */
static int orc_world_to_map(int size_x, int size_y, double res, const double* origin, double wx, double wy, int* mx, int* my)
{
    if (wx < origin[0] || wy < origin[1]) return 0;
    *mx = (int)((wx - origin[0]) / res);
    *my = (int)((wy - origin[1]) / res);
    return *mx < size_x && *my < size_y;
}
static double orc_line_cost(int size_x, const unsigned char* cost, int x0, int x1, int y0, int y1)
{
    /* base_local_planner::LineIterator */
    const int deltax = abs(x1 - x0), deltay = abs(y1 - y0);
    int x = x0, y = y0, xinc1, xinc2, yinc1, yinc2, den, num, numadd, numpixels;
    if (x1 >= x0) { xinc1 = 1; xinc2 = 1; } else { xinc1 = -1; xinc2 = -1; }
    if (y1 >= y0) { yinc1 = 1; yinc2 = 1; } else { yinc1 = -1; yinc2 = -1; }
    if (deltax >= deltay) { xinc1 = 0; yinc2 = 0; den = deltax; num = deltax / 2; numadd = deltay; numpixels = deltax; }
    else { xinc2 = 0; yinc1 = 0; den = deltay; num = deltay / 2; numadd = deltax; numpixels = deltay; }
    double line_cost = 0.0;
    for (int cur = 0; cur <= numpixels; ++cur)
    {
        const unsigned char c = cost[(size_t)y * size_x + x];
        const double pc = c == 255 ? -2.0 : (c == 254 ? -1.0 : (double)c);   /* pointCost */
        if (pc < 0) return pc;
        if (line_cost < pc) line_cost = pc;
        num += numadd;
        if (num >= den) { num -= den; x += xinc1; y += yinc1; }
        x += xinc2; y += yinc2;
    }
    return line_cost;
}
static double orc_footprint_cost(int size_x, int size_y, double res, const double* origin, const unsigned char* cost, double px, double py,
                                 double th, const double* fp, int n_fp)
{
    int cx, cy;
    if (!orc_world_to_map(size_x, size_y, res, origin, px, py, &cx, &cy)) return -1.0;
    if (n_fp < 3)
    {
        const unsigned char c = cost[(size_t)cy * size_x + cx];
        if (c == 255) return -2.0;
        if (c == 254 || c == 253) return -1.0;
        return (double)c;
    }
    const double co = cos(th), si = sin(th);
    double fc = 0.0;
    for (int i = 0; i < n_fp; ++i)
    {
        const int j = (i + 1) % n_fp;   /* edges 0-1, 1-2, ..., then the closing edge last -> first */
        const double ax = px + (fp[2 * i] * co - fp[2 * i + 1] * si), ay = py + (fp[2 * i] * si + fp[2 * i + 1] * co);
        const double bx = px + (fp[2 * j] * co - fp[2 * j + 1] * si), by = py + (fp[2 * j] * si + fp[2 * j + 1] * co);
        int x0, y0, x1, y1;
        if (!orc_world_to_map(size_x, size_y, res, origin, ax, ay, &x0, &y0)) return -3.0;
        if (!orc_world_to_map(size_x, size_y, res, origin, bx, by, &x1, &y1)) return -3.0;
        const double lc = orc_line_cost(size_x, cost, x0, x1, y0, y1);
        if (fc < lc) fc = lc;
        if (lc < 0) return lc;
    }
    return fc;
}
int orc_pose_trajectory_feasible(int size_x, int size_y, double resolution, const double* origin, const unsigned char* cost,
                                 const double* x_seq, int n, const double* footprint, int n_fp, double inscribed_radius,
                                 double min_resolution_angular, int look_ahead_idx)
{
    if (n < 2) return 0;
    if (look_ahead_idx < 0 || look_ahead_idx >= n) look_ahead_idx = n - 1;
    for (int i = 0; i <= look_ahead_idx; ++i)
    {
        const double* p = x_seq + 3 * i;
        if (orc_footprint_cost(size_x, size_y, resolution, origin, cost, p[0], p[1], p[2], footprint, n_fp) == -1.0) return 0;
        if (i < look_ahead_idx)
        {
            const double* q = x_seq + 3 * (i + 1);
            const double delta_rot = orc_normalize_theta(q[2] - p[2]);
            const double dx = q[0] - p[0], dy = q[1] - p[1];
            const double dist = sqrt(dx * dx + dy * dy);
            if (fabs(delta_rot) > min_resolution_angular || dist > inscribed_radius)
            {
                const double a = ceil(fabs(delta_rot) / min_resolution_angular), b = ceil(dist / inscribed_radius);
                const int n_add = (int)(a > b ? a : b) - 1;
                double ix = p[0], iy = p[1], ith = p[2];
                for (int step = 0; step < n_add; ++step)
                {
                    ix = ix + dx / (n_add + 1.0);
                    iy = iy + dy / (n_add + 1.0);
                    ith = orc_normalize_theta(ith + delta_rot / (n_add + 1.0));
                    if (orc_footprint_cost(size_x, size_y, resolution, origin, cost, ix, iy, ith, footprint, n_fp) == -1.0) return 0;
                }
            }
        }
    }
    return 1;
}

/*
 * FullDiscretizationGridBaseSE2::resampleTrajectory(n_new)
 * [R/src/optimal_control/full_discretization_grid_base_se2.cpp:440-524], the operation the grid adaptation applies with
 * n_new = n +- 1 [R/src/optimal_control/finite_differences_variable_grid_se2.cpp:99-121].
 *   dt_new = dt (n-1)/(n_new-1)                                                              (:467)
 *   for idx_new = 1 .. n_new-2: t = idx_new dt_new; idx_old = first old sample with idx_old dt >= t (search resumes where
 *   the previous sample stopped, :478-481); x = x_prev + f (x_cur - x_prev), f = (t - (idx_old-1) dt)/dt, with
 *   x_prev = old sample idx_old-1, x_cur = old sample idx_old or the final-state vertex (:484-485); heading by
 *   interpolate_angle (:493); control = old control of interval idx_old-1 (:497, :507)
 *   sample 0 and the final state are not touched (:472-474).
 * X / U: component-major, n columns (column n-1 of X = final state); Xn / Un: n_new columns.  Returns dt_new.
 */
double orc_resample_trajectory(int n, const double* X, const double* U, double dt, int n_new, double* Xn, double* Un)
{
    const double dt_new = dt * (double)(n - 1) / (double)(n_new - 1);
    int idx_old = 1;
    for (int c = 0; c < 3; ++c) Xn[c * n_new + 0] = X[c * n + 0];
    for (int c = 0; c < 2; ++c) Un[c * n_new + 0] = U[c * n + 0];
    for (int idx_new = 1; idx_new < n_new - 1; ++idx_new)
    {
        const double t_new = dt_new * (double)idx_new;
        while (t_new > (double)idx_old * dt && idx_old < n) ++idx_old;
        const double t_old_p1 = (double)idx_old * dt;
        int prev = idx_old - 1, cur = idx_old;
        if (cur > n - 1) cur = n - 1;   /* beyond the last interval: the final-state vertex */
        if (prev > n - 1) prev = n - 1;
        const double frac = dt > 0.0 ? (t_new - (t_old_p1 - dt)) / dt : 0.0;
        for (int c = 0; c < 2; ++c) Xn[c * n_new + idx_new] = X[c * n + prev] + frac * (X[c * n + cur] - X[c * n + prev]);
        Xn[2 * n_new + idx_new] = orc_interpolate_angle(X[2 * n + prev], X[2 * n + cur], frac);
        int iu = idx_old - 1;
        if (iu > n - 2) iu = n - 2;     /* the control time series repeats its last sample */
        for (int c = 0; c < 2; ++c) Un[c * n_new + idx_new] = U[c * n + iu];
    }
    for (int c = 0; c < 3; ++c) Xn[c * n_new + n_new - 1] = X[c * n + n - 1];
    for (int c = 0; c < 2; ++c) Un[c * n_new + n_new - 1] = 0.0;
    return dt_new;
}

/*
 * Warm start: FullDiscretizationGridBaseSE2::warmStartShifting + findNearestState
 * [R/src/optimal_control/full_discretization_grid_base_se2.cpp:241-339], then x_0 <- measured state and the fixed
 * components of x_f <- goal [:104-109].
 */
void orc_warm_shift(const orc_problem* p, orc_ws* ws)
{
    const int N = ws->N;
    double* X = ws->X; double* U = ws->U;
    /* findNearestState: greedy descent over at most 20 samples, plain Euclidean norm incl. unwrapped theta */
    int num_shift = 0;
    {
        double d0 = 0;
        for (int i = 0; i < 3; ++i) { double e = p->x0[i] - X[IX(i, 0)]; d0 += e * e; }
        d0 = sqrt(d0);
        if (fabs(d0) >= 1e-12)
        {
            int num_interv = N - 1, lookahead = num_interv - 1 < 20 ? num_interv - 1 : 20;
            double cache = d0;
            for (int i = 1; i <= lookahead; ++i)
            {
                double d = 0;
                for (int j = 0; j < 3; ++j) { double e = p->x0[j] - X[IX(j, i)]; d += e * e; }
                d = sqrt(d);
                if (d < cache) { cache = d; num_shift = i; }
                else break;
            }
        }
    }
    if (num_shift > 0 && num_shift <= N - 2)
    {
        for (int i = 0; i < N - num_shift; ++i)
        {
            int idx = i + num_shift;
            for (int j = 0; j < 3; ++j) X[IX(j, i)] = X[IX(j, idx)];
            if (idx != N - 1)
                for (int j = 0; j < 2; ++j) U[IX(j, i)] = U[IX(j, idx)];
        }
        int idx = N - num_shift;
        for (int i = 0; i < num_shift; ++i, ++idx)
        {
            for (int j = 0; j < 2; ++j) X[IX(j, idx)] = X[IX(j, idx - 2)] + 2.0 * (X[IX(j, idx - 1)] - X[IX(j, idx - 2)]);
            X[IX(2, idx)] = orc_interpolate_angle(X[IX(2, idx - 2)], X[IX(2, idx - 1)], 2.0);
            for (int j = 0; j < 2; ++j) U[IX(j, idx - 1)] = U[IX(j, idx - 2)];
        }
    }
    for (int i = 0; i < 3; ++i) X[IX(i, 0)] = p->x0[i];
    for (int i = 0; i < 3; ++i)
        if (p->cfg->xf_fixed[i]) X[IX(i, N - 1)] = p->xf[i];
}

/*
 * Obstacle association: StageInequalitySE2::update [R/src/optimal_control/stage_inequality_se2.cpp:50-162]:
 * per stage k = 1..N-1, every obstacle closer than force_inclusion_dist is kept; of the others within cutoff_dist the
 * nearest on the left and the nearest on the right are kept, the side being the sign of cross2d(heading, obstacle
 * centroid IN WORLD COORDINATES) (reference quirk, SURVEY App. C.5 -- mirrored).  Rows exist only for k = 1..N-2
 * (SURVEY 8a, a13).  Row budget K: candidates are appended in reference order; when the K slots are full a candidate
 * replaces the farthest kept obstacle if it is nearer (no-op whenever the reference's list fits in K).
 * Via-point association: MinTimeViaPointsCost::update [R/src/optimal_control/min_time_via_points_cost.cpp:40-118] with
 * findClosestPose [R/src/optimal_control/full_discretization_grid_base_se2.cpp:364-388].
 */
static void assoc_insert(orc_ws* ws, int k, int* cnt, double* dists, int j, double dist)
{
    const int N = ws->N, K = ws->K;
    if (K <= 0) return;
    if (*cnt < K)
    {
        ws->OBSIDX[IX(*cnt, k)] = (double)j;
        dists[*cnt] = dist;
        ++*cnt;
        return;
    }
    int far = 0;
    for (int i = 1; i < K; ++i)
        if (dists[i] > dists[far]) far = i;
    if (dist < dists[far])
    {
        ws->OBSIDX[IX(far, k)] = (double)j;
        dists[far] = dist;
    }
}

void orc_associate(const orc_problem* p, orc_ws* ws)
{
    const mpcb200_config* c = p->cfg;
    const int N = ws->N, K = ws->K;
    for (int i = 0; i < (K > 0 ? K : 1) * N; ++i) ws->OBSIDX[i] = -1.0;
    for (int k = 1; k <= N - 2 && K > 0; ++k)
    {
        double pose[3] = {ws->X[IX(0, k)], ws->X[IX(1, k)], ws->X[IX(2, k)]};
        double ox = cos(pose[2]), oy = sin(pose[2]);
        double left_min = 1e300, right_min = 1e300;
        int left = -1, right = -1, cnt = 0;
        double dists[64];
        for (int j = 0; j < p->n_obst; ++j)
        {
            double ob[MPCB200_OBST_STRIDE];
            const double* op0 = p->obst_params + j * MPCB200_OBST_STRIDE;
            const double* op = obstacle_at(c, op0, k, ws->SCAL[MPCB200_SC_DT], ob);
            double dist = orc_footprint_distance(c, pose, p->obst_type[j], op, NULL, NULL);
            /* dynamic obstacles are kept at every stage (stage_inequality_se2.cpp:99-106) */
            if (dist < c->force_inclusion_dist || obstacle_is_dynamic(c, op0)) { assoc_insert(ws, k, &cnt, dists, j, dist); continue; }
            if (dist > c->cutoff_dist) continue;
            /* teb getCentroid(): point / circle centre, segment midpoint; quirk 5: world coordinates, not relative to the pose */
            const int is_line = p->obst_type[j] == MPCB200_OBST_LINE;
            const double ccx = is_line ? 0.5 * (op[0] + op[2]) : op[0], ccy = is_line ? 0.5 * (op[1] + op[3]) : op[1];
            if (ox * ccy - ccx * oy > 0)
            {
                if (dist < left_min) { left_min = dist; left = j; }
            }
            else
            {
                if (dist < right_min) { right_min = dist; right = j; }
            }
        }
        if (left >= 0) assoc_insert(ws, k, &cnt, dists, left, left_min);
        if (right >= 0) assoc_insert(ws, k, &cnt, dists, right, right_min);
    }
    /* via-points */
    for (int j = 0; j < 64; ++j) ws->vp_stage[j] = -1;
    if (has_viapoints(c))
    {
        int start_idx = 0;
        for (int j = 0; j < p->n_vp && j < 64; ++j)
        {
            double min_dist = 1e300;
            int idx = -1;
            for (int i = start_idx; i < N - 1; ++i)
            {
                double dx = ws->X[IX(0, i)] - p->vp[3 * j], dy = ws->X[IX(1, i)] - p->vp[3 * j + 1];
                double d = sqrt(dx * dx + dy * dy);
                if (d < min_dist) { min_dist = d; idx = i; }
            }
            {
                double dx = ws->X[IX(0, N - 1)] - p->vp[3 * j], dy = ws->X[IX(1, N - 1)] - p->vp[3 * j + 1];
                double d = sqrt(dx * dx + dy * dy);
                if (d < min_dist) { min_dist = d; idx = N - 1; }
            }
            if (c->vp_ordered) start_idx = idx + 2;
            if (idx > N - 2) idx = N - 2;
            if (idx < 1)
            {
                if (c->vp_ordered) idx = 1;
                else continue;
            }
            ws->vp_stage[j] = idx;
        }
    }
}

/*
 * Initial-guess repair (solver-side, not in the reference): poses of the initial guess that violate an associated
 * obstacle row (the straight-line guess usually cuts through obstacles) are pushed out along the row's position
 * gradient until the row holds with margin.  Interior-point iterations started from violated nonconvex rows jam
 * (Waechter & Biegler 2000); Ipopt escapes through its restoration phase, we avoid the situation up front.
 * A pose that stays pinched between obstacles after these sweeps is moved sideways instead: along the normal of the
 * start -> goal line, in 0.1 m increments up to +-2.5 m, to the clear position whose lateral offset is closest to the one
 * of the previous stage (so that consecutive poses pass an obstacle on the same side).
 */
#define ORC_PROJ_MARGIN 0.05
#define ORC_PROJ_SWEEPS 6
#define ORC_LAT_STEP 0.1
#define ORC_LAT_MAX_STEPS 25
static double row_value(const orc_problem* p, const orc_ws* ws, int k, int slot, const double* X, const double* U,
                        double dt, double* grad8, double* hess6);
static double stage_max_row(const orc_problem* p, orc_ws* ws, int k)
{
    const int N = ws->N;
    double m = -1e300;
    for (int sl = 8; sl < 8 + ws->K; ++sl)
    {
        if (ws->OBSIDX[IX(sl - 8, k)] < 0.0) continue;
        double g = row_value(p, ws, k, sl, ws->X, ws->U, ws->SCAL[MPCB200_SC_DT], NULL, NULL);
        if (g > m) m = g;
    }
    return m;
}
void orc_project_init(const orc_problem* p, orc_ws* ws)
{
    const int N = ws->N, K = ws->K;
    const double margin = ORC_PROJ_MARGIN;
    /* unit normal of the start -> goal line */
    double nx = -(ws->X[IX(1, N - 1)] - ws->X[IX(1, 0)]), ny = ws->X[IX(0, N - 1)] - ws->X[IX(0, 0)];
    { double nn = sqrt(nx * nx + ny * ny); if (nn < 1e-12) { nx = 0; ny = 1; } else { nx /= nn; ny /= nn; } }
    double o_prev = 0.0;
    for (int k = 1; k <= N - 2; ++k)
    {
        const double bx = ws->X[IX(0, k)], by = ws->X[IX(1, k)];
        for (int sweep = 0; sweep < ORC_PROJ_SWEEPS; ++sweep)
        {
            int moved = 0;
            for (int sl = 8; sl < 8 + K; ++sl)
            {
                if (ws->OBSIDX[IX(sl - 8, k)] < 0.0) continue;
                double grad[8];
                double g = row_value(p, ws, k, sl, ws->X, ws->U, ws->SCAL[MPCB200_SC_DT], grad, NULL);
                if (g <= -margin) continue;
                double n2 = grad[0] * grad[0] + grad[1] * grad[1];
                if (n2 < 1e-16) { grad[0] = 1.0; grad[1] = 0.0; n2 = 1.0; }
                double step = (g + margin) / n2;
                ws->X[IX(0, k)] -= step * grad[0];
                ws->X[IX(1, k)] -= step * grad[1];
                moved = 1;
            }
            if (!moved) break;
        }
        if (stage_max_row(p, ws, k) > -0.5 * margin)
        {
            /* the sweeps are pinched between obstacles: look sideways for the clear lateral offset closest to the previous stage's */
            double best = 1e300, best_o = 0.0; int found = 0;
            for (int m = -ORC_LAT_MAX_STEPS; m <= ORC_LAT_MAX_STEPS; ++m)
            {
                const double o = ORC_LAT_STEP * (double)m;
                ws->X[IX(0, k)] = bx + o * nx; ws->X[IX(1, k)] = by + o * ny;
                if (stage_max_row(p, ws, k) > -margin) continue;
                const double cost = fabs(o - o_prev) + 1e-3 * fabs(o);
                if (cost < best) { best = cost; best_o = o; found = 1; }
            }
            if (found) { ws->X[IX(0, k)] = bx + best_o * nx; ws->X[IX(1, k)] = by + best_o * ny; }
            else { ws->X[IX(0, k)] = bx; ws->X[IX(1, k)] = by; }
        }
        o_prev = (ws->X[IX(0, k)] - bx) * nx + (ws->X[IX(1, k)] - by) * ny;
    }
}

/*
 * Initial controls (solver-side, not in the reference, which starts from u = u_ref = 0): the linearisation of a
 * nonholonomic model at rest is uncontrollable (df/dx ~ v = 0), which makes the very first KKT system singular for
 * fixed terminal states.  The controls are therefore seeded by inverting the dynamics along the state guess,
 * u_k ~ argmin |f(x_k,u) - (x_{k+1}-x_k)/dt|, and then clipped strictly inside the control bounds and the
 * control-rate rows (forward pass from u_prev, backward pass from the final rate rows towards u_ref = 0).
 */
#define ORC_INIT_SHRINK 0.9
void orc_init_controls(const orc_problem* p, orc_ws* ws)
{
    const mpcb200_config* c = p->cfg;
    const int N = ws->N;
    const double dt = ws->SCAL[MPCB200_SC_DT];
    double lo[2], hi[2];
    for (int i = 0; i < 2; ++i)
    {
        double l = c->u_lb[i] > -MPCB200_INF ? c->u_lb[i] : -1e6, h = c->u_ub[i] < MPCB200_INF ? c->u_ub[i] : 1e6;
        double mid = 0.5 * (l + h), half = 0.5 * (h - l) * ORC_INIT_SHRINK;
        lo[i] = mid - half; hi[i] = mid + half;
    }
    for (int k = 0; k <= N - 2; ++k)
    {
        double th = ws->X[IX(2, k)];
        double dx = ws->X[IX(0, k + 1)] - ws->X[IX(0, k)], dy = ws->X[IX(1, k + 1)] - ws->X[IX(1, k)];
        double dth = orc_normalize_theta(ws->X[IX(2, k + 1)] - th);
        double v = (dx * cos(th) + dy * sin(th)) / dt, w = dth / dt, u1 = 0.0;
        switch (c->robot_type)
        {
            case MPCB200_ROBOT_UNICYCLE: u1 = w; break;
            case MPCB200_ROBOT_SIMPLE_CAR: u1 = fabs(v) > 1e-3 ? atan(c->wheelbase * w / v) : 0.0; break;
            case MPCB200_ROBOT_SIMPLE_CAR_FRONT:
            {
                double a = fabs(v) > 1e-3 ? c->wheelbase * w / v : 0.0;
                u1 = asin(a > 1 ? 1 : (a < -1 ? -1 : a));
                break;
            }
            default:
            {
                double a = fabs(v) > 1e-3 ? c->length_rear * w / v : 0.0;
                double beta = asin(a > 0.99 ? 0.99 : (a < -0.99 ? -0.99 : a));
                u1 = atan(tan(beta) * (c->length_front + c->length_rear) / c->length_rear);
            }
        }
        ws->U[IX(0, k)] = v < lo[0] ? lo[0] : (v > hi[0] ? hi[0] : v);
        ws->U[IX(1, k)] = u1 < lo[1] ? lo[1] : (u1 > hi[1] ? hi[1] : u1);
    }
    /* control-rate rows: forward pass from u_prev, backward pass from u_ref = 0 */
    for (int i = 0; i < 2; ++i)
    {
        double dl = c->du_lb[i] > -MPCB200_INF ? c->du_lb[i] * ORC_INIT_SHRINK : -1e6;
        double dh = c->du_ub[i] < MPCB200_INF ? c->du_ub[i] * ORC_INIT_SHRINK : 1e6;
        double prev = p->u_prev[i], T = p->u_prev_dt;
        for (int k = 0; k <= N - 2; ++k)
        {
            if (!(k == 0 && T == 0.0))
            {
                double a = prev + dl * T, b = prev + dh * T;
                double u = ws->U[IX(i, k)];
                ws->U[IX(i, k)] = u < a ? a : (u > b ? b : u);
            }
            prev = ws->U[IX(i, k)];
            T = dt;
        }
        /* backward: (u_ref - u_{N-2})/dt in [dl, dh]  ->  u_{N-2} in [-dh dt, -dl dt]; then u_{k-1} in [u_k - dh dt, u_k - dl dt] */
        double next = 0.0;
        for (int k = N - 2; k >= 0; --k)
        {
            double a = next - dh * dt, b = next - dl * dt;
            double u = ws->U[IX(i, k)];
            ws->U[IX(i, k)] = u < a ? a : (u > b ? b : u);
            next = ws->U[IX(i, k)];
        }
    }
}

/* ======================================================================================================= */
/* Part 2: interior-point method                                                                            */
/* ======================================================================================================= */

/* Clipped slack steps: the fraction-to-the-boundary rule lets ONE nearly active row cut the step of all variables.  The rows
   that block most -- at most 1/ORC_CLIP_DIV of the rows, taken in whole sqrt(2)-wide bins of their step ratio -- are
   excluded from the rule; their slacks are clipped at ORC_CLIP_FLOOR times their value instead (their rows then lose less
   infeasibility than the Newton step promised, which the merit line search sees). */
#define ORC_CLIP_DIV 8
#define ORC_CLIP_FLOOR 0.01
#define ORC_CLIP_BINS 40
static int clip_bin(double ratio) /* bin j holds the ratios in (2^(-(j+1)/2), 2^(-j/2)] */
{
    int j = (int)floor(-2.0 * log2(ratio));
    return j < 0 ? 0 : (j >= ORC_CLIP_BINS ? ORC_CLIP_BINS - 1 : j);
}
#define ORC_MU_AUTO_MIN 0.1
#define ORC_MU_AUTO_MAX 1.0
#define ORC_KAPPA_EPS 10.0
#define ORC_KAPPA_MU 0.2
#define ORC_THETA_MU 1.5
#define ORC_TAU_MIN 0.99
#define ORC_SLACK_PUSH 1e-2
#define ORC_ARMIJO 1e-4
#define ORC_MAX_BACKTRACK 3
#define ORC_MAX_INERTIA_TRIES 2 /* factorisations per IPM iteration; escalation continues in the next iteration */
#define ORC_MAX_DELTA 1e8
#define ORC_DELTA_FLOOR 1e-5
#define ORC_TINY_STEP 1e-8
#define ORC_TINY_STEP_COUNT 2
#define ORC_KAPPA_SIGMA 1e10
#define ORC_SMAX 100.0

/* slack / multiplier initialisation: s = max(-g, push), lambda = mu/s, nu = 0 */
void orc_init_duals(const orc_problem* p, orc_ws* ws)
{
    const int N = ws->N, RS = ws->RS;
    double dt = ws->SCAL[MPCB200_SC_DT];
    double mu = p->cfg->mu_init;
    if (!(mu > 0.0))
    {
        /* automatic initial barrier parameter: the barrier mu * sum(ln s) has one term per inequality row; it is balanced
           against the objective at the initial guess, mu_0 = |f(x_0)| / m clamped to [0.1, 1] (0.1 = Ipopt's mu_init).
           Quadratic-form problems (f ~ 10^3) start at 1, minimum-time problems (f ~ 10^1) at 0.1. */
        int m = 0;
        for (int k = 0; k < N; ++k)
            for (int sl = 0; sl < RS; ++sl) m += row_active(p, ws, k, sl);
        const double f0 = fabs(orc_objective(p, ws, ws->X, ws->U, dt));
        mu = m > 0 ? f0 / (double)m : ORC_MU_AUTO_MIN;
        if (mu < ORC_MU_AUTO_MIN) mu = ORC_MU_AUTO_MIN;
        if (mu > ORC_MU_AUTO_MAX) mu = ORC_MU_AUTO_MAX;
    }
    for (int k = 0; k < N; ++k)
    {
        for (int sl = 0; sl < RS; ++sl)
        {
            double s = 1.0, lam = 0.0;
            if (row_active(p, ws, k, sl))
            {
                double g = row_value(p, ws, k, sl, ws->X, ws->U, dt, NULL, NULL);
                s = -g > ORC_SLACK_PUSH ? -g : ORC_SLACK_PUSH;
                lam = mu / s;
            }
            ws->S[IX(sl, k)] = s;
            ws->LAM[IX(sl, k)] = lam;
        }
        for (int i = 0; i < 3; ++i) ws->NU[IX(i, k)] = 0.0;
    }
    ws->SCAL[MPCB200_SC_MU] = mu;
    ws->SCAL[MPCB200_SC_RHO] = 1.0;
    ws->SCAL[MPCB200_SC_DELTA] = 0.0;
    ws->SCAL[MPCB200_SC_DELTA_LAST] = 0.0;
    ws->SCAL[MPCB200_SC_ITER] = 0.0;
    ws->SCAL[MPCB200_SC_STATUS] = -1.0;
    ws->SCAL[MPCB200_SC_NREG] = 0.0;
    ws->SCAL[MPCB200_SC_TINY] = 0.0;
}

static inline int hidx(int i, int j) /* upper-triangle packed index, i <= j */
{
    return i * 5 - (i * (i - 1)) / 2 + (j - i);
}
static inline void hadd(double* KKT, int N, int k, int i, int j, double v)
{
    if (i > j) { int t = i; i = j; j = t; }
    KKT[(MPCB200_K_H + hidx(i, j)) * N + k] += v;
}

/* extra evaluation outputs kept in the oracle only */
typedef struct {
    double dual_inf, prim_inf, sl_max, sl_min, sum_nu, sum_lam, inf1, barrier_log, obj;
    int m_eq, m_ineq;
} eval_info;

static eval_info g_last_info;

static void eval_impl(const orc_problem* p, orc_ws* ws, eval_info* info, double* G1 /*5xN + 1*/)
{
    const mpcb200_config* c = p->cfg;
    const int N = ws->N, RS = ws->RS;
    double* KKT = ws->KKT;
    double* GL = ws->GL;
    const double dt = ws->SCAL[MPCB200_SC_DT];
    memset(KKT, 0, sizeof(double) * KW * N);
    memset(GL, 0, sizeof(double) * 5 * N);
    memset(G1, 0, sizeof(double) * (5 * N + 1));
    double htt = 0.0, gt0 = 0.0, gl_dt = 0.0;
    double prim_inf = 0.0, inf1 = 0.0, sum_nu = 0.0;
    int m_eq = 0;

#define KK(f, k) KKT[((f)) * N + (k)]
    /* ---- cost + dynamics ---- */
    if (has_mintime(c)) { gt0 += (double)(N - 1); gl_dt += (double)(N - 1); }
    for (int k = 0; k <= N - 2; ++k)
    {
        double x[3] = {ws->X[IX(0, k)], ws->X[IX(1, k)], ws->X[IX(2, k)]};
        double u[2] = {ws->U[IX(0, k)], ws->U[IX(1, k)]};
        double nu[3] = {ws->NU[IX(0, k)], ws->NU[IX(1, k)], ws->NU[IX(2, k)]};
        double f[3], Jf[9], Hc[6];
        const int mid = is_midpoint(c);
        {
            double xn[3] = {ws->X[IX(0, k + 1)], ws->X[IX(1, k + 1)], ws->X[IX(2, k + 1)]}, xe[3];
            collocation_pose(c, x, xn, xe);
            orc_dynamics_derivs(c, xe, u, nu, f, Jf, Hc);
        }
        double e[3];
        e[0] = x[0] + dt * f[0] - ws->X[IX(0, k + 1)];
        e[1] = x[1] + dt * f[1] - ws->X[IX(1, k + 1)];
        e[2] = dt * f[2] - orc_normalize_theta(ws->X[IX(2, k + 1)] - x[2]);
        for (int i = 0; i < 3; ++i)
        {
            KK(MPCB200_K_E + i, k) = e[i];
            KK(MPCB200_K_D + i, k) = f[i];
            KK(MPCB200_K_A + i, k) = dt * Jf[i * 3 + 0];
            KK(MPCB200_K_B + 2 * i, k) = dt * Jf[i * 3 + 1];
            KK(MPCB200_K_B + 2 * i + 1, k) = dt * Jf[i * 3 + 2];
            if (fabs(e[i]) > prim_inf) prim_inf = fabs(e[i]);
            inf1 += fabs(e[i]);
            sum_nu += fabs(nu[i]);
        }
        m_eq += 3;
        if (mid)
        {
            /* Midpoint differences: the mean heading makes the defect depend on x_{k+1} too,
             *   de/dx_k = I + a e_th',  de/dx_{k+1} = -(I - a e_th'),  a = dt/2 df/dtheta  (a_theta = 0 for all models).
             * The linearised row is multiplied by (I - a e_th')^{-1} = I + a e_th' so that the record keeps the explicit form
             *   dx_{k+1} = (I + 2a e_th') dx_k + B~ du_k + d~ d(dt) + e~      (K_A = 2a is what the loop above stored)
             * the Riccati sweep works on; the multiplier of the transformed row is nu~ = (I - a e_th')' nu. */
            for (int i = 0; i < 2; ++i)
            {
                const double ai = 0.5 * KK(MPCB200_K_A + i, k);
                KK(MPCB200_K_B + 2 * i, k) += ai * KK(MPCB200_K_B + 4, k);
                KK(MPCB200_K_B + 2 * i + 1, k) += ai * KK(MPCB200_K_B + 5, k);
                KK(MPCB200_K_D + i, k) += ai * KK(MPCB200_K_D + 2, k);
                KK(MPCB200_K_E + i, k) += ai * KK(MPCB200_K_E + 2, k);
            }
        }
        /* quadratic cost */
        double gx[3] = {0, 0, 0}, gu[2] = {0, 0};
        if (has_quadratic(c))
        {
            double d[3] = {x[0] - p->xf[0], x[1] - p->xf[1], orc_normalize_theta(x[2] - p->xf[2])};
            const int integ = c->quadratic_integral_form != 0;
            const double fx = integ ? integral_state_weight(c, N, k) : 1.0;
            const double wx = integ ? dt * fx : 1.0, wu = integ ? dt : 1.0;
            double gxl[3] = {0, 0, 0}, gul[2] = {0, 0};
            for (int i = 0; i < 3; ++i)
            {
                for (int j = 0; j < 3; ++j) gxl[i] += (c->Q[i * 3 + j] + c->Q[j * 3 + i]) * d[j];
                for (int j = i; j < 3; ++j) hadd(KKT, N, k, i, j, wx * (c->Q[i * 3 + j] + c->Q[j * 3 + i]));
                gx[i] = wx * gxl[i];
            }
            for (int i = 0; i < 2; ++i)
            {
                for (int j = 0; j < 2; ++j) gul[i] += (c->R[i * 2 + j] + c->R[j * 2 + i]) * u[j];
                for (int j = i; j < 2; ++j) hadd(KKT, N, k, 3 + i, 3 + j, wu * (c->R[i * 2 + j] + c->R[j * 2 + i]));
                gu[i] = wu * gul[i];
            }
            if (integ && c->variable_dt)
            {
                /* d/ddt of dt * (w l_x + l_u) is w l_x + l_u; the w-dt cross Hessian is its gradient */
                const double l = fx * quad3(c->Q, d) + quad2(c->R, u);
                gt0 += l; gl_dt += l;
                for (int i = 0; i < 3; ++i) KK(MPCB200_K_HB + i, k) += fx * gxl[i];
                for (int i = 0; i < 2; ++i) KK(MPCB200_K_HB + 3 + i, k) += gul[i];
            }
        }
        for (int i = 0; i < 3; ++i) { KK(MPCB200_K_G + i, k) += gx[i]; GL[IX(i, k)] += gx[i]; }
        for (int i = 0; i < 2; ++i) { KK(MPCB200_K_G + 3 + i, k) += gu[i]; GL[IX(3 + i, k)] += gu[i]; }
        /* Lagrangian terms of the dynamics: nu_k^T e_k */
        double fx_nu = nu[0] * Jf[0] + nu[1] * Jf[3] + nu[2] * Jf[6];
        double fu_nu[2] = {nu[0] * Jf[1] + nu[1] * Jf[4] + nu[2] * Jf[7], nu[0] * Jf[2] + nu[1] * Jf[5] + nu[2] * Jf[8]};
        /* the heading of x_k enters the dynamics with weight wth (1, or 1/2 through the mean heading) */
        const double wth = mid ? 0.5 : 1.0;
        GL[IX(0, k)] += nu[0];
        GL[IX(1, k)] += nu[1];
        GL[IX(2, k)] += nu[2] + wth * dt * fx_nu;
        GL[IX(3, k)] += dt * fu_nu[0];
        GL[IX(4, k)] += dt * fu_nu[1];
        for (int i = 0; i < 3; ++i) GL[IX(i, k + 1)] -= nu[i];
        gl_dt += nu[0] * f[0] + nu[1] * f[1] + nu[2] * f[2];
        hadd(KKT, N, k, 2, 2, wth * wth * dt * Hc[0]);
        hadd(KKT, N, k, 2, 3, wth * dt * Hc[1]);
        hadd(KKT, N, k, 2, 4, wth * dt * Hc[2]);
        hadd(KKT, N, k, 3, 3, dt * Hc[3]);
        hadd(KKT, N, k, 3, 4, dt * Hc[4]);
        hadd(KKT, N, k, 4, 4, dt * Hc[5]);
        if (c->variable_dt)
        {
            KK(MPCB200_K_HB + 2, k) += wth * fx_nu;
            KK(MPCB200_K_HB + 3, k) += fu_nu[0];
            KK(MPCB200_K_HB + 4, k) += fu_nu[1];
        }
        if (mid)
        {
            /* the other half of the mean heading belongs to theta_{k+1} */
            GL[IX(2, k + 1)] += 0.5 * dt * fx_nu;
            hadd(KKT, N, k + 1, 2, 2, 0.25 * dt * Hc[0]);
            if (c->variable_dt) KK(MPCB200_K_HB + 2, k + 1) += 0.5 * fx_nu;
            /* ... and the Lagrangian has a Hessian block between theta_{k+1} and w_k = (theta_k, u_k),
             *   q = (dt/4 Hc_tt, dt/2 Hc_t0, dt/2 Hc_t1),
             * which the stage-wise record cannot hold.  It is condensed into stage k with the linearised heading row
             *   d(theta_{k+1}) = r' dw_k + e~_2 + d~_2 d(dt),  r = (1, B~_20, B~_21):
             * H_k += q r' + r q', Newton gradient += q e~_2, dt border += q d~_2 -- the same Newton step; the multiplier the
             * sweep returns is nu+ - e_th (q' dw_k) (undone in orc_kkt_solve). */
            const double q[3] = {0.25 * dt * Hc[0], 0.5 * dt * Hc[1], 0.5 * dt * Hc[2]};
            const double r[3] = {1.0, KK(MPCB200_K_B + 4, k), KK(MPCB200_K_B + 5, k)};
            for (int i = 0; i < 3; ++i)
            {
                for (int j = i; j < 3; ++j) hadd(KKT, N, k, 2 + i, 2 + j, q[i] * r[j] + r[i] * q[j]);
                KK(MPCB200_K_G + 2 + i, k) += q[i] * KK(MPCB200_K_E + 2, k);
                if (c->variable_dt) KK(MPCB200_K_HB + 2 + i, k) += q[i] * KK(MPCB200_K_D + 2, k);
            }
        }
    }
    /* terminal cost */
    if (has_terminal_cost(c))
    {
        int k = N - 1;
        double d[3] = {ws->X[IX(0, k)] - p->xf[0], ws->X[IX(1, k)] - p->xf[1], orc_normalize_theta(ws->X[IX(2, k)] - p->xf[2])};
        for (int i = 0; i < 3; ++i)
        {
            double g = 0;
            for (int j = 0; j < 3; ++j) g += (c->Qf[i * 3 + j] + c->Qf[j * 3 + i]) * d[j];
            KK(MPCB200_K_G + i, k) += g;
            GL[IX(i, k)] += g;
            for (int j = i; j < 3; ++j) hadd(KKT, N, k, i, j, c->Qf[i * 3 + j] + c->Qf[j * 3 + i]);
        }
    }
    /* end term of the trapezoidal rule: dt/2 l_x(x_{N-1}) */
    if (has_trapezoid(c))
    {
        int k = N - 1;
        const double fx = integral_state_weight(c, N, k);
        double d[3] = {ws->X[IX(0, k)] - p->xf[0], ws->X[IX(1, k)] - p->xf[1], orc_normalize_theta(ws->X[IX(2, k)] - p->xf[2])};
        for (int i = 0; i < 3; ++i)
        {
            double g = 0;
            for (int j = 0; j < 3; ++j) g += (c->Q[i * 3 + j] + c->Q[j * 3 + i]) * d[j];
            KK(MPCB200_K_G + i, k) += dt * fx * g;
            GL[IX(i, k)] += dt * fx * g;
            for (int j = i; j < 3; ++j) hadd(KKT, N, k, i, j, dt * fx * (c->Q[i * 3 + j] + c->Q[j * 3 + i]));
            if (c->variable_dt) KK(MPCB200_K_HB + i, k) += fx * g;
        }
        if (c->variable_dt) { gt0 += fx * quad3(c->Q, d); gl_dt += fx * quad3(c->Q, d); }
    }
    /* via-points */
    if (has_viapoints(c))
    {
        for (int j = 0; j < p->n_vp && j < 64; ++j)
        {
            int k = ws->vp_stage[j];
            if (k < 0) continue;
            double w = c->vp_position_weight;
            double ex = p->vp[3 * j] - ws->X[IX(0, k)], ey = p->vp[3 * j + 1] - ws->X[IX(1, k)];
            KK(MPCB200_K_G + 0, k) += -2 * w * ex; GL[IX(0, k)] += -2 * w * ex;
            KK(MPCB200_K_G + 1, k) += -2 * w * ey; GL[IX(1, k)] += -2 * w * ey;
            hadd(KKT, N, k, 0, 0, 2 * w);
            hadd(KKT, N, k, 1, 1, 2 * w);
            if (c->vp_orientation_weight > 0)
            {
                KK(MPCB200_K_G + 2, k) += -c->vp_orientation_weight;
                GL[IX(2, k)] += -c->vp_orientation_weight;
            }
        }
    }
    /* ---- inequality rows ---- */
    double sl_max = -1e300, sl_min = 1e300, sum_lam = 0.0, blog = 0.0, gt1 = 0.0;
    int m_ineq = 0;
    for (int k = 0; k < N; ++k)
    {
        for (int sl = 0; sl < RS; ++sl)
        {
            if (!row_active(p, ws, k, sl)) { ws->G[IX(sl, k)] = 0.0; continue; }
            double grad[8], h6[6];
            const int state_row = sl >= 8 || (k == N - 1 && sl == 2); /* rows on x_k: obstacles, terminal ball */
            double hdt[4] = {0, 0, 0, 0};
            g_row_hdt = sl >= 8 ? hdt : NULL;
            double g = row_value(p, ws, k, sl, ws->X, ws->U, dt, grad, state_row ? h6 : NULL);
            g_row_hdt = NULL;
            ws->G[IX(sl, k)] = g;
            double s = ws->S[IX(sl, k)], lam = ws->LAM[IX(sl, k)];
            double r = g + s, sig = lam / s;
            if (fabs(r) > prim_inf) prim_inf = fabs(r);
            inf1 += fabs(r);
            if (s * lam > sl_max) sl_max = s * lam;
            if (s * lam < sl_min) sl_min = s * lam;
            sum_lam += fabs(lam);
            blog += log(s);
            ++m_ineq;
            double c0 = sig * r; /* mu-independent part of gamma = mu/s + sig r */
            double c1 = 1.0 / s; /* coefficient of mu */
            /* x_k part (obstacle rows, terminal ball) */
            if (state_row)
            {
                for (int i = 0; i < 3; ++i)
                {
                    KK(MPCB200_K_G + i, k) += c0 * grad[i];
                    G1[IX(i, k)] += c1 * grad[i];
                    GL[IX(i, k)] += lam * grad[i];
                }
                int q = 0;
                for (int i = 0; i < 3; ++i)
                    for (int j = i; j < 3; ++j, ++q) hadd(KKT, N, k, i, j, lam * h6[q] + sig * grad[i] * grad[j]);
                if (grad[7] != 0.0)
                {
                    /* dynamic obstacle with a free dt: the row also depends on dt */
                    for (int i = 0; i < 3; ++i) KK(MPCB200_K_HB + i, k) += lam * hdt[i] + sig * grad[i] * grad[7];
                    gt0 += c0 * grad[7];
                    gt1 += c1 * grad[7];
                    gl_dt += lam * grad[7];
                    htt += lam * hdt[3] + sig * grad[7] * grad[7];
                }
                continue;
            }
            /* u_k part */
            for (int i = 0; i < 2; ++i)
            {
                double gi = grad[3 + i];
                if (gi == 0.0 || k > N - 2) continue;
                KK(MPCB200_K_G + 3 + i, k) += c0 * gi;
                G1[IX(3 + i, k)] += c1 * gi;
                GL[IX(3 + i, k)] += lam * gi;
                hadd(KKT, N, k, 3 + i, 3 + i, sig * gi * gi);
                if (grad[7] != 0.0) KK(MPCB200_K_HB + 3 + i, k) += sig * gi * grad[7];
                if (grad[5 + i] != 0.0) KK(MPCB200_K_C + i, k) += sig * gi * grad[5 + i];
            }
            /* u_{k-1} part */
            for (int i = 0; i < 2; ++i)
            {
                double gi = grad[5 + i];
                if (gi == 0.0 || k < 1) continue;
                KK(MPCB200_K_G + 3 + i, k - 1) += c0 * gi;
                G1[IX(3 + i, k - 1)] += c1 * gi;
                GL[IX(3 + i, k - 1)] += lam * gi;
                hadd(KKT, N, k - 1, 3 + i, 3 + i, sig * gi * gi);
                if (grad[7] != 0.0) KK(MPCB200_K_HB + 3 + i, k - 1) += sig * gi * grad[7];
            }
            /* dt part */
            if (grad[7] != 0.0)
            {
                gt0 += c0 * grad[7];
                gt1 += c1 * grad[7];
                gl_dt += lam * grad[7];
                htt += sig * grad[7] * grad[7];
            }
        }
    }
    G1[5 * N] = gt1;
    /* ---- errors (Ipopt's scaled optimality error) ---- */
    double dual_inf = 0.0;
    for (int k = 0; k < N; ++k)
    {
        for (int i = 0; i < 5; ++i)
        {
            if (i < 3 && k == 0) continue;                     /* x_0 fixed */
            if (i < 3 && k == N - 1 && c->xf_fixed[i]) continue; /* fixed terminal components */
            if (i >= 3 && k == N - 1) continue;
            double v = fabs(GL[IX(i, k)]);
            if (v > dual_inf) dual_inf = v;
        }
    }
    if (c->variable_dt && fabs(gl_dt) > dual_inf) dual_inf = fabs(gl_dt);
    ws->gl_dt = gl_dt;
    ws->SCAL[MPCB200_SC_HTT] = htt;
    ws->SCAL[MPCB200_SC_GT] = gt0;
    info->dual_inf = dual_inf; info->prim_inf = prim_inf; info->sl_max = sl_max; info->sl_min = sl_min;
    info->sum_nu = sum_nu; info->sum_lam = sum_lam; info->inf1 = inf1; info->barrier_log = blog;
    info->m_eq = m_eq; info->m_ineq = m_ineq;
    info->obj = orc_objective(p, ws, ws->X, ws->U, dt);
    ws->SCAL[MPCB200_SC_OBJ] = info->obj;
    ws->SCAL[MPCB200_SC_INF] = inf1;
    ws->SCAL[MPCB200_SC_BLOG] = blog;
    ws->SCAL[MPCB200_SC_GLDT] = gl_dt;
#undef KK
}

static double scaled_error(const eval_info* in, double mu)
{
    double sd = (in->sum_nu + in->sum_lam) / (double)(in->m_eq + in->m_ineq > 0 ? in->m_eq + in->m_ineq : 1);
    sd = (sd > ORC_SMAX ? sd : ORC_SMAX) / ORC_SMAX;
    double sc = in->m_ineq > 0 ? in->sum_lam / (double)in->m_ineq : 0.0;
    sc = (sc > ORC_SMAX ? sc : ORC_SMAX) / ORC_SMAX;
    double compl = 0.0;
    if (in->m_ineq > 0)
    {
        double a = in->sl_max - mu, b = mu - in->sl_min;
        compl = (a > b ? a : b);
        if (compl < 0) compl = 0;
    }
    double e = in->dual_inf / sd;
    if (in->prim_inf > e) e = in->prim_inf;
    if (compl / sc > e) e = compl / sc;
    return e;
}

/* finalize the condensed gradient with the barrier parameter: g = g0 + mu * g1 */
static void finalize_gradient(orc_ws* ws, const double* G1, double mu)
{
    const int N = ws->N;
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < 5; ++i) ws->KKT[(MPCB200_K_G + i) * N + k] += mu * G1[IX(i, k)];
    ws->SCAL[MPCB200_SC_GT] += mu * G1[5 * N];
}

/* public eval: records + errors; the barrier update rule is applied here so that the records carry the final mu */
static void eval_and_update_mu(const orc_problem* p, orc_ws* ws, eval_info* info, int allow_mu_update)
{
    const int N = ws->N;
    double* G1 = (double*)malloc(sizeof(double) * (5 * N + 1));
    eval_impl(p, ws, info, G1);
    double mu = ws->SCAL[MPCB200_SC_MU];
    double tol = p->cfg->tol;
    double mu_min = tol / 10.0;
    double e0 = scaled_error(info, 0.0);
    double emu = scaled_error(info, mu);
    if (allow_mu_update && e0 > tol)
    {
        /* monotone Fiacco-McCormick update (Ipopt's mu_strategy monotone) */
        while (emu <= ORC_KAPPA_EPS * mu && mu > mu_min)
        {
            double m1 = ORC_KAPPA_MU * mu, m2 = pow(mu, ORC_THETA_MU);
            mu = m1 < m2 ? m1 : m2;
            if (mu < mu_min) mu = mu_min;
            emu = scaled_error(info, mu);
        }
    }
    ws->SCAL[MPCB200_SC_MU] = mu;
    ws->SCAL[MPCB200_SC_ERR0] = e0;
    ws->SCAL[MPCB200_SC_ERRMU] = emu;
    finalize_gradient(ws, G1, mu);
    free(G1);
}

void orc_eval(const orc_problem* p, orc_ws* ws)
{
    eval_info info;
    eval_and_update_mu(p, ws, &info, 1);
    g_last_info = info;
}

/* ---- Riccati solve of the bordered block-tridiagonal KKT system ---------------------------------------- */
/*
 * Unknowns: dw_k = (dx_k, du_k), k = 0..N-2, dx_{N-1}, d(dt); multipliers nu+_k of
 *      dx_{k+1} = A_k dx_k + B_k du_k + d_k d(dt) + e_k,   dx_0 = 0, dx_{N-1,j} = 0 for fixed terminal components.
 * Stage state y_k = (dx_k, du_{k-1}) in R^5 (the previous control enters through the rate cross block C_k),
 * parameters theta_hat = (1, d(dt), pi_0, pi_1, pi_2): pi_j are the multipliers of the fixed terminal components.
 * Value function V_k(y) = 1/2 y'P y + y' PI theta_hat + 1/2 theta_hat' TH theta_hat  (DESIGN.md "Riccati").
 * Returns 1 if the inertia is wrong (some M_vv not positive definite, or the reduced (dt, pi) system has the wrong signs).
 */
static void mat_zero(double* a, int n) { for (int i = 0; i < n; ++i) a[i] = 0.0; }

int orc_kkt_solve(const orc_problem* p, orc_ws* ws, double delta)
{
    const mpcb200_config* c = p->cfg;
    const int N = ws->N;
    const double* KKT = ws->KKT;
#define KK(f, k) KKT[((f)) * N + (k)]
    double P[25], PI[25], TH[25];
    mat_zero(P, 25); mat_zero(PI, 25); mat_zero(TH, 25);
    const int dt_free = c->variable_dt;
    /* terminal stage k = N-1: y = (dx_{N-1}, du_{N-2}) */
    {
        int k = N - 1;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j)
            {
                double v = KK(MPCB200_K_H + hidx(i, j), k) + (i == j ? delta : 0.0);
                if (c->xf_fixed[i] || c->xf_fixed[j]) v = 0.0;
                P[i * 5 + j] = v; P[j * 5 + i] = v;
            }
        for (int i = 0; i < 3; ++i)
        {
            PI[i * 5 + 0] = c->xf_fixed[i] ? 0.0 : KK(MPCB200_K_G + i, k);
            if (c->xf_fixed[i]) PI[i * 5 + 2 + i] = 1.0;
            /* x_{N-1} - dt cross term: end term of the trapezoidal cost rule, heading of the last midpoint defect */
            if (dt_free && (has_trapezoid(c) || is_midpoint(c)) && !c->xf_fixed[i]) PI[i * 5 + 1] = KK(MPCB200_K_HB + i, k);
        }
        TH[0 * 5 + 1] = TH[1 * 5 + 0] = ws->SCAL[MPCB200_SC_GT];
        TH[1 * 5 + 1] = ws->SCAL[MPCB200_SC_HTT] + delta;
    }
    memcpy(ws->P + 25 * (N - 1), P, sizeof(P));
    memcpy(ws->PI + 25 * (N - 1), PI, sizeof(PI));
    for (int k = N - 2; k >= 0; --k)
    {
        /* F = [Abar Bbar] (5x7), Chat (5x5) */
        double F[5][7], Ch[5][5];
        memset(F, 0, sizeof(F)); memset(Ch, 0, sizeof(Ch));
        for (int i = 0; i < 3; ++i)
        {
            F[i][i] = 1.0;
            F[i][2] += KK(MPCB200_K_A + i, k);
            F[i][5] = KK(MPCB200_K_B + 2 * i, k);
            F[i][6] = KK(MPCB200_K_B + 2 * i + 1, k);
            Ch[i][0] = KK(MPCB200_K_E + i, k);
            if (dt_free) Ch[i][1] = KK(MPCB200_K_D + i, k);
        }
        F[3][5] = 1.0; F[4][6] = 1.0;
        /* M (7x7) in z = (x(0..2), uprev(3..4), u(5..6)) */
        double M[7][7], Mh[7][5];
        memset(M, 0, sizeof(M)); memset(Mh, 0, sizeof(Mh));
        static const int map[5] = {0, 1, 2, 5, 6};
        for (int i = 0; i < 5; ++i)
            for (int j = i; j < 5; ++j)
            {
                double v = KK(MPCB200_K_H + hidx(i, j), k) + (i == j ? delta : 0.0);
                M[map[i]][map[j]] = v; M[map[j]][map[i]] = v;
            }
        for (int i = 0; i < 2; ++i) { M[3 + i][5 + i] = KK(MPCB200_K_C + i, k); M[5 + i][3 + i] = KK(MPCB200_K_C + i, k); }
        for (int i = 0; i < 5; ++i)
        {
            Mh[map[i]][0] = KK(MPCB200_K_G + i, k);
            if (dt_free) Mh[map[i]][1] = KK(MPCB200_K_HB + i, k);
        }
        /* PF = P F (5x7); W = P Chat + PI (5x5) */
        double PF[5][7], W[5][5];
        for (int i = 0; i < 5; ++i)
        {
            for (int j = 0; j < 7; ++j) { double s = 0; for (int l = 0; l < 5; ++l) s += P[i * 5 + l] * F[l][j]; PF[i][j] = s; }
            for (int j = 0; j < 5; ++j) { double s = PI[i * 5 + j]; for (int l = 0; l < 5; ++l) s += P[i * 5 + l] * Ch[l][j]; W[i][j] = s; }
        }
        /* MM = M + F'PF ; NN = Mh + F' W ; TT = TH + Chat' (P Chat + PI) + PI' Chat */
        double MM[7][7], NN[7][5], TT[5][5];
        for (int i = 0; i < 7; ++i)
        {
            for (int j = 0; j < 7; ++j) { double s = M[i][j]; for (int l = 0; l < 5; ++l) s += F[l][i] * PF[l][j]; MM[i][j] = s; }
            for (int j = 0; j < 5; ++j) { double s = Mh[i][j]; for (int l = 0; l < 5; ++l) s += F[l][i] * W[l][j]; NN[i][j] = s; }
        }
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j)
            {
                double s = TH[i * 5 + j];
                for (int l = 0; l < 5; ++l) s += Ch[l][i] * W[l][j] + PI[l * 5 + i] * Ch[l][j];
                TT[i][j] = s;
            }
        /* eliminate v = z[5..6] */
        double a = MM[5][5], b = 0.5 * (MM[5][6] + MM[6][5]), d = MM[6][6];
        double det = a * d - b * b;
        if (!(a > 0.0) || !(det > 1e-14 * a * (d > 0 ? d : 1.0)) || !(d > 0.0))
            return 1;
        double i00 = d / det, i01 = -b / det, i11 = a / det;
        double KGm[2][5], KTm[2][5]; /* Mvv^-1 Mvy, Mvv^-1 Nv */
        for (int j = 0; j < 5; ++j)
        {
            KGm[0][j] = i00 * MM[5][j] + i01 * MM[6][j];
            KGm[1][j] = i01 * MM[5][j] + i11 * MM[6][j];
            KTm[0][j] = i00 * NN[5][j] + i01 * NN[6][j];
            KTm[1][j] = i01 * NN[5][j] + i11 * NN[6][j];
        }
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j)
            {
                P[i * 5 + j] = MM[i][j] - (MM[i][5] * KGm[0][j] + MM[i][6] * KGm[1][j]);
                PI[i * 5 + j] = NN[i][j] - (MM[i][5] * KTm[0][j] + MM[i][6] * KTm[1][j]);
                TH[i * 5 + j] = TT[i][j] - (NN[5][i] * KTm[0][j] + NN[6][i] * KTm[1][j]);
            }
        /* symmetrise P */
        for (int i = 0; i < 5; ++i)
            for (int j = i + 1; j < 5; ++j) { double s = 0.5 * (P[i * 5 + j] + P[j * 5 + i]); P[i * 5 + j] = s; P[j * 5 + i] = s; }
        memcpy(ws->P + 25 * k, P, sizeof(P));
        memcpy(ws->PI + 25 * k, PI, sizeof(PI));
        memcpy(ws->KG + 10 * k, KGm, sizeof(KGm));
        memcpy(ws->KT + 10 * k, KTm, sizeof(KTm));
    }
    /* root: y_0 = 0 -> stationarity of 1/2 th' TH th over the active theta entries.  Eliminate pi (negative
       definite block) first, then dt (must leave a positive pivot). */
    double th[5] = {1.0, 0, 0, 0, 0};
    {
        int act[4], na = 0; /* active pi entries */
        for (int j = 0; j < 3; ++j) if (c->xf_fixed[j]) act[na++] = 2 + j;
        /* negative-definite Cholesky of the pi block: -TH_pipi = L L' */
        double L[3][3]; memset(L, 0, sizeof(L));
        for (int i = 0; i < na; ++i)
            for (int j = 0; j <= i; ++j)
            {
                double s = -0.5 * (TH[act[i] * 5 + act[j]] + TH[act[j] * 5 + act[i]]);
                for (int l = 0; l < j; ++l) s -= L[i][l] * L[j][l];
                if (i == j) { if (!(s > 0.0)) return 1; L[i][i] = sqrt(s); }
                else L[i][j] = s / L[j][j];
            }
        /* solve for pi as affine function of (1, dt): TH_pipi pi = -(TH_pi0 + TH_pidt ddt) */
        double rhs0[3], rhs1[3], sol0[3], sol1[3];
        for (int i = 0; i < na; ++i) { rhs0[i] = TH[act[i] * 5 + 0]; rhs1[i] = TH[act[i] * 5 + 1]; }
        /* (-TH_pipi) pi = rhs  ->  L L' pi = rhs */
        for (int pass = 0; pass < 2; ++pass)
        {
            double* r = pass ? rhs1 : rhs0; double* s = pass ? sol1 : sol0;
            double y[3];
            for (int i = 0; i < na; ++i) { double t = r[i]; for (int l = 0; l < i; ++l) t -= L[i][l] * y[l]; y[i] = t / L[i][i]; }
            for (int i = na - 1; i >= 0; --i) { double t = y[i]; for (int l = i + 1; l < na; ++l) t -= L[l][i] * s[l]; s[i] = t / L[i][i]; }
        }
        double ddt = 0.0;
        if (dt_free)
        {
            /* reduced dt equation: (TH_tt + TH_tpi sol1) ddt = -(TH_t0 + TH_tpi sol0) */
            double htt = TH[1 * 5 + 1], gt = TH[1 * 5 + 0];
            for (int i = 0; i < na; ++i) { htt += TH[1 * 5 + act[i]] * sol1[i]; gt += TH[1 * 5 + act[i]] * sol0[i]; }
            if (!(htt > 0.0)) return 1;
            ddt = -gt / htt;
        }
        th[1] = ddt;
        for (int i = 0; i < na; ++i) th[act[i]] = sol0[i] + sol1[i] * ddt;
    }
    /* forward pass */
    double y[5] = {0, 0, 0, 0, 0};
    double* STEP = ws->STEP;
    for (int k = 0; k <= N - 2; ++k)
    {
        const double* KGm = ws->KG + 10 * k; const double* KTm = ws->KT + 10 * k;
        double v[2];
        for (int i = 0; i < 2; ++i)
        {
            double s = 0;
            for (int j = 0; j < 5; ++j) s -= KGm[i * 5 + j] * y[j] + KTm[i * 5 + j] * th[j];
            v[i] = s;
        }
        for (int i = 0; i < 3; ++i) STEP[IX(i, k)] = y[i];
        STEP[IX(3, k)] = v[0]; STEP[IX(4, k)] = v[1];
        double yn[5];
        for (int i = 0; i < 3; ++i)
        {
            yn[i] = y[i] + KK(MPCB200_K_A + i, k) * y[2] + KK(MPCB200_K_B + 2 * i, k) * v[0] + KK(MPCB200_K_B + 2 * i + 1, k) * v[1] +
                    KK(MPCB200_K_E + i, k) + (dt_free ? KK(MPCB200_K_D + i, k) * th[1] : 0.0);
        }
        yn[3] = v[0]; yn[4] = v[1];
        /* nu+_k = (P_{k+1} y_{k+1} + PI_{k+1} theta_hat)[0:3] */
        const double* Pn = ws->P + 25 * (k + 1); const double* PIn = ws->PI + 25 * (k + 1);
        for (int i = 0; i < 3; ++i)
        {
            double s = 0;
            for (int j = 0; j < 5; ++j) s += Pn[i * 5 + j] * yn[j] + PIn[i * 5 + j] * th[j];
            STEP[IX(5 + i, k)] = s;
        }
        memcpy(y, yn, sizeof(y));
    }
    for (int i = 0; i < 3; ++i) STEP[IX(i, N - 1)] = y[i];
    for (int i = 3; i < 8; ++i) STEP[IX(i, N - 1)] = 0.0;
    if (is_midpoint(c))
    {
        /* midpoint differences (orc_eval): the sweep worked on the transformed, condensed system,
         *   nu+ = (I + e_th a') nu~ + e_th (q' dw_k),  a = K_A / 2,  q from the Hessian of nu' f at the current iterate */
        const double dtc = ws->SCAL[MPCB200_SC_DT];
        for (int k = 0; k <= N - 2; ++k)
        {
            double x[3] = {ws->X[IX(0, k)], ws->X[IX(1, k)], ws->X[IX(2, k)]};
            double xn[3] = {ws->X[IX(0, k + 1)], ws->X[IX(1, k + 1)], ws->X[IX(2, k + 1)]}, xe[3];
            double u[2] = {ws->U[IX(0, k)], ws->U[IX(1, k)]};
            double nu[3] = {ws->NU[IX(0, k)], ws->NU[IX(1, k)], ws->NU[IX(2, k)]};
            double f[3], Jf[9], Hc[6];
            collocation_pose(c, x, xn, xe);
            orc_dynamics_derivs(c, xe, u, nu, f, Jf, Hc);
            double corr = 0.5 * KK(MPCB200_K_A + 0, k) * STEP[IX(5, k)] + 0.5 * KK(MPCB200_K_A + 1, k) * STEP[IX(6, k)];
            corr += 0.25 * dtc * Hc[0] * STEP[IX(2, k)] + 0.5 * dtc * (Hc[1] * STEP[IX(3, k)] + Hc[2] * STEP[IX(4, k)]);
            STEP[IX(7, k)] += corr;
        }
    }
    ws->SCAL[MPCB200_SC_DDT] = th[1];
    ws->SCAL[MPCB200_SC_DELTA] = delta;
#undef KK
    return 0;
}

/* merit pieces at a trial point: objective, l1 infeasibility, sum log s */
static void trial_eval(const orc_problem* p, orc_ws* ws, const double* Xt, const double* Ut, double dtt, double alpha,
                       double* obj, double* inf1, double* blog)
{
    const mpcb200_config* c = p->cfg;
    const int N = ws->N, RS = ws->RS;
    double inf = 0.0, bl = 0.0;
    for (int k = 0; k <= N - 2; ++k)
    {
        double x1[3] = {Xt[IX(0, k)], Xt[IX(1, k)], Xt[IX(2, k)]}, u1[2] = {Ut[IX(0, k)], Ut[IX(1, k)]};
        double x2[3] = {Xt[IX(0, k + 1)], Xt[IX(1, k + 1)], Xt[IX(2, k + 1)]}, e[3];
        orc_defect(c, x1, u1, x2, dtt, e);
        inf += fabs(e[0]) + fabs(e[1]) + fabs(e[2]);
    }
    for (int k = 0; k < N; ++k)
        for (int sl = 0; sl < RS; ++sl)
        {
            if (!row_active(p, ws, k, sl)) continue;
            double g = row_value(p, ws, k, sl, Xt, Ut, dtt, NULL, NULL);
            double s = ws->S[IX(sl, k)] + alpha * ws->DS[IX(sl, k)];
            if (s < ORC_CLIP_FLOOR * ws->S[IX(sl, k)]) s = ORC_CLIP_FLOOR * ws->S[IX(sl, k)];
            inf += fabs(g + s);
            bl += log(s);
        }
    *obj = orc_objective(p, ws, Xt, Ut, dtt);
    *inf1 = inf;
    *blog = bl;
}


static int orc_solve_monotone(const orc_problem* p, orc_ws* ws, orc_result* res, int iter0);
int orc_solve(const orc_problem* p, orc_ws* ws, orc_result* res)
{
    return orc_solve_monotone(p, ws, res, 0);
}
static int orc_solve_monotone(const orc_problem* p, orc_ws* ws, orc_result* res, int iter0)
{
    const mpcb200_config* c = p->cfg;
    const int N = ws->N, RS = ws->RS;
    int status = MPCB200_STATUS_MAX_ITER;
    int iter = iter0, nreg = 0, nbt = 0;
    eval_info info;
    for (;;)
    {
        eval_and_update_mu(p, ws, &info, 1);
        g_last_info = info;
        ws->SCAL[MPCB200_SC_ITER] = (double)iter;
        if (ws->SCAL[MPCB200_SC_ERR0] <= c->tol) { status = MPCB200_STATUS_CONVERGED; break; }
        if (iter >= c->max_iter) { status = MPCB200_STATUS_MAX_ITER; break; }
        const double mu = ws->SCAL[MPCB200_SC_MU];
        const double dt = ws->SCAL[MPCB200_SC_DT];
        /* ---- Newton step with inertia correction (Ipopt Algorithm IC) ---- */
        double delta = 0.0, dlast = ws->SCAL[MPCB200_SC_DELTA_LAST];
        /* after a regularised iteration the first attempt is delta_last/3 (not 0): such instances need regularisation
           again ~70% of the time and every failed attempt costs a full sweep; the value decays back to 0 */
        if (dlast > 0.0)
        {
            delta = dlast / 3.0;
            if (delta < ORC_DELTA_FLOOR) delta = 0.0;
        }
        int ok = 0;
        for (int tries = 0; tries < ORC_MAX_INERTIA_TRIES; ++tries)
        {
            if (orc_kkt_solve(p, ws, delta) == 0) { ok = 1; break; }
            ++nreg;
            if (delta == 0.0) delta = (dlast == 0.0) ? 1e-4 : (dlast / 3.0 > 1e-20 ? dlast / 3.0 : 1e-20);
            else delta *= (dlast == 0.0 ? 100.0 : 8.0);
            if (delta > ORC_MAX_DELTA) break;
        }
        if (!ok && delta <= ORC_MAX_DELTA)
        {
            /* deferred escalation: the factorisation budget of this iteration is spent (on the GPU every extra sweep of
               one instance stalls the whole launch).  Null step; the next iteration resumes at this delta
               (DELTA_LAST / 3 is its first attempt). */
            ws->SCAL[MPCB200_SC_DELTA_LAST] = 3.0 * delta;
            ws->SCAL[MPCB200_SC_DEFER] = 0.0; /* (the GPU path raises it in the KKT phase and clears it in the line-search phase) */
            ws->SCAL[MPCB200_SC_ALPHA] = 0.0;
            ++iter;
            continue;
        }
        ws->SCAL[MPCB200_SC_DEFER] = 0.0;
        if (!ok) { status = MPCB200_STATUS_NUMERICAL_ERROR; break; }
        if (delta > 0.0) ws->SCAL[MPCB200_SC_DELTA_LAST] = delta;
        else ws->SCAL[MPCB200_SC_DELTA_LAST] = 0.0;
        const double ddt = ws->SCAL[MPCB200_SC_DDT];
        /* ---- slack / multiplier steps and fraction to the boundary ---- */
        const double tau = (1.0 - mu > ORC_TAU_MIN) ? 1.0 - mu : ORC_TAU_MIN;
        double a_p = 1.0, a_d = 1.0;
        double dphi_bar = 0.0; /* -mu sum ds/s */
        double curv = 0.0;     /* sum sigma ds^2 */
        int hist[ORC_CLIP_BINS], m_rows = 0;
        for (int j = 0; j < ORC_CLIP_BINS; ++j) hist[j] = 0;
        for (int k = 0; k < N; ++k)
            for (int sl = 0; sl < RS; ++sl)
            {
                if (!row_active(p, ws, k, sl)) { ws->DS[IX(sl, k)] = 0; ws->DLAM[IX(sl, k)] = 0; continue; }
                ++m_rows;
                double grad[8];
                double g = row_value(p, ws, k, sl, ws->X, ws->U, dt, grad, NULL);
                double gdz = 0.0;
                for (int i = 0; i < 3; ++i) gdz += grad[i] * ws->STEP[IX(i, k)];
                if (k <= N - 2) for (int i = 0; i < 2; ++i) gdz += grad[3 + i] * ws->STEP[IX(3 + i, k)];
                if (k >= 1) for (int i = 0; i < 2; ++i) gdz += grad[5 + i] * ws->STEP[IX(3 + i, k - 1)];
                gdz += grad[7] * ddt;
                double s = ws->S[IX(sl, k)], lam = ws->LAM[IX(sl, k)];
                double ds = -(g + s) - gdz;
                double dl = mu / s - lam - (lam / s) * ds;
                ws->DS[IX(sl, k)] = ds;
                ws->DLAM[IX(sl, k)] = dl;
                if (ds < 0 && -tau * s / ds < 1.0) hist[clip_bin(-tau * s / ds)]++;
                if (dl < 0 && -tau * lam / dl < a_d) a_d = -tau * lam / dl;
                dphi_bar += -mu * ds / s;
                curv += (lam / s) * ds * ds;
            }
        {
            /* threshold bin: the bins jt.. (smallest ratios) hold at most m_rows / ORC_CLIP_DIV rows: those are clipped; the
               primal step length is the smallest ratio of the other rows */
            int jt = ORC_CLIP_BINS, cum = 0;
            for (int j = ORC_CLIP_BINS - 1; j >= 0; --j)
            {
                if (cum + hist[j] > m_rows / ORC_CLIP_DIV) break;
                cum += hist[j]; jt = j;
            }
            for (int k = 0; k < N; ++k)
                for (int sl = 0; sl < RS; ++sl)
                {
                    const double ds = ws->DS[IX(sl, k)];
                    if (!(ds < 0)) continue;
                    const double r = -tau * ws->S[IX(sl, k)] / ds;
                    if (r < a_p && clip_bin(r) < jt) a_p = r;
                }
        }
        /* ---- l1 merit function and its directional derivative ---- */
        /* grad J . dz : cost gradient = condensed gradient minus the row terms; recompute from the objective pieces:
           use GL - (dynamics + row multiplier terms) is messy, so accumulate directly: dJ = sum (g_k - rowpart).  Simpler and
           exact: dJ := directional derivative of the objective by its own gradient, evaluated below. */
        double dJ = 0.0;
        {
            if (has_mintime(c)) dJ += (double)(N - 1) * ddt;
            if (has_quadratic(c))
                for (int k = 0; k <= N - 2; ++k)
                {
                    double d[3] = {ws->X[IX(0, k)] - p->xf[0], ws->X[IX(1, k)] - p->xf[1], orc_normalize_theta(ws->X[IX(2, k)] - p->xf[2])};
                    double u[2] = {ws->U[IX(0, k)], ws->U[IX(1, k)]};
                    const int integ = c->quadratic_integral_form != 0;
                    const double fx = integ ? integral_state_weight(c, N, k) : 1.0;
                    const double wx = integ ? dt * fx : 1.0, wu = integ ? dt : 1.0;
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j) dJ += wx * (c->Q[i * 3 + j] + c->Q[j * 3 + i]) * d[j] * ws->STEP[IX(i, k)];
                    for (int i = 0; i < 2; ++i)
                        for (int j = 0; j < 2; ++j) dJ += wu * (c->R[i * 2 + j] + c->R[j * 2 + i]) * u[j] * ws->STEP[IX(3 + i, k)];
                    if (integ && c->variable_dt) dJ += (fx * quad3(c->Q, d) + quad2(c->R, u)) * ddt;
                }
            if (has_trapezoid(c))
            {
                int k = N - 1;
                const double fx = integral_state_weight(c, N, k);
                double d[3] = {ws->X[IX(0, k)] - p->xf[0], ws->X[IX(1, k)] - p->xf[1], orc_normalize_theta(ws->X[IX(2, k)] - p->xf[2])};
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) dJ += dt * fx * (c->Q[i * 3 + j] + c->Q[j * 3 + i]) * d[j] * ws->STEP[IX(i, k)];
                if (c->variable_dt) dJ += fx * quad3(c->Q, d) * ddt;
            }
            if (has_terminal_cost(c))
            {
                int k = N - 1;
                double d[3] = {ws->X[IX(0, k)] - p->xf[0], ws->X[IX(1, k)] - p->xf[1], orc_normalize_theta(ws->X[IX(2, k)] - p->xf[2])};
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) dJ += (c->Qf[i * 3 + j] + c->Qf[j * 3 + i]) * d[j] * ws->STEP[IX(i, k)];
            }
            if (has_viapoints(c))
                for (int j = 0; j < p->n_vp && j < 64; ++j)
                {
                    int k = ws->vp_stage[j];
                    if (k < 0) continue;
                    double w = c->vp_position_weight;
                    dJ += -2 * w * (p->vp[3 * j] - ws->X[IX(0, k)]) * ws->STEP[IX(0, k)];
                    dJ += -2 * w * (p->vp[3 * j + 1] - ws->X[IX(1, k)]) * ws->STEP[IX(1, k)];
                    if (c->vp_orientation_weight > 0) dJ += -c->vp_orientation_weight * ws->STEP[IX(2, k)];
                }
        }
        /* curvature dz' H dz (block tridiagonal + border) */
        {
            const double* KKT = ws->KKT;
            for (int k = 0; k < N; ++k)
            {
                int nv = (k <= N - 2) ? 5 : 3;
                for (int i = 0; i < nv; ++i)
                    for (int j = 0; j < nv; ++j)
                    {
                        int a = i < j ? i : j, b = i < j ? j : i;
                        curv += ws->STEP[IX(i, k)] * (KKT[(MPCB200_K_H + hidx(a, b)) * N + k] + (a == b ? ws->SCAL[MPCB200_SC_DELTA] : 0.0)) * ws->STEP[IX(j, k)];
                    }
                if (k >= 1 && k <= N - 2)
                    for (int i = 0; i < 2; ++i) curv += 2.0 * ws->STEP[IX(3 + i, k - 1)] * KKT[(MPCB200_K_C + i) * N + k] * ws->STEP[IX(3 + i, k)];
                if (c->variable_dt) /* border; the terminal record carries one only with the trapezoidal rule */
                    for (int i = 0; i < nv; ++i) curv += 2.0 * ddt * KKT[(MPCB200_K_HB + i) * N + k] * ws->STEP[IX(i, k)];
            }
            if (c->variable_dt) curv += ddt * ddt * (ws->SCAL[MPCB200_SC_HTT] + ws->SCAL[MPCB200_SC_DELTA]);
        }
        const double inf1 = info.inf1;
        double rho = 1.0; /* memoryless penalty parameter: recomputed every iteration */
        {
            double num = dJ + dphi_bar + 0.5 * (curv > 0 ? curv : 0.0);
            if (inf1 > 1e-14)
            {
                double rho_trial = num / ((1.0 - 0.1) * inf1);
                if (rho < rho_trial) rho = rho_trial + 1.0;
            }
        }
        ws->SCAL[MPCB200_SC_RHO] = rho;
        const double phi0 = info.obj - mu * info.barrier_log + rho * inf1;
        const double dphi = dJ + dphi_bar - rho * inf1;
        /* ---- backtracking line search ---- */
        double alpha = a_p;
        int accepted = 0;
        for (int bt = 0; bt < ORC_MAX_BACKTRACK; ++bt)
        {
            for (int k = 0; k < N; ++k)
            {
                for (int i = 0; i < 3; ++i) ws->XT[IX(i, k)] = ws->X[IX(i, k)] + alpha * ws->STEP[IX(i, k)];
                for (int i = 0; i < 2; ++i) ws->UT[IX(i, k)] = ws->U[IX(i, k)] + (k <= N - 2 ? alpha * ws->STEP[IX(3 + i, k)] : 0.0);
            }
            double dtt = dt + alpha * ddt;
            double obj, inft, blog;
            trial_eval(p, ws, ws->XT, ws->UT, dtt, alpha, &obj, &inft, &blog);
            double phi = obj - mu * blog + rho * inft;
            if (phi <= phi0 + ORC_ARMIJO * alpha * dphi || (bt > 0 && fabs(phi - phi0) <= 1e-13 * (1.0 + fabs(phi0))))
            {
                accepted = 1;
                break;
            }
            alpha *= 0.5;
            ++nbt;
        }
        if (!accepted)
        {
            /* tiny step: accept anyway (keeps the iteration going; counted) */
        }
        /* ---- update ---- */
        double a_dual = a_d > alpha ? alpha : a_d; /* the multipliers never move further than the primal step taken */
        for (int k = 0; k < N; ++k)
        {
            for (int i = 0; i < 3; ++i) ws->X[IX(i, k)] += alpha * ws->STEP[IX(i, k)];
            if (k <= N - 2)
            {
                for (int i = 0; i < 2; ++i) ws->U[IX(i, k)] += alpha * ws->STEP[IX(3 + i, k)];
                for (int i = 0; i < 3; ++i) ws->NU[IX(i, k)] += alpha * (ws->STEP[IX(5 + i, k)] - ws->NU[IX(i, k)]);
            }
            for (int sl = 0; sl < RS; ++sl)
            {
                if (!row_active(p, ws, k, sl)) continue;
                double s = ws->S[IX(sl, k)] + alpha * ws->DS[IX(sl, k)];
                if (s < ORC_CLIP_FLOOR * ws->S[IX(sl, k)]) s = ORC_CLIP_FLOOR * ws->S[IX(sl, k)];
                double lam = ws->LAM[IX(sl, k)] + a_dual * ws->DLAM[IX(sl, k)];
                double lo = mu / (ORC_KAPPA_SIGMA * s), hi = ORC_KAPPA_SIGMA * mu / s;
                if (lam < lo) lam = lo;
                if (lam > hi) lam = hi;
                ws->S[IX(sl, k)] = s;
                ws->LAM[IX(sl, k)] = lam;
            }
        }
        if (c->variable_dt) ws->SCAL[MPCB200_SC_DT] = dt + alpha * ddt;
        ws->SCAL[MPCB200_SC_ALPHA] = alpha;
        {
            /* jam detection: two consecutive steps shorter than 1e-8 -> give the instance up */
            const double tiny = alpha < ORC_TINY_STEP ? ws->SCAL[MPCB200_SC_TINY] + 1.0 : 0.0;
            ws->SCAL[MPCB200_SC_TINY] = tiny;
            if (tiny >= (double)ORC_TINY_STEP_COUNT) { status = MPCB200_STATUS_NUMERICAL_ERROR; ++iter; break; }
        }
#ifdef ORC_TRACE
            fprintf(stderr, "it %3d mu %.1e E0 %.2e Emu %.2e obj %.5f inf1 %.2e dinf %.2e delta %.1e a_p %.3f a_d %.3f alpha %.4f rho %.2e dphi %.2e acc %d dt %.4f\n",
                    iter, mu, ws->SCAL[MPCB200_SC_ERR0], ws->SCAL[MPCB200_SC_ERRMU], info.obj, inf1, info.dual_inf, delta, a_p, a_d, alpha, rho, dphi, accepted, dt);
#endif
        ++iter;
    }
    ws->SCAL[MPCB200_SC_STATUS] = (double)status;
    ws->SCAL[MPCB200_SC_NREG] = (double)nreg;
    if (res)
    {
        res->status = status; res->iters = iter; res->kkt_err = ws->SCAL[MPCB200_SC_ERR0];
        res->objective = ws->SCAL[MPCB200_SC_OBJ]; res->dt = ws->SCAL[MPCB200_SC_DT];
        res->n_regularised = nreg; res->n_backtracks = nbt;
    }
    return status;
}

/*
 * One Controller::step for one instance (R/src/controller.cpp:111-179 + corbo::PredictiveController::step):
 * cold: init from the initial plan; warm: shift.  outer_iterations x (association + solve); new_run only on the first.
 * Output time series as FullDiscretizationGridBaseSE2::getStateAndControlTimeSeries [full_discretization_grid_base_se2.cpp:579-615]:
 * u_seq has N samples, the last control duplicated.
 */
int orc_step(const orc_problem* p, orc_ws* ws, const double* x_init, int reinit, double* u_seq, double* x_seq,
             orc_result* res)
{
    const int N = ws->N;
    const int is_cold = (ws->cold || reinit);
    if (is_cold) orc_init_cold(p, x_init, ws);
    else if (p->cfg->warm_start && !p->cfg->variable_dt) orc_warm_shift(p, ws);
    else
    {
        for (int i = 0; i < 3; ++i) ws->X[IX(i, 0)] = p->x0[i];
        for (int i = 0; i < 3; ++i)
            if (p->cfg->xf_fixed[i]) ws->X[IX(i, N - 1)] = p->xf[i];
    }
    int status = 0;
    orc_result r, racc;
    memset(&racc, 0, sizeof(racc));
    int outer = p->cfg->outer_iterations > 0 ? p->cfg->outer_iterations : 1;
    for (int it = 0; it < outer; ++it)
    {
        orc_associate(p, ws);
        if (it == 0 && is_cold && !p->cfg->reference_initial_guess)   /* solver-side preprocessing; off: the reference's guess */
        {
            orc_project_init(p, ws);
            orc_init_controls(p, ws);
        }
        orc_init_duals(p, ws);
        status = orc_solve(p, ws, &r);
        racc.iters += r.iters; racc.n_regularised += r.n_regularised; racc.n_backtracks += r.n_backtracks;
    }
    r.iters = racc.iters; r.n_regularised = racc.n_regularised; r.n_backtracks = racc.n_backtracks;
    ws->cold = 0;
    for (int k = 0; k < N; ++k)
    {
        if (x_seq)
        {
            x_seq[3 * k + 0] = ws->X[IX(0, k)];
            x_seq[3 * k + 1] = ws->X[IX(1, k)];
            x_seq[3 * k + 2] = orc_normalize_theta(ws->X[IX(2, k)]);
        }
        if (u_seq)
        {
            int kk = k <= N - 2 ? k : N - 2;
            u_seq[2 * k + 0] = ws->U[IX(0, kk)];
            u_seq[2 * k + 1] = ws->U[IX(1, kk)];
        }
    }
    if (res) *res = r;
    return status;
}

/* ---- batch driver (CPU baseline): one instance per task, static partition over threads ------------------ */
typedef struct {
    const mpcb200_config* cfg; int B; const double *x0, *xf, *u_prev; double u_prev_dt;
    const mpcb200_obstacles* obst; const mpcb200_viapoints* vp; const double* x_init;
    double *u_seq, *x_seq, *dt_out; int* status; double* kkt_err; int* iters;
    int tid, nthreads;
} batch_arg;

static void* batch_worker(void* vp_)
{
    batch_arg* a = (batch_arg*)vp_;
    const mpcb200_config* cfg = a->cfg;
    const int N = cfg->n, K = cfg->k_max_obstacles_per_stage;
    orc_ws* ws = orc_ws_alloc(N, K);
    for (int b = a->tid; b < a->B; b += a->nthreads)
    {
        orc_problem p;
        memset(&p, 0, sizeof(p));
        p.cfg = cfg;
        for (int i = 0; i < 3; ++i) { p.x0[i] = a->x0[3 * b + i]; p.xf[i] = a->xf[3 * b + i]; }
        for (int i = 0; i < 2; ++i) p.u_prev[i] = a->u_prev ? a->u_prev[2 * b + i] : 0.0;
        p.u_prev_dt = a->u_prev_dt;
        if (a->obst && a->obst->count)
        {
            int M = a->obst->max_per_instance;
            p.n_obst = a->obst->count[b];
            p.obst_type = a->obst->type + (size_t)b * M;
            p.obst_params = a->obst->params + (size_t)b * M * MPCB200_OBST_STRIDE;
        }
        if (a->vp && a->vp->count)
        {
            p.n_vp = a->vp->count[b];
            p.vp = a->vp->poses + (size_t)b * a->vp->max_per_instance * 3;
        }
        orc_result r;
        ws->cold = 1;
        orc_step(&p, ws, a->x_init ? a->x_init + (size_t)b * N * 3 : NULL, 1, a->u_seq ? a->u_seq + (size_t)b * N * 2 : NULL,
                 a->x_seq ? a->x_seq + (size_t)b * N * 3 : NULL, &r);
        if (a->dt_out) a->dt_out[b] = r.dt;
        if (a->status) a->status[b] = r.status;
        if (a->kkt_err) a->kkt_err[b] = r.kkt_err;
        if (a->iters) a->iters[b] = r.iters;
    }
    orc_ws_free(ws);
    return NULL;
}

int orc_step_batch(const mpcb200_config* cfg, int B, const double* x0, const double* xf, const double* u_prev,
                   double u_prev_dt, const mpcb200_obstacles* obst, const mpcb200_viapoints* vp, const double* x_init,
                   double* u_seq, double* x_seq, double* dt_out, int* status, double* kkt_err, int* iters,
                   int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    batch_arg args[256];
    for (int t = 0; t < n_threads; ++t)
    {
        args[t] = (batch_arg){cfg, B, x0, xf, u_prev, u_prev_dt, obst, vp, x_init, u_seq, x_seq, dt_out, status, kkt_err, iters, t, n_threads};
        if (n_threads == 1) batch_worker(&args[t]);
        else pthread_create(&th[t], NULL, batch_worker, &args[t]);
    }
    if (n_threads > 1)
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    return 0;
}
