/*
 * mpc_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the optimal control problem that mpc_local_planner poses behind Controller::step()
 * (reference files cited per function in mpc_oracle.c) plus a serial primal-dual interior-point solver that
 * stands in for the reference's control_box_rst + Ipopt path.
 *
 * PARITY UNPINNED: the reference ships no tests, no golden vectors and no recorded outputs,
 * and its numerical core (control_box_rst, Ipopt/MUMPS, teb_local_planner geometry) is third-party code that is
 * absent from /root/reference and unpinned (mpc_local_planner/package.xml:29,45).  The reference cannot be compiled
 * here (needs ROS1, Eigen, corbo, Ipopt).  This oracle is therefore cross-validated against two independent scipy
 * solvers (SLSQP, trust-constr) on the same restated functions (tests/test_oracle_scipy.py, tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (libmpcb200.so) never links or calls it.
 */
#ifndef MPC_ORACLE_H_
#define MPC_ORACLE_H_

#include "../include/mpcb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One OCP instance (inputs of Controller::step for one robot). */
typedef struct orc_problem {
    const mpcb200_config* cfg;
    double x0[3], xf[3], u_prev[2], u_prev_dt;
    int n_obst;
    const int* obst_type;      /* [n_obst] */
    const double* obst_params; /* [n_obst][MPCB200_OBST_STRIDE] */
    int n_vp;
    const double* vp;          /* [n_vp][3] */
} orc_problem;

/* Per-instance workspace; array fields use the SAME [component][k] layout as the device workspace (mpcb200.h). */
typedef struct orc_ws {
    int N, K, RS;
    double* X;      /* 3 x N */
    double* U;      /* 2 x N */
    double* NU;     /* 3 x N */
    double* S;      /* RS x N */
    double* LAM;    /* RS x N */
    double* KKT;    /* 42 x N */
    double* STEP;   /* 8 x N */
    double* OBSIDX; /* K x N (as double, -1 = empty) */
    double SCAL[MPCB200_SCAL_WORDS];
    int vp_stage[64]; /* stage index per via-point, -1 = skipped */
    /* scratch (not mirrored on the device) */
    double* GL;     /* 5 x N  gradient of the Lagrangian per stage (for the KKT error) */
    double* G;      /* RS x N inequality row values g */
    double* DS;     /* RS x N slack step */
    double* DLAM;   /* RS x N multiplier step */
    double* XT;     /* trial point scratch 3 x N, 2 x N */
    double* UT;
    double* P;      /* N x 25 Riccati value matrices */
    double* PI;     /* N x 25 Riccati [p | Pi] */
    double* KG;     /* N x 10 feedback gains */
    double* KT;     /* N x 10 theta-hat feedforward gains */
    double gl_dt;   /* dL/d(dt) */
    int cold;       /* 1 until the first successful solve */
} orc_ws;

typedef struct orc_result {
    int status;
    int iters;
    double kkt_err;
    double objective;
    double dt;
    int n_regularised;  /* inertia-correction refactorisations */
    int n_backtracks;
} orc_result;

/* ---- reference restatement: elementary functions ------------------------------------------------------- */
double orc_normalize_theta(double theta);
double orc_interpolate_angle(double a1, double a2, double factor);
void orc_dynamics(const mpcb200_config* cfg, const double* x, const double* u, double* f);
/* derivatives wrt (theta, u0, u1): J[3][3]; Hc[6] = sum_j nu_j * Hess f_j packed (tt,t0,t1,00,01,11) */
void orc_dynamics_derivs(const mpcb200_config* cfg, const double* x, const double* u, const double* nu, double* f,
                         double J[9], double Hc[6]);
/* ForwardDiffCollocationSE2::computeEqualityConstraint exactly as coded in the reference (divides by dt). */
void orc_defect_reference(const mpcb200_config* cfg, const double* x1, const double* u1, const double* x2, double dt,
                          double* e);
/* The solver's multiplied form: e = x1 + dt f(x1,u1) - x2 (theta row wrapped) == dt * reference defect. */
void orc_defect(const mpcb200_config* cfg, const double* x1, const double* u1, const double* x2, double dt, double* e);
/* RobotFootprintModel::calculateDistance(pose, obstacle) [teb, SURVEY App. B.3]; optional gradient/Hessian wrt (x,y,theta). */
double orc_footprint_distance(const mpcb200_config* cfg, const double* pose, int obst_type, const double* obst_params,
                              double* grad3, double* hess6);
/* objective value of an iterate (sum of all cost terms incl. constants at k = 0) */
double orc_objective(const orc_problem* p, const orc_ws* ws, const double* X, const double* U, double dt);

/* ---- workspace ------------------------------------------------------------------------------------------ */
orc_ws* orc_ws_alloc(int N, int K);
void orc_ws_free(orc_ws* ws);

/* ---- solver phases (mirrors the device phases) ---------------------------------------------------------- */
/* A.6 cold initial guess (x_init optional [N][3]) + dt = dt_ref + slack/multiplier init is done in orc_init_duals */
void orc_init_cold(const orc_problem* p, const double* x_init, orc_ws* ws);
/* A.7 warm-start shift (FullDiscretizationGridBaseSE2::warmStartShifting) */
void orc_warm_shift(const orc_problem* p, orc_ws* ws);
/* updateObstacleContainerWithCostmap for one robot: mpc_local_planner_ros.cpp:474-499 */
int orc_costmap_obstacles(int size_x, int size_y, double resolution, const double* origin, const unsigned char* cost,
                          const double* robot_pose, double behind_dist, int max_out, double* xy);
/* Controller::isPoseTrajectoryFeasible for one robot: controller.cpp:859-917 (+ [EXT] CostmapModel::footprintCost) */
int orc_pose_trajectory_feasible(int size_x, int size_y, double resolution, const double* origin, const unsigned char* cost,
                                 const double* x_seq, int n, const double* footprint, int n_fp, double inscribed_radius,
                                 double min_resolution_angular, int look_ahead_idx);
/* resampleTrajectory(n_new): full_discretization_grid_base_se2.cpp:440-524 */
double orc_resample_trajectory(int n, const double* X, const double* U, double dt, int n_new, double* Xn, double* Un);
/* a12/a11: obstacle + via-point association from the current trajectory */
void orc_associate(const orc_problem* p, orc_ws* ws);
/* solver-side initial-guess repair (cold start only): push poses out of violated obstacle rows; seed the controls by
   inverting the dynamics along the state guess, clipped inside bounds and rate rows */
void orc_project_init(const orc_problem* p, orc_ws* ws);
void orc_init_controls(const orc_problem* p, orc_ws* ws);
void orc_init_duals(const orc_problem* p, orc_ws* ws);
/* stage functions + derivatives -> KKT records, errors (SCAL[ERR0], SCAL[ERRMU]) */
void orc_eval(const orc_problem* p, orc_ws* ws);
/* Riccati factorisation + solve; returns 0 ok, 1 wrong inertia */
int orc_kkt_solve(const orc_problem* p, orc_ws* ws, double delta);
/* full interior-point solve from the current workspace state */
int orc_solve(const orc_problem* p, orc_ws* ws, orc_result* res);

/* Convenience: whole Controller::step for one instance (cold or warm according to ws->cold / reinit). */
int orc_step(const orc_problem* p, orc_ws* ws, const double* x_init, int reinit, double* u_seq, double* x_seq,
             orc_result* res);

/* Batch driver used as the CPU baseline: same argument meaning as mpcb200_step_batch, always cold. */
int orc_step_batch(const mpcb200_config* cfg, int B, const double* x0, const double* xf, const double* u_prev,
                   double u_prev_dt, const mpcb200_obstacles* obst, const mpcb200_viapoints* vp, const double* x_init,
                   double* u_seq, double* x_seq, double* dt_out, int* status, double* kkt_err, int* iters,
                   int n_threads);

#ifdef __cplusplus
}
#endif
#endif
