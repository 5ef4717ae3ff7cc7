/* oracle/clrrt_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the CL-RRT tree-expansion hot path of vdBerg93/cl-rrt
 * (MotionPlanner::planMotion -> expandTree -> Simulation::propagate, with nearest-node
 * selection and the OBB/SAT collision check).  Every function cites the reference
 * file:line it follows (paths relative to /root/reference/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (cl-rrt_b200/) never links or calls it.
 *
 * Parity pin: tests/test_oracle_vs_reference.py and tests/golden/ check this file
 * bit-for-bit against the reference's own sources compiled by oracle/build_ref.sh
 * (known answers of rrt/src/testers.cpp:129-160 included).  Both run on the same glibc
 * libm / rand() and neither is FMA-contracted, so agreement is exact, not approximate.
 *
 * Semantics are those of the "defined" reference variant: identical to
 * the unmodified sources wherever those are defined, and where the reference reads out
 * of bounds it uses
 *   - ref.v[min(IDwp+LAlong, N-1)]      (rrt/src/controller.cpp:39, rrt/src/simulation.cpp:66)
 *   - the (N-3,N-2,N-1) lateral-error window when IDwp==N-1 (rrt/src/controller.cpp:76)
 *   - Euler integration over the 7 ODE states only (rrt/src/simulation.cpp:28)
 *   - normsY[3] = 0 for the never-written fourth SAT axis (rrt/src/old_collisioncheck.cpp:74-75)
 *
 * Arithmetic types follow the reference's mixture exactly: rollout in
 * double; Dubins metric and SAT geometry in float; node costs stored as float.
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fPIC -shared (no -march=native, no fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_NODE_STRIDE 20
#define ORC_OUT_STRIDE 24
#define ORC_SORT_LIMIT 10 /* rrt/src/rrtplanner.cpp:13 */
#define ORC_LALONG 2      /* rrt/src/controller.cpp:35 */

typedef struct { /* rrt/include/rrt/vehicle.h:5-21 */
  double dmax, ddmax, Td, Ta, amin, amax, L, w, Lrear, Lfront, b, Vch, rho, Kus;
} orc_vehicle;

typedef struct {
  double cx, cy, theta, size_x, size_y, vx, vy; /* car_msgs/msg/Obstacle2D.msg */
} orc_obstacle;

typedef struct {
  double state[10];
  double ref_fx, ref_fy, ref_bx, ref_by, ref_vback;
  float costE, costS; /* rrt/include/rrt/rrtplanner.h:40-41 */
  int parent, goal, nref;
} orc_node;

/* file-scope state mirrors the reference's globals (rrt/src/rrt_node.cpp:2-24) */
static double sim_dt, ctrl_tla, ctrl_dla, ctrl_mindla, ctrl_dlavmin, ctrl_Kp, ctrl_Ki;
static double ref_res, ref_int, ref_mindist, vmax, ay_road_max;
static int obs_use_pred = 1;
static int fail_iterlimit, fail_collision, fail_acclimit, sim_count;
static double Wcost[5];
/* curved-road mode: MyRRT::bend, laneShifts[0], Cxy (rrt/include/rrt/rrtplanner.h:55-58) */
static int road_bend = 0;
static double road_S = 0, road_Cxy[3] = {0, 0, 0};
/* getDistToLane, rrt/src/simulation.cpp:49-53 */
static double getDistToLane(double x, double y, double S, const double* Cxy) {
  double Lx = (x - S * Cxy[1] + y * Cxy[1] - Cxy[1] * Cxy[2]) / (Cxy[1] * Cxy[1] + 1);
  double Ly = S + Cxy[2] + (Cxy[1] * (x - S * Cxy[1] + y * Cxy[1] - Cxy[1] * Cxy[2])) / (Cxy[1] * Cxy[1] + 1);
  return sqrt((Lx - x) * (Lx - x) + (Ly - y) * (Ly - y));
}
static double goalPose[4];
static orc_vehicle veh;
static orc_obstacle* det = NULL;
static int n_det = 0; /* 0 => shipped stub, rrt/src/collisioncheck.cpp:6-8 */
static orc_node* tree = NULL;
static int n_tree = 0, cap_tree = 0;
static long sat_calls = 0, sat_axes = 0, wp_scanned = 0;

static double std_max(double a, double b) { return (a < b) ? b : a; } /* std::max(a,b) */
static double std_min(double a, double b) { return (b < a) ? b : a; } /* std::min(a,b) */
static double sq(double x) { return x * x; }                         /* pow(x,2), folded by GCC */

/* rrt/include/rrt/functions.h:49-56 */
static double angleDiff(double a, double b) {
  double dif = fmod(b - a + M_PI, 2 * M_PI);
  if (dif < 0) dif += 2 * M_PI;
  return dif - M_PI;
}
/* rrt/include/rrt/functions.h:42-47 */
static double wrapToPi(double x) {
  x = fmod(x + M_PI, 2 * M_PI);
  if (x < 0) x += 2 * M_PI;
  return x - M_PI;
}
/* rrt/include/rrt/functions.h:60-62, rrt/src/simulation.cpp:7-9 */
static double saturate(double lo, double hi, double val) { return std_max(std_min(val, hi), lo); }

/* rrt/include/rrt/functions.h:11-21: values by accumulation val += h */
static void linspace(double a, double b, size_t N, double* out) {
  double h = (b - a) / (double)(N - 1);
  double val = a;
  for (size_t i = 0; i < N; i++, val += h) out[i] = val;
}

/* rrt/src/controller.cpp:13-16 */
static void updateLookahead(double v) {
  double dla_c = ctrl_mindla - ctrl_tla * ctrl_dlavmin;
  ctrl_dla = std_max(ctrl_mindla, dla_c + ctrl_tla * fabs(v));
}
/* rrt/src/controller.cpp:18-21 */
static void updateReferenceResolution(double v) { ref_res = std_max(fabs(v) * ref_int, ref_mindist); }

typedef struct {
  int N;
  double *x, *y, *v;
  int dir;
} orc_ref;
static void ref_free(orc_ref* r) { free(r->x); free(r->y); free(r->v); r->x = r->y = r->v = NULL; }

/* rrt/src/reference.cpp:9-22 */
static orc_ref getReference(double sx, double sy, const orc_node* node, int dir) {
  orc_ref ref;
  double L = sqrt(sq(sx - node->ref_bx) + sq(sy - node->ref_by));
  int N = (int)(round(L / ref_res) + 1);
  ref.N = N;
  ref.x = malloc(sizeof(double) * (size_t)N);
  ref.y = malloc(sizeof(double) * (size_t)N);
  ref.v = NULL;
  linspace(node->ref_bx, sx, (size_t)N, ref.x);
  linspace(node->ref_by, sy, (size_t)N, ref.y);
  ref.dir = dir;
  return ref;
}

/* rrt/src/reference.cpp:25-70 */
static orc_ref getGoalReference(const orc_node* node, const double* goal) {
  double dla_c = ctrl_mindla - ctrl_tla * ctrl_dlavmin;
  double dla_end = std_max(ctrl_mindla, dla_c + ctrl_tla * fabs(goal[3]));
  double Dextend = dla_end;
  double Dalign = 1;
  double P1x = goal[0] + Dalign * cos(goal[2]), P1y = goal[1] + Dalign * sin(goal[2]);
  double P2x = goal[0] - Dalign * cos(goal[2]), P2y = goal[1] - Dalign * sin(goal[2]);
  double Pcx, Pcy, Pfx, Pfy;
  if (sqrt(sq(P1x - node->ref_bx) + sq(P1y - node->ref_by)) <
      sqrt(sq(P2x - node->ref_bx) + sq(P2y - node->ref_by))) {
    Pcx = P1x; Pcy = P1y; Pfx = P1x; Pfy = P1y;
  } else {
    Pcx = P2x; Pcy = P2y; Pfx = P2x; Pfy = P2y;
  }
  Pfx += (Dextend + Dalign) * cos(goal[2]);
  Pfy += (Dextend + Dalign) * sin(goal[2]);
  double N1 = round(sqrt(sq(Pcx - node->ref_bx) + sq(Pcy - node->ref_by)) / ref_res) + 1;
  double N2 = round(sqrt(sq(Pfx - Pcx) + sq(Pfy - Pcy)) / ref_res) + 1;
  size_t n1 = (size_t)N1, n2 = (size_t)N2;
  orc_ref ref;
  ref.N = (int)(n1 + n2);
  ref.x = malloc(sizeof(double) * (n1 + n2));
  ref.y = malloc(sizeof(double) * (n1 + n2));
  ref.v = NULL;
  linspace(node->ref_bx, Pcx, n1, ref.x);
  linspace(Pcx, Pfx, n2, ref.x + n1);
  linspace(node->ref_by, Pcy, n1, ref.y);
  linspace(Pcy, Pfy, n2, ref.y + n1);
  ref.dir = 1; /* upstream leaves ref.dir unset here (reference.cpp:34); every caller drives forward */
  return ref;
}

/* rrt/src/reference.cpp:73-170 */
static void generateVelocityProfile(orc_ref* ref, double v0, double vmax_, const double* goal, int GB) {
  double vend = goal[3];
  double a_acc = 1, a_dec = -1, tmin = 1;
  double Lp, res;
  const int N = ref->N;
  if (GB) {
    double Dgoal = sqrt(sq(goal[0] - ref->x[0]) + sq(goal[1] - ref->y[0]));
    Lp = Dgoal + ctrl_mindla;
    res = Lp / (double)((size_t)N - 1);
  } else {
    double Dgoal = sqrt(sq(goal[0] - ref->x[N - 1]) + sq(goal[1] - ref->y[N - 1]));
    double Lref = sqrt(sq(ref->x[0] - ref->x[N - 1]) + sq(ref->y[0] - ref->y[N - 1]));
    res = Lref / (double)((size_t)N - 1);
    Lp = Lref + Dgoal + ctrl_mindla;
  }
  double Daccel = (sq(vmax_) - sq(v0)) / (2 * a_acc);
  double Dcoast = vmax_ * tmin;
  double Dbrake = (sq(vend) - sq(vmax_)) / (2 * a_dec);
  int D_vmax_bool = (Daccel + Dcoast + Dbrake) < Lp;
  double Vcoast;
  if (vend > (v0 + 0.1)) {
    Vcoast = vend;
  } else if (D_vmax_bool) {
    Vcoast = vmax_;
  } else {
    double D = Lp;
    double v1 = (sqrt(sq(a_acc) * sq(a_dec) * sq(tmin) - 2 * D * sq(a_acc) * a_dec + sq(a_acc) * sq(vend) +
                      2 * D * a_acc * sq(a_dec) - a_acc * a_dec * sq(v0) - a_acc * a_dec * sq(vend) +
                      sq(a_dec) * sq(v0)) +
                 a_acc * a_dec * tmin) /
                (a_acc - a_dec);
    Vcoast = v1;
  }
  Daccel = (sq(Vcoast) - sq(v0)) / (2 * a_acc);
  if (Daccel < 0) { Daccel = 0; Vcoast = v0; }
  Dbrake = std_max(0.0, (sq(vend) - sq(Vcoast)) / (2 * a_dec));
  Dcoast = std_max(0.0, Lp - Daccel - Dbrake);
  double tbrake = (vend - Vcoast) / a_dec;
  ref->v = malloc(sizeof(double) * (size_t)N);
  for (int i = 0; i != N; i++) {
    double D = i * res;
    if (D < Daccel) {
      double t1 = -(v0 - sqrt(sq(v0) + 2 * a_acc * D)) / a_acc;
      double t2 = -(v0 + sqrt(sq(v0) + 2 * a_acc * D)) / a_acc;
      double t = (t1 >= 0) * t1 + (t2 >= 0) * t2;
      ref->v[i] = v0 + a_acc * t;
    } else if (D <= (Daccel + Dcoast)) {
      ref->v[i] = Vcoast;
    } else {
      double t1 = -(Vcoast + sqrt(sq(Vcoast) + 2 * D * a_dec - 2 * Daccel * a_dec - 2 * Dcoast * a_dec)) / a_dec;
      double t2 = -(Vcoast - sqrt(sq(Vcoast) + 2 * D * a_dec - 2 * Daccel * a_dec - 2 * Dcoast * a_dec)) / a_dec;
      double dt = (t1 != tbrake) * (t1 >= 0) * (t1 <= tbrake) * t1 + (t2 >= 0) * (t2 <= tbrake) * t2;
      ref->v[i] = std_max(0.0, Vcoast + a_dec * dt);
    }
  }
}

/* ---------------- controller: rrt/src/controller.cpp ---------------- */
typedef struct {
  int IDwp;
  double Px, Py, ym, iE;
  int endreached;
} orc_ctrl;

/* rrt/src/controller.cpp:96-113 (the early exit is commented out upstream: full scan) */
static int findClosestPoint(const orc_ref* ref, double px, double py, int ID) {
  double dmin = INFINITY, di;
  int idmin = 0;
  for (int i = ID; i < ref->N; i++) {
    di = (ref->x[i] - px) * (ref->x[i] - px) + (ref->y[i] - py) * (ref->y[i] - py);
    if (di < dmin) { dmin = di; idmin = i; }
    wp_scanned++;
  }
  return idmin;
}
/* rrt/src/controller.cpp:53-68 */
static void updateWaypoint(orc_ctrl* c, const orc_ref* ref, const double* x) {
  updateLookahead(x[4]);
  c->Px = x[0] + ctrl_dla * ref->dir * cos(x[2]);
  c->Py = x[1] + ctrl_dla * ref->dir * sin(x[2]);
  c->IDwp = findClosestPoint(ref, c->Px, c->Py, c->IDwp);
  if ((size_t)c->IDwp >= (size_t)ref->N - 1 - ORC_LALONG) c->endreached = 1;
  if ((ref->x[c->IDwp] == ref->x[ref->N - 1]) && (ref->y[c->IDwp] == ref->y[ref->N - 1])) c->endreached = 1;
}
/* rrt/src/controller.cpp:70-93, :115-148 */
static double getLateralError(const orc_ref* ref, const double* x, int IDwp, double Px, double Py) {
  int IDmin, IDmax;
  if (IDwp == 0) { IDmin = IDwp; IDmax = IDwp + 2; }
  else if (IDwp == ref->N - 1) { IDmin = IDwp - 2; IDmax = IDwp; } /* defined variant of :76 */
  else { IDmin = IDwp - 1; IDmax = IDwp + 1; }
  double xval[3] = {ref->x[IDmin], ref->x[IDmin + 1], ref->x[IDmax]};
  double yval[3] = {ref->y[IDmin], ref->y[IDmin + 1], ref->y[IDmax]};
  double X[3] = {Px, Py, x[2]};
  double Tx[3], Ty[3];
  for (int i = 0; i <= 2; i++) { /* transformToVehicle :115-132 */
    Tx[i] = xval[i] * cos(X[2]) - X[0] * cos(X[2]) - yval[i] * sin(X[2]) + X[1] * sin(X[2]);
    Ty[i] = yval[i] * cos(X[2]) - X[1] * cos(X[2]) + xval[i] * sin(X[2]) - X[0] * sin(X[2]);
  }
  double y = 0, L; /* interpolate :134-148 */
  for (int i = 0; i <= 2; i++) {
    L = 1;
    for (int j = 0; j <= 2; j++)
      if (i != j) L = L * (Tx[j]) / (Tx[i] - Tx[j]);
    y = y + Ty[i] * L;
  }
  return y;
}
static int clampi(int i, int hi) { return i < hi ? i : hi; }

/* ---------------- collision: rrt/src/old_collisioncheck.cpp, rrt/include/rrt/collision.h ------------ */
typedef struct {
  double px, py;
  float w, h, o;
  float normsX[4], normsY[4], vx[4], vy[4], maxMin[2];
} orc_obb;
/* OBB ctor collision.h:24-27, setVertices :56-65, setNorms :67-76 */
static void obb_init(orc_obb* b, double px, double py, float w, float h, float o) {
  b->px = px; b->py = py; b->w = w; b->h = h; b->o = o;
  b->vx[0] = px + cosf(o) * (h / 2) - sinf(o) * (w / 2);
  b->vy[0] = py + sinf(o) * (h / 2) + cosf(o) * (w / 2);
  b->vx[1] = px + cosf(o) * (h / 2) - sinf(o) * (-w / 2);
  b->vy[1] = py + sinf(o) * (h / 2) + cosf(o) * (-w / 2);
  b->vx[2] = px + cosf(o) * (-h / 2) - sinf(o) * (-w / 2);
  b->vy[2] = py + sinf(o) * (-h / 2) + cosf(o) * (-w / 2);
  b->vx[3] = px + cosf(o) * (-h / 2) - sinf(o) * (w / 2);
  b->vy[3] = py + sinf(o) * (-h / 2) + cosf(o) * (w / 2);
  for (int i = 0; i < 3; i++) {
    b->normsX[i] = b->vy[i + 1] - b->vy[i];
    b->normsY[i] = -(b->vx[i + 1] - b->vx[i]);
  }
  b->normsX[3] = -(b->vx[0] - b->vx[3]); /* :74 is overwritten by :75 */
  b->normsY[3] = 0;                      /* never written upstream; defined variant */
}
/* findMaxMin :78-95 */
static void obb_maxmin(orc_obb* b, float x, float y) {
  b->maxMin[0] = b->vx[0] * x + b->vy[0] * y;
  b->maxMin[1] = b->maxMin[0];
  for (int i = 1; i <= 3; i++) {
    float proj = b->vx[i] * x + b->vy[i] * y;
    if (proj > b->maxMin[0]) b->maxMin[0] = proj;
    else if (proj < b->maxMin[1]) b->maxMin[1] = proj;
  }
}
/* getOBBdist :98-148 */
static double getOBBdist(orc_obb a, orc_obb b) {
  sat_calls++;
  for (int pass = 0; pass < 2; pass++) {
    const orc_obb* ax = pass == 0 ? &a : &b;
    for (int i = 0; i <= 3; i++) {
      float nx = ax->normsX[i], ny = ax->normsY[i];
      sat_axes++;
      obb_maxmin(&a, nx, ny);
      float aP0 = a.maxMin[0], aP1 = a.maxMin[1];
      obb_maxmin(&b, nx, ny);
      float bP0 = b.maxMin[0], bP1 = b.maxMin[1];
      float D1 = bP1 - aP0;
      float D2 = aP1 - bP0;
      if (D1 > 0) return D1;
      else if (D2 > 0) return D2;
    }
  }
  return 0;
}
/* checkObsDistance(states, det, carState) :24-51 with getOBBvector :6-22 */
static double obsDistance3(const double* states) {
  double t = obs_use_pred ? states[6] : 0;
  orc_obb vOBB;
  obb_init(&vOBB, states[0] + 1.424 * cos(states[2]), states[1] + 1.424 * sin(states[2]), 2, 4.848, states[2]);
  double dist2closest = 10000;
  for (int j = 0; j != n_det; j++) {
    orc_obb obs;
    obb_init(&obs, det[j].cx + det[j].vx * t, det[j].cy + det[j].vy * t, det[j].size_x / 2, det[j].size_y / 2,
             det[j].theta);
    double D = getOBBdist(vOBB, obs);
    if (D == 0) return 0;
    else if (D < dist2closest) dist2closest = D;
  }
  return dist2closest;
}
/* hook at rrt/src/simulation.cpp:83 */
static double checkObsDistance(const double* x) {
  if (n_det == 0) return 100; /* rrt/src/collisioncheck.cpp:6-8 */
  return obsDistance3(x);
}

/* ---------------- simulation: rrt/src/simulation.cpp ---------------- */
typedef struct {
  double xf[10], costE, costS;
  int endReached, goalReached, n_steps, fail, N, tainted;
  double ref_bx, ref_by, ref_vback, trace, idwp0;
} orc_sim;

static int gen_profile_flag = 1; /* the constructor's genProfile argument; every call site upstream passes true */
/* Simulation ctor :36-47 + propagate :55-143.  traj (optional) receives stateArray, (n_steps+1) x 10. */
static void simulate(const double* state0, orc_ref* ref, int GoalBiased, double Vstart, orc_sim* out, double* traj,
                     int traj_cap) {
  double x[10];
  memcpy(x, state0, sizeof x);
  out->costE = 0; out->costS = 0; out->goalReached = 0; out->endReached = 0;
  out->fail = 0; out->tainted = 0; out->trace = 0;
  orc_ctrl c; /* Controller ctor, controller.cpp:23-28 */
  updateLookahead(x[4]);
  c.IDwp = 0; c.endreached = 0; c.iE = 0;
  updateWaypoint(&c, ref, x);
  x[7] = c.IDwp; /* :41 */
  out->idwp0 = c.IDwp;
  if (gen_profile_flag) generateVelocityProfile(ref, Vstart, vmax, goalPose, GoalBiased); /* :42-45, genProfile */
  const int N = ref->N;
  out->N = N; out->ref_bx = ref->x[N - 1]; out->ref_by = ref->y[N - 1]; out->ref_vback = ref->v[N - 1];
  if (traj && traj_cap > 0) memcpy(traj, x, sizeof x);
  int i;
  for (i = 0; i < (20 / sim_dt); i++) { /* :58 */
    sim_count++;
    /* control.getControls :61 -> controller.cpp:30-34 */
    updateWaypoint(&c, ref, x);
    c.ym = getLateralError(ref, x, c.IDwp, c.Px, c.Py); /* getSteerCommand :47-51 */
    double cmdDelta = 2 * ((veh.L + veh.Kus * x[4] * x[4]) / sq(ctrl_dla)) * c.ym;
    double dc = saturate(-veh.dmax, veh.dmax, cmdDelta);
    double vref = ref->v[clampi(c.IDwp + ORC_LALONG, N - 1)]; /* getAccelerationCommand :37-45 */
    double E = vref - x[4];
    c.iE = c.iE + E * sim_dt;
    double ac = saturate(veh.amin, veh.amax, ctrl_Kp * E + ctrl_Ki * c.iE);
    /* VehicleODE :11-25 */
    double dx[7];
    double Gss = 1 / (1 + sq(x[4] / veh.Vch));
    dx[0] = x[4] * cos(x[2]);
    dx[1] = x[4] * sin(x[2]);
    dx[2] = (x[4] / veh.L) * tan(x[3]) * Gss;
    dx[3] = (1 / veh.Td) * (dc - x[3]);
    dx[4] = x[5];
    dx[5] = (1 / veh.Ta) * (ac - x[5]);
    dx[6] = 1;
    dx[4] = saturate(veh.amin, veh.amax, dx[4]);
    dx[3] = saturate(-veh.ddmax, veh.ddmax, dx[3]);
    /* IntegrateEuler :27-34 */
    for (int k = 0; k < 7; k++) x[k] = x[k] + dx[k] * sim_dt;
    x[3] = saturate(-veh.dmax, veh.dmax, x[3]);
    x[7] = c.IDwp;                                             /* :64 */
    x[8] = ref->v[clampi(c.IDwp + ORC_LALONG, N - 1)];          /* :66 */
    x[9] = dc;                                                 /* :67 */
    if (c.IDwp >= N - 2) out->tainted = 1;
    out->trace += (double)(i + 1) * (double)c.IDwp;
    if (traj && (i + 1) < traj_cap) memcpy(traj + 10 * (size_t)(i + 1), x, sizeof x);
    double Dobs = checkObsDistance(x); /* :83 */
    if (Dobs == 0) { out->endReached = 0; fail_collision++; out->fail = 1; i++; goto done; }
    out->costE += x[4] * sim_dt; /* :89-91 */
    double kappa = tan(x[3]) / veh.L;
    out->costS += Wcost[0] * x[4] * sim_dt + Wcost[1] * fabs(kappa) + Wcost[2] * exp(-Wcost[3] * Dobs);
    if (road_bend) { /* rrt/src/simulation.cpp:92-95 */
      double Dgoallane = getDistToLane(x[0], x[1], road_S, road_Cxy);
      out->costS += Wcost[4] * Dgoallane;
    }
    double ay = fabs(x[4] * dx[2]); /* :98-104 */
    if (ay + ay_road_max > 3) { out->endReached = 0; fail_acclimit++; out->fail = 2; i++; goto done; }
    double dist_to_goal = sqrt(sq(x[0] - goalPose[0]) + sq(x[1] - goalPose[1])); /* :110-111 */
    double goal_heading_error = fabs(angleDiff(x[2], goalPose[2]));
    double Verror = (x[4] - ref->v[N - 1]); /* :114-122; abs(Verror<0.1) is abs(bool): one-sided */
    if (c.endreached && (Verror < 0.1)) { out->endReached = 1; i++; goto done; }
    if ((dist_to_goal <= 1) && (goal_heading_error < 0.05)) { out->goalReached = 1; i++; goto done; } /* :125-133 */
  }
  fail_iterlimit++; /* :142 */
  out->fail = 3;
done:
  out->n_steps = i;
  memcpy(out->xf, x, sizeof x);
}

/* ---------------- tree / nearest: rrt/src/rrtplanner.cpp ---------------- */
/* dubinsDistance :371-406 — float throughout; double enters via M_PI terms and S - N.state */
static float dubinsDistance(double Sx, double Sy, const double* Nstate, int dir) {
  float rho = 4.77;
  float qw_x = Sx - Nstate[0];
  float qw_y = Sy - Nstate[1];
  float ang = -Nstate[2] - M_PI * (dir != 1);
  float tmp = cosf(ang) * qw_x - sinf(ang) * qw_y;
  qw_y = fabsf(sinf(ang) * qw_x + cosf(ang) * qw_y);
  qw_x = tmp;
  float dc = sqrtf(qw_x * qw_x + (qw_y - rho) * (qw_y - rho));
  float thetac = atan2f(qw_x, rho - qw_y);
  while (thetac < 0) thetac = thetac + 2 * M_PI;
  float df = sqrtf(qw_x * qw_x + (qw_y + rho) * (qw_y + rho));
  float alpha = 2 * M_PI - acosf((5 * rho * rho - df * df) / (4 * rho * rho));
  int q_in_Dp = 0;
  if ((qw_x * qw_x + (qw_y + rho) * (qw_y + rho) <= rho * rho) |
      (qw_x * qw_x + (qw_y - rho) * (qw_y - rho) <= rho * rho))
    q_in_Dp = 1;
  if (!q_in_Dp) return sqrtf(dc * dc - rho * rho) + rho * (thetac - acosf(rho / dc));
  else return rho * (alpha + asinf(qw_x / df) - asinf(rho * sinf(alpha) / df));
}
/* feasibleNode :271-289 */
static int feasibleNode(const orc_node* node, double sx, double sy) {
  double angPar = atan2(node->ref_by - node->ref_fy, node->ref_bx - node->ref_fx);
  double angNew = atan2(sy - node->ref_by, sx - node->ref_bx);
  double Lref = sqrt(sq(node->ref_bx - sx) + sq(node->ref_by - sy));
  if (fabs(angleDiff(angNew, angPar)) > (M_PI / 4)) return 0;
  else if (Lref < (2.1 * ref_res)) return 0;
  else return 1;
}
typedef struct { int id; float key; } orc_pair;
static int pair_cmp(const void* a, const void* b) {
  const orc_pair *p = a, *q = b;
  if (p->key < q->key) return -1;
  if (q->key < p->key) return 1;
  return (p->id > q->id) - (p->id < q->id); /* ties: lower node id first */
}
/* The reference sorts with std::sort and a key-only comparator (rrtplanner.cpp:234, :256).  std::sort is
 * unstable, and the reference's trees hold many nodes with identical poses (one-step rollouts from the root),
 * hence identical keys, so WHICH of several tied parents is tried first is decided by the standard library's
 * algorithm, not by the reference's source.  Two tie rules are therefore provided:
 *   tie_mode 0 (default): ties broken by lower node id — the rule the GPU product implements and documents;
 *   tie_mode 1: a restatement of libstdc++'s std::sort (GCC 13 bits/stl_algo.h: introsort with median-of-3
 *               pivot, 16-element threshold, final insertion sort, heapsort fallback), which reproduces the
 *               tie order of the reference binary built in this container; used to pin the oracle against
 *               whole-tree growth of the reference for arbitrary seeds (tests/test_oracle_vs_reference.py). */
static int tie_mode = 0;
static int n_search = -1; /* nodes visible to the candidate search; -1 = whole tree (snapshot rounds set it) */
#define LESS(a, b) ((a).key < (b).key)
static void sl_swap(orc_pair* a, orc_pair* b) { orc_pair t = *a; *a = *b; *b = t; }
static void sl_unguarded_linear_insert(orc_pair* last) {
  orc_pair val = *last;
  orc_pair* next = last - 1;
  while (LESS(val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void sl_insertion_sort(orc_pair* first, orc_pair* last) {
  if (first == last) return;
  for (orc_pair* i = first + 1; i != last; ++i) {
    if (LESS(*i, *first)) {
      orc_pair val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof(orc_pair));
      *first = val;
    } else sl_unguarded_linear_insert(i);
  }
}
static void sl_push_heap(orc_pair* first, long hole, long top, orc_pair value) {
  long parent = (hole - 1) / 2;
  while (hole > top && LESS(first[parent], value)) { first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2; }
  first[hole] = value;
}
static void sl_adjust_heap(orc_pair* first, long hole, long len, orc_pair value) {
  const long top = hole;
  long second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (LESS(first[second], first[second - 1])) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  sl_push_heap(first, hole, top, value);
}
static void sl_heapsort(orc_pair* first, orc_pair* last) { /* __partial_sort(first,last,last) */
  long len = last - first;
  if (len >= 2)
    for (long parent = (len - 2) / 2;; parent--) {
      sl_adjust_heap(first, parent, len, first[parent]);
      if (parent == 0) break;
    }
  while (last - first > 1) {
    --last;
    orc_pair value = *last;
    *last = *first;
    sl_adjust_heap(first, 0, last - first, value);
  }
}
static void sl_introsort_loop(orc_pair* first, orc_pair* last, long depth) {
  while (last - first > 16) {
    if (depth == 0) { sl_heapsort(first, last); return; }
    --depth;
    orc_pair *mid = first + (last - first) / 2, *a = first + 1, *b = mid, *c = last - 1;
    if (LESS(*a, *b)) { /* __move_median_to_first */
      if (LESS(*b, *c)) sl_swap(first, b);
      else if (LESS(*a, *c)) sl_swap(first, c);
      else sl_swap(first, a);
    } else if (LESS(*a, *c)) sl_swap(first, a);
    else if (LESS(*b, *c)) sl_swap(first, c);
    else sl_swap(first, b);
    orc_pair *lo = first + 1, *hi = last; /* __unguarded_partition(first+1, last, first) */
    for (;;) {
      while (LESS(*lo, *first)) ++lo;
      --hi;
      while (LESS(*first, *hi)) --hi;
      if (!(lo < hi)) break;
      sl_swap(lo, hi);
      ++lo;
    }
    sl_introsort_loop(lo, last, depth);
    last = lo;
  }
}
static void libstdcxx_sort(orc_pair* first, long n) {
  if (n == 0) return;
  long lg = 0;
  for (long m = n; m > 1; m >>= 1) lg++;
  sl_introsort_loop(first, first + n, 2 * lg);
  if (n > 16) {
    sl_insertion_sort(first, first + 16);
    for (orc_pair* i = first + 16; i != first + n; ++i) sl_unguarded_linear_insert(i);
  } else sl_insertion_sort(first, first + n);
}
/* sortNodesExplore :227-247 (heuristic 0) / sortNodesOptimize :250-268 (heuristic 1) */
static int sortNodes(double sx, double sy, int heuristic, int* out, float* keyout) {
  const int n_vis = n_search >= 0 ? n_search : n_tree;
  orc_pair* d = malloc(sizeof(orc_pair) * (size_t)n_vis);
  for (int i = 0; i < n_vis; i++) {
    d[i].id = i;
    float k = dubinsDistance(sx, sy, tree[i].state, 1);
    d[i].key = heuristic ? tree[i].costE + k : k;
  }
  if (tie_mode == 1) libstdcxx_sort(d, n_vis);
  else qsort(d, (size_t)n_vis, sizeof(orc_pair), pair_cmp);
  int n = 0;
  for (int i = 0; i < n_vis; i++) {
    if (feasibleNode(&tree[d[i].id], sx, sy)) {
      out[n] = d[i].id;
      if (keyout) keyout[n] = d[i].key;
      n++;
    }
    if (n == ORC_SORT_LIMIT) break;
  }
  free(d);
  return n;
}
/* feasibleGoalBias :292-315 */
static int feasibleGoalBias(void) {
  double R1 = 4.77, R2 = R1 - 0.3;
  double clx = goalPose[0] + R1 * cos(goalPose[2] - M_PI_2);
  double cly = goalPose[1] + R1 * cos(goalPose[2] - M_PI_2);
  double crx = goalPose[0] + R1 * cos(goalPose[2] + M_PI_2);
  double cry = goalPose[1] + R1 * cos(goalPose[2] + M_PI_2);
  const orc_node* node = &tree[n_tree - 1];
  int outside_left = sqrt(sq(node->state[0] - clx) + sq(node->state[1] - cly)) > R2;
  int outside_right = sqrt(sq(node->state[0] - crx) + sq(node->state[1] - cry)) > R2;
  double angleRef = atan2(goalPose[1] - node->ref_by, goalPose[0] - node->ref_bx);
  double dHead1 = fabs(wrapToPi(goalPose[2] - angleRef));
  double dHead2 = fabs(wrapToPi(goalPose[2] + M_PI - angleRef));
  double minAngleDiff = std_min(dHead1, dHead2);
  double c = cos(goalPose[2] + M_PI_2 - angleRef);
  double sgn = (double)((0.0 < c) - (c < 0.0));
  double angle = sgn * minAngleDiff;
  int within = fabs(angle) < (M_PI_4 / 2);
  return outside_left * outside_right * within;
}
static void addNode(const orc_node* nd) {
  if (n_tree == cap_tree) {
    cap_tree = cap_tree ? 2 * cap_tree : 1024;
    tree = realloc(tree, sizeof(orc_node) * (size_t)cap_tree);
  }
  tree[n_tree++] = *nd;
}
/* Node ctor rrtplanner.h:45 from a finished Simulation, as at rrtplanner.cpp:156 / :170 */
static void nodeFromSim(orc_node* nd, const orc_sim* s, const orc_ref* ref, int parent) {
  memcpy(nd->state, s->xf, sizeof nd->state);
  nd->ref_fx = ref->x[0]; nd->ref_fy = ref->y[0];
  nd->ref_bx = ref->x[ref->N - 1]; nd->ref_by = ref->y[ref->N - 1];
  nd->ref_vback = ref->v[ref->N - 1];
  nd->costE = (float)(s->costE + tree[parent].costE);
  nd->costS = (float)(s->costS + tree[parent].costS);
  nd->parent = parent; nd->goal = s->goalReached; nd->nref = ref->N;
}
/* sampleAroundVehicle :187-201 */
static void sampleAroundVehicle(const double* goal, double* sx, double* sy) {
  double dGoal = sqrt(sq(goal[0]) + sq(goal[1]));
  double goalHeading = atan2(goal[1], goal[0]);
  double latMin = -7, latMax = 7;
  double rLong = (float)(rand()) / ((float)(RAND_MAX / (dGoal + 10)));
  double rLat = latMin + (float)(rand()) / ((float)(RAND_MAX / (latMax - latMin)));
  *sx = rLong * cos(goalHeading) + rLat * cos(goalHeading + M_PI / 2);
  *sy = rLong * sin(goalHeading) + rLat * sin(goalHeading + M_PI / 2);
}
/* expandTree :123-174 with a caller-supplied sample / heuristic draw */
static void expandTreeWith(double sx, double sy, int heuristic) {
  int sorted[ORC_SORT_LIMIT];
  int ns = sortNodes(sx, sy, heuristic, sorted, NULL);
  int node_added = 0;
  for (int r = 0; r < ns; r++) {
    const int p = sorted[r];
    orc_ref ref = getReference(sx, sy, &tree[p], 1);
    orc_sim sim;
    simulate(tree[p].state, &ref, 0, tree[p].ref_vback, &sim, NULL, 0);
    if (sim.endReached || sim.goalReached) {
      orc_node nd;
      nodeFromSim(&nd, &sim, &ref, p);
      addNode(&nd);
      node_added = 1;
      ref_free(&ref);
      break;
    }
    ref_free(&ref);
  }
  if (node_added && feasibleGoalBias()) {
    const int p = n_tree - 1;
    orc_ref ref = getGoalReference(&tree[p], goalPose);
    orc_sim sim;
    simulate(tree[p].state, &ref, 1, tree[p].ref_vback, &sim, NULL, 0);
    if (sim.endReached || sim.goalReached) {
      orc_node nd;
      nodeFromSim(&nd, &sim, &ref, p);
      addNode(&nd);
    }
    ref_free(&ref);
  }
}
static void expandTree(void) {
  double sx, sy;
  sampleAroundVehicle(goalPose, &sx, &sy);
  double r = (double)(rand()) / ((double)(RAND_MAX / (1)));
  expandTreeWith(sx, sy, !(r <= 0.7)); /* :142-147, RRT.goalReached stays 0 */
}

/* =============================== flat C interface (mirrors oracle/ref_driver.cpp) ===================== */
static void sim_to_out(double* o, const orc_sim* s) {
  for (int k = 0; k < 10; k++) o[k] = s->xf[k];
  o[10] = s->costE; o[11] = s->costS; o[12] = s->endReached; o[13] = s->goalReached;
  o[14] = s->n_steps; o[15] = s->fail; o[16] = s->N; o[17] = s->ref_bx; o[18] = s->ref_by; o[19] = s->ref_vback;
  o[20] = s->tainted; o[21] = s->trace; o[22] = s->idwp0; o[23] = 0;
}
static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void orc_init(void) {
  /* rrt/launch/parameters.launch:3-20 */
  ctrl_tla = 1.4; ctrl_mindla = 3.2; ctrl_dlavmin = 3; ref_int = 0.02; ref_mindist = 0.2;
  sim_dt = 0.04; ctrl_Kp = 8; ctrl_Ki = 0.05;
  Wcost[0] = 10; Wcost[1] = 5; Wcost[2] = 0; Wcost[3] = 4; Wcost[4] = 1;
  ay_road_max = 0; vmax = 5; obs_use_pred = 1;
  /* Vehicle::setPrius, rrt/include/rrt/vehicle.h:39-60 */
  veh.dmax = 0.52; veh.ddmax = 0.3294; veh.Td = 0.3; veh.Ta = 0.3; veh.amin = -6; veh.amax = 2; veh.L = 2.7;
  double lf = 1.0868, lr = 1.6132;
  veh.Lrear = 1; veh.Lfront = 2.7 + 0.5; veh.w = 2; veh.b = lr; veh.rho = 5.95;
  double Cf = 22201, Cr = 22201, m = 950 + 640;
  veh.Kus = (m / veh.L) * (lr / Cf - lf / Cr);
  veh.Vch = 20;
  /* the library keeps file-scope state (like the reference): a new planner object starts from the defaults */
  tie_mode = 0; gen_profile_flag = 1; road_bend = 0;
}
void orc_set_tie_mode(int m) { tie_mode = m; }
void orc_set_weights(const double* w5) { memcpy(Wcost, w5, sizeof Wcost); }
/* MotionRequest.bend / laneShifts[0] / Cxy, rrt/src/motionplanner.cpp:23 */
void orc_set_road(int bend, const double* Cxy3, double lane_shift) {
  road_bend = bend; road_S = lane_shift; memcpy(road_Cxy, Cxy3, sizeof road_Cxy);
}
void orc_get_vehicle(double* v14) { memcpy(v14, &veh, sizeof veh); }
void orc_srand(unsigned seed) { srand(seed); }
void orc_set_obstacles(const double* o7, int n) {
  free(det);
  det = NULL;
  n_det = n;
  if (n > 0) {
    det = malloc(sizeof(orc_obstacle) * (size_t)n);
    memcpy(det, o7, sizeof(orc_obstacle) * (size_t)n);
  }
}
/* MotionPlanner::planMotion prologue, rrt/src/motionplanner.cpp:9-32 (commit_path=false),
   MyRRT::addInitialNode rrtplanner.cpp:21-37, initializeTree :39-48 */
void orc_tree_init(const double* car_state6, const double* goal4, double vmax_) {
  fail_acclimit = fail_collision = fail_iterlimit = sim_count = 0;
  double carPose[10] = {0, 0, 0, car_state6[3], car_state6[4], car_state6[5], 0, 0, 0, 0};
  updateLookahead(carPose[4]);
  updateReferenceResolution(carPose[4]);
  memcpy(goalPose, goal4, sizeof goalPose);
  vmax = vmax_;
  n_tree = 0;
  orc_node root;
  memcpy(root.state, carPose, sizeof carPose);
  double xend = 1, yend = 0, res = 0.1;
  int N = (int)floor(sqrt(sq(xend) + sq(yend)) / res);
  double* tmp = malloc(sizeof(double) * (size_t)N);
  linspace(0, xend, (size_t)N, tmp);
  root.ref_fx = tmp[0]; root.ref_bx = tmp[N - 1];
  linspace(0, yend, (size_t)N, tmp);
  root.ref_fy = tmp[0]; root.ref_by = tmp[N - 1];
  free(tmp);
  root.ref_vback = carPose[4];
  root.costE = 0; root.costS = 0; root.parent = -1; root.goal = 0; root.nref = N;
  addNode(&root);
}
double orc_get_ref_res(void) { return ref_res; }
int orc_expand(int iters) {
  for (int i = 0; i < iters; i++) expandTree();
  return n_tree;
}
/* Timer, rrt/include/rrt/rrtplanner.h:11-25: CPU time via clock(), polled once per iteration */
int orc_expand_timed(double budget_ms, int* iters_out) {
  clock_t t0 = clock();
  int iter = 0;
  for (;; iter++) {
    double diffms = ((double)(clock() - t0)) / (CLOCKS_PER_SEC / 1000);
    if (!(diffms < budget_ms)) break;
    expandTree();
  }
  if (iters_out) *iters_out = iter;
  return n_tree;
}
/* expandTree with caller-supplied samples (K=1 sequential semantics, one sample after the other) */
int orc_expand_with(const double* sample_xy, const unsigned char* heuristic, int K) {
  for (int j = 0; j < K; j++) expandTreeWith(sample_xy[2 * j], sample_xy[2 * j + 1], heuristic[j]);
  return n_tree;
}
/* A round of K samples against ONE tree snapshot (the batched formulation of the GPU product, SURVEY.md §0
   fact 4): every sample sees only the nodes present when the round started; accepted nodes (and their goal-biased
   children) are appended in sample order.  K == 1 is expandTree itself. */
int orc_expand_round(const double* sample_xy, const unsigned char* heuristic, int K) {
  n_search = n_tree;
  for (int j = 0; j < K; j++) expandTreeWith(sample_xy[2 * j], sample_xy[2 * j + 1], heuristic[j]);
  n_search = -1;
  return n_tree;
}
int orc_tree_size(void) { return n_tree; }
void orc_counters(int* c4) { c4[0] = fail_collision; c4[1] = fail_acclimit; c4[2] = fail_iterlimit; c4[3] = sim_count; }
void orc_work_counters(long* c3) { c3[0] = sat_calls; c3[1] = sat_axes; c3[2] = wp_scanned; }
void orc_tree_export(double* out, int cap) {
  int n = cap < n_tree ? cap : n_tree;
  for (int i = 0; i < n; i++) {
    const orc_node* nd = &tree[i];
    double* o = out + (size_t)ORC_NODE_STRIDE * i;
    for (int k = 0; k < 10; k++) o[k] = nd->state[k];
    o[10] = nd->ref_fx; o[11] = nd->ref_fy; o[12] = nd->ref_bx; o[13] = nd->ref_by; o[14] = nd->ref_vback;
    o[15] = nd->costE; o[16] = nd->costS; o[17] = nd->parent; o[18] = nd->goal; o[19] = nd->nref;
  }
}
void orc_tree_import(const double* in, int n) {
  n_tree = 0;
  for (int i = 0; i < n; i++) {
    const double* o = in + (size_t)ORC_NODE_STRIDE * i;
    orc_node nd;
    memcpy(nd.state, o, sizeof nd.state);
    nd.ref_fx = o[10]; nd.ref_fy = o[11]; nd.ref_bx = o[12]; nd.ref_by = o[13]; nd.ref_vback = o[14];
    nd.costE = (float)o[15]; nd.costS = (float)o[16]; nd.parent = (int)o[17]; nd.goal = o[18] != 0; nd.nref = (int)o[19];
    addNode(&nd);
  }
}
double orc_rollout_batch(const int* parent, const double* sample_xy, const unsigned char* gb, int M, double* out) {
  double t0 = now_s();
  for (int j = 0; j < M; j++) {
    const orc_node* p = &tree[parent[j]];
    const int g = gb && gb[j];
    orc_ref ref = g ? getGoalReference(p, goalPose) : getReference(sample_xy[2 * j], sample_xy[2 * j + 1], p, 1);
    orc_sim sim;
    simulate(p->state, &ref, g, p->ref_vback, &sim, NULL, 0);
    sim_to_out(out + (size_t)ORC_OUT_STRIDE * j, &sim);
    ref_free(&ref);
  }
  return now_s() - t0;
}
int orc_rollout_traj(int parent, const double* sample_xy, int gb, double* traj, int cap, double* refv, int vcap) {
  const orc_node* p = &tree[parent];
  orc_ref ref = gb ? getGoalReference(p, goalPose) : getReference(sample_xy[0], sample_xy[1], p, 1);
  orc_sim sim;
  simulate(p->state, &ref, gb != 0, p->ref_vback, &sim, traj, cap);
  if (refv) for (int i = 0; i < (vcap < ref.N ? vcap : ref.N); i++) refv[i] = ref.v[i];
  ref_free(&ref);
  return sim.n_steps + 1;
}
double orc_nearest_batch(const double* sample_xy, const unsigned char* heuristic, int K, int* cand, float* key,
                         int* count) {
  double t0 = now_s();
  for (int j = 0; j < K; j++) {
    int ids[ORC_SORT_LIMIT];
    float ks[ORC_SORT_LIMIT];
    int n = sortNodes(sample_xy[2 * j], sample_xy[2 * j + 1], heuristic[j], ids, ks);
    count[j] = n;
    for (int r = 0; r < ORC_SORT_LIMIT; r++) {
      cand[ORC_SORT_LIMIT * j + r] = r < n ? ids[r] : -1;
      key[ORC_SORT_LIMIT * j + r] = r < n ? ks[r] : 0;
    }
  }
  return now_s() - t0;
}
void orc_keys(const double* sample_xy, int heuristic, float* key, unsigned char* feas) {
  for (int i = 0; i < n_tree; i++) {
    float k = dubinsDistance(sample_xy[0], sample_xy[1], tree[i].state, 1);
    key[i] = heuristic ? tree[i].costE + k : k;
    feas[i] = (unsigned char)feasibleNode(&tree[i], sample_xy[0], sample_xy[1]);
  }
}
float orc_dubins(double sx, double sy, double nx, double ny, double nth, int dir) {
  double st[3] = {nx, ny, nth};
  return dubinsDistance(sx, sy, st, dir);
}
double orc_obb_dist(const double* a5, const double* b5) {
  orc_obb a, b;
  obb_init(&a, a5[0], a5[1], (float)a5[2], (float)a5[3], (float)a5[4]);
  obb_init(&b, b5[0], b5[1], (float)b5[2], (float)b5[3], (float)b5[4]);
  return getOBBdist(a, b);
}
double orc_obs_distance(const double* x10) { return obsDistance3(x10); }
/* checkObsDistance for n poses (x, y, theta, t): out[i] = the distance the hook at simulation.cpp:83 sees */
void orc_obs_distance_batch(const double* pose4, int n, double* out) {
  for (int i = 0; i < n; i++) {
    double x[10] = {pose4[4 * i], pose4[4 * i + 1], pose4[4 * i + 2], 0, 0, 0, pose4[4 * i + 3], 0, 0, 0};
    out[i] = checkObsDistance(x);
  }
}
/* Simulation::Simulation(RRT, state, ref, veh, GoalBiased, genProfile, Vstart) on a caller-supplied reference
 * (simulation.h:18-19): rx, ry of n points, rv filled when genProfile (else read).  out: 24 doubles (sim_to_out);
 * traj optional, cap x 10.  Returns the number of stateArray rows. */
int orc_simulate(const double* state10, const double* rx, const double* ry, double* rv, int n, int dir, int gb, int genProfile,
                 double Vstart, double* out, double* traj, int cap) {
  orc_ref ref;
  ref.N = n; ref.dir = dir;
  ref.x = malloc(sizeof(double) * (size_t)n); ref.y = malloc(sizeof(double) * (size_t)n); ref.v = NULL;
  memcpy(ref.x, rx, sizeof(double) * (size_t)n); memcpy(ref.y, ry, sizeof(double) * (size_t)n);
  if (!genProfile) { ref.v = malloc(sizeof(double) * (size_t)n); memcpy(ref.v, rv, sizeof(double) * (size_t)n); }
  orc_sim sim;
  gen_profile_flag = genProfile;
  simulate(state10, &ref, gb, Vstart, &sim, traj, cap);
  gen_profile_flag = 1;
  sim_to_out(out, &sim);
  memcpy(rv, ref.v, sizeof(double) * (size_t)n);
  ref_free(&ref);
  return sim.n_steps + 1;
}
void orc_draw_samples(int K, double* sample_xy, unsigned char* heuristic, double* r_out) {
  for (int j = 0; j < K; j++) {
    sampleAroundVehicle(goalPose, &sample_xy[2 * j], &sample_xy[2 * j + 1]);
    double r = (double)(rand()) / ((double)(RAND_MAX / (1)));
    heuristic[j] = !(r <= 0.7);
    if (r_out) r_out[j] = r;
  }
}
int orc_feasible_goal_bias(void) { return feasibleGoalBias(); }
/* extractBestPath, rrt/src/rrtplanner.cpp:318-368: cheapest (float costS) goal node, then parents */
int orc_best_path(int* ids, int cap) {
  int best = -1;
  for (int i = 0; i < n_tree; i++)
    if (tree[i].goal && (best < 0 || (double)tree[i].costS < (double)tree[best].costS)) best = i;
  if (best < 0) return 0;
  int len = 0;
  for (int id = best; id != -1; id = tree[id].parent) len++;
  int k = len;
  for (int id = best; id != -1; id = tree[id].parent) {
    k--;
    if (k < cap) ids[k] = id;
  }
  return len;
}
