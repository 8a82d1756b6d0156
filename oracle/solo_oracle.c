/*
 * solo_oracle.c — CPU ORACLE (test infrastructure, NOT product code). See solo_oracle.h.
 *
 * PARITY UNPINNED for the physics (no PyBullet in this image, no golden vectors in the
 * reference).  Env arithmetic follows the reference line by line; citations are
 * file:line into the reference tree.  "[3P]" marks behaviour of the third-party
 * PyBullet/Bullet dependency restated from its published algorithm.
 */
#include "solo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define MAXL SOLO_MAX_LINKS
#define MAXD ORACLE_MAX_DOF
#define MAXC 16                   /* contact points: 4 feet + 4 knees + 8 base-box corners */
#define MAXROWS (3 * MAXC)

/* ------------------------------------------------------------------ small algebra */
typedef double v3[3];
typedef double m3[3][3];
typedef double v6[6];     /* (angular, linear) */
typedef double m6[6][6];

static void v3set(v3 a, double x, double y, double z) { a[0] = x; a[1] = y; a[2] = z; }
static void v3cpy(v3 a, const v3 b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static double v3dot(const v3 a, const v3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double v3norm(const v3 a) { return sqrt(v3dot(a, a)); }
static void v3cross(v3 o, const v3 a, const v3 b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static void m3mulv(v3 o, const m3 A, const v3 x) {
  v3 t;
  for (int i = 0; i < 3; i++) t[i] = A[i][0] * x[0] + A[i][1] * x[1] + A[i][2] * x[2];
  v3cpy(o, t);
}
static void m3Tmulv(v3 o, const m3 A, const v3 x) {
  v3 t;
  for (int i = 0; i < 3; i++) t[i] = A[0][i] * x[0] + A[1][i] * x[1] + A[2][i] * x[2];
  v3cpy(o, t);
}
static void m3mul(m3 o, const m3 A, const m3 B) {
  m3 t;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
  memcpy(o, t, sizeof(m3));
}
static void m3T(m3 o, const m3 A) {
  m3 t;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i][j] = A[j][i];
  memcpy(o, t, sizeof(m3));
}
static void m3sym6(m3 o, const double I[6]) { /* ixx ixy ixz iyy iyz izz */
  o[0][0] = I[0]; o[0][1] = I[1]; o[0][2] = I[2];
  o[1][0] = I[1]; o[1][1] = I[3]; o[1][2] = I[4];
  o[2][0] = I[2]; o[2][1] = I[4]; o[2][2] = I[5];
}
static void m3skew(m3 o, const v3 a) {
  o[0][0] = 0; o[0][1] = -a[2]; o[0][2] = a[1];
  o[1][0] = a[2]; o[1][1] = 0; o[1][2] = -a[0];
  o[2][0] = -a[1]; o[2][1] = a[0]; o[2][2] = 0;
}
/* rotation matrix of angle th about unit axis a (Rodrigues): maps link coords at q=th to parent */
static void m3axisangle(m3 R, const v3 a, double th) {
  double c = cos(th), s = sin(th), C = 1.0 - c;
  R[0][0] = c + a[0] * a[0] * C;        R[0][1] = a[0] * a[1] * C - a[2] * s; R[0][2] = a[0] * a[2] * C + a[1] * s;
  R[1][0] = a[1] * a[0] * C + a[2] * s; R[1][1] = c + a[1] * a[1] * C;        R[1][2] = a[1] * a[2] * C - a[0] * s;
  R[2][0] = a[2] * a[0] * C - a[1] * s; R[2][1] = a[2] * a[1] * C + a[0] * s; R[2][2] = c + a[2] * a[2] * C;
}
/* quaternion (x,y,z,w), local->world */
static void quat_to_m3(m3 R, const double q[4]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - w * z);     R[0][2] = 2 * (x * z + w * y);
  R[1][0] = 2 * (x * y + w * z);     R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - w * x);
  R[2][0] = 2 * (x * z - w * y);     R[2][1] = 2 * (y * z + w * x);     R[2][2] = 1 - 2 * (x * x + y * y);
}
static void quat_mul(double o[4], const double a[4], const double b[4]) { /* a (x) b */
  double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
/* [3P] p.getEulerFromQuaternion = btQuaternion::getEulerZYX -> (roll, pitch, yaw) */
static void quat_to_euler(v3 rpy, const double q[4]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double sqx = x * x, sqy = y * y, sqz = z * z, sqw = w * w;
  double sarg = -2.0 * (x * z - w * y);
  if (sarg <= -0.99999) {
    rpy[1] = -0.5 * M_PI; rpy[0] = 0; rpy[2] = 2 * atan2(x, -y);
  } else if (sarg >= 0.99999) {
    rpy[1] = 0.5 * M_PI; rpy[0] = 0; rpy[2] = 2 * atan2(-x, y);
  } else {
    rpy[1] = asin(sarg);
    rpy[0] = atan2(2 * (y * z + w * x), sqw - sqx - sqy + sqz);
    rpy[2] = atan2(2 * (x * y + w * z), sqw + sqx - sqy - sqz);
  }
}
static double pymod(double a, double b) { /* Python float modulo, result has the sign of b */
  double r = fmod(a, b);
  if (r != 0.0 && ((r < 0) != (b < 0))) r += b;
  return r;
}
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

static void v6zero(v6 a) { for (int i = 0; i < 6; i++) a[i] = 0; }
static double v6dot(const v6 a, const v6 b) {
  double s = 0; for (int i = 0; i < 6; i++) s += a[i] * b[i]; return s;
}
static void m6zero(m6 A) { memset(A, 0, sizeof(m6)); }
static void m6mulv(v6 o, const m6 A, const v6 x) {
  v6 t;
  for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += A[i][j] * x[j]; t[i] = s; }
  memcpy(o, t, sizeof(v6));
}
static void m6Tmulv_add(v6 o, const m6 A, const v6 x) { /* o += A^T x */
  for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += A[j][i] * x[j]; o[i] += s; }
}
static void m6_XtAX_add(m6 O, const m6 X, const m6 A) { /* O += X^T A X */
  m6 T;
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += A[i][k] * X[k][j]; T[i][j] = s; }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += X[k][i] * T[k][j]; O[i][j] += s; }
}
/* spatial motion cross product  o = a x b  (both motion vectors) */
static void v6crossm(v6 o, const v6 a, const v6 b) {
  v3 t1, t2, t3;
  v3cross(t1, a, b);
  v3cross(t2, a, b + 3);
  v3cross(t3, a + 3, b);
  o[0] = t1[0]; o[1] = t1[1]; o[2] = t1[2];
  o[3] = t2[0] + t3[0]; o[4] = t2[1] + t3[1]; o[5] = t2[2] + t3[2];
}
/* dense symmetric positive definite solve (n <= MAXD), A destroyed */
static void chol_solve(int n, double* A, int lda, double* b) {
  for (int j = 0; j < n; j++) {
    double d = A[j * lda + j];
    for (int k = 0; k < j; k++) d -= A[j * lda + k] * A[j * lda + k];
    d = sqrt(d);
    A[j * lda + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i * lda + j];
      for (int k = 0; k < j; k++) s -= A[i * lda + k] * A[j * lda + k];
      A[i * lda + j] = s / d;
    }
  }
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= A[i * lda + k] * b[k];
    b[i] = s / A[i * lda + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= A[k * lda + i] * b[k];
    b[i] = s / A[i * lda + i];
  }
}

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

/* ------------------------------------------------------------------ env */
struct OracleEnv {
  SoloModelTable m;
  SoloSimParams p;
  int nj, act_dim, d0, d;
  int dof_of_link[MAXL]; /* -1 for fixed */
  int link_of_dof[MAXL];
  /* state */
  v3 pos; double quat[4]; v3 vlin; v3 vang; /* world */
  double q[MAXL], qd[MAXL];                 /* per link (0 for fixed) */
  /* contact record of the last substep, per foot */
  int c_has[SOLO_MAX_FEET];
  double c_force[SOLO_MAX_FEET];
  /* env bookkeeping (baseEnv.py:30-38) */
  int need_reset, timestep, goals_reached_env;
  double reward_sum, dr[5]; /* stand, joint_pose, torque, balance, progress */
  double hist[ORACLE_MAX_HIST][ORACLE_MAX_D0]; /* hist[0] = newest */
  /* pointgoal (solo.py:138-145) */
  double goal[2], potential, progress, goal_radius;
  int goals_reached;
  /* rng */
  uint64_t seed; int64_t env_id; uint32_t episode, draw;
  int settle_last;
  int last_iters;        /* PGS iterations used by the last substep with contacts */
  int last_limit_rows;   /* joint-limit rows of the last substep */
  v3 fext;               /* external force on the base, base axes, at the base origin (gait-env pushes) */
};

/* per-step workspace of the articulated-body algorithm, Bullet-style: every spatial
 * quantity of link i lives in the link's COM frame (axes = link axes) */
typedef struct {
  m3 E0;                 /* world -> base */
  m3 Rw[MAXL];           /* link -> world */
  v3 pw[MAXL];           /* link COM in world */
  m6 X[MAXL];            /* motion transform parent COM frame -> link COM frame */
  v6 S[MAXL], v[MAXL], c[MAXL], Z[MAXL], h[MAXL], a[MAXL];
  m6 IA[MAXL];
  double D[MAXL], u[MAXL];
  v6 v0, Z0, a0;
  m6 IA0;
  double L0[36];         /* cholesky factor of IA0 */
} AbaWork;

void oracle_default_params(SoloSimParams* p) {
  memset(p, 0, sizeof(*p));
  p->abi_version = SOLO_ABI_VERSION;
  p->dt = 1.0 / 240.0;            /* solo.py:22 */
  p->frame_skip = 4;              /* baseEnv.py:9 */
  p->gravity_z = -9.81;           /* simulation.py:19 */
  p->lin_damping = 0.04;          /* [3P] btMultiBody default */
  p->ang_damping = 0.04;          /* [3P] */
  p->max_coord_vel = 100.0;       /* [3P] m_maxCoordinateVelocity */
  p->solver_iters = 50;           /* [3P] PyBullet numSolverIterations */
  p->solver_residual_threshold = 1e-7; /* [3P] PyBullet solverResidualThreshold */
  p->contact_erp = 0.2;           /* [3P] */
  p->contact_slop = 1e-5;         /* [3P] */
  p->contact_margin = 0.02;       /* [3P] contact breaking threshold */
  p->friction = 1.0;              /* solo.urdf:37-40 x plane.urdf */
  p->cone_friction = 1;           /* [3P] */
  p->torque_hold = 0;             /* SURVEY F4 */
  p->control = SOLO_CONTROL_TORQUE;
  p->kp = 0; p->kd = 0;
  p->max_torque = 3.0;            /* solo.py:53 */
  p->joint_state_limit = 10.0;    /* solo.py:109 */
  p->joint_vel_limit = 100.0;     /* solo.py:110 */
  p->task = SOLO_TASK_STAND;      /* baseEnv.py:11 */
  p->episode_length = 400;
  p->num_history_stack = 0;       /* baseEnv.py:15 */
  p->initial_z = 0.35;            /* solo.py:52 */
  p->settle_min = 5; p->settle_max = 12; /* baseEnv.py:79 */
  p->goal_radius = 2.0;           /* solo.py:141 */
  p->goal_reach_dist = 0.5;       /* solo.py:270 */
  p->pointgoal_dt = 4.0 / 240.0;  /* intent of the undefined scene.dt, baseEnv.py:137 */
  p->contact_flag_force = 0.2;    /* solo.py:320 */
  p->fall_z = 0.05;               /* baseEnv.py:169 */
  p->stand_z = 0.2;               /* baseEnv.py:96 */
  p->reset_mode = SOLO_RESET_CACHED;
  p->joint_limits = 1;
  p->limit_rows_per_leg = 1;
  p->joint_limit_erp = 0.2;
  p->joint_limit_max_impulse = 100.0;
  p->split_impulse_threshold = -0.04;
  p->body_contacts = 0;
  p->knee_radius = 0.015;          /* lower-leg mesh z in [-0.1675, 0.0135] about the KFE axis */
  p->base_half_x = 0.2241; p->base_half_y = 0.1095;   /* Solo12 base box (Solo8: 0.212 x 0.1046) */
  p->base_z_lo = -0.025; p->base_z_hi = 0.028;
}

OracleEnv* oracle_env_create(const SoloModelTable* m, const SoloSimParams* p, uint64_t seed,
                             int64_t env_id) {
  OracleEnv* e = (OracleEnv*)calloc(1, sizeof(OracleEnv));
  e->m = *m;
  e->p = *p;
  int nj = 0;
  for (int i = 0; i < m->num_links; i++) {
    if (m->jtype[i] == SOLO_JOINT_REVOLUTE) { e->dof_of_link[i] = nj; e->link_of_dof[nj] = i; nj++; }
    else e->dof_of_link[i] = -1;
  }
  e->nj = nj;
  e->act_dim = nj + (p->control == SOLO_CONTROL_VPD ? 2 : 0); /* baseEnv.py:21-24 */
  e->d0 = 1 + 3 + 6 + 2 * nj + 4 + (p->task == SOLO_TASK_POINTGOAL ? 4 : 0); /* solo.py:198-222 */
  e->d = e->d0 * (1 + p->num_history_stack);                  /* solo.py:186-196 */
  e->quat[3] = 1.0;
  e->pos[2] = p->initial_z;
  e->need_reset = 1;  /* baseEnv.py:31 */
  e->goal_radius = p->goal_radius;
  e->seed = seed; e->env_id = env_id; e->episode = 0; e->draw = 0;
  return e;
}
void oracle_env_destroy(OracleEnv* e) { free(e); }
int oracle_nj(const OracleEnv* e) { return e->nj; }
int oracle_act_dim(const OracleEnv* e) { return e->act_dim; }
int oracle_obs_dim0(const OracleEnv* e) { return e->d0; }
int oracle_obs_dim(const OracleEnv* e) { return e->d; }
int oracle_settle_count_last(const OracleEnv* e) { return e->settle_last; }
int oracle_last_solver_iters(const OracleEnv* e) { return e->last_iters; }
int oracle_last_limit_rows(const OracleEnv* e) { return e->last_limit_rows; }

static void env_rng(OracleEnv* e, uint32_t w[4]) {
  w[0] = (uint32_t)((uint64_t)e->env_id & 0xffffffffu);
  w[1] = (uint32_t)((uint64_t)e->env_id >> 32);
  w[2] = e->episode;
  w[3] = e->draw++;
  philox4x32_10(w, (uint32_t)(e->seed & 0xffffffffu), (uint32_t)(e->seed >> 32));
}
static double u01(uint32_t w) { return (double)(w >> 8) * (1.0 / 16777216.0); }
/* solo.py:325-330: xy ~ U(1, goal_radius), sign ~ {-1, +1} per axis */
static void sample_goal_from(OracleEnv* e, const uint32_t w[4]) {
  double gx = 1.0 + (e->goal_radius - 1.0) * u01(w[1]);
  double gy = 1.0 + (e->goal_radius - 1.0) * u01(w[2]);
  e->goal[0] = (w[3] & 1u) ? gx : -gx;
  e->goal[1] = (w[3] & 2u) ? gy : -gy;
}

/* ------------------------------------------------------------------ kinematics + ABA */
static void build_kinematics(const OracleEnv* e, AbaWork* W) {
  const SoloModelTable* m = &e->m;
  m3 Rb;
  quat_to_m3(Rb, e->quat);
  m3T(W->E0, Rb);
  for (int i = 0; i < m->num_links; i++) {
    int par = m->parent[i];
    m3 Rj, E;
    if (m->jtype[i] == SOLO_JOINT_REVOLUTE) m3axisangle(Rj, m->axis[i], e->q[i]);
    else { memset(Rj, 0, sizeof(m3)); Rj[0][0] = Rj[1][1] = Rj[2][2] = 1; }
    m3T(E, Rj); /* parent -> this */
    /* r: parent COM -> this COM, in this frame  ([3P] btMultibodyLink cachedRVector) */
    v3 e_par, r;
    const double* pcom = (par < 0) ? m->base_com : m->com[par];
    for (int k = 0; k < 3; k++) e_par[k] = m->origin[i][k] - pcom[k];
    m3mulv(r, E, e_par);
    for (int k = 0; k < 3; k++) r[k] += m->com[i][k];
    /* X = [[E,0],[-r x E, E]] */
    m3 rx, rxE;
    m3skew(rx, r);
    m3mul(rxE, rx, E);
    m6zero(W->X[i]);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        W->X[i][a][b] = E[a][b];
        W->X[i][3 + a][3 + b] = E[a][b];
        W->X[i][3 + a][b] = -rxE[a][b];
      }
    /* S = (axis, axis x d), d = pivot -> COM  ([3P] btMultiBody::setupRevolute) */
    v6zero(W->S[i]);
    if (m->jtype[i] == SOLO_JOINT_REVOLUTE) {
      v3 t;
      v3cross(t, m->axis[i], m->com[i]);
      for (int k = 0; k < 3; k++) { W->S[i][k] = m->axis[i][k]; W->S[i][3 + k] = t[k]; }
    }
    /* world pose of the COM frame */
    if (par < 0) m3mul(W->Rw[i], Rb, Rj);
    else m3mul(W->Rw[i], W->Rw[par], Rj);
    v3 rw;
    m3mulv(rw, W->Rw[i], r);
    const double* ppw = (par < 0) ? e->pos : W->pw[par];
    for (int k = 0; k < 3; k++) W->pw[i][k] = ppw[k] + rw[k];
    if (par < 0) { /* base COM offset (0 for both URDFs) */
      v3 bc; m3mulv(bc, Rb, m->base_com);
      for (int k = 0; k < 3; k++) W->pw[i][k] += bc[k];
    }
  }
}

/* damping + gyroscopic bias of one body in its COM frame ([3P] btMultiBody ABA):
 *   Z += ( I w (k + k|w|) + w x I w ,  m v (k + k|v|) + m w x v )                     */
static void body_bias(v6 Z, const v6 vel, double mass, const m3 I, double klin, double kang) {
  v3 Iw, t;
  m3mulv(Iw, I, vel);
  double wn = v3norm(vel), vn = v3norm(vel + 3);
  v3cross(t, vel, Iw);
  for (int k = 0; k < 3; k++) Z[k] += Iw[k] * (kang + kang * wn) + t[k];
  v3cross(t, vel, vel + 3);
  for (int k = 0; k < 3; k++) Z[3 + k] += mass * vel[3 + k] * (klin + klin * vn) + mass * t[k];
}

/* Forward dynamics by the articulated-body algorithm. tau is per DOF.  with_gravity
 * adds m*g on every body as an external force ([3P] btMultiBodyDynamicsWorld gravity).
 * out[6+nj] = (angular acc world, linear acc world, joint acc). */
static void aba_forward(const OracleEnv* e, AbaWork* W, const double* tau, double* out) {
  const SoloModelTable* m = &e->m;
  const SoloSimParams* p = &e->p;
  int n = m->num_links;
  build_kinematics(e, W);
  v3 gw = {0, 0, p->gravity_z};

  /* base */
  m3mulv(W->v0, W->E0, e->vang);
  m3mulv(W->v0 + 3, W->E0, e->vlin);
  {
    v3 f, fl;
    for (int k = 0; k < 3; k++) f[k] = m->base_mass * gw[k];
    m3mulv(fl, W->E0, f);
    v6zero(W->Z0);
    for (int k = 0; k < 3; k++) W->Z0[3 + k] = -fl[k] - e->fext[k];   /* gravity + external push (base axes) */
    m3 Ib; m3sym6(Ib, m->base_inertia);
    body_bias(W->Z0, W->v0, m->base_mass, Ib, p->lin_damping, p->ang_damping);
    m6zero(W->IA0);
    for (int a = 0; a < 3; a++) {
      for (int b = 0; b < 3; b++) W->IA0[a][b] = Ib[a][b];
      W->IA0[3 + a][3 + a] = m->base_mass;
    }
  }
  /* outward pass: velocities, coriolis, bias forces, rigid inertias */
  for (int i = 0; i < n; i++) {
    int par = m->parent[i];
    const double* vp = (par < 0) ? W->v0 : W->v[par];
    m6mulv(W->v[i], W->X[i], vp);
    v6 vj;
    for (int k = 0; k < 6; k++) vj[k] = W->S[i][k] * e->qd[i];
    for (int k = 0; k < 6; k++) W->v[i][k] += vj[k];
    v6crossm(W->c[i], W->v[i], vj);
    v3 f, fl;
    for (int k = 0; k < 3; k++) f[k] = m->mass[i] * gw[k];
    m3Tmulv(fl, W->Rw[i], f);
    v6zero(W->Z[i]);
    for (int k = 0; k < 3; k++) W->Z[i][3 + k] = -fl[k];
    m3 Il; m3sym6(Il, m->inertia[i]);
    body_bias(W->Z[i], W->v[i], m->mass[i], Il, p->lin_damping, p->ang_damping);
    m6zero(W->IA[i]);
    for (int a = 0; a < 3; a++) {
      for (int b = 0; b < 3; b++) W->IA[i][a][b] = Il[a][b];
      W->IA[i][3 + a][3 + a] = m->mass[i];
    }
  }
  /* inward pass */
  for (int i = n - 1; i >= 0; i--) {
    int par = m->parent[i];
    m6 Ia;
    v6 pa, Ic;
    memcpy(Ia, W->IA[i], sizeof(m6));
    m6mulv(Ic, W->IA[i], W->c[i]);
    for (int k = 0; k < 6; k++) pa[k] = W->Z[i][k] + Ic[k];
    if (m->jtype[i] == SOLO_JOINT_REVOLUTE) {
      m6mulv(W->h[i], W->IA[i], W->S[i]);
      W->D[i] = v6dot(W->S[i], W->h[i]);
      /* [3P] Y = tau - S.Z - h.c */
      W->u[i] = tau[e->dof_of_link[i]] - v6dot(W->S[i], W->Z[i]) - v6dot(W->h[i], W->c[i]);
      double invD = 1.0 / W->D[i];
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) Ia[a][b] -= W->h[i][a] * W->h[i][b] * invD;
      for (int k = 0; k < 6; k++) pa[k] += W->h[i][k] * (W->u[i] * invD);
    } else {
      W->D[i] = 1.0; W->u[i] = 0.0; v6zero(W->h[i]);
    }
    if (par < 0) { m6_XtAX_add(W->IA0, W->X[i], Ia); m6Tmulv_add(W->Z0, W->X[i], pa); }
    else { m6_XtAX_add(W->IA[par], W->X[i], Ia); m6Tmulv_add(W->Z[par], W->X[i], pa); }
  }
  /* base acceleration */
  {
    double A[36];
    for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) A[a * 6 + b] = W->IA0[a][b];
    for (int k = 0; k < 6; k++) W->a0[k] = -W->Z0[k];
    chol_solve(6, A, 6, W->a0);
    memcpy(W->L0, A, sizeof(A));
  }
  /* outward pass: accelerations */
  for (int i = 0; i < n; i++) {
    int par = m->parent[i];
    const double* ap = (par < 0) ? W->a0 : W->a[par];
    v6 Xa;
    m6mulv(Xa, W->X[i], ap);
    if (m->jtype[i] == SOLO_JOINT_REVOLUTE) {
      double qdd = (W->u[i] - v6dot(W->h[i], Xa)) / W->D[i];
      out[6 + e->dof_of_link[i]] = qdd;
      for (int k = 0; k < 6; k++) W->a[i][k] = Xa[k] + W->c[i][k] + W->S[i][k] * qdd;
    } else {
      for (int k = 0; k < 6; k++) W->a[i][k] = Xa[k];
    }
  }
  /* base acceleration back to the world frame; the linear part of the spatial
   * acceleration is the derivative of body-frame velocity coordinates, so the classical
   * acceleration of the base origin is a_lin + w x v. */
  {
    v3 t, al;
    v3cross(t, W->v0, W->v0 + 3);
    for (int k = 0; k < 3; k++) al[k] = W->a0[3 + k] + t[k];
    m3Tmulv(out, W->E0, W->a0);
    m3Tmulv(out + 3, W->E0, al);
  }
}

/* Velocity response M^-1 J^T to a unit impulse `dir` (world) applied at world point `pt`
 * on link `link` ([3P] btMultiBody::calcAccelerationDeltasMultiDof); uses the articulated
 * inertias cached by the last aba_forward.  out[6+nj] = (dw world, dv world, dqd). */
/* velocity response (angular, linear, joints) to a unit impulse along `dir` at world point `pt` of `link`
 * (link -1: the base, -2: none) plus a generalized impulse `jimp` on the joint of `jlink` (jlink < 0: none) */
static void impulse_response(const OracleEnv* e, const AbaWork* W, int link, const v3 pt,
                             const v3 dir, int jlink, double jimp, double* out) {
  const SoloModelTable* m = &e->m;
  int n = m->num_links;
  v6 Zt[MAXL], Z0, at[MAXL], a0;
  double ut[MAXL];
  for (int i = 0; i < n; i++) v6zero(Zt[i]);
  v6zero(Z0);
  if (link >= 0) {
    v3 arm, tq, fl, tl;
    for (int k = 0; k < 3; k++) arm[k] = pt[k] - W->pw[link][k];
    v3cross(tq, arm, dir);
    m3Tmulv(fl, W->Rw[link], dir);
    m3Tmulv(tl, W->Rw[link], tq);
    for (int k = 0; k < 3; k++) { Zt[link][k] = -tl[k]; Zt[link][3 + k] = -fl[k]; }
  } else if (link == -1) {              /* a point of the base itself (base COM = base origin) */
    v3 arm, tq, fl, tl;
    for (int k = 0; k < 3; k++) arm[k] = pt[k] - e->pos[k];
    v3cross(tq, arm, dir);
    m3mulv(fl, W->E0, dir);
    m3mulv(tl, W->E0, tq);
    for (int k = 0; k < 3; k++) { Z0[k] = -tl[k]; Z0[3 + k] = -fl[k]; }
  }
  for (int i = n - 1; i >= 0; i--) {
    int par = m->parent[i];
    v6 pa;
    memcpy(pa, Zt[i], sizeof(v6));
    if (m->jtype[i] == SOLO_JOINT_REVOLUTE) {
      ut[i] = -v6dot(W->S[i], Zt[i]) + (i == jlink ? jimp : 0.0);
      for (int k = 0; k < 6; k++) pa[k] += W->h[i][k] * (ut[i] / W->D[i]);
    } else ut[i] = 0;
    if (par < 0) m6Tmulv_add(Z0, W->X[i], pa);
    else m6Tmulv_add(Zt[par], W->X[i], pa);
  }
  /* base: solve with the cached cholesky factor */
  for (int k = 0; k < 6; k++) a0[k] = -Z0[k];
  for (int i = 0; i < 6; i++) {
    double s = a0[i];
    for (int k = 0; k < i; k++) s -= W->L0[i * 6 + k] * a0[k];
    a0[i] = s / W->L0[i * 6 + i];
  }
  for (int i = 5; i >= 0; i--) {
    double s = a0[i];
    for (int k = i + 1; k < 6; k++) s -= W->L0[k * 6 + i] * a0[k];
    a0[i] = s / W->L0[i * 6 + i];
  }
  for (int i = 0; i < n; i++) {
    int par = m->parent[i];
    const double* ap = (par < 0) ? a0 : at[par];
    v6 Xa;
    m6mulv(Xa, W->X[i], ap);
    if (m->jtype[i] == SOLO_JOINT_REVOLUTE) {
      double qdd = (ut[i] - v6dot(W->h[i], Xa)) / W->D[i];
      out[6 + e->dof_of_link[i]] = qdd;
      for (int k = 0; k < 6; k++) at[i][k] = Xa[k] + W->S[i][k] * qdd;
    } else memcpy(at[i], Xa, sizeof(v6));
  }
  m3Tmulv(out, W->E0, a0);
  m3Tmulv(out + 3, W->E0, a0 + 3);
}

/* Jacobian row: velocity of world point pt on `link` along dir = J . (w, v, qd)
 * ([3P] btMultiBody::fillContactJacobianMultiDof) */
static void contact_jacobian(const OracleEnv* e, const AbaWork* W, int link, const v3 pt,
                             const v3 dir, double* J) {
  const SoloModelTable* m = &e->m;
  for (int k = 0; k < 6 + e->nj; k++) J[k] = 0;
  v3 r, t;
  for (int k = 0; k < 3; k++) r[k] = pt[k] - e->pos[k];
  v3cross(t, r, dir);
  for (int k = 0; k < 3; k++) { J[k] = t[k]; J[3 + k] = dir[k]; }
  for (int i = link; i >= 0; i = m->parent[i]) {
    if (m->jtype[i] != SOLO_JOINT_REVOLUTE) continue;
    v3 aw, ow, comw, arm, vel;
    m3mulv(aw, W->Rw[i], m->axis[i]);
    m3mulv(comw, W->Rw[i], m->com[i]);
    for (int k = 0; k < 3; k++) ow[k] = W->pw[i][k] - comw[k]; /* joint pivot in world */
    for (int k = 0; k < 3; k++) arm[k] = pt[k] - ow[k];
    v3cross(vel, aw, arm);
    J[6 + e->dof_of_link[i]] = v3dot(dir, vel);
  }
}

static void clamp_velocities(OracleEnv* e) { /* [3P] btMultiBody::applyDeltaVeeMultiDof clamp */
  double lim = e->p.max_coord_vel;
  for (int k = 0; k < 3; k++) { e->vang[k] = clampd(e->vang[k], -lim, lim); e->vlin[k] = clampd(e->vlin[k], -lim, lim); }
  for (int i = 0; i < e->m.num_links; i++) e->qd[i] = clampd(e->qd[i], -lim, lim);
}

/* [3P] btPlaneSpace1 */
static void plane_space(const v3 n, v3 p, v3 q) {
  if (fabs(n[2]) > 0.7071067811865475244008443621048490) {
    double a = n[1] * n[1] + n[2] * n[2];
    double k = 1.0 / sqrt(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    double a = n[0] * n[0] + n[1] * n[1];
    double k = 1.0 / sqrt(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}

/* Collision detection on the current poses against the plane z = 0: the foot spheres, and with body_contacts the
 * knee spheres (centre = KFE joint origin, carried by the lower-leg link) and the corners of the base box.
 * Point order = row order: feet, knees, lower base corners, upper base corners, legs 0..3 inside each group.
 * Out per contact: link (-1 = base), index of the point (0..15), distance, world contact point. */
static int detect_contacts(const OracleEnv* e, const AbaWork* W, int* clink, int* cidx, double* cdist, v3* cpt) {
  const SoloModelTable* m = &e->m;
  const SoloSimParams* p = &e->p;
  int nc = 0;
  for (int f = 0; f < m->num_feet; f++) {
    int l = m->foot_link[f];
    v3 off, c;
    for (int k = 0; k < 3; k++) off[k] = m->foot_center[f][k] - m->com[l][k];
    m3mulv(c, W->Rw[l], off);
    for (int k = 0; k < 3; k++) c[k] += W->pw[l][k];
    double dist = c[2] - m->foot_radius;
    if (dist < p->contact_margin) {
      clink[nc] = l; cidx[nc] = f; cdist[nc] = dist;
      v3set(cpt[nc], c[0], c[1], c[2] - m->foot_radius);
      nc++;
    }
  }
  if (!p->body_contacts) return nc;
  for (int f = 0; f < m->num_feet; f++) {               /* knees */
    int l = m->parent[m->foot_link[f]];                  /* lower leg: its joint origin is the KFE axis */
    v3 off, c;
    for (int k = 0; k < 3; k++) off[k] = -m->com[l][k];
    m3mulv(c, W->Rw[l], off);
    for (int k = 0; k < 3; k++) c[k] += W->pw[l][k];
    double dist = c[2] - p->knee_radius;
    if (dist < p->contact_margin) {
      clink[nc] = l; cidx[nc] = 4 + f; cdist[nc] = dist;
      v3set(cpt[nc], c[0], c[1], c[2] - p->knee_radius);
      nc++;
    }
  }
  m3 Rb;
  quat_to_m3(Rb, e->quat);
  for (int g = 0; g < 2; g++)                             /* lower corners, then upper corners */
    for (int f = 0; f < 4; f++) {                         /* leg order: FL (+x,+y) FR (+x,-y) HL (-x,+y) HR (-x,-y) */
      v3 loc = {(f < 2 ? 1.0 : -1.0) * p->base_half_x, ((f & 1) ? -1.0 : 1.0) * p->base_half_y,
                g == 0 ? p->base_z_lo : p->base_z_hi};
      v3 c;
      m3mulv(c, Rb, loc);
      for (int k = 0; k < 3; k++) c[k] += e->pos[k];
      if (c[2] < p->contact_margin) {
        clink[nc] = -1; cidx[nc] = 8 + 4 * g + f; cdist[nc] = c[2];
        v3cpy(cpt[nc], c);
        nc++;
      }
    }
  return nc;
}

typedef struct {
  double J[MAXD], u[MAXD]; /* jacobian row and M^-1 J^T */
  double dinv, rhs, lambda;
} Row;

static double rowdot(int n, const double* a, const double* b) {
  double s = 0; for (int k = 0; k < n; k++) s += a[k] * b[k]; return s;
}

/* One Bullet multibody step ([3P], restated):
 *  1 collision detection on the current poses (foot sphere vs plane z=0)
 *  2 ABA with gravity/damping/applied joint torque -> v* = v + dt*qdd (clamped)
 *  3 contact rows: normal + 2 friction directions (btPlaneSpace1), M^-1 J^T per row,
 *    rhs with speculative / ERP penetration term, no warm start
 *  4 PGS, solver_iters iterations: all normals, then friction pairs (implicit cone)
 *  5 v = v* + M^-1 J^T lambda (clamped); q += dt*v (base quaternion by exponential map) */
void oracle_substep(OracleEnv* e, const double* tau) {
  const SoloModelTable* m = &e->m;
  const SoloSimParams* p = &e->p;
  AbaWork W;
  int nd = 6 + e->nj;
  double qdd[MAXD];
  double dt = p->dt;
  aba_forward(e, &W, tau, qdd);

  /* 1 contacts at the pre-integration poses */
  int clink[MAXC], cidx[MAXC];
  v3 cpt[MAXC];
  double cdist[MAXC];
  for (int f = 0; f < m->num_feet; f++) { e->c_has[f] = 0; e->c_force[f] = 0; }
  const int nc = detect_contacts(e, &W, clink, cidx, cdist, cpt);
  for (int c = 0; c < nc; c++) if (cidx[c] < 4) e->c_has[cidx[c]] = 1;
  /* 1b joint-limit rows ([3P] btMultiBodyJointLimitConstraint::createConstraintRows): a row exists while the
   * joint position (start of the step) is at or beyond a limit; row 0 = lower bound, row 1 = upper bound */
  int nl = 0, llink[2 * MAXL];
  double ldir[2 * MAXL], lpen[2 * MAXL];
  if (p->joint_limits) {
    const double lim = p->joint_state_limit;
    for (int i = 0; i < m->num_links; i++) {
      if (m->jtype[i] != SOLO_JOINT_REVOLUTE) continue;
      const double pen_lo = e->q[i] + lim, pen_hi = lim - e->q[i];
      if (!(pen_lo > 0)) { llink[nl] = i; ldir[nl] = 1.0; lpen[nl] = pen_lo; nl++; }
      if (!(pen_hi > 0)) { llink[nl] = i; ldir[nl] = -1.0; lpen[nl] = pen_hi; nl++; }
    }
    if (p->limit_rows_per_leg == 1 && nl > 1) {
      /* kernel-compatible restriction: one row per leg, the most violated joint (first on ties) */
      int keep[2 * MAXL], nk = 0;
      for (int a = 0; a < nl; a++) {
        int ra = llink[a];
        while (m->parent[ra] >= 0) ra = m->parent[ra];
        int best = 1;
        for (int b = 0; b < nl; b++) {
          if (b == a) continue;
          int rb = llink[b];
          while (m->parent[rb] >= 0) rb = m->parent[rb];
          if (rb != ra) continue;
          if (lpen[b] < lpen[a] || (lpen[b] == lpen[a] && b < a)) best = 0;
        }
        if (best) keep[nk++] = a;
      }
      for (int a = 0; a < nk; a++) { llink[a] = llink[keep[a]]; ldir[a] = ldir[keep[a]]; lpen[a] = lpen[keep[a]]; }
      nl = nk;
    }
  }
  e->last_limit_rows = nl;
  /* 2 unconstrained velocity */
  for (int k = 0; k < 3; k++) { e->vang[k] += dt * qdd[k]; e->vlin[k] += dt * qdd[3 + k]; }
  for (int j = 0; j < e->nj; j++) e->qd[e->link_of_dof[j]] += dt * qdd[6 + j];
  clamp_velocities(e);

  if (nc > 0 || nl > 0) {
    double vel[MAXD], dv[MAXD];
    for (int k = 0; k < 3; k++) { vel[k] = e->vang[k]; vel[3 + k] = e->vlin[k]; }
    for (int j = 0; j < e->nj; j++) vel[6 + j] = e->qd[e->link_of_dof[j]];
    for (int k = 0; k < nd; k++) dv[k] = 0;
    /* 3 rows */
    Row rn[MAXC], rf[MAXC][2];
    const v3 nrm = {0, 0, 1};
    v3 t1, t2;
    plane_space(nrm, t1, t2);
    for (int c = 0; c < nc; c++) {
      int l = clink[c];
      const double* dirs[3] = {nrm, t1, t2};
      for (int r = 0; r < 3; r++) {
        Row* row = (r == 0) ? &rn[c] : &rf[c][r - 1];
        contact_jacobian(e, &W, l, cpt[c], dirs[r], row->J);
        impulse_response(e, &W, l, cpt[c], dirs[r], -1, 0.0, row->u);
        double d = rowdot(nd, row->J, row->u);
        row->dinv = 1.0 / d;
        double rel_vel = rowdot(nd, row->J, vel);
        double velocity_error = -rel_vel; /* restitution 0.5 x 0 = 0 */
        double positional_error = 0;
        if (r == 0) {
          double pen = cdist[c] + p->contact_slop;
          if (pen > 0) velocity_error -= pen / dt;
          else positional_error = -pen * p->contact_erp / dt;
        }
        row->rhs = (positional_error + velocity_error) * row->dinv;
        row->lambda = 0; /* [3P] warm starting is disabled for multibody contacts */
      }
    }
    Row rl[2 * MAXL];
    for (int c = 0; c < nl; c++) {
      Row* row = &rl[c];
      const int dof = 6 + e->dof_of_link[llink[c]];
      const v3 zero = {0, 0, 0};
      for (int k = 0; k < nd; k++) row->J[k] = 0;
      row->J[dof] = ldir[c];
      impulse_response(e, &W, -2, zero, zero, llink[c], ldir[c], row->u);
      row->dinv = 1.0 / rowdot(nd, row->J, row->u);
      const double rel_vel = rowdot(nd, row->J, vel);
      /* splitImpulse is on by default: shallow violations combine the ERP push-back with the velocity target,
       * deeper ones put it into m_rhsPenetration, which the multibody solver never applies */
      const double positional_error = (lpen[c] > p->split_impulse_threshold) ? -lpen[c] * p->joint_limit_erp / dt : 0.0;
      row->rhs = (positional_error - rel_vel) * row->dinv;
      row->lambda = 0;
    }
    /* 4 PGS ([3P] btMultiBodyConstraintSolver::solveSingleIteration; the loop of
     * solveGroupCacheFriendlyIterations stops after the iteration whose largest squared row
     * residual, in velocity units deltaImpulse / jacDiagABInv, is <= the threshold) */
    e->last_iters = 0;
    for (int it = 0; it < p->solver_iters; it++) {
      double res2 = 0;
      e->last_iters = it + 1;
      /* non-contact rows come first in every iteration.  limit_rows_per_leg == 1 (what the kernels solve):
       * the rows, one per leg at most, are relaxed as one simultaneous group (all candidates from the same
       * state) -- identical to Bullet's sequential order when an env holds a single limit row; otherwise
       * Bullet's order, row by row */
      {
        double dl[2 * MAXL];
        const int simultaneous = (p->limit_rows_per_leg == 1);
        for (int c = 0; c < nl; c++) {
          Row* r = &rl[c];
          double delta = r->rhs - rowdot(nd, r->J, dv) * r->dinv;
          double sum = r->lambda + delta;
          if (sum < 0) { delta = -r->lambda; r->lambda = 0; }
          else if (sum > p->joint_limit_max_impulse) { delta = p->joint_limit_max_impulse - r->lambda; r->lambda = p->joint_limit_max_impulse; }
          else r->lambda = sum;
          dl[c] = delta;
          if (!simultaneous) for (int k = 0; k < nd; k++) dv[k] += r->u[k] * delta;
          double rv = delta / r->dinv;
          if (rv * rv > res2) res2 = rv * rv;
        }
        if (simultaneous)
          for (int c = 0; c < nl; c++)
            for (int k = 0; k < nd; k++) dv[k] += rl[c].u[k] * dl[c];
      }
      for (int c = 0; c < nc; c++) {
        Row* r = &rn[c];
        double delta = r->rhs - rowdot(nd, r->J, dv) * r->dinv;
        double sum = r->lambda + delta;
        if (sum < 0) { delta = -r->lambda; r->lambda = 0; } else r->lambda = sum;
        for (int k = 0; k < nd; k++) dv[k] += r->u[k] * delta;
        double rv = delta / r->dinv;
        if (rv * rv > res2) res2 = rv * rv;
      }
      for (int c = 0; c < nc; c++) {
        double lim = p->friction * rn[c].lambda;
        Row* a = &rf[c][0];
        Row* b = &rf[c][1];
        if (p->cone_friction) { /* [3P] resolveConeFrictionConstraintRows */
          double dA = a->rhs - rowdot(nd, a->J, dv) * a->dinv;
          double dB = b->rhs - rowdot(nd, b->J, dv) * b->dinv;
          double sumA = a->lambda + dA, sumB = b->lambda + dB;
          double ang = atan2(sumA, sumB);
          double limA = fabs(lim * sin(ang)), limB = fabs(lim * cos(ang));
          sumA = clampd(sumA, -limA, limA);
          sumB = clampd(sumB, -limB, limB);
          dA = sumA - a->lambda; a->lambda = sumA;
          dB = sumB - b->lambda; b->lambda = sumB;
          for (int k = 0; k < nd; k++) dv[k] += a->u[k] * dA + b->u[k] * dB;
          double rv = dA / a->dinv + dB / b->dinv;
          if (rv * rv > res2) res2 = rv * rv;
        } else {
          for (int s = 0; s < 2; s++) {
            Row* r = s ? b : a;
            double delta = r->rhs - rowdot(nd, r->J, dv) * r->dinv;
            double sum = clampd(r->lambda + delta, -lim, lim);
            delta = sum - r->lambda; r->lambda = sum;
            for (int k = 0; k < nd; k++) dv[k] += r->u[k] * delta;
            double rv = delta / r->dinv;
            if (rv * rv > res2) res2 = rv * rv;
          }
        }
      }
      if (res2 <= p->solver_residual_threshold) break;
    }
    /* 5 apply */
    for (int k = 0; k < 3; k++) { e->vang[k] += dv[k]; e->vlin[k] += dv[3 + k]; }
    for (int j = 0; j < e->nj; j++) e->qd[e->link_of_dof[j]] += dv[6 + j];
    clamp_velocities(e);
    for (int c = 0; c < nc; c++) if (cidx[c] < 4) e->c_force[cidx[c]] = rn[c].lambda / dt;
  }
  /* integrate positions ([3P] btMultiBody::stepPositionsMultiDof) */
  for (int k = 0; k < 3; k++) e->pos[k] += dt * e->vlin[k];
  {
    double wn = v3norm(e->vang), ha = 0.5 * wn * dt, s;
    double dq[4];
    if (wn > 1e-12) s = sin(ha) / wn; else s = 0.5 * dt;
    dq[0] = e->vang[0] * s; dq[1] = e->vang[1] * s; dq[2] = e->vang[2] * s; dq[3] = cos(ha);
    double qn[4];
    quat_mul(qn, dq, e->quat);
    double nn = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
    for (int k = 0; k < 4; k++) e->quat[k] = qn[k] / nn;
  }
  for (int i = 0; i < m->num_links; i++) e->q[i] += dt * e->qd[i];
}

/* Independent check hook for the contact pipeline (tests/test_oracle_physics.py): the constraint rows the
 * next oracle_substep(tau) would build, WITHOUT advancing the env.  Row order = the sweep order (limit rows,
 * contact normals, then the friction pair of each contact).  Per row r: J[r][nd] (Jacobian), U[r][nd]
 * (M^-1 J^T by the ABA impulse-response pass), target[r] = positional + velocity error (rhs before the
 * 1/diag scaling), kind[r] = 0 normal / 1,2 friction / 3 joint limit, owner[r] = contact index (limit rows:
 * the link).  vstar[nd] = the unconstrained velocity v + dt*qdd (clamped).  Returns the number of rows.
 * The test recomputes J M^-1 J^T from the dense CRBA mass matrix and solves the same LCP with 10^4 sweeps. */
int oracle_contact_rows(const OracleEnv* e0, const double* tau, double* J, double* U, double* target,
                        int* kind, int* owner, double* vstar) {
  OracleEnv ecopy = *e0;
  OracleEnv* e = &ecopy;
  const SoloModelTable* m = &e->m;
  const SoloSimParams* p = &e->p;
  AbaWork W;
  int nd = 6 + e->nj, nr = 0;
  double qdd[MAXD], dt = p->dt;
  aba_forward(e, &W, tau, qdd);
  int clink[MAXC], cidx[MAXC];
  v3 cpt[MAXC];
  double cdist[MAXC];
  const v3 nrm = {0, 0, 1};
  const int nc = detect_contacts(e, &W, clink, cidx, cdist, cpt);
  for (int k = 0; k < 3; k++) { e->vang[k] += dt * qdd[k]; e->vlin[k] += dt * qdd[3 + k]; }
  for (int j = 0; j < e->nj; j++) e->qd[e->link_of_dof[j]] += dt * qdd[6 + j];
  clamp_velocities(e);
  double vel[MAXD];
  for (int k = 0; k < 3; k++) { vel[k] = e->vang[k]; vel[3 + k] = e->vlin[k]; }
  for (int j = 0; j < e->nj; j++) vel[6 + j] = e->qd[e->link_of_dof[j]];
  for (int k = 0; k < nd; k++) vstar[k] = vel[k];
  if (p->joint_limits) {
    const double lim = p->joint_state_limit;
    for (int i = 0; i < m->num_links; i++) {
      if (m->jtype[i] != SOLO_JOINT_REVOLUTE) continue;
      for (int side = 0; side < 2; side++) {
        const double pen = side == 0 ? e0->q[i] + lim : lim - e0->q[i];
        const double dir = side == 0 ? 1.0 : -1.0;
        if (pen > 0) continue;
        double* Jr = J + (size_t)nr * nd; double* Ur = U + (size_t)nr * nd;
        const v3 zero = {0, 0, 0};
        for (int k = 0; k < nd; k++) Jr[k] = 0;
        Jr[6 + e->dof_of_link[i]] = dir;
        impulse_response(e, &W, -2, zero, zero, i, dir, Ur);
        const double positional = (pen > p->split_impulse_threshold) ? -pen * p->joint_limit_erp / dt : 0.0;
        target[nr] = positional - rowdot(nd, Jr, vel);
        kind[nr] = 3; owner[nr] = i; nr++;
      }
    }
  }
  v3 t1, t2;
  plane_space(nrm, t1, t2);
  const double* dirs[3] = {nrm, t1, t2};
  for (int pass = 0; pass < 2; pass++) {          /* normals of every contact first, then the friction pairs */
    for (int c = 0; c < nc; c++) {
      int l = clink[c];
      for (int r = (pass == 0 ? 0 : 1); r < (pass == 0 ? 1 : 3); r++) {
        double* Jr = J + (size_t)nr * nd; double* Ur = U + (size_t)nr * nd;
        contact_jacobian(e, &W, l, cpt[c], dirs[r], Jr);
        impulse_response(e, &W, l, cpt[c], dirs[r], -1, 0.0, Ur);
        double verr = -rowdot(nd, Jr, vel), perr = 0;
        if (r == 0) {
          double pen = cdist[c] + p->contact_slop;
          if (pen > 0) verr -= pen / dt; else perr = -pen * p->contact_erp / dt;
        }
        target[nr] = perr + verr;
        kind[nr] = r; owner[nr] = c; nr++;
      }
    }
  }
  return nr;
}

void oracle_set_external_force(OracleEnv* e, const double* f3) { for (int k = 0; k < 3; k++) e->fext[k] = f3[k]; }

/* parity hook: overwrite the contact record (what p.getContactPoints returned, solo.py:313-317):
 * force[f] < 0 = no contact point on foot f */
void oracle_set_contacts(OracleEnv* e, const double* force) {
  for (int f = 0; f < e->m.num_feet; f++) {
    e->c_has[f] = force[f] >= 0;
    e->c_force[f] = force[f] >= 0 ? force[f] : 0.0;
  }
}

void oracle_forward_dynamics(OracleEnv* e, const double* tau, double* qdd) {
  AbaWork W;
  aba_forward(e, &W, tau, qdd);
}

/* Independent cross-check: composite mass matrix from world-frame body Jacobians
 * (M = sum_b J_b^T I_b J_b), bias by a world-frame recursive Newton-Euler pass, dense
 * Cholesky.  Spatial vectors here are about the WORLD origin in world axes, i.e. a
 * different formulation from aba_forward's link-COM frames. */
void oracle_forward_dynamics_crba(OracleEnv* e, const double* tau, double* qdd, double* Mout) {
  const SoloModelTable* m = &e->m;
  const SoloSimParams* p = &e->p;
  int n = m->num_links, nd = 6 + e->nj;
  AbaWork W;
  build_kinematics(e, &W);
  m3 Rb; quat_to_m3(Rb, e->quat);
  /* body list: index 0 = base, 1+i = link i */
  v6 Sw[MAXL], vb[MAXL + 1], ab[MAXL + 1];
  /* base spatial velocity about the world origin */
  {
    v3 t; v3cross(t, e->pos, e->vang); /* v_O = v + p x w */
    for (int k = 0; k < 3; k++) { vb[0][k] = e->vang[k]; vb[0][3 + k] = e->vlin[k] + t[k]; }
    v6zero(ab[0]);
  }
  for (int i = 0; i < n; i++) {
    int par = m->parent[i];
    v6zero(Sw[i]);
    if (m->jtype[i] == SOLO_JOINT_REVOLUTE) {
      v3 aw, comw, ow, t;
      m3mulv(aw, W.Rw[i], m->axis[i]);
      m3mulv(comw, W.Rw[i], m->com[i]);
      for (int k = 0; k < 3; k++) ow[k] = W.pw[i][k] - comw[k];
      v3cross(t, ow, aw);
      for (int k = 0; k < 3; k++) { Sw[i][k] = aw[k]; Sw[i][3 + k] = t[k]; }
    }
    v6 vj, cj;
    for (int k = 0; k < 6; k++) vj[k] = Sw[i][k] * e->qd[i];
    for (int k = 0; k < 6; k++) vb[1 + i][k] = vb[1 + par][k] + vj[k];
    v6crossm(cj, vb[1 + i], vj);
    for (int k = 0; k < 6; k++) ab[1 + i][k] = ab[1 + par][k] + cj[k];
  }
  double M[MAXD * MAXD], rhs[MAXD];
  for (int k = 0; k < nd * nd; k++) M[k] = 0;
  for (int k = 0; k < 6; k++) rhs[k] = 0;
  for (int j = 0; j < e->nj; j++) rhs[6 + j] = tau[j];
  for (int b = 0; b <= n; b++) {
    double mass = b == 0 ? m->base_mass : m->mass[b - 1];
    m3 Il, Iw, R, T;
    v3 c;
    if (b == 0) { m3sym6(Il, m->base_inertia); memcpy(R, Rb, sizeof(m3)); m3mulv(c, Rb, m->base_com); for (int k = 0; k < 3; k++) c[k] += e->pos[k]; }
    else { m3sym6(Il, m->inertia[b - 1]); memcpy(R, W.Rw[b - 1], sizeof(m3)); v3cpy(c, W.pw[b - 1]); }
    m3mul(T, R, Il); m3 Rt; m3T(Rt, R); m3mul(Iw, T, Rt);
    /* spatial inertia about the world origin */
    m6 I6; m6zero(I6);
    m3 cx, cxcx; m3skew(cx, c); m3mul(cxcx, cx, cx);
    for (int a = 0; a < 3; a++)
      for (int d = 0; d < 3; d++) {
        I6[a][d] = Iw[a][d] - mass * cxcx[a][d];
        I6[a][3 + d] = mass * cx[a][d];
        I6[3 + a][d] = -mass * cx[a][d];
      }
    for (int a = 0; a < 3; a++) I6[3 + a][3 + a] = mass;
    /* body jacobian columns: base 6 (identity), joints on the path */
    double Jb[6][MAXD];
    for (int a = 0; a < 6; a++) for (int k = 0; k < nd; k++) Jb[a][k] = (k == a) ? 1.0 : 0.0;
    for (int i = b - 1; i >= 0; i = m->parent[i])
      if (m->jtype[i] == SOLO_JOINT_REVOLUTE)
        for (int a = 0; a < 6; a++) Jb[a][6 + e->dof_of_link[i]] = Sw[i][a];
    /* M += J^T I J */
    double IJ[6][MAXD];
    for (int a = 0; a < 6; a++)
      for (int k = 0; k < nd; k++) { double s = 0; for (int d = 0; d < 6; d++) s += I6[a][d] * Jb[d][k]; IJ[a][k] = s; }
    for (int k = 0; k < nd; k++)
      for (int l = 0; l < nd; l++) { double s = 0; for (int a = 0; a < 6; a++) s += Jb[a][k] * IJ[a][l]; M[k * nd + l] += s; }
    /* bias wrench: I a0 + v x* I v - f_ext */
    v6 Iv, Ia, f;
    m6mulv(Iv, I6, vb[b]);
    m6mulv(Ia, I6, ab[b]);
    {
      v3 t1, t2, t3;
      v3cross(t1, vb[b], Iv); v3cross(t2, vb[b] + 3, Iv + 3); v3cross(t3, vb[b], Iv + 3);
      for (int k = 0; k < 3; k++) { f[k] = Ia[k] + t1[k] + t2[k]; f[3 + k] = Ia[3 + k] + t3[k]; }
    }
    /* external: gravity + Bullet damping at the COM */
    {
      v3 w, vc, t, fe, te, Iww;
      v3cpy(w, vb[b]);
      v3cross(t, w, c);
      for (int k = 0; k < 3; k++) vc[k] = vb[b][3 + k] + t[k];
      double vn = v3norm(vc), wn = v3norm(w);
      m3mulv(Iww, Iw, w);
      for (int k = 0; k < 3; k++) {
        fe[k] = -mass * vc[k] * (p->lin_damping + p->lin_damping * vn);
        te[k] = -Iww[k] * (p->ang_damping + p->ang_damping * wn);
      }
      fe[2] += mass * p->gravity_z;
      if (b == 0) {                         /* external push on the base: base axes -> world, at the base origin */
        v3 fw; m3mulv(fw, Rb, e->fext);
        for (int k = 0; k < 3; k++) fe[k] += fw[k];
      }
      v3cross(t, c, fe);
      for (int k = 0; k < 3; k++) { f[k] -= te[k] + t[k]; f[3 + k] -= fe[k]; }
    }
    for (int k = 0; k < nd; k++) { double s = 0; for (int a = 0; a < 6; a++) s += Jb[a][k] * f[a]; rhs[k] -= s; }
  }
  if (Mout) memcpy(Mout, M, sizeof(double) * nd * nd);
  chol_solve(nd, M, nd, rhs);
  /* spatial base acceleration about world origin -> classical acceleration of the base origin:
   * v = v_O + w x p  =>  vdot = vdot_O + wdot x p + w x v */
  {
    v3 t1, t2;
    v3cross(t1, rhs, e->pos);
    v3cross(t2, e->vang, e->vlin);
    for (int k = 0; k < 3; k++) { qdd[k] = rhs[k]; qdd[3 + k] = rhs[3 + k] + t1[k] + t2[k]; }
    for (int j = 0; j < e->nj; j++) qdd[6 + j] = rhs[6 + j];
  }
}

double oracle_energy(OracleEnv* e) {
  const SoloModelTable* m = &e->m;
  AbaWork W;
  build_kinematics(e, &W);
  m3 Rb; quat_to_m3(Rb, e->quat);
  double E = 0;
  /* base */
  {
    v3 wl, Iw; m3 Ib; m3sym6(Ib, m->base_inertia);
    m3Tmulv(wl, Rb, e->vang); m3mulv(Iw, Ib, wl);
    E += 0.5 * m->base_mass * v3dot(e->vlin, e->vlin) + 0.5 * v3dot(wl, Iw) - m->base_mass * e->p.gravity_z * e->pos[2];
  }
  v6 v0;
  m3Tmulv(v0, Rb, e->vang); m3Tmulv(v0 + 3, Rb, e->vlin);
  v6 vl[MAXL];
  for (int i = 0; i < m->num_links; i++) {
    int par = m->parent[i];
    m6mulv(vl[i], W.X[i], par < 0 ? v0 : vl[par]);
    for (int k = 0; k < 6; k++) vl[i][k] += W.S[i][k] * e->qd[i];
    v3 Iw; m3 Il; m3sym6(Il, m->inertia[i]); m3mulv(Iw, Il, vl[i]);
    E += 0.5 * m->mass[i] * v3dot(vl[i] + 3, vl[i] + 3) + 0.5 * v3dot(vl[i], Iw) - m->mass[i] * e->p.gravity_z * W.pw[i][2];
  }
  return E;
}

void oracle_foot_positions(OracleEnv* e, double* out) {
  const SoloModelTable* m = &e->m;
  AbaWork W;
  build_kinematics(e, &W);
  for (int f = 0; f < m->num_feet; f++) {
    int l = m->foot_link[f];
    v3 off, c;
    for (int k = 0; k < 3; k++) off[k] = m->foot_center[f][k] - m->com[l][k];
    m3mulv(c, W.Rw[l], off);
    for (int k = 0; k < 3; k++) out[f * 3 + k] = c[k] + W.pw[l][k];
  }
}

/* ------------------------------------------------------------------ env layer */
void oracle_get_state(const OracleEnv* e, double* s) {
  for (int k = 0; k < 3; k++) { s[k] = e->pos[k]; s[7 + k] = e->vlin[k]; s[10 + k] = e->vang[k]; }
  for (int k = 0; k < 4; k++) s[3 + k] = e->quat[k];
  for (int j = 0; j < e->nj; j++) { s[13 + j] = e->q[e->link_of_dof[j]]; s[13 + e->nj + j] = e->qd[e->link_of_dof[j]]; }
}

/* solo.py:310-323: flag = 1 iff some ground-foot contact point has tuple[9] (normal
 * force [3P]) < 0.2 */
static void contact_flags(const OracleEnv* e, double* f) {
  for (int i = 0; i < e->m.num_feet; i++)
    f[i] = (e->c_has[i] && e->c_force[i] < e->p.contact_flag_force) ? 1.0 : 0.0;
}

void oracle_get_contacts(const OracleEnv* e, double* out) {
  double f[SOLO_MAX_FEET];
  contact_flags(e, f);
  for (int i = 0; i < e->m.num_feet; i++) { out[i * 3] = f[i]; out[i * 3 + 1] = e->c_has[i]; out[i * 3 + 2] = e->c_force[i]; }
}

/* solo.py:198-222 */
void oracle_get_current_state(OracleEnv* e, double* s) {
  int nj = e->nj, k = 0;
  s[k++] = e->pos[2];                                          /* :202 */
  v3 eul; quat_to_euler(eul, e->quat);                         /* :203 */
  for (int i = 0; i < 3; i++) s[k++] = (pymod(eul[i], 2.0) * M_PI) / (2 * M_PI); /* :206, SURVEY F6 */
  for (int i = 0; i < 3; i++) s[k++] = e->vlin[i];             /* :204 getBaseVelocity = (lin, ang) world */
  for (int i = 0; i < 3; i++) s[k++] = e->vang[i];
  for (int j = 0; j < nj; j++) s[k++] = e->q[e->link_of_dof[j]] / e->p.joint_state_limit;  /* :210 */
  for (int j = 0; j < nj; j++) s[k++] = e->qd[e->link_of_dof[j]] / e->p.joint_vel_limit;   /* :211 */
  contact_flags(e, s + k); k += 4;                             /* :215 */
  if (e->p.task == SOLO_TASK_POINTGOAL) {                      /* :218-220, :337-340 */
    s[k++] = e->pos[0] / 2.0; s[k++] = e->pos[1] / 2.0;
    s[k++] = e->goal[0] / 2.0; s[k++] = e->goal[1] / 2.0;
  }
}

/* solo.py:186-196 */
void oracle_get_observation(OracleEnv* e, double* obs) {
  double cur[ORACLE_MAX_D0];
  oracle_get_current_state(e, cur);
  for (int k = 0; k < e->d0; k++) obs[k] = cur[k];
  for (int h = 0; h < e->p.num_history_stack; h++) /* reversed(deque): newest first */
    for (int k = 0; k < e->d0; k++) obs[(1 + h) * e->d0 + k] = cur[k] - e->hist[h][k];
}

static void history_push(OracleEnv* e) { /* deque(maxlen=H).append, solo.py:262 */
  int H = e->p.num_history_stack;
  if (H <= 0) return;
  for (int h = H - 1; h > 0; h--) memcpy(e->hist[h], e->hist[h - 1], sizeof(double) * e->d0);
  oracle_get_current_state(e, e->hist[0]);
}

static double calc_potential(const OracleEnv* e) { /* solo.py:277-279 */
  double dx = e->pos[0] - e->goal[0], dy = e->pos[1] - e->goal[1];
  return sqrt(dx * dx + dy * dy);
}

/* solo.py:261-274; torque acts during the first substep only ([3P] applied joint torques
 * are cleared after every stepSimulation) unless torque_hold */
static void simulator_step(OracleEnv* e, const double* tau) {
  double zero[MAXL] = {0};
  history_push(e);
  for (int s = 0; s < e->p.frame_skip; s++)
    oracle_substep(e, (tau && (s == 0 || e->p.torque_hold)) ? tau : zero);
  if (e->p.task == SOLO_TASK_POINTGOAL) {
    double oldp = e->potential;              /* solo.py:281-284 */
    e->potential = calc_potential(e);
    e->progress = -1.0 * (e->potential - oldp);
    if (e->potential < e->p.goal_reach_dist) { /* :270-272 */
      e->goals_reached += 1;
      uint32_t w[4]; env_rng(e, w); sample_goal_from(e, w);
    }
  }
}

void oracle_set_goal(OracleEnv* e, double gx, double gy) {
  e->goal[0] = gx; e->goal[1] = gy; e->potential = calc_potential(e); e->progress = 0;
}
void oracle_get_goal(const OracleEnv* e, double* g) { g[0] = e->goal[0]; g[1] = e->goal[1]; }
void oracle_set_goal_radius(OracleEnv* e, double r) { e->goal_radius = r; }

static void clear_episode(OracleEnv* e) { /* baseEnv.py:72-77 */
  e->need_reset = 0; e->timestep = 0; e->reward_sum = 0; e->goals_reached_env = 0;
  for (int k = 0; k < 5; k++) e->dr[k] = 0;
}

void oracle_set_state(OracleEnv* e, const double* s) {
  for (int k = 0; k < 3; k++) { e->pos[k] = s[k]; e->vlin[k] = s[7 + k]; e->vang[k] = s[10 + k]; }
  {
    double nn = sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
    for (int k = 0; k < 4; k++) e->quat[k] = s[3 + k] / nn;   /* orientation is a unit quaternion */
  }
  for (int i = 0; i < e->m.num_links; i++) { e->q[i] = 0; e->qd[i] = 0; }
  for (int j = 0; j < e->nj; j++) { e->q[e->link_of_dof[j]] = s[13 + j]; e->qd[e->link_of_dof[j]] = s[13 + e->nj + j]; }
  for (int f = 0; f < SOLO_MAX_FEET; f++) { e->c_has[f] = 0; e->c_force[f] = 0; }
  for (int h = 0; h < e->p.num_history_stack; h++) oracle_get_current_state(e, e->hist[h]);
  if (e->p.task == SOLO_TASK_POINTGOAL) { e->potential = calc_potential(e); e->progress = 0; }
  clear_episode(e);
}

/* baseEnv.py:70-82 + solo.py:166-181,291-296 */
void oracle_env_reset(OracleEnv* e, double* obs) {
  /* robot_specific_reset: base -> (0,0,initial_z), identity, q = 0; [3P] velocities zeroed */
  v3set(e->pos, 0, 0, e->p.initial_z);
  e->quat[0] = e->quat[1] = e->quat[2] = 0; e->quat[3] = 1;
  v3set(e->vlin, 0, 0, 0); v3set(e->vang, 0, 0, 0);
  for (int i = 0; i < e->m.num_links; i++) { e->q[i] = 0; e->qd[i] = 0; }
  for (int f = 0; f < SOLO_MAX_FEET; f++) { e->c_has[f] = 0; e->c_force[f] = 0; } /* contact set cleared */
  e->episode += 1; e->draw = 0;
  uint32_t w[4]; env_rng(e, w);
  if (e->p.task == SOLO_TASK_POINTGOAL) sample_goal_from(e, w); /* needed before get_current_state */
  for (int h = 0; h < e->p.num_history_stack; h++) oracle_get_current_state(e, e->hist[h]); /* solo.py:170-171 */
  if (e->p.task == SOLO_TASK_POINTGOAL) {                                                     /* :173-177 */
    e->goals_reached = 0; e->potential = calc_potential(e); e->progress = 0;
  }
  clear_episode(e);
  int span = e->p.settle_max - e->p.settle_min;
  int k = e->p.settle_min + (span > 0 ? (int)(w[0] % (uint32_t)span) : 0); /* baseEnv.py:79 */
  e->settle_last = k;
  for (int i = 0; i < k; i++) simulator_step(e, NULL);
  if (obs) oracle_get_observation(e, obs);
}

/* solo.py:224-259, controllers/PD.py:3-10 */
void oracle_action_to_torque(const OracleEnv* e, const double* a, double* tau) {
  const SoloSimParams* p = &e->p;
  for (int j = 0; j < e->nj; j++) {
    int l = e->link_of_dof[j];
    if (p->control == SOLO_CONTROL_TORQUE) {
      tau[j] = clampd(a[j], -1, 1) * p->max_torque;                       /* :229 */
    } else {
      double kp = p->kp, kd = p->kd;
      if (p->control == SOLO_CONTROL_VPD) { kp = a[e->nj]; kd = a[e->nj + 1]; } /* :250 */
      double q_ref = clampd(a[j], -1, 1) * p->joint_state_limit;          /* :234,:246 */
      double t = kp * (q_ref - e->q[l]) - kd * e->qd[l];                   /* PD.py:5 */
      tau[j] = clampd(t, -p->max_torque, p->max_torque);                  /* PD.py:8 */
    }
  }
}

/* baseEnv.py:91-157 */
static double get_reward(OracleEnv* e, const double* action) {
  const SoloSimParams* p = &e->p;
  int nj = e->nj;
  double stand = 0, jp = 0, balance = 0, progress = 0, torque = 0;
  double pos_z = e->pos[2];
  stand = (pos_z > p->stand_z ? 1.0 : 0.0) * 0.5;                          /* :96,:109,:127 */
  double s = 0;
  for (int j = 0; j < nj; j++) {
    double q = e->q[e->link_of_dof[j]];
    s += (p->task == SOLO_TASK_STAND) ? fabs(q) : q * q;                    /* :101 | :113,:131 */
  }
  jp = -0.1 * (s / nj);
  if (p->task == SOLO_TASK_WALK) {                                          /* :115-119 */
    if (pos_z > p->stand_z) {
      double vx = e->vlin[0];
      double sg = (vx > 0) - (vx < 0);
      progress = 2 * sg * vx * vx;
    }
  } else if (p->task == SOLO_TASK_POINTGOAL) {                              /* :133-140 */
    v3 eul; quat_to_euler(eul, e->quat);
    balance = -0.1 * (fabs(eul[0]) + fabs(eul[1]));
    if (pos_z > p->stand_z) progress = 1 * e->progress * (1.0 / p->pointgoal_dt);
  }
  if (p->control == SOLO_CONTROL_TORQUE) {                                  /* :142-146, SURVEY F9a */
    double tp = 0;
    for (int j = 0; j < e->act_dim; j++) tp += action[j] * action[j];
    torque = -0.01 * tp;
  }
  e->dr[0] += stand; e->dr[1] += jp; e->dr[2] += torque; e->dr[3] += balance; e->dr[4] += progress; /* :182-187 */
  return stand + jp + balance + progress + torque;
}

int oracle_env_step(OracleEnv* e, const double* action, int auto_reset, double* obs,
                    double* reward, int* done, OracleInfo* info) {
  const SoloSimParams* p = &e->p;
  if (e->need_reset) return -1;                                /* baseEnv.py:43 */
  double tau[MAXL];
  oracle_action_to_torque(e, action, tau);                     /* :45 */
  simulator_step(e, tau);                                      /* :46 */
  e->timestep += 1;                                            /* :47 */
  if (obs) oracle_get_observation(e, obs);                     /* :49 */
  double r = get_reward(e, action);                            /* :50 */
  int d = 0, success = 0, timeout = 0;                         /* :162-180 */
  if (e->timestep >= p->episode_length) { d = 1; timeout = 1; success = (p->task != SOLO_TASK_POINTGOAL); }
  else if (e->pos[2] < p->fall_z) { d = 1; }
  else if (p->task == SOLO_TASK_POINTGOAL && e->goals_reached > e->goals_reached_env) {
    e->goals_reached_env = e->goals_reached; d = 1; success = 1;
  }
  if (d) {                                                     /* :52-60 */
    e->need_reset = 1;
    if (success) { if (p->task == SOLO_TASK_POINTGOAL) r = 0.1 * (p->episode_length - e->timestep); }
    else if (!timeout) r = -10;
  }
  e->reward_sum += r;                                          /* :62 */
  if (info) {                                                  /* :63-66 */
    info->episode_reward = r; info->episode_return = e->reward_sum;
    info->episode_length = e->timestep; info->success = success; info->timeout = timeout;
    info->goals_reached = e->goals_reached_env;
    info->dr_stand = e->dr[0]; info->dr_joint_pose = e->dr[1]; info->dr_torque = e->dr[2];
    info->dr_balance = e->dr[3]; info->dr_progress = e->dr[4];
  }
  *reward = r; *done = d;
  if (d && auto_reset) oracle_env_reset(e, obs);               /* agents/ppo/envs.py:38-40 */
  return 0;
}

/* agents/ppo/storage.py:35-55, float32, same op order */
void oracle_gae(const float* rewards, const float* values, const float* masks, float* returns,
                int T, int N, float gamma, float lam, int use_gae) {
  for (int n = 0; n < N; n++) {
    if (use_gae) {
      float gae = 0.f;
      for (int t = T - 1; t >= 0; t--) {
        float delta = rewards[t * N + n] + gamma * values[(t + 1) * N + n] * masks[(t + 1) * N + n] - values[t * N + n];
        gae = delta + gamma * lam * masks[(t + 1) * N + n] * gae;
        returns[t * N + n] = gae + values[t * N + n];
      }
    } else {
      for (int t = T - 1; t >= 0; t--)
        returns[t * N + n] = returns[(t + 1) * N + n] * gamma * masks[(t + 1) * N + n] + rewards[t * N + n];
    }
  }
}

/* ------------------------------------------------------------------ batched (CPU baseline) */
/* A persistent pthread pool (no OpenMP): bench.py runs under torchrun at N > 1, which exports
 * OMP_NUM_THREADS=1 to every rank; libgomp then keeps a one-thread pool and builds / tears down the other
 * team threads in EVERY `parallel num_threads(n)` region, which cut the CPU arm to a third of its
 * stand-alone rate in round 1.  Workers pull chunks of envs from an atomic counter (dynamic schedule: env
 * steps differ in cost by the number of contacts, PGS sweeps and resets) and sleep on a condition variable
 * between batches. */
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct {
  OracleEnv** envs; int n; const float* actions; float* obs; float* reward; float* done; int mode;
  const double* states; const double* taus; double* out_states; double* out_contacts; int* out_iters;  /* mode 2 */
} BatchJob;

static struct {
  pthread_t th[256];
  int nth, started;
  pthread_mutex_t mu;
  pthread_cond_t go, fin;
  unsigned long gen;
  int running;
  atomic_int next;
  BatchJob job;
} g_pool = {.mu = PTHREAD_MUTEX_INITIALIZER, .go = PTHREAD_COND_INITIALIZER, .fin = PTHREAD_COND_INITIALIZER};

static void batch_run_chunks(const BatchJob* j) {
  const int chunk = 4;
  for (;;) {
    int i0 = atomic_fetch_add(&g_pool.next, chunk);
    if (i0 >= j->n) break;
    int i1 = i0 + chunk < j->n ? i0 + chunk : j->n;
    for (int i = i0; i < i1; i++) {
      OracleEnv* e = j->envs[i];
      double o[ORACLE_MAX_D0 * (1 + ORACLE_MAX_HIST)];
      if (j->mode == 2) {            /* parity tests: set_state -> one substep -> get_state / contacts */
        const int W = 13 + 2 * e->nj;
        oracle_set_state(e, j->states + (size_t)i * W);
        oracle_substep(e, j->taus + (size_t)i * e->nj);
        oracle_get_state(e, j->out_states + (size_t)i * W);
        if (j->out_contacts) oracle_get_contacts(e, j->out_contacts + (size_t)i * 12);
        if (j->out_iters) j->out_iters[i] = e->last_iters;
      } else if (j->mode == 0) {
        oracle_env_reset(e, o);
        if (j->obs) for (int k = 0; k < e->d; k++) j->obs[(size_t)i * e->d + k] = (float)o[k];
      } else {
        double a[MAXL + 2], r;
        int d;
        for (int k = 0; k < e->act_dim; k++) a[k] = j->actions[(size_t)i * e->act_dim + k];
        oracle_env_step(e, a, 1, o, &r, &d, NULL);
        if (j->obs) for (int k = 0; k < e->d; k++) j->obs[(size_t)i * e->d + k] = (float)o[k];
        j->reward[i] = (float)r; j->done[i] = (float)d;
      }
    }
  }
}

static void* pool_worker(void* arg) {
  (void)arg;
  unsigned long seen = 0;
  for (;;) {
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.gen == seen) pthread_cond_wait(&g_pool.go, &g_pool.mu);
    seen = g_pool.gen;
    BatchJob job = g_pool.job;
    pthread_mutex_unlock(&g_pool.mu);
    batch_run_chunks(&job);
    pthread_mutex_lock(&g_pool.mu);
    if (--g_pool.running == 0) pthread_cond_signal(&g_pool.fin);
    pthread_mutex_unlock(&g_pool.mu);
  }
  return NULL;
}

static void pool_run(const BatchJob* job, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_mutex_lock(&g_pool.mu);
  while (g_pool.nth < nthreads - 1) {          /* the caller is the n-th worker */
    if (pthread_create(&g_pool.th[g_pool.nth], NULL, pool_worker, NULL) != 0) break;
    g_pool.nth++;
  }
  const int helpers = g_pool.nth < nthreads - 1 ? g_pool.nth : nthreads - 1;
  g_pool.job = *job;
  atomic_store(&g_pool.next, 0);
  /* every pooled thread wakes on the broadcast; more threads than asked for only happens when an earlier
   * call asked for more, and they all pull from the same counter, so the result does not depend on it */
  g_pool.running = g_pool.nth;
  g_pool.gen++;
  pthread_cond_broadcast(&g_pool.go);
  pthread_mutex_unlock(&g_pool.mu);
  (void)helpers;
  batch_run_chunks(job);
  pthread_mutex_lock(&g_pool.mu);
  while (g_pool.running > 0) pthread_cond_wait(&g_pool.fin, &g_pool.mu);
  pthread_mutex_unlock(&g_pool.mu);
}

int oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

void oracle_batch_substep(OracleEnv** envs, int n, const double* states, const double* taus, double* out_states,
                          double* out_contacts, int* out_iters, int nthreads) {
  BatchJob j = {envs, n, NULL, NULL, NULL, NULL, 2, states, taus, out_states, out_contacts, out_iters};
  pool_run(&j, nthreads);
}

void oracle_batch_reset(OracleEnv** envs, int n, float* obs, int nthreads) {
  BatchJob j = {envs, n, NULL, obs, NULL, NULL, 0, NULL, NULL, NULL, NULL, NULL};
  pool_run(&j, nthreads);
}

void oracle_batch_step(OracleEnv** envs, int n, const float* actions, float* obs, float* reward,
                       float* done, int nthreads) {
  BatchJob j = {envs, n, actions, obs, reward, done, 1, NULL, NULL, NULL, NULL, NULL};
  pool_run(&j, nthreads);
}
