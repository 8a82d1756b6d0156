/*
 * solo_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Double-precision, single-env, clarity-over-speed restatement of the reference's
 * env-step path:  baseEnv.py:42-187, solo.py:186-340, controllers/PD.py:3-10,
 * simulation.py:13-35, plus the third-party PyBullet `stepSimulation` the reference
 * calls at solo.py:66,265.
 *
 * PARITY UNPINNED for the physics: PyBullet (pip `pybullet`, version unpinned by the
 * reference: no requirements file, README.md:4-8) is not installable in this image and
 * the reference ships no tests / golden vectors.  The Bullet step is restated from its
 * published algorithm (btMultiBody ABA with 0.04 linear/angular damping and gyroscopic
 * term, speculative contact rows with ERP, btMultiBodyConstraintSolver sequential-
 * impulse PGS, 50 iterations, implicit cone friction, semi-implicit Euler).  The env
 * arithmetic (observation, reward, termination, reset, PD, GAE) IS pinned: against
 * closed-form known answers from the reference text and against golden vectors
 * produced by importing the reference's own PD.py / storage.py (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use this library.
 */
#ifndef SOLO_ORACLE_H
#define SOLO_ORACLE_H

#include "../include/solo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_HIST 8
#define ORACLE_MAX_D0 64
#define ORACLE_MAX_DOF (6 + SOLO_MAX_LINKS)

typedef struct OracleInfo {
  double episode_reward, episode_return;
  int episode_length, success, timeout, goals_reached;
  double dr_stand, dr_joint_pose, dr_torque, dr_balance, dr_progress;
} OracleInfo;

typedef struct OracleEnv OracleEnv;

void oracle_default_params(SoloSimParams* p);
OracleEnv* oracle_env_create(const SoloModelTable* m, const SoloSimParams* p,
                             uint64_t seed, int64_t env_id);
void oracle_env_destroy(OracleEnv* e);
int oracle_nj(const OracleEnv* e);
int oracle_act_dim(const OracleEnv* e);
int oracle_obs_dim0(const OracleEnv* e);
int oracle_obs_dim(const OracleEnv* e);

/* baseEnv.py:70-82 */
void oracle_env_reset(OracleEnv* e, double* obs);
/* baseEnv.py:42-68 + worker auto-reset agents/ppo/envs.py:38-40 (auto_reset != 0).
 * returns 0, or -1 when called before reset (baseEnv.py:43). */
int oracle_env_step(OracleEnv* e, const double* action, int auto_reset, double* obs,
                    double* reward, int* done, OracleInfo* info);
void oracle_get_observation(OracleEnv* e, double* obs);   /* solo.py:186-196 */
void oracle_get_current_state(OracleEnv* e, double* s0);  /* solo.py:198-222 */

/* state = pos(3) quat xyzw(4) linvel(3) angvel(3) q(nj) qd(nj) */
void oracle_get_state(const OracleEnv* e, double* state);
/* also clears the contact set, refills history with the current state (solo.py:170-171)
 * and marks the env as reset */
void oracle_set_state(OracleEnv* e, const double* state);
void oracle_set_goal(OracleEnv* e, double gx, double gy);
void oracle_get_goal(const OracleEnv* e, double* gxy);
void oracle_set_goal_radius(OracleEnv* e, double r);
int oracle_settle_count_last(const OracleEnv* e);
int oracle_last_solver_iters(const OracleEnv* e);
int oracle_last_limit_rows(const OracleEnv* e);

/* solo.py:224-259 + PD.py:3-10 on the env's current state; tau[nj] */
void oracle_action_to_torque(const OracleEnv* e, const double* action, double* tau);
/* contact-free ABA on the env's current state: qdd[6+nj] = (angacc world, linacc world, joints) */
void oracle_forward_dynamics(OracleEnv* e, const double* tau, double* qdd);
/* same accelerations via CRBA + RNEA + dense solve (independent cross-check); also
 * returns the joint-space mass matrix M[(6+nj)^2] when M != NULL */
void oracle_forward_dynamics_crba(OracleEnv* e, const double* tau, double* qdd, double* M);
/* one Bullet-equivalent step (p.stepSimulation, solo.py:265) */
void oracle_substep(OracleEnv* e, const double* tau);
/* contact record of the last substep: out[4][3] = flag, has_point, normal force */
void oracle_get_contacts(const OracleEnv* e, double* out);
void oracle_set_contacts(OracleEnv* e, const double* force4);
/* external force on the base (base axes, at the base origin) for the following substeps; gait-env pushes */
void oracle_set_external_force(OracleEnv* e, const double* f3);
/* constraint rows of the next substep without advancing the env (see solo_oracle.c); J, U: [rows][6+nj] */
int oracle_contact_rows(const OracleEnv* e, const double* tau, double* J, double* U, double* target,
                        int* kind, int* owner, double* vstar);
/* total mechanical energy (kinetic + potential) of the current state */
double oracle_energy(OracleEnv* e);
/* world position of foot sphere centres: out[4][3] */
void oracle_foot_positions(OracleEnv* e, double* out);

/* float32 GAE exactly in the op order of agents/ppo/storage.py:35-55 */
void oracle_gae(const float* rewards, const float* values, const float* masks, float* returns,
                int T, int N, float gamma, float lam, int use_gae);

/* Batched stepping over an array of envs with OpenMP (CPU baseline leg of bench.py):
 * actions [n, A] double, obs [n, D] float, reward [n] float, done [n] float. */
void oracle_batch_reset(OracleEnv** envs, int n, float* obs, int nthreads);
void oracle_batch_step(OracleEnv** envs, int n, const float* actions, float* obs,
                       float* reward, float* done, int nthreads);
/* parity tests at scale: env i <- states[i]; one substep with taus[i]; out_states[i], out_contacts[i][4][3]
 * (flag, has_point, force), out_iters[i] = PGS sweeps used (any output may be NULL except out_states) */
void oracle_batch_substep(OracleEnv** envs, int n, const double* states, const double* taus, double* out_states,
                          double* out_contacts, int* out_iters, int nthreads);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
