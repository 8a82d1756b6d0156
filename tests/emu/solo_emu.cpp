/*
 * solo_emu.cpp — TEST HARNESS: replays the four-lanes-per-env program of the CUDA step
 * kernel (solorl_b200/csrc/solo_core.cuh + solo_env.cuh, compiled for the host) on the
 * CPU in fp32, lanes executed one after another and shuffles replaced by array reads.
 * It lets the `-m "not gpu"` suite check the kernel math against the fp64 oracle in a
 * container without a GPU.  It is never built into, or imported by, the product.
 */
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/solo_b200.h"
#include "../../solorl_b200/csrc/solo_core.cuh"
#include "../../solorl_b200/csrc/solo_body.cuh"
#include "../../solorl_b200/csrc/solo_env.cuh"
#include "../../solorl_b200/csrc/solo_host_model.h"

using namespace solo;

struct Emu {
  ModelConst mc;
  SimConst sc;
  int njl, nj, d0, d, act_dim;
  BaseState st;
  float q[12], qd[12];
  float cforce[4];
  EnvBook bk;
  float hist[8][64];
  float goal[2], potential, progress, goal_radius;
  uint64_t seed;
  int64_t env_id;
  int need_reset;
  int settle_last;
};

static float sum4(const float* x) { return (x[0] + x[1]) + (x[2] + x[3]); } /* xor-butterfly order */

template <int NJL>
static void emu_fd(Emu* e, const float* tau, float* out) {
  BaseWork bw;
  base_prepare(e->st, bw);
  Lane<NJL> ln[4];
  Sym6 IA[4];
  float pA[4][6];
  for (int l = 0; l < 4; l++) {
    for (int k = 0; k < NJL; k++) { ln[l].q[k] = e->q[l * NJL + k]; ln[l].qd[k] = e->qd[l * NJL + k]; }
    leg_inward<NJL>(e->mc.leg[l], e->sc, bw, ln[l], tau + l * NJL, IA[l], pA[l]);
  }
  Sym6 I0;
  float p0[6], a0[6];
  for (int i = 0; i < 6; i++) { float x[4] = {IA[0].A[i], IA[1].A[i], IA[2].A[i], IA[3].A[i]}; I0.A[i] = sum4(x); }
  for (int i = 0; i < 9; i++) { float x[4] = {IA[0].H[i], IA[1].H[i], IA[2].H[i], IA[3].H[i]}; I0.H[i] = sum4(x); }
  for (int i = 0; i < 6; i++) { float x[4] = {IA[0].M[i], IA[1].M[i], IA[2].M[i], IA[3].M[i]}; I0.M[i] = sum4(x); }
  for (int i = 0; i < 6; i++) { float x[4] = {pA[0][i], pA[1][i], pA[2][i], pA[3][i]}; p0[i] = sum4(x); }
  base_solve(e->mc, e->sc, bw, I0, p0, a0);
  for (int l = 0; l < 4; l++) leg_outward<NJL>(ln[l], a0, out + 6 + l * NJL);
  base_world_acc(e->sc, bw, a0, out, out + 3);
}

template <int NJL>
static void emu_substep(Emu* e, const float* tau) {
  const SimConst& sc = e->sc;
  BaseWork bw;
  base_prepare(e->st, bw);
  Lane<NJL> ln[4];
  Sym6 IA[4];
  float pA[4][6];
  for (int l = 0; l < 4; l++) {
    for (int k = 0; k < NJL; k++) { ln[l].q[k] = e->q[l * NJL + k]; ln[l].qd[k] = e->qd[l * NJL + k]; }
    leg_inward<NJL>(e->mc.leg[l], sc, bw, ln[l], tau + l * NJL, IA[l], pA[l]);
  }
  Sym6 I0;
  float p0[6], a0[6];
  for (int i = 0; i < 6; i++) { float x[4] = {IA[0].A[i], IA[1].A[i], IA[2].A[i], IA[3].A[i]}; I0.A[i] = sum4(x); }
  for (int i = 0; i < 9; i++) { float x[4] = {IA[0].H[i], IA[1].H[i], IA[2].H[i], IA[3].H[i]}; I0.H[i] = sum4(x); }
  for (int i = 0; i < 6; i++) { float x[4] = {IA[0].M[i], IA[1].M[i], IA[2].M[i], IA[3].M[i]}; I0.M[i] = sum4(x); }
  for (int i = 0; i < 6; i++) { float x[4] = {pA[0][i], pA[1][i], pA[2][i], pA[3][i]}; p0[i] = sum4(x); }
  base_solve(e->mc, sc, bw, I0, p0, a0);
  float aw[3], al[3];
  base_world_acc(sc, bw, a0, aw, al);
  for (int l = 0; l < 4; l++) {
    float qdd[NJL];
    leg_outward<NJL>(ln[l], a0, qdd);
    for (int k = 0; k < NJL; k++) ln[l].qd[k] = clampf(ln[l].qd[k] + sc.dt * qdd[k], -sc.vmax, sc.vmax);
  }
  base_add_velocity(sc, e->st, aw, al, sc.dt);
  unsigned mask = 0;
  for (int l = 0; l < 4; l++) {
    contact_setup<NJL>(e->mc.leg[l], e->mc, sc, e->st, bw, ln[l]);
    if (ln[l].active) mask |= 1u << l;
  }
  float lam[kRows];
  for (int r = 0; r < kRows; r++) lam[r] = 0.f;
  /* joint-limit rows: the four-row lane program (group_substep's `lwarp` branch) */
  unsigned lmask = 0;
  LimitRow<NJL> lr[4];
  float lamL[4] = {0.f, 0.f, 0.f, 0.f};
  for (int l = 0; l < 4; l++) {
    int kL; float dirL, penL;
    const bool any = limit_select<NJL>(sc, ln[l], kL, dirL, penL);
    limit_setup<NJL>(sc, bw, ln[l], any, kL, dirL, penL, lr[l]);
    if (any) lmask |= 1u << l;
  }
  /* knees / base corners on the ground: the env takes the general row-record path (solo_body.cuh) */
  unsigned pmask = mask;
  BodyLaneGeom<NJL> bg[4];
  if (sc.body_contacts) {
    for (int l = 0; l < 4; l++) {
      body_lane_geometry<NJL>(sc, e->st, bw, ln[l], l, bg[l]);
      if (bg[l].kn_on) pmask |= 1u << (4 + l);
      if (bg[l].c_on[0]) pmask |= 1u << (8 + l);
      if (bg[l].c_on[1]) pmask |= 1u << (12 + l);
    }
  }
  const bool body_env = (pmask >> 4) != 0;
  if (body_env) {
    static float rows[kBodyEnvStride];
    for (int l = 0; l < 4; l++) body_lane_fill<NJL>(sc, e->st, bw, ln[l], &lr[l], l, ln[l].active != 0, bg[l], rows);
    float dv0[6], s[4][3];
    if (sc.cone) body_pgs<true>(rows, lmask, pmask, sc, 0xFu, dv0, s);
    else body_pgs<false>(rows, lmask, pmask, sc, 0xFu, dv0, s);
    for (int l = 0; l < 4; l++) {
      body_apply_leg<NJL>(ln[l], sc, s[l], dv0);
      e->cforce[l] = ln[l].active ? rows[body_slot(l, 0) * kBodyRowW + kBrLam] * sc.inv_dt : -1.0f;
    }
    float dw[3], dvl[3];
    mat3_mulv(bw.R, dv0, dw);
    mat3_mulv(bw.R, dv0 + 3, dvl);
    base_add_velocity(sc, e->st, dw, dvl, 1.0f);
    integrate_base(sc, e->st);
    for (int l = 0; l < 4; l++)
      for (int k = 0; k < NJL; k++) {
        e->qd[l * NJL + k] = ln[l].qd[k];
        e->q[l * NJL + k] = ln[l].q[k] + sc.dt * ln[l].qd[k];
      }
    return;
  }
  if (lmask) {
    PgsLane4 pl[4];
    for (int l = 0; l < 4; l++) {
      float rows[4][kRowsL];
      for (int j = 0; j < 4; j++) assemble_block4<NJL>(ln[l], lr[l], l, j, ln[j].K, lr[j].K, ((lmask >> j) & 1u) != 0, rows);
      pgs_lane_init4<NJL>(ln[l], lr[l], l, rows, mask, lmask, sc.lim_max_impulse, pl[l]);
    }
    for (int it = 0; it < sc.iters; it++) {
      float res2 = 0.f;
      {   /* the limit rows of the four legs: one simultaneous group */
        float dl[4];
        for (int f = 0; f < 4; f++) {
          float nv, rv;
          pgs_limit_candidate(pl[f], sc.lim_max_impulse, nv, dl[f], rv);
          pgs_limit_commit(pl[f], sc.lim_max_impulse, nv);
          res2 = fmaxf(res2, rv * rv);
        }
        for (int f = 0; f < 4; f++)
          for (int l = 0; l < 4; l++) pgs_apply(pl[l], limit_col(f), dl[f]);
      }
      for (int f = 0; f < 4; f++) {
        if (!((mask >> f) & 1u)) continue;
        float nv, d, rv;
        pgs_normal_candidate(pl[f], nv, d, rv);
        pgs_normal_commit(pl[f], nv);
        for (int l = 0; l < 4; l++) pgs_apply(pl[l], row_of(f, 0), d);
        res2 = fmaxf(res2, rv * rv);
      }
      for (int f = 0; f < 4; f++) {
        if (!((mask >> f) & 1u)) continue;
        if (sc.cone) {
          float nA, nB, dA, dB, rv;
          pgs_cone_candidate(pl[f], sc.mu, nA, nB, dA, dB, rv);
          pl[f].lam[1] = nA; pl[f].lam[2] = nB;
          for (int l = 0; l < 4; l++) { pgs_apply(pl[l], row_of(f, 1), dA); pgs_apply(pl[l], row_of(f, 2), dB); }
          res2 = fmaxf(res2, rv * rv);
        } else {
          for (int q = 0; q < 2; q++) {
            float nv, d, rv;
            pgs_pyramid_candidate(pl[f], sc.mu, q, nv, d, rv);
            pl[f].lam[1 + q] = nv;
            for (int l = 0; l < 4; l++) pgs_apply(pl[l], row_of(f, 1 + q), d);
            res2 = fmaxf(res2, rv * rv);
          }
        }
      }
      if (res2 <= sc.res_thr) break;
    }
    for (int l = 0; l < 4; l++) {
      for (int m = 0; m < 3; m++) lam[row_of(l, m)] = pl[l].lam[m];
      lamL[l] = pl[l].lam[3];
    }
  } else if (mask) {
    PgsLane pl[4];
    for (int l = 0; l < 4; l++) {
      float rows[3][kRows];
      for (int j = 0; j < 4; j++) assemble_block<NJL>(ln[l], l, j, ln[j].K, rows);
      pgs_lane_init<NJL>(ln[l], l, rows, mask, pl[l]);
    }
    for (int it = 0; it < sc.iters; it++) {
      float res2 = 0.f;
      for (int f = 0; f < 4; f++) {
        if (!((mask >> f) & 1u)) continue;
        float nv, d, rv;
        pgs_normal_candidate(pl[f], nv, d, rv);   /* owner lane; "shuffle" = plain read */
        pgs_normal_commit(pl[f], nv);
        for (int l = 0; l < 4; l++) pgs_apply(pl[l], row_of(f, 0), d);
        res2 = fmaxf(res2, rv * rv);
      }
      for (int f = 0; f < 4; f++) {
        if (!((mask >> f) & 1u)) continue;
        if (sc.cone) {
          float nA, nB, dA, dB, rv;
          pgs_cone_candidate(pl[f], sc.mu, nA, nB, dA, dB, rv);
          pl[f].lam[1] = nA; pl[f].lam[2] = nB;
          for (int l = 0; l < 4; l++) { pgs_apply(pl[l], row_of(f, 1), dA); pgs_apply(pl[l], row_of(f, 2), dB); }
          res2 = fmaxf(res2, rv * rv);
        } else {
          for (int q = 0; q < 2; q++) {
            float nv, d, rv;
            pgs_pyramid_candidate(pl[f], sc.mu, q, nv, d, rv);
            pl[f].lam[1 + q] = nv;
            for (int l = 0; l < 4; l++) pgs_apply(pl[l], row_of(f, 1 + q), d);
            res2 = fmaxf(res2, rv * rv);
          }
        }
      }
      if (res2 <= sc.res_thr) break;
    }
    for (int l = 0; l < 4; l++)
      for (int m = 0; m < 3; m++) lam[row_of(l, m)] = pl[l].lam[m];
  }
  float dv0[6];
  {
    float part[4][6];
    for (int l = 0; l < 4; l++) {
      float lam4[4] = {lam[row_of(l, 0)], lam[row_of(l, 1)], lam[row_of(l, 2)], lamL[l]};
      if (lmask) impulse_base_part4<NJL>(ln[l], lr[l], lam4, part[l]);
      else impulse_base_part<NJL>(ln[l], lam4, part[l]);
    }
    for (int i = 0; i < 6; i++) { float x[4] = {part[0][i], part[1][i], part[2][i], part[3][i]}; dv0[i] = sum4(x); }
  }
  for (int l = 0; l < 4; l++) {
    float lam3[4] = {lam[row_of(l, 0)], lam[row_of(l, 1)], lam[row_of(l, 2)], lamL[l]};
    if (lmask) impulse_leg4<NJL>(ln[l], lr[l], sc, lam3, dv0);
    else if (mask) impulse_leg<NJL>(ln[l], sc, lam3, dv0);
    e->cforce[l] = ln[l].active ? lam3[0] * sc.inv_dt : -1.0f;
  }
  if (mask || lmask) {
    float dw[3], dvl[3];
    mat3_mulv(bw.R, dv0, dw);
    mat3_mulv(bw.R, dv0 + 3, dvl);
    base_add_velocity(sc, e->st, dw, dvl, 1.0f);
  }
  integrate_base(sc, e->st);
  for (int l = 0; l < 4; l++)
    for (int k = 0; k < NJL; k++) {
      e->qd[l * NJL + k] = ln[l].qd[k];
      e->q[l * NJL + k] = ln[l].q[k] + sc.dt * ln[l].qd[k];
    }
}

static void emu_cur_state(Emu* e, float* s) { /* solo.py:198-222 */
  int njl = e->njl;
  cur_base(e->st, s);
  for (int l = 0; l < 4; l++)
    for (int k = 0; k < njl; k++) {
      s[idx_q(njl, l, k)] = e->q[l * njl + k] / e->sc.q_limit;
      s[idx_qd(njl, l, k)] = e->qd[l * njl + k] / e->sc.qd_limit;
    }
  for (int l = 0; l < 4; l++) s[idx_flag(njl, l)] = contact_flag(e->sc, e->cforce[l]);
  if (e->sc.task == 2) cur_pointgoal(e->st, e->goal, s + idx_pg(njl));
}

static void emu_obs(Emu* e, float* obs) { /* solo.py:186-196 */
  float cur[64];
  emu_cur_state(e, cur);
  for (int k = 0; k < e->d0; k++) obs[k] = cur[k];
  for (int h = 0; h < e->sc.H; h++)
    for (int k = 0; k < e->d0; k++) obs[(1 + h) * e->d0 + k] = cur[k] - e->hist[h][k];
}

static void emu_rng(Emu* e, uint32_t* w) {
  w[0] = (uint32_t)((uint64_t)e->env_id & 0xffffffffu);
  w[1] = (uint32_t)((uint64_t)e->env_id >> 32);
  w[2] = e->bk.episode;
  w[3] = e->bk.draw++;
  philox4x32_10(w, (uint32_t)(e->seed & 0xffffffffu), (uint32_t)(e->seed >> 32));
}

static void emu_sim_step(Emu* e, const float* tau) { /* solo.py:261-274 */
  float zero[12] = {0};
  for (int h = e->sc.H - 1; h > 0; h--) memcpy(e->hist[h], e->hist[h - 1], sizeof(float) * e->d0);
  if (e->sc.H > 0) emu_cur_state(e, e->hist[0]);
  for (int s = 0; s < e->sc.frame_skip; s++) {
    const float* t = (tau && (s == 0 || e->sc.torque_hold)) ? tau : zero;
    if (e->njl == 3) emu_substep<3>(e, t); else emu_substep<2>(e, t);
  }
  if (e->sc.task == 2) {
    float oldp = e->potential;
    e->potential = calc_potential(e->st, e->goal);
    e->progress = oldp - e->potential;
    if (e->potential < e->sc.goal_reach) {
      e->bk.goals += 1;
      uint32_t w[4];
      emu_rng(e, w);
      sample_goal(w, e->goal_radius, e->goal);
    }
  }
}

extern "C" {

void* emu_create(const SoloModelTable* m, const SoloSimParams* p, uint64_t seed, int64_t env_id) {
  Emu* e = (Emu*)calloc(1, sizeof(Emu));
  std::string err;
  if (build_model_const(*m, e->mc, err) != 0 || build_sim_const(*p, e->sc, err) != 0) { free(e); return nullptr; }
  e->njl = e->mc.njl; e->nj = 4 * e->njl;
  e->act_dim = e->nj + (e->sc.control == 2 ? 2 : 0);
  e->d0 = obs_dim0(e->njl, e->sc.task);
  e->d = e->d0 * (1 + e->sc.H);
  reset_pose(e->sc, e->st);
  for (int l = 0; l < 4; l++) e->cforce[l] = -1.f;
  e->goal_radius = (float)p->goal_radius;
  e->seed = seed; e->env_id = env_id;
  e->need_reset = 1;
  return e;
}
void emu_destroy(void* h) { free(h); }
int emu_obs_dim(void* h) { return ((Emu*)h)->d; }

void emu_set_state(void* h, const float* s) {
  Emu* e = (Emu*)h;
  for (int k = 0; k < 3; k++) { e->st.p[k] = s[k]; e->st.v[k] = s[7 + k]; e->st.w[k] = s[10 + k]; }
  for (int k = 0; k < 4; k++) e->st.q[k] = s[3 + k];
  normalize_quat(e->st.q);
  for (int j = 0; j < e->nj; j++) { e->q[j] = s[13 + j]; e->qd[j] = s[13 + e->nj + j]; }
  for (int l = 0; l < 4; l++) e->cforce[l] = -1.f;
  for (int hh = 0; hh < e->sc.H; hh++) emu_cur_state(e, e->hist[hh]);
  if (e->sc.task == 2) { e->potential = calc_potential(e->st, e->goal); e->progress = 0; }
  book_clear_episode(e->bk);
  e->need_reset = 0;
}
void emu_get_state(void* h, float* s) {
  Emu* e = (Emu*)h;
  for (int k = 0; k < 3; k++) { s[k] = e->st.p[k]; s[7 + k] = e->st.v[k]; s[10 + k] = e->st.w[k]; }
  for (int k = 0; k < 4; k++) s[3 + k] = e->st.q[k];
  for (int j = 0; j < e->nj; j++) { s[13 + j] = e->q[j]; s[13 + e->nj + j] = e->qd[j]; }
}
void emu_set_goal(void* h, float gx, float gy) {
  Emu* e = (Emu*)h;
  e->goal[0] = gx; e->goal[1] = gy; e->potential = calc_potential(e->st, e->goal); e->progress = 0;
}
void emu_forward_dynamics(void* h, const float* tau, float* out) {
  Emu* e = (Emu*)h;
  if (e->njl == 3) emu_fd<3>(e, tau, out); else emu_fd<2>(e, tau, out);
}
void emu_substep(void* h, const float* tau) {
  Emu* e = (Emu*)h;
  if (e->njl == 3) emu_substep<3>(e, tau); else emu_substep<2>(e, tau);
}
void emu_get_contacts(void* h, float* out) {
  Emu* e = (Emu*)h;
  for (int l = 0; l < 4; l++) {
    out[l * 3] = contact_flag(e->sc, e->cforce[l]);
    out[l * 3 + 1] = e->cforce[l] >= 0.f ? 1.f : 0.f;
    out[l * 3 + 2] = e->cforce[l] >= 0.f ? e->cforce[l] : 0.f;
  }
}
void emu_action_to_torque(void* h, const float* a, float* tau) {
  Emu* e = (Emu*)h;
  float kp = e->sc.kp, kd = e->sc.kd;
  if (e->sc.control == 2) { kp = a[e->nj]; kd = a[e->nj + 1]; }
  for (int j = 0; j < e->nj; j++) tau[j] = action_to_torque(e->sc, a[j], e->q[j], e->qd[j], kp, kd);
}
void emu_get_observation(void* h, float* obs) { emu_obs((Emu*)h, obs); }

void emu_reset(void* h, float* obs) { /* baseEnv.py:70-82 */
  Emu* e = (Emu*)h;
  reset_pose(e->sc, e->st);
  for (int j = 0; j < e->nj; j++) { e->q[j] = 0; e->qd[j] = 0; }
  for (int l = 0; l < 4; l++) e->cforce[l] = -1.f;
  e->bk.episode += 1; e->bk.draw = 0;
  uint32_t w[4];
  emu_rng(e, w);
  if (e->sc.task == 2) sample_goal(w, e->goal_radius, e->goal);
  for (int hh = 0; hh < e->sc.H; hh++) emu_cur_state(e, e->hist[hh]);
  if (e->sc.task == 2) { e->bk.goals = 0; e->potential = calc_potential(e->st, e->goal); e->progress = 0; }
  book_clear_episode(e->bk);
  e->need_reset = 0;
  int k = e->sc.settle_min + (e->sc.settle_span > 0 ? (int)(w[0] % (uint32_t)e->sc.settle_span) : 0);
  e->settle_last = k;
  for (int i = 0; i < k; i++) emu_sim_step(e, nullptr);
  if (obs) emu_obs(e, obs);
}
int emu_settle_count_last(void* h) { return ((Emu*)h)->settle_last; }

int emu_step(void* h, const float* action, int auto_reset, float* obs, float* reward, int* done,
             SoloEpisodeStats* stats) { /* baseEnv.py:42-68 */
  Emu* e = (Emu*)h;
  if (e->need_reset) return SOLO_E_STATE;
  float tau[12];
  emu_action_to_torque(h, action, tau);
  emu_sim_step(e, tau);
  e->bk.timestep += 1;
  emu_obs(e, obs);
  float sq = 0, sa = 0;
  for (int l = 0; l < 4; l++) { /* per-lane partial sums, then the butterfly */
  }
  float pq[4] = {0, 0, 0, 0}, pa[4] = {0, 0, 0, 0};
  for (int l = 0; l < 4; l++)
    for (int k = 0; k < e->njl; k++) {
      float q = e->q[l * e->njl + k], a = action[l * e->njl + k];
      pq[l] += (e->sc.task == 0) ? fabsf(q) : q * q;
      pa[l] += a * a;
    }
  sq = sum4(pq); sa = sum4(pa);
  StepOutcome o = step_outcome(e->sc, e->st, e->nj, sq, sa, e->progress, e->bk);
  if (stats) {
    stats->episode_reward = o.reward; stats->episode_return = e->bk.reward_sum;
    stats->episode_length = e->bk.timestep; stats->success = o.success; stats->timeout = o.timeout;
    stats->goals_reached = e->bk.goals_env;
    stats->dr_stand = e->bk.dr[0]; stats->dr_joint_pose = e->bk.dr[1]; stats->dr_torque = e->bk.dr[2];
    stats->dr_balance = e->bk.dr[3]; stats->dr_progress = e->bk.dr[4];
    stats->nan = 0;
  }
  *reward = o.reward; *done = o.done;
  if (o.done) { e->need_reset = 1; if (auto_reset) emu_reset(h, obs); }
  return 0;
}

}  // extern "C"
