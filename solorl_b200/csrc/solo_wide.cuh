/*
 * solo_wide.cuh — the "wide" build of the step kernel: 16 threads per environment.
 *
 * Why.  At the headline batch (4096 envs per GPU) the four-lanes-per-env kernel puts 16 384 threads on a
 * machine with 148 x 128 = 18 944 FP32 lanes: one warp per scheduler, every warp running one long dependent
 * chain (ncu, round 1: 5.8 % achieved occupancy, 21 % of the issue slots).  The work inside an env that IS
 * parallel -- the per-link kinematics / rotated inertias / bias wrenches of the 8 or 12 links, and the 12
 * contact rows (unit impulse propagated up the leg, then IA0^-1 P) -- is handed to three helper warpgroups,
 * so that each scheduler holds four warps instead of one and the serial part of a lane shrinks to what is
 * serial in the physics: the inward recursion of a leg, the 6x6 base factorisation, and the Gauss-Seidel sweeps.
 *
 * Block = 32 envs = 512 threads = 4 warpgroups of 128 threads, thread t4 = (env, leg) in each:
 *   warpgroup 0      the (env, leg) lanes of the narrow builds: state in registers, prologue / epilogue of the
 *                    env step, leg recursion + base solve, joint-limit rows, Delassus rows + PGS, impulses;
 *   warpgroup 1+k    link k of every leg (phase P1) and contact direction m = k of every foot (phase P3).
 * The warpgroups exchange through shared memory (structure of arrays over t4: conflict-free) at four block
 * barriers per substep.  Registers are re-partitioned with setmaxnreg (warp specialisation): the kernel is
 * launched at 128 registers per thread, warpgroup 0 grows to 232, the helpers shrink to 88
 * (232 x 128 + 88 x 384 = 63 488 <= 65 536).
 *
 * The arithmetic is the same functions of solo_core.cuh the narrow builds run (legkin_joint, link_bias,
 * leg_recursion, contact_row, ...), called per link / per row instead of per leg.
 */
#pragma once

namespace solo {

constexpr int kWEnvs = 32;             /* envs per block */
constexpr int kWE4 = kWEnvs * 4;       /* (env, leg) pairs per block = threads per warpgroup */
constexpr int kWThreads = 4 * kWE4;
constexpr int kWRegsLeg = 232, kWRegsHelper = 88;

struct WideShared {
  /* per env */
  float quat[4][kWEnvs], vel[6][kWEnvs];   /* base orientation, base (angular, linear) velocity, world frame */
  float Rb[9][kWEnvs];                     /* base -> world at the start of the substep */
  float F[21][kWEnvs];                     /* LDL^T factor of the base articulated inertia: L[15], dinv[6] */
  /* per (link k, env, leg) */
  float q[kMaxJL][kWE4], qd[kMaxJL][kWE4];
  float r[kMaxJL][3][kWE4], ax[kMaxJL][3][kWE4], cJ[kMaxJL][6][kWE4];
  float pk[kMaxJL][6][kWE4], com[kMaxJL][3][kWE4], Ic[kMaxJL][6][kWE4];
  float h[kMaxJL][6][kWE4], invD[kMaxJL][kWE4];
  /* per (env, leg) */
  float Rl[9][kWE4], ol[3][kWE4];
  float rc[3][kWE4], v6[6][kWE4], dist[kWE4];
  int active[kWE4];
  /* per (contact direction m, env, leg) */
  float sP[3][kMaxJL][kWE4], P[3][6][kWE4], K[3][6][kWE4], b[3][kWE4];
};

/* SOLO_TRACE builds (tools only): cycle stamps of the phase boundaries, leg lane 0 of every block, last substep */
#ifdef SOLO_TRACE
__device__ long long g_wide_trace[1024][8];
__device__ long long g_wide_trace_h[1024][3][8];
#define WIDE_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_wide_trace[blockIdx.x][i] = clock64(); } while (0)
#define WIDE_HSTAMP(role, i) do { if ((threadIdx.x & 127) == 0 && blockIdx.x < 1024) g_wide_trace_h[blockIdx.x][role][i] = clock64(); } while (0)
#else
#define WIDE_STAMP(i) do { } while (0)
#define WIDE_HSTAMP(role, i) do { } while (0)
#endif

__device__ __forceinline__ void wide_regs_grow() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kWRegsLeg));
}
__device__ __forceinline__ void wide_regs_shrink() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kWRegsHelper));
}

/* P1, helper warpgroup 1+K: kinematics down to link K of one leg, then that link's COM, rotated inertia and
 * bias wrench (the outward half of leg_inward, one link per thread). */
template <int NJL, int K>
__device__ __forceinline__ void wide_link(const LegConst& lc, const SimConst& sc, WideShared& S, int t4) {
  const int el = t4 >> 2;
  BaseState st;
#pragma unroll
  for (int i = 0; i < 4; i++) st.q[i] = S.quat[i][el];
#pragma unroll
  for (int i = 0; i < 3; i++) { st.w[i] = S.vel[i][el]; st.v[i] = S.vel[3 + i][el]; st.p[i] = 0.f; }
  BaseWork bw;
  base_prepare(st, bw);
  LegKin kn;
  legkin_init(bw, kn);
  float r[3], a[3], cJ[6];
#pragma unroll
  for (int j = 0; j <= K; j++) legkin_joint<NJL>(lc, j, S.q[j][t4], S.qd[j][t4], kn, r, a, cJ);
  float com[3], Ic[6], pk[6];
  link_bias<NJL>(lc, sc, K, kn, com, Ic, pk);
#pragma unroll
  for (int i = 0; i < 3; i++) { S.r[K][i][t4] = r[i]; S.ax[K][i][t4] = a[i]; S.com[K][i][t4] = com[i]; }
#pragma unroll
  for (int i = 0; i < 6; i++) { S.cJ[K][i][t4] = cJ[i]; S.pk[K][i][t4] = pk[i]; S.Ic[K][i][t4] = Ic[i]; }
  if (K == NJL - 1) {
#pragma unroll
    for (int i = 0; i < 9; i++) S.Rl[i][t4] = kn.R[i];
#pragma unroll
    for (int i = 0; i < 3; i++) S.ol[i][t4] = kn.o[i];
  }
  if (K == 0 && (t4 & 3) == 0) {
#pragma unroll
    for (int i = 0; i < 9; i++) S.Rb[i][el] = bw.R[i];
  }
}

/* P3, helper warpgroup 1+m: contact row m of one foot (contact_row of solo_core.cuh fed from shared memory). */
template <int NJL>
__device__ __forceinline__ void wide_row(const SimConst& sc, WideShared& S, int t4, int m) {
  if (!__any_sync(0xffffffffu, S.active[t4] != 0)) return;   /* warp-uniform: none of these 8 envs touches down */
  const int el = t4 >> 2;
  float R[9], rc[3], v6[6], ax[NJL][3], r[NJL][3], h[NJL][6], invD[NJL];
  Ldl6 F;
#pragma unroll
  for (int i = 0; i < 9; i++) R[i] = S.Rb[i][el];
#pragma unroll
  for (int i = 0; i < 15; i++) F.L[i] = S.F[i][el];
#pragma unroll
  for (int i = 0; i < 6; i++) F.dinv[i] = S.F[15 + i][el];
#pragma unroll
  for (int i = 0; i < 3; i++) rc[i] = S.rc[i][t4];
#pragma unroll
  for (int i = 0; i < 6; i++) v6[i] = S.v6[i][t4];
#pragma unroll
  for (int k = 0; k < NJL; k++) {
    invD[k] = S.invD[k][t4];
#pragma unroll
    for (int i = 0; i < 3; i++) { ax[k][i] = S.ax[k][i][t4]; r[k][i] = S.r[k][i][t4]; }
#pragma unroll
    for (int i = 0; i < 6; i++) h[k][i] = S.h[k][i][t4];
  }
  float sPm[NJL], P[6], K[6], b;
  contact_row<NJL>(sc, R, F, ax, r, h, invD, rc, v6, S.dist[t4], m, sPm, P, K, b);
#pragma unroll
  for (int k = 0; k < NJL; k++) S.sP[m][k][t4] = sPm[k];
#pragma unroll
  for (int i = 0; i < 6; i++) { S.P[m][i][t4] = P[i]; S.K[m][i][t4] = K[i]; }
  S.b[m][t4] = b;
}

/* One substep for an (env, leg) lane of warpgroup 0; the helpers run wide_link / wide_row between the
 * barriers marked (1)..(4), in lockstep with wide_helper_substep below. */
template <int NJL, bool LIMITS>
__device__ __forceinline__ void wide_leg_substep(const LegConst& lc, const ModelConst& mc, const SimConst& sc,
                                                 WideShared& S, int t4, int leg, BaseState& st, Lane<NJL>& ln,
                                                 const float* tau, float& cforce, int& nc_sum, int& sweep_feet) {
  const int el = t4 >> 2;
  WIDE_STAMP(0);
  /* publish what the link threads need */
#pragma unroll
  for (int k = 0; k < NJL; k++) { S.q[k][t4] = ln.q[k]; S.qd[k][t4] = ln.qd[k]; }
  if (leg == 0) {
#pragma unroll
    for (int i = 0; i < 4; i++) S.quat[i][el] = st.q[i];
#pragma unroll
    for (int i = 0; i < 3; i++) { S.vel[i][el] = st.w[i]; S.vel[3 + i][el] = st.v[i]; }
  }
  __syncthreads();                                   /* (1) -> P1 */
  WIDE_STAMP(1);
  BaseWork bw;
  base_prepare(st, bw);
  __syncthreads();                                   /* (2) link data ready */
  WIDE_STAMP(2);
  float pk[NJL][6], com[NJL][3], Ic[NJL][6];
#pragma unroll
  for (int k = 0; k < NJL; k++) {
#pragma unroll
    for (int i = 0; i < 3; i++) { ln.r[k][i] = S.r[k][i][t4]; ln.ax[k][i] = S.ax[k][i][t4]; com[k][i] = S.com[k][i][t4]; }
#pragma unroll
    for (int i = 0; i < 6; i++) { ln.cJ[k][i] = S.cJ[k][i][t4]; pk[k][i] = S.pk[k][i][t4]; Ic[k][i] = S.Ic[k][i][t4]; }
  }
#pragma unroll
  for (int i = 0; i < 9; i++) ln.Rl[i] = S.Rl[i][t4];
#pragma unroll
  for (int i = 0; i < 3; i++) ln.ol[i] = S.ol[i][t4];
  Sym6 IA;
  float pA[6], a0[6];
  leg_recursion<NJL>(lc, ln, tau, pk, com, Ic, IA, pA);
  sum4_sym6(IA);
#pragma unroll
  for (int i = 0; i < 6; i++) pA[i] = sum4(pA[i]);
  base_solve(mc, sc, bw, IA, pA, a0);
  {
    float qdd[NJL], aw[3], al[3];
    leg_outward<NJL>(ln, a0, qdd);
    base_world_acc(sc, bw, a0, aw, al);
    base_add_velocity(sc, st, aw, al, sc.dt);
#pragma unroll
    for (int k = 0; k < NJL; k++) ln.qd[k] = clampf(ln.qd[k] + sc.dt * qdd[k], -sc.vmax, sc.vmax);
  }
  float rc[3], v6[6];
  contact_geometry<NJL>(lc, mc, sc, st, bw, ln, rc, v6);
  /* publish what the row threads need */
#pragma unroll
  for (int k = 0; k < NJL; k++) {
    S.invD[k][t4] = ln.invD[k];
#pragma unroll
    for (int i = 0; i < 6; i++) S.h[k][i][t4] = ln.h[k][i];
  }
#pragma unroll
  for (int i = 0; i < 3; i++) S.rc[i][t4] = rc[i];
#pragma unroll
  for (int i = 0; i < 6; i++) S.v6[i][t4] = v6[i];
  S.dist[t4] = ln.dist;
  S.active[t4] = ln.active;
  if (leg == 0) {
#pragma unroll
    for (int i = 0; i < 15; i++) S.F[i][el] = bw.F.L[i];
#pragma unroll
    for (int i = 0; i < 6; i++) S.F[15 + i][el] = bw.F.dinv[i];
  }
  int kL;
  float dirL, penL;
  const bool lim_any = LIMITS && limit_select<NJL>(sc, ln, kL, dirL, penL);
  WIDE_STAMP(3);
  const int any_contact = __syncthreads_or(ln.active);   /* (3) -> P3, skipped by everybody when nobody touches down */
  WIDE_STAMP(4);
  if (any_contact) __syncthreads();                  /* (4) rows ready */
  WIDE_STAMP(5);
  /* the same vote the row warp of these 8 envs took: without a foot near the ground there are no rows to read,
   * and the lanes hold exact zeros (a joint-limit row may still couple with them) */
  const bool have_rows = any_contact && __any_sync(0xffffffffu, ln.active != 0);
#pragma unroll
  for (int m = 0; m < 3; m++) {
#pragma unroll
    for (int i = 0; i < 6; i++) { ln.P[m][i] = have_rows ? S.P[m][i][t4] : 0.f; ln.K[m][i] = have_rows ? S.K[m][i][t4] : 0.f; }
#pragma unroll
    for (int k = 0; k < NJL; k++) ln.sP[k][m] = have_rows ? S.sP[m][k][t4] : 0.f;
    ln.b[m] = have_rows ? S.b[m][t4] : 0.f;
  }
  contact_local_block<NJL>(ln);
  contact_solve<NJL, LIMITS, false>(sc, leg, st, bw, ln, lim_any, kL, dirL, penL, cforce, nc_sum, sweep_feet);
  WIDE_STAMP(6);
}

template <int NJL>
__device__ __forceinline__ void wide_helper_substep(const LegConst& lc, const SimConst& sc, WideShared& S, int t4,
                                                    int role) {
  WIDE_HSTAMP(role, 0);
  __syncthreads();                                   /* (1) */
  WIDE_HSTAMP(role, 1);
  if (role == 0) wide_link<NJL, 0>(lc, sc, S, t4);
  else if (role == 1) wide_link<NJL, 1>(lc, sc, S, t4);
  else if (NJL == 3) wide_link<NJL, NJL - 1>(lc, sc, S, t4);
#ifdef SOLO_TRACE_TWICE     /* I-cache experiment: the same phase again, now with its code resident */
  WIDE_HSTAMP(role, 7);
#pragma unroll 1
  for (int rep = 0; rep < 1; rep++) {
    if (role == 0) wide_link<NJL, 0>(lc, sc, S, t4);
    else if (role == 1) wide_link<NJL, 1>(lc, sc, S, t4);
    else if (NJL == 3) wide_link<NJL, NJL - 1>(lc, sc, S, t4);
  }
#endif
  WIDE_HSTAMP(role, 2);
  __syncthreads();                                   /* (2) */
  WIDE_HSTAMP(role, 3);
  const int any_contact = __syncthreads_or(0);       /* (3) */
  WIDE_HSTAMP(role, 4);
  if (any_contact) {
    wide_row<NJL>(sc, S, t4, role);
    WIDE_HSTAMP(role, 5);
    __syncthreads();                                 /* (4) */
    WIDE_HSTAMP(role, 6);
  }
}

}  // namespace solo
