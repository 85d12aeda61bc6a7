// SQP iteration, version 2: the per-node work is split from the sequential recursion so that every SM runs many warps.
//
//   K0 lin_kernel      one warp = TWO horizon nodes (half-warp each). Lane-parallel unit / dual sweeps of the kinematic tree
//                      give, per node and RK2 stage, the non-trivial rows of d f/dx, d f/du and the contact kinematics with
//                      their Jacobians. Output: a compact "linearisation record" per node (LIN_STRIDE doubles).
//   K1 lq_kernel       one warp = one node. RK2 sensitivities using the row sparsity of the flow-map Jacobian, cost and soft
//                      constraints, equality constraints, least-squares projection of the contact-velocity rows (10x10 normal
//                      equations), projected LQ model. Output: "projected record" (PROJ_STRIDE doubles).
//   K2 riccati_kernel  one warp = one instance, sequential in k, only the value-function recursion: S, s, K, k.
//   K3 forward_ls      one warp = one instance: forward pass through the projected model, then the filter line search with
//                      lanes = nodes (shared with version 1: flow_map_lane / node_values_lane in hb_mpc.cuh).
//
// Block structure used throughout (x = [hbar(6) | p(3) | theta(3) | qj(10)], u = [F(12) | vj(10)]):
//   d f/dx has non-zero rows 3..11 only; d f/dF is (1/m) I in rows 0..2, (r_c - com)x / m in rows 3..5;
//   d f/dvj is -A_b^-1 A_j in rows 6..11 and I in rows 12..21;
//   the state-input equalities split into "swing force = 0" (selector rows) and contact-velocity rows that involve vj only,
//   so the projection is a 10-dimensional problem; free inputs are the stance forces and the null space of the velocity rows.
#pragma once
#include "hb_common.cuh"
#include "hb_mpc.cuh"
#include "hb_rbd.cuh"

namespace hb {

// ---------------------------------------------------------------- record layouts (doubles)
constexpr int LIN_F1 = 0, LIN_F2 = 22, LIN_A1 = 44, LIN_A2 = 242, LIN_BF1 = 440, LIN_BF2 = 476, LIN_BV1 = 512, LIN_BV2 = 572,
              LIN_EPOS = 632, LIN_EVEL = 644, LIN_DPQ = 656, LIN_DVX = 812, LIN_DVV = 1076, LIN_STRIDE = 1200;
constexpr int NTMAX = 16;   // free inputs after projection: 3 n_stance + (10 - rank of the velocity rows); 12 / 9 / 6 in regular poses
constexpr int NVMAX = 8;    // null-space columns kept for the velocity rows
constexpr int PJ_AT = 0, PJ_BT = 484, PJ_BTV = 836, PJ_QT = 858, PJ_PT = 1342, PJ_RT = 1694, PJ_QV = 1950, PJ_RV = 1972,
              PJ_PXV = 1988, PJ_NV = 2208, PJ_PEV = 2288, PJ_META = 2298, PJ_STRIDE = 2320;
// META: [0] nt, [1] n_stance_force_dims, [2] nv, [3] cost, [4] defect^2, [5] eq^2, [6] overflow flag
constexpr int RK_STRIDE = NTMAX * NX + NTMAX;  // K (nt x 22, ld 22) + kff
__host__ __device__ inline int ntp_of(int nt) { return nt <= 6 ? 6 : (nt <= 10 ? 10 : (nt <= 12 ? 12 : 16)); }

// ---------------------------------------------------------------- K0
struct LinHalf {
  double x[NX], u[NU], x2[NX], f[NX];
  double Acm[6 * 16], Jc[12 * 16], dh[6 * NDIR], dcom[3 * NDIR], dp[12 * NDIR], dv[12 * NDIR];
  double Abinv[36], AbinvAj[6 * NJ], dvb[6 * NDIR], vgen[16], epos[12], evel[12], com[3];
  double sn[NDIR], cs[NDIR];   // sines / cosines of yaw, pitch, roll and the ten joint angles, shared by all sweeps of a stage
  double part[16 * 9];         // per-lane partial sums of a chain-split sweep: P(3), Lo(3), mc(3)
  double vals[2 * 6];          // values from the two base-seeded lanes 0 (base + left chain) and 3 (right chain): P(3), mc(3)
};

// Linearise the flow map of one node at state xs (half-warp cooperative; `hl` = lane within the half, `act` = node exists).
// Writes f, compact A rows 3..11 (9x22), Bf rows 3..5 (3x12), Bv rows 6..11 (6x10); with want_ee the contact kinematics record.
__device__ __noinline__ void lin_half(LinHalf& sh, const ChainModel& cm, const double* xs, int hl, bool act, double* rec_f, double* rec_A, double* rec_Bf,
                                      double* rec_Bv, bool want_ee, double* rec) {
  const double m = c_model.total_mass;
  if (hl < NDIR) { double s, c; sincos(xs[9 + hl], &s, &c); sh.sn[hl] = s; sh.cs[hl] = c; }
  __syncwarp();
  // Lane tasks of the two chain-split sweeps (16 lanes, all busy):
  //   hl 0..2  : Euler coordinate hl,     base body + left chain        hl 3..5  : Euler coordinate hl-3, right chain
  //   hl 6..10 : left joint hl-6, left chain                            hl 11..15: right joint hl-11, right chain
  // `gen` is the generalised coordinate the lane seeds (for joints it equals the lane index). A joint coordinate only moves its own
  // chain; the three translation coordinates have closed-form columns.
  const bool euler = hl < 6;
  const int leg = (hl < 3 || (hl >= 6 && hl < 11)) ? 0 : 1;
  const int gen = euler ? 3 + (hl % 3) : hl;
  const bool with_base = hl < 3;
  const bool vlane = (hl == 0 || hl == 3);   // lanes whose VALUES (not seeds) are used: base + left chain, right chain
  {
    // sweep 1: unit generalised velocity e_gen -> column gen of the centroidal momentum matrix and of the contact Jacobians
    ChainOut<double> co;
    kin_chain_f<double>(cm, leg, with_base, [&](int i) { return xs[6 + i]; }, [&](int i) { return (i == gen) ? 1.0 : 0.0; },
                        [&](int k, double& s, double& c) { s = sh.sn[k]; c = sh.cs[k]; }, co);
#pragma unroll
    for (int i = 0; i < 3; ++i) { sh.part[hl * 9 + i] = co.P[i]; sh.part[hl * 9 + 3 + i] = co.Lo[i]; }
    if (vlane) {
#pragma unroll
      for (int i = 0; i < 3; ++i) sh.vals[(hl / 3) * 6 + 3 + i] = co.mc[i];
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int c = leg + 2 * t;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        sh.Jc[(3 * c + i) * 16 + gen] = co.cvel[3 * t + i];
        if (!euler) sh.Jc[(3 * (c ^ 1) + i) * 16 + gen] = 0.0;      // the other leg's contacts do not move
        if (vlane) sh.epos[3 * c + i] = co.cpos[3 * t + i];
      }
    }
    if (hl < 3) {   // translation columns: h = [m e; 0], every contact moves with e
#pragma unroll
      for (int r = 0; r < 6; ++r) sh.Acm[r * 16 + hl] = (r == hl) ? m : 0.0;
#pragma unroll
      for (int r = 0; r < 12; ++r) sh.Jc[r * 16 + hl] = ((r % 3) == hl) ? 1.0 : 0.0;
    }
  }
  __syncwarp();
  if (hl < NDIR) {
    // column g = 3 + hl: sum the chain parts, then refer the angular momentum to the centre of mass
    const int g = 3 + hl, la = (hl < 3) ? hl : g;
    double P[3], Lo[3], com[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      P[i] = sh.part[la * 9 + i]; Lo[i] = sh.part[la * 9 + 3 + i];
      if (hl < 3) { P[i] += sh.part[(3 + hl) * 9 + i]; Lo[i] += sh.part[(3 + hl) * 9 + 3 + i]; }
      com[i] = (sh.vals[3 + i] + sh.vals[9 + i]) / m;
    }
    sh.Acm[0 * 16 + g] = P[0]; sh.Acm[1 * 16 + g] = P[1]; sh.Acm[2 * 16 + g] = P[2];
    sh.Acm[3 * 16 + g] = Lo[0] - (com[1] * P[2] - com[2] * P[1]);
    sh.Acm[4 * 16 + g] = Lo[1] - (com[2] * P[0] - com[0] * P[2]);
    sh.Acm[5 * 16 + g] = Lo[2] - (com[0] * P[1] - com[1] * P[0]);
    if (hl == 0) { sh.com[0] = com[0]; sh.com[1] = com[1]; sh.com[2] = com[2]; }
  }
  __syncwarp();
  if (hl < 7) {
    double Ab[36], rhs[6], y[6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) Ab[6 * r + c] = sh.Acm[r * 16 + c];
    if (hl < 6) {
#pragma unroll
      for (int r = 0; r < 6; ++r) rhs[r] = (r == hl) ? 1.0 : 0.0;
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) { double s = m * xs[r]; for (int j = 0; j < NJ; ++j) s -= sh.Acm[r * 16 + 6 + j] * sh.u[12 + j]; rhs[r] = s; }
    }
    solve6(Ab, rhs, y);
    if (hl < 6) {
#pragma unroll
      for (int r = 0; r < 6; ++r) sh.Abinv[6 * r + hl] = y[r];
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r) sh.vgen[r] = y[r];
    }
  }
  if (hl >= 6) sh.vgen[hl] = sh.u[6 + hl];   // vgen[6..15] = joint velocities u[12..21]
  __syncwarp();
  if (hl < NJ) {
#pragma unroll
    for (int r = 0; r < 6; ++r) { double s = 0.0; for (int c = 0; c < 6; ++c) s += sh.Abinv[6 * r + c] * sh.Acm[c * 16 + 6 + hl]; sh.AbinvAj[r * NJ + hl] = s; }
  }
  {
    // sweep 2: dual numbers seeded along configuration coordinate gen (direction gen - 3), generalised velocity held fixed
    const int dir = gen - 3;
    ChainOut<D1> co;
    kin_chain_f<D1>(cm, leg, with_base, [&](int i) { return D1(xs[6 + i], (i == gen) ? 1.0 : 0.0); }, [&](int i) { return D1(sh.vgen[i], 0.0); },
                    [&](int k, D1& s, D1& c) { const double sv = sh.sn[k], cv = sh.cs[k]; const double on = (k == dir) ? 1.0 : 0.0; s = D1(sv, cv * on); c = D1(cv, -sv * on); },
                    co);
#pragma unroll
    for (int i = 0; i < 3; ++i) { sh.part[hl * 9 + i] = co.P[i].d; sh.part[hl * 9 + 3 + i] = co.Lo[i].d; sh.part[hl * 9 + 6 + i] = co.mc[i].d; }
    if (vlane) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { sh.vals[(hl / 3) * 6 + i] = co.P[i].v; sh.vals[(hl / 3) * 6 + 3 + i] = co.mc[i].v; }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int c = leg + 2 * t;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int row = 3 * c + i;
        sh.dp[row * NDIR + dir] = co.cpos[3 * t + i].d;
        sh.dv[row * NDIR + dir] = co.cvel[3 * t + i].d;
        if (!euler) { sh.dp[(3 * (c ^ 1) + i) * NDIR + dir] = 0.0; sh.dv[(3 * (c ^ 1) + i) * NDIR + dir] = 0.0; }
        if (vlane) sh.evel[row] = co.cvel[3 * t + i].v;
      }
    }
  }
  __syncwarp();
  if (hl < NDIR) {
    const int la = (hl < 3) ? hl : 3 + hl;
    double Pd[3], Ld[3], dc[3], Pv[3], cv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Pd[i] = sh.part[la * 9 + i]; Ld[i] = sh.part[la * 9 + 3 + i]; dc[i] = sh.part[la * 9 + 6 + i];
      if (hl < 3) { Pd[i] += sh.part[(3 + hl) * 9 + i]; Ld[i] += sh.part[(3 + hl) * 9 + 3 + i]; dc[i] += sh.part[(3 + hl) * 9 + 6 + i]; }
      dc[i] /= m;
      Pv[i] = sh.vals[i] + sh.vals[6 + i];
      cv[i] = (sh.vals[3 + i] + sh.vals[9 + i]) / m;
    }
    // d/dq of h = [P ; Lo - com x P]
    sh.dh[0 * NDIR + hl] = Pd[0]; sh.dh[1 * NDIR + hl] = Pd[1]; sh.dh[2 * NDIR + hl] = Pd[2];
    sh.dh[3 * NDIR + hl] = Ld[0] - ((dc[1] * Pv[2] - dc[2] * Pv[1]) + (cv[1] * Pd[2] - cv[2] * Pd[1]));
    sh.dh[4 * NDIR + hl] = Ld[1] - ((dc[2] * Pv[0] - dc[0] * Pv[2]) + (cv[2] * Pd[0] - cv[0] * Pd[2]));
    sh.dh[5 * NDIR + hl] = Ld[2] - ((dc[0] * Pv[1] - dc[1] * Pv[0]) + (cv[0] * Pd[1] - cv[1] * Pd[0]));
    sh.dcom[0 * NDIR + hl] = dc[0]; sh.dcom[1 * NDIR + hl] = dc[1]; sh.dcom[2 * NDIR + hl] = dc[2];
  }
  __syncwarp();
  if (hl < NDIR) {
#pragma unroll
    for (int r = 0; r < 6; ++r) { double s = 0.0; for (int c = 0; c < 6; ++c) s += sh.Abinv[6 * r + c] * sh.dh[c * NDIR + hl]; sh.dvb[r * NDIR + hl] = -s; }
  }
  // flow map value (all 16 lanes of the half take part; entries 16..21 by lanes 0..5 in a second round)
  for (int i = hl; i < NX; i += 16) {
    double val;
    if (i < 3) {
      double s = 0.0;
      for (int c = 0; c < NC; ++c) s += sh.u[3 * c + i];
      val = s / m - (i == 2 ? HB_GRAVITY : 0.0);
    } else if (i < 6) {
      const int a = i - 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
      double s = 0.0;
      for (int c = 0; c < NC; ++c) s += (sh.epos[3 * c + a1] - sh.com[a1]) * sh.u[3 * c + a2] - (sh.epos[3 * c + a2] - sh.com[a2]) * sh.u[3 * c + a1];
      val = s / m;
    } else if (i < 12) val = sh.vgen[i - 6];
    else val = sh.u[i];
    sh.f[i] = val;
    if (act) rec_f[i] = val;
  }
  __syncwarp();
  if (!act) return;
  // compact Jacobian blocks; column j of the 9x22 block by lane (two rounds for 22 columns)
  for (int j = hl; j < NX; j += 16) {
    double ca[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) ca[i] = 0.0;
    if (j < 6) {
#pragma unroll
      for (int r = 0; r < 6; ++r) ca[3 + r] = m * sh.Abinv[6 * r + j];
    } else if (j >= 9) {
      const int k = j - 9;
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
      for (int c = 0; c < NC; ++c) {
        const double d0 = sh.dp[(3 * c) * NDIR + k] - sh.dcom[k], d1 = sh.dp[(3 * c + 1) * NDIR + k] - sh.dcom[NDIR + k], d2 = sh.dp[(3 * c + 2) * NDIR + k] - sh.dcom[2 * NDIR + k];
        const double* F = sh.u + 3 * c;
        t0 += d1 * F[2] - d2 * F[1]; t1 += d2 * F[0] - d0 * F[2]; t2 += d0 * F[1] - d1 * F[0];
      }
      ca[0] = t0 / m; ca[1] = t1 / m; ca[2] = t2 / m;
#pragma unroll
      for (int r = 0; r < 6; ++r) ca[3 + r] = sh.dvb[r * NDIR + k];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) rec_A[i * NX + j] = ca[i];
  }
  if (hl < 12) {
    const int c = hl / 3, a = hl - 3 * c;
    const double r0 = (sh.epos[3 * c] - sh.com[0]) / m, r1 = (sh.epos[3 * c + 1] - sh.com[1]) / m, r2 = (sh.epos[3 * c + 2] - sh.com[2]) / m;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;   // (r - com) x e_a / m
    if (a == 0) { b1 = r2; b2 = -r1; } else if (a == 1) { b0 = -r2; b2 = r0; } else { b0 = r1; b1 = -r0; }
    rec_Bf[0 * 12 + hl] = b0; rec_Bf[1 * 12 + hl] = b1; rec_Bf[2 * 12 + hl] = b2;
  }
  if (hl < NJ) {
#pragma unroll
    for (int r = 0; r < 6; ++r) rec_Bv[r * NJ + hl] = -sh.AbinvAj[r * NJ + hl];
  }
  if (want_ee) {
    if (hl < 12) { rec[LIN_EPOS + hl] = sh.epos[hl]; rec[LIN_EVEL + hl] = sh.evel[hl]; }
    for (int j = hl; j < NX; j += 16) {
      for (int r = 0; r < 12; ++r) {
        double vx = 0.0;
        if (j < 6) { for (int c = 0; c < 6; ++c) vx += sh.Jc[r * 16 + c] * sh.Abinv[6 * c + j]; vx *= m; }
        else if (j >= 9) { const int k = j - 9; vx = sh.dv[r * NDIR + k]; for (int c = 0; c < 6; ++c) vx += sh.Jc[r * 16 + c] * sh.dvb[c * NDIR + k]; rec[LIN_DPQ + r * NDIR + k] = sh.dp[r * NDIR + k]; }
        rec[LIN_DVX + r * NX + j] = vx;
        if (j >= 12) { const int jj = j - 12; double vu = sh.Jc[r * 16 + 6 + jj]; for (int c = 0; c < 6; ++c) vu -= sh.Jc[r * 16 + c] * sh.AbinvAj[c * NJ + jj]; rec[LIN_DVV + r * NJ + jj] = vu; }
      }
    }
  }
}

struct SqpArgs {
  int B, N;
  double dt;
  const double* x_ref; const double* swing; const int32_t* mode;
  double* xt; double* ut;
  double* lin;     // B x N x LIN_STRIDE
  double* proj;    // B x N x PROJ_STRIDE
  double* rk;      // B x N x RK_STRIDE
  double* dxt; double* dut; double* perf; int32_t* flags;
  const double* x0;
};

__global__ void __launch_bounds__(64) lin_kernel(SqpArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp_in_block = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  LinHalf& sh = reinterpret_cast<LinHalf*>(smem_raw)[warp_in_block * 2 + half];
  ChainModel& cm = *reinterpret_cast<ChainModel*>(smem_raw + 4 * sizeof(LinHalf));
  chain_model_load(cm, threadIdx.x, blockDim.x);
  __syncthreads();
  const int N = a.N, NP = (N + 1) >> 1;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  if (w >= (long long)a.B * NP) return;
  const int inst = (int)(w / NP), pr = (int)(w - (long long)inst * NP);
  const int k = 2 * pr + half;
  const bool act = k < N;
  const int kk = act ? k : N - 1;
  const double* xk = (kk == 0) ? a.x0 + (size_t)inst * NX : a.xt + ((size_t)inst * (N + 1) + kk) * NX;   // node 0 is pinned to the measured state
  const double* uk = a.ut + ((size_t)inst * N + kk) * NU;
  for (int i = hl; i < NX; i += 16) { sh.x[i] = xk[i]; sh.u[i] = uk[i]; }
  __syncwarp();
  double* rec = a.lin + ((size_t)inst * N + kk) * LIN_STRIDE;
  lin_half(sh, cm, sh.x, hl, act, rec + LIN_F1, rec + LIN_A1, rec + LIN_BF1, rec + LIN_BV1, true, rec);
  for (int i = hl; i < NX; i += 16) sh.x2[i] = sh.x[i] + a.dt * sh.f[i];
  __syncwarp();
  lin_half(sh, cm, sh.x2, hl, act, rec + LIN_F2, rec + LIN_A2, rec + LIN_BF2, rec + LIN_BV2, false, rec);
}

// ---------------------------------------------------------------- K1
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NKEEP> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NKEEP) : "memory"); }

// Shared memory of the node LQ kernel. `rec` first holds the linearisation record (brought in with cp.async); once the
// discretisation, cost and constraint rows are built it is dead and re-used for the projection workspace and the projected model.
struct LqShared {
  double rec[LIN_STRIDE];
  double Ad[9 * NX];                   // rows 3..11 of the discrete A, then of At (the other rows are identity / identity + dt Pxv and are formed on the way out)
  double BdF[9 * 12], Bdv[9 * NJ];     // rows 3..11 of Bd
  double Cv[12 * NX], Dv[12 * NJ], ev[12];
  double Nv[NJ * NVMAX], pev[NJ];
  double Rvv[NJ * NJ], Pv[NJ * NX], RFF[12 * 12];
  double gx[8 * NX], gv[8 * NJ], gh[8];
  double x[NX], u[NU], xn[NX], xref[NX], swing[24], b[NX], Qd[NX], q[NX], r[NU], bt[NX], rRpe[NJ];
  int rowc[12], rowa[12], rowt[12], piv[NJ], freev[NJ], stidx[12];
};
// aliases inside `rec` after it is dead
constexpr int LQ_G = 0, LQ_GL = 34, LQ_T1 = 340, LQ_BT = 560, LQ_PXV = 912;   // G 10x34, T1 10x22, Bt 22x16, Pxv 10x22 (ends at 1132)
static_assert(LQ_PXV + NJ * NX <= LIN_STRIDE, "aliased area overflow");

template <int MR>
__device__ __forceinline__ void lq_build_tableau(const LqShared& sh, double* __restrict__ G, int lane) {
  constexpr int GL = LQ_GL;
  double xr[MR];
#pragma unroll
  for (int r = 0; r < MR; ++r) xr[r] = (lane < NJ) ? sh.Dv[r * NJ + lane] : -sh.Cv[r * NX + lane - NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < MR; ++r) s = fma(sh.Dv[r * NJ + i], xr[r], s);
    G[i * GL + lane] = s;
  }
  if (lane < NJ) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < MR; ++r) s = fma(-sh.Dv[r * NJ + lane], sh.ev[r], s);
    G[lane * GL + NJ + NX] = s;
  }
}

__global__ void __launch_bounds__(32) lq_kernel(SqpArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LqShared& sh = *reinterpret_cast<LqShared*>(smem_raw);
  const int lane = threadIdx.x;
  const int N = a.N;
  const long long w = blockIdx.x;
  const int inst = (int)(w / N), k = (int)(w - (long long)inst * N);
  const double dt = a.dt;
  const Model& md = c_model;
  const double* grec = a.lin + ((size_t)inst * N + k) * LIN_STRIDE;
  double* out = a.proj + ((size_t)inst * N + k) * PJ_STRIDE;
  const double* xk = (k == 0) ? a.x0 + (size_t)inst * NX : a.xt + ((size_t)inst * (N + 1) + k) * NX;
  // ---- asynchronous loads: record + node vectors
  for (int i = 2 * lane; i < LIN_DVV + 12 * NJ; i += 64) cp_async16(sh.rec + i, grec + i);
  if (lane < 11) {
    cp_async16(sh.x + 2 * lane, xk + 2 * lane);
    cp_async16(sh.u + 2 * lane, a.ut + ((size_t)inst * N + k) * NU + 2 * lane);
    cp_async16(sh.xn + 2 * lane, a.xt + ((size_t)inst * (N + 1) + k + 1) * NX + 2 * lane);
    cp_async16(sh.xref + 2 * lane, a.x_ref + ((size_t)inst * (N + 1) + k) * NX + 2 * lane);
  } else if (lane >= 16 && lane < 28) {
    cp_async16(sh.swing + 2 * (lane - 16), a.swing + ((size_t)inst * (N + 1) + k) * 24 + 2 * (lane - 16));
  }
  cp_async_commit();
  const int mode = a.mode[(size_t)inst * (N + 1) + k];
  bool fl[4]; int ns = 0;
  for (int c = 0; c < 4; ++c) { fl[c] = contact_flag(mode, c); ns += fl[c]; }
  // constant parts that do not need the record
  for (int idx = lane; idx < NJ * NJ; idx += 32) { const int i = idx / NJ, j = idx - i * NJ; sh.Rvv[idx] = md.R[(12 + i) * NU + 12 + j]; }
  for (int idx = lane; idx < 144; idx += 32) { const int i = idx / 12, j = idx - i * 12; sh.RFF[idx] = (i == j) ? md.R[i * NU + i] : 0.0; }
  for (int idx = lane; idx < NJ * NX; idx += 32) sh.Pv[idx] = 0.0;
  cp_async_wait<0>();
  __syncwarp();
  const double* A1c = sh.rec + LIN_A1; const double* A2c = sh.rec + LIN_A2;
  const double* Bf1 = sh.rec + LIN_BF1; const double* Bf2 = sh.rec + LIN_BF2;
  const double* Bv1 = sh.rec + LIN_BV1; const double* Bv2 = sh.rec + LIN_BV2;
  const double* f1 = sh.rec + LIN_F1; const double* f2 = sh.rec + LIN_F2;
  const double* epos = sh.rec + LIN_EPOS; const double* evel = sh.rec + LIN_EVEL;
  const double* dpq = sh.rec + LIN_DPQ; const double* dvx = sh.rec + LIN_DVX; const double* dvv = sh.rec + LIN_DVV;
  // ---- RK2 sensitivities (S2) on the non-trivial rows 3..11
  if (lane < NX) {
    const int j = lane;
    double c1[9];
#pragma unroll
    for (int kk = 0; kk < 9; ++kk) c1[kk] = A1c[kk * NX + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      double s = 0.0;
#pragma unroll
      for (int kk = 0; kk < 9; ++kk) s = fma(A2c[i * NX + 3 + kk], c1[kk], s);
      sh.Ad[i * NX + j] = 0.5 * dt * (c1[i] + A2c[i * NX + j] + dt * s) + ((3 + i == j) ? 1.0 : 0.0);
    }
    if (j < 12) {
      // B1 force column j = [e_a / m ; Bf1[:, j] ; 0]: (A2 B1)[i][j] = A2c[i][a] / m + A2c[i][3:6] Bf1[:, j]
      const double b0 = Bf1[j], b1 = Bf1[12 + j], b2 = Bf1[24 + j];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const double s = A2c[i * NX + 3] * b0 + A2c[i * NX + 4] * b1 + A2c[i * NX + 5] * b2 + A2c[i * NX + (j % 3)] / md.total_mass;
        const double base = (i < 3) ? (Bf1[i * 12 + j] + Bf2[i * 12 + j]) : 0.0;
        sh.BdF[i * 12 + j] = 0.5 * dt * (base + dt * s);
      }
    } else {
      const int jj = j - 12;
      double cb[6];
#pragma unroll
      for (int kk = 0; kk < 6; ++kk) cb[kk] = Bv1[kk * NJ + jj];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        double s = A2c[i * NX + 12 + jj];
#pragma unroll
        for (int kk = 0; kk < 6; ++kk) s = fma(A2c[i * NX + 6 + kk], cb[kk], s);
        const double base = (i >= 3) ? (Bv1[(i - 3) * NJ + jj] + Bv2[(i - 3) * NJ + jj]) : 0.0;
        sh.Bdv[i * NJ + jj] = 0.5 * dt * (base + dt * s);
      }
    }
  }
  double d2 = 0.0;
  if (lane < NX) { const double bb = sh.x[lane] + 0.5 * dt * (f1[lane] + f2[lane]) - sh.xn[lane]; sh.b[lane] = bb; d2 = bb * bb; }
  d2 = warp_sum(d2);
  // ---- cost (M2, M6, M7, M8), scaled by dt at the end
  const double fz = ns > 0 ? md.total_mass * HB_GRAVITY / ns : 0.0;
  double cost = 0.0;
  if (lane < NX) {
    const double d = sh.x[lane] - sh.xref[lane];
    sh.q[lane] = md.Q[lane] * d;
    sh.Qd[lane] = md.Q[lane];
    cost += 0.5 * md.Q[lane] * d * d;
    double s = 0.0;
    if (lane < 12) { double dul = sh.u[lane]; if ((lane % 3) == 2 && fl[lane / 3]) dul -= fz; s = md.R[lane * NU + lane] * dul; cost += 0.5 * dul * s; }
    else { for (int j = 12; j < NU; ++j) s += md.R[lane * NU + j] * sh.u[j]; cost += 0.5 * sh.u[lane] * s; }
    sh.r[lane] = s;
  }
  __syncwarp();
  // all scalar penalties in ONE pass: lanes 0-9 joint position limits, 10-19 joint velocity limits, 20-23 normal-force limits
  // (double sided), 24-27 friction cones of stance contacts (one sided)
  double shiftsum = 0.0;
  {
    double h = 1.0, lo = 0.0, hi = 2.0, pmu = 0.0, pdl = 1.0;
    bool two = true, on = false;
    double Fx = 0.0, Fy = 0.0, tn = 1.0, t2 = 1.0;
    if (lane < 10) { h = sh.x[12 + lane]; lo = md.joint_lower[lane]; hi = md.joint_upper[lane]; pmu = HB_LIMIT_POS_MU; pdl = HB_LIMIT_POS_DELTA; on = true; }
    else if (lane < 20) { const int j = lane - 10; h = sh.u[12 + j]; lo = -md.joint_vel_limit[j]; hi = md.joint_vel_limit[j]; pmu = HB_LIMIT_VEL_MU; pdl = HB_LIMIT_VEL_DELTA; on = true; }
    else if (lane < 24) { const int c = lane - 20; h = sh.u[3 * c + 2]; lo = 0.0; hi = HB_LIMIT_FORCE_MAX; pmu = HB_LIMIT_FORCE_MU; pdl = HB_LIMIT_FORCE_DELTA; on = true; }
    else if (lane < 28) {
      const int c = lane - 24;
      if (fl[c]) {
        Fx = sh.u[3 * c]; Fy = sh.u[3 * c + 1];
        t2 = Fx * Fx + Fy * Fy + HB_FRICTION_REGULARIZATION; tn = sqrt(t2);
        h = HB_FRICTION_MU * sh.u[3 * c + 2] - tn; lo = 0.0; pmu = HB_FRICTION_BARRIER_MU; pdl = HB_FRICTION_BARRIER_DELTA; two = false; on = true;
      }
    }
    const Pen pa = relaxed_barrier(h - lo, pmu, pdl);
    const Pen pb = relaxed_barrier(two ? hi - h : 1.0, two ? pmu : 0.0, pdl);
    const double pv = pa.v + pb.v, p1 = pa.d1 - pb.d1, p2 = pa.d2 + pb.d2;
    __syncwarp();
    if (on) {
      cost += pv;
      if (lane < 10) { sh.q[12 + lane] += p1; sh.Qd[12 + lane] += p2; }
      else if (lane < 20) { const int j = lane - 10; sh.r[12 + j] += p1; sh.Rvv[j * NJ + j] += p2; }
      else if (lane < 24) { const int c = lane - 20; sh.r[3 * c + 2] += p1; sh.RFF[(3 * c + 2) * 12 + 3 * c + 2] += p2; }
    }
    __syncwarp();
    if (on && lane >= 24) {
      const int c = lane - 24;
      const double t32 = tn * t2;
      const double gr[3] = {-Fx / tn, -Fy / tn, HB_FRICTION_MU};
      const double Hh[9] = {-(Fy * Fy + HB_FRICTION_REGULARIZATION) / t32, Fx * Fy / t32, 0.0, Fx * Fy / t32,
                            -(Fx * Fx + HB_FRICTION_REGULARIZATION) / t32, 0.0, 0.0, 0.0, 0.0};
      for (int i = 0; i < 3; ++i) {
        sh.r[3 * c + i] += p1 * gr[i];
        for (int j = 0; j < 3; ++j) sh.RFF[(3 * c + i) * 12 + 3 * c + j] += p2 * gr[i] * gr[j] + p1 * Hh[3 * i + j];
      }
      shiftsum = -p1 * HB_FRICTION_HESSIAN_SHIFT;
    }
  }
  shiftsum = warp_sum(shiftsum);
  __syncwarp();
  if (lane < NX) sh.Qd[lane] += shiftsum;
  if (lane < 12) sh.RFF[lane * 12 + lane] += shiftsum;
  else if (lane >= 16 && lane < 16 + NJ) sh.Rvv[(lane - 16) * NJ + lane - 16] += shiftsum;
  // xy swing soft constraint: gradients gx (22) / gv (10, vj columns only)
  int npair = 0;
  for (int c = 0; c < 4; ++c) {
    if (fl[c]) continue;
    for (int ax = 0; ax < 2; ++ax) {
      const int row = 3 * c + ax;
      if (lane < NX) {
        double px = 0.0;
        if (lane >= 6 && lane < 9) px = (lane - 6 == ax) ? 1.0 : 0.0;
        else if (lane >= 9) px = dpq[row * NDIR + lane - 9];
        sh.gx[npair * NX + lane] = dvx[row * NX + lane] + HB_XY_POSITION_GAIN * px;
        if (lane < NJ) sh.gv[npair * NJ + lane] = dvv[row * NJ + lane];
      }
      if (lane == 0) sh.gh[npair] = evel[row] - sh.swing[6 * c + 3 + ax] + HB_XY_POSITION_GAIN * (epos[row] - sh.swing[6 * c + ax]);
      ++npair;
    }
  }
  for (int idx = npair * NX + lane; idx < 8 * NX; idx += 32) sh.gx[idx] = 0.0;   // unused pairs must be exact zeros (they enter Qt with weight 0)
  __syncwarp();
  if (npair > 0) {
    const double w = HB_SOFT_SWING_WEIGHT;
    if (lane < NX) {
      double qa = 0.0;
      double pvc[NJ];
#pragma unroll
      for (int i = 0; i < NJ; ++i) pvc[i] = 0.0;
      for (int p = 0; p < npair; ++p) {
        const double h = sh.gh[p], gxj = sh.gx[p * NX + lane];
        qa += w * h * gxj;
#pragma unroll
        for (int i = 0; i < NJ; ++i) pvc[i] = fma(w * sh.gv[p * NJ + i], gxj, pvc[i]);
      }
      sh.q[lane] += qa;
#pragma unroll
      for (int i = 0; i < NJ; ++i) sh.Pv[i * NX + lane] = pvc[i];
      if (lane < NJ) {
        double ra = 0.0;
        for (int p = 0; p < npair; ++p) {
          const double gvj = sh.gv[p * NJ + lane];
          ra += w * sh.gh[p] * gvj;
#pragma unroll
          for (int i = 0; i < NJ; ++i) sh.Rvv[i * NJ + lane] += w * sh.gv[p * NJ + i] * gvj;
        }
        sh.r[12 + lane] += ra;
      }
    }
    if (lane < npair) cost += 0.5 * w * sh.gh[lane] * sh.gh[lane];
  }
  cost = warp_sum(cost);
  // ---- contact-velocity equality rows (M4, M5); swing forces handled separately (M3)
  int mr = 0;
  for (int c = 0; c < 4; ++c) {
    if (fl[c]) { for (int ax = 0; ax < 3; ++ax) { if (lane == 0) { sh.rowc[mr] = c; sh.rowa[mr] = ax; sh.rowt[mr] = 0; } ++mr; } }
    else { if (lane == 0) { sh.rowc[mr] = c; sh.rowa[mr] = 2; sh.rowt[mr] = 2; } ++mr; }
  }
  __syncwarp();
  double e2 = 0.0;
  if (lane < NX) {
    for (int rr = 0; rr < mr; ++rr) {
      const int ra = sh.rowa[rr], row = 3 * sh.rowc[rr] + ra, t = sh.rowt[rr];
      const double gain = (t == 0) ? ((ra == 2) ? HB_ZEROVEL_Z_GAIN : 0.0) : HB_POSITION_ERROR_GAIN;
      double px = 0.0;
      if (lane >= 6 && lane < 9) px = (lane - 6 == ra) ? 1.0 : 0.0;
      else if (lane >= 9) px = dpq[row * NDIR + lane - 9];
      sh.Cv[rr * NX + lane] = dvx[row * NX + lane] + gain * px;
      if (lane < NJ) sh.Dv[rr * NJ + lane] = dvv[row * NJ + lane];
    }
  }
  if (lane < mr) {
    const int c = sh.rowc[lane], ax = sh.rowa[lane], row = 3 * c + ax;
    double evv;
    if (sh.rowt[lane] == 0) evv = evel[row] + (ax == 2 ? HB_ZEROVEL_Z_GAIN * epos[row] + HB_ZEROVEL_Z_OFFSET : 0.0);
    else evv = evel[row] - sh.swing[6 * c + 5] + HB_POSITION_ERROR_GAIN * (epos[row] - sh.swing[6 * c + 2]);
    sh.ev[lane] = evv;
    e2 = evv * evv;
  }
  if (lane >= 16 && lane < 28) { const int j = lane - 16; if (!fl[j / 3]) e2 += sh.u[j] * sh.u[j]; }
  e2 = warp_sum(e2);
  __syncwarp();
  // ================= the record is dead from here on: its storage holds G, T1, Bt, Pxv =================
  double* G = sh.rec + LQ_G; double* T1 = sh.rec + LQ_T1; double* Bt = sh.rec + LQ_BT; double* Pxv = sh.rec + LQ_PXV;
  // ---- least-squares projection on vj: G = Dv'Dv, rhs = -Dv'[Cv | ev]  (10 x 33, ld 34), entries spread over all lanes
  constexpr int GW = NJ + NX + 1, GL = LQ_GL;
  // lane j owns column j of the tableau: its right operand X[r][j] (Dv, -Cv) stays in registers, Dv[r][i] is a broadcast load.
  // mr is 4 (flight), 8 (single support) or 12 (stance) for the contact pairs of this robot.
  if (mr == 4) lq_build_tableau<4>(sh, G, lane);
  else if (mr == 8) lq_build_tableau<8>(sh, G, lane);
  else if (mr == 12) lq_build_tableau<12>(sh, G, lane);
  else {
    for (int idx = lane; idx < NJ * GW; idx += 32) {
      const int i = idx / GW, j = idx - i * GW;
      double s = 0.0;
      if (j < NJ) { for (int r = 0; r < mr; ++r) s = fma(sh.Dv[r * NJ + i], sh.Dv[r * NJ + j], s); }
      else if (j < NJ + NX) { for (int r = 0; r < mr; ++r) s = fma(-sh.Dv[r * NJ + i], sh.Cv[r * NX + j - NJ], s); }
      else { for (int r = 0; r < mr; ++r) s = fma(-sh.Dv[r * NJ + i], sh.ev[r], s); }
      G[i * GL + j] = s;
    }
  }
  if (lane < NJ) sh.piv[lane] = 0;
  __syncwarp();
  double dmax = lane < NJ ? G[lane * GL + lane] : 0.0;
  if (ns < 4) dmax = fmax(dmax, 1.0);     // the swing-force selector rows of D have unit diagonal in D'D
  dmax = warp_max(dmax);
  const double tol = 1e-9 * fmax(dmax, 1e-300);
  for (int step = 0; step < NJ; ++step) {
    double dv = (lane < NJ && !sh.piv[lane]) ? G[lane * GL + lane] : -1.0;
    int pi = lane;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {      // candidates live in lanes 0..9: a 16-lane butterfly suffices
      const double ov = __shfl_xor_sync(HB_FULL_MASK, dv, o);
      const int oi = __shfl_xor_sync(HB_FULL_MASK, pi, o);
      if (ov > dv || (ov == dv && oi < pi)) { dv = ov; pi = oi; }
    }
    dv = __shfl_sync(HB_FULL_MASK, dv, 0); pi = __shfl_sync(HB_FULL_MASK, pi, 0);
    if (!(dv > tol)) break;
    const int p = pi;
    const double inv = 1.0 / dv;
    // lane j owns column j (j < 32) of the 10 x 33 tableau; the last column (index 32) is updated by lane i for row i
    static_assert(GW == 33, "column ownership below assumes 33 columns");
    const double pj = G[p * GL + lane] * inv;                   // scaled pivot-row entry of this lane's column
    const double p32 = G[p * GL + 32] * inv;
    const double colv = (lane < NJ) ? G[lane * GL + p] : 0.0;   // column p before elimination, one entry per lane
    __syncwarp();
    G[p * GL + lane] = pj;
    if (lane == 0) { G[p * GL + 32] = p32; sh.piv[p] = 1; }
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const double ci = __shfl_sync(HB_FULL_MASK, colv, i);
      if (i != p) G[i * GL + lane] -= ci * pj;
    }
    if (lane < NJ && lane != p) G[lane * GL + 32] -= colv * p32;
    __syncwarp();
  }
  int nv = 0;
  for (int i = 0; i < NJ; ++i) if (!sh.piv[i]) { if (lane == 0) sh.freev[nv] = i; ++nv; }
  int nf = 0;
  for (int j = 0; j < 12; ++j) if (fl[j / 3]) { if (lane == 0) sh.stidx[nf] = j; ++nf; }
  bool overflow = false;
  if (nv > NVMAX) { nv = NVMAX; overflow = true; }       // degenerate pose (velocity rows lost rank): flagged, instance reported as failed
  if (nf + nv > NTMAX) { nv = NTMAX - nf; overflow = true; }
  const int nt = nf + nv;
  const int ntp = ntp_of(nt);
  __syncwarp();
  if (lane < NX) {
#pragma unroll
    for (int i = 0; i < NJ; ++i) Pxv[i * NX + lane] = sh.piv[i] ? G[i * GL + NJ + lane] : 0.0;
    if (lane < NJ) sh.pev[lane] = sh.piv[lane] ? G[lane * GL + NJ + NX] : 0.0;
    if (lane < nv) { const int fc = sh.freev[lane]; for (int i = 0; i < NJ; ++i) sh.Nv[i * NVMAX + lane] = sh.piv[i] ? -G[i * GL + fc] : ((i == fc) ? 1.0 : 0.0); }
  }
  for (int idx = lane; idx < NX * NTMAX; idx += 32) Bt[idx] = 0.0;
  __syncwarp();
  // ---- projected model (cost scaled by dt)
  if (lane < NX) {
    const int j = lane;
    double pc[NJ];
#pragma unroll
    for (int kk = 0; kk < NJ; ++kk) pc[kk] = Pxv[kk * NX + j];
    // At = Ad + Bd_v Pxv : rows 3..11 += Bdv Pxv ; rows 12..21 += dt Pxv
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      double s = sh.Ad[i * NX + j];
#pragma unroll
      for (int kk = 0; kk < NJ; ++kk) s = fma(sh.Bdv[i * NJ + kk], pc[kk], s);
      sh.Ad[i * NX + j] = s;
    }
    // PRPx_v = Pv + Rvv Pxv (10 x 22)
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      double s = sh.Pv[i * NX + j];
#pragma unroll
      for (int kk = 0; kk < NJ; ++kk) s = fma(sh.Rvv[i * NJ + kk], pc[kk], s);
      T1[i * NX + j] = s;
    }
  }
  if (lane < NJ) {
    double s = sh.r[12 + lane];
#pragma unroll
    for (int kk = 0; kk < NJ; ++kk) s = fma(sh.Rvv[lane * NJ + kk], sh.pev[kk], s);
    sh.rRpe[lane] = s;
  }
  __syncwarp();
  // Qt = diag(Qd) + sum_p w gx_p gx_p' + Pxv' PRPx_v + Pv' Pxv ; qt = q + Pxv' rRpe + Pv' pev   (column j by lane j)
  if (lane < NX) {
    const int j = lane;
    double tc[NJ], pc[NJ], gj[8];
#pragma unroll
    for (int kk = 0; kk < NJ; ++kk) { tc[kk] = T1[kk * NX + j]; pc[kk] = Pxv[kk * NX + j]; }
#pragma unroll
    for (int p = 0; p < 8; ++p) gj[p] = (p < npair) ? HB_SOFT_SWING_WEIGHT * sh.gx[p * NX + j] : 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double s = (i == j) ? sh.Qd[i] : 0.0;
#pragma unroll
      for (int kk = 0; kk < NJ; ++kk) { s = fma(Pxv[kk * NX + i], tc[kk], s); s = fma(sh.Pv[kk * NX + i], pc[kk], s); }
      if (npair > 0) {
#pragma unroll
        for (int p = 0; p < 8; ++p) s = fma(sh.gx[p * NX + i], gj[p], s);
      }
      out[PJ_QT + i * NX + j] = dt * s;
    }
    double s = sh.q[j];
#pragma unroll
    for (int kk = 0; kk < NJ; ++kk) s += pc[kk] * sh.rRpe[kk] + sh.Pv[kk * NX + j] * sh.pev[kk];
    out[PJ_QV + j] = dt * s;
  }
  // Bt (22 x nt, ld NTMAX): stance-force columns then null-space columns; padded columns stay zero
  if (lane < nt) {
    const int c = lane;
    if (c < nf) {
      const int j = sh.stidx[c];
      Bt[(j % 3) * NTMAX + c] = dt / md.total_mass;
#pragma unroll
      for (int i = 0; i < 9; ++i) Bt[(3 + i) * NTMAX + c] = sh.BdF[i * 12 + j];
    } else {
      const int cc = c - nf;
#pragma unroll
      for (int i = 0; i < 9; ++i) { double s = 0.0; for (int kk = 0; kk < NJ; ++kk) s = fma(sh.Bdv[i * NJ + kk], sh.Nv[kk * NVMAX + cc], s); Bt[(3 + i) * NTMAX + c] = s; }
      for (int kk = 0; kk < NJ; ++kk) Bt[(12 + kk) * NTMAX + c] = dt * sh.Nv[kk * NVMAX + cc];
    }
  }
  // bt = b + Bd_v pev - Bd_F[:, swing] F_swing
  if (lane < NX) {
    const int i = lane;
    double s = sh.b[i];
    if (i >= 3 && i < 12) {
      for (int kk = 0; kk < NJ; ++kk) s = fma(sh.Bdv[(i - 3) * NJ + kk], sh.pev[kk], s);
      for (int j = 0; j < 12; ++j) if (!fl[j / 3]) s -= sh.BdF[(i - 3) * 12 + j] * sh.u[j];
    } else if (i >= 12) s += dt * sh.pev[i - 12];
    else { for (int c = 0; c < 4; ++c) if (!fl[c]) s -= dt / md.total_mass * sh.u[3 * c + i]; }
    sh.bt[i] = s;
  }
  __syncwarp();
  // Rt (ntp x ntp), Pt (ntp x 22), rt (ntp); RN = Rvv Nv (10 x nv) goes into the dead constraint-row storage first
  double* RN = sh.Cv;
  for (int idx = lane; idx < NJ * NVMAX; idx += 32) {
    const int p = idx / NVMAX, cj = idx - p * NVMAX;
    double t = 0.0;
    if (cj < nv) { for (int qq = 0; qq < NJ; ++qq) t = fma(sh.Rvv[p * NJ + qq], sh.Nv[qq * NVMAX + cj], t); }
    RN[idx] = t;
  }
  __syncwarp();
  for (int idx = lane; idx < NTMAX * NTMAX; idx += 32) {
    const int i = idx / NTMAX, j = idx - i * NTMAX;
    double s = 0.0;
    if (i >= nt || j >= nt) s = (i == j) ? 1.0 / dt : 0.0;       // identity on the padded diagonal
    else if (i < nf && j < nf) s = sh.RFF[sh.stidx[i] * 12 + sh.stidx[j]];
    else if (i >= nf && j >= nf) {
      const int ci = i - nf, cj = j - nf;
      for (int p = 0; p < NJ; ++p) s = fma(sh.Nv[p * NVMAX + ci], RN[p * NVMAX + cj], s);
    }
    out[PJ_RT + i * NTMAX + j] = dt * s;
  }
  if (lane < NX) {
    // Pt is stored transposed (22 x NTMAX) so that the Riccati kernel can stage it with 16-byte async copies
    for (int c = 0; c < NTMAX; ++c) {
      double s = 0.0;
      if (c >= nf && c < nt) { const int cc = c - nf; for (int kk = 0; kk < NJ; ++kk) s = fma(sh.Nv[kk * NVMAX + cc], T1[kk * NX + lane], s); }
      out[PJ_PT + lane * NTMAX + c] = dt * s;
    }
  }
  if (lane < ntp) {
    double s;
    if (lane >= nt) s = 0.0;
    else if (lane < nf) s = sh.r[sh.stidx[lane]];
    else { s = 0.0; const int cc = lane - nf; for (int kk = 0; kk < NJ; ++kk) s = fma(sh.Nv[kk * NVMAX + cc], sh.rRpe[kk], s); }
    out[PJ_RV + lane] = dt * s;
  }
  // ---- write the rest of the record
  for (int idx = lane; idx < TS; idx += 32) {
    const int i = idx / NX, j = idx - i * NX;
    double v;
    if (i < 3) v = (i == j) ? 1.0 : 0.0;                                   // momentum rows: identity
    else if (i < 12) v = sh.Ad[idx - 3 * NX];
    else v = ((i == j) ? 1.0 : 0.0) + dt * Pxv[(i - 12) * NX + j];         // joint rows: I + dt Pxv
    out[PJ_AT + idx] = v;
  }
  for (int idx = lane; idx < NX * NTMAX; idx += 32) out[PJ_BT + idx] = Bt[idx];
  if (lane < NX) out[PJ_BTV + lane] = sh.bt[lane];
  for (int idx = lane; idx < NJ * NX; idx += 32) out[PJ_PXV + idx] = Pxv[idx];
  for (int idx = lane; idx < NJ * NVMAX; idx += 32) { const int c = idx % NVMAX; out[PJ_NV + idx] = (c < nv) ? sh.Nv[idx] : 0.0; }
  if (lane < NJ) out[PJ_PEV + lane] = sh.pev[lane];
  if (lane == 0) { out[PJ_META] = nt; out[PJ_META + 1] = nf; out[PJ_META + 2] = nv; out[PJ_META + 3] = cost; out[PJ_META + 4] = d2; out[PJ_META + 5] = e2; out[PJ_META + 6] = overflow ? 1.0 : 0.0; }
}

// ---------------------------------------------------------------- K2: value-function recursion
// Row-owner products: lane i owns row i of the result, the right operand is streamed from shared memory with 128-bit
// broadcast loads (two doubles per load), N accumulators stay in registers. NTP = number of free inputs padded to an even
// compile-time size (6 flight, 10 single support, 12 stance, 16 degenerate); the padded rows/columns of the projected model
// are zero (identity on the diagonal of R~), so the padded gains are zero.
template <int N, bool TA, int MODE>
__device__ __forceinline__ void rowmm(double* __restrict__ C, int ldc, const double* __restrict__ A, int lda,
                                      const double* __restrict__ B, int ldb, int m, int kdim) {
  const int i = lane_id();
  if (i < m) {
    double c[N];
#pragma unroll
    for (int j = 0; j < N; ++j) c[j] = (MODE == 1) ? C[i * ldc + j] : 0.0;
#pragma unroll 2
    for (int k = 0; k < kdim; ++k) {
      const double a = TA ? A[k * lda + i] : A[i * lda + k];
      const double* br = B + k * ldb;
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        const double2 b2 = *reinterpret_cast<const double2*>(br + j);
        c[j] = fma(a, b2.x, c[j]);
        c[j + 1] = fma(a, b2.y, c[j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) C[i * ldc + j] = c[j];
  }
  __syncwarp();
}


struct RicNodeIn { double At[TS], Bt[NX * NTMAX], bt[NX], qt[NX], rt[NTMAX], meta[8]; };
struct RicShared {
  double S[TS], SA[TS];
  RicNodeIn in[2];                                   // node data, staged one node ahead with cp.async
  double SBK[NX * NTMAX];                            // SB (22 x NTMAX), later K (NTMAX x 22)
  double HuxT[NX * NTMAX], Huu[NTMAX * 18];
  double sv[NX], sb[NX], hu[NTMAX], kff[NTMAX], idg[NTMAX];
};

__device__ __forceinline__ void ric_prefetch(RicNodeIn& n, const double* __restrict__ rec, int lane) {
  for (int i = 2 * lane; i < TS; i += 64) cp_async16(n.At + i, rec + PJ_AT + i);
  for (int i = 2 * lane; i < NX * NTMAX; i += 64) cp_async16(n.Bt + i, rec + PJ_BT + i);
  if (lane < 11) { cp_async16(n.bt + 2 * lane, rec + PJ_BTV + 2 * lane); cp_async16(n.qt + 2 * lane, rec + PJ_QV + 2 * lane); }
  else if (lane < 19) cp_async16(n.rt + 2 * (lane - 11), rec + PJ_RV + 2 * (lane - 11));
  else if (lane < 23) cp_async16(n.meta + 2 * (lane - 19), rec + PJ_META + 2 * (lane - 19));
  cp_async_commit();
}

// One node of the recursion, executed by the TWO warps of the block. The products that do not depend on each other are split
// between the warps (by result columns, so that every row-owner product keeps its full lane utilisation); the Cholesky of Huu and
// the gain solve (one warp, latency bound) overlap with the largest product At' S At of the other warp.
template <int NTP>
__device__ __noinline__ void riccati_node(RicShared& sh, const RicNodeIn& in, const double* __restrict__ rec, double* __restrict__ rk, bool& fail,
                                          int warp) {
  const int lane = lane_id();
  double* SB = sh.SBK; double* K = sh.SBK;
  // ---- phase A: [SA | SB | sb] = S [At | Bt | bt] (+ s): 22 + NTP + 1 result columns, split 16 / rest
  if (warp == 0) {
    rowmm<16, false, 0>(sh.SA, NX, sh.S, NX, in.At, NX, NX, NX);
    cp_async_wait<0>();                       // Pt^T / Rt staged by this warp at the top of the node
  } else {
    rowmm<6, false, 0>(sh.SA + 16, NX, sh.S, NX, in.At + 16, NX, NX, NX);
    rowmm<NTP, false, 0>(SB, NTMAX, sh.S, NX, in.Bt, NTMAX, NX, NX);
    if (lane < NX) {
      double s0 = sh.sv[lane], s1 = 0.0;
#pragma unroll
      for (int k = 0; k < NX; k += 2) { s0 = fma(sh.S[lane * NX + k], in.bt[k], s0); s1 = fma(sh.S[lane * NX + k + 1], in.bt[k + 1], s1); }
      sh.sb[lane] = s0 + s1;
    }
  }
  __syncthreads();
  // ---- phase B: Hux^T (22 x NTP) = Pt^T + SA^T Bt (warp 0) ; Huu = Rt + Bt^T SB, hu = rt + Bt^T sb (warp 1)
  if (warp == 0) {
    rowmm<NTP, true, 1>(sh.HuxT, NTMAX, sh.SA, NX, in.Bt, NTMAX, NX, NX);
  } else {
    // S is dead until phase D: stage Qt into it now (arrives while Huu is formed)
    for (int i = 2 * lane; i < TS; i += 64) cp_async16(sh.S + i, rec + PJ_QT + i);
    cp_async_commit();
    rowmm<NTP, true, 1>(sh.Huu, 18, in.Bt, NTMAX, SB, NTMAX, NTP, NX);
    if (lane < NTP) {
      double s0 = in.rt[lane];
#pragma unroll
      for (int k = 0; k < NX; ++k) s0 = fma(in.Bt[k * NTMAX + lane], sh.sb[k], s0);
      sh.hu[lane] = s0;
    }
  }
  __syncthreads();
  // ---- phase C: gains (warp 0) || S = Qt + At' SA (warp 1)
  if (warp == 0) {
    // Cholesky of the symmetrised Huu entirely in registers: lane i owns row i (right-looking, column by column, the pivot column is
    // broadcast with shuffles), then forward / backward substitution of the 22 + 1 right-hand sides, one per lane, with the factor
    // entries fetched from their owner lanes. No shared-memory round trips on this latency-bound stretch.
    {
      double a[NTP], rinv[NTP];
#pragma unroll
      for (int c = 0; c < NTP; ++c) a[c] = (lane < NTP && c <= lane) ? 0.5 * (sh.Huu[lane * 18 + c] + sh.Huu[c * 18 + lane]) : 0.0;
#pragma unroll
      for (int j = 0; j < NTP; ++j) {
        double d = __shfl_sync(HB_FULL_MASK, a[j], j);
        if (!(d > 0.0)) { fail = true; d = 1.0; }
        const double r = 1.0 / sqrt(d);
        rinv[j] = r;
        const double l = a[j] * r;              // L[i][j] on lane i >= j
        a[j] = l;
#pragma unroll
        for (int k = j + 1; k < NTP; ++k) a[k] = fma(-l, __shfl_sync(HB_FULL_MASK, l, k), a[k]);
      }
      double col[NTP], y[NTP];
#pragma unroll
      for (int c = 0; c < NTP; ++c) col[c] = (lane < NX) ? sh.HuxT[lane * NTMAX + c] : ((lane == NX) ? sh.hu[c] : 0.0);
#pragma unroll
      for (int c = 0; c < NTP; ++c) {
        double sacc = col[c];
#pragma unroll
        for (int kk = 0; kk < c; ++kk) sacc = fma(-__shfl_sync(HB_FULL_MASK, a[kk], c), y[kk], sacc);
        y[c] = sacc * rinv[c];
      }
#pragma unroll
      for (int c = NTP - 1; c >= 0; --c) {
        double sacc = y[c];
#pragma unroll
        for (int kk = c + 1; kk < NTP; ++kk) sacc = fma(-__shfl_sync(HB_FULL_MASK, a[c], kk), col[kk], sacc);
        col[c] = sacc * rinv[c];                 // solution of Huu x = rhs; the gain is its negative
      }
      if (lane < NX) {
#pragma unroll
        for (int c = 0; c < NTP; ++c) { K[c * NX + lane] = -col[c]; rk[c * NX + lane] = -col[c]; }     // SB is dead: K takes its place
      } else if (lane == NX) {
#pragma unroll
        for (int c = 0; c < NTP; ++c) { sh.kff[c] = -col[c]; rk[NTMAX * NX + c] = -col[c]; }
      }
    }
    __syncwarp();
    // s <- qt + At' sb + Hux' kff
    if (lane < NX) {
      double s0 = in.qt[lane], s1 = 0.0;
#pragma unroll
      for (int k = 0; k < NX; k += 2) { s0 = fma(in.At[k * NX + lane], sh.sb[k], s0); s1 = fma(in.At[(k + 1) * NX + lane], sh.sb[k + 1], s1); }
#pragma unroll
      for (int c = 0; c < NTP; ++c) s0 = fma(sh.HuxT[lane * NTMAX + c], sh.kff[c], s0);
      sh.sv[lane] = s0 + s1;
    }
  } else {
    cp_async_wait<0>();   // Qt has landed in S (and the next node's inputs, issued by this warp at the top of the node)
    __syncwarp();
    rowmm<NX, true, 1>(sh.S, NX, in.At, NX, sh.SA, NX, NX, NX);
  }
  __syncthreads();
  // ---- phase D: S += Hux' K, result columns split 12 / 10
  if (warp == 0) rowmm<12, false, 1>(sh.S, NX, sh.HuxT, NTMAX, K, NX, NX, NTP);
  else rowmm<10, false, 1>(sh.S + 12, NX, sh.HuxT, NTMAX, K + 12, NX, NX, NTP);
  __syncthreads();
  for (int idx = threadIdx.x; idx < TS; idx += 64) { const int i = idx / NX, j = idx - i * NX; if (j > i) { const double v = 0.5 * (sh.S[idx] + sh.S[j * NX + i]); sh.S[idx] = v; sh.S[j * NX + i] = v; } }
  __syncthreads();
}

__global__ void __launch_bounds__(64) riccati_kernel(SqpArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RicShared& sh = *reinterpret_cast<RicShared*>(smem_raw);
  const int inst = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, N = a.N;
  const double* proj = a.proj + (size_t)inst * N * PJ_STRIDE;
  // warp 1 owns the node-input prefetch (and the Qt staging), warp 0 the Pt^T / Rt staging; each waits for its own groups
  if (warp == 1) ric_prefetch(sh.in[(N - 1) & 1], proj + (size_t)(N - 1) * PJ_STRIDE, lane);
  for (int idx = threadIdx.x; idx < TS; idx += 64) sh.S[idx] = 0.0;   // no terminal cost
  if (threadIdx.x < NX) sh.sv[threadIdx.x] = 0.0;
  bool fail = false;
  double merit = 0.0, dyn = 0.0, eqs = 0.0;
  for (int k = N - 1; k >= 0; --k) {
    const double* rec = proj + (size_t)k * PJ_STRIDE;
    double* rk = a.rk + ((size_t)inst * N + k) * RK_STRIDE;
    if (warp == 1) cp_async_wait<0>();        // inputs of node k (prefetched one node ahead)
    __syncthreads();
    if (warp == 0) {
      // stage Pt^T -> HuxT (22 x 16) and Rt -> Huu (16 rows of 16, leading dimension 18) for this node
      for (int i = 2 * lane; i < NX * NTMAX; i += 64) cp_async16(sh.HuxT + i, rec + PJ_PT + i);
      for (int c = lane; c < NTMAX * 8; c += 32) { const int r = c >> 3, q = c & 7; cp_async16(sh.Huu + r * 18 + 2 * q, rec + PJ_RT + r * NTMAX + 2 * q); }
      cp_async_commit();
    } else if (k > 0) {
      ric_prefetch(sh.in[(k - 1) & 1], proj + (size_t)(k - 1) * PJ_STRIDE, lane);
    }
    const RicNodeIn& in = sh.in[k & 1];
    const int nt = (int)in.meta[0];
    merit += a.dt * in.meta[3]; dyn += a.dt * in.meta[4]; eqs += a.dt * in.meta[5];
    if (in.meta[6] != 0.0) fail = true;
    const int ntp = ntp_of(nt);
    if (ntp == 12) riccati_node<12>(sh, in, rec, rk, fail, warp);
    else if (ntp == 10) riccati_node<10>(sh, in, rec, rk, fail, warp);
    else if (ntp == 6) riccati_node<6>(sh, in, rec, rk, fail, warp);
    else riccati_node<16>(sh, in, rec, rk, fail, warp);
  }
  if (threadIdx.x == 0) {
    double* pf = a.perf + (size_t)inst * 4;
    pf[0] = merit; pf[1] = dyn; pf[2] = eqs; pf[3] = 0.0;
    a.flags[inst] = fail ? 1 : 0;
  }
}

// ---------------------------------------------------------------- K3: forward pass + filter line search
// per-node data of the forward pass, double-buffered and filled with cp.async (16-byte LDGSTS) one node ahead
struct FwNode {
  double At[TS], Bt[NX * NTMAX], K[NTMAX * NX], Pxv[NJ * NX], Nv[NJ * NVMAX];
  double bt[NX], qt[NX], kff[NTMAX], rt[NTMAX], pev[NJ], meta[8], u[NU];
};
struct Fw2Shared {
  FwNode nd[2];
  double dx[NX], dxn[NX], w[NTMAX];
};

__device__ __forceinline__ void fw_prefetch(FwNode& n, const double* __restrict__ rec, const double* __restrict__ rk, const double* __restrict__ uk, int lane) {
  auto copy = [&](double* dst, const double* src, int ndbl) { for (int i = 2 * lane; i < ndbl; i += 64) cp_async16(dst + i, src + i); };
  copy(n.At, rec + PJ_AT, TS);
  copy(n.Bt, rec + PJ_BT, NX * NTMAX);
  copy(n.K, rk, NTMAX * NX);
  copy(n.Pxv, rec + PJ_PXV, NJ * NX);
  copy(n.Nv, rec + PJ_NV, NJ * NVMAX);
  copy(n.bt, rec + PJ_BTV, NX);
  copy(n.qt, rec + PJ_QV, NX);
  copy(n.kff, rk + NTMAX * NX, NTMAX);
  copy(n.rt, rec + PJ_RV, NTMAX);
  copy(n.pev, rec + PJ_PEV, NJ);
  copy(n.meta, rec + PJ_META, 8);
  copy(n.u, uk, NU);
  cp_async_commit();
}

__global__ void __launch_bounds__(32) forward_linesearch2_kernel(SqpArgs a, int max_trials, void* info_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Fw2Shared& sh = *reinterpret_cast<Fw2Shared*>(smem_raw);
  const int inst = blockIdx.x, lane = threadIdx.x, N = a.N;
  const double dt = a.dt;
  double* xt = a.xt + (size_t)inst * (N + 1) * NX;
  double* ut = a.ut + (size_t)inst * N * NU;
  double* dxt = a.dxt + (size_t)inst * (N + 1) * NX;
  double* dut = a.dut + (size_t)inst * N * NU;
  const double* xref = a.x_ref + (size_t)inst * (N + 1) * NX;
  const double* swing = a.swing + (size_t)inst * (N + 1) * 24;
  const int32_t* mode = a.mode + (size_t)inst * (N + 1);
  const double* proj = a.proj + (size_t)inst * N * PJ_STRIDE;
  const double* rkb = a.rk + (size_t)inst * N * RK_STRIDE;
  fw_prefetch(sh.nd[0], proj, rkb, ut, lane);
  if (lane < NX) { xt[lane] = a.x0[(size_t)inst * NX + lane]; sh.dx[lane] = 0.0; dxt[lane] = 0.0; }
  double armijo = 0.0;
  bool finite = (a.flags[inst] == 0);
  for (int k = 0; k < N; ++k) {
    if (k + 1 < N) { fw_prefetch(sh.nd[(k + 1) & 1], proj + (size_t)(k + 1) * PJ_STRIDE, rkb + (size_t)(k + 1) * RK_STRIDE, ut + (k + 1) * NU, lane); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncwarp();
    const FwNode& nd = sh.nd[k & 1];
    const int nt = (int)nd.meta[0], nf = (int)nd.meta[1], nv = (int)nd.meta[2];
    const int md_k = mode[k];
    double arm = 0.0;
    if (lane < nt) {
      double s = nd.kff[lane];
#pragma unroll
      for (int j = 0; j < NX; ++j) s = fma(nd.K[lane * NX + j], sh.dx[j], s);
      sh.w[lane] = s;
      arm = nd.rt[lane] * s;
    }
    __syncwarp();
    if (lane < NX) {
      double s = nd.bt[lane], s1 = 0.0;
#pragma unroll
      for (int j = 0; j < NX; j += 2) { s = fma(nd.At[lane * NX + j], sh.dx[j], s); s1 = fma(nd.At[lane * NX + j + 1], sh.dx[j + 1], s1); }
      s += s1;
      for (int c = 0; c < nt; ++c) s = fma(nd.Bt[lane * NTMAX + c], sh.w[c], s);
      sh.dxn[lane] = s;
      dxt[(k + 1) * NX + lane] = s;
      arm += nd.qt[lane] * sh.dx[lane];
      // du: stance forces are free variables, swing forces go to zero, vj from the projection
      double du;
      if (lane < 12) {
        if (contact_flag(md_k, lane / 3)) { int c = 0; for (int j = 0; j < lane; ++j) c += contact_flag(md_k, j / 3); du = sh.w[c]; }
        else du = -nd.u[lane];
      } else {
        const int i = lane - 12;
        du = nd.pev[i];
#pragma unroll
        for (int j = 0; j < NX; ++j) du = fma(nd.Pxv[i * NX + j], sh.dx[j], du);
        for (int c = 0; c < nv; ++c) du = fma(nd.Nv[i * NVMAX + c], sh.w[nf + c], du);
      }
      dut[k * NU + lane] = du;
      if (!isfinite(s) || !isfinite(du)) finite = false;
    }
    armijo += warp_sum(arm);
    __syncwarp();
    if (lane < NX) sh.dx[lane] = sh.dxn[lane];
    __syncwarp();
  }
  finite = __all_sync(HB_FULL_MASK, finite);
  const double* pf = a.perf + (size_t)inst * 4;
  const double merit0 = pf[0], v0 = sqrt(pf[1] + pf[2]);
  const double gamma_c = 1e-6, armijoFactor = 1e-4, alpha_decay = 0.5, alpha_min = 1e-4;
  double alpha = 1.0, merit1 = merit0, v1 = v0;
  bool accepted = false;
  int trials = 0;
  if (finite) {
    while (alpha >= alpha_min && trials < max_trials) {
      double ms = 0.0, ds = 0.0, es = 0.0;
      for (int k = lane; k < N; k += 32) {
        double x[NX], u[NU], xn[NX], f1[NX], f2[NX], x2[NX], ep[12], ev[12], xr[NX], sw[24];
        for (int i = 0; i < NX; ++i) { x[i] = xt[k * NX + i] + alpha * dxt[k * NX + i]; xn[i] = xt[(k + 1) * NX + i] + alpha * dxt[(k + 1) * NX + i]; xr[i] = xref[k * NX + i]; }
        for (int i = 0; i < NU; ++i) u[i] = ut[k * NU + i] + alpha * dut[k * NU + i];
        for (int i = 0; i < 24; ++i) sw[i] = swing[k * 24 + i];
        flow_map_lane(x, u, f1, ep, ev);
        for (int i = 0; i < NX; ++i) x2[i] = x[i] + dt * f1[i];
        flow_map_lane(x2, u, f2, nullptr, nullptr);
        double d2 = 0.0;
        for (int i = 0; i < NX; ++i) { const double d = x[i] + 0.5 * dt * (f1[i] + f2[i]) - xn[i]; d2 += d * d; }
        double cost, e2;
        node_values_lane(x, u, xr, sw, mode[k], ep, ev, cost, e2);
        ms += dt * cost; ds += dt * d2; es += dt * e2;
      }
      ms = warp_sum(ms); ds = warp_sum(ds); es = warp_sum(es);
      ++trials;
      const double vn = sqrt(ds + es);
      const double am = alpha * armijo;
      bool acc;
      if (vn > HB_SQP_G_MAX) acc = vn < (1.0 - gamma_c) * v0;
      else if (vn < HB_SQP_G_MIN && v0 < HB_SQP_G_MIN && am < 0.0) acc = ms < merit0 + armijoFactor * am;
      else acc = ms < (merit0 - gamma_c * v0) || vn < (1.0 - gamma_c) * v0;
      if (isfinite(ms) && isfinite(vn) && acc) { accepted = true; merit1 = ms; v1 = vn; break; }
      alpha *= alpha_decay;
    }
  }
  if (accepted) {
    // accepted step: x += alpha dx, u += alpha du; four independent loads in flight per lane (one warp streams 70 KB here)
    auto axpy = [&](double* y, const double* d, int n) {
      int idx = lane;
      for (; idx + 96 < n; idx += 128) {
        const double y0 = y[idx], y1 = y[idx + 32], y2 = y[idx + 64], y3 = y[idx + 96];
        const double d0 = d[idx], d1 = d[idx + 32], d2 = d[idx + 64], d3 = d[idx + 96];
        y[idx] = fma(alpha, d0, y0); y[idx + 32] = fma(alpha, d1, y1); y[idx + 64] = fma(alpha, d2, y2); y[idx + 96] = fma(alpha, d3, y3);
      }
      for (; idx < n; idx += 32) y[idx] = fma(alpha, d[idx], y[idx]);
    };
    axpy(xt, dxt, (N + 1) * NX);
    axpy(ut, dut, N * NU);
  }
  if (lane == 0 && info_out) {
    struct Info { double alpha, merit0, merit1, viol0, viol1, armijo; int32_t status, n_trials; };
    Info* io = reinterpret_cast<Info*>(info_out) + inst;
    io->alpha = accepted ? alpha : 0.0; io->merit0 = merit0; io->merit1 = merit1; io->viol0 = v0; io->viol1 = v1;
    io->armijo = armijo; io->status = finite ? 0 : 3; io->n_trials = trials;
  }
}

}  // namespace hb
