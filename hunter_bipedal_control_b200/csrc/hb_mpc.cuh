// K1-K4: one SQP iteration of the Hunter centroidal NMPC (what ocs2::SqpSolver::run does for sqpIteration = 1,
// legged_controllers/src/LeggedController.cpp:378-379,406), one warp per instance.
//
//   backward kernel  k = N-1 .. 0 : LQ model of node k (centroidal dynamics + EE kinematics by lane-parallel unit /
//                    dual passes, RK2 sensitivities, cost, soft constraints, equality constraints), least-squares
//                    projection of the state-input equalities, one Riccati step. Only the closed-loop maps of the
//                    node (K_x, k_x, K_u, k_u) and a 22-vector for the Armijo metric leave the SM.
//   forward kernel   dx_{k+1} = K_x dx_k + k_x, du_k = K_u dx_k + k_u (sequential in k, 22x22 mat-vecs from shared memory)
//   line search      lanes = nodes: every lane integrates one shooting interval of the candidate (RK2), evaluates
//                    its constraints and costs; warp reductions give merit / violation; filter acceptance.
#pragma once
#include "hb_common.cuh"
#include "hb_rbd.cuh"
#include "hb_qp.cuh"   // warp_chol_inv

namespace hb {

constexpr int TS = NX * NX;        // one 22x22 tile
constexpr int GAIN_STRIDE = 1040;  // doubles stored per node: Kx(484) kx(22) Ku(484) ku(22) ga(22) a0(1) (+pad)
constexpr int NDIR = 13;           // non-trivial configuration directions: euler(3) + joints(10)

// ------------------------------------------------------------------ penalties
struct Pen { double v, d1, d2; };
__device__ __forceinline__ Pen relaxed_barrier(double h, double mu, double delta) {
  Pen p;
  if (h > delta) { p.v = -mu * log(h); p.d1 = -mu / h; p.d2 = mu / (h * h); }
  else {
    const double z = (h - 2.0 * delta) / delta;
    p.v = mu * (-log(delta) + 0.5 * z * z - 0.5);
    p.d1 = mu * (h - 2.0 * delta) / (delta * delta);
    p.d2 = mu / (delta * delta);
  }
  return p;
}
__device__ __forceinline__ Pen double_sided(double h, double lo, double hi, double mu, double delta) {
  const Pen a = relaxed_barrier(h - lo, mu, delta), b = relaxed_barrier(hi - h, mu, delta);
  Pen p; p.v = a.v + b.v; p.d1 = a.d1 - b.d1; p.d2 = a.d2 + b.d2;
  return p;
}

// ------------------------------------------------------------------ shared memory of the backward kernel
struct MpcShared {
  double S[TS], A1[TS], B1[TS];
  double A2[TS], B2[TS], T1[TS];   // contiguous: reused as the 22x45 normal-equation workspace of the projection
  double T2[TS];
  double Q[TS], R[TS], P[TS], PX[TS], PU[TS];
  double Cm[16 * NX], Dm[16 * NX];
  double EJ[3 * 12 * NX];          // EE Jacobians dpos_dx, dvel_dx, dvel_du (12x22 each)
  // kinematics scratch
  double Acm[6 * 16], Jc[12 * 16], dh[6 * NDIR], dcom[3 * NDIR], dp[12 * NDIR], dv[12 * NDIR];
  double Abinv[36], AbinvAj[6 * NJ], dvb[6 * NDIR];
  // vectors
  double x[NX], u[NX], xn[NX], x2[NX], f1[NX], f2[NX], b[NX], xref[NX], swing[24];
  double q[NX], r[NX], e[16], pe[NX], sv[NX], bt[NX], sb[NX], qt[NX], rt[NX], rRpe[NX], hu[NX], kff[NX], idg[NX];
  double epos[12], evel[12], com[3], vgen[16], gxy[8 * 2 * NX + 8];
  int freeidx[NX], pivflag[NX], rowc[16], rowa[16], rowt[16];
};

// C[m x n] (+)= op(A)[m x k] * B[k x n]; lanes own columns of C (n <= 32), m <= 22.
// TA: A is stored k x m (use A^T). MODE 0: C = AB, 1: C += AB, 2: C = -AB
template <bool TA, int MODE>
__device__ __forceinline__ void wmm(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int m, int k, int n) {
  const int j = lane_id();
  if (j < n) {
    double acc[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) acc[i] = 0.0;
    for (int kk = 0; kk < k; ++kk) {
      const double bv = B[kk * ldb + j];
#pragma unroll
      for (int i = 0; i < NX; ++i)
        if (i < m) acc[i] = fma(TA ? A[kk * lda + i] : A[i * lda + kk], bv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
      if (i < m) {
        if (MODE == 0) C[i * ldc + j] = acc[i];
        else if (MODE == 1) C[i * ldc + j] += acc[i];
        else C[i * ldc + j] = -acc[i];
      }
  }
  __syncwarp();
}
// y[m] (+)= op(A)[m x k] x[k]; lanes own rows
template <bool TA, int MODE>
__device__ __forceinline__ void wmv(double* y, const double* A, int lda, const double* x, int m, int k) {
  const int i = lane_id();
  if (i < m) {
    double a0 = 0.0, a1 = 0.0;
    int kk = 0;
    for (; kk + 1 < k; kk += 2) {
      a0 = fma(TA ? A[kk * lda + i] : A[i * lda + kk], x[kk], a0);
      a1 = fma(TA ? A[(kk + 1) * lda + i] : A[i * lda + kk + 1], x[kk + 1], a1);
    }
    if (kk < k) a0 = fma(TA ? A[kk * lda + i] : A[i * lda + kk], x[kk], a0);
    if (MODE == 0) y[i] = a0 + a1; else y[i] += a0 + a1;
  }
  __syncwarp();
}

// ------------------------------------------------------------------ centroidal flow map, one lane = one evaluation
// (used by the line search: lane-per-node). x = [hbar(6), q(16)], u = [F(12), qj_dot(10)]   
// Uses the block structure of the centroidal momentum matrix: A[:, 0:3] = [m I; 0] (computeFloatingBaseCentroidalMomentumMatrixInverse).
__device__ inline void flow_map_lane(const double* x, const double* u, double* f, double* epos, double* evel) {
  const double m = c_model.total_mass;
  double q[NQ], v[NQ];
  for (int i = 0; i < NQ; ++i) q[i] = x[6 + i];
  KinOut<double> o;
  for (int i = 0; i < 6; ++i) v[i] = 0.0;
  for (int j = 0; j < NJ; ++j) v[6 + j] = u[12 + j];
  kin_pass<double>(q, v, o);
  double hj[6], com[3], cp[12];
  for (int i = 0; i < 6; ++i) hj[i] = o.h[i];
  for (int i = 0; i < 3; ++i) com[i] = o.com[i];
  for (int i = 0; i < 12; ++i) cp[i] = o.cpos[i];
  double Ae[6][3];  // euler columns of A
  for (int c = 0; c < 3; ++c) {
    for (int i = 0; i < NQ; ++i) v[i] = (i == 3 + c) ? 1.0 : 0.0;
    kin_pass<double>(q, v, o);
    for (int r = 0; r < 6; ++r) Ae[r][c] = o.h[r];
  }
  double rl[3], ra[3];
  for (int i = 0; i < 3; ++i) { rl[i] = m * x[i] - hj[i]; ra[i] = m * x[3 + i] - hj[3 + i]; }
  // 3x3 solve A22 th = ra (Cramer)
  const double a = Ae[3][0], b = Ae[3][1], c = Ae[3][2], d = Ae[4][0], e = Ae[4][1], g = Ae[4][2], h = Ae[5][0], k = Ae[5][1], l = Ae[5][2];
  const double det = a * (e * l - g * k) - b * (d * l - g * h) + c * (d * k - e * h);
  const double id = 1.0 / det;
  double th[3];
  th[0] = ((e * l - g * k) * ra[0] - (b * l - c * k) * ra[1] + (b * g - c * e) * ra[2]) * id;
  th[1] = (-(d * l - g * h) * ra[0] + (a * l - c * h) * ra[1] - (a * g - c * d) * ra[2]) * id;
  th[2] = ((d * k - e * h) * ra[0] - (a * k - b * h) * ra[1] + (a * e - b * d) * ra[2]) * id;
  double vb[6];
  for (int i = 0; i < 3; ++i) { vb[i] = (rl[i] - Ae[i][0] * th[0] - Ae[i][1] * th[1] - Ae[i][2] * th[2]) / m; vb[3 + i] = th[i]; }
  if (evel) {
    for (int i = 0; i < 6; ++i) v[i] = vb[i];
    for (int j = 0; j < NJ; ++j) v[6 + j] = u[12 + j];
    kin_pass<double>(q, v, o);
    for (int i = 0; i < 12; ++i) evel[i] = o.cvel[i];
  }
  if (epos) for (int i = 0; i < 12; ++i) epos[i] = cp[i];
  const double im = 1.0 / m;
  double fl[3] = {0, 0, 0}, fa[3] = {0, 0, 0};
  for (int cc = 0; cc < NC; ++cc) {
    const double* F = u + 3 * cc;
    const double r0 = cp[3 * cc] - com[0], r1 = cp[3 * cc + 1] - com[1], r2 = cp[3 * cc + 2] - com[2];
    fl[0] += F[0]; fl[1] += F[1]; fl[2] += F[2];
    fa[0] += r1 * F[2] - r2 * F[1]; fa[1] += r2 * F[0] - r0 * F[2]; fa[2] += r0 * F[1] - r1 * F[0];
  }
  for (int i = 0; i < 3; ++i) { f[i] = fl[i] * im; f[3 + i] = fa[i] * im; }
  f[2] -= HB_GRAVITY;
  for (int i = 0; i < 6; ++i) f[6 + i] = vb[i];
  for (int j = 0; j < NJ; ++j) f[12 + j] = u[12 + j];
}

// Stage cost (unscaled) and equality-constraint values of one node, one lane = one node (values only).
__device__ inline void node_values_lane(const double* x, const double* u, const double* xref, const double* swing, int mode,
                                        const double* epos, const double* evel, double& cost, double& eq_sq) {
  const Model& md = c_model;
  bool fl[4]; int ns = 0;
  for (int c = 0; c < 4; ++c) { fl[c] = contact_flag(mode, c); ns += fl[c]; }
  double cst = 0.0;
  for (int i = 0; i < NX; ++i) { const double d = x[i] - xref[i]; cst += 0.5 * md.Q[i] * d * d; }
  double du[NU];
  const double fz = ns > 0 ? md.total_mass * HB_GRAVITY / ns : 0.0;
  for (int i = 0; i < NU; ++i) du[i] = u[i];
  for (int c = 0; c < 4; ++c) if (fl[c]) du[3 * c + 2] -= fz;
  for (int i = 0; i < NU; ++i) { double s = 0.0; for (int j = 0; j < NU; ++j) s += md.R[i * NU + j] * du[j]; cst += 0.5 * du[i] * s; }
  double e2 = 0.0;
  for (int c = 0; c < 4; ++c) {
    if (fl[c]) {
      const double Fx = u[3 * c], Fy = u[3 * c + 1], Fz = u[3 * c + 2];
      const double h = HB_FRICTION_MU * Fz - sqrt(Fx * Fx + Fy * Fy + HB_FRICTION_REGULARIZATION);
      cst += relaxed_barrier(h, HB_FRICTION_BARRIER_MU, HB_FRICTION_BARRIER_DELTA).v;
      const double e0 = evel[3 * c], e1 = evel[3 * c + 1], e2z = evel[3 * c + 2] + HB_ZEROVEL_Z_GAIN * epos[3 * c + 2] + HB_ZEROVEL_Z_OFFSET;
      e2 += e0 * e0 + e1 * e1 + e2z * e2z;
    } else {
      for (int a = 0; a < 2; ++a) {
        const double h = evel[3 * c + a] - swing[6 * c + 3 + a] + HB_XY_POSITION_GAIN * (epos[3 * c + a] - swing[6 * c + a]);
        cst += 0.5 * HB_SOFT_SWING_WEIGHT * h * h;
      }
      const double en = evel[3 * c + 2] - swing[6 * c + 5] + HB_POSITION_ERROR_GAIN * (epos[3 * c + 2] - swing[6 * c + 2]);
      e2 += u[3 * c] * u[3 * c] + u[3 * c + 1] * u[3 * c + 1] + u[3 * c + 2] * u[3 * c + 2] + en * en;
    }
  }
  for (int j = 0; j < NJ; ++j) {
    cst += double_sided(x[12 + j], md.joint_lower[j], md.joint_upper[j], HB_LIMIT_POS_MU, HB_LIMIT_POS_DELTA).v;
    cst += double_sided(u[12 + j], -md.joint_vel_limit[j], md.joint_vel_limit[j], HB_LIMIT_VEL_MU, HB_LIMIT_VEL_DELTA).v;
  }
  for (int c = 0; c < 4; ++c) cst += double_sided(u[3 * c + 2], 0.0, HB_LIMIT_FORCE_MAX, HB_LIMIT_FORCE_MU, HB_LIMIT_FORCE_DELTA).v;
  cost = cst; eq_sq = e2;
}

// ------------------------------------------------------------------ warp-cooperative linearisation of the flow map at (sh.x|x2, sh.u)
// Writes f (22) and the tiles Aout, Bout (d f/dx, d f/du). With want_ee also the contact kinematics and their Jacobians (sh.EJ).
__device__ inline void flow_lin_warp(MpcShared& sh, const double* xs, double* f, double* Aout, double* Bout, bool want_ee) {
  const int lane = lane_id();
  const double m = c_model.total_mass;
  // pass 1: unit generalised velocities -> columns of the centroidal momentum matrix and of the contact Jacobians
  {
    double q[NQ], e[NQ];
    const int k = lane & 15;
    for (int i = 0; i < NQ; ++i) { q[i] = xs[6 + i]; e[i] = (i == k) ? 1.0 : 0.0; }
    KinOut<double> o;
    kin_pass<double>(q, e, o);
    if (lane < 16) {
      for (int r = 0; r < 6; ++r) sh.Acm[r * 16 + k] = o.h[r];
      for (int r = 0; r < 12; ++r) sh.Jc[r * 16 + k] = o.cvel[r];
    }
    if (lane == 0) { for (int r = 0; r < 12; ++r) sh.epos[r] = o.cpos[r]; for (int r = 0; r < 3; ++r) sh.com[r] = o.com[r]; }
  }
  __syncwarp();
  // A_b^-1 (lanes 0-5, one column each) and v_b (lane 6)
  if (lane < 7) {
    double Ab[36], rhs[6], y[6];
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) Ab[6 * r + c] = sh.Acm[r * 16 + c];
    if (lane < 6) { for (int r = 0; r < 6; ++r) rhs[r] = (r == lane) ? 1.0 : 0.0; }
    else {
      for (int r = 0; r < 6; ++r) { double s = m * xs[r]; for (int j = 0; j < NJ; ++j) s -= sh.Acm[r * 16 + 6 + j] * sh.u[12 + j]; rhs[r] = s; }
    }
    solve6(Ab, rhs, y);
    if (lane < 6) { for (int r = 0; r < 6; ++r) sh.Abinv[6 * r + lane] = y[r]; }
    else { for (int r = 0; r < 6; ++r) sh.vgen[r] = y[r]; }
  }
  if (lane >= 8 && lane < 18) sh.vgen[6 + lane - 8] = sh.u[12 + lane - 8];
  __syncwarp();
  if (lane < NJ) {
    for (int r = 0; r < 6; ++r) { double s = 0.0; for (int c = 0; c < 6; ++c) s += sh.Abinv[6 * r + c] * sh.Acm[c * 16 + 6 + lane]; sh.AbinvAj[r * NJ + lane] = s; }
  }
  // pass 2: dual numbers seeded along configuration direction 3+lane, generalised velocity held fixed
  if (lane < NDIR) {
    D1 q[NQ], v[NQ];
    for (int i = 0; i < NQ; ++i) { q[i] = D1(xs[6 + i], (i == 3 + lane) ? 1.0 : 0.0); v[i] = D1(sh.vgen[i], 0.0); }
    KinOut<D1> o;
    kin_pass<D1>(q, v, o);
    for (int r = 0; r < 6; ++r) sh.dh[r * NDIR + lane] = o.h[r].d;
    for (int r = 0; r < 3; ++r) sh.dcom[r * NDIR + lane] = o.com[r].d;
    for (int r = 0; r < 12; ++r) { sh.dp[r * NDIR + lane] = o.cpos[r].d; sh.dv[r * NDIR + lane] = o.cvel[r].d; }
    if (lane == 0) for (int r = 0; r < 12; ++r) sh.evel[r] = o.cvel[r].v;
  }
  __syncwarp();
  if (lane < NDIR) {
    for (int r = 0; r < 6; ++r) { double s = 0.0; for (int c = 0; c < 6; ++c) s += sh.Abinv[6 * r + c] * sh.dh[c * NDIR + lane]; sh.dvb[r * NDIR + lane] = -s; }
  }
  // flow map value
  if (lane < 3) {
    double s = 0.0;
    for (int c = 0; c < NC; ++c) s += sh.u[3 * c + lane];
    f[lane] = s / m - (lane == 2 ? HB_GRAVITY : 0.0);
  } else if (lane < 6) {
    const int a = lane - 3, a1 = (a + 1) % 3, a2 = (a + 2) % 3;
    double s = 0.0;
    for (int c = 0; c < NC; ++c) s += (sh.epos[3 * c + a1] - sh.com[a1]) * sh.u[3 * c + a2] - (sh.epos[3 * c + a2] - sh.com[a2]) * sh.u[3 * c + a1];
    f[lane] = s / m;
  } else if (lane < NX) {
    f[lane] = sh.vgen[lane - 6];
  }
  __syncwarp();
  // Jacobian columns: lane j owns column j of A and of B
  if (lane < NX) {
    const int j = lane;
    double ca[NX], cb[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) { ca[i] = 0.0; cb[i] = 0.0; }
    if (j < 6) {
      for (int r = 0; r < 6; ++r) ca[6 + r] = m * sh.Abinv[6 * r + j];
    } else if (j >= 9) {
      const int k = j - 9;
      double t[3] = {0, 0, 0};
      for (int c = 0; c < NC; ++c) {
        const double d0 = sh.dp[(3 * c) * NDIR + k] - sh.dcom[k], d1 = sh.dp[(3 * c + 1) * NDIR + k] - sh.dcom[NDIR + k], d2 = sh.dp[(3 * c + 2) * NDIR + k] - sh.dcom[2 * NDIR + k];
        const double* F = sh.u + 3 * c;
        t[0] += d1 * F[2] - d2 * F[1]; t[1] += d2 * F[0] - d0 * F[2]; t[2] += d0 * F[1] - d1 * F[0];
      }
      for (int r = 0; r < 3; ++r) ca[3 + r] = t[r] / m;
      for (int r = 0; r < 6; ++r) ca[6 + r] = sh.dvb[r * NDIR + k];
    }
    if (j < 12) {
      const int c = j / 3, a = j - 3 * c;
      cb[a] = 1.0 / m;
      // (r - com) x e_a
      const double r0 = sh.epos[3 * c] - sh.com[0], r1 = sh.epos[3 * c + 1] - sh.com[1], r2 = sh.epos[3 * c + 2] - sh.com[2];
      if (a == 0) { cb[4] = r2 / m; cb[5] = -r1 / m; }
      else if (a == 1) { cb[3] = -r2 / m; cb[5] = r0 / m; }
      else { cb[3] = r1 / m; cb[4] = -r0 / m; }
    } else {
      const int jj = j - 12;
      for (int r = 0; r < 6; ++r) cb[6 + r] = -sh.AbinvAj[r * NJ + jj];
      cb[12 + jj] = 1.0;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) { Aout[i * NX + j] = ca[i]; Bout[i * NX + j] = cb[i]; }
    if (want_ee) {
      double* dpos_dx = sh.EJ; double* dvel_dx = sh.EJ + 12 * NX; double* dvel_du = sh.EJ + 24 * NX;
      for (int r = 0; r < 12; ++r) {
        double px = 0.0, vx = 0.0, vu = 0.0;
        if (j < 6) { for (int c = 0; c < 6; ++c) vx += sh.Jc[r * 16 + c] * sh.Abinv[6 * c + j]; vx *= m; }
        else if (j < 9) { px = ((r % 3) == (j - 6)) ? 1.0 : 0.0; }
        else { const int k = j - 9; px = sh.dp[r * NDIR + k]; vx = sh.dv[r * NDIR + k]; for (int c = 0; c < 6; ++c) vx += sh.Jc[r * 16 + c] * sh.dvb[c * NDIR + k]; }
        if (j >= 12) { const int jj = j - 12; vu = sh.Jc[r * 16 + 6 + jj]; for (int c = 0; c < 6; ++c) vu -= sh.Jc[r * 16 + c] * sh.AbinvAj[c * NJ + jj]; }
        dpos_dx[r * NX + j] = px; dvel_dx[r * NX + j] = vx; dvel_du[r * NX + j] = vu;
      }
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------ cost + constraints LQ model of the node held in sh (M2-M8)
// Fills Q, R, P (u x x), q, r (unscaled), Cm, Dm, e, returns m; cost value through cost_out (all lanes).
__device__ inline int node_cost_constraints_warp(MpcShared& sh, int mode, double& cost_out) {
  const int lane = lane_id();
  const Model& md = c_model;
  const double* dpos_dx = sh.EJ; const double* dvel_dx = sh.EJ + 12 * NX; const double* dvel_du = sh.EJ + 24 * NX;
  bool fl[4]; int ns = 0;
  for (int c = 0; c < 4; ++c) { fl[c] = contact_flag(mode, c); ns += fl[c]; }
  const double fz = ns > 0 ? md.total_mass * HB_GRAVITY / ns : 0.0;
  // tiles: Q diag, R = cost R, P = 0
  for (int idx = lane; idx < TS; idx += 32) {
    const int i = idx / NX, j = idx - i * NX;
    sh.Q[idx] = (i == j) ? md.Q[i] : 0.0;
    sh.R[idx] = md.R[idx];
    sh.P[idx] = 0.0;
  }
  double cost = 0.0;
  if (lane < NX) {
    const double d = sh.x[lane] - sh.xref[lane];
    sh.q[lane] = md.Q[lane] * d;
    cost += 0.5 * md.Q[lane] * d * d;
    double s = 0.0;
    for (int j = 0; j < NU; ++j) {
      double duj = sh.u[j];
      if (j < 12 && (j % 3) == 2 && fl[j / 3]) duj -= fz;
      s += md.R[lane * NU + j] * duj;
    }
    double dul = sh.u[lane];
    if (lane < 12 && (lane % 3) == 2 && fl[lane / 3]) dul -= fz;
    sh.r[lane] = s;
    cost += 0.5 * dul * s;
  }
  __syncwarp();
  // limits (M8): lanes 0-9 joint position, 10-19 joint velocity, 20-23 normal force; friction cone (M6): lanes 24-27
  double shiftsum = 0.0;  // sum over stance contacts of p'(h) * (-hessianDiagonalShift)
  if (lane < 10) {
    const Pen p = double_sided(sh.x[12 + lane], md.joint_lower[lane], md.joint_upper[lane], HB_LIMIT_POS_MU, HB_LIMIT_POS_DELTA);
    cost += p.v; sh.q[12 + lane] += p.d1; sh.Q[(12 + lane) * NX + 12 + lane] += p.d2;
  } else if (lane < 20) {
    const int j = lane - 10;
    const Pen p = double_sided(sh.u[12 + j], -md.joint_vel_limit[j], md.joint_vel_limit[j], HB_LIMIT_VEL_MU, HB_LIMIT_VEL_DELTA);
    cost += p.v; sh.r[12 + j] += p.d1; sh.R[(12 + j) * NU + 12 + j] += p.d2;
  }
  __syncwarp();
  if (lane >= 20 && lane < 24) {
    const int c = lane - 20;
    const Pen p = double_sided(sh.u[3 * c + 2], 0.0, HB_LIMIT_FORCE_MAX, HB_LIMIT_FORCE_MU, HB_LIMIT_FORCE_DELTA);
    cost += p.v; sh.r[3 * c + 2] += p.d1; sh.R[(3 * c + 2) * NU + 3 * c + 2] += p.d2;
  }
  __syncwarp();
  if (lane >= 24 && lane < 28) {
    const int c = lane - 24;
    if (fl[c]) {
      const double Fx = sh.u[3 * c], Fy = sh.u[3 * c + 1], Fz = sh.u[3 * c + 2];
      const double t2 = Fx * Fx + Fy * Fy + HB_FRICTION_REGULARIZATION, tn = sqrt(t2), t32 = tn * t2;
      const double h = HB_FRICTION_MU * Fz - tn;
      const Pen p = relaxed_barrier(h, HB_FRICTION_BARRIER_MU, HB_FRICTION_BARRIER_DELTA);
      cost += p.v;
      const double gr[3] = {-Fx / tn, -Fy / tn, HB_FRICTION_MU};
      const double Hh[9] = {-(Fy * Fy + HB_FRICTION_REGULARIZATION) / t32, Fx * Fy / t32, 0.0, Fx * Fy / t32,
                            -(Fx * Fx + HB_FRICTION_REGULARIZATION) / t32, 0.0, 0.0, 0.0, 0.0};
      for (int i = 0; i < 3; ++i) {
        sh.r[3 * c + i] += p.d1 * gr[i];
        for (int j = 0; j < 3; ++j) sh.R[(3 * c + i) * NU + 3 * c + j] += p.d2 * gr[i] * gr[j] + p.d1 * Hh[3 * i + j];
      }
      shiftsum = -p.d1 * HB_FRICTION_HESSIAN_SHIFT;
    }
  }
  shiftsum = warp_sum(shiftsum);
  __syncwarp();
  if (lane < NX) { sh.Q[lane * NX + lane] += shiftsum; sh.R[lane * NU + lane] += shiftsum; }
  __syncwarp();
  // xy swing reference soft constraint (M7): Gauss-Newton terms, lane j owns column j
  int npair = 0;
  for (int c = 0; c < 4; ++c) {
    if (fl[c]) continue;
    for (int a = 0; a < 2; ++a) {
      const int row = 3 * c + a;
      double* gx = sh.gxy + npair * 2 * NX; double* gu = gx + NX;
      if (lane < NX) { gx[lane] = dvel_dx[row * NX + lane] + HB_XY_POSITION_GAIN * dpos_dx[row * NX + lane]; gu[lane] = dvel_du[row * NX + lane]; }
      if (lane == 0) sh.gxy[16 * NX + npair] = sh.evel[row] - sh.swing[6 * c + 3 + a] + HB_XY_POSITION_GAIN * (sh.epos[row] - sh.swing[6 * c + a]);
      ++npair;
    }
  }
  __syncwarp();
  if (npair > 0) {
    const double w = HB_SOFT_SWING_WEIGHT;
    for (int pidx = 0; pidx < npair; ++pidx) {
      const double* gx = sh.gxy + pidx * 2 * NX; const double* gu = gx + NX;
      const double h = sh.gxy[16 * NX + pidx];
      if (lane == 0) cost += 0.5 * w * h * h;
      if (lane < NX) {
        const int j = lane;
        const double gxj = gx[j], guj = gu[j];
        sh.q[j] += w * h * gxj; sh.r[j] += w * h * guj;
        for (int i = 0; i < NX; ++i) {
          sh.Q[i * NX + j] += w * gx[i] * gxj;
          sh.R[i * NU + j] += w * gu[i] * guj;
          sh.P[i * NX + j] += w * gu[i] * gxj;
        }
      }
      __syncwarp();
    }
  }
  cost_out = warp_sum(cost);
  // equality constraints: row table (uniform)
  int mrows = 0;
  for (int c = 0; c < 4; ++c) {
    if (fl[c]) { for (int a = 0; a < 3; ++a) { if (lane == 0) { sh.rowc[mrows] = c; sh.rowa[mrows] = a; sh.rowt[mrows] = 0; } ++mrows; } }
    else {
      for (int a = 0; a < 3; ++a) { if (lane == 0) { sh.rowc[mrows] = c; sh.rowa[mrows] = a; sh.rowt[mrows] = 1; } ++mrows; }
      if (lane == 0) { sh.rowc[mrows] = c; sh.rowa[mrows] = 2; sh.rowt[mrows] = 2; } ++mrows;
    }
  }
  __syncwarp();
  if (lane < NX) {
    const int j = lane;
    for (int rr = 0; rr < mrows; ++rr) {
      const int c = sh.rowc[rr], a = sh.rowa[rr], t = sh.rowt[rr], row = 3 * c + a;
      double cv = 0.0, dvv = 0.0;
      if (t == 0) { cv = dvel_dx[row * NX + j] + (a == 2 ? HB_ZEROVEL_Z_GAIN * dpos_dx[row * NX + j] : 0.0); dvv = dvel_du[row * NX + j]; }
      else if (t == 1) { dvv = (j == row) ? 1.0 : 0.0; }
      else { cv = dvel_dx[row * NX + j] + HB_POSITION_ERROR_GAIN * dpos_dx[row * NX + j]; dvv = dvel_du[row * NX + j]; }
      sh.Cm[rr * NX + j] = cv; sh.Dm[rr * NX + j] = dvv;
    }
  }
  if (lane < mrows) {
    const int c = sh.rowc[lane], a = sh.rowa[lane], t = sh.rowt[lane], row = 3 * c + a;
    double ev;
    if (t == 0) ev = sh.evel[row] + (a == 2 ? HB_ZEROVEL_Z_GAIN * sh.epos[row] + HB_ZEROVEL_Z_OFFSET : 0.0);
    else if (t == 1) ev = sh.u[row];
    else ev = sh.evel[row] - sh.swing[6 * c + 5] + HB_POSITION_ERROR_GAIN * (sh.epos[row] - sh.swing[6 * c + 2]);
    sh.e[lane] = ev;
  }
  __syncwarp();
  return mrows;
}

// ------------------------------------------------------------------ least-squares projection of C dx + D du + e = 0 (S4)
// Normal equations D'D du = -D'(C dx + e), Gauss-Jordan with diagonal pivoting and rank threshold (D is rank deficient for
// Hunter: two point contacts on one rigid foot). Produces PX (22x22), PU (22 x nt, ld 22), pe; returns nt.
__device__ inline int project_warp(MpcShared& sh, int mrows) {
  const int lane = lane_id();
  constexpr int W = NU + NX + 1;  // 45
  double* G = sh.A2;              // 22 x 45 spans A2, B2, T1
  for (int j = lane; j < W; j += 32) {
    for (int i = 0; i < NU; ++i) {
      double s = 0.0;
      if (j < NU) { for (int r = 0; r < mrows; ++r) s += sh.Dm[r * NX + i] * sh.Dm[r * NX + j]; }
      else if (j < NU + NX) { for (int r = 0; r < mrows; ++r) s -= sh.Dm[r * NX + i] * sh.Cm[r * NX + j - NU]; }
      else { for (int r = 0; r < mrows; ++r) s -= sh.Dm[r * NX + i] * sh.e[r]; }
      G[i * W + j] = s;
    }
  }
  if (lane < NU) sh.pivflag[lane] = 0;
  __syncwarp();
  double dmax = lane < NU ? G[lane * W + lane] : 0.0;
  dmax = warp_max(dmax);
  const double tol = 1e-9 * fmax(dmax, 1e-300);
  for (int step = 0; step < NU; ++step) {
    double dv = (lane < NU && !sh.pivflag[lane]) ? G[lane * W + lane] : -1.0;
    int pi = lane;
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(HB_FULL_MASK, dv, o);
      const int oi = __shfl_xor_sync(HB_FULL_MASK, pi, o);
      if (ov > dv || (ov == dv && oi < pi)) { dv = ov; pi = oi; }
    }
    if (!(dv > tol)) break;
    const int p = pi;
    const double inv = 1.0 / dv;
    __syncwarp();
    for (int j = lane; j < W; j += 32) G[p * W + j] *= inv;
    if (lane == 0) sh.pivflag[p] = 1;
    __syncwarp();
    // eliminate column p from all other rows: lanes own columns, loop rows
    for (int j = lane; j < W; j += 32) {
      if (j == p) continue;
      const double gp = G[p * W + j];
      for (int i = 0; i < NU; ++i) if (i != p) G[i * W + j] -= G[i * W + p] * gp;
    }
    __syncwarp();
    if (lane < NU && lane != p) G[lane * W + p] = 0.0;
    __syncwarp();
  }
  int nt = 0;
  for (int i = 0; i < NU; ++i) if (!sh.pivflag[i]) { if (lane == 0) sh.freeidx[nt] = i; ++nt; }
  __syncwarp();
  if (lane < NX) {
    const int j = lane;
    for (int i = 0; i < NU; ++i) sh.PX[i * NX + j] = sh.pivflag[i] ? G[i * W + NU + j] : 0.0;
    sh.pe[j] = sh.pivflag[j] ? G[j * W + NU + NX] : 0.0;
    if (j < nt) {
      const int fc = sh.freeidx[j];
      for (int i = 0; i < NU; ++i) sh.PU[i * NX + j] = sh.pivflag[i] ? -G[i * W + fc] : ((i == fc) ? 1.0 : 0.0);
    }
  }
  __syncwarp();
  return nt;
}

// ------------------------------------------------------------------ backward kernel
struct MpcArgs {
  int B, N;
  double dt;
  const double* x_ref;   // B x (N+1) x 22
  const double* swing;   // B x (N+1) x 24
  const int32_t* mode;   // B x (N+1)
  double* xt;            // B x (N+1) x 22
  double* ut;            // B x N x 22
  double* gains;         // B x N x GAIN_STRIDE
  double* dxt;           // B x (N+1) x 22
  double* dut;           // B x N x 22
  double* perf;          // B x 4: merit0, dynSSE0, eqSSE0, armijo
  int32_t* flags;        // B: bit0 = numerical failure in the backward pass
  const double* x0;      // B x 22
};

__global__ void __launch_bounds__(32) mpc_backward_kernel(MpcArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MpcShared& sh = *reinterpret_cast<MpcShared*>(smem_raw);
  const int inst = blockIdx.x;
  if (inst >= a.B) return;
  const int lane = lane_id();
  const int N = a.N;
  const double dt = a.dt;
  double* xt = a.xt + (size_t)inst * (N + 1) * NX;
  double* ut = a.ut + (size_t)inst * N * NU;
  const double* xref = a.x_ref + (size_t)inst * (N + 1) * NX;
  const double* swing = a.swing + (size_t)inst * (N + 1) * 24;
  const int32_t* mode = a.mode + (size_t)inst * (N + 1);
  double* gains = a.gains + (size_t)inst * N * GAIN_STRIDE;
  // the first node is pinned to the measured state (SqpSolver initialises x[0] = initState)
  if (lane < NX) xt[lane] = a.x0[(size_t)inst * NX + lane];
  for (int idx = lane; idx < TS; idx += 32) sh.S[idx] = 0.0;   // no terminal cost
  if (lane < NX) sh.sv[lane] = 0.0;
  __syncwarp();
  double merit = 0.0, dyn = 0.0, eqs = 0.0;
  bool fail = false;
  for (int k = N - 1; k >= 0; --k) {
    if (lane < NX) {
      sh.x[lane] = xt[k * NX + lane]; sh.u[lane] = ut[k * NU + lane]; sh.xn[lane] = xt[(k + 1) * NX + lane];
      sh.xref[lane] = xref[k * NX + lane];
    }
    if (lane < 24) sh.swing[lane] = swing[k * 24 + lane];
    const int md_k = mode[k];
    __syncwarp();
    // ---- stage 1 of RK2 + contact kinematics
    flow_lin_warp(sh, sh.x, sh.f1, sh.A1, sh.B1, true);
    double cost;
    const int mrows = node_cost_constraints_warp(sh, md_k, cost);
    if (lane < NX) sh.x2[lane] = sh.x[lane] + dt * sh.f1[lane];
    __syncwarp();
    // ---- stage 2
    flow_lin_warp(sh, sh.x2, sh.f2, sh.A2, sh.B2, false);
    // ---- RK2 sensitivities (S2): Ad = I + dt/2 (A1 + A2 + dt A2 A1), Bd = dt/2 (B1 + B2 + dt A2 B1)
    wmm<false, 0>(sh.T1, NX, sh.A2, NX, sh.A1, NX, NX, NX, NX);
    wmm<false, 0>(sh.T2, NX, sh.A2, NX, sh.B1, NX, NX, NX, NU);
    for (int idx = lane; idx < TS; idx += 32) {
      const int i = idx / NX, j = idx - i * NX;
      sh.A1[idx] = ((i == j) ? 1.0 : 0.0) + 0.5 * dt * (sh.A1[idx] + sh.A2[idx] + dt * sh.T1[idx]);
      sh.B1[idx] = 0.5 * dt * (sh.B1[idx] + sh.B2[idx] + dt * sh.T2[idx]);
    }
    double d2 = 0.0, e2 = 0.0;
    if (lane < NX) { const double bb = sh.x[lane] + 0.5 * dt * (sh.f1[lane] + sh.f2[lane]) - sh.xn[lane]; sh.b[lane] = bb; d2 = bb * bb; }
    if (lane < mrows) e2 = sh.e[lane] * sh.e[lane];
    d2 = warp_sum(d2); e2 = warp_sum(e2);
    merit += dt * cost; dyn += dt * d2; eqs += dt * e2;
    __syncwarp();
    // ---- projection (S4)
    const int nt = project_warp(sh, mrows);
    // ---- scale the stage cost by dt (S3)
    for (int idx = lane; idx < TS; idx += 32) { sh.Q[idx] *= dt; sh.R[idx] *= dt; sh.P[idx] *= dt; }
    if (lane < NX) { sh.q[lane] *= dt; sh.r[lane] *= dt; }
    __syncwarp();
    // ---- projected model
    wmm<false, 0>(sh.T1, NX, sh.B1, NX, sh.PX, NX, NX, NU, NX);           // Bd Px
    for (int idx = lane; idx < TS; idx += 32) sh.A1[idx] += sh.T1[idx];   // At
    __syncwarp();
    wmm<false, 0>(sh.T2, NX, sh.B1, NX, sh.PU, NX, NX, NU, nt);           // Bt (22 x nt, ld 22)
    wmv<false, 0>(sh.bt, sh.B1, NX, sh.pe, NX, NU);
    if (lane < NX) sh.bt[lane] += sh.b[lane];
    wmm<false, 0>(sh.T1, NX, sh.R, NU, sh.PX, NX, NU, NU, NX);            // R Px
    for (int idx = lane; idx < TS; idx += 32) sh.T1[idx] += sh.P[idx];    // PRPx = P + R Px
    wmv<false, 0>(sh.rRpe, sh.R, NU, sh.pe, NU, NU);
    if (lane < NU) sh.rRpe[lane] += sh.r[lane];
    __syncwarp();
    wmm<true, 0>(sh.A2, NX, sh.PX, NX, sh.T1, NX, NX, NU, NX);            // Px' PRPx
    for (int idx = lane; idx < TS; idx += 32) sh.Q[idx] += sh.A2[idx];
    __syncwarp();
    wmm<true, 0>(sh.A2, NX, sh.P, NX, sh.PX, NX, NX, NU, NX);             // P' Px
    for (int idx = lane; idx < TS; idx += 32) sh.Q[idx] += sh.A2[idx];    // Qt
    wmv<true, 0>(sh.qt, sh.PX, NX, sh.rRpe, NX, NU);
    wmv<true, 1>(sh.qt, sh.P, NX, sh.pe, NX, NU);
    if (lane < NX) sh.qt[lane] += sh.q[lane];
    __syncwarp();
    wmm<true, 0>(sh.A2, NX, sh.PU, NX, sh.T1, NX, nt, NU, NX);            // Pt = Pu' PRPx   (nt x 22)
    wmm<false, 0>(sh.B1, NX, sh.R, NU, sh.PU, NX, NU, NU, nt);            // R Pu            (22 x nt)   [Bd is dead]
    wmm<true, 0>(sh.B2, NX, sh.PU, NX, sh.B1, NX, nt, NU, nt);            // Rt = Pu' R Pu   (nt x nt)
    wmv<true, 0>(sh.rt, sh.PU, NX, sh.rRpe, nt, NU);
    // ---- Riccati step (S5)
    wmm<false, 0>(sh.T1, NX, sh.S, NX, sh.A1, NX, NX, NX, NX);            // SA
    wmm<false, 0>(sh.B1, NX, sh.S, NX, sh.T2, NX, NX, NX, nt);            // SB (22 x nt)
    wmv<false, 0>(sh.sb, sh.S, NX, sh.bt, NX, NX);
    if (lane < NX) sh.sb[lane] += sh.sv[lane];
    __syncwarp();
    wmm<true, 1>(sh.A2, NX, sh.T2, NX, sh.T1, NX, nt, NX, NX);            // Hux = Pt + Bt' SA
    wmm<true, 1>(sh.B2, NX, sh.T2, NX, sh.B1, NX, nt, NX, nt);            // Huu = Rt + Bt' SB
    wmv<true, 0>(sh.hu, sh.T2, NX, sh.sb, nt, NX);
    if (lane < nt) sh.hu[lane] += sh.rt[lane];
    __syncwarp();
    // symmetrise Huu, factorise, K = -Huu^-1 Hux, kff = -Huu^-1 hu
    for (int idx = lane; idx < nt * nt; idx += 32) { const int i = idx / nt, j = idx - i * nt; if (j < i) sh.B2[i * NX + j] = 0.5 * (sh.B2[i * NX + j] + sh.B2[j * NX + i]); }
    __syncwarp();
    if (nt > 0) {
      if (!warp_chol_inv(sh.B2, nt, NX, sh.idg, lane)) fail = true;
      // columns of Hux (22) and hu (1): lane j solves its own column with Li, Li'
      if (lane <= NX) {
        double col[NU], y[NU];
        for (int i = 0; i < nt; ++i) col[i] = (lane < NX) ? sh.A2[i * NX + lane] : sh.hu[i];
        for (int i = 0; i < nt; ++i) { double s = sh.idg[i] * col[i]; for (int kk = 0; kk < i; ++kk) s += sh.B2[kk * NX + i] * col[kk]; y[i] = s; }
        for (int i = 0; i < nt; ++i) { double s = sh.idg[i] * y[i]; for (int kk = i + 1; kk < nt; ++kk) s += sh.B2[i * NX + kk] * y[kk]; col[i] = -s; }
        if (lane < NX) { for (int i = 0; i < nt; ++i) sh.B1[i * NX + lane] = col[i]; }   // K (nt x 22)
        else { for (int i = 0; i < nt; ++i) sh.kff[i] = col[i]; }
      }
    }
    __syncwarp();
    // S <- Qt + At' SA + Hux' K ; s <- qt + At' sb + Hux' kff
    wmm<true, 0>(sh.P, NX, sh.A1, NX, sh.T1, NX, NX, NX, NX);
    wmm<true, 1>(sh.P, NX, sh.A2, NX, sh.B1, NX, NX, nt, NX);
    wmv<true, 0>(sh.sv, sh.A1, NX, sh.sb, NX, NX);
    wmv<true, 1>(sh.sv, sh.A2, NX, sh.kff, NX, nt);
    if (lane < NX) sh.sv[lane] += sh.qt[lane];
    for (int idx = lane; idx < TS; idx += 32) sh.P[idx] += sh.Q[idx];
    __syncwarp();
    for (int idx = lane; idx < TS; idx += 32) { const int i = idx / NX, j = idx - i * NX; sh.S[idx] = 0.5 * (sh.P[idx] + sh.P[j * NX + i]); }
    // ---- closed loop in original coordinates: Ku = Px + Pu K, ku = pe + Pu kff, Kx = At + Bt K, kx = bt + Bt kff
    wmm<false, 1>(sh.PX, NX, sh.PU, NX, sh.B1, NX, NU, nt, NX);
    wmm<false, 1>(sh.A1, NX, sh.T2, NX, sh.B1, NX, NX, nt, NX);
    wmv<false, 1>(sh.pe, sh.PU, NX, sh.kff, NU, nt);
    wmv<false, 1>(sh.bt, sh.T2, NX, sh.kff, NX, nt);
    // Armijo pieces: ga = qt + K' rt ; a0 = rt' kff
    wmv<true, 1>(sh.qt, sh.B1, NX, sh.rt, NX, nt);
    double a0 = (lane < nt) ? sh.rt[lane] * sh.kff[lane] : 0.0;
    a0 = warp_sum(a0);
    double* gk = gains + (size_t)k * GAIN_STRIDE;
    for (int idx = lane; idx < TS; idx += 32) { gk[idx] = sh.A1[idx]; gk[TS + NX + idx] = sh.PX[idx]; }
    if (lane < NX) { gk[TS + lane] = sh.bt[lane]; gk[2 * TS + NX + lane] = sh.pe[lane]; gk[2 * TS + 2 * NX + lane] = sh.qt[lane]; }
    if (lane == 0) gk[2 * TS + 3 * NX] = a0;
    __syncwarp();
  }
  if (lane == 0) {
    double* pf = a.perf + (size_t)inst * 4;
    pf[0] = merit; pf[1] = dyn; pf[2] = eqs; pf[3] = 0.0;
    a.flags[inst] = fail ? 1 : 0;
  }
}

// ------------------------------------------------------------------ forward pass + filter line search (S5 forward, S6)
struct LsShared {
  double Kx[TS], Ku[TS];
  double kx[NX], ku[NX], ga[NX], dx[NX], dxn[NX], du[NX];
};

__global__ void __launch_bounds__(32) mpc_forward_linesearch_kernel(MpcArgs a, int max_trials, void* info_out /* hb_solve_info* */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LsShared& sh = *reinterpret_cast<LsShared*>(smem_raw);
  const int inst = blockIdx.x;
  if (inst >= a.B) return;
  const int lane = lane_id();
  const int N = a.N;
  const double dt = a.dt;
  double* xt = a.xt + (size_t)inst * (N + 1) * NX;
  double* ut = a.ut + (size_t)inst * N * NU;
  double* dxt = a.dxt + (size_t)inst * (N + 1) * NX;
  double* dut = a.dut + (size_t)inst * N * NU;
  const double* xref = a.x_ref + (size_t)inst * (N + 1) * NX;
  const double* swing = a.swing + (size_t)inst * (N + 1) * 24;
  const int32_t* mode = a.mode + (size_t)inst * (N + 1);
  const double* gains = a.gains + (size_t)inst * N * GAIN_STRIDE;
  // ---- forward pass
  if (lane < NX) { sh.dx[lane] = 0.0; dxt[lane] = 0.0; }
  __syncwarp();
  double armijo = 0.0;
  bool finite = (a.flags[inst] == 0);
  for (int k = 0; k < N; ++k) {
    const double* gk = gains + (size_t)k * GAIN_STRIDE;
    for (int idx = lane; idx < TS; idx += 32) { sh.Kx[idx] = gk[idx]; sh.Ku[idx] = gk[TS + NX + idx]; }
    if (lane < NX) { sh.kx[lane] = gk[TS + lane]; sh.ku[lane] = gk[2 * TS + NX + lane]; sh.ga[lane] = gk[2 * TS + 2 * NX + lane]; }
    const double a0 = gk[2 * TS + 3 * NX];
    __syncwarp();
    double arm = 0.0;
    if (lane < NX) {
      double s0 = sh.kx[lane], s1 = sh.ku[lane];
      for (int j = 0; j < NX; ++j) { const double d = sh.dx[j]; s0 = fma(sh.Kx[lane * NX + j], d, s0); s1 = fma(sh.Ku[lane * NX + j], d, s1); }
      sh.dxn[lane] = s0; sh.du[lane] = s1;
      dxt[(k + 1) * NX + lane] = s0; dut[k * NU + lane] = s1;
      arm = sh.ga[lane] * sh.dx[lane];
      if (!isfinite(s0) || !isfinite(s1)) finite = false;
    }
    armijo += warp_sum(arm) + a0;
    __syncwarp();
    if (lane < NX) sh.dx[lane] = sh.dxn[lane];
    __syncwarp();
  }
  finite = __all_sync(HB_FULL_MASK, finite);
  // ---- filter line search (FilterLinesearch::acceptStep)
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
    for (int idx = lane; idx < (N + 1) * NX; idx += 32) xt[idx] += alpha * dxt[idx];
    for (int idx = lane; idx < N * NU; idx += 32) ut[idx] += alpha * dut[idx];
  }
  if (lane == 0 && info_out) {
    struct Info { double alpha, merit0, merit1, viol0, viol1, armijo; int32_t status, n_trials; };
    Info* io = reinterpret_cast<Info*>(info_out) + inst;
    io->alpha = accepted ? alpha : 0.0; io->merit0 = merit0; io->merit1 = merit1; io->viol0 = v0; io->viol1 = v1;
    io->armijo = armijo; io->status = finite ? 0 : 3; io->n_trials = trials;
  }
}

}  // namespace hb
