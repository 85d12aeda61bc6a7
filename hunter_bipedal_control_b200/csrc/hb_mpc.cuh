// Lane-level pieces of the SQP iteration shared by the node-parallel pipeline in hb_sqp.cuh (what ocs2::SqpSolver::run does for
// sqpIteration = 1, legged_controllers/src/LeggedController.cpp:378-379,406):
//   penalties (relaxed barrier, double sided), and the VALUE-only evaluations the filter line search runs with one lane per
//   horizon node: the centroidal flow map (flow_map_lane) and the stage cost / equality-constraint values (node_values_lane).
#pragma once
#include "hb_common.cuh"
#include "hb_rbd.cuh"

namespace hb {

constexpr int TS = NX * NX;        // one 22x22 tile
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


}  // namespace hb
