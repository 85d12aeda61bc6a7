// ORACLE (test infrastructure only; never linked into the product path).
// Rigid-body algorithms for the fixed Hunter tree, restating what the reference obtains from Pinocchio
// (un-vendored; call sites legged_wbc/src/WbcBase.cpp:85-116, legged_interface/src/LeggedRobotPreComputation.cpp:158-170)
// and from ocs2_centroidal_model (un-vendored; call sites WbcBase.cpp:125-135, LeggedRobotDynamicsAD.cpp:57-71).
// Generalised coordinates: q = [p(3), yaw, pitch, roll, q_j(10)], v = q_dot.
// The floating base is treated as three world-axis prismatic joints followed by revolute z, y, x joints,
// so q_dot = v holds exactly as for Pinocchio's (Translation, SphericalZYX) composite joint.
// All quantities are expressed in world axes. Pinned against MuJoCo 3.0.1 (tests/test_oracle_rbd.py).
#pragma once
#include "../include/hunter_model_constants.h"
#include "hb_dual.hpp"

namespace hbo {

template <class T> inline void cross3(const T* a, const T* b, T* c) {
  T c0 = a[1] * b[2] - a[2] * b[1];
  T c1 = a[2] * b[0] - a[0] * b[2];
  T c2 = a[0] * b[1] - a[1] * b[0];
  c[0] = c0; c[1] = c1; c[2] = c2;
}
template <class T> inline T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> inline void matvec3(const T* R, const T* x, T* y) {
  T y0 = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
  T y1 = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
  T y2 = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
  y[0] = y0; y[1] = y1; y[2] = y2;
}
template <class T> inline void matTvec3(const T* R, const T* x, T* y) {
  T y0 = R[0] * x[0] + R[3] * x[1] + R[6] * x[2];
  T y1 = R[1] * x[0] + R[4] * x[1] + R[7] * x[2];
  T y2 = R[2] * x[0] + R[5] * x[1] + R[8] * x[2];
  y[0] = y0; y[1] = y1; y[2] = y2;
}
template <class T> inline void matmul3(const T* A, const T* B, T* C) {
  T r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  for (int i = 0; i < 9; ++i) C[i] = r[i];
}
// Rotation about a signed coordinate axis (ax has exactly one entry of +-1) by angle th.
template <class T> inline void axis_rotation(const double* ax, const T& th, T* R) {
  T c = cos(th), s = sin(th);
  for (int i = 0; i < 9; ++i) R[i] = T(0.0);
  if (ax[0] != 0.0) { T ss = s * ax[0]; R[0] = T(1.0); R[4] = c; R[5] = -ss; R[7] = ss; R[8] = c; }
  else if (ax[1] != 0.0) { T ss = s * ax[1]; R[4] = T(1.0); R[0] = c; R[2] = ss; R[6] = -ss; R[8] = c; }
  else { T ss = s * ax[2]; R[8] = T(1.0); R[0] = c; R[1] = -ss; R[3] = ss; R[4] = c; }
}

template <class T>
struct Kin {
  T R[HB_NBODY][9];   // world <- body rotation
  T p[HB_NBODY][3];   // body origin in world
  T ax[HB_NQ][3];     // world axis of generalised velocity k (k<3 translation, 3..5 euler z,y,x, 6.. joints)
};

// rotation matrix of ZYX Euler angles (yaw, pitch, roll): R = Rz Ry Rx
template <class T> inline void euler_zyx_to_R(const T* e, T* R) {
  T cz = cos(e[0]), sz = sin(e[0]), cy = cos(e[1]), sy = sin(e[1]), cx = cos(e[2]), sx = sin(e[2]);
  R[0] = cz * cy; R[1] = cz * sy * sx - sz * cx; R[2] = cz * sy * cx + sz * sx;
  R[3] = sz * cy; R[4] = sz * sy * sx + cz * cx; R[5] = sz * sy * cx - cz * sx;
  R[6] = -sy;     R[7] = cy * sx;                R[8] = cy * cx;
}

template <class T>
void forward_kinematics(const T* q, Kin<T>& k) {
  for (int i = 0; i < 3; ++i) {
    k.p[0][i] = q[i];
    for (int j = 0; j < 3; ++j) k.ax[i][j] = T(i == j ? 1.0 : 0.0);
  }
  euler_zyx_to_R(q + 3, k.R[0]);
  T cz = cos(q[3]), sz = sin(q[3]), cy = cos(q[4]), sy = sin(q[4]);
  k.ax[3][0] = T(0.0); k.ax[3][1] = T(0.0); k.ax[3][2] = T(1.0);      // yaw about world z
  k.ax[4][0] = -sz; k.ax[4][1] = cz; k.ax[4][2] = T(0.0);            // pitch about Rz e_y
  k.ax[5][0] = cz * cy; k.ax[5][1] = sz * cy; k.ax[5][2] = -sy;      // roll about Rz Ry e_x
  for (int b = 1; b < HB_NBODY; ++b) {
    const int P = HB_PARENT[b];
    T off[3] = {T(HB_JOINT_XYZ[3 * b]), T(HB_JOINT_XYZ[3 * b + 1]), T(HB_JOINT_XYZ[3 * b + 2])};
    T o[3];
    matvec3(k.R[P], off, o);
    for (int i = 0; i < 3; ++i) k.p[b][i] = k.p[P][i] + o[i];
    T Rj[9];
    axis_rotation(&HB_JOINT_AXIS[3 * b], q[5 + b], Rj);
    matmul3(k.R[P], Rj, k.R[b]);
    T a[3] = {T(HB_JOINT_AXIS[3 * b]), T(HB_JOINT_AXIS[3 * b + 1]), T(HB_JOINT_AXIS[3 * b + 2])};
    matvec3(k.R[b], a, k.ax[5 + b]);
  }
}

template <class T>
struct Vel {
  T w[HB_NBODY][3];  // angular velocity (world)
  T v[HB_NBODY][3];  // linear velocity of body origin (world)
};

template <class T>
void velocities(const Kin<T>& k, const T* vg, Vel<T>& u) {
  for (int i = 0; i < 3; ++i) {
    u.v[0][i] = vg[i];
    u.w[0][i] = k.ax[3][i] * vg[3] + k.ax[4][i] * vg[4] + k.ax[5][i] * vg[5];
  }
  for (int b = 1; b < HB_NBODY; ++b) {
    const int P = HB_PARENT[b];
    T d[3], wxd[3];
    for (int i = 0; i < 3; ++i) d[i] = k.p[b][i] - k.p[P][i];
    cross3(u.w[P], d, wxd);
    for (int i = 0; i < 3; ++i) {
      u.v[b][i] = u.v[P][i] + wxd[i];
      u.w[b][i] = u.w[P][i] + k.ax[5 + b][i] * vg[5 + b];
    }
  }
}

template <class T> void contact_position(const Kin<T>& k, int c, T* r) {
  const int b = HB_CONTACT_BODY[c];
  T off[3] = {T(HB_CONTACT_OFFSET[3 * c]), T(HB_CONTACT_OFFSET[3 * c + 1]), T(HB_CONTACT_OFFSET[3 * c + 2])};
  T o[3];
  matvec3(k.R[b], off, o);
  for (int i = 0; i < 3; ++i) r[i] = k.p[b][i] + o[i];
}
template <class T> void contact_velocity(const Kin<T>& k, const Vel<T>& u, int c, T* v) {
  const int b = HB_CONTACT_BODY[c];
  T off[3] = {T(HB_CONTACT_OFFSET[3 * c]), T(HB_CONTACT_OFFSET[3 * c + 1]), T(HB_CONTACT_OFFSET[3 * c + 2])};
  T o[3], wxo[3];
  matvec3(k.R[b], off, o);
  cross3(u.w[b], o, wxo);
  for (int i = 0; i < 3; ++i) v[i] = u.v[b][i] + wxo[i];
}

template <class T> void center_of_mass(const Kin<T>& k, T* c) {
  c[0] = c[1] = c[2] = T(0.0);
  for (int b = 0; b < HB_NBODY; ++b) {
    T cb[3] = {T(HB_BODY_COM[3 * b]), T(HB_BODY_COM[3 * b + 1]), T(HB_BODY_COM[3 * b + 2])};
    T o[3];
    matvec3(k.R[b], cb, o);
    for (int i = 0; i < 3; ++i) c[i] += (k.p[b][i] + o[i]) * (HB_BODY_MASS[b] / HB_TOTAL_MASS);
  }
}

// Centroidal momentum h = [linear; angular about the CoM], world axes: h = A(q) v.
template <class T> void centroidal_momentum(const Kin<T>& k, const Vel<T>& u, const T* com, T* h) {
  for (int i = 0; i < 6; ++i) h[i] = T(0.0);
  for (int b = 0; b < HB_NBODY; ++b) {
    T cb[3] = {T(HB_BODY_COM[3 * b]), T(HB_BODY_COM[3 * b + 1]), T(HB_BODY_COM[3 * b + 2])};
    T r[3], wxr[3], vc[3], rc[3], l[3], Iw[3], wl[3], Iwl[3];
    matvec3(k.R[b], cb, r);
    cross3(u.w[b], r, wxr);
    for (int i = 0; i < 3; ++i) { vc[i] = (u.v[b][i] + wxr[i]) * HB_BODY_MASS[b]; rc[i] = k.p[b][i] + r[i] - com[i]; }
    cross3(rc, vc, l);
    matTvec3(k.R[b], u.w[b], wl);
    for (int i = 0; i < 3; ++i)
      Iwl[i] = wl[0] * HB_BODY_INERTIA[9 * b + 3 * i] + wl[1] * HB_BODY_INERTIA[9 * b + 3 * i + 1] + wl[2] * HB_BODY_INERTIA[9 * b + 3 * i + 2];
    matvec3(k.R[b], Iwl, Iw);
    for (int i = 0; i < 3; ++i) { h[i] += vc[i]; h[3 + i] += l[i] + Iw[i]; }
  }
}

// Centroidal momentum matrix A(q) (6x16, row-major), column k = momentum of unit generalised velocity k.
template <class T> void centroidal_matrix(const Kin<T>& k, const T* com, T* A) {
  for (int c = 0; c < HB_NQ; ++c) {
    T e[HB_NQ];
    for (int i = 0; i < HB_NQ; ++i) e[i] = T(i == c ? 1.0 : 0.0);
    Vel<T> u;
    velocities(k, e, u);
    T h[6];
    centroidal_momentum(k, u, com, h);
    for (int r = 0; r < 6; ++r) A[r * HB_NQ + c] = h[r];
  }
}

// Solve the 6x6 system Ab y = rhs by Gaussian elimination with partial pivoting (pivot chosen on values).
template <class T> void solve6(const T* Ab /*6x16 row-major, uses first 6 cols*/, const T* rhs, T* y) {
  T M[6][7];
  for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) M[i][j] = Ab[i * HB_NQ + j]; M[i][6] = rhs[i]; }
  for (int c = 0; c < 6; ++c) {
    int piv = c; double best = std::fabs(value_of(M[c][c]));
    for (int r = c + 1; r < 6; ++r) { double a = std::fabs(value_of(M[r][c])); if (a > best) { best = a; piv = r; } }
    if (piv != c) for (int j = 0; j < 7; ++j) { T t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
    for (int r = c + 1; r < 6; ++r) {
      T f = M[r][c] / M[c][c];
      for (int j = c; j < 7; ++j) M[r][j] = M[r][j] - f * M[c][j];
    }
  }
  for (int r = 5; r >= 0; --r) {
    T s = M[r][6];
    for (int j = r + 1; j < 6; ++j) s = s - M[r][j] * y[j];
    y[r] = s / M[r][r];
  }
}

// Inverse dynamics tau = M(q) a + C(q,v) v + g(q) in the coordinates above (recursive Newton-Euler, world frame).
template <class T>
void rnea(const T* q, const T* vg, const T* ag, bool gravity, T* tau, const Kin<T>* kin_in = nullptr) {
  Kin<T> kl;
  if (!kin_in) forward_kinematics(q, kl);
  const Kin<T>& k = kin_in ? *kin_in : kl;
  Vel<T> u;
  velocities(k, vg, u);
  T wd[HB_NBODY][3], pd[HB_NBODY][3];  // angular acceleration, classical acceleration of the body origin
  {
    // base chain: yaw, pitch, roll about the same point
    T w1[3], w2[3], t1[3], t2[3];
    for (int i = 0; i < 3; ++i) { w1[i] = k.ax[3][i] * vg[3]; w2[i] = w1[i] + k.ax[4][i] * vg[4]; }
    cross3(w1, k.ax[4], t1);
    cross3(w2, k.ax[5], t2);
    for (int i = 0; i < 3; ++i) {
      wd[0][i] = k.ax[3][i] * ag[3] + k.ax[4][i] * ag[4] + t1[i] * vg[4] + k.ax[5][i] * ag[5] + t2[i] * vg[5];
      pd[0][i] = ag[i];
    }
  }
  for (int b = 1; b < HB_NBODY; ++b) {
    const int P = HB_PARENT[b];
    T d[3], t[3], t2[3], t3[3], t4[3];
    for (int i = 0; i < 3; ++i) d[i] = k.p[b][i] - k.p[P][i];
    cross3(wd[P], d, t);
    cross3(u.w[P], d, t2);
    cross3(u.w[P], t2, t3);
    cross3(u.w[P], k.ax[5 + b], t4);
    for (int i = 0; i < 3; ++i) {
      pd[b][i] = pd[P][i] + t[i] + t3[i];
      wd[b][i] = wd[P][i] + k.ax[5 + b][i] * ag[5 + b] + t4[i] * vg[5 + b];
    }
  }
  T f[HB_NBODY][3], n[HB_NBODY][3];  // net force, net moment about the body origin (this body + subtree)
  for (int b = 0; b < HB_NBODY; ++b) {
    T cb[3] = {T(HB_BODY_COM[3 * b]), T(HB_BODY_COM[3 * b + 1]), T(HB_BODY_COM[3 * b + 2])};
    T r[3], t[3], t2[3], t3[3], F[3], wl[3], wdl[3], Iw[3], Iwd[3], Nl[3], N[3], rxF[3];
    matvec3(k.R[b], cb, r);
    cross3(wd[b], r, t);
    cross3(u.w[b], r, t2);
    cross3(u.w[b], t2, t3);
    for (int i = 0; i < 3; ++i) F[i] = (pd[b][i] + t[i] + t3[i]) * HB_BODY_MASS[b];
    if (gravity) F[2] = F[2] + T(HB_BODY_MASS[b] * HB_GRAVITY);
    matTvec3(k.R[b], u.w[b], wl);
    matTvec3(k.R[b], wd[b], wdl);
    for (int i = 0; i < 3; ++i) {
      Iw[i] = wl[0] * HB_BODY_INERTIA[9 * b + 3 * i] + wl[1] * HB_BODY_INERTIA[9 * b + 3 * i + 1] + wl[2] * HB_BODY_INERTIA[9 * b + 3 * i + 2];
      Iwd[i] = wdl[0] * HB_BODY_INERTIA[9 * b + 3 * i] + wdl[1] * HB_BODY_INERTIA[9 * b + 3 * i + 1] + wdl[2] * HB_BODY_INERTIA[9 * b + 3 * i + 2];
    }
    cross3(wl, Iw, Nl);
    for (int i = 0; i < 3; ++i) Nl[i] = Nl[i] + Iwd[i];
    matvec3(k.R[b], Nl, N);
    cross3(r, F, rxF);
    for (int i = 0; i < 3; ++i) { f[b][i] = F[i]; n[b][i] = N[i] + rxF[i]; }
  }
  for (int b = HB_NBODY - 1; b >= 1; --b) {
    const int P = HB_PARENT[b];
    T d[3], dxf[3];
    for (int i = 0; i < 3; ++i) d[i] = k.p[b][i] - k.p[P][i];
    cross3(d, f[b], dxf);
    for (int i = 0; i < 3; ++i) { f[P][i] += f[b][i]; n[P][i] += n[b][i] + dxf[i]; }
    tau[5 + b] = dot3(k.ax[5 + b], n[b]);
  }
  for (int i = 0; i < 3; ++i) { tau[i] = f[0][i]; tau[3 + i] = dot3(k.ax[3 + i], n[0]); }
}

}  // namespace hbo
