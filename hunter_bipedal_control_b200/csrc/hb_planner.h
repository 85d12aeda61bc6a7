// Host-side reference preprocessing of one MPC solve, mirroring the reference's own objects:
//   P1  GaitSchedule::{insertModeSequenceTemplate, tileModeSequenceTemplate}      legged_interface/src/gait/GaitSchedule.cpp:57-161
//       gait templates                                                            legged_controllers/config/hunter/reference.info:54-118
//   P3  SwingTrajectoryPlanner::{update, calNextFootPos, genSwingTrajs}           legged_interface/src/foot_planner/SwingTrajectoryPlanner.cpp:164-358
//       CubicSpline nodes (time, position, velocity)                              legged_interface/src/foot_planner/CubicSpline.cpp:46-70
//   P5  cmdVelToTargetTrajectories / targetPoseToTargetTrajectories               legged_controllers/src/TargetTrajectoriesPublisher.cpp:41-130
// The output is the compact hb_reference consumed by hb_reference_expand_batch (device). P2 (speed-based gait switching) is a
// caller decision (the gait id is an input) and P4 (IK joint references) is "next" (targets carry the default joint angles).
#pragma once
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/hunter_b200.h"
#include "../../include/hunter_model_constants.h"

namespace hbplan {

struct Vec3 { double x, y, z; };
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// getRotationMatrixFromZyxEulerAngles applied to a vector
inline Vec3 rot_zyx(const double* e, Vec3 v) {
  const double cz = cos(e[0]), sz = sin(e[0]), cy = cos(e[1]), sy = sin(e[1]), cx = cos(e[2]), sx = sin(e[2]);
  const double R[9] = {cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx, sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx, -sy, cy * sx, cy * cx};
  return {R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z};
}

// MotionPhaseDefinition.h:55-87
inline bool contact_flag(int mode, int c) { return (c & 1) ? (mode == 1 || mode == 3) : (mode == 2 || mode == 3); }

struct ModeSchedule { std::vector<double> events; std::vector<int> modes; };

// reference.info:54-118
inline void gait_template(int gait, std::vector<int>& modes, std::vector<double>& times) {
  switch (gait) {
    case 1: modes = {2, 1}; times = {0.0, 0.3, 0.6}; break;                       // trot
    case 2: modes = {2, 3, 1, 3}; times = {0.0, 0.25, 0.3, 0.55, 0.6}; break;     // standing_trot
    case 3: modes = {2, 0, 1, 0}; times = {0.0, 0.15, 0.2, 0.35, 0.4}; break;     // flying_trot
    default: modes = {3}; times = {0.0, 0.5}; break;                              // stance
  }
}

// A schedule that is STANCE (two phases split at `prev_event`, like the reference's initialModeSchedule {STANCE, STANCE},
// reference.info:21-32) until `start`, then the template tiled up to `final_time` and closed by a STANCE phase:
// GaitSchedule.cpp:57-93 (insert; the last mode before insertion is STANCE, so no extra transition phase) and :123-161 (tile).
inline ModeSchedule tile_gait(int gait, double prev_event, double start, double final_time) {
  ModeSchedule ms;
  ms.modes.push_back(3);
  ms.events.push_back(prev_event);
  ms.modes.push_back(3);
  std::vector<int> tm; std::vector<double> tt;
  gait_template(gait, tm, tt);
  ms.events.push_back(start);
  while (ms.events.back() < final_time) {
    for (size_t i = 0; i < tm.size(); ++i) {
      ms.modes.push_back(tm[i]);
      ms.events.push_back(ms.events.back() + (tt[i + 1] - tt[i]));
    }
  }
  ms.modes.push_back(3);
  return ms;
}

// ModeSchedule::modeAtTime: lower_bound on the event times (an event time itself belongs to the earlier mode)
inline int mode_at(const ModeSchedule& ms, double t) {
  const size_t idx = std::lower_bound(ms.events.begin(), ms.events.end(), t) - ms.events.begin();
  return ms.modes[idx];
}

struct Target { double t[2]; double x[2][22]; };

// cmdVelToTargetTrajectories (TargetTrajectoriesPublisher.cpp:102-130) with targetPoseToTargetTrajectories (:41-62)
inline Target cmd_vel_to_target(const double* cmd /*vx,vy,vz,wz*/, double time, const double* state, double time_to_target) {
  const double* pose = state + 6;
  Vec3 v = rot_zyx(pose + 3, {cmd[0], cmd[1], cmd[2]});
  if (fabs(v.x) < 0.06) v.x = 0.0;
  else if (fabs(v.y) < 0.06) v.y = 0.0;
  double target[6] = {pose[0] + v.x * time_to_target, pose[1] + v.y * time_to_target, HB_COM_HEIGHT, pose[3] + cmd[3] * time_to_target, 0.0, 0.0};
  double cur[6] = {pose[0], pose[1], pose[2], pose[3], 0.0, 0.0};
  double dz = HB_COM_HEIGHT - pose[2];
  dz = dz > 0 ? fmin(dz, 0.04) : fmax(dz, -0.04);      // changeLimit_[2] (TargetTrajectoriesPublisher.h:97)
  cur[2] = pose[2] + dz;
  Target tg;
  tg.t[0] = time; tg.t[1] = time + time_to_target;
  for (int k = 0; k < 2; ++k) {
    for (int i = 0; i < 22; ++i) tg.x[k][i] = 0.0;
    for (int i = 0; i < 6; ++i) tg.x[k][6 + i] = (k == 0) ? cur[i] : target[i];
    for (int j = 0; j < 10; ++j) tg.x[k][12 + j] = HB_DEFAULT_JOINT_STATE[j];
    tg.x[k][0] = v.x; tg.x[k][1] = v.y; tg.x[k][2] = v.z;    // stateTrajectory[.].head(3) = cmdVelRot (:127-128)
  }
  return tg;
}

inline void target_state(const Target& tg, double t, double* x) {
  if (t <= tg.t[0]) { memcpy(x, tg.x[0], sizeof(double) * 22); return; }
  if (t >= tg.t[1]) { memcpy(x, tg.x[1], sizeof(double) * 22); return; }
  const double a = (t - tg.t[0]) / (tg.t[1] - tg.t[0]);
  for (int i = 0; i < 22; ++i) x[i] = (1.0 - a) * tg.x[0][i] + a * tg.x[1][i];
}

struct Node { double t, p, v; };
struct Seg { double t0, t1, p0, v0, p1, v1; };

struct SwingPlan { std::vector<Seg> seg[4][3]; };

// SwingTrajectoryPlanner::findIndex (SwingTrajectoryPlanner.cpp:394-419)
inline void find_index(size_t index, const std::vector<bool>& stock, int& start_idx, int& final_idx) {
  const int n = (int)stock.size();
  start_idx = 0;
  for (int ip = (int)index - 1; ip >= 0; --ip) if (stock[ip] != stock[index]) { start_idx = ip; break; }
  final_idx = n - 2;
  for (int ip = (int)index + 1; ip < n; ++ip) if (stock[ip] != stock[index]) { final_idx = ip - 1; break; }
}

// SwingTrajectoryPlanner::calNextFootPos (:289-312). body_vel_cmd = [vx, vy, vz, wz, 0, 0] as set from /cmd_vel_filtered
// (SwitchedModelReferenceManager.cpp:91-101): its tail(3) = (wz, 0, 0) is used as the commanded angular velocity, as the reference does.
inline Vec3 next_foot_pos(int foot, double current_time, double stop_time, double next_middle_time, const double* next_middle_body_pos,
                          const double* current_body_pos, Vec3 current_body_vel, const double* body_vel_cmd) {
  const Vec3 bias[4] = {{HB_FEET_BIAS_X1, HB_FEET_BIAS_Y, HB_FEET_BIAS_Z}, {HB_FEET_BIAS_X1, -HB_FEET_BIAS_Y, HB_FEET_BIAS_Z},
                        {HB_FEET_BIAS_X2, HB_FEET_BIAS_Y, HB_FEET_BIAS_Z}, {HB_FEET_BIAS_X2, -HB_FEET_BIAS_Y, HB_FEET_BIAS_Z}};
  const Vec3 roted_bias = rot_zyx(next_middle_body_pos + 3, bias[foot]);
  const Vec3 vel_cmd_linear = rot_zyx(current_body_pos + 3, {body_vel_cmd[0], body_vel_cmd[1], body_vel_cmd[2]});
  const Vec3 vel_cmd_angular = rot_zyx(current_body_pos + 3, {body_vel_cmd[3], body_vel_cmd[4], body_vel_cmd[5]});
  Vec3 vel_linear = current_body_vel; vel_linear.z = 0.0;
  const double k = 0.03;
  const Vec3 p_shoulder = (stop_time - current_time) * (0.5 * vel_linear + 0.5 * vel_cmd_linear) + roted_bias;
  const Vec3 p_symmetry = (next_middle_time - stop_time) * vel_linear + k * (vel_linear - vel_cmd_linear);
  const Vec3 p_centrifugal = (0.5 * sqrt(current_body_pos[2] / 9.81)) * cross(vel_linear, vel_cmd_angular);
  Vec3 r = Vec3{current_body_pos[0], current_body_pos[1], current_body_pos[2]} + p_shoulder + p_symmetry + p_centrifugal;
  r.z = HB_NEXT_POSITION_Z;
  return r;
}

// SwingTrajectoryPlanner::genSwingTrajs (:314-358): x/y three-node, z four-node Hermite splines with the reference's shape constants
inline void gen_swing(SwingPlan& sp, int foot, double t0, double t1, Vec3 a, Vec3 b) {
  const double xy_a1 = 0.417, xy_l1 = 0.650, xy_k1 = 1.770;
  const double pa[3] = {a.x, a.y, a.z}, pb[3] = {b.x, b.y, b.z};
  for (int ax = 0; ax < 2; ++ax) {
    const Node n0{t0, pa[ax], 0.0}, n1{(1 - xy_a1) * t0 + xy_a1 * t1, (1 - xy_l1) * pa[ax] + xy_l1 * pb[ax], xy_k1 * (pb[ax] - pa[ax]) / (t1 - t0)}, n2{t1, pb[ax], 0.0};
    sp.seg[foot][ax].push_back({n0.t, n1.t, n0.p, n0.v, n1.p, n1.v});
    sp.seg[foot][ax].push_back({n1.t, n2.t, n1.p, n1.v, n2.p, n2.v});
  }
  const double scaling = std::min(1.0, (t1 - t0) / HB_SWING_TIME_SCALE);
  const double max_z = std::max(a.z, b.z) + scaling * HB_SWING_HEIGHT;
  const double z_a1 = 0.251, z_l1 = 0.749, z_k1 = 1.338, z_a2 = 0.630, z_l2 = 0.570, z_k2 = 1.633, z_k3 = 0.0;
  const Node n0{t0, a.z, 0.0};
  const Node n1{(1 - z_a1) * t0 + z_a1 * t1, z_l1 * max_z, z_k1 * (z_l1 * (max_z - a.z)) / (z_a1 * (t1 - t0))};
  const Node n2{(1 - z_a2) * t0 + z_a2 * t1, z_l2 * max_z + (1 - z_l2) * b.z, z_k2 * z_l2 * (b.z - max_z) / ((1 - z_a2) * (t1 - t0))};
  const Node n3{t1, b.z, z_k3 * z_l2 * (b.z - max_z) / ((1 - z_a2) * (t1 - t0))};
  sp.seg[foot][2].push_back({n0.t, n1.t, n0.p, n0.v, n1.p, n1.v});
  sp.seg[foot][2].push_back({n1.t, n2.t, n1.p, n1.v, n2.p, n2.v});
  sp.seg[foot][2].push_back({n2.t, n3.t, n2.p, n2.v, n3.p, n3.v});
}

// SwingTrajectoryPlanner::update (:164-286). latest_stance (4x3) is the planner's state (in/out).
// Returns false where the reference would throw (swing phase without a defined take-off / touch-down, :421-458).
inline bool plan_swing(const ModeSchedule& ms, const Target& tg, double init_time, const double* current_feet /*12*/, const double* body_vel_cmd /*6*/,
                       double* latest_stance /*12*/, SwingPlan& sp) {
  const int np = (int)ms.modes.size();
  const int mode_now = mode_at(ms, init_time + 0.001);
  for (int i = 0; i < 4; ++i) {
    if (contact_flag(mode_now, i)) for (int a = 0; a < 3; ++a) latest_stance[3 * i + a] = current_feet[3 * i + a];
    latest_stance[3 * i + 2] = HB_NEXT_POSITION_Z;
  }
  for (int j = 0; j < 4; ++j) {
    std::vector<bool> stock(np);
    for (int p = 0; p < np; ++p) stock[p] = contact_flag(ms.modes[p], j);
    Vec3 last{latest_stance[3 * j], latest_stance[3 * j + 1], latest_stance[3 * j + 2]}, next = last;
    int last_final_idx = 0;
    for (int p = 0; p < np; ++p) {
      int si, fi;
      find_index(p, stock, si, fi);
      if (!stock[p]) {
        if (si < 0 || fi >= np - 1) return false;      // checkThatIndicesAreValid
        const double t_start = ms.events[si], t_final = ms.events[fi];
        if (init_time < t_final && fi > last_final_idx) {
          last = next;
          double next_middle_time;
          if (fi < np - 1) {
            int si2, fi2;
            find_index(fi + 1, stock, si2, fi2);
            next_middle_time = 0.5 * (t_final + ms.events[fi2]);
          } else next_middle_time = t_final;
          double xm[22], xc[22];
          target_state(tg, next_middle_time, xm);
          target_state(tg, init_time, xc);
          const Vec3 body_vel{tg.x[0][0], tg.x[0][1], tg.x[0][2]};
          next = next_foot_pos(j, init_time, t_final, next_middle_time, xm + 6, xc + 6, body_vel, body_vel_cmd);
          last_final_idx = fi;
        }
        // every phase of a swing interval pushes the same spline set (one MultiCubicSpline per phase index in the reference);
        // only emit it once per swing interval
        if (p == 0 || stock[p - 1]) gen_swing(sp, j, t_start, t_final, last, next);
      } else {
        if (p == 0 || !stock[p - 1]) {
          const double t_start = ms.events[si], t_final = (fi >= 0 && fi < (int)ms.events.size()) ? ms.events[fi] : ms.events.back();
          const double pn[3] = {next.x, next.y, next.z};
          for (int a = 0; a < 3; ++a) sp.seg[j][a].push_back({t_start, t_final, pn[a], 0.0, pn[a], 0.0});
        }
      }
    }
  }
  return true;
}

inline int fill_reference(const ModeSchedule& ms, const Target& tg, const SwingPlan& sp, double t_lo, double t_hi, hb_reference* out) {
  memset(out, 0, sizeof(*out));
  // mode schedule restricted to the window (events strictly inside), keeping the mode in force at t_lo
  size_t first = 0;
  while (first < ms.events.size() && ms.events[first] <= t_lo) ++first;
  size_t last = first;
  while (last < ms.events.size() && ms.events[last] < t_hi) ++last;
  const int ne = (int)(last - first);
  if (ne > HB_MAX_EVENTS) return -1;
  out->n_events = ne;
  for (int i = 0; i < ne; ++i) out->event_times[i] = ms.events[first + i];
  for (int i = 0; i <= ne; ++i) out->modes[i] = ms.modes[first + i];
  out->n_targets = 2;
  for (int k = 0; k < 2; ++k) { out->target_times[k] = tg.t[k]; memcpy(out->target_states[k], tg.x[k], sizeof(double) * 22); }
  for (int c = 0; c < 4; ++c)
    for (int a = 0; a < 3; ++a) {
      int n = 0;
      for (const Seg& s : sp.seg[c][a]) {
        if (s.t1 <= t_lo || s.t0 >= t_hi) continue;
        if (n >= HB_MAX_SEGMENTS) return -1;
        double* d = out->segments[c][a][n++];
        d[0] = s.t0; d[1] = s.t1; d[2] = s.p0; d[3] = s.v0; d[4] = s.p1; d[5] = s.v1;
      }
      out->n_segments[c][a] = n;
    }
  return 0;
}

}  // namespace hbplan
