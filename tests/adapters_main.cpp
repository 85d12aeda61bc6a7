// Test driver of the reference-side adapters (adapters/B200Wbc.h, adapters/B200Mpc.h) compiled against the stand-in headers of
// tests/adapter_stubs/: reads a scenario written by tests/test_adapters.py, drives the adapters the way LeggedController does
// (wbc_->loadTasksSetting / setStanceMode / setKpKd / update; mpc_->getSolverPtr()->setReferenceManager; mpc_->run; primalSolution)
// and writes what they return.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "B200Mpc.h"
#include "B200Wbc.h"

using namespace legged;

struct Seg { double t0, t1, p0, v0, p1, v1; };
static std::vector<Seg> g_segs[4][3];

static double hermite(const Seg& s, double t, bool vel) {
  const double T = s.t1 - s.t0, tn = (t - s.t0) / T, dp = s.p1 - s.p0, dv = s.v1 - s.v0;
  const double c0 = s.p0, c1 = s.v0 * T, c2 = -(3.0 * s.v0 + dv) * T + 3.0 * dp, c3 = (2.0 * s.v0 + dv) * T - 2.0 * dp;
  return vel ? ((3.0 * c3 * tn + 2.0 * c2) * tn + c1) / T : ((c3 * tn + c2) * tn + c1) * tn + c0;
}
static double swing_value(size_t leg, int axis, double t) {
  const auto& v = g_segs[leg][axis % 3];
  size_t s = 0;
  while (s + 1 < v.size() && t >= v[s].t1) ++s;
  return hermite(v[s], t, axis >= 3);
}
static vector_t read_vec(std::istream& in, int n) { vector_t v(n); for (int i = 0; i < n; ++i) in >> v[i]; return v; }

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: adapters_main scenario.txt out.txt\n"); return 2; }
  std::ifstream in(argv[1]);
  std::ofstream out(argv[2]);
  out.precision(17);
  std::string taskFile;
  in >> taskFile;
  // ---------------- WBC
  {
    B200Wbc wbc(PinocchioInterface{}, CentroidalModelInfo{}, PinocchioEndEffectorKinematics{});
    wbc.loadTasksSetting(taskFile, false);
    int ncase; in >> ncase;
    for (int c = 0; c < ncase; ++c) {
      int mode, stance; double kp, kd;
      in >> mode >> stance >> kp >> kd;
      const vector_t x = read_vec(in, 22), u = read_vec(in, 22), rbd = read_vec(in, 32);
      wbc.setStanceMode(stance != 0);
      if (kp > 0.0) wbc.setKpKd(kp, kd);
      const vector_t sol = wbc.update(x, u, rbd, static_cast<size_t>(mode), 0.002);
      out << "wbc";
      for (int i = 0; i < 38; ++i) out << " " << sol[i];
      out << "\n";
    }
  }
  // ---------------- MPC
  {
    double dt, horizon; in >> dt >> horizon;
    ModeSchedule ms;
    int nev; in >> nev;
    ms.eventTimes.resize(nev); ms.modeSequence.resize(nev + 1);
    for (int i = 0; i < nev; ++i) in >> ms.eventTimes[i];
    for (int i = 0; i <= nev; ++i) in >> ms.modeSequence[i];
    TargetTrajectories tg;
    int nt; in >> nt;
    for (int i = 0; i < nt; ++i) { double t; in >> t; tg.timeTrajectory.push_back(t); tg.stateTrajectory.push_back(read_vec(in, 22)); }
    for (int c = 0; c < 4; ++c) for (int a = 0; a < 3; ++a) {
      int ns; in >> ns;
      g_segs[c][a].resize(ns);
      for (auto& s : g_segs[c][a]) in >> s.t0 >> s.t1 >> s.p0 >> s.v0 >> s.p1 >> s.v1;
    }
    auto planner = std::make_shared<SwingTrajectoryPlanner>(swing_value);
    auto refMgr = std::make_shared<SwitchedModelReferenceManager>(planner);
    refMgr->setModeSchedule(ms);
    refMgr->setTargetTrajectories(tg);
    mpc::Settings mpcSettings; mpcSettings.timeHorizon_ = horizon;
    sqp::Settings sqpSettings; sqpSettings.dt = dt;
    std::shared_ptr<MPC_BASE> mpc = std::make_shared<B200Mpc>(mpcSettings, sqpSettings, refMgr);
    mpc->getSolverPtr()->setReferenceManager(refMgr);                  // LeggedController.cpp:386
    int ncyc; in >> ncyc;
    for (int c = 0; c < ncyc; ++c) {
      double t; in >> t;
      const vector_t x0 = read_vec(in, 22);
      mpc->run(t, x0);                                                  // MPC_MRT_Interface::advanceMpc
      const PrimalSolution ps = mpc->getSolverPtr()->primalSolution(mpc->getSolverPtr()->getFinalTime());
      out << "mpc " << ps.timeTrajectory_.size();
      for (size_t k = 0; k < ps.timeTrajectory_.size(); ++k) {
        out << " " << ps.timeTrajectory_[k];
        for (int i = 0; i < 22; ++i) out << " " << ps.stateTrajectory_[k][i];
        for (int i = 0; i < 22; ++i) out << " " << ps.inputTrajectory_[k][i];
      }
      const vector_t u_mid = ps.controllerPtr_->computeInput(t + 0.002, x0);   // evaluatePolicy with the feed-forward controller
      for (int i = 0; i < 22; ++i) out << " " << u_mid[i];
      out << "\n";
    }
  }
  return 0;
}
