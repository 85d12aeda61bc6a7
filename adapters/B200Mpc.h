// B200Mpc -- drop-in replacement of ocs2::SqpMpc for the Hunter controller on top of libhunter_b200.so (C ABI: include/hunter_b200.h).
//
// The controller owns its MPC as std::shared_ptr<ocs2::MPC_BASE> (legged_controllers/include/legged_controllers/LeggedController.h:94), builds
// it in setupMpc() (LeggedController.cpp:376-388) and only ever talks to it through MPC_MRT_Interface: advanceMpc() -> MPC_BASE::run(t, x) ->
// calculateController(t0, x0, tf) -> SolverBase::run, then SolverBase::getPrimalSolution (LeggedController.cpp:144-156, 406).
// B200Mpc keeps that surface: it is an MPC_BASE whose solver is a SolverBase that runs ONE SQP iteration on the GPU per call
// (sqp.sqpIteration 1, task.info:83) with the solver's own warm start (mpc.coldStart false, task.info:146).
//
// What stays the reference's: the SwitchedModelReferenceManager attached with setReferenceManager (LeggedController.cpp:383-386) still runs
// its preSolverRun (gait schedule, swing planner, joint references); the adapter samples its products at the node times:
//   mode schedule      -> node modes (post-event mode at a switch), event times -> node grid (hb_time_grid_batch, event nodes as OCS2)
//   target trajectories-> x_ref per node (TargetTrajectories::getDesiredState)
//   swing planner      -> [px py pz vx vy vz] per contact and node (SwingTrajectoryPlanner::get{X,Y,Z}{position,velocity}Constraint)
// Controller change: LeggedController.cpp:378
//   mpc_ = std::make_shared<B200Mpc>(leggedInterface_->mpcSettings(), leggedInterface_->sqpSettings(),
//                                    leggedInterface_->getSwitchedModelReferenceManagerPtr());
//
// tests/test_adapters.py compiles this header against minimal stand-ins of the OCS2 / reference headers (tests/adapter_stubs/) and runs two
// MPC cycles through it on the GPU, checking the published PrimalSolution against the same cycles driven through the Python binding.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <hunter_b200.h>

#include <ocs2_core/Types.h>
#include <ocs2_mpc/MPC_BASE.h>
#include <ocs2_oc/oc_data/PrimalSolution.h>
#include <ocs2_oc/oc_solver/SolverBase.h>
#include <ocs2_sqp/SqpSettings.h>

#include "legged_interface/SwitchedModelReferenceManager.h"

namespace legged
{
using namespace ocs2;

class B200Solver final : public SolverBase
{
public:
  B200Solver(const sqp::Settings& sqpSettings, scalar_t timeHorizon, std::shared_ptr<SwitchedModelReferenceManager> switchedReferenceManager,
             int device = 0)
    : switched_(std::move(switchedReferenceManager)), dt_(sqpSettings.dt), horizon_(timeHorizon)
  {
    hb_config cfg;
    hb_default_config(&cfg);
    cfg.dt = dt_;
    cfg.time_horizon = horizon_;
    cfg.event_nodes = 1;
    cfg.horizon_N = static_cast<int32_t>(std::ceil(horizon_ / dt_)) + 12;  // node capacity: the grid plus a node per mode switch
    cfg.max_batch = 1;
    N_ = cfg.horizon_N;
    const int rc = hb_create(&cfg, device, &ctx_);
    if (rc != 0)
      throw std::runtime_error(std::string("[B200Mpc] hb_create: ") + hb_strerror(rc));
    tk_.assign(N_ + 1, 0.0);
    xRef_.assign((N_ + 1) * 22, 0.0);
    swing_.assign((N_ + 1) * 24, 0.0);
    mode_.assign(N_ + 1, 3);
    xTraj_.assign((N_ + 1) * 22, 0.0);
    uTraj_.assign(N_ * 22, 0.0);
  }

  ~B200Solver() override
  {
    if (ctx_ != nullptr)
      hb_destroy(ctx_);
  }

  void reset() override
  {
    first_ = true;
    primalSolution_ = PrimalSolution();
    iterations_ = 0;
  }

  scalar_t getFinalTime() const override
  {
    return primalSolution_.timeTrajectory_.empty() ? 0.0 : primalSolution_.timeTrajectory_.back();
  }

  void getPrimalSolution(scalar_t /*finalTime*/, PrimalSolution* primalSolutionPtr) const override
  {
    *primalSolutionPtr = primalSolution_;
  }

  size_t getNumIterations() const override
  {
    return iterations_;
  }

  const hb_solve_info& lastSolveInfo() const
  {
    return info_;
  }

private:
  void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime) override
  {
    // SolverBase::run has already called the reference manager's preSolverRun(initTime, finalTime, initState)
    const ModeSchedule& modeSchedule = getReferenceManager().getModeSchedule();
    const TargetTrajectories& target = getReferenceManager().getTargetTrajectories();
    (void)finalTime;  // the horizon is mpc.timeHorizon, as for SqpMpc

    // 1. node grid: only the event part of hb_reference is needed for hb_time_grid_batch
    hb_reference ref{};
    const auto& ev = modeSchedule.eventTimes;
    int first = 0;
    while (first < static_cast<int>(ev.size()) && ev[first] <= initTime - horizon_)
      ++first;  // keep the capacity for the events that matter
    const int nev = std::min<int>(static_cast<int>(ev.size()) - first, HB_MAX_EVENTS);
    ref.n_events = nev;
    for (int i = 0; i < nev; ++i)
      ref.event_times[i] = ev[first + i];
    for (int i = 0; i <= nev; ++i)
      ref.modes[i] = static_cast<int32_t>(modeSchedule.modeSequence[first + i]);
    ref.n_targets = 1;
    int32_t nn = 0, gridStatus = 0;
    std::vector<scalar_t> tkNew(N_ + 1);
    check(hb_time_grid_batch(ctx_, 1, &initTime, &ref, tkNew.data(), &nn, &gridStatus), "hb_time_grid_batch");
    if (gridStatus != 0)
      throw std::runtime_error("[B200Mpc] more mode switches inside the horizon than the node capacity allows");

    // 2. references sampled at the node times
    const auto& planner = *switched_->getSwingTrajectoryPlanner();
    for (int k = 0; k <= N_; ++k)
    {
      const scalar_t t = tkNew[std::min<int>(k, nn)];
      const vector_t xd = target.getDesiredState(t);
      for (int i = 0; i < 22; ++i)
        xRef_[k * 22 + i] = xd[i];
      mode_[k] = static_cast<int32_t>(modeAtNode(modeSchedule, t));
      for (size_t c = 0; c < 4; ++c)
      {
        scalar_t* s = &swing_[k * 24 + 6 * c];
        s[0] = planner.getXpositionConstraint(c, t);
        s[1] = planner.getYpositionConstraint(c, t);
        s[2] = planner.getZpositionConstraint(c, t);
        s[3] = planner.getXvelocityConstraint(c, t);
        s[4] = planner.getYvelocityConstraint(c, t);
        s[5] = planner.getZvelocityConstraint(c, t);
      }
    }

    // 3. warm start (SqpSolver::initializeStateInputTrajectories): previous solution interpolated on the new grid, initializer beyond it
    if (first_)
    {
      check(hb_mpc_cold_start_batch(ctx_, 1, initState.data(), mode_.data(), xTraj_.data(), uTraj_.data()), "hb_mpc_cold_start_batch");
    }
    else
    {
      shiftPreviousSolution(tkNew, nn, initState);
    }

    // 4. one SQP iteration on the device
    check(hb_mpc_solve_grid_batch(ctx_, 1, initState.data(), tkNew.data(), &nn, xRef_.data(), swing_.data(), mode_.data(), xTraj_.data(),
                                  uTraj_.data(), &info_),
          "hb_mpc_solve_grid_batch");
    if (info_.status != 0)
      throw std::runtime_error("[B200Mpc] numerical failure in the SQP iteration");  // stops the MPC thread, LeggedController.cpp:413-418
    tk_ = tkNew;
    nn_ = nn;
    first_ = false;
    ++iterations_;

    // 5. publish (feed-forward policy, sqp.useFeedbackPolicy false, task.info:93)
    primalSolution_.timeTrajectory_.assign(tk_.begin(), tk_.begin() + nn_ + 1);
    primalSolution_.stateTrajectory_.resize(nn_ + 1);
    primalSolution_.inputTrajectory_.resize(nn_ + 1);
    for (int k = 0; k <= nn_; ++k)
    {
      primalSolution_.stateTrajectory_[k] = vector_t(22);
      primalSolution_.inputTrajectory_[k] = vector_t(22);
      const int ku = std::min(k, nn_ - 1);  // the input trajectory repeats its last sample at the final node
      for (int i = 0; i < 22; ++i)
      {
        primalSolution_.stateTrajectory_[k][i] = xTraj_[k * 22 + i];
        primalSolution_.inputTrajectory_[k][i] = uTraj_[ku * 22 + i];
      }
    }
    primalSolution_.modeSchedule_ = modeSchedule;
    primalSolution_.controllerPtr_.reset(new FeedforwardController(primalSolution_.timeTrajectory_, primalSolution_.inputTrajectory_));
  }

  // mode in force on the interval that starts at t (the post-event mode when t is a switching time)
  static size_t modeAtNode(const ModeSchedule& ms, scalar_t t)
  {
    size_t idx = 0;
    while (idx < ms.eventTimes.size() && ms.eventTimes[idx] <= t + 1e-9)
      ++idx;
    return ms.modeSequence[idx];
  }

  void shiftPreviousSolution(const std::vector<scalar_t>& tkNew, int nnNew, const vector_t& initState)
  {
    const std::vector<scalar_t> xp = xTraj_, up = uTraj_;
    const scalar_t tEnd = tk_[nn_];
    auto locate = [&](scalar_t t, scalar_t& al) {
      int k = static_cast<int>(std::upper_bound(tk_.begin(), tk_.begin() + nn_ + 1, t) - tk_.begin()) - 1;
      k = std::max(0, std::min(k, nn_ - 1));
      const scalar_t d = tk_[k + 1] - tk_[k];
      al = d > 0.0 ? std::min(1.0, std::max(0.0, (t - tk_[k]) / d)) : 0.0;
      return k;
    };
    int istar = nnNew;
    for (int i = 0; i < nnNew; ++i)
      if (tkNew[i + 1] > tEnd + 1e-9)
      {
        istar = i;
        break;
      }
    for (int k = 0; k <= N_; ++k)
    {
      const int ks = std::min(std::min(k, nnNew), istar);
      if (ks == 0)
      {
        for (int i = 0; i < 22; ++i)
          xTraj_[k * 22 + i] = initState[i];
        continue;
      }
      scalar_t al;
      const int j = locate(tkNew[ks], al);
      for (int i = 0; i < 22; ++i)
        xTraj_[k * 22 + i] = (1.0 - al) * xp[j * 22 + i] + al * xp[(j + 1) * 22 + i];
    }
    for (int k = 0; k < N_; ++k)
    {
      if (k < istar && k < nnNew)
      {
        scalar_t al;
        const int j = locate(tkNew[k], al);
        const int j1 = std::min(j + 1, nn_ - 1);
        for (int i = 0; i < 22; ++i)
          uTraj_[k * 22 + i] = (1.0 - al) * up[j * 22 + i] + al * up[j1 * 22 + i];
      }
      else
      {  // LeggedRobotInitializer: weight-compensating input
        const int32_t m = mode_[std::min(k, nnNew)];
        const bool fl[4] = { m == 2 || m == 3, m == 1 || m == 3, m == 2 || m == 3, m == 1 || m == 3 };
        const int ns = fl[0] + fl[1] + fl[2] + fl[3];
        for (int i = 0; i < 22; ++i)
          uTraj_[k * 22 + i] = 0.0;
        for (int c = 0; c < 4; ++c)
          if (fl[c])
            uTraj_[k * 22 + 3 * c + 2] = robotWeight_ / ns;
      }
    }
  }

  void check(int rc, const char* what) const
  {
    if (rc != 0)
      throw std::runtime_error(std::string("[B200Mpc] ") + what + ": " + hb_strerror(rc));
  }

  std::shared_ptr<SwitchedModelReferenceManager> switched_;
  hb_ctx* ctx_ = nullptr;
  scalar_t dt_, horizon_;
  int N_ = 0, nn_ = 0;
  bool first_ = true;
  size_t iterations_ = 0;
  const scalar_t robotWeight_ = 12.586944 * 9.81;  // total mass of the URDF (the initializer's weight compensation)
  std::vector<scalar_t> tk_, xRef_, swing_, xTraj_, uTraj_;
  std::vector<int32_t> mode_;
  hb_solve_info info_{};
  PrimalSolution primalSolution_;
};

class B200Mpc final : public MPC_BASE
{
public:
  B200Mpc(mpc::Settings mpcSettings, const sqp::Settings& sqpSettings, std::shared_ptr<SwitchedModelReferenceManager> switchedReferenceManager,
          int device = 0)
    : MPC_BASE(std::move(mpcSettings))
  {
    solverPtr_.reset(new B200Solver(sqpSettings, this->settings().timeHorizon_, std::move(switchedReferenceManager), device));
  }

  ~B200Mpc() override = default;

  B200Solver* getSolverPtr() override
  {
    return solverPtr_.get();
  }
  const B200Solver* getSolverPtr() const override
  {
    return solverPtr_.get();
  }

protected:
  void calculateController(scalar_t initTime, const vector_t& initState, scalar_t finalTime) override
  {
    if (settings().coldStart_)
      solverPtr_->reset();
    solverPtr_->run(initTime, initState, finalTime);
  }

private:
  std::unique_ptr<B200Solver> solverPtr_;
};

}  // namespace legged
