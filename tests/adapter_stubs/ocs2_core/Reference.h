// stand-ins of ocs2_core/reference/{ModeSchedule,TargetTrajectories}.h and ocs2_oc/synchronized_module/ReferenceManagerInterface.h
#pragma once
#include <algorithm>
#include "ocs2_core/Types.h"
namespace ocs2 {
struct ModeSchedule {
  scalar_array_t eventTimes;
  size_array_t modeSequence = size_array_t{0};
  size_t modeAtTime(scalar_t t) const { return modeSequence[static_cast<size_t>(std::lower_bound(eventTimes.begin(), eventTimes.end(), t) - eventTimes.begin())]; }
};
struct TargetTrajectories {
  scalar_array_t timeTrajectory;
  vector_array_t stateTrajectory, inputTrajectory;
  vector_t getDesiredState(scalar_t t) const {   // LinearInterpolation::interpolate, clamped
    if (timeTrajectory.size() == 1 || t <= timeTrajectory.front()) return stateTrajectory.front();
    if (t >= timeTrajectory.back()) return stateTrajectory.back();
    size_t s = 0;
    while (s + 2 < timeTrajectory.size() && timeTrajectory[s + 1] <= t) ++s;
    const scalar_t al = (t - timeTrajectory[s]) / (timeTrajectory[s + 1] - timeTrajectory[s]);
    vector_t x(stateTrajectory[s].size());
    for (long i = 0; i < x.size(); ++i) x[i] = (1.0 - al) * stateTrajectory[s][i] + al * stateTrajectory[s + 1][i];
    return x;
  }
};
class ReferenceManagerInterface {
 public:
  virtual ~ReferenceManagerInterface() = default;
  virtual void preSolverRun(scalar_t initTime, scalar_t finalTime, const vector_t& initState) = 0;
  virtual const ModeSchedule& getModeSchedule() const = 0;
  virtual const TargetTrajectories& getTargetTrajectories() const = 0;
};
class ReferenceManager : public ReferenceManagerInterface {
 public:
  void preSolverRun(scalar_t initTime, scalar_t finalTime, const vector_t& initState) override { modifyReferences(initTime, finalTime, initState, targetTrajectories_, modeSchedule_); }
  const ModeSchedule& getModeSchedule() const override { return modeSchedule_; }
  const TargetTrajectories& getTargetTrajectories() const override { return targetTrajectories_; }
  void setModeSchedule(ModeSchedule m) { modeSchedule_ = std::move(m); }
  void setTargetTrajectories(TargetTrajectories t) { targetTrajectories_ = std::move(t); }
 protected:
  virtual void modifyReferences(scalar_t, scalar_t, const vector_t&, TargetTrajectories&, ModeSchedule&) {}
  ModeSchedule modeSchedule_;
  TargetTrajectories targetTrajectories_;
};
}  // namespace ocs2
