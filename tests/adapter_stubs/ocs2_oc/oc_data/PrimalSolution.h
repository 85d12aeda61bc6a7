// stand-in of ocs2_oc/oc_data/PrimalSolution.h and ocs2_core/control/{ControllerBase,FeedforwardController}.h
#pragma once
#include <memory>
#include "ocs2_core/Reference.h"
namespace ocs2 {
class ControllerBase {
 public:
  virtual ~ControllerBase() = default;
  virtual vector_t computeInput(scalar_t t, const vector_t& x) = 0;
  virtual ControllerBase* clone() const = 0;
};
class FeedforwardController final : public ControllerBase {
 public:
  FeedforwardController(scalar_array_t time, vector_array_t input) : timeStamp_(std::move(time)), uffArray_(std::move(input)) {}
  vector_t computeInput(scalar_t t, const vector_t&) override {
    size_t s = 0;
    while (s + 2 < timeStamp_.size() && timeStamp_[s + 1] <= t) ++s;
    if (timeStamp_.size() < 2) return uffArray_.front();
    const scalar_t al = std::min(1.0, std::max(0.0, (t - timeStamp_[s]) / (timeStamp_[s + 1] - timeStamp_[s])));
    vector_t u(uffArray_[s].size());
    for (long i = 0; i < u.size(); ++i) u[i] = (1.0 - al) * uffArray_[s][i] + al * uffArray_[s + 1][i];
    return u;
  }
  FeedforwardController* clone() const override { return new FeedforwardController(*this); }
  scalar_array_t timeStamp_;
  vector_array_t uffArray_;
};
struct PrimalSolution {
  PrimalSolution() = default;
  PrimalSolution(const PrimalSolution& o)
    : timeTrajectory_(o.timeTrajectory_), stateTrajectory_(o.stateTrajectory_), inputTrajectory_(o.inputTrajectory_), modeSchedule_(o.modeSchedule_),
      controllerPtr_(o.controllerPtr_ ? o.controllerPtr_->clone() : nullptr) {}
  PrimalSolution& operator=(const PrimalSolution& o) { PrimalSolution t(o); swap(t); return *this; }
  void swap(PrimalSolution& o) {
    timeTrajectory_.swap(o.timeTrajectory_); stateTrajectory_.swap(o.stateTrajectory_); inputTrajectory_.swap(o.inputTrajectory_);
    std::swap(modeSchedule_, o.modeSchedule_); controllerPtr_.swap(o.controllerPtr_);
  }
  scalar_array_t timeTrajectory_;
  vector_array_t stateTrajectory_, inputTrajectory_;
  ModeSchedule modeSchedule_;
  std::unique_ptr<ControllerBase> controllerPtr_;
};
}  // namespace ocs2
