// stand-in of legged_interface/include/legged_interface/SwitchedModelReferenceManager.h:67-119
#pragma once
#include <memory>
#include "legged_interface/foot_planner/SwingTrajectoryPlanner.h"
#include "ocs2_core/Reference.h"
namespace legged {
using namespace ocs2;
class SwitchedModelReferenceManager : public ReferenceManager {
 public:
  explicit SwitchedModelReferenceManager(std::shared_ptr<SwingTrajectoryPlanner> p) : swingTrajectoryPtr_(std::move(p)) {}
  const std::shared_ptr<SwingTrajectoryPlanner>& getSwingTrajectoryPlanner() { return swingTrajectoryPtr_; }
 private:
  std::shared_ptr<SwingTrajectoryPlanner> swingTrajectoryPtr_;
};
}  // namespace legged
