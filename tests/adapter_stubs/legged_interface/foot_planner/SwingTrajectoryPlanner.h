// stand-in of legged_interface/include/legged_interface/foot_planner/SwingTrajectoryPlanner.h:116-122 (the six constraint accessors the solver
// evaluates per node); the test implementation returns the references of the workload generator
#pragma once
#include <functional>
#include "ocs2_core/Types.h"
namespace legged {
using namespace ocs2;
class SwingTrajectoryPlanner {
 public:
  // test hook: f(leg, axis (0..2 position, 3..5 velocity), time)
  explicit SwingTrajectoryPlanner(std::function<scalar_t(size_t, int, scalar_t)> f) : f_(std::move(f)) {}
  scalar_t getXvelocityConstraint(size_t leg, scalar_t time) const { return f_(leg, 3, time); }
  scalar_t getYvelocityConstraint(size_t leg, scalar_t time) const { return f_(leg, 4, time); }
  scalar_t getZvelocityConstraint(size_t leg, scalar_t time) const { return f_(leg, 5, time); }
  scalar_t getXpositionConstraint(size_t leg, scalar_t time) const { return f_(leg, 0, time); }
  scalar_t getYpositionConstraint(size_t leg, scalar_t time) const { return f_(leg, 1, time); }
  scalar_t getZpositionConstraint(size_t leg, scalar_t time) const { return f_(leg, 2, time); }
 private:
  std::function<scalar_t(size_t, int, scalar_t)> f_;
};
}  // namespace legged
