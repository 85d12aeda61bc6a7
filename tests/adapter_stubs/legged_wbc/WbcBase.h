// stand-in of legged_wbc/include/legged_wbc/WbcBase.h:31-139: the virtuals, the setters and the protected members the adapter touches
#pragma once
#include <string>
#include "ocs2_core/Types.h"
namespace legged {
using namespace ocs2;
struct PinocchioInterface {};
struct CentroidalModelInfo {};
struct PinocchioEndEffectorKinematics {};
class WbcBase {
 public:
  WbcBase(const PinocchioInterface&, CentroidalModelInfo, const PinocchioEndEffectorKinematics&) {}
  virtual ~WbcBase() = default;
  virtual void loadTasksSetting(const std::string& taskFile, bool verbose) = 0;
  virtual vector_t update(const vector_t& stateDesired, const vector_t& inputDesired, const vector_t& rbdStateMeasured, size_t mode, scalar_t period) = 0;
  void setKpKd(scalar_t swingKp, scalar_t swingKd) { swingKp_ = swingKp; swingKd_ = swingKd; }
  size_t getContactForceSize() { return contact_force_size_; }
  void setStanceMode(bool stance_mode) { stance_mode_ = stance_mode; }
 protected:
  scalar_t swingKp_{}, swingKd_{};
  size_t contact_force_size_ = 12;
  bool stance_mode_ = false;
};
}  // namespace legged
