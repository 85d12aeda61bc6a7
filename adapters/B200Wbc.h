// B200Wbc -- drop-in replacement of legged::WeightedWbc on top of libhunter_b200.so (C ABI: include/hunter_b200.h).
//
// Lives in the reference's catkin workspace next to legged_wbc/include/legged_wbc/WeightedWbc.h. It derives from the reference's own
// legged::WbcBase (legged_wbc/include/legged_wbc/WbcBase.h:31-139) and overrides the two virtuals the controller calls:
//   loadTasksSetting(taskFile, verbose)   LeggedController.cpp:86   -> hb_load_task_info (gains, limits, weights read from task.info)
//   update(stateDesired, inputDesired, rbdStateMeasured, mode, period)   LeggedController.cpp:180 -> hb_wbc_solve_batch (B = 1)
// The non-virtual setters of WbcBase the controller uses keep working because the adapter forwards the members they set:
//   setStanceMode (WbcBase.h:73-76, LeggedController.cpp:161-173) -> stance_mode flag of the call
//   setKpKd       (WbcBase.h:65-69)                              -> hb_wbc_set_kp_kd before every solve
// Controller change: LeggedController.cpp:85  wbc_ = std::make_shared<B200Wbc>(leggedInterface_->getPinocchioInterface(), ...same arguments...);
//
// tests/test_adapters.py compiles this header against minimal stand-ins of the reference headers (tests/adapter_stubs/) and runs it on the GPU.
#pragma once

#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <string>

#include <hunter_b200.h>

#include "legged_wbc/WbcBase.h"

namespace legged
{
class B200Wbc : public WbcBase
{
public:
  using WbcBase::WbcBase;

  ~B200Wbc() override
  {
    if (ctx_ != nullptr)
      hb_destroy(ctx_);
  }

  // device index of the GPU that solves this robot's QPs (call before loadTasksSetting; default 0)
  void setDevice(int device)
  {
    device_ = device;
  }

  void loadTasksSetting(const std::string& taskFile, bool verbose) override
  {
    ensureContext();
    const int rc = hb_load_task_info(ctx_, taskFile.c_str());
    if (rc != 0)
      throw std::runtime_error(std::string("[B200Wbc] cannot load task settings from ") + taskFile + ": " + hb_strerror(rc));
    hb_wbc_settings s;
    hb_wbc_get_settings(ctx_, &s);
    swingKp_ = s.swing_kp;  // keep the base-class members in step, setKpKd() edits them later
    swingKd_ = s.swing_kd;
    kpKdLoaded_ = true;
    if (verbose)
    {
      std::cerr << "\n #### B200Wbc settings (from " << taskFile << "):"
                << "\n #### torque limits " << s.torque_limits[0] << " " << s.torque_limits[1] << " " << s.torque_limits[2] << " "
                << s.torque_limits[3] << " " << s.torque_limits[4] << ", friction " << s.friction_coefficient << "\n #### swing kp/kd "
                << s.swing_kp << "/" << s.swing_kd << ", base height kp/kd " << s.base_height_kp << "/" << s.base_height_kd
                << ", base angular kp/kd " << s.base_angular_kp << "/" << s.base_angular_kd << "\n #### weights swing " << s.weight_swing_leg
                << ", base " << s.weight_base_accel << ", contact force " << s.weight_contact_force << "\n";
    }
  }

  vector_t update(const vector_t& stateDesired, const vector_t& inputDesired, const vector_t& rbdStateMeasured, size_t mode,
                  scalar_t /*period*/) override
  {
    ensureContext();
    if (kpKdLoaded_)
      hb_wbc_set_kp_kd(ctx_, swingKp_, swingKd_);  // WbcBase::setKpKd writes these members
    const int32_t m = static_cast<int32_t>(mode);
    const uint8_t stance = stance_mode_ ? 1 : 0;
    vector_t sol(38);
    int32_t status = 0;
    const int rc = hb_wbc_solve_batch(ctx_, 1, stateDesired.data(), inputDesired.data(), rbdStateMeasured.data(), &m, &stance, sol.data(),
                                      &status);
    if (rc != 0)
      throw std::runtime_error(std::string("[B200Wbc] hb_wbc_solve_batch: ") + hb_strerror(rc));
    if (status != 0)
    {  // same fallback as WeightedWbc::update, WeightedWbc.cpp:57-62
      std::cout << "ERROR: WeightWBC Not Solved!!!" << std::endl;
      if (last_qpSol_.size() > 0)
        sol = last_qpSol_;
    }
    last_qpSol_ = sol;
    return sol;  // [qdd(16) | F(12) | tau(10)]
  }

private:
  void ensureContext()
  {
    if (ctx_ != nullptr)
      return;
    hb_config cfg;
    hb_default_config(&cfg);
    cfg.max_batch = 1;
    cfg.horizon_N = 1;  // WBC only: no MPC scratch
    const int rc = hb_create(&cfg, device_, &ctx_);
    if (rc != 0)
      throw std::runtime_error(std::string("[B200Wbc] hb_create: ") + hb_strerror(rc));
  }

  hb_ctx* ctx_ = nullptr;
  int device_ = 0;
  bool kpKdLoaded_ = false;
  vector_t last_qpSol_;
};

}  // namespace legged
