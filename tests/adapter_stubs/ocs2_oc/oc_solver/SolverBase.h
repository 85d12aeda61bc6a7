// stand-in of ocs2_oc/oc_solver/SolverBase.h: run() = preRun (reference manager's preSolverRun) + runImpl; the accessors the MRT interface uses
#pragma once
#include <memory>
#include <stdexcept>
#include "ocs2_oc/oc_data/PrimalSolution.h"
namespace ocs2 {
class SolverBase {
 public:
  virtual ~SolverBase() = default;
  virtual void reset() = 0;
  void run(scalar_t initTime, const vector_t& initState, scalar_t finalTime) {
    if (!referenceManagerPtr_) throw std::runtime_error("[SolverBase] no reference manager");
    referenceManagerPtr_->preSolverRun(initTime, finalTime, initState);
    runImpl(initTime, initState, finalTime);
  }
  void setReferenceManager(std::shared_ptr<ReferenceManagerInterface> p) { referenceManagerPtr_ = std::move(p); }
  const ReferenceManagerInterface& getReferenceManager() const { return *referenceManagerPtr_; }
  ReferenceManagerInterface& getReferenceManager() { return *referenceManagerPtr_; }
  virtual scalar_t getFinalTime() const = 0;
  virtual void getPrimalSolution(scalar_t finalTime, PrimalSolution* primalSolutionPtr) const = 0;
  PrimalSolution primalSolution(scalar_t finalTime) const { PrimalSolution p; getPrimalSolution(finalTime, &p); return p; }
  virtual size_t getNumIterations() const = 0;
 private:
  virtual void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime) = 0;
  std::shared_ptr<ReferenceManagerInterface> referenceManagerPtr_;
};
}  // namespace ocs2
