// stand-in of ocs2_core/Types.h (used by every reference header through legged_wbc/Task.h:17): scalar_t, vector_t (Eigen::VectorXd upstream)
#pragma once
#include <cstddef>
#include <vector>
namespace ocs2 {
using scalar_t = double;
class vector_t {   // the subset of Eigen::VectorXd the adapters use: size(), data(), operator[], construction from a size
 public:
  vector_t() = default;
  explicit vector_t(long n) : v_(static_cast<size_t>(n), 0.0) {}
  long size() const { return static_cast<long>(v_.size()); }
  scalar_t* data() { return v_.data(); }
  const scalar_t* data() const { return v_.data(); }
  scalar_t& operator[](long i) { return v_[static_cast<size_t>(i)]; }
  const scalar_t& operator[](long i) const { return v_[static_cast<size_t>(i)]; }
 private:
  std::vector<scalar_t> v_;
};
using scalar_array_t = std::vector<scalar_t>;
using vector_array_t = std::vector<vector_t>;
using size_array_t = std::vector<size_t>;
}  // namespace ocs2
