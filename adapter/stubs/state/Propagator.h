// STAND-IN for ov_plane/src/state/Propagator.h:45-160 (constructor, feed_imu, propagate_and_clone, fast_state_propagate).  Syntax check only.
#pragma once
#include <Eigen/Dense>
#include <memory>

#include "ovp.h"
namespace ov_core {
struct ImuData {
  double timestamp;
  Eigen::Vector3d wm, am;
};
} // namespace ov_core
namespace ov_plane {
class State;
struct NoiseManager { // utils/NoiseManager.h:41-63
  double sigma_w = 1.6968e-04, sigma_wb = 1.9393e-05, sigma_a = 2.0000e-3, sigma_ab = 3.0000e-03;
};
class Propagator {
public:
  Propagator(NoiseManager noises, double gravity_mag);
  void feed_imu(const ov_core::ImuData &message, double oldest_time = -1);
  void propagate_and_clone(std::shared_ptr<State> state, double timestamp);
  bool fast_state_propagate(std::shared_ptr<State> state, double timestamp, Eigen::Matrix<double, 13, 1> &state_plus,
                            Eigen::Matrix<double, 12, 12> &covariance);

protected:
  NoiseManager _noises;
  double _gravity_mag;
  // added by the integration: feed_imu has no State argument in the reference, the samples go to the filter context directly
public:
  static ovp_ctx *ctx;
};
} // namespace ov_plane
