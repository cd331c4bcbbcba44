// MINIMAL STAND-IN for ov_core feat/FeatureInitializer.h: ClonePose as used by PlaneFitting::optimize_plane.  Syntax check only.
#pragma once
#include <Eigen/Dense>
namespace ov_core {
class FeatureInitializer {
public:
  struct ClonePose {
    Eigen::Matrix3d _Rot;
    Eigen::Vector3d _pos;
    const Eigen::Matrix3d &Rot() { return _Rot; }
    const Eigen::Vector3d &pos() { return _pos; }
  };
};
} // namespace ov_core
