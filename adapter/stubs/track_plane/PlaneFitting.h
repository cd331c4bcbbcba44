// STAND-IN for ov_plane/src/track_plane/PlaneFitting.h:52-104 (the two entry points the updaters call), same signatures.
#pragma once
#include <Eigen/Dense>
#include <memory>
#include <unordered_map>
#include <vector>

#include "feat/FeatureInitializer.h"
#include "ovp.h"

namespace ov_core {
class Feature;
}
namespace ov_plane {
class PlaneFitting {
public:
  static bool plane_fitting(std::vector<std::shared_ptr<ov_core::Feature>> &feats, Eigen::Vector4d &plane_abcd, int min_inlier_num = 5,
                            double max_plane_solver_condition_number = 200.0);
  static bool optimize_plane(std::vector<std::shared_ptr<ov_core::Feature>> &feats, Eigen::Vector3d &cp_inG,
                             std::unordered_map<size_t, std::unordered_map<double, ov_core::FeatureInitializer::ClonePose>> &clonesCAM,
                             double sigma_px_norm, double sigma_c, bool fix_plane, const Eigen::VectorXd &stateI, const Eigen::VectorXd &calib0);
  // added by the integration: the filter context the estimator state lives in (set once by VioManager after ovp_create)
  static ovp_ctx *ctx;
};
} // namespace ov_plane
