// STAND-IN for ov_plane/src/update/UpdaterMSCKF.h:52-93: the class as the reference declares it (same members, same signatures).
#pragma once
#include <Eigen/Dense>
#include <map>
#include <memory>
#include <vector>

#include "feat/FeatureInitializer.h"
#include "update/UpdaterOptions.h"

namespace ov_core {
class Feature;
}
namespace ov_plane {
class State;
class UpdaterMSCKF {
public:
  UpdaterMSCKF(UpdaterOptions &options, ov_core::FeatureInitializerOptions &feat_init_options);
  void update(std::shared_ptr<State> state, std::vector<std::shared_ptr<ov_core::Feature>> &feature_vec,
              std::vector<std::shared_ptr<ov_core::Feature>> &feature_vec_extra,
              std::vector<std::shared_ptr<ov_core::Feature>> &feature_vec_used, const std::map<size_t, size_t> &feat2plane);

protected:
  UpdaterOptions _options;
  std::shared_ptr<ov_core::FeatureInitializer> initializer_feat;
  std::map<int, double> chi_squared_table;
};
} // namespace ov_plane
