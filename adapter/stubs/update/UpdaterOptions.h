// STAND-IN for ov_plane/src/update/UpdaterOptions.h:38-54.  Syntax check only.
#pragma once
namespace ov_plane {
struct UpdaterOptions {
  double chi2_multipler = 5;
  double sigma_pix = 1;
  double sigma_pix_sq = 1;
};
} // namespace ov_plane
namespace ov_core {
struct FeatureInitializerOptions {};
} // namespace ov_core
