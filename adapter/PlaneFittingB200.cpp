// Reference-side adapter for ov_plane/src/track_plane/PlaneFitting.cpp: the two entry points the updaters call
// (UpdaterMSCKF.cpp:279,321,353, UpdaterPlane.cpp:230,261, UpdaterSLAM.cpp:171) with the reference's own signatures, forwarding to the
// C ABI (include/ovp.h: ovp_plane_fitting, ovp_optimize_plane).  Built instead of PlaneFitting.cpp and linked with -lovp; no Ceres needed.
// This one-plane form keeps the call sites unchanged; a caller that gathers the candidate planes of a frame first passes them as ONE batch
// (n_planes > 1) and gets them evaluated concurrently.
#include "track_plane/PlaneFitting.h"

#include <cstdio>
#include <cstdlib>

#include "feat/Feature.h"

using namespace ov_plane;

ovp_ctx *PlaneFitting::ctx = nullptr;

static void ck(int st) {
  if (st) {
    std::fprintf(stderr, "[PLANE-FIT]: %s\n", ovp_last_error(PlaneFitting::ctx));
    std::exit(EXIT_FAILURE);
  }
}

bool PlaneFitting::plane_fitting(std::vector<std::shared_ptr<ov_core::Feature>> &feats, Eigen::Vector4d &plane_abcd, int min_inlier_num,
                                 double max_plane_solver_condition_number) {
  const int F = (int)feats.size();
  std::vector<double> p(3 * (size_t)F);
  for (int i = 0; i < F; i++)
    for (int k = 0; k < 3; k++)
      p[3 * (size_t)i + k] = feats[i]->p_FinG(k);
  const int offs[2] = {0, F};
#if defined(__GLIBCXX__) && defined(_GLIBCXX_RELEASE) && _GLIBCXX_RELEASE >= 11
  const int shuffle_kind = 1; // the std::shuffle this very build of the reference would have used
#else
  const int shuffle_kind = 0;
#endif
  ovp_plane_fit_options o = {min_inlier_num, max_plane_solver_condition_number, shuffle_kind};
  int status = 0;
  double abcd[4];
  std::vector<int> inl((size_t)(F > 0 ? F : 1));
  ck(ovp_plane_fitting(ctx, 1, offs, p.data(), &o, &status, abcd, inl.data()));
  if (!status)
    return false;
  for (int k = 0; k < 4; k++)
    plane_abcd(k) = abcd[k];
  std::vector<std::shared_ptr<ov_core::Feature>> keep; // feats = best_inliers (PlaneFitting.cpp:186)
  for (int i = 0; i < F; i++)
    if (inl[i])
      keep.push_back(feats[i]);
  feats = keep;
  return true;
}

bool PlaneFitting::optimize_plane(std::vector<std::shared_ptr<ov_core::Feature>> &feats, Eigen::Vector3d &cp_inG,
                                  std::unordered_map<size_t, std::unordered_map<double, ov_core::FeatureInitializer::ClonePose>> &clonesCAM,
                                  double sigma_px_norm, double sigma_c, bool fix_plane, const Eigen::VectorXd &stateI, const Eigen::VectorXd &calib0) {
  // clonesCAM / stateI / calib0 are derived from the filter state (UpdaterMSCKF.cpp:122-140, 273-274), which lives in the context: the camera
  // poses are rebuilt on the device from the clone and extrinsics values, so only the time stamps of the measurements travel
  (void)clonesCAM;
  (void)stateI;
  (void)calib0;
  const int F = (int)feats.size();
  std::vector<int> meas_offset(1, 0), meas_clone;
  std::vector<float> uvn;
  std::vector<double> p(3 * (size_t)F);
  for (int i = 0; i < F; i++) {
    for (int k = 0; k < 3; k++)
      p[3 * (size_t)i + k] = feats[i]->p_FinG(k);
    for (auto const &pair : feats[i]->timestamps) // mono: one camera id
      for (size_t m = 0; m < pair.second.size(); m++) {
        meas_clone.push_back(ovp_clone_handle(ctx, pair.second[m]));
        uvn.push_back(feats[i]->uvs_norm.at(pair.first).at(m)(0));
        uvn.push_back(feats[i]->uvs_norm.at(pair.first).at(m)(1));
      }
    meas_offset.push_back((int)meas_clone.size());
  }
  const int offs[2] = {0, F}, fix = fix_plane ? 1 : 0;
  ovp_plane_refine_options o = {sigma_px_norm, sigma_c, 0};
  int status = 0;
  double cp_in[3] = {cp_inG(0), cp_inG(1), cp_inG(2)}, cp_out[3];
  std::vector<double> p_out(3 * (size_t)(F > 0 ? F : 1));
  std::vector<int> inl((size_t)(F > 0 ? F : 1));
  ck(ovp_optimize_plane(ctx, 1, offs, meas_offset.data(), meas_clone.data(), uvn.data(), p.data(), cp_in, &fix, &o, p_out.data(), cp_out, inl.data(),
                        &status, nullptr));
  for (int k = 0; k < 3; k++)
    cp_inG(k) = cp_out[k]; // unchanged when the solver did not converge (:431-438), refined otherwise (:441-445)
  std::vector<std::shared_ptr<ov_core::Feature>> keep;
  for (int i = 0; i < F; i++) {
    if (!inl[i])
      continue;
    for (int k = 0; k < 3; k++)
      feats[i]->p_FinG(k) = p_out[3 * (size_t)i + k]; // the inliers receive their refined position even when the call fails (:476)
    keep.push_back(feats[i]);
  }
  if (!status)
    return false;
  feats = keep;
  return true;
}
