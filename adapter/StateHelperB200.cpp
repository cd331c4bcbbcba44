// Drop-in body of ov_plane::StateHelper on top of the B200 C ABI (include/ovp.h).
//
// Add this file to the reference tree as ov_plane/src/state/StateHelperB200.cpp and build it INSTEAD of StateHelper.cpp
// (ov_plane/CMakeLists.txt:84-110); link with -lovp.  Every function keeps the reference's signature
// (ov_plane/src/state/StateHelper.h:82-245) so VioManager, the updaters and the ROS visualiser compile unchanged.  The covariance
// and the variable table live on the device inside State::ctx; the ov_type objects remain the host-side view of the values and are
// refreshed from the device after every call that changes them (one small device->host copy), instead of the reference's
// `var->update(dx)` loop (StateHelper.cpp:190-193).
//
// In this repository the file is only SYNTAX-CHECKED against minimal stand-in headers (adapter/stubs/, tests/test_cpu_adapter.py):
// Eigen and OpenVINS ov_core are not available in the build container.
#include "state/StateHelper.h"
#include "utils/print.h"
#include <cstdlib>

using namespace ov_plane;
using ov_type::Type;

namespace {
void ck(const std::shared_ptr<State> &s, int st) {
  if (st != OVP_OK) { // the reference prints and calls std::exit(EXIT_FAILURE) (StateHelper.cpp:46-49, 116-118, 185-187, ...)
    PRINT_ERROR(RED "StateHelper (B200): %s\n" RESET, ovp_last_error(s->ctx));
    std::exit(EXIT_FAILURE);
  }
}
std::vector<int> handles(const std::shared_ptr<State> &s, const std::vector<std::shared_ptr<Type>> &vars) {
  std::vector<int> h;
  for (auto &v : vars)
    h.push_back(s->handle.at(v.get()));
  return h;
}
int kind_of(const std::shared_ptr<Type> &v) {
  if (std::dynamic_pointer_cast<ov_type::IMU>(v))
    return OVP_KIND_IMU;
  if (std::dynamic_pointer_cast<ov_type::PoseJPL>(v))
    return OVP_KIND_POSE;
  if (std::dynamic_pointer_cast<ov_type::Landmark>(v))
    return OVP_KIND_LANDMARK;
  return OVP_KIND_VEC;
}
// ids (Type::id()) and values of every variable <- device
void refresh(const std::shared_ptr<State> &s, bool values) {
  double v[16], f[16];
  for (auto &kv : s->by_handle) {
    kv.second->set_local_id(ovp_var_id(s->ctx, kv.first));
    if (!values)
      continue;
    ck(s, ovp_var_get(s->ctx, kv.first, v, f));
    const int n = ovp_var_value_size(s->ctx, kv.first);
    Eigen::MatrixXd val(n, 1), fej(n, 1);
    for (int i = 0; i < n; i++) {
      val(i, 0) = v[i];
      fej(i, 0) = f[i];
    }
    kv.second->set_value(val);
    kv.second->set_fej(fej);
  }
}
void adopt(const std::shared_ptr<State> &s, const std::shared_ptr<Type> &v, int h) {
  s->handle[v.get()] = h;
  s->by_handle[h] = v;
  v->set_local_id(ovp_var_id(s->ctx, h));
}
void forget(const std::shared_ptr<State> &s, const std::shared_ptr<Type> &v) {
  auto it = s->handle.find(v.get());
  if (it != s->handle.end()) {
    s->by_handle.erase(it->second);
    s->handle.erase(it);
  }
  v->set_local_id(-1); // StateHelper.cpp:340
}
} // namespace

void StateHelper::EKFPropagation(std::shared_ptr<State> state, const std::vector<std::shared_ptr<Type>> &order_NEW,
                                 const std::vector<std::shared_ptr<Type>> &order_OLD, const Eigen::MatrixXd &Phi, const Eigen::MatrixXd &Q) {
  auto hn = handles(state, order_NEW), ho = handles(state, order_OLD);
  ck(state, ovp_ekf_propagation(state->ctx, hn.data(), (int)hn.size(), ho.data(), (int)ho.size(), Phi.data(), (int)Phi.rows(), (int)Phi.cols(),
                                Q.data())); // Eigen is column-major like the ABI: no copy
}

void StateHelper::EKFUpdate(std::shared_ptr<State> state, const std::vector<std::shared_ptr<Type>> &H_order, const Eigen::MatrixXd &H,
                            const Eigen::VectorXd &res, const Eigen::MatrixXd &R) {
  auto h = handles(state, H_order);
  // every caller passes a diagonal R (identity from the MSCKF / SLAM / plane updaters, a non-identity diagonal from ZUPT)
  Eigen::VectorXd Rd = R.diagonal();
  bool identity = true;
  for (long i = 0; i < Rd.rows(); i++)
    identity = identity && Rd(i) == 1.0;
  ck(state, ovp_ekf_update(state->ctx, h.data(), (int)h.size(), H.data(), (int)H.rows(), res.data(), identity ? nullptr : Rd.data()));
  refresh(state, true); // values after `var->update(dx)` (StateHelper.cpp:190-193); camera objects are refreshed by the caller's hook (:197-201)
}

void StateHelper::set_initial_covariance(std::shared_ptr<State> state, const Eigen::MatrixXd &covariance,
                                         const std::vector<std::shared_ptr<Type>> &order) {
  auto h = handles(state, order);
  ck(state, ovp_set_initial_covariance(state->ctx, covariance.data(), (int)covariance.rows(), h.data(), (int)h.size()));
}

Eigen::MatrixXd StateHelper::get_marginal_covariance(std::shared_ptr<State> state, const std::vector<std::shared_ptr<Type>> &small_variables) {
  auto h = handles(state, small_variables);
  int n = 0;
  for (auto &v : small_variables)
    n += v->size();
  Eigen::MatrixXd out(n, n);
  ck(state, ovp_get_marginal_covariance(state->ctx, h.data(), (int)h.size(), out.data()));
  return out;
}

Eigen::MatrixXd StateHelper::get_full_covariance(std::shared_ptr<State> state) {
  const int n = ovp_cov_rows(state->ctx);
  Eigen::MatrixXd out(n, n);
  ck(state, ovp_cov_download(state->ctx, out.data(), n));
  return out;
}

void StateHelper::marginalize(std::shared_ptr<State> state, std::shared_ptr<Type> marg) {
  ck(state, ovp_marginalize(state->ctx, state->handle.at(marg.get()))); // the ids of later variables shift exactly as :325-334
  forget(state, marg);
  refresh(state, false);
}

std::shared_ptr<Type> StateHelper::clone(std::shared_ptr<State> state, std::shared_ptr<Type> variable_to_clone) {
  int h = -1;
  ck(state, ovp_clone(state->ctx, state->handle.at(variable_to_clone.get()), &h));
  std::shared_ptr<Type> c;
  if (std::dynamic_pointer_cast<ov_type::PoseJPL>(variable_to_clone))
    c = std::make_shared<ov_type::PoseJPL>();
  else
    c = std::make_shared<ov_type::Vec>(variable_to_clone->size());
  c->set_value(variable_to_clone->value());
  c->set_fej(variable_to_clone->fej());
  adopt(state, c, h);
  return c;
}

bool StateHelper::initialize(std::shared_ptr<State> state, std::shared_ptr<Type> new_variable, const std::vector<std::shared_ptr<Type>> &H_order,
                             Eigen::MatrixXd &H_R, Eigen::MatrixXd &H_L, Eigen::MatrixXd &R, Eigen::VectorXd &res, double chi_2_mult,
                             bool do_update) {
  auto h = handles(state, H_order);
  int accepted = 0, nh = -1;
  size_t tag = 0;
  if (auto lm = std::dynamic_pointer_cast<ov_type::Landmark>(new_variable))
    tag = lm->_featid;
  // R must be isotropic (StateHelper.cpp:413-425): the ABI takes sigma^2 and reports OVP_ERR_NON_ISOTROPIC otherwise
  ck(state, ovp_initialize(state->ctx, kind_of(new_variable), new_variable->size(), new_variable->value().data(), new_variable->fej().data(),
                           (int64_t)tag, h.data(), (int)h.size(), H_R.data(), H_L.data(), res.data(), (int)res.rows(), R(0, 0), chi_2_mult,
                           do_update ? 1 : 0, &accepted, &nh));
  if (!accepted)
    return false; // chi2 test failed: the state is untouched (:473-475)
  adopt(state, new_variable, nh);
  refresh(state, true);
  return true;
}

void StateHelper::initialize_invertible(std::shared_ptr<State> state, std::shared_ptr<Type> new_variable,
                                        const std::vector<std::shared_ptr<Type>> &H_order, const Eigen::MatrixXd &H_R, const Eigen::MatrixXd &H_L,
                                        const Eigen::MatrixXd &R, const Eigen::VectorXd &res) {
  auto h = handles(state, H_order);
  int nh = -1;
  ck(state, ovp_initialize_invertible(state->ctx, kind_of(new_variable), new_variable->size(), new_variable->value().data(),
                                      new_variable->fej().data(), 0, h.data(), (int)h.size(), H_R.data(), H_L.data(), res.data(), R(0, 0),
                                      &nh)); // rows == size of the new variable (H_L is square and invertible, StateHelper.cpp:489-586)
  adopt(state, new_variable, nh);
  refresh(state, true);
}

void StateHelper::augment_clone(std::shared_ptr<State> state, Eigen::Matrix<double, 3, 1> last_w) {
  int h = -1;
  ck(state, ovp_augment_clone(state->ctx, state->_timestamp, last_w.data(), &h)); // OVP_ERR_TIME: a clone at this time exists (:592-596)
  auto pose = std::make_shared<ov_type::PoseJPL>();
  adopt(state, pose, h);
  state->_clones_IMU[state->_timestamp] = pose;
  refresh(state, true);
}

void StateHelper::marginalize_old_clone(std::shared_ptr<State> state) {
  if ((int)state->_clones_IMU.size() <= state->_options.max_clone_size)
    return;
  const double t = state->margtimestep();
  StateHelper::marginalize(state, state->_clones_IMU.at(t)); // :627-636
  state->_clones_IMU.erase(t);
}

void StateHelper::marginalize_slam(std::shared_ptr<State> state) {
  ck(state, ovp_marginalize_slam(state->ctx)); // every landmark with should_marg, ArUco ids excluded (:638-652)
  for (auto it = state->_features_SLAM.begin(); it != state->_features_SLAM.end();) {
    if (ovp_slam_handle(state->ctx, (int64_t)it->first) < 0) {
      forget(state, it->second);
      state->_features_SLAM_to_PLANE.erase(it->first);
      it = state->_features_SLAM.erase(it);
    } else {
      ++it;
    }
  }
  refresh(state, false);
}

void StateHelper::merge_planes_and_marginalize(std::shared_ptr<State> state, const std::map<size_t, size_t> &feat2plane,
                                               const std::map<size_t, std::set<size_t>> &plane2oldplane) {
  std::vector<int64_t> ff, fp, mn, mo;
  for (auto &kv : feat2plane) {
    ff.push_back((int64_t)kv.first);
    fp.push_back((int64_t)kv.second);
  }
  for (auto &kv : plane2oldplane)
    for (size_t old : kv.second) {
      mn.push_back((int64_t)kv.first);
      mo.push_back((int64_t)old);
    }
  ck(state, ovp_merge_planes_and_marginalize(state->ctx, ff.data(), fp.data(), (int)ff.size(), mn.data(), mo.data(), (int)mn.size()));
  // mirror the renames / removals of _features_PLANE (StateHelper.cpp:654-758)
  std::unordered_map<size_t, std::shared_ptr<ov_type::Vec>> kept;
  for (auto &kv : state->_features_PLANE) {
    size_t id = kv.first;
    for (auto &m : plane2oldplane)
      if (m.second.count(id) && ovp_plane_handle(state->ctx, (int64_t)id) < 0 && ovp_plane_handle(state->ctx, (int64_t)m.first) >= 0 &&
          !state->_features_PLANE.count(m.first))
        id = m.first; // renamed old -> new (the new plane was not in the state)
    if (ovp_plane_handle(state->ctx, (int64_t)id) >= 0)
      kept[id] = kv.second;
    else
      forget(state, kv.second);
  }
  state->_features_PLANE.swap(kept);
  refresh(state, true);
}
