// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).  Flat C API over the CPU restatement so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can drive it through ctypes.
// Nothing under ov_plane_b200/ or include/ may link or load this library.
#include "oracle.hpp"
#include "oracle_planefit.hpp"
#include <string>

using namespace orc;

namespace {
struct Ctx {
  StateP state;
  Chi2Table chi2;
  std::string last_error;
  Propagator prop;
  UpdaterZeroVelocity zupt;
  MsckfResult last_msckf;
};
inline Mat from_colmajor(const double *p, int r, int c) {
  Mat m(r, c);
  if (r * c > 0)
    std::memcpy(m.a.data(), p, sizeof(double) * (size_t)r * c);
  return m;
}
inline std::vector<VarP> handles_to_vars(Ctx *c, const int *h, int k) {
  std::vector<VarP> v;
  for (int i = 0; i < k; i++)
    v.push_back(c->state->by_handle.at(h[i]));
  return v;
}
template <class F> int guarded(Ctx *c, F f) {
  try {
    f();
    return 0;
  } catch (const OracleExit &e) {
    c->last_error = e.what();
    return 2;
  } catch (const std::exception &e) {
    c->last_error = e.what();
    return 1;
  }
}
} // namespace

extern "C" {

// test instrumentation (oracle.hpp GaugeProbe): process-global, like the reference's single-threaded update path
void orc_set_gauge_probe(double (*fn)(int, int, const double *, const double *), int gate_without) {
  gauge_probe().fn = fn;
  gauge_probe().gate_without = gate_without != 0;
  gauge_probe().junk.clear();
}
int orc_gauge_probe_values(double *out, int cap) {
  auto &j = gauge_probe().junk;
  int n = (int)j.size();
  for (int i = 0; i < n && i < cap; i++)
    out[i] = j[i];
  j.clear();
  return n;
}

void *orc_create(int do_fej, int use_rk4, int imu_avg, int calib_pose, int calib_intr, int calib_dt, int max_clone_size,
                 double sigma_constraint, double const_init_multi, double const_init_chi2) {
  StateOptions o;
  o.do_fej = do_fej;
  o.use_rk4_integration = use_rk4;
  o.imu_avg = imu_avg;
  o.do_calib_camera_pose = calib_pose;
  o.do_calib_camera_intrinsics = calib_intr;
  o.do_calib_camera_timeoffset = calib_dt;
  o.max_clone_size = max_clone_size;
  o.sigma_constraint = sigma_constraint;
  o.const_init_multi = const_init_multi;
  o.const_init_chi2 = const_init_chi2;
  Ctx *c = new Ctx();
  c->state = std::make_shared<State>(o);
  return c;
}
void orc_destroy(void *p) { delete (Ctx *)p; }
const char *orc_last_error(void *p) { return ((Ctx *)p)->last_error.c_str(); }
void orc_set_chi2_table(void *p, const double *q, int n) { ((Ctx *)p)->chi2.q.assign(q, q + n); }
void orc_set_plane_merge_options(void *p, double sigma, double chi2, double deg) {
  auto &o = ((Ctx *)p)->state->_options;
  o.sigma_plane_merge = sigma;
  o.plane_merge_chi2 = chi2;
  o.plane_merge_deg_max = deg;
}

int orc_cov_rows(void *p) { return ((Ctx *)p)->state->_Cov.rows(); }
void orc_get_cov(void *p, double *out) {
  Mat &P = ((Ctx *)p)->state->_Cov;
  std::memcpy(out, P.a.data(), sizeof(double) * P.a.size());
}
// test helper: overwrite the covariance (must be n x n with n == current rows)
int orc_set_cov(void *p, const double *in, int n) {
  Ctx *c = (Ctx *)p;
  if (n != c->state->_Cov.rows())
    return 1;
  std::memcpy(c->state->_Cov.a.data(), in, sizeof(double) * (size_t)n * n);
  return 0;
}
int orc_handle_imu(void *p) { return ((Ctx *)p)->state->_imu->handle; }
int orc_handle_dt(void *p) { return ((Ctx *)p)->state->_calib_dt_CAMtoIMU->handle; }
int orc_handle_calib(void *p) { return ((Ctx *)p)->state->_calib_IMUtoCAM->handle; }
int orc_handle_intr(void *p) { return ((Ctx *)p)->state->_cam_intrinsics->handle; }
int orc_var_id(void *p, int h) { return ((Ctx *)p)->state->by_handle.at(h)->id; }
int orc_var_size(void *p, int h) { return ((Ctx *)p)->state->by_handle.at(h)->sz; }
int orc_var_nvalue(void *p, int h) { return (int)((Ctx *)p)->state->by_handle.at(h)->value.size(); }
void orc_var_set(void *p, int h, const double *value, const double *fej) {
  VarP v = ((Ctx *)p)->state->by_handle.at(h);
  if (value)
    std::copy(value, value + v->value.size(), v->value.begin());
  if (fej)
    std::copy(fej, fej + v->fej.size(), v->fej.begin());
}
void orc_var_get(void *p, int h, double *value, double *fej) {
  VarP v = ((Ctx *)p)->state->by_handle.at(h);
  if (value)
    std::copy(v->value.begin(), v->value.end(), value);
  if (fej)
    std::copy(v->fej.begin(), v->fej.end(), fej);
}
int orc_num_variables(void *p) { return (int)((Ctx *)p)->state->_variables.size(); }
void orc_variable_order(void *p, int *handles) {
  auto &v = ((Ctx *)p)->state->_variables;
  for (size_t i = 0; i < v.size(); i++)
    handles[i] = v[i]->handle;
}
void orc_set_timestamp(void *p, double t) { ((Ctx *)p)->state->_timestamp = t; }
double orc_get_timestamp(void *p) { return ((Ctx *)p)->state->_timestamp; }

// ---- raw construction helpers (tests build a synthetic state, then overwrite the covariance) -----------------
static int add_raw(Ctx *c, VarP v) {
  c->state->reg(v);
  int N = c->state->_Cov.rows();
  c->state->_Cov.conservativeResize(N + v->sz, N + v->sz);
  v->id = N;
  c->state->_variables.push_back(v);
  return v->handle;
}
int orc_add_clone_raw(void *p, double timestamp, const double *value7, const double *fej7) {
  Ctx *c = (Ctx *)p;
  VarP v = Var::makePose();
  std::copy(value7, value7 + 7, v->value.begin());
  std::copy(fej7, fej7 + 7, v->fej.begin());
  int h = add_raw(c, v);
  c->state->_clones_IMU[timestamp] = v;
  return h;
}
int orc_add_plane_raw(void *p, long long planeid, const double *cp, const double *cp_fej) {
  Ctx *c = (Ctx *)p;
  VarP v = Var::makeVec(3);
  std::copy(cp, cp + 3, v->value.begin());
  std::copy(cp_fej, cp_fej + 3, v->fej.begin());
  int h = add_raw(c, v);
  c->state->_features_PLANE[(size_t)planeid] = v;
  return h;
}
int orc_add_slam_raw(void *p, long long featid, const double *pf, const double *pf_fej) {
  Ctx *c = (Ctx *)p;
  VarP v = Var::makeLandmark(3);
  std::copy(pf, pf + 3, v->value.begin());
  std::copy(pf_fej, pf_fej + 3, v->fej.begin());
  v->featid = (size_t)featid;
  int h = add_raw(c, v);
  c->state->_features_SLAM[(size_t)featid] = v;
  return h;
}

// ---- StateHelper ----------------------------------------------------------------------------------------------
int orc_get_marginal_covariance(void *p, const int *handles, int k, double *out) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    Mat m = StateHelper::get_marginal_covariance(c->state, handles_to_vars(c, handles, k));
    std::memcpy(out, m.a.data(), sizeof(double) * m.a.size());
  });
}
int orc_set_initial_covariance(void *p, const double *cov, int n, const int *handles, int k) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] { StateHelper::set_initial_covariance(c->state, from_colmajor(cov, n, n), handles_to_vars(c, handles, k)); });
}
int orc_ekf_propagation(void *p, const int *new_h, int kn, const int *old_h, int ko, const double *Phi, int phi_rows, int phi_cols,
                        const double *Q) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    StateHelper::EKFPropagation(c->state, handles_to_vars(c, new_h, kn), handles_to_vars(c, old_h, ko),
                                from_colmajor(Phi, phi_rows, phi_cols), from_colmajor(Q, phi_rows, phi_rows));
  });
}
// H: rows x n col-major (ld = rows); Rdiag: NULL => identity
int orc_ekf_update(void *p, const int *handles, int k, const double *H, int rows, const double *res, const double *Rdiag) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    auto order = handles_to_vars(c, handles, k);
    int n = 0;
    for (auto &v : order)
      n += v->sz;
    Mat R = Mat::Identity(rows);
    if (Rdiag)
      for (int i = 0; i < rows; i++)
        R(i, i) = Rdiag[i];
    StateHelper::EKFUpdate(c->state, order, from_colmajor(H, rows, n), from_colmajor(res, rows, 1), R);
  });
}
int orc_marginalize(void *p, int h) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    VarP v = c->state->by_handle.at(h);
    StateHelper::marginalize(c->state, v);
    for (auto it = c->state->_clones_IMU.begin(); it != c->state->_clones_IMU.end(); ++it)
      if (it->second == v) {
        c->state->_clones_IMU.erase(it);
        break;
      }
    for (auto it = c->state->_features_PLANE.begin(); it != c->state->_features_PLANE.end(); ++it)
      if (it->second == v) {
        c->state->_features_PLANE.erase(it);
        break;
      }
    for (auto it = c->state->_features_SLAM.begin(); it != c->state->_features_SLAM.end(); ++it)
      if (it->second == v) {
        c->state->_features_SLAM.erase(it);
        break;
      }
  });
}
int orc_marginalize_old_clone(void *p) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] { StateHelper::marginalize_old_clone(c->state); });
}
// sets state timestamp (as propagate_and_clone does, Propagator.cpp:121) then augment_clone; returns new handle
int orc_augment_clone(void *p, double timestamp, const double *last_w, int *new_handle) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    c->state->_timestamp = timestamp;
    VarP v = StateHelper::augment_clone(c->state, vec3(last_w[0], last_w[1], last_w[2]));
    *new_handle = v->handle;
  });
}
// initialize a new Vec/Landmark variable of size s (3 or 1) with isotropic noise sigma2 (R = sigma2 * I)
int orc_initialize(void *p, int kind, int s, const double *value, const double *fej, long long tag, const int *handles, int k,
                   const double *H_R, const double *H_L, const double *res, int rows, double sigma2, double chi2_mult, int do_update,
                   int *accepted, int *new_handle) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    auto order = handles_to_vars(c, handles, k);
    int n = 0;
    for (auto &v : order)
      n += v->sz;
    VarP nv = (kind == KIND_LANDMARK) ? Var::makeLandmark(s) : Var::makeVec(s);
    std::copy(value, value + s, nv->value.begin());
    std::copy(fej, fej + s, nv->fej.begin());
    nv->featid = (size_t)tag;
    Mat HR = from_colmajor(H_R, rows, n), HL = from_colmajor(H_L, rows, s), r = from_colmajor(res, rows, 1);
    Mat R = sigma2 * Mat::Identity(rows);
    bool ok = StateHelper::initialize(c->state, nv, order, HR, HL, R, r, chi2_mult, c->chi2, do_update != 0);
    *accepted = ok ? 1 : 0;
    *new_handle = -1;
    if (ok) {
      c->state->reg(nv);
      *new_handle = nv->handle;
      if (kind == KIND_LANDMARK)
        c->state->_features_SLAM[(size_t)tag] = nv;
      else
        c->state->_features_PLANE[(size_t)tag] = nv;
    }
  });
}
// StateHelper::initialize_invertible called directly: H_L square (rows == size of the new variable), no chi2 test, no update
int orc_initialize_invertible(void *p, int kind, int s, const double *value, const double *fej, long long tag, const int *handles, int k,
                              const double *H_R, const double *H_L, const double *res, double sigma2, int *new_handle) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    auto order = handles_to_vars(c, handles, k);
    int n = 0;
    for (auto &v : order)
      n += v->sz;
    VarP nv = (kind == KIND_LANDMARK) ? Var::makeLandmark(s) : Var::makeVec(s);
    std::copy(value, value + s, nv->value.begin());
    std::copy(fej, fej + s, nv->fej.begin());
    nv->featid = (size_t)tag;
    StateHelper::initialize_invertible(c->state, nv, order, from_colmajor(H_R, s, n), from_colmajor(H_L, s, s), sigma2 * Mat::Identity(s),
                                       from_colmajor(res, s, 1));
    c->state->reg(nv);
    *new_handle = nv->handle;
    if (kind == KIND_LANDMARK)
      c->state->_features_SLAM[(size_t)tag] = nv;
    else
      c->state->_features_PLANE[(size_t)tag] = nv;
  });
}
int orc_merge_planes_and_marginalize(void *p, const long long *f2p_feat, const long long *f2p_plane, int nf, const long long *merge_new,
                                     const long long *merge_old, int nm) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    std::map<size_t, size_t> feat2plane;
    for (int i = 0; i < nf; i++)
      feat2plane[(size_t)f2p_feat[i]] = (size_t)f2p_plane[i];
    std::map<size_t, std::set<size_t>> plane2old;
    for (int i = 0; i < nm; i++)
      plane2old[(size_t)merge_new[i]].insert((size_t)merge_old[i]);
    StateHelper::merge_planes_and_marginalize(c->state, feat2plane, plane2old, c->chi2);
  });
}
int orc_plane_handle(void *p, long long planeid) {
  Ctx *c = (Ctx *)p;
  auto it = c->state->_features_PLANE.find((size_t)planeid);
  return it == c->state->_features_PLANE.end() ? -1 : it->second->handle;
}

// ---- stateless matrix helpers (all col-major, in place; *_rows_out receives the new row count) ----------------
void orc_nullspace_project_inplace(double *H_f, int hf_cols, double *H_x, int hx_cols, double *res, int rows, int *rows_out) {
  Mat Hf = from_colmajor(H_f, rows, hf_cols), Hx = from_colmajor(H_x, rows, hx_cols), r = from_colmajor(res, rows, 1);
  UpdaterHelper::nullspace_project_inplace(Hf, Hx, r);
  *rows_out = Hx.rows();
  std::memcpy(H_x, Hx.a.data(), sizeof(double) * Hx.a.size());
  std::memcpy(res, r.a.data(), sizeof(double) * r.a.size());
}
void orc_measurement_compress_inplace(double *H_x, int cols, double *res, int rows, int *rows_out) {
  Mat Hx = from_colmajor(H_x, rows, cols), r = from_colmajor(res, rows, 1);
  UpdaterHelper::measurement_compress_inplace(Hx, r);
  *rows_out = Hx.rows();
  std::memcpy(H_x, Hx.a.data(), sizeof(double) * Hx.a.size());
  std::memcpy(res, r.a.data(), sizeof(double) * r.a.size());
}
void orc_plane_nullspace_project_inplace(double *H_f, int hf_cols, double *H_x, int hx_cols, double *H_cp, double *res, int rows,
                                         int *rows_out) {
  Mat Hf = from_colmajor(H_f, rows, hf_cols), Hx = from_colmajor(H_x, rows, hx_cols), Hcp = from_colmajor(H_cp, rows, 3),
      r = from_colmajor(res, rows, 1);
  UpdaterPlane::nullspace_project_inplace(Hf, Hx, Hcp, r);
  *rows_out = Hx.rows();
  std::memcpy(H_x, Hx.a.data(), sizeof(double) * Hx.a.size());
  std::memcpy(H_cp, Hcp.a.data(), sizeof(double) * Hcp.a.size());
  std::memcpy(res, r.a.data(), sizeof(double) * r.a.size());
}
void orc_plane_measurement_compress_inplace(double *H_x, int cols, double *H_cp, double *res, int rows, int *rows_out) {
  Mat Hx = from_colmajor(H_x, rows, cols), Hcp = from_colmajor(H_cp, rows, 3), r = from_colmajor(res, rows, 1);
  UpdaterPlane::measurement_compress_inplace(Hx, Hcp, r);
  *rows_out = Hx.rows();
  std::memcpy(H_x, Hx.a.data(), sizeof(double) * Hx.a.size());
  std::memcpy(H_cp, Hcp.a.data(), sizeof(double) * Hcp.a.size());
  std::memcpy(res, r.a.data(), sizeof(double) * r.a.size());
}

// ---- UpdaterHelper::get_feature_jacobian_full -----------------------------------------------------------------
// clone_handles[m] identify the clone of each measurement.  Outputs are written col-major with ld = rows_out;
// buffers must hold 3*m(+1) rows; x_order receives variable handles.
int orc_feature_jacobian_representation(int rep, int do_fej, const double *p_FinG, const double *p_FinG_fej, const double *p_FinA,
                                        const double *anchor7, const double *anchor_fej7, const double *calib7, double *H_f, int *hf_cols,
                                        double *H_anc, double *H_calib, int *has_anchor) {
  try {
    Mat Hf, Ha, Hc;
    bool anc = UpdaterHelper::get_feature_jacobian_representation(rep, do_fej != 0, from_colmajor(p_FinG, 3, 1), from_colmajor(p_FinG_fej, 3, 1),
                                                                  from_colmajor(p_FinA, 3, 1), from_colmajor(anchor7, 7, 1),
                                                                  from_colmajor(anchor_fej7, 7, 1), from_colmajor(calib7, 7, 1), Hf, Ha, Hc);
    *hf_cols = Hf.cols();
    std::memcpy(H_f, Hf.a.data(), sizeof(double) * 3 * Hf.cols());
    *has_anchor = anc ? 1 : 0;
    if (anc) {
      std::memcpy(H_anc, Ha.a.data(), sizeof(double) * 18);
      std::memcpy(H_calib, Hc.a.data(), sizeof(double) * 18);
    }
    return 0;
  } catch (...) {
    return 1;
  }
}
int orc_feature_jacobian_full(void *p, int m, const int *clone_handles, const float *uv, const double *p_FinG, const double *p_FinG_fej,
                              long long planeid, const double *cp, const double *cp_fej, double sigma_px, double sigma_c, double *H_f,
                              int *hf_cols, double *H_x, int *hx_cols, double *res, int *rows_out, int *x_order, int *x_order_n) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    Feature f;
    for (int i = 0; i < m; i++) {
      VarP cl = c->state->by_handle.at(clone_handles[i]);
      double ts = NAN;
      for (auto &kv : c->state->_clones_IMU)
        if (kv.second == cl)
          ts = kv.first;
      f.timestamps.push_back(ts);
      f.uvs.push_back(uv[2 * i]);
      f.uvs.push_back(uv[2 * i + 1]);
    }
    f.p_FinG = vec3(p_FinG[0], p_FinG[1], p_FinG[2]);
    f.p_FinG_fej = vec3(p_FinG_fej[0], p_FinG_fej[1], p_FinG_fej[2]);
    f.planeid = (size_t)planeid;
    if (planeid != 0) {
      f.cp_FinG = vec3(cp[0], cp[1], cp[2]);
      f.cp_FinG_fej = vec3(cp_fej[0], cp_fej[1], cp_fej[2]);
    }
    Mat Hf, Hx, r;
    std::vector<VarP> order;
    UpdaterHelper::get_feature_jacobian_full(c->state, f, sigma_px, sigma_c, Hf, Hx, r, order);
    *hf_cols = Hf.cols();
    *hx_cols = Hx.cols();
    *rows_out = r.rows();
    std::memcpy(H_f, Hf.a.data(), sizeof(double) * Hf.a.size());
    std::memcpy(H_x, Hx.a.data(), sizeof(double) * Hx.a.size());
    std::memcpy(res, r.a.data(), sizeof(double) * r.a.size());
    *x_order_n = (int)order.size();
    for (size_t i = 0; i < order.size(); i++)
      x_order[i] = order[i]->handle;
  });
}

// ---- UpdaterMSCKF::update core --------------------------------------------------------------------------------
// SoA batch: meas_offset[F+1]; meas_clone[Σm] clone handles; uv[2Σm] float; p_FinG / p_FinG_original [3F];
// planeid[F] (0 = not on a plane, i.e. not in feat2plane); plane_est_ids/plane_est_cp: the planes that got a
// linearisation point (UpdaterMSCKF.cpp:198-404), in-state planes use the state value (cp entries ignored).
// Outputs: feat_status[F]: 1 accepted (point path), 0 chi2-rejected, 2 consumed by a passed plane update;
// plane_status[nplanes]: 1 pass, 0 fail, -1 not visited.
int orc_msckf_update(void *p, int F, const int *meas_offset, const int *meas_clone, const float *uv, const double *p_FinG,
                     const double *p_FinG_original, const long long *featid, const long long *planeid, int nplanes,
                     const long long *plane_est_ids, const double *plane_est_cp, double sigma_pix, double chi2_mult, int *feat_status,
                     double *feat_chi2, int *plane_status, double *plane_chi2, int *hx_order, int *hx_order_n, double *timers4) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    std::vector<Feature> fv(F);
    std::map<size_t, size_t> feat2plane;
    std::unordered_map<Var *, double> clone_ts;
    for (auto &kv : c->state->_clones_IMU)
      clone_ts[kv.second.get()] = kv.first;
    for (int i = 0; i < F; i++) {
      Feature &f = fv[i];
      f.featid = (size_t)featid[i];
      for (int k = meas_offset[i]; k < meas_offset[i + 1]; k++) {
        f.timestamps.push_back(clone_ts.at(c->state->by_handle.at(meas_clone[k]).get()));
        f.uvs.push_back(uv[2 * k]);
        f.uvs.push_back(uv[2 * k + 1]);
      }
      f.p_FinG = vec3(p_FinG[3 * i], p_FinG[3 * i + 1], p_FinG[3 * i + 2]);
      f.p_FinG_original = vec3(p_FinG_original[3 * i], p_FinG_original[3 * i + 1], p_FinG_original[3 * i + 2]);
      if (planeid[i] != 0)
        feat2plane[f.featid] = (size_t)planeid[i];
    }
    std::map<size_t, Mat> plane_est;
    for (int i = 0; i < nplanes; i++) {
      size_t pid = (size_t)plane_est_ids[i];
      auto it = c->state->_features_PLANE.find(pid);
      if (it != c->state->_features_PLANE.end())
        plane_est[pid] = it->second->vecvalue(false);
      else
        plane_est[pid] = vec3(plane_est_cp[3 * i], plane_est_cp[3 * i + 1], plane_est_cp[3 * i + 2]);
    }
    UpdaterMSCKF up;
    up.sigma_pix = sigma_pix;
    up.chi2_multipler = chi2_mult;
    up.chi2tab = c->chi2;
    MsckfResult r = up.update(c->state, fv, feat2plane, plane_est);
    std::unordered_map<size_t, int> idx;
    for (int i = 0; i < F; i++) {
      idx[(size_t)featid[i]] = i;
      feat_status[i] = -1;
      if (feat_chi2)
        feat_chi2[i] = NAN;
    }
    for (size_t k = 0; k < r.feat_status.size(); k++) {
      int i = idx.at(r.feat_status[k].first);
      feat_status[i] = r.feat_status[k].second;
      if (feat_chi2)
        feat_chi2[i] = r.feat_chi2[k];
    }
    for (size_t fid : r.used_plane_featids)
      feat_status[idx.at(fid)] = 2;
    for (int i = 0; i < nplanes; i++) {
      plane_status[i] = -1;
      if (plane_chi2)
        plane_chi2[i] = NAN;
    }
    for (size_t k = 0; k < r.plane_status.size(); k++)
      for (int i = 0; i < nplanes; i++)
        if ((size_t)plane_est_ids[i] == r.plane_status[k].first) {
          plane_status[i] = r.plane_status[k].second;
          if (plane_chi2)
            plane_chi2[i] = r.plane_chi2[k];
        }
    *hx_order_n = (int)r.Hx_order_handles.size();
    for (size_t i = 0; i < r.Hx_order_handles.size(); i++)
      hx_order[i] = r.Hx_order_handles[i];
    if (timers4) {
      timers4[0] = r.t.plane_updates;
      timers4[1] = r.t.feat_system;
      timers4[2] = r.t.compression;
      timers4[3] = r.t.update;
    }
    c->last_msckf = r;
  });
}

// ---- UpdaterPlane::init_vio_plane core ------------------------------------------------------------------------------
int orc_plane_init(void *p, int F, const int *meas_offset, const int *meas_clone, const float *uv, const double *p_FinG,
                   const long long *featid, const long long *planeid, int nplanes, const long long *plane_est_ids,
                   const double *plane_est_cp, double sigma_pix, int *plane_status, int *new_handles) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    std::vector<Feature> fv(F);
    std::map<size_t, size_t> feat2plane;
    std::unordered_map<Var *, double> clone_ts;
    for (auto &kv : c->state->_clones_IMU)
      clone_ts[kv.second.get()] = kv.first;
    for (int i = 0; i < F; i++) {
      Feature &f = fv[i];
      f.featid = (size_t)featid[i];
      for (int k = meas_offset[i]; k < meas_offset[i + 1]; k++) {
        f.timestamps.push_back(clone_ts.at(c->state->by_handle.at(meas_clone[k]).get()));
        f.uvs.push_back(uv[2 * k]);
        f.uvs.push_back(uv[2 * k + 1]);
      }
      f.p_FinG = vec3(p_FinG[3 * i], p_FinG[3 * i + 1], p_FinG[3 * i + 2]);
      if (planeid[i] != 0)
        feat2plane[f.featid] = (size_t)planeid[i];
    }
    std::map<size_t, Mat> plane_est;
    for (int i = 0; i < nplanes; i++) {
      plane_est[(size_t)plane_est_ids[i]] = vec3(plane_est_cp[3 * i], plane_est_cp[3 * i + 1], plane_est_cp[3 * i + 2]);
      plane_status[i] = -1;
      new_handles[i] = -1;
    }
    UpdaterPlaneInit up;
    up.sigma_pix = sigma_pix;
    up.chi2tab = c->chi2;
    PlaneInitResult r = up.init_vio_plane(c->state, fv, feat2plane, plane_est);
    for (size_t k = 0; k < r.plane_status.size(); k++)
      for (int i = 0; i < nplanes; i++)
        if ((size_t)plane_est_ids[i] == r.plane_status[k].first) {
          plane_status[i] = r.plane_status[k].second;
          new_handles[i] = r.new_handles[k];
        }
  });
}

// ---- UpdaterSLAM cores -----------------------------------------------------------------------------------------------
static void fill_features(Ctx *c, int F, const int *meas_offset, const int *meas_clone, const float *uv, const double *p_FinG,
                          const double *p_FinG_original, const long long *featid, const long long *planeid, std::vector<Feature> &fv,
                          std::map<size_t, size_t> &feat2plane) {
  std::unordered_map<Var *, double> clone_ts;
  for (auto &kv : c->state->_clones_IMU)
    clone_ts[kv.second.get()] = kv.first;
  fv.resize(F);
  for (int i = 0; i < F; i++) {
    Feature &f = fv[i];
    f.featid = (size_t)featid[i];
    for (int k = meas_offset[i]; k < meas_offset[i + 1]; k++) {
      f.timestamps.push_back(clone_ts.at(c->state->by_handle.at(meas_clone[k]).get()));
      f.uvs.push_back(uv[2 * k]);
      f.uvs.push_back(uv[2 * k + 1]);
    }
    if (p_FinG)
      f.p_FinG = vec3(p_FinG[3 * i], p_FinG[3 * i + 1], p_FinG[3 * i + 2]);
    if (p_FinG_original)
      f.p_FinG_original = vec3(p_FinG_original[3 * i], p_FinG_original[3 * i + 1], p_FinG_original[3 * i + 2]);
    if (planeid[i] != 0)
      feat2plane[f.featid] = (size_t)planeid[i];
  }
}
int orc_slam_update(void *p, int F, const int *meas_offset, const int *meas_clone, const float *uv, const long long *featid,
                    const long long *planeid, double sigma_pix, double chi2_mult, int *feat_status, double *feat_chi2) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    std::vector<Feature> fv;
    std::map<size_t, size_t> f2p;
    fill_features(c, F, meas_offset, meas_clone, uv, nullptr, nullptr, featid, planeid, fv, f2p);
    UpdaterSLAM up;
    up.opt.sigma_pix = sigma_pix;
    up.opt.chi2_multipler = chi2_mult;
    up.chi2tab = c->chi2;
    std::vector<double> chi;
    std::vector<int> st = up.update(c->state, fv, f2p, &chi);
    for (int i = 0; i < F; i++) {
      feat_status[i] = st[i];
      if (feat_chi2)
        feat_chi2[i] = chi[i];
    }
  });
}
int orc_slam_delayed_init(void *p, int F, const int *meas_offset, const int *meas_clone, const float *uv, const double *p_FinG,
                          const double *p_FinG_original, const long long *featid, const long long *planeid, double sigma_pix,
                          double chi2_mult, int *feat_status, int *new_handles) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    std::vector<Feature> fv;
    std::map<size_t, size_t> f2p;
    fill_features(c, F, meas_offset, meas_clone, uv, p_FinG, p_FinG_original, featid, planeid, fv, f2p);
    UpdaterSLAM up;
    up.opt.sigma_pix = sigma_pix;
    up.opt.chi2_multipler = chi2_mult;
    up.chi2tab = c->chi2;
    std::vector<int> hs;
    std::vector<int> st = up.delayed_init(c->state, fv, f2p, &hs);
    for (int i = 0; i < F; i++) {
      feat_status[i] = st[i];
      new_handles[i] = hs[i];
    }
  });
}
int orc_marginalize_slam(void *p) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] { StateHelper::marginalize_slam(c->state); });
}
int orc_slam_handle(void *p, long long featid) {
  Ctx *c = (Ctx *)p;
  auto it = c->state->_features_SLAM.find((size_t)featid);
  return it == c->state->_features_SLAM.end() ? -1 : it->second->handle;
}
int orc_slam_should_marg(void *p, long long featid) {
  Ctx *c = (Ctx *)p;
  auto it = c->state->_features_SLAM.find((size_t)featid);
  return it == c->state->_features_SLAM.end() ? -1 : (it->second->should_marg ? 1 : 0);
}

// ---- anchored representations: get_feature_jacobian_full with a representation, Landmark bookkeeping, anchor change ----
static double clone_time_of(Ctx *c, int handle) {
  VarP cl = c->state->by_handle.at(handle);
  for (auto &kv : c->state->_clones_IMU)
    if (kv.second == cl)
      return kv.first;
  throw std::runtime_error("handle is not a clone");
}
int orc_feature_jacobian_full_rep(void *p, int m, const int *clone_handles, const float *uv, int representation, int anchor_clone_handle,
                                  const double *p_F, const double *p_F_fej, double sigma_px, double *H_f, int *hf_cols, double *H_x, int *hx_cols,
                                  double *res, int *rows_out, int *x_order, int *x_order_n) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    Feature f;
    for (int i = 0; i < m; i++) {
      f.timestamps.push_back(clone_time_of(c, clone_handles[i]));
      f.uvs.push_back(uv[2 * i]);
      f.uvs.push_back(uv[2 * i + 1]);
    }
    f.feat_representation = representation;
    if (representation >= 2) {
      f.anchor_clone_timestamp = clone_time_of(c, anchor_clone_handle);
      f.p_FinA = vec3(p_F[0], p_F[1], p_F[2]);
      f.p_FinA_fej = vec3(p_F_fej[0], p_F_fej[1], p_F_fej[2]);
    } else {
      f.p_FinG = vec3(p_F[0], p_F[1], p_F[2]);
      f.p_FinG_fej = vec3(p_F_fej[0], p_F_fej[1], p_F_fej[2]);
    }
    Mat Hf, Hx, r;
    std::vector<VarP> order;
    UpdaterHelper::get_feature_jacobian_full(c->state, f, sigma_px, 1.0, Hf, Hx, r, order);
    *hf_cols = Hf.cols();
    *hx_cols = Hx.cols();
    *rows_out = r.rows();
    std::memcpy(H_f, Hf.a.data(), sizeof(double) * Hf.a.size());
    std::memcpy(H_x, Hx.a.data(), sizeof(double) * Hx.a.size());
    std::memcpy(res, r.a.data(), sizeof(double) * r.a.size());
    *x_order_n = (int)order.size();
    for (size_t i = 0; i < order.size(); i++)
      x_order[i] = order[i]->handle;
  });
}
int orc_slam_set_representation(void *p, long long featid, int representation, int anchor_clone_handle) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    VarP lm = c->state->_features_SLAM.at((size_t)featid);
    lm->feat_representation = representation;
    lm->anchor_clone_timestamp = representation >= 2 ? clone_time_of(c, anchor_clone_handle) : -1;
  });
}
int orc_slam_get_representation(void *p, long long featid, int *representation, int *anchor_clone_handle) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    VarP lm = c->state->_features_SLAM.at((size_t)featid);
    *representation = lm->feat_representation;
    *anchor_clone_handle = lm->feat_representation >= 2 ? c->state->_clones_IMU.at(lm->anchor_clone_timestamp)->handle : -1;
  });
}
int orc_slam_perform_anchor_change(void *p, long long featid, int new_anchor_clone_handle) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] { LandmarkOps::perform_anchor_change(c->state, c->state->_features_SLAM.at((size_t)featid), clone_time_of(c, new_anchor_clone_handle)); });
}
int orc_slam_change_anchors(void *p, int *n_changed) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] { *n_changed = LandmarkOps::change_anchors(c->state); });
}

// ---- Propagator -----------------------------------------------------------------------------------------------
void orc_prop_set(void *p, double sigma_w, double sigma_wb, double sigma_a, double sigma_ab, double gravity_mag) {
  Ctx *c = (Ctx *)p;
  c->prop._noises.sigma_w = sigma_w;
  c->prop._noises.sigma_wb = sigma_wb;
  c->prop._noises.sigma_a = sigma_a;
  c->prop._noises.sigma_ab = sigma_ab;
  c->prop._gravity = vec3(0, 0, gravity_mag);
}
void orc_prop_feed_imu(void *p, double t, const double *wm, const double *am) {
  Ctx *c = (Ctx *)p;
  ImuData d;
  d.timestamp = t;
  d.wm = vec3(wm[0], wm[1], wm[2]);
  d.am = vec3(am[0], am[1], am[2]);
  c->prop.imu_data.push_back(d);
}
int orc_prop_propagate_and_clone(void *p, double timestamp, double *Phi15, double *Q15, int *new_handle) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    Mat Phi, Q;
    c->prop.propagate_and_clone(c->state, timestamp, &Phi, &Q);
    if (Phi15)
      std::memcpy(Phi15, Phi.a.data(), sizeof(double) * 225);
    if (Q15)
      std::memcpy(Q15, Q.a.data(), sizeof(double) * 225);
    *new_handle = c->state->_clones_IMU.at(timestamp)->handle;
  });
}

int orc_prop_fast_state_propagate(void *p, double timestamp, double *state_plus13, double *cov144, int *ok) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    Mat sp, cv;
    *ok = c->prop.fast_state_propagate(c->state, timestamp, sp, cv) ? 1 : 0;
    if (*ok) {
      std::memcpy(state_plus13, sp.a.data(), sizeof(double) * 13);
      std::memcpy(cov144, cv.a.data(), sizeof(double) * 144);
    }
  });
}

// ---- small math exports for unit tests of the restated ov_core pieces ------------------------------------------
void orc_quat_2_Rot(const double *q, double *R) {
  Mat r = quat_2_Rot(from_colmajor(q, 4, 1));
  std::memcpy(R, r.a.data(), 72);
}
void orc_rot_2_quat(const double *R, double *q) {
  Mat r = rot_2_quat(from_colmajor(R, 3, 3));
  std::memcpy(q, r.a.data(), 32);
}
void orc_quat_multiply(const double *q, const double *p, double *out) {
  Mat r = quat_multiply(from_colmajor(q, 4, 1), from_colmajor(p, 4, 1));
  std::memcpy(out, r.a.data(), 32);
}
void orc_exp_so3(const double *w, double *R) {
  Mat r = exp_so3(from_colmajor(w, 3, 1));
  std::memcpy(R, r.a.data(), 72);
}
void orc_Jr_so3(const double *w, double *R) {
  Mat r = Jr_so3(from_colmajor(w, 3, 1));
  std::memcpy(R, r.a.data(), 72);
}
void orc_radtan_distort(const double *cam, double x, double y, double *uv) { radtan_distort_d(cam, x, y, uv[0], uv[1]); }
void orc_radtan_jacobian(const double *cam, double x, double y, double *dzn4, double *dzeta16) {
  Mat a, b;
  radtan_distort_jacobian(cam, x, y, a, b);
  std::memcpy(dzn4, a.a.data(), 32);
  std::memcpy(dzeta16, b.a.data(), 128);
}
void orc_make_givens(double p, double q, double *cs) {
  Givens g;
  g.make(p, q);
  cs[0] = g.c;
  cs[1] = g.s;
}
void orc_var_update(int kind, double *value, const double *dx) {
  Var v;
  v.kind = (Kind)kind;
  int nv = kind == KIND_IMU ? 16 : (kind == KIND_POSE ? 7 : 3);
  v.sz = kind == KIND_IMU ? 15 : (kind == KIND_POSE ? 6 : 3);
  v.value.assign(value, value + nv);
  v.update(dx);
  std::copy(v.value.begin(), v.value.end(), value);
}

// ---- UpdaterZeroVelocity ----
void orc_zupt_set(void *p, double sigma_w, double sigma_wb, double sigma_a, double sigma_ab, double gravity_mag, double max_velocity,
                  double noise_multiplier, double max_disparity, double chi2_mult) {
  Ctx *c = (Ctx *)p;
  c->zupt._noises.sigma_w = sigma_w;
  c->zupt._noises.sigma_wb = sigma_wb;
  c->zupt._noises.sigma_a = sigma_a;
  c->zupt._noises.sigma_ab = sigma_ab;
  c->zupt._gravity = vec3(0, 0, gravity_mag);
  c->zupt._zupt_max_velocity = max_velocity;
  c->zupt._zupt_noise_multiplier = noise_multiplier;
  c->zupt._zupt_max_disparity = max_disparity;
  c->zupt.chi2_multipler = chi2_mult;
}
void orc_zupt_feed_imu(void *p, double t, const double *wm, const double *am) {
  Ctx *c = (Ctx *)p;
  ImuData d;
  d.timestamp = t;
  d.wm = vec3(wm[0], wm[1], wm[2]);
  d.am = vec3(am[0], am[1], am[2]);
  c->zupt.imu_data.push_back(d);
}
int orc_zupt_try_update(void *p, double timestamp, double average_disparity, int num_features, int *accepted, double *chi2) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    *accepted = c->zupt.try_update(c->state, timestamp, average_disparity, num_features, c->chi2) ? 1 : 0;
    *chi2 = c->zupt.last_chi2;
  });
}

// ---- FeatureInitializer (ov_core, restated) ----
int orc_triangulate_features(void *p, int F, const int *meas_offset, const int *meas_clone, const float *uv_norm, double *p_FinG, int *status) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    FeatureInitializer fi;
    VarP calib = c->state->_calib_IMUtoCAM;
    for (int f = 0; f < F; f++) {
      std::vector<FeatureInitializer::ClonePose> cams;
      std::vector<float> uvn;
      for (int k = meas_offset[f]; k < meas_offset[f + 1]; k++) {
        VarP cl = c->state->by_handle.at(meas_clone[k]);
        FeatureInitializer::ClonePose cp;
        cp.R = calib->Rot() * cl->Rot(); // UpdaterMSCKF.cpp:131-132
        cp.p = cl->pos() - cp.R.T() * calib->pos();
        cams.push_back(cp);
        uvn.push_back(uv_norm[2 * k]);
        uvn.push_back(uv_norm[2 * k + 1]);
      }
      Mat pf(3, 1);
      status[f] = (cams.size() >= 2 && fi.triangulate_and_refine(cams, uvn, pf)) ? 1 : 0;
      for (int i = 0; i < 3; i++)
        p_FinG[3 * f + i] = status[f] ? pf(i, 0) : 0.0;
    }
  });
}
// ---- PlaneFitting (track_plane/PlaneFitting.cpp) ----
// shuffle_kind: 0 libstdc++ GCC 7..10 (the reference's documented toolchains), 1 libstdc++ GCC >= 11, 2 this build's std::shuffle
int orc_plane_shuffle(int n, int n_shuffles, int shuffle_kind, int *out) { // n_shuffles successive shuffles of 0..n-1 from std::mt19937(8888)
  std::mt19937 g(8888);
  for (int k = 0; k < n_shuffles; k++) {
    std::vector<int> v(n);
    for (int i = 0; i < n; i++)
      v[i] = i;
    PlaneFitting::shuffle(v, g, (ShuffleKind)shuffle_kind);
    for (int i = 0; i < n; i++)
      out[(size_t)k * n + i] = v[i];
  }
  return 0;
}
int orc_fit_plane(int K, const double *pts, double cond_thresh, int cond_check, double *abcd, int *ok) {
  std::vector<int> idx(K);
  for (int i = 0; i < K; i++)
    idx[i] = i;
  *ok = PlaneFitting::fit_plane(pts, idx, abcd, cond_thresh, cond_check != 0) ? 1 : 0;
  return 0;
}
int orc_plane_fitting(int F, const double *pts, int min_inlier_num, double max_cond, int shuffle_kind, double *abcd, int *inlier, int *ok) {
  std::vector<int> inl;
  *ok = PlaneFitting::plane_fitting(F, pts, abcd, min_inlier_num, max_cond, inl, (ShuffleKind)shuffle_kind) ? 1 : 0;
  for (int f = 0; f < F; f++)
    inlier[f] = inl[f];
  return 0;
}
// optimize_plane against the clone / extrinsics / IMU values of the oracle state.  info[4] = {converged, iterations, initial cost, final cost}
int orc_optimize_plane(void *p, int F, const int *meas_offset, const int *meas_clone, const float *uv_norm, const double *p_FinG,
                       const double *cp_inG, double sigma_px_norm, double sigma_c, int fix_plane, int max_num_iterations, double *p_out,
                       double *cp_out, int *inlier, int *ok, double *info) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    PlaneFitting::Problem P;
    P.F = F;
    P.meas_offset.assign(meas_offset, meas_offset + F + 1);
    VarP calib = c->state->_calib_IMUtoCAM;
    for (int k = 0; k < meas_offset[F]; k++) {
      VarP cl = c->state->by_handle.at(meas_clone[k]);
      FeatureInitializer::ClonePose cp;
      cp.R = calib->Rot() * cl->Rot(); // UpdaterMSCKF.cpp:131-132
      cp.p = cl->pos() - cp.R.T() * calib->pos();
      P.cam.push_back(cp);
      P.uvn.push_back((double)uv_norm[2 * k]);
      P.uvn.push_back((double)uv_norm[2 * k + 1]);
    }
    P.p0.assign(p_FinG, p_FinG + 3 * F);
    for (int i = 0; i < 3; i++)
      P.cp0[i] = cp_inG[i];
    P.sigma_px_norm = sigma_px_norm;
    P.sigma_c = sigma_c;
    P.fix_plane = fix_plane != 0;
    std::vector<double> pv;
    std::vector<int> inl;
    PlaneFitting::Summary S;
    VarP imu = c->state->_imu;
    *ok = PlaneFitting::optimize_plane(P, imu->Rot(), imu->pos(), calib->Rot(), calib->pos(), pv, cp_out, inl, &S, max_num_iterations > 0 ? max_num_iterations : 12) ? 1 : 0;
    for (int i = 0; i < 3 * F; i++)
      p_out[i] = pv[i];
    for (int f = 0; f < F; f++)
      inlier[f] = inl[f];
    if (info) {
      info[0] = S.converged ? 1.0 : 0.0;
      info[1] = S.iterations;
      info[2] = S.initial_cost;
      info[3] = S.final_cost;
      info[4] = S.reason;
    }
  });
}
// robustified cost 1/2 sum rho(s) of the refinement problem at given feature positions / plane (for the optimality checks of the tests)
int orc_optimize_plane_cost(void *p, int F, const int *meas_offset, const int *meas_clone, const float *uv_norm, const double *p_FinG,
                            const double *cp_inG, double sigma_px_norm, double sigma_c, int fix_plane, double *cost) {
  Ctx *c = (Ctx *)p;
  return guarded(c, [&] {
    PlaneFitting::Problem P;
    P.F = F;
    P.meas_offset.assign(meas_offset, meas_offset + F + 1);
    VarP calib = c->state->_calib_IMUtoCAM;
    for (int k = 0; k < meas_offset[F]; k++) {
      VarP cl = c->state->by_handle.at(meas_clone[k]);
      FeatureInitializer::ClonePose cp;
      cp.R = calib->Rot() * cl->Rot();
      cp.p = cl->pos() - cp.R.T() * calib->pos();
      P.cam.push_back(cp);
      P.uvn.push_back((double)uv_norm[2 * k]);
      P.uvn.push_back((double)uv_norm[2 * k + 1]);
    }
    P.p0.assign(p_FinG, p_FinG + 3 * F);
    for (int i = 0; i < 3; i++)
      P.cp0[i] = cp_inG[i];
    P.sigma_px_norm = sigma_px_norm;
    P.sigma_c = sigma_c;
    P.fix_plane = fix_plane != 0;
    PlaneFitting::layout(P);
    std::vector<double> x(P.n, 0.0);
    for (int f = 0; f < F; f++)
      if (P.feat_col[f] >= 0)
        for (int i = 0; i < 3; i++)
          x[P.feat_col[f] + i] = p_FinG[3 * f + i];
    if (P.cp_col >= 0)
      for (int i = 0; i < 3; i++)
        x[P.cp_col + i] = cp_inG[i];
    *cost = PlaneFitting::evaluate(P, x, nullptr, nullptr);
  });
}
} // extern "C"
