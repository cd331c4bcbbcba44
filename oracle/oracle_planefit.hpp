// TEST INFRASTRUCTURE (CPU oracle) - never linked into or called from the product path.
//
// ov_plane::PlaneFitting restated on the CPU (reference: ov_plane/src/track_plane/PlaneFitting.cpp):
//   fit_plane       :44-81    linear plane fit  A n = -1  (condition check by singular values, column-pivoted Householder QR solve)
//   plane_fitting   :83-195   RANSAC over 5-point sets drawn with std::shuffle(std::mt19937(8888)), inlier test 0.05 m, final fit
//   optimize_plane  :197-514  joint refinement of the on-plane features (and the plane) with Ceres: reprojection factors
//                             (ov_init::Factor_ImageReprojCalib, OpenVINS @74a63cf, not in the tree: camera-frame poses, identity
//                             extrinsics, intrinsics [1 1 0 ...] => residual (pi(R (p_f - p_C)) - uv_norm) / sigma) and point-on-plane
//                             factors (ceres/Factor_PointOnPlane.cpp:39-70), Cauchy loss a = 1 on every block, DENSE_SCHUR, DOGLEG,
//                             12 iterations, failure unless the solver reports CONVERGENCE (:431)
//
// Ceres itself is a third-party dependency that is NOT under /root/reference (package.xml:47 `libceres-dev`: 1.14.0 on the Ubuntu
// 20.04 image of Dockerfile_ros1_20_04:20).  MiniCeres below restates the published algorithm of that version's
// TrustRegionMinimizer + DoglegStrategy(TRADITIONAL_DOGLEG) with the solver defaults (Jacobi scaling, function_tolerance 1e-6,
// gradient_tolerance 1e-10, parameter_tolerance 1e-8, initial radius 1e4, min_relative_decrease 1e-3, mu in [1e-8, 1] x10,
// min_lm_diagonal 1e-6) and the loss correction of corrector.cc (rho'' <= 0 for Cauchy: residual and Jacobian scaled by sqrt(rho')).
// PARITY UNPINNED: there is no Ceres here to run; tests/test_cpu_planefit.py pins the restatement by (i) the recovery of the true plane on
// noise-free data, (ii) an independent NumPy transcription of the objective from the factor definitions (equal to evaluate() to 1e-10) and
// generic minimisers of it (SciPy L-BFGS-B / trust-region least squares) that approach the end point from above and never undercut it,
// (iii) the CUDA path, which forms the same iteration through per-feature blocks and a Schur complement instead of the dense normal
// equations used here (identical iteration counts and termination reasons, results to 1e-14).
//
// std::shuffle / std::uniform_int_distribution are implementation-defined: libstdc++ changed uniform_int_distribution in GCC 11
// (Lemire's method for 32-bit generators).  The reference's documented toolchains (Ubuntu 18.04 / 20.04: GCC 7 / 9) use the older
// down-scaling loop.  Both are restated (ShuffleKind) so that the result does not depend on the compiler that builds THIS file;
// tests/test_cpu_planefit.py checks the LEMIRE restatement against this container's own std::shuffle.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>
#include <vector>

#include "oracle.hpp"

namespace orc {

enum ShuffleKind { SHUFFLE_LIBSTDCXX_CLASSIC = 0 /* GCC 7..10 */, SHUFFLE_LIBSTDCXX_LEMIRE = 1 /* GCC >= 11 */, SHUFFLE_STD = 2 /* this build's std::shuffle */ };

struct PlaneFitting {
  // ---- libstdc++ std::shuffle restated (bits/stl_algo.h) over std::mt19937 (result_type = uint_fast32_t, range 2^32 - 1) ----
  static uint64_t uniform_below(std::mt19937 &g, uint64_t n, ShuffleKind kind) { // uniform_int_distribution<unsigned long>{0, n - 1}
    if (n == 0)
      return 0;
    const uint64_t urngrange = 0xFFFFFFFFull, urange = n - 1;
    if (urngrange > urange) {
      if (kind == SHUFFLE_LIBSTDCXX_LEMIRE) {
        const uint32_t range = (uint32_t)n;
        uint64_t product = (uint64_t)g() * (uint64_t)range;
        uint32_t low = (uint32_t)product;
        if (low < range) {
          const uint32_t threshold = (uint32_t)(-range) % range;
          while (low < threshold) {
            product = (uint64_t)g() * (uint64_t)range;
            low = (uint32_t)product;
          }
        }
        return product >> 32;
      }
      const uint64_t uerange = urange + 1, scaling = urngrange / uerange, past = uerange * scaling;
      uint64_t ret;
      do
        ret = (uint64_t)g();
      while (ret >= past);
      return ret / scaling;
    }
    return (uint64_t)g(); // urngrange == urange (never for these sizes)
  }
  static void shuffle(std::vector<int> &v, std::mt19937 &g, ShuffleKind kind) {
    if (kind == SHUFFLE_STD) {
      std::shuffle(v.begin(), v.end(), g);
      return;
    }
    const size_t n = v.size();
    if (n == 0)
      return;
    const uint64_t urngrange = 0xFFFFFFFFull, urange = n;
    if (urngrange / urange >= urange) {
      size_t i = 1;
      if ((urange % 2) == 0) {
        std::swap(v[i], v[uniform_below(g, 2, kind)]);
        i++;
      }
      while (i != n) {
        const uint64_t swap_range = (uint64_t)i + 1;
        const uint64_t x = uniform_below(g, swap_range * (swap_range + 1), kind);
        std::swap(v[i], v[x / (swap_range + 1)]);
        i++;
        std::swap(v[i], v[x % (swap_range + 1)]);
        i++;
      }
      return;
    }
    for (size_t i = 1; i < n; i++)
      std::swap(v[i], v[uniform_below(g, (uint64_t)i + 1, kind)]);
  }

  // ---- small dense helpers with Eigen-like algorithms -------------------------------------------------------------------------
  // singular values of a K x 3 matrix by one-sided (Hestenes) Jacobi, like Eigen::JacobiSVD's two-sided sweeps in spirit
  static void singular_values3(Mat A, double s[3]) {
    const int K = A.r;
    for (int sweep = 0; sweep < 60; sweep++) {
      double off = 0;
      for (int p = 0; p < 3; p++)
        for (int q = p + 1; q < 3; q++) {
          double a = 0, b = 0, c = 0;
          for (int i = 0; i < K; i++) {
            a += A(i, p) * A(i, p);
            b += A(i, q) * A(i, q);
            c += A(i, p) * A(i, q);
          }
          if (std::abs(c) <= 1e-300 || std::abs(c) <= 1e-17 * std::sqrt(a * b))
            continue;
          off = std::max(off, std::abs(c) / std::sqrt(a * b));
          const double zeta = (b - a) / (2 * c);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::abs(zeta) + std::sqrt(1 + zeta * zeta));
          const double cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
          for (int i = 0; i < K; i++) {
            const double x = A(i, p), y = A(i, q);
            A(i, p) = cs * x - sn * y;
            A(i, q) = sn * x + cs * y;
          }
        }
      if (off < 1e-15)
        break;
    }
    for (int j = 0; j < 3; j++) {
      double n = 0;
      for (int i = 0; i < K; i++)
        n += A(i, j) * A(i, j);
      s[j] = std::sqrt(n);
    }
    std::sort(s, s + 3, [](double x, double y) { return x > y; });
  }
  // least squares A x = b, A K x 3, by Householder QR with column pivoting (Eigen::ColPivHouseholderQR::solve)
  static void lstsq3_colpiv(Mat A, Mat b, double x[3]) {
    const int K = A.r;
    int perm[3] = {0, 1, 2};
    for (int j = 0; j < 3 && j < K; j++) {
      int best = j;
      double bn = -1;
      for (int c = j; c < 3; c++) {
        double n = 0;
        for (int i = j; i < K; i++)
          n += A(i, c) * A(i, c);
        if (n > bn) {
          bn = n;
          best = c;
        }
      }
      if (best != j) {
        for (int i = 0; i < K; i++)
          std::swap(A(i, j), A(i, best));
        std::swap(perm[j], perm[best]);
      }
      double sigma = 0;
      for (int i = j; i < K; i++)
        sigma += A(i, j) * A(i, j);
      sigma = std::sqrt(sigma);
      if (sigma == 0.0)
        continue;
      const double alpha = (A(j, j) > 0) ? -sigma : sigma;
      std::vector<double> v(K, 0.0);
      for (int i = j; i < K; i++)
        v[i] = A(i, j);
      v[j] -= alpha;
      double vn = 0;
      for (int i = j; i < K; i++)
        vn += v[i] * v[i];
      if (vn == 0.0)
        continue;
      for (int c = j; c < 3; c++) {
        double d = 0;
        for (int i = j; i < K; i++)
          d += v[i] * A(i, c);
        d = 2 * d / vn;
        for (int i = j; i < K; i++)
          A(i, c) -= d * v[i];
      }
      double d = 0;
      for (int i = j; i < K; i++)
        d += v[i] * b(i, 0);
      d = 2 * d / vn;
      for (int i = j; i < K; i++)
        b(i, 0) -= d * v[i];
    }
    double y[3] = {0, 0, 0};
    for (int j = 2; j >= 0; j--) {
      double s = b(j, 0);
      for (int c = j + 1; c < 3; c++)
        s -= A(j, c) * y[c];
      y[j] = s / A(j, j);
    }
    for (int j = 0; j < 3; j++)
      x[perm[j]] = y[j];
  }
  static double point_to_plane_distance(const double *p, const double *abcd) { // PlaneFitting.h:73-75
    return p[0] * abcd[0] + p[1] * abcd[1] + p[2] * abcd[2] + abcd[3];
  }

  // PlaneFitting::fit_plane (:44-81).  pts: 3 doubles per point; idx: the subset.
  static bool fit_plane(const double *pts, const std::vector<int> &idx, double abcd[4], double cond_thresh, bool cond_check = true) {
    const int K = (int)idx.size();
    if (K < 3)
      return false;
    Mat A(K, 3), b(K, 1);
    for (int i = 0; i < K; i++) {
      for (int j = 0; j < 3; j++)
        A(i, j) = pts[3 * idx[i] + j];
      b(i, 0) = -1.0;
    }
    if (cond_check) {
      double s[3];
      singular_values3(A, s);
      const double cond = s[0] / s[2];
      if (cond > cond_thresh)
        return false;
    }
    double n[3];
    lstsq3_colpiv(A, b, n);
    const double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int j = 0; j < 3; j++)
      abcd[j] = n[j] / nn;
    abcd[3] = 1.0 / nn;
    const double cpn = std::abs(abcd[3]) * std::sqrt(abcd[0] * abcd[0] + abcd[1] * abcd[1] + abcd[2] * abcd[2]);
    return cpn > 0.02;
  }

  // PlaneFitting::plane_fitting (:83-195).  inlier: F flags (the reference replaces `feats` by the inlier set, original order kept).
  static bool plane_fitting(int F, const double *pts, double abcd[4], int min_inlier_num, double max_cond, std::vector<int> &inlier,
                            ShuffleKind kind = SHUFFLE_LIBSTDCXX_CLASSIC) {
    const int ransac_solver_feat_num = 5;
    const int max_iter_num = 200;
    const double min_inlier_ratio = 0.80, max_error_threshold = 0.05, min_distance_between_points = 0.05;
    const size_t min_feat_on_plane_num_threshold = (size_t)std::max(min_inlier_num, (int)((double)F * min_inlier_ratio));
    std::mt19937 rand_gen(8888);
    inlier.assign(F, 0);
    if (F < min_inlier_num)
      return false;
    double best_error = -1;
    std::vector<int> best_inliers;
    for (int n = 0; n < max_iter_num; n++) {
      std::vector<int> copy(F);
      for (int i = 0; i < F; i++)
        copy[i] = i;
      shuffle(copy, rand_gen, kind);
      std::vector<int> ransac_set;
      size_t it = 0;
      while ((int)ransac_set.size() < ransac_solver_feat_num && it != copy.size()) {
        if (ransac_set.empty()) {
          ransac_set.push_back(copy[it]);
        } else {
          bool good = true;
          const double *p = pts + 3 * copy[it];
          for (int q : ransac_set) {
            const double dx = pts[3 * q] - p[0], dy = pts[3 * q + 1] - p[1], dz = pts[3 * q + 2] - p[2];
            if (std::sqrt(dx * dx + dy * dy + dz * dz) < min_distance_between_points) {
              good = false;
              break;
            }
          }
          if (good)
            ransac_set.push_back(copy[it]);
        }
        it++;
      }
      if ((int)ransac_set.size() != ransac_solver_feat_num)
        return false;
      if (fit_plane(pts, ransac_set, abcd, max_cond)) {
        double avg = 0.0;
        std::vector<int> inl;
        for (int f = 0; f < F; f++) {
          const double e = point_to_plane_distance(pts + 3 * f, abcd);
          if (std::abs(e) < max_error_threshold) {
            inl.push_back(f);
            avg += std::abs(e);
          }
        }
        avg /= (double)inl.size();
        const bool valid_set = (inl.size() > min_feat_on_plane_num_threshold && avg < max_error_threshold);
        const bool better_set = ((best_inliers.size() < inl.size()) || (best_inliers.size() == inl.size() && avg < best_error));
        if (valid_set && better_set) {
          best_inliers = inl;
          best_error = avg;
        }
      }
    }
    if (!best_inliers.empty()) {
      if (fit_plane(pts, best_inliers, abcd, max_cond, false)) {
        for (int f : best_inliers)
          inlier[f] = 1;
        return true;
      }
    }
    return false;
  }

  // ---- optimize_plane ----------------------------------------------------------------------------------------------------------
  struct Problem {
    int F = 0;
    std::vector<int> meas_offset;                       // F + 1
    std::vector<FeatureInitializer::ClonePose> cam;     // per measurement: R_GtoCi, p_CiinG (clonesCAM)
    std::vector<double> uvn;                            // per measurement: uv_norm cast to double
    std::vector<double> p0;                             // 3F initial feature positions
    double cp0[3];
    double sigma_px_norm, sigma_c;
    bool fix_plane;
    // derived: free-parameter layout
    std::vector<int> feat_col; // F: column offset of the feature block, -1 when constant (no measurements)
    int cp_col = -1, n = 0;
  };
  static void layout(Problem &P) {
    P.feat_col.assign(P.F, -1);
    P.n = 0;
    for (int f = 0; f < P.F; f++)
      if (P.meas_offset[f + 1] > P.meas_offset[f]) {
        P.feat_col[f] = P.n;
        P.n += 3;
      }
    P.cp_col = -1;
    if (!P.fix_plane) {
      P.cp_col = P.n;
      P.n += 3;
    }
  }
  // residuals (loss-corrected), Jacobian (loss-corrected, unscaled), cost = 1/2 sum rho(s).  x = free parameters.
  static double evaluate(const Problem &P, const std::vector<double> &x, std::vector<double> *res, Mat *J) {
    const double slam_inflation = 2.0;
    double cp[3];
    for (int i = 0; i < 3; i++)
      cp[i] = (P.cp_col >= 0) ? x[P.cp_col + i] : P.cp0[i];
    const double d = std::sqrt(cp[0] * cp[0] + cp[1] * cp[1] + cp[2] * cp[2]);
    const double nrm[3] = {cp[0] / d, cp[1] / d, cp[2] / d};
    int rows = 0;
    for (int f = 0; f < P.F; f++) {
      const int m = P.meas_offset[f + 1] - P.meas_offset[f];
      if (m == 0)
        rows += (P.cp_col >= 0) ? 1 : 0; // constant feature + constant plane: the block has no free parameter (Ceres drops it)
      else
        rows += 3 * m;
    }
    if (res)
      res->assign(rows, 0.0);
    if (J)
      *J = Mat(rows, P.n);
    double cost = 0;
    int row = 0;
    auto plane_block = [&](int f, const double *p, double sigma) {
      const double wc = 1.0 / sigma;
      const double ndp = nrm[0] * p[0] + nrm[1] * p[1] + nrm[2] * p[2];
      const double r = -1.0 * wc * (0.0 - (ndp - d)); // Factor_PointOnPlane.cpp:52
      const double s = r * r, rho1 = std::max(std::numeric_limits<double>::min(), 1.0 / (1.0 + s)), sc = std::sqrt(rho1);
      cost += 0.5 * std::log(1.0 + s);
      if (res)
        (*res)[row] = sc * r;
      if (J) {
        if (P.feat_col[f] >= 0)
          for (int i = 0; i < 3; i++)
            (*J)(row, P.feat_col[f] + i) = sc * wc * nrm[i]; // :58-61
        if (P.cp_col >= 0)
          for (int i = 0; i < 3; i++)
            (*J)(row, P.cp_col + i) = sc * wc * 1.0 / d * (p[i] - ndp * nrm[i] - d * nrm[i]); // :64-68
      }
      row++;
    };
    for (int f = 0; f < P.F; f++) {
      const int m0 = P.meas_offset[f], m = P.meas_offset[f + 1] - m0;
      double p[3];
      for (int i = 0; i < 3; i++)
        p[i] = (P.feat_col[f] >= 0) ? x[P.feat_col[f] + i] : P.p0[3 * f + i];
      if (m == 0) { // SLAM feature: constant, one inflated constraint (PlaneFitting.cpp:274-277)
        if (P.cp_col >= 0)
          plane_block(f, p, slam_inflation * P.sigma_c);
        continue;
      }
      for (int k = 0; k < m; k++) {
        const FeatureInitializer::ClonePose &c = P.cam[m0 + k];
        double pc[3];
        for (int i = 0; i < 3; i++)
          pc[i] = c.R(i, 0) * (p[0] - c.p(0, 0)) + c.R(i, 1) * (p[1] - c.p(1, 0)) + c.R(i, 2) * (p[2] - c.p(2, 0));
        const double w = 1.0 / P.sigma_px_norm;
        const double r0 = w * (pc[0] / pc[2] - P.uvn[2 * (m0 + k)]), r1 = w * (pc[1] / pc[2] - P.uvn[2 * (m0 + k) + 1]);
        const double s = r0 * r0 + r1 * r1, rho1 = std::max(std::numeric_limits<double>::min(), 1.0 / (1.0 + s)), sc = std::sqrt(rho1);
        cost += 0.5 * std::log(1.0 + s);
        if (res) {
          (*res)[row] = sc * r0;
          (*res)[row + 1] = sc * r1;
        }
        if (J) {
          const double dz[2][3] = {{1.0 / pc[2], 0.0, -pc[0] / (pc[2] * pc[2])}, {0.0, 1.0 / pc[2], -pc[1] / (pc[2] * pc[2])}};
          for (int a = 0; a < 2; a++)
            for (int i = 0; i < 3; i++)
              (*J)(row + a, P.feat_col[f] + i) = sc * w * (dz[a][0] * c.R(0, i) + dz[a][1] * c.R(1, i) + dz[a][2] * c.R(2, i));
        }
        row += 2;
        plane_block(f, p, P.sigma_c); // one constraint per measurement (:367-369)
      }
    }
    return cost;
  }

  struct Summary {
    bool converged = false;
    int iterations = 0; // summary.iterations.size() - 1
    double initial_cost = 0, final_cost = 0;
    int reason = 0; // 1 gradient, 2 parameter, 3 function tolerance, 4 no free parameters, -1 max iterations, -2 invalid steps / failure
  };
  // Ceres 1.14 TrustRegionMinimizer + DoglegStrategy (TRADITIONAL_DOGLEG), dense normal equations for the Gauss-Newton solve
  static Summary mini_ceres_dogleg(const Problem &P, std::vector<double> &x, int max_num_iterations = 12) {
    Summary S;
    const int n = P.n;
    if (n == 0) { // "No non-constant parameter blocks found": Ceres reports CONVERGENCE
      S.converged = true;
      S.reason = 4;
      return S;
    }
    const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8, min_relative_decrease = 1e-3;
    const double min_diagonal = 1e-6, max_diagonal = 1e32, min_mu = 1e-8, max_mu = 1.0, mu_increase_factor = 10.0;
    double radius = 1e4, mu = min_mu, alpha = 0, dogleg_step_norm = 0;
    bool reuse = false;
    std::vector<double> res, scale(n), diagonal(n), gradient(n), gn(n), step(n), delta(n), cand(n);
    Mat J;
    double x_cost = evaluate(P, x, &res, &J);
    S.initial_cost = S.final_cost = x_cost;
    for (int j = 0; j < n; j++) { // Jacobi scaling from the first Jacobian
      double s = 0;
      for (int i = 0; i < J.r; i++)
        s += J(i, j) * J(i, j);
      scale[j] = 1.0 / (1.0 + std::sqrt(s));
    }
    auto scale_columns = [&]() {
      for (int j = 0; j < n; j++)
        for (int i = 0; i < J.r; i++)
          J(i, j) *= scale[j];
    };
    auto gradient_max_norm = [&]() { // of the UNSCALED Jacobian: call before scale_columns
      double g = 0;
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int i = 0; i < J.r; i++)
          s += J(i, j) * res[i];
        g = std::max(g, std::abs(s));
      }
      return g;
    };
    double gmax = gradient_max_norm();
    scale_columns();
    double x_norm = 0;
    for (double v : x)
      x_norm += v * v;
    x_norm = std::sqrt(x_norm);
    bool last_successful = false;
    int num_consecutive_invalid = 0;
    int iteration = 0;
    (void)gmax;
    while (true) {
      // FinalizeIterationAndCheckIfMinimizerCanContinue
      if (iteration >= max_num_iterations) {
        S.reason = -1;
        break;
      }
      if (last_successful && gmax <= gradient_tolerance) {
        S.converged = true;
        S.reason = 1;
        break;
      }
      iteration++;
      last_successful = false;
      // ---- DoglegStrategy::ComputeStep ----
      bool solve_ok = true;
      auto traditional_dogleg = [&]() {
        double gnorm = 0, gnn = 0;
        for (int j = 0; j < n; j++) {
          gnorm += gradient[j] * gradient[j];
          gnn += gn[j] * gn[j];
        }
        gnorm = std::sqrt(gnorm);
        gnn = std::sqrt(gnn);
        if (gnn <= radius) {
          for (int j = 0; j < n; j++)
            step[j] = gn[j];
          dogleg_step_norm = gnn;
        } else if (gnorm * alpha >= radius) {
          for (int j = 0; j < n; j++)
            step[j] = -(radius / gnorm) * gradient[j];
          dogleg_step_norm = radius;
        } else {
          double gdot = 0;
          for (int j = 0; j < n; j++)
            gdot += gradient[j] * gn[j];
          const double b_dot_a = -alpha * gdot;
          const double a_sq = std::pow(alpha * gnorm, 2.0);
          const double bma_sq = a_sq - 2 * b_dot_a + std::pow(gnn, 2);
          const double c = b_dot_a - a_sq;
          const double dd = std::sqrt(c * c + bma_sq * (std::pow(radius, 2.0) - a_sq));
          const double beta = (c <= 0) ? (dd - c) / bma_sq : (radius * radius - a_sq) / (dd + c);
          double sn = 0;
          for (int j = 0; j < n; j++) {
            step[j] = (-alpha * (1.0 - beta)) * gradient[j] + beta * gn[j];
            sn += step[j] * step[j];
          }
          dogleg_step_norm = std::sqrt(sn);
        }
        for (int j = 0; j < n; j++)
          step[j] /= diagonal[j];
      };
      if (reuse) {
        traditional_dogleg();
      } else {
        reuse = true;
        for (int j = 0; j < n; j++) {
          double s = 0;
          for (int i = 0; i < J.r; i++)
            s += J(i, j) * J(i, j);
          diagonal[j] = std::sqrt(std::min(std::max(s, min_diagonal), max_diagonal));
        }
        for (int j = 0; j < n; j++) { // gradient = D^-1 J^T r
          double s = 0;
          for (int i = 0; i < J.r; i++)
            s += J(i, j) * res[i];
          gradient[j] = s / diagonal[j];
        }
        { // Cauchy point
          double g2 = 0, jg2 = 0;
          for (int j = 0; j < n; j++)
            g2 += gradient[j] * gradient[j];
          for (int i = 0; i < J.r; i++) {
            double s = 0;
            for (int j = 0; j < n; j++)
              s += J(i, j) * (gradient[j] / diagonal[j]);
            jg2 += s * s;
          }
          alpha = g2 / jg2;
        }
        // Gauss-Newton step: (J^T J + mu D^2) y = J^T r, gn = -D y
        solve_ok = false;
        Mat JtJ = J.T() * J, Jtr(n, 1);
        for (int j = 0; j < n; j++) {
          double s = 0;
          for (int i = 0; i < J.r; i++)
            s += J(i, j) * res[i];
          Jtr(j, 0) = s;
        }
        while (mu < max_mu) {
          Mat A = JtJ, L;
          for (int j = 0; j < n; j++)
            A(j, j) += mu * diagonal[j] * diagonal[j];
          bool ok = chol_lower(A, L);
          Mat y = Jtr;
          if (ok) {
            chol_solve_inplace(L, y);
            for (int j = 0; j < n; j++)
              if (!std::isfinite(y(j, 0)))
                ok = false;
          }
          if (!ok) {
            mu *= mu_increase_factor;
            continue;
          }
          for (int j = 0; j < n; j++)
            gn[j] = -diagonal[j] * y(j, 0);
          solve_ok = true;
          break;
        }
        if (solve_ok)
          traditional_dogleg();
      }
      // ---- TrustRegionMinimizer::ComputeTrustRegionStep ----
      bool step_is_valid = false;
      double model_cost_change = 0;
      if (solve_ok) {
        for (int i = 0; i < J.r; i++) {
          double mr = 0;
          for (int j = 0; j < n; j++)
            mr += J(i, j) * step[j];
          model_cost_change += -mr * (res[i] + mr / 2.0);
        }
        step_is_valid = (model_cost_change > 0.0);
      }
      if (!step_is_valid) { // HandleInvalidStep
        if (++num_consecutive_invalid >= 5) {
          S.reason = -2;
          break;
        }
        mu *= mu_increase_factor; // StepIsInvalid
        reuse = false;
        continue;
      }
      num_consecutive_invalid = 0;
      double step_norm = 0;
      for (int j = 0; j < n; j++) {
        delta[j] = step[j] * scale[j];
        cand[j] = x[j] + delta[j];
        step_norm += delta[j] * delta[j];
      }
      step_norm = std::sqrt(step_norm);
      double cand_cost = evaluate(P, cand, nullptr, nullptr);
      if (!std::isfinite(cand_cost))
        cand_cost = std::numeric_limits<double>::max();
      if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) { // ParameterToleranceReached
        S.converged = true;
        S.reason = 2;
        break;
      }
      if (std::abs(x_cost - cand_cost) <= function_tolerance * x_cost) { // FunctionToleranceReached (the candidate is not taken)
        S.converged = true;
        S.reason = 3;
        break;
      }
      const double relative_decrease = (x_cost - cand_cost) / model_cost_change;
      if (getenv("ORC_PF_TRACE")) // per-iteration trace of the checker (tools / debugging)
        fprintf(stderr, "it %d cost %.6e cand %.6e model %.3e rel %.3f radius %.3e step %.3e dogleg %.3e mu %.1e\n", iteration, x_cost, cand_cost, model_cost_change,
                relative_decrease, radius, step_norm, dogleg_step_norm, mu);
      if (relative_decrease > min_relative_decrease) { // HandleSuccessfulStep
        x = cand;
        x_norm = 0;
        for (double v : x)
          x_norm += v * v;
        x_norm = std::sqrt(x_norm);
        x_cost = evaluate(P, x, &res, &J);
        gmax = gradient_max_norm();
        scale_columns();
        last_successful = true;
        if (relative_decrease < 0.25)
          radius *= 0.5;
        if (relative_decrease > 0.75)
          radius = std::max(radius, 3.0 * dogleg_step_norm);
        mu = std::max(min_mu, 2.0 * mu / mu_increase_factor);
        reuse = false;
      } else { // HandleUnsuccessfulStep
        radius *= 0.5;
        reuse = true;
      }
    }
    S.iterations = iteration;
    S.final_cost = x_cost;
    return S;
  }

  // PlaneFitting::optimize_plane (:197-514).  Outputs: p (3F, updated for the inliers like the reference's side effect), cp (updated
  // whenever the solver converged), inlier flags.  stateI7 = [q_GtoI, p_IinG] of the CURRENT IMU pose, calib7 = [q_ItoC, p_IinC].
  static bool optimize_plane(Problem P, const Mat &R_GtoI, const Mat &p_IinG, const Mat &R_ItoC, const Mat &p_IinC, std::vector<double> &p,
                             double cp[3], std::vector<int> &inlier, Summary *sum_out = nullptr, int max_num_iterations = 12) {
    const double min_inlier_ratio = 0.80, max_error_threshold = 0.03;
    const int F = P.F;
    const size_t min_feat_on_plane_num_threshold = (size_t)std::max(4, (int)((double)F * min_inlier_ratio));
    p = P.p0;
    for (int i = 0; i < 3; i++)
      cp[i] = P.cp0[i];
    inlier.assign(F, 0);
    if ((!P.fix_plane && F < 4) || (P.fix_plane && F == 0))
      return false;
    layout(P);
    std::vector<double> x(P.n, 0.0);
    for (int f = 0; f < F; f++)
      if (P.feat_col[f] >= 0)
        for (int i = 0; i < 3; i++)
          x[P.feat_col[f] + i] = P.p0[3 * f + i];
    if (P.cp_col >= 0)
      for (int i = 0; i < 3; i++)
        x[P.cp_col + i] = P.cp0[i];
    Summary S = mini_ceres_dogleg(P, x, max_num_iterations);
    if (sum_out)
      *sum_out = S;
    if (!S.converged)
      return false;
    if (P.cp_col >= 0)
      for (int i = 0; i < 3; i++)
        cp[i] = x[P.cp_col + i];
    const double cn = std::sqrt(cp[0] * cp[0] + cp[1] * cp[1] + cp[2] * cp[2]);
    const double abcd[4] = {cp[0] / cn, cp[1] / cn, cp[2] / cn, -cn};
    Mat R_GtoCi = R_ItoC * R_GtoI;
    Mat p_CiinG = p_IinG - R_GtoCi.T() * p_IinC;
    size_t n_inl = 0;
    for (int f = 0; f < F; f++) {
      double after[3];
      for (int i = 0; i < 3; i++)
        after[i] = (P.feat_col[f] >= 0) ? x[P.feat_col[f] + i] : P.p0[3 * f + i];
      const double error = point_to_plane_distance(&P.p0[3 * f], abcd); // the position BEFORE the refinement (:459)
      if (std::abs(error) >= max_error_threshold)
        continue;
      if (std::isnan(std::sqrt(after[0] * after[0] + after[1] * after[1] + after[2] * after[2])))
        continue;
      const double z = R_GtoCi(2, 0) * (after[0] - p_CiinG(0, 0)) + R_GtoCi(2, 1) * (after[1] - p_CiinG(1, 0)) + R_GtoCi(2, 2) * (after[2] - p_CiinG(2, 0));
      if (z < 0.1)
        continue;
      for (int i = 0; i < 3; i++)
        p[3 * f + i] = after[i];
      inlier[f] = 1;
      n_inl++;
    }
    if ((F != 1 && n_inl < min_feat_on_plane_num_threshold) || (P.fix_plane && F == 1 && n_inl == 0))
      return false;
    return true;
  }
};

} // namespace orc
