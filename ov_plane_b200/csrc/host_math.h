// Host-side 3x3 / 4x4 / 15x15 arithmetic of the Propagator (state/Propagator.cpp:343-569): IMU mean integration (RK4 /
// discrete) and the per-interval F, Qd.  These are O(15^3) per IMU interval and strictly sequential (each interval's FEJ
// linearisation point is the previous interval's result, Propagator.cpp:448-453), so they stay on the host; the N-sized
// work (EKFPropagation + augment_clone) runs on the device.  Row-major small matrices.
#pragma once
#include <cmath>
#include <cstring>

namespace ovp {
namespace hm {

// 0.95 quantile of the chi-squared distribution with `dof` degrees of freedom: what boost::math::quantile(chi_squared(dof), 0.95)
// returns in the reference (UpdaterMSCKF.cpp:59-62, on the fly beyond its 500-entry table, :616-619).  Newton iteration on the
// regularised lower incomplete gamma function P(dof/2, x/2) (series below a+1, Lentz continued fraction above), started from the
// Wilson-Hilferty approximation; relative accuracy ~1e-13.
inline double gamma_p(double a, double x) {
  if (x <= 0.0)
    return 0.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 10000; n++) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17)
        break;
    }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 10000; i++) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < tiny)
      d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny)
      c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-17)
      break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}
inline double chi2_quantile95(int dof) {
  const double k = dof, a = 0.5 * k, z = 1.6448536269514722;
  const double t = 1.0 - 2.0 / (9.0 * k) + z * std::sqrt(2.0 / (9.0 * k));
  double x = std::fmax(1e-3, k * t * t * t);
  for (int it = 0; it < 100; it++) {
    const double f = gamma_p(a, 0.5 * x) - 0.95;
    const double pdf = 0.5 * std::exp(-0.5 * x + (a - 1.0) * std::log(0.5 * x) - std::lgamma(a));
    const double dx = f / pdf;
    x -= dx;
    if (x <= 0.0)
      x = 1e-6;
    if (std::fabs(dx) < 1e-14 * x)
      break;
  }
  return x;
}

struct M3 {
  double a[9];
  double &operator()(int i, int j) { return a[3 * i + j]; }
  double operator()(int i, int j) const { return a[3 * i + j]; }
};
struct V3 {
  double v[3];
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
struct V4 {
  double v[4];
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};

inline V3 v3(double x, double y, double z) { return V3{{x, y, z}}; }
inline V3 operator+(const V3 &a, const V3 &b) { return v3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3 operator-(const V3 &a, const V3 &b) { return v3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3 operator*(double s, const V3 &a) { return v3(s * a[0], s * a[1], s * a[2]); }
inline double norm(const V3 &a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
inline M3 eye3() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline M3 operator*(const M3 &A, const M3 &B) {
  M3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
  return C;
}
inline V3 operator*(const M3 &A, const V3 &x) {
  return v3(A(0, 0) * x[0] + A(0, 1) * x[1] + A(0, 2) * x[2], A(1, 0) * x[0] + A(1, 1) * x[1] + A(1, 2) * x[2],
            A(2, 0) * x[0] + A(2, 1) * x[1] + A(2, 2) * x[2]);
}
inline M3 operator*(double s, const M3 &A) {
  M3 C;
  for (int i = 0; i < 9; i++)
    C.a[i] = s * A.a[i];
  return C;
}
inline M3 operator+(const M3 &A, const M3 &B) {
  M3 C;
  for (int i = 0; i < 9; i++)
    C.a[i] = A.a[i] + B.a[i];
  return C;
}
inline M3 transpose(const M3 &A) {
  M3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C(i, j) = A(j, i);
  return C;
}
inline M3 skew(const V3 &w) { return M3{{0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}}; }
inline M3 outer(const V3 &a, const V3 &b) {
  M3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C(i, j) = a[i] * b[j];
  return C;
}

// ov_core quat_ops.h (JPL)
inline M3 quat_2_Rot(const V4 &q) {
  V3 v = v3(q[0], q[1], q[2]);
  double w = q[3];
  return (2 * w * w - 1) * eye3() + (-2 * w) * skew(v) + 2.0 * outer(v, v);
}
// rotation matrix -> JPL quaternion (ov_core quat_ops.h rot_2_quat: the largest of {R00, R11, R22, trace} picks the branch)
inline V4 rot_2_quat(const M3 &R) {
  V4 q;
  const double T = R(0, 0) + R(1, 1) + R(2, 2);
  int br = 3;
  if (R(0, 0) >= T && R(0, 0) >= R(1, 1) && R(0, 0) >= R(2, 2))
    br = 0;
  else if (R(1, 1) >= T && R(1, 1) >= R(0, 0) && R(1, 1) >= R(2, 2))
    br = 1;
  else if (R(2, 2) >= T && R(2, 2) >= R(0, 0) && R(2, 2) >= R(1, 1))
    br = 2;
  if (br == 3) {
    q[3] = std::sqrt((1 + T) / 4);
    const double s = 1 / (4 * q[3]);
    q[0] = s * (R(1, 2) - R(2, 1));
    q[1] = s * (R(2, 0) - R(0, 2));
    q[2] = s * (R(0, 1) - R(1, 0));
  } else {
    const int i = br, j = (br + 1) % 3, k = (br + 2) % 3;
    q[i] = std::sqrt((1 + 2 * R(i, i) - T) / 4);
    const double s = 1 / (4 * q[i]);
    q[j] = s * (R(i, j) + R(j, i));
    q[k] = s * (R(i, k) + R(k, i));
    q[3] = s * (R(j, k) - R(k, j));
  }
  if (q[3] < 0)
    for (int a = 0; a < 4; a++)
      q[a] = -q[a];
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int a = 0; a < 4; a++)
    q[a] /= n;
  return q;
}
inline V4 quatnorm(V4 q) {
  if (q[3] < 0)
    for (int i = 0; i < 4; i++)
      q[i] = -q[i];
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++)
    q[i] /= n;
  return q;
}
inline V4 quat_multiply(const V4 &q, const V4 &p) {
  // Qm = [q3 I - skew(qv), qv; -qv^T, q3]
  V4 r;
  r[0] = q[3] * p[0] + q[2] * p[1] - q[1] * p[2] + q[0] * p[3];
  r[1] = -q[2] * p[0] + q[3] * p[1] + q[0] * p[2] + q[1] * p[3];
  r[2] = q[1] * p[0] - q[0] * p[1] + q[3] * p[2] + q[2] * p[3];
  r[3] = -q[0] * p[0] - q[1] * p[1] - q[2] * p[2] + q[3] * p[3];
  if (r[3] < 0)
    for (int i = 0; i < 4; i++)
      r[i] = -r[i];
  double n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
  for (int i = 0; i < 4; i++)
    r[i] /= n;
  return r;
}
// 0.5 * Omega(w) * q
inline V4 half_omega_times(const V3 &w, const V4 &q) {
  // Omega = [-skew(w), w; -w^T, 0]
  V4 r;
  r[0] = 0.5 * (w[2] * q[1] - w[1] * q[2] + w[0] * q[3]);
  r[1] = 0.5 * (-w[2] * q[0] + w[0] * q[2] + w[1] * q[3]);
  r[2] = 0.5 * (w[1] * q[0] - w[0] * q[1] + w[2] * q[3]);
  r[3] = 0.5 * (-w[0] * q[0] - w[1] * q[1] - w[2] * q[2]);
  return r;
}
inline M3 exp_so3(const V3 &w) {
  M3 wx = skew(w);
  double theta = norm(w);
  double A, B;
  if (theta < 1e-7) {
    A = 1;
    B = 0.5;
  } else {
    A = std::sin(theta) / theta;
    B = (1 - std::cos(theta)) / (theta * theta);
  }
  if (theta == 0)
    return eye3();
  return eye3() + A * wx + B * (wx * wx);
}
inline M3 Jl_so3(const V3 &w) {
  double theta = norm(w);
  if (theta < 1e-6)
    return eye3();
  V3 a = (1.0 / theta) * w;
  return (std::sin(theta) / theta) * eye3() + (1 - std::sin(theta) / theta) * outer(a, a) + ((1 - std::cos(theta)) / theta) * skew(a);
}
inline M3 Jr_so3(const V3 &w) { return Jl_so3(-1.0 * w); }

// 15x15 row-major helpers
struct M15 {
  double a[225];
  M15() { std::memset(a, 0, sizeof(a)); }
  double &operator()(int i, int j) { return a[15 * i + j]; }
  double operator()(int i, int j) const { return a[15 * i + j]; }
  void setBlock3(int i0, int j0, const M3 &B) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        (*this)(i0 + i, j0 + j) = B(i, j);
  }
};
inline M15 mul(const M15 &A, const M15 &B) {
  M15 C;
  for (int i = 0; i < 15; i++)
    for (int k = 0; k < 15; k++) {
      double aik = A(i, k);
      if (aik == 0.0)
        continue;
      for (int j = 0; j < 15; j++)
        C(i, j) += aik * B(k, j);
    }
  return C;
}
inline M15 mulT(const M15 &A, const M15 &B) { // A * B^T
  M15 C;
  for (int i = 0; i < 15; i++)
    for (int j = 0; j < 15; j++) {
      double s = 0;
      for (int k = 0; k < 15; k++)
        s += A(i, k) * B(j, k);
      C(i, j) = s;
    }
  return C;
}

} // namespace hm
} // namespace ovp
