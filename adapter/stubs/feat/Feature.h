// MINIMAL STAND-IN for ov_core feat/Feature.h (OpenVINS @74a63cf): the members PlaneFitting reads.  Syntax check only.
#pragma once
#include <Eigen/Dense>
#include <unordered_map>
#include <vector>
namespace ov_core {
class Feature {
public:
  size_t featid;
  bool to_delete = false;
  std::unordered_map<size_t, std::vector<Eigen::Vector2f>> uvs;      // raw pixels (the reference holds Eigen::VectorXf of size 2)
  std::unordered_map<size_t, std::vector<Eigen::Vector2f>> uvs_norm; // undistorted normalised coordinates
  void clean_old_measurements(const std::vector<double> &valid_times);
  std::unordered_map<size_t, std::vector<double>> timestamps;
  Eigen::Vector3d p_FinG;
};
} // namespace ov_core
