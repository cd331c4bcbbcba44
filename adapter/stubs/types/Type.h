// MINIMAL STAND-IN for OpenVINS ov_core types/{Type,Vec,JPLQuat,PoseJPL,IMU,Landmark}.h (syntax check of adapter/*.cpp only).
#pragma once
#include <Eigen/Dense>
#include <memory>
namespace ov_type {
class Type {
public:
  explicit Type(int size) : _size(size) {}
  virtual ~Type() {}
  virtual void set_local_id(int new_id) { _id = new_id; }
  int id() { return _id; }
  int size() { return _size; }
  virtual const Eigen::MatrixXd &value() const { return _value; }
  virtual const Eigen::MatrixXd &fej() const { return _fej; }
  virtual void set_value(const Eigen::MatrixXd &new_value) { _value = new_value; }
  virtual void set_fej(const Eigen::MatrixXd &new_value) { _fej = new_value; }

protected:
  Eigen::MatrixXd _value, _fej;
  int _id = -1, _size = -1;
};
class Vec : public Type {
public:
  explicit Vec(int dim) : Type(dim) {}
};
class JPLQuat : public Type {
public:
  JPLQuat() : Type(3) {}
};
class PoseJPL : public Type {
public:
  PoseJPL() : Type(6) {}
};
class IMU : public Type {
public:
  IMU() : Type(15) {}
  std::shared_ptr<PoseJPL> pose() { return _pose; }

private:
  std::shared_ptr<PoseJPL> _pose;
};
class Landmark : public Vec {
public:
  explicit Landmark(int dim) : Vec(dim) {}
  size_t _featid = 0;
  bool should_marg = false;
};
} // namespace ov_type
