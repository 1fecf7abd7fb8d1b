// Work-alike of the `state` carrier of reference faster/include/faster_types.hpp:79-165 (same member names and
// setters), for building solverGurobi.hpp without the reference tree.  When dropping solverGurobi.hpp into the
// reference, keep the reference's own faster_types.hpp: this file is then not used.
#pragma once
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#else
#include "fq_compat.hpp"
#endif
#include <iostream>

struct state
{
  Eigen::Vector3d pos = Eigen::Vector3d::Zero();
  Eigen::Vector3d vel = Eigen::Vector3d::Zero();
  Eigen::Vector3d accel = Eigen::Vector3d::Zero();
  Eigen::Vector3d jerk = Eigen::Vector3d::Zero();
  double yaw = 0;
  double dyaw = 0;

  void setPos(const double x, const double y, const double z) { pos = Eigen::Vector3d(x, y, z); }
  void setVel(const double x, const double y, const double z) { vel = Eigen::Vector3d(x, y, z); }
  void setAccel(const double x, const double y, const double z) { accel = Eigen::Vector3d(x, y, z); }
  void setJerk(const double x, const double y, const double z) { jerk = Eigen::Vector3d(x, y, z); }
  void setPos(const Eigen::Vector3d& d) { pos = d; }
  void setVel(const Eigen::Vector3d& d) { vel = d; }
  void setAccel(const Eigen::Vector3d& d) { accel = d; }
  void setJerk(const Eigen::Vector3d& d) { jerk = d; }
  void setYaw(const double& d) { yaw = d; }
  void setZero()
  {
    pos = vel = accel = jerk = Eigen::Vector3d::Zero();
    yaw = 0;
    dyaw = 0;
  }
  void printPos() { std::cout << "Pos= " << pos.x() << " " << pos.y() << " " << pos.z() << std::endl; }
};
